#!/bin/bash
# Round 2, GPU call 1 (1 GPU): whole GPU suite in ONE process (no -x: every failure is listed), the remaining golden
# fixtures from the reference binary, and kernel-lab sweeps for 256 B / 512 B / 1 KiB rows.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=r2
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clocks_call1.csv &
SMI=$!
echo "=== 1. pytest -m gpu, one process"
timeout 2400 python -X faulthandler -m pytest tests -m gpu -q -p no:cacheprovider --durations=15 > gpurun_out/pytest_gpu_$R.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/pytest_gpu_$R.log | cut -c1-400
echo "=== 2. golden fixtures from the reference binary"
bash tools/make_golden.sh 2>&1 | tail -8
echo "=== 3. kernel lab, local HBM"
L=wholegraph_b200/lib/rowmove_lab
for rb in 256 512 1024; do
  timeout 600 $L --row-bytes $rb --rows $((20000000*1024/rb/4)) --n 1048576 --mode local --set small > gpurun_out/lab_local_rb${rb}_random.txt 2>&1; cat gpurun_out/lab_local_rb${rb}_random.txt
done
timeout 300 $L --row-bytes 256 --rows 20000000 --n 1048576 --mode local --set default --pattern seq
timeout 300 $L --row-bytes 1024 --rows 5000000 --n 1048576 --mode local --set default --pattern seq
timeout 300 $L --row-bytes 256 --rows 20000000 --n 4194304 --mode local --set default
timeout 300 $L --row-bytes 256 --rows 20000000 --n 1048576 --mode local --set default --op scatter
echo "=== 4. smoke + bench arms (old bench.py contract)"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference_$R.json 2> gpurun_out/bench_reference_$R.err; tail -1 gpurun_out/bench_reference_$R.json | cut -c1-300
timeout 900 python bench.py > gpurun_out/bench_ours_$R.json 2> gpurun_out/bench_ours_$R.err; tail -1 gpurun_out/bench_ours_$R.json | cut -c1-600
kill $SMI
