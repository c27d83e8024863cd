#!/bin/bash
# round-end style verification on one GPU: full gpu test suite, smoke, both bench arms, ncu evidence of the optimizer kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -X faulthandler -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_final.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference > gpurun_out/bench_reference_r1.json 2> gpurun_out/bench_reference_r1.err; tail -1 gpurun_out/bench_reference_r1.json | cut -c1-400
python bench.py > gpurun_out/bench_ours_r1.json 2> gpurun_out/bench_ours_r1.err; tail -1 gpurun_out/bench_ours_r1.json | cut -c1-1800
ncu --set full --clock-control none --import-source on -k regex:fused_merge_update -s 3 -c 1 -o gpurun_out/prof_adam_r1 -f python tools/bench_ops.py --what adam > gpurun_out/ncu_adam.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_adam_r1.csv python tools/bench_ops.py --what adam > /dev/null 2>&1
python tools/bench_ops.py 2>&1 | tail -8
ls -la gpurun_out | tail -5
