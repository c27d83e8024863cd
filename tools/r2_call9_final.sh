#!/bin/bash
# Round 2, GPU call 9 (1 GPU): what the driver does at round end -- the GPU suite in ONE process with -x, smoke(), both bench arms --
# plus refreshed ncu evidence and a sanitizer pass.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=r2
echo "=== 1. python -m pytest tests/ -x -q -m gpu"
timeout 2400 python -X faulthandler -m pytest tests/ -x -q -m gpu -p no:cacheprovider --durations=5 > gpurun_out/pytest_gpu_${R}_final.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_gpu_${R}_final.log | cut -c1-300
echo "=== 2. smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== 3. bench arms (driver flags)"
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_reference_$R.json 2> gpurun_out/bench_reference_$R.err; tail -1 gpurun_out/bench_reference_$R.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('REF', d['value'], d['ms_per_step'], d['e2e']['value'], [(s['key'], s['value'], s['ms_per_step'], s['frac']) for s in d['shapes']])"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ours_$R.json 2> gpurun_out/bench_ours_$R.err; tail -1 gpurun_out/bench_ours_$R.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('OURS', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('device_resident_output',{}).get('value'), [(s['key'], s['value'], s['ms_per_step'], s['frac']) for s in d['shapes']], d['clocks'], d['cpu_baseline']['value'], d['roofline']['traffic'])"
echo "=== 4. ncu refresh"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_bench_c2.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --shapes c2 > /dev/null 2>&1
for k in c2 c3 ns; do
  CMD="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --shapes $k"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:row_move_vec -s 4 -c 1 -o gpurun_out/${R}_gather_${k}_full -f $CMD > gpurun_out/ncu_gather_${k}_$R.log 2>&1
  ncu -i gpurun_out/${R}_gather_${k}_full.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py "ncu --set full --clock-control none --import-source on -k regex:row_move_vec -s 4 -c 1  $CMD   (round 2, B200, bench-sized table; per-launch values)" > gpurun_out/${R}_gather_${k}_full_summary.txt
  grep -E "gpu__time_duration|dram__bytes" gpurun_out/${R}_gather_${k}_full_summary.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_merge_update -s 2 -c 1 -o gpurun_out/${R}_fused_adam_full -f python tools/bench_ops.py --what adam > /dev/null 2>&1
ncu -i gpurun_out/${R}_fused_adam_full.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py "ncu --set full -k regex:fused_merge_update -s 2 -c 1 python tools/bench_ops.py --what adam  (LazyAdam 5M x 512 fp32, 262,144 uniform gradient rows)" > gpurun_out/${R}_fused_adam_full_summary.txt
grep -E "gpu__time_duration|dram__bytes|registers" gpurun_out/${R}_fused_adam_full_summary.txt
echo "=== 5. compute-sanitizer memcheck on the smoke pass"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck_$R.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|smoke OK" gpurun_out/sanitizer_memcheck_$R.log | tail -3
