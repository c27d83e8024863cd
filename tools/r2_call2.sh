#!/bin/bash
# Round 2, GPU call 2 (1 GPU): full GPU suite again after the fixes, optimizer golden fixture, both bench arms with the
# three BASELINE shapes, ncu launch list + full captures on the bench-sized tables, C1 / overlap / reference columns.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=r2
REF=oracle/_ref/libwholegraph_ref.so
echo "=== 1. pytest -m gpu, one process"
timeout 2400 python -X faulthandler -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > gpurun_out/pytest_gpu_${R}b.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_gpu_${R}b.log | cut -c1-600
grep -n "differs from the oracle" -A12 gpurun_out/pytest_gpu_${R}b.log | head -30
echo "=== 2. optimizer golden fixture from the reference binary"
WG_GOLDEN_SMALL=1 WHOLEGRAPH_B200_LIB=$REF timeout 900 python tests/ref_optimizer_worker.py gpurun_out/reference_optimizer_golden.npz 2>&1 | tail -3
ls -la gpurun_out/reference_optimizer_golden.npz
echo "=== 3. bench arms"
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference_$R.json 2> gpurun_out/bench_reference_$R.err; tail -1 gpurun_out/bench_reference_$R.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('REF', d['value'], d['ms_per_step'], d['e2e']['value'], [(s['key'], s['value'], s['ms_per_step'], s['frac']) for s in d['shapes']])"
timeout 900 python bench.py > gpurun_out/bench_ours_$R.json 2> gpurun_out/bench_ours_$R.err; tail -1 gpurun_out/bench_ours_$R.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('OURS', d['value'], d['ms_per_step'], d['e2e'], [(s['key'], s['value'], s['ms_per_step'], s['frac']) for s in d['shapes']], d['clocks'], d['cpu_baseline']['value'])"
tail -3 gpurun_out/bench_ours_$R.err
echo "=== 4. ncu: launch list of the bench command, full captures on the bench-sized tables"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_bench_c2.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --shapes c2 > /dev/null 2>&1
grep -c row_move gpurun_out/${R}_launches_bench_c2.csv
for k in c2 c3 ns; do
  CMD="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --shapes $k"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:row_move_vec -s 4 -c 1 -o gpurun_out/${R}_gather_${k}_full -f $CMD > gpurun_out/ncu_gather_${k}_$R.log 2>&1
  ncu -i gpurun_out/${R}_gather_${k}_full.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py "ncu --set full --clock-control none --import-source on -k regex:row_move_vec -s 4 -c 1  $CMD   (round 2, B200, bench-sized table; per-launch values)" > gpurun_out/${R}_gather_${k}_full_summary.txt
  grep -E "gpu__time_duration|dram__bytes|launch__grid|registers_per|warps_active" gpurun_out/${R}_gather_${k}_full_summary.txt
done
echo "=== 5. C1 (HOST table) on both libraries, gather||sample overlap, SM budgets, reference columns"
timeout 600 python tools/bench_ops.py --what c1 2>&1 | tail -2
WHOLEGRAPH_B200_LIB=$REF timeout 600 python tools/bench_ops.py --what c1 2>&1 | tail -2
timeout 900 python tools/bench_ops.py --what overlap,budget 2>&1 | tail -14
timeout 600 python tools/bench_ops.py --what scatter,adam,sample 2>&1 | tail -8
WHOLEGRAPH_B200_LIB=$REF timeout 600 python tools/bench_ops.py --what adam,sample 2>&1 | tail -8
WHOLEGRAPH_B200_LIB=$REF timeout 600 python tools/bench_ops.py --what refadam 2>&1 | tail -3
timeout 900 python tools/bench_ops.py --what unique,selfloop,weighted 2>&1 | tail -10
