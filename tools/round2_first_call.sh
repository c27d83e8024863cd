#!/bin/bash
# First GPU call of the next round (1 GPU, ~25 min): everything that was written after round 1's GPU budget ran out gets
# its first execution, then the headline numbers and ncu evidence are refreshed under r2 names.
#   gpurun --timeout 2400 -- 'bash tools/round2_first_call.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=${ROUND_TAG:-r2}
step() { echo "=== $*"; }

step "1. full GPU suite (incl. tests/test_zz_*: native env, reference-binary optimizer + graph-ops parity, full-size configs)"
timeout 1500 python -X faulthandler -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_$R.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$R.log | cut -c1-300
# if -x stopped early, still learn what the new files do on their own
for f in tests/test_zz_cpp_abi.py tests/test_zz_native_env.py tests/test_zz_ref_optimizer_parity_gpu.py tests/test_zz_ref_graph_ops_parity_gpu.py tests/test_zz_ref_sampling_parity_gpu.py tests/test_zz_reference_binding_full_gpu.py tests/test_zz_vec32_gather_gpu.py tests/test_zz_training_autograd_gpu.py tests/test_zz_file_io_grid_gpu.py tests/test_zz_sampling_grid_gpu.py tests/test_zz_gather_scatter_functors_gpu.py tests/test_zz_one_dim_table_gpu.py tests/test_zz_baseline_configs_gpu.py; do
  timeout 900 python -X faulthandler -m pytest $f -m gpu -q -s > gpurun_out/$(basename $f .py)_$R.log 2>&1; echo "$f rc=$?"; grep -E "bit-identical|passed|failed|Error" gpurun_out/$(basename $f .py)_$R.log | tail -6 | cut -c1-300
done

step "2. smoke + both bench arms"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference_$R.json 2> gpurun_out/bench_reference_$R.err; tail -1 gpurun_out/bench_reference_$R.json | cut -c1-400
timeout 900 python bench.py > gpurun_out/bench_ours_$R.json 2> gpurun_out/bench_ours_$R.err; tail -1 gpurun_out/bench_ours_$R.json | cut -c1-1800

echo "-- A/B: 256-bit accesses (WG_VEC32=1), C2 and the 256-byte-row shape"
WG_VEC32=1 timeout 600 python bench.py --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-330
timeout 600 python bench.py --no-e2e --no-cpu-baseline --dim 128 --dtype fp16 --rows-per-gpu 125000000 2>/dev/null | tail -1 | cut -c1-330
WG_VEC32=1 timeout 600 python bench.py --no-e2e --no-cpu-baseline --dim 128 --dtype fp16 --rows-per-gpu 125000000 2>/dev/null | tail -1 | cut -c1-330

step "3. ncu: launch list of the bench command + one full capture of the gather kernel"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_bench_c2.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:row_move_vec -s 4 -c 1 -o gpurun_out/${R}_gather_c2_full -f python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --rows-per-gpu 20000000 > gpurun_out/ncu_gather_$R.log 2>&1
ncu -i gpurun_out/${R}_gather_c2_full.ncu-rep --page raw --csv 2>/dev/null | python - <<'PY' > gpurun_out/${R}_gather_c2_full_summary.txt
import csv, sys
rows = list(csv.reader(sys.stdin))
if len(rows) >= 3:
    hdr, units, vals = rows[0], rows[1], rows[2]
    want = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
            "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct")
    for h, u, v in zip(hdr, units, vals):
        if h in want:
            print("%-60s %s %s" % (h, v, u))
PY
cat gpurun_out/${R}_gather_c2_full_summary.txt

step "4. other kernels, the reference's optimizer kernels next to ours, native env A/B on the sampler"
timeout 600 python tools/bench_ops.py 2>&1 | tail -8
# the rows DESIGN.md section 9 still lists as "not timed": append_unique, self loop, weighted sampler, gather under an SM budget
timeout 900 python tools/bench_ops.py --what unique,selfloop,weighted,budget 2>&1 | tail -16
WHOLEGRAPH_B200_LIB=oracle/_ref/libwholegraph_ref.so timeout 600 python tools/bench_ops.py --what refadam 2>&1 | tail -3
# the reference's WHOLE wholememory_embedding_gather_gradient_apply (its embedding layer is built into oracle/_ref): same tool, other library
WHOLEGRAPH_B200_LIB=oracle/_ref/libwholegraph_ref.so timeout 600 python tools/bench_ops.py --what adam 2>&1 | tail -3
WG_TORCH_NATIVE_ENV=1 timeout 600 python tools/bench_ops.py --what sample 2>&1 | tail -4
# the reference's own sampler kernels (restated PCG stand-in) on the same workload: the C5 "reference" column
WHOLEGRAPH_B200_LIB=oracle/_ref/libwholegraph_ref.so timeout 600 python tools/bench_ops.py --what sample 2>&1 | tail -4
timeout 600 python tools/bench_sample_multi.py 2>&1 | tail -1
WG_TORCH_NATIVE_ENV=1 timeout 600 python tools/bench_sample_multi.py 2>&1 | tail -1
ls -la gpurun_out | tail -12

step "4b. C++ bench with the reference's command line, same binary on both libraries (config C2)"
B=wholegraph_b200/lib/gather_scatter_bench
[ -x $B ] || g++ -std=c++17 -O2 -Iinclude -I/usr/local/cuda/include tools/gather_scatter_bench.cpp -o $B -L/usr/local/cuda/lib64 -lcudart -ldl
timeout 600 $B -t 1 -l 1 -e 102400000000 -g 1073741824 -d 256 -c 20 -n 1 2>&1 | tail -3
timeout 600 $B -t 1 -l 1 -e 102400000000 -g 1073741824 -d 256 -c 20 -n 1 --lib oracle/_ref/libwholegraph_ref.so 2>&1 | tail -3

step "5. sanitizers on the smoke pass (every kernel of the hot path once; SURVEY section 5)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck_$R.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|smoke OK" gpurun_out/sanitizer_memcheck_$R.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck_$R.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|smoke OK" gpurun_out/sanitizer_racecheck_$R.log | tail -3

step "6. golden vectors from the reference binary for the optimizer / sampler / graph-op kernels (copy into tests/golden/ afterwards)"
bash tools/make_golden.sh 2>&1 | tail -6
