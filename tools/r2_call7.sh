#!/bin/bash
# Round 2, GPU call 7 (1 GPU): hot-row kernel (list-based dispatch) tests + timing; multi-hop sampling at FULL graph size on one GPU,
# both libraries, with per-kernel launch lists (why is the small-batch step slower than the reference's at size?).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
REF=oracle/_ref/libwholegraph_ref.so
timeout 1200 python -m pytest tests/test_sparse_optimizer_gpu.py tests/test_zz_ref_optimizer_parity_gpu.py "tests/test_multi_rank_gpu.py::test_ranks_sharing_one_gpu_mapped_memory" -m gpu -q -p no:cacheprovider -k "not sampling and not file_io and not gather_scatter" > gpurun_out/pytest_call7.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_call7.log | cut -c1-500
timeout 600 python tools/bench_ops.py --what adam 2>&1 | tail -3
S="tools/bench_sample_multi.py --nodes 111059956 --edges 1000000000"
timeout 600 python $S 2>&1 | grep -E '^\{' | cut -c1-500
WHOLEGRAPH_B200_LIB=$REF timeout 600 python $S 2>&1 | grep -E '^\{' | cut -c1-500
WG_TORCH_NATIVE_ENV=1 timeout 600 python $S 2>&1 | grep -E '^\{' | cut -c1-500
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_sample_full_ours.csv python $S --steps 3 --warmup 3 > /dev/null 2>&1
WHOLEGRAPH_B200_LIB=$REF timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_sample_full_ref.csv python $S --steps 3 --warmup 3 > /dev/null 2>&1
for w in ours ref; do echo "== launch list tail, $w"; python - gpurun_out/r2_launches_sample_full_$w.csv <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ii = {h: i for i, h in enumerate(hdr)}
items = [(r[ii["Kernel Name"]][:90], float(r[ii["Metric Value"]])) for r in rows[1:]]
# the last step = everything after the last-but-one occurrence pattern; simply print the final 40 launches
for name, ns in items[-44:]:
    print("%9.1f us  %s" % (ns / 1e3, name))
PY
done
