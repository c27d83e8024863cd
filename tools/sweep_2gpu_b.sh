T1="python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline"
T2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e"
run() { echo "== $*"; env "${@:2}" $1 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('   aggregate %.1f GB/s  ms %.4f  per-GPU %.1f GB/s frac %.4f' % (d['value'], d['ms_per_step'], d['value']/d['n_gpus'], d['roofline']['frac']))
    elif 'rror' in l: print('   '+l[:300])
"; }
run "$T1" A=1
run "$T1 --dim 128 --dtype fp16 --rows-per-gpu 125000000" A=1
run "$T1 --dim 128 --dtype fp16 --rows-per-gpu 125000000" WG_BATCH_ROWS=16
run "$T1 --dim 128 --dtype fp16 --rows-per-gpu 125000000" WG_BATCH_ROWS=8
run "$T1 --dim 64 --dtype fp32 --rows-per-gpu 100000000" A=1
run "$T1 --dim 1024 --dtype fp32 --rows-per-gpu 25000000" A=1
run "$T2" A=1
run "$T2" WG_BULK=0
run "$T2" WG_BULK=0 WG_BATCH_ROWS=4
run "$T2 --index-pattern remote" WG_BULK=0
python -m pytest tests/test_gather_scatter_gpu.py tests/test_ref_parity_gpu.py -m gpu -q -x 2>&1 | tail -2
python tools/bench_ops.py --what scatter 2>&1 | tail -1
