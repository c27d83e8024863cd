T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e"
run() { echo "== $*"; env "${@:2}" $T $1 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('   aggregate %.1f GB/s  ms %.4f  per-GPU %.1f GB/s' % (d['value'], d['ms_per_step'], d['value']/d['n_gpus']))
    elif 'rror' in l: print('   '+l[:200])
"; }
run "--index-pattern remote" A=1
run "--index-pattern local" A=1
run "--index-pattern random" A=1
run "--index-pattern remote" WG_UNROLL=8
run "--index-pattern remote" WG_UNROLL=2
run "--index-pattern remote" WG_BATCH_ROWS=32
run "--index-pattern remote" WG_BATCH_ROWS=4
run "--index-pattern random" WG_UNROLL=8
run "--index-pattern random" WG_UNROLL=8 WG_BATCH_ROWS=32
run "--index-pattern random" WG_BATCH_ROWS=4
run "--index-pattern remote" WG_CACHE_POLICY=1
run "--index-pattern remote" WG_CACHE_POLICY=2
run "--index-pattern random --dim 128 --dtype fp16 --rows-per-gpu 125000000" A=1
run "--index-pattern remote --dim 128 --dtype fp16 --rows-per-gpu 125000000" A=1
run "--index-pattern random --dim 256 --dtype fp16 --rows-per-gpu 125000000" A=1
