B="python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --per-step-events"
run() { echo "== $*"; env "$@" $B 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('   value %.1f GB/s  ms %.4f  frac %.4f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))
    elif 'per-step' in l or 'rror' in l: print('   '+l[:200])
"; }
run A=1
run WG_GRID_MODE=1
run WG_GRID_MODE=1 WG_BATCH_ROWS=4
run WG_GRID_MODE=1 WG_BATCH_ROWS=32
run WG_GRID_MODE=1 WG_THREADS=128
run WG_GRID_MODE=1 WG_THREADS=512 WG_BATCH_ROWS=8
run WG_THREADS=512 WG_BLOCKS_PER_SM=2
run WG_THREADS=1024 WG_BLOCKS_PER_SM=1
run WG_THREADS=128 WG_BLOCKS_PER_SM=10
run WG_GRID_MODE=1 WG_UNROLL=2
