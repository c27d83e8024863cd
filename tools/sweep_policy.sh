B="python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --per-step-events"
run() { echo "== $*"; env "$@" $B 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('   value %.1f GB/s  ms %.4f  frac %.4f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))
    elif 'per-step' in l or 'rror' in l: print('   '+l[:200])
"; }
# store variants (high nibble) with default load
for st in 0 1 2 3 4; do run WG_CACHE_POLICY=$((st*16)); done
# load variants with default store
for ld in 1 2 3 4; do run WG_CACHE_POLICY=$ld; done
run WG_CACHE_POLICY=17
run WG_CACHE_POLICY=18 WG_UNROLL=2
python - <<'PY'
import torch
a = torch.empty(1<<28, dtype=torch.float32, device='cuda'); b = torch.empty_like(a)
big = torch.empty(1<<30, dtype=torch.float32, device='cuda'); big2 = torch.empty_like(big)
for name, x, y in (('1 GiB copy (2 GiB traffic)', a, b), ('4 GiB copy (8 GiB traffic)', big, big2)):
    for _ in range(3): y.copy_(x)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); y.copy_(x); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print('%s: best %.4f ms -> %.1f GB/s' % (name, best, 2 * x.numel() * 4 / best / 1e6))
PY
