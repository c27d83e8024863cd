N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
python -m pytest tests/test_multi_rank_gpu.py -m gpu -q -x -k "one_rank_per_gpu" > gpurun_out/pytest_${N}gpu.log 2>&1; tail -4 gpurun_out/pytest_${N}gpu.log | cut -c1-300
run() { echo "== N=$1 ${@:2}"; python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --steps 20 --warmup 5 "${@:2}" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('   aggregate %.1f GB/s  ms %.4f  per-GPU %.1f GB/s  e2e %s' % (d['value'], d['ms_per_step'], d['value']/d['n_gpus'], d.get('e2e',{}).get('value')))
        open('gpurun_out/bench_lines.jsonl','a').write(l+'\n')
    elif 'rror' in l: print('   '+l[:300])
"; }
for n in 2 4 8; do [ $n -le $N ] && run $n; done
[ 8 -le $N ] && WG_BULK=0 run 8 --no-e2e
[ 8 -le $N ] && WG_BULK_SLOT_KB=4 run 8 --no-e2e
[ 8 -le $N ] && run 8 --no-e2e --dim 256 --dtype fp16 --rows-per-gpu 125000000
[ 8 -le $N ] && run 8 --no-e2e --dim 128 --dtype fp16 --rows-per-gpu 125000000
[ 8 -le $N ] && run 8 --no-e2e --impl reference
[ 4 -le $N ] && run 4 --no-e2e --impl reference
run 2 --no-e2e --impl reference
