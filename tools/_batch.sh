mkdir -p gpurun_out
for c in c2 c3; do
HF6D_BENCH_WATCHDOG=500 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 3 --warmup 3 --no-refine --config $c > gpurun_out/r02k_bench_4gpu_$c.json 2> gpurun_out/r02k_bench_4gpu_$c.err
done
python - <<'P'
import json
for c in ['c2','c3']:
  for line in open(f'gpurun_out/r02k_bench_4gpu_{c}.json'):
    if line.startswith('{'):
        d=json.loads(line); print(c, d['value'], d['e2e']['value'])
        for m,v in d['sharded']['modes'].items(): print(m, v.get('frames_per_s'), v.get('bit_identical'))
P
