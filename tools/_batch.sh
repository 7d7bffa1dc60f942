mkdir -p gpurun_out
HF6D_BENCH_WATCHDOG=500 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02o_bench_8gpu.json 2> gpurun_out/r02o_bench_8gpu.err
python - <<'P'
import json
for line in open('gpurun_out/r02o_bench_8gpu.json'):
    if line.startswith('{'):
        d=json.loads(line); print(d['value'], d['e2e']['value'])
        for m,v in d.get('sharded',{}).get('modes',{}).items(): print(m, v.get('frames_per_s'), v.get('bit_identical'), v.get('unavailable'))
P
tail -c 400 gpurun_out/r02o_bench_8gpu.err
