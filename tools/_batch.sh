mkdir -p gpurun_out
for c in c1 c3 c4 c5; do
  HF6D_BENCH_WATCHDOG=500 timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_bench_$c.json 2> gpurun_out/r02h_bench_$c.err
  head -c 300 gpurun_out/r02h_bench_$c.json; echo; tail -c 300 gpurun_out/r02h_bench_$c.err
done
