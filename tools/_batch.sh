mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02n_pytest_all.log 2>&1; tail -6 gpurun_out/r02n_pytest_all.log
HF6D_BENCH_WATCHDOG=500 timeout 600 python bench.py > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err; tail -c 300 gpurun_out/r02n_bench.err; head -c 200 gpurun_out/r02n_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02n_bench_ref.json 2> gpurun_out/r02n_bench_ref.err; head -c 300 gpurun_out/r02n_bench_ref.json
