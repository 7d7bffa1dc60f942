mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cli.py tests/test_refine.py -m gpu -q > gpurun_out/r02d_pytest_cli.log 2>&1; tail -8 gpurun_out/r02d_pytest_cli.log
