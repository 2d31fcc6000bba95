mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_env.py tests/test_gpu_cli.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -6
