timeout 1200 python -m pytest tests/test_gpu_configs.py tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -15
