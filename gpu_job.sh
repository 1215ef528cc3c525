timeout 900 python -m pytest tests/test_gpu_edges.py tests/test_gpu_api.py -m gpu -q 2>&1 | tail -5
python scripts/probe_latency.py 2>&1 | head -3
