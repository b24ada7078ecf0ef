timeout 600 python -m pytest tests/test_softmax_misc_gpu.py -m gpu -q 2>&1 | grep -E "Error|assert|passed|failed" | head
timeout 300 python tools/band_sweep.py
