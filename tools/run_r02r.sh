timeout 600 python -m pytest tests/test_softmax_misc_gpu.py -m gpu -q 2>&1 | grep -E "Error|assert |passed|failed" | head -5
timeout 200 python tools/band_sweep.py 2>&1 | tail -6
HPTB_TUNE_NO_BAND_PIPE=1 timeout 200 python tools/band_sweep.py 2>&1 | grep "axis 0"
