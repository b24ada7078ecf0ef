timeout 300 python -m pytest tests/test_reduce_gpu.py -m gpu -q 2>&1 | grep -E "Error|assert |passed|failed" | head -5
timeout 300 python tools/layout_survey.py --out gpurun_out/r02n_layout_survey.txt > /dev/null 2>&1; grep -i "arg" gpurun_out/r02n_layout_survey.txt
