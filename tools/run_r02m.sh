timeout 600 python -m pytest tests/test_reduce_gpu.py tests/test_riders_gpu.py tests/test_edge_cases_gpu.py -m gpu -q 2>&1 | grep -E "Error|assert |passed|failed" | head
timeout 300 python tools/layout_survey.py --out gpurun_out/r02m_layout_survey.txt > /dev/null 2>&1; grep -i "arg\|mean_var\|sum(ax" gpurun_out/r02m_layout_survey.txt
