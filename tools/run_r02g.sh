mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_softmax_misc_gpu.py tests/test_edge_cases_gpu.py -m gpu -q 2>&1 | grep -E "Error|assert|passed|failed" | head -20
python tools/dbg_softmax.py 2>&1 | tail -8
timeout 600 python tools/layout_survey.py --out gpurun_out/r02g_layout_survey.txt > /dev/null 2>&1; grep -i "softmax\|layernorm" gpurun_out/r02g_layout_survey.txt
