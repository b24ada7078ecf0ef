mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_reduce_gpu.py tests/test_edge_cases_gpu.py -m gpu -q 2>&1 | grep -E "Error|assert |passed|failed" | head
for w in 2 4; do CUDA_MODULE_LOADING=EAGER timeout 120 python tests/virtual_ranks_worker.py $w 2>&1 | tail -3; echo "virtual world $w rc=$?"; done
CUDA_MODULE_LOADING=EAGER timeout 120 python tests/virtual_ranks_worker.py 2 2>&1 | tail -1
timeout 120 python tests/virtual_ranks_worker.py 2 2>&1 | tail -1; echo "lazy rc=$?"
timeout 300 python tools/layout_survey.py --out gpurun_out/r02k_layout_survey.txt > /dev/null 2>&1; tail -12 gpurun_out/r02k_layout_survey.txt
