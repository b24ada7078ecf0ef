mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_reduce_gpu.py tests/test_full_size_gpu.py tests/test_edge_cases_gpu.py tests/test_reference_kernels_gpu.py -m gpu -q 2>&1 | grep -E "Error|assert |passed|failed" | head -5
timeout 300 python tools/layout_survey.py --out gpurun_out/r02o_layout_survey.txt > /dev/null 2>&1; tail -12 gpurun_out/r02o_layout_survey.txt
# memcheck over the new kernels (TMA tiles, softmax bands, virtual-rank exchange)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_tma_tile_gpu.py -m gpu -q -x -k "copy or unary" > gpurun_out/r02o_memcheck_tma.log 2>&1; echo "memcheck tma rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02o_memcheck_tma.log | tail -3
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_softmax_misc_gpu.py -m gpu -q -x -k "band" > gpurun_out/r02o_memcheck_band.log 2>&1; echo "memcheck band rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02o_memcheck_band.log | tail -3
CUDA_MODULE_LOADING=EAGER timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/virtual_ranks_worker.py 2 > gpurun_out/r02o_memcheck_xchg.log 2>&1; echo "memcheck xchg rc=$?"; grep -E "ERROR SUMMARY|virtual ranks ok" gpurun_out/r02o_memcheck_xchg.log | tail -3
