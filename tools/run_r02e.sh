mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tma_tile_gpu.py tests/test_binary_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/tma_compare.py > gpurun_out/r02e_tma_compare.txt 2>&1; cat gpurun_out/r02e_tma_compare.txt
HPTB_LIB_VARIANT=sub4 timeout 300 python tools/tma_compare.py > gpurun_out/r02e_tma_compare_sub4.txt 2>&1; cat gpurun_out/r02e_tma_compare_sub4.txt
