mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_tma_tile_gpu.py -m gpu -q -x -k "copy" > gpurun_out/r02t_racecheck_tma.log 2>&1; echo "racecheck tma rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r02t_racecheck_tma.log | tail -4
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_softmax_misc_gpu.py -m gpu -q -x -k "band" > gpurun_out/r02t_racecheck_band.log 2>&1; echo "racecheck band rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r02t_racecheck_band.log | tail -4
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_softmax_misc_gpu.py tests/test_tma_tile_gpu.py -m gpu -q -x -k "band or copy" > gpurun_out/r02t_synccheck.log 2>&1; echo "synccheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02t_synccheck.log | tail -3
python -m pytest tests -m gpu -x -q 2>&1 | tail -1
