mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02b_topo.txt 2>&1
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q > gpurun_out/r02b_pytest_sharded_2gpu.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r02b_pytest_sharded_2gpu.log
timeout 300 python tools/sweep_shard.py > gpurun_out/r02b_sweep_shard.txt 2>&1; cat gpurun_out/r02b_sweep_shard.txt
