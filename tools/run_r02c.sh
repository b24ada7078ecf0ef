mkdir -p gpurun_out
timeout 600 python tools/sweep_shard.py > gpurun_out/r02c_sweep_shard.txt 2>&1; cat gpurun_out/r02c_sweep_shard.txt
