set -x
mkdir -p gpurun_out
nproc; free -g | head -2; nvidia-smi topo -m 2>&1 | head -12
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest_2gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02a_pytest_2gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r02a_bench_1gpu.json 2> gpurun_out/r02a_bench_1gpu.err; echo "bench1 rc=$?"
tail -3 gpurun_out/r02a_bench_1gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02a_bench_2gpu.json 2> gpurun_out/r02a_bench_2gpu.err; echo "bench2 rc=$?"
tail -3 gpurun_out/r02a_bench_2gpu.err
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02a_bench_reference.json 2>/dev/null; echo "ref rc=$?"
timeout 200 python tools/shard_probe.py > gpurun_out/r02a_shard_probe.txt 2>&1; cat gpurun_out/r02a_shard_probe.txt
cat gpurun_out/r02a_bench_1gpu.json | cut -c1-1500
cat gpurun_out/r02a_bench_2gpu.json | cut -c1-1500
