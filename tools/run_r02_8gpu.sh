# one 8-GPU box: sharded parity at world 8, the bench at N = 8, 4, 2, 1, the reference arm under torchrun, the PCIe probe
mkdir -p gpurun_out
nproc; free -g | head -2
nvidia-smi topo -m > gpurun_out/r02_8gpu_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29601 tests/sharded_worker.py > gpurun_out/r02_sharded_worker_8gpu.log 2>&1; echo "sharded8 rc=$?"; grep "sharded ok" gpurun_out/r02_sharded_worker_8gpu.log | head -3; tail -3 gpurun_out/r02_sharded_worker_8gpu.log | cut -c1-300
HPTB_NO_P2P=1 timeout 900 $TR --nproc-per-node 8 --master-port 29602 tests/sharded_worker.py > gpurun_out/r02_sharded_worker_8gpu_nccl.log 2>&1; echo "sharded8 nccl rc=$?"; grep -c "sharded ok" gpurun_out/r02_sharded_worker_8gpu_nccl.log
for N in 8 4 2; do
  timeout 400 $TR --nproc-per-node $N --master-port 2961$N bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "bench$N rc=$?"
done
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench1 rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29621 bench.py --impl reference --gpus 8 --steps 10 --warmup 2 > gpurun_out/r02_bench_reference_8.json 2>/dev/null; echo "ref8 rc=$?"; cut -c1-300 gpurun_out/r02_bench_reference_8.json
timeout 300 $TR --nproc-per-node 8 --master-port 29622 tools/pcie_probe.py > gpurun_out/r02_pcie_probe_8gpu.txt 2>&1; cat gpurun_out/r02_pcie_probe_8gpu.txt | grep "^rank"
python - <<'PY'
import json
base=None
for n in (1,2,4,8):
    try:
        d=json.load(open(f'gpurun_out/r02_bench_{n}gpu.json'))
    except Exception as ex:
        print(n, 'failed', ex); continue
    if n==1: base=d['value']
    print(n, 'value', d['value'], 'ms', d['ms_per_step'], 'speedup', round(d['value']/base,3) if base else None, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('pcie_h2d_gbs_per_gpu'), [(k['op'],k['us']) for k in d['kernels']], d['parity']['ok_all_ranks'], d['collective'][:40], d['gpu_launches'])
PY
