# tools/run_scaling_8gpu.sh — one 8-GPU gpurun call: sharded parity at world 8, then bench.py at N = 8, 4, 2, 1 (strong scaling of config 5)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29601 tests/sharded_worker.py > gpurun_out/r02s_sharded_worker_8gpu.log 2>&1; echo "sharded8 rc=$?"; grep -c "sharded ok" gpurun_out/r02s_sharded_worker_8gpu.log; tail -2 gpurun_out/r02s_sharded_worker_8gpu.log | cut -c1-300
for N in 8 4 2; do
  timeout 400 $TR --nproc-per-node $N --master-port 2961$N bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r02s_bench_${N}gpu.json 2> gpurun_out/r02s_bench_${N}gpu.err; echo "bench$N rc=$?"
done
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/r02s_bench_1gpu.json 2> gpurun_out/r02s_bench_1gpu.err; echo "bench1 rc=$?"
python - <<'PY'
import json
base=None
for n in (1,2,4,8):
    try:
        d=json.load(open(f'gpurun_out/r02s_bench_{n}gpu.json'))
    except Exception as ex:
        print(n, 'failed', ex); continue
    if n==1: base=d['value']
    print(n, 'value', d['value'], 'ms', d['ms_per_step'], 'speedup', round(d['value']/base,3) if base else None, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], [(k['op'],k['us']) for k in d['kernels']], d['parity']['ok_all_ranks'], d['gpu_launches'])
PY
