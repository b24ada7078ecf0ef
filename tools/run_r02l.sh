mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_reduce_gpu.py tests/test_sharded_gpu.py tests/test_full_size_gpu.py -m gpu -q 2>&1 | grep -E "Error|assert |passed|failed" | head
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 50 --warmup 5 --no-rows 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], [(k['op'],k['us']) for k in d['kernels']], d['parity']['ok_all_ranks'])"
timeout 300 python bench.py --steps 50 --warmup 5 --no-rows 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], [(k['op'],k['us']) for k in d['kernels']], d['parity']['ok_all_ranks'])"
