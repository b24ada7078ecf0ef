"""Worker of tests/test_sharded_gpu.py::test_exchange_protocol_with_virtual_ranks_on_one_gpu: `world` virtual ranks on
cuda:0 (hptb_comm_init_local_group), one stream per rank, all ranks' kernels of a call in flight together."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2

import numpy as np
import hpt_b200 as hb
from util import ENUM, O, assert_reduce_bar, rand, to_numpy, to_torch

ctx = hb.context(0)
comms = hb.Comm.local_group(ctx, world)
streams = [torch.cuda.Stream() for _ in range(world)]
rng = np.random.default_rng(61)

def run(x, d, sax, op, axes):
    n = x.shape[sax]
    shards = []
    for r in range(world):
        off, ln = hb.shard_bounds(n, world, r)
        block = np.ascontiguousarray(np.take(x, range(off, off + ln), axis=sax))
        local = hb.Tensor.to_cuda(to_torch(block, d), 0, stream=streams[r].cuda_stream)
        shards.append(hb.ShardedTensor(local, comms[r], sax, n, off))
    outs = [shards[r]._reduce(op, axes if not op.startswith("arg") else axes[0], stream=streams[r].cuda_stream) for r in range(world)]
    torch.cuda.synchronize()
    want, od, exact = O.reduce(op, x, d, axes)
    got = [to_numpy(o.to_cpu(), od) for o in outs]
    for r in range(world):
        assert_reduce_bar(op, x, d, axes, got[r], want, od, exact, f"{op} {d} {x.shape} axes={axes} virtual rank {r}/{world}")
        assert got[r].tobytes() == got[0].tobytes(), f"{op} {d}: rank {r} differs from rank 0"

for d in ("f32", "bf16", "f16", "i32", "i64", "u8", "bool", "f64"):
    for shape, sax in (((37, 24), 0), ((6, 40), 1)):
        x = rand(rng, shape, d)
        for op in ("sum", "mean", "max", "min", "logsumexp", "sum_square", "reducel2", "all", "any", "argmax", "argmin"):
            if op == "logsumexp" and d in O.INTS:
                continue
            for axes in ([[sax]] if op.startswith("arg") else [[sax], list(range(len(shape)))]):
                run(x, d, sax, op, axes)
# config 5's shape in small: the fused epilogues (one CTA for sum(), one per column tile for sum(axis 0))
x = rand(rng, (2048, 16384), "f32")
for op, axes in (("sum", [0, 1]), ("mean", [0, 1]), ("sum", [0]), ("argmax", [0]), ("logsumexp", [0])):
    run(x, "f32", 0, op, axes)
xh = rand(rng, (512, 16384), "bf16")
for op, axes in (("sum", [0]), ("mean", [0, 1])):
    run(xh, "bf16", 0, op, axes)
# ties, NaNs and an all-NaN column across the shard boundary
x = rng.integers(0, 3, size=(16 * world, 300)).astype(np.float32)
x[:, 5] = np.nan
x[: 10 * world, 7] = np.nan
x[:, 9] = -np.inf
run(x, "f32", 0, "argmax", [0])
run(x, "f32", 0, "argmin", [0])
for c in comms:
    c.destroy()

print(f"virtual ranks ok (world {world})")
