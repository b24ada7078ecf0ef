"""One rank of tests/test_sharded_gpu.py (launched with torch.distributed.run, one process per GPU):
hptb_reduce_sharded / ShardedTensor against the oracle's GLOBAL reduction, every op, NCCL exchange included."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import hpt_b200 as hb
    from oracle import hpt_oracle as O
    from util import rand, to_numpy, to_torch

    hb.set_stream(torch.cuda.current_stream().cuda_stream)
    comm = hb.Comm.from_torch_distributed(hb.context(local))
    p2p = hb.lib.hptb_comm_uses_peer_memory(comm.handle)
    rng = np.random.default_rng(11)  # identical data on every rank
    checked = 0
    for d in ("f32", "f64", "bf16", "f16", "i32", "i64", "u8", "bool"):
        for shape, sax in (((37, 24), 0), ((16, 5, 12), 0), ((6, 40), 1)):
            x = rand(rng, shape, d)
            if d in ("f32", "f64"):
                x.flat[::7] = x.flat[3]  # ties
            X = hb.ShardedTensor.scatter_from_host(to_torch(x, d), comm, sax, device=local)
            for op in ("sum", "mean", "max", "min", "prod", "logsumexp", "sum_square", "reducel1", "reducel2", "reducel3", "nansum",
                       "all", "any", "argmax", "argmin"):
                if op == "prod" and d in ("f16", "bf16"):
                    continue  # products of hundreds of half-precision values over/underflow: nothing to compare
                axes_list = [[sax]] if op.startswith("arg") else [[sax], list(range(len(shape)))]
                for axes in axes_list:
                    want, od, exact = O.reduce(op, x, d, axes)
                    got_t = getattr(X, op)(axes if not op.startswith("arg") else axes[0])
                    assert isinstance(got_t, hb.Tensor)
                    got = to_numpy(got_t.to_cpu(), od)
                    if exact or od in O.INTS or od in ("bool", "i64"):
                        np.testing.assert_array_equal(got, want, err_msg=f"{op} {d} {shape} axes={axes} rank={rank}")
                    else:
                        # partial sums are exchanged in the output dtype; bound relative to Σ|x| as in test_reduce_gpu
                        tol = {"f64": 1e-12, "f32": 2e-5, "f16": 2e-2, "bf16": 6e-2}[od]
                        np.testing.assert_allclose(got.astype(np.float64), np.asarray(want, dtype=np.float64), rtol=tol, atol=tol,
                                                   err_msg=f"{op} {d} {shape} axes={axes} rank={rank}")
                    checked += 1
            # a reduction that keeps the shard axis stays sharded and local
            if d == "f32":
                keep = X.sum([len(shape) - 1 if sax == 0 else 0])
                assert isinstance(keep, hb.ShardedTensor)
                want, _, _ = O.reduce("sum", x, d, [len(shape) - 1 if sax == 0 else 0])
                off, ln = hb.shard_bounds(shape[sax], world, rank)
                got = keep.local.to_cpu().numpy()
                np.testing.assert_allclose(got, np.take(want, range(off, off + ln), axis=keep.shard_axis), rtol=1e-5, atol=1e-5)
    dist.barrier()
    comm.destroy()
    dist.destroy_process_group()
    print(f"rank {rank}: sharded ok ({checked} cases, peer memory {p2p})")


if __name__ == "__main__":
    main()
