"""One rank of tests/test_sharded_gpu.py (launched with torch.distributed.run, one process per GPU):
hptb_reduce_sharded / ShardedTensor against the oracle's GLOBAL reduction at the single-GPU bars
(tests/util.py assert_reduce_bar: integers and indices bit-exact, sums within 1e-6·log2 n of an f64 accumulation,
f16/bf16 outputs within 1 output ulp) — every op and dtype on small uneven shards, BASELINE config 5's shape
([rows,16384] f32 sharded over rows: sum(), mean(), sum(axis 0)), rows of 2^20 elements and more, outputs too many
for the mailboxes (NCCL all-gather path), strided outputs, ties and NaNs for the arg ops, run-to-run determinism and
bit-identical results on every rank."""
import os
import sys
from ctypes import byref, c_int32

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
    from hpt_b200 import _ffi
    from oracle import hpt_oracle as O
    from util import ENUM, assert_reduce_bar, rand, to_numpy, to_torch

    hb.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx = hb.context(local)
    comm = hb.Comm.from_torch_distributed(ctx)
    p2p = hb.lib.hptb_comm_uses_peer_memory(comm.handle)
    rng = np.random.default_rng(11)  # identical data on every rank
    checked = 0
    failures = []  # every case runs; all failures are reported together (a GPU box is too expensive for one at a time)

    def same_on_every_rank(got, what):
        """the rank-ordered combine makes the result bit-identical everywhere"""
        t = torch.from_numpy(np.ascontiguousarray(got).view(np.uint8).copy()).cuda()
        ref = t.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(t, ref), f"{what}: rank {rank} differs from rank 0"

    def check(x, d, sax, op, axes, what=""):
        nonlocal checked
        X = hb.ShardedTensor.scatter_from_host(to_torch(x, d), comm, sax, device=local)
        want, od, exact = O.reduce(op, x, d, axes)
        got_t = getattr(X, op)(axes if not op.startswith("arg") else axes[0])
        assert isinstance(got_t, hb.Tensor) and got_t.dtype == ENUM[od]
        got = to_numpy(got_t.to_cpu(), od)
        try:
            assert_reduce_bar(op, x, d, axes, got, want, od, exact, f"{what}{op} {d} {x.shape} axes={axes} shard_axis={sax} rank={rank}")
        except AssertionError as ex:
            failures.append(str(ex))
        same_on_every_rank(got, f"{op} {d} {x.shape}")
        checked += 1
        return X, got

    # ---- every op and dtype, small uneven shards ---------------------------------------------------------------------
    for d in ("f32", "f64", "bf16", "f16", "i8", "i16", "i32", "i64", "u8", "u32", "bool"):
        for shape, sax in (((37, 24), 0), ((16, 5, 12), 0), ((6, 40), 1)):
            x = rand(rng, shape, d)
            if d in ("f32", "f64"):
                x.flat[::7] = x.flat[3]  # ties
            # products: magnitudes near 1 so that hundreds of factors neither overflow nor vanish (as test_reduce_gpu.py)
            xp = x if d not in O.FLOATS else (np.sign(x) * (0.8 + 0.4 * np.abs(np.tanh(x)))).astype(x.dtype)
            for op in ("sum", "mean", "max", "min", "prod", "logsumexp", "sum_square", "reducel1", "reducel2", "reducel3", "nansum",
                       "nanprod", "all", "any", "argmax", "argmin"):
                if op in ("prod", "nanprod") and d in ("f16", "bf16"):
                    continue  # products of hundreds of half-precision values over/underflow: nothing to compare
                if op == "logsumexp" and d in O.INTS:
                    continue  # exp of full-range integers overflows
                axes_list = [[sax]] if op.startswith("arg") else [[sax], list(range(len(shape)))]
                for axes in axes_list:
                    X, _ = check(xp if op in ("prod", "nanprod") else x, d, sax, op, axes)
            # a reduction that keeps the shard axis stays sharded and local
            if d == "f32":
                keep = X.sum([len(shape) - 1 if sax == 0 else 0])
                assert isinstance(keep, hb.ShardedTensor)
                want, _, _ = O.reduce("sum", x, d, [len(shape) - 1 if sax == 0 else 0])
                off, ln = hb.shard_bounds(shape[sax], world, rank)
                got = keep.local.to_cpu().numpy()
                np.testing.assert_allclose(got, np.take(want, range(off, off + ln), axis=keep.shard_axis), rtol=1e-5, atol=1e-5)

    # ---- BASELINE config 5's shape, scaled to what the CPU oracle finishes quickly: [4096, 16384] f32 over rows -------------
    x = rand(rng, (4096, 16384), "f32")
    for op, axes in (("sum", [0, 1]), ("mean", [0, 1]), ("sum", [0]), ("max", [0]), ("logsumexp", [0]), ("argmax", [0]), ("reducel2", [0, 1])):
        check(x, "f32", 0, op, axes, "cfg5-shaped ")
    # determinism: the same call twice gives the same bits (fixed slots, rank-ordered combine)
    X = hb.ShardedTensor.scatter_from_host(to_torch(x, "f32"), comm, 0, device=local)
    a, b = X.sum([0, 1]).to_cpu().numpy(), X.sum([0, 1]).to_cpu().numpy()
    assert a.tobytes() == b.tobytes()
    a, b = X.sum([0]).to_cpu().numpy(), X.sum([0]).to_cpu().numpy()
    assert a.tobytes() == b.tobytes()
    # halves of the same shape: accumulators are exchanged in f32, the result is rounded once
    for d in ("bf16", "f16"):
        xh = rand(rng, (1024, 16384), d)
        for op, axes in (("sum", [0]), ("mean", [0, 1]), ("mean", [0]), ("sum", [0, 1])):
            check(xh, d, 0, op, axes, "cfg5-shaped ")
    del x, X

    # ---- rows of 2^20 elements and more -------------------------------------------------------------------------------
    x = rand(rng, (4, 1 << 22), "f32")          # sharded along the LONG axis: each rank reduces runs of 2^22/world elements
    for op, axes in (("sum", [1]), ("mean", [1]), ("argmin", [1]), ("sum", [0, 1]), ("logsumexp", [1])):
        check(x, "f32", 1, op, axes, "long rows ")
    xh = rand(rng, (3, 1 << 22), "bf16")
    check(xh, "bf16", 1, "sum", [1], "long rows ")
    check(xh, "bf16", 1, "mean", [1], "long rows ")
    xi = rand(rng, (2, (1 << 21) + 5), "i64", -1000, 1000)
    check(xi, "i64", 1, "sum", [1], "long rows ")
    check(xi, "i64", 1, "argmax", [1], "long rows ")
    # more outputs than the mailboxes hold (2^20 > 65536): NCCL all-gather of accumulators + combine kernel
    x = rand(rng, (16, 1 << 20), "f32")
    for op, axes in (("sum", [0]), ("argmax", [0]), ("mean", [0])):
        check(x, "f32", 0, op, axes, "many outputs ")

    # ---- arg ops: ties everywhere, NaNs, all-NaN columns, ±inf ----------------------------------------------------------------
    x = rng.integers(0, 3, size=(64 * world, 300)).astype(np.float32)
    x[:, 5] = np.nan                      # all-NaN column → 0
    x[: 40 * world, 7] = np.nan           # NaN never wins, the first real maximum does
    x[:, 9] = -np.inf                     # all equal to the identity → 0
    x[3, 11], x[50 * world, 11] = np.inf, np.inf
    for op in ("argmax", "argmin"):
        check(x, "f32", 0, op, [0], "ties/NaN ")
    xi = rng.integers(0, 2, size=(33 * world + 1, 200)).astype(np.int64)
    for op in ("argmax", "argmin"):
        check(xi, "i64", 0, op, [0], "ties ")

    # ---- a strided `out` through the C ABI: accumulators go to a scratch, the combine kernel writes the view -------------------
    x = rand(rng, (50 * world, 96), "f32")
    X = hb.ShardedTensor.scatter_from_host(to_torch(x, "f32"), comm, 0, device=local)
    wide = hb.Tensor.zeros((96, 2), hb.F32, local)
    view = wide[:, 1]
    ax = (c_int32 * 1)(0)
    _ffi.check(hb.lib.hptb_reduce_sharded(comm.handle, _ffi.REDUCE_OPS["sum"], byref(X.local._c()), ax, 1, 0, X.offset, X.global_len,
                                           byref(view._c()), hb.get_stream()))
    got = wide.to_cpu().numpy()
    want, od, exact = O.reduce("sum", x, "f32", [0])
    assert_reduce_bar("sum", x, "f32", [0], got[:, 1], want, od, exact, "strided out")
    assert (got[:, 0] == 0).all(), "strided out: neighbours clobbered"
    checked += 1

    dist.barrier()
    comm.destroy()
    dist.destroy_process_group()
    assert not failures, f"{len(failures)} of {checked} cases failed:\n" + "\n".join(failures[:40])
    print(f"rank {rank}: sharded ok ({checked} cases, peer memory {p2p})")


if __name__ == "__main__":
    main()
