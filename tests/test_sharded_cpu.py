"""World-size-2 (gloo, CPU) test of the multi-GPU path's host logic.

The NCCL data path needs GPUs (tests/test_sharded_gpu.py); what can be checked here is that the library's own
partition (`hptb_shard_bounds`) and exchange plan (`hptb_shard_plan_reduce` — the plan `hptb_reduce_sharded`
follows in comm.cpp) reproduce the global reduction when two ranks each reduce their shard and exchange exactly
what the plan says.  Local reductions are the oracle's (this is a test of the plan, not of a kernel); the exchange
runs through torch.distributed/gloo on 127.0.0.1.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [  # (op, dtype, shape, axes, shard_axis)
    ("sum", "f32", (10, 7), [0, 1], 0), ("sum", "i32", (9, 4, 5), [0], 0), ("sum", "f32", (9, 4, 5), [1], 0),
    ("mean", "f32", (11, 6), [0], 0), ("mean", "f64", (8, 5), [0, 1], 0), ("logsumexp", "f32", (13, 5), [0], 0),
    ("max", "f32", (10, 3), [0], 0), ("min", "i64", (7, 3), [0, 1], 0), ("prod", "i32", (6, 4), [0], 0),
    ("sum_square", "f32", (10, 4), [0], 0), ("argmax", "f32", (9, 6), [0], 0), ("argmin", "i32", (12, 5), [0], 0),
    ("argmax", "f32", (9, 6), [1], 0), ("sum", "f32", (5, 12), [1], 1), ("argmin", "f32", (5, 12), [1], 1),
    ("reducel2", "f32", (10, 7), [0], 0), ("reducel3", "f64", (9, 5), [0, 1], 0), ("reducel2", "i32", (8, 6), [0], 0),
    ("reducel1", "f32", (10, 7), [0, 1], 0),
]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import hpt_b200 as hb
        from oracle import hpt_oracle as O
        rng = np.random.default_rng(7)  # same data on every rank
        for op, d, shape, axes, sax in CASES:
            if d in O.INTS:
                x = rng.integers(0, 4, size=shape).astype(O.NP[d])  # many ties for the arg rule
            else:
                x = rng.integers(-3, 4, size=shape).astype(O.NP[d]) if op.startswith("arg") else rng.standard_normal(shape).astype(O.NP[d])
            if op.startswith("arg") and d == "f32":
                x[0, :] = np.nan  # NaN never wins
            want, od, exact = O.reduce(op, x, d, axes)
            off, ln = hb.shard_bounds(shape[sax], world, rank)
            local = np.take(x, range(off, off + ln), axis=sax)
            plan = hb.shard_plan(op, axes, sax, world)
            assert plan["crosses"] == (sax in axes)
            if not plan["crosses"]:
                got_local, _, _ = O.reduce(op, local, d, axes)
                new_ax = sax - sum(1 for a in axes if a < sax)
                np.testing.assert_array_equal(got_local, np.take(want, range(off, off + ln), axis=new_ax))
                continue
            coll = plan["collective"]
            if coll == "allgather_arg":
                (ax,) = axes
                v, _, _ = O.reduce("max" if op == "argmax" else "min", local, d, axes)
                i, _, _ = O.reduce(op, local, d, axes)
                i = i + off
                vs = [torch.empty_like(torch.from_numpy(np.ascontiguousarray(v))) for _ in range(world)]
                is_ = [torch.empty_like(torch.from_numpy(i)) for _ in range(world)]
                dist.all_gather(vs, torch.from_numpy(np.ascontiguousarray(v)))
                dist.all_gather(is_, torch.from_numpy(i))
                # rank-ordered strict combine from the identity with index 0 (sharded.cu arg_combine_kernel)
                ident = (-np.inf if op == "argmax" else np.inf) if d not in O.INTS else (np.iinfo(O.NP[d]).min if op == "argmax" else np.iinfo(O.NP[d]).max)
                best = np.full(v.shape, ident, dtype=np.float64 if d not in O.INTS else O.NP[d])
                bi = np.zeros(v.shape, dtype=np.int64)
                for r in range(world):
                    vr, ir = vs[r].numpy(), is_[r].numpy()
                    better = (vr > best) if op == "argmax" else (vr < best)
                    best = np.where(better, vr, best)
                    bi = np.where(better, ir, bi)
                np.testing.assert_array_equal(bi, want)
                continue
            # allreduce family
            if plan["global_count"]:
                n = 1
                for a in axes:
                    n *= shape[a]
                part = O.reduce_f64("sum", local, d, axes) / n
            elif plan["pre_exp"]:
                part = np.exp(O.reduce_f64("logsumexp", local, d, axes))
            elif plan["post_root"]:
                part = O.reduce_f64(op, local, d, axes) ** plan["post_root"]  # the unrooted power sum
            elif d in O.INTS or op in ("max", "min"):
                part, _, _ = O.reduce(op, local, d, axes)
            else:
                part = O.reduce_f64(op, local, d, axes)
            t = torch.from_numpy(np.ascontiguousarray(part).copy())
            dist.all_reduce(t, op={"allreduce_sum": dist.ReduceOp.SUM, "allreduce_prod": dist.ReduceOp.PRODUCT,
                                   "allreduce_max": dist.ReduceOp.MAX, "allreduce_min": dist.ReduceOp.MIN}[coll])
            got = t.numpy()
            if plan["post_ln"]:
                got = np.log(got)
            if plan["post_root"]:
                got = got ** (1.0 / plan["post_root"])
            if exact:
                np.testing.assert_array_equal(got.astype(want.dtype), want)
            else:
                np.testing.assert_allclose(got, want.astype(np.float64), rtol=1e-5, atol=1e-6)
        q.put((rank, "ok"))
    except Exception as ex:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_the_axis():
    import hpt_b200 as hb
    for n in (0, 1, 7, 8, 262144):
        for world in (1, 2, 3, 8):
            blocks = [hb.shard_bounds(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and sum(b[1] for b in blocks) == n
            for a, b in zip(blocks, blocks[1:]):
                assert a[0] + a[1] == b[0] and a[1] >= b[1] >= a[1] - 1
    with pytest.raises(hb.HptError):
        hb.shard_bounds(4, 2, 2)


def test_plan_is_local_when_the_shard_axis_is_kept_or_world_is_one():
    import hpt_b200 as hb
    assert hb.shard_plan("sum", [1], 0, 8)["collective"] == "none"
    assert hb.shard_plan("sum", [0], 0, 1)["collective"] == "none"
    assert hb.shard_plan("argmax", [0], 0, 2)["collective"] == "allgather_arg"
    p = hb.shard_plan("logsumexp", [0, 1], 0, 4)
    assert p["collective"] == "allreduce_sum" and p["pre_exp"] and p["post_ln"]


def test_two_ranks_follow_the_plan_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}:\n{msg}"
