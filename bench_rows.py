"""bench_rows.py — the other BASELINE.json configurations, reported under "rows" by `bench.py --rows`.

Every row: device-resident inputs, CUDA-event timing on the launch stream, ≥3 warm-ups, median-free mean over
`reps` launches; working sets that fit the 126 MB L2 (configs 1 and 4) rotate over 4 independent buffer sets so
the number is an HBM number.  Bytes are the algorithmic bytes of SURVEY.md §8d / BASELINE.md §3.
torch is used only to generate synthetic inputs on the device and for the barrier/max-over-ranks plumbing.
"""
import ctypes
from ctypes import byref, c_int32, c_int64, c_void_p

import torch


def _timeit(torch, stream, fns, reps, warm=3):
    """fns: list of callables rotated round-robin; returns mean µs per call."""
    for i in range(warm * len(fns)):
        fns[i % len(fns)]()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(reps):
        fns[i % len(fns)]()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def run_rows(hb, torch, dist, world, rank, local, stream, peak):
    from hpt_b200 import _ffi
    rows = []
    dev = torch.device("cuda", local)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    T = hb.Tensor
    F32, F64, I64, BF16, F16 = hb.F32, hb.F64, hb.I64, hb.BF16, hb.F16
    tdt = {F32: torch.float32, BF16: torch.bfloat16, F16: torch.float16}

    def dev_randn(shape, dt=F32):
        t = torch.randn(shape, generator=g, device=dev, dtype=torch.float32).to(tdt[dt])
        return T.from_device_ptr(t.data_ptr(), dt, tuple(shape), device=local, keepalive=t)

    def add_row(cfg, op, us, nbytes, nelem, note=""):
        if world > 1:
            t = torch.tensor([us], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            us = t.item()
        gbs = world * nbytes / (us * 1e-6) / 1e9
        rows.append({"config": cfg, "op": op, "us": round(us, 2), "gbs": round(gbs, 1), "frac_of_measured_peak": round(gbs / (peak * world), 4),
                     "gelem_per_s": round(world * nelem / (us * 1e-6) / 1e9, 2), "algorithmic_bytes": nbytes, "note": note})

    import os
    only = set(c for c in os.environ.get("HPTB_ROWS", "").split(",") if c)  # e.g. HPTB_ROWS=cfg5 (development)

    def want(cfg):
        return not only or cfg in only

    R = 4
    # ---- config 1: f32 [4096,4096] + [1,4096], then sum over axis 1 -------------------------------------
    def rows_cfg1():
        A = [dev_randn((4096, 4096)) for _ in range(R)]
        B = [dev_randn((1, 4096)) for _ in range(R)]
        C = [T.empty((4096, 4096), F32, local) for _ in range(R)]
        S = [T.empty((4096,), F32, local) for _ in range(R)]
        us = _timeit(torch, stream, [lambda i=i: A[i].add_(B[i], C[i]) for i in range(R)], 200)
        add_row("cfg1", "add f32 [4096,4096]+[1,4096]", us, 134234112, 16777216, "rotating 4 buffer sets")
        us = _timeit(torch, stream, [lambda i=i: C[i].sum_([1], False, True, S[i]) for i in range(R)], 200)
        add_row("cfg1", "sum(axis 1) f32 [4096,4096]", us, 67125248, 16777216, "rotating 4 buffer sets")
        ax1 = (c_int32 * 1)(1)
        fused = [lambda i=i: _ffi.check(hb.lib.hptb_binary_reduce(A[i].ctx.handle, _ffi.BINARY_OPS["add"], _ffi.REDUCE_OPS["sum"], byref(A[i]._c()),
                                                                  byref(B[i]._c()), ax1, 1, byref(S[i]._c()), 1, hb.get_stream())) for i in range(R)]
        us = _timeit(torch, stream, fused, 200)
        add_row("cfg1", "EXTRA fused (A + B).sum(1) in one pass (hptb_binary_reduce)", us, 67125248 + 16384, 16777216,
                "bytes = one read of A and B + the [4096] result; the unfused pair above moves 201 MB")
        del A, B, C, S

    # ---- config 3: bf16/f16 [64,512,56,56] NCHW viewed NHWC, mean over (0,1,2) ---------------------------------
    def rows_cfg3():
        for dt, name in ((BF16, "bf16"), (F16, "f16")):
            X = [dev_randn((64, 512, 56, 56), dt) for _ in range(2)]
            V = [x.permute([0, 2, 3, 1]) for x in X]
            Mo = [T.empty((512,), dt, local) for _ in range(2)]
            us = _timeit(torch, stream, [lambda i=i: V[i]._reduce("mean", [0, 1, 2], out=Mo[i]) for i in range(2)], 100)
            add_row("cfg3", f"mean(0,1,2) {name} NCHW→NHWC view [64,56,56,512]", us, 205521920, 102760448, "f32 accumulate")
            us = _timeit(torch, stream, [lambda i=i: V[i].mean_var([0, 1, 2]) for i in range(2)], 100)
            add_row("cfg3", f"mean_var(0,1,2) {name} (extension, fused single read)", us, 205521920 + 1024, 102760448, "one launch")
            del X, V

    # ---- config 4: f32 [32,128,4096] softmax / logsumexp over the last axis, + i64 → f64 ------------------------
    def rows_cfg4():
        X = [dev_randn((32, 128, 4096)) for _ in range(R)]
        Y = [T.empty((32, 128, 4096), F32, local) for _ in range(R)]

        def softmax_into(i):
            _ffi.check(hb.lib.hptb_softmax(X[i].ctx.handle, byref(X[i]._c()), 2, 0, byref(Y[i]._c()), hb.get_stream()))
        us = _timeit(torch, stream, [lambda i=i: softmax_into(i) for i in range(R)], 200)
        add_row("cfg4", "softmax(-1) f32 [32,128,4096]", us, 134217728, 16777216, "rotating 4 buffer sets")
        L = [T.empty((32, 128), F32, local) for _ in range(R)]
        us = _timeit(torch, stream, [lambda i=i: X[i]._reduce("logsumexp", [-1], out=L[i]) for i in range(R)], 200)
        add_row("cfg4", "logsumexp(-1) f32 [32,128,4096]", us, 67125248, 16777216, "rotating 4 buffer sets")
        kt = torch.randint(-1000, 1000, (4096,), generator=g, device=dev, dtype=torch.int64)
        Kt = T.from_device_ptr(kt.data_ptr(), I64, (4096,), device=local, keepalive=kt)
        Z = [T.empty((32, 128, 4096), F64, local) for _ in range(3)]
        us = _timeit(torch, stream, [lambda i=i: X[i].add_(Kt, Z[i]) for i in range(3)], 200)
        add_row("cfg4", "add f32 [32,128,4096] + i64 [4096] → f64", us, 201359360, 16777216, "normal_promote f32⊕i64→f64")
        del X, Y, Z

    # ---- config 5: f32 [262144,16384] sharded over the outer axis, full sum / mean, sum(0) -------------------------
    def rows_cfg5():
        rows_total, cols = 262144, 16384
        rows_local = rows_total // world
        big = torch.empty((rows_local, cols), device=dev, dtype=torch.float32)
        for r0 in range(0, rows_local, 16384):  # generate in slabs to bound the temporary
            big[r0:r0 + 16384].normal_(generator=g)
        Xs = T.from_device_ptr(big.data_ptr(), F32, (rows_local, cols), device=local, keepalive=big)
        comm = None
        if world > 1:
            idbuf = (ctypes.c_char * 128)()
            if rank == 0:
                _ffi.check(hb.lib.hptb_comm_unique_id(idbuf))
            obj = [bytes(idbuf)]
            dist.broadcast_object_list(obj, src=0)
            comm = c_void_p()
            _ffi.check(hb.lib.hptb_comm_init_rank(Xs.ctx.handle, world, rank, ctypes.c_char_p(obj[0]), byref(comm)))

        def sharded(op, axes, out):
            ax = (c_int32 * len(axes))(*axes)
            if comm is None:
                _ffi.check(hb.lib.hptb_reduce(Xs.ctx.handle, _ffi.REDUCE_OPS[op], byref(Xs._c()), ax, len(axes), byref(out._c()), 1, hb.get_stream()))
            else:
                _ffi.check(hb.lib.hptb_reduce_sharded(comm, _ffi.REDUCE_OPS[op], byref(Xs._c()), ax, len(axes), 0, rank * rows_local,
                                                       rows_total, byref(out._c()), hb.get_stream()))
        o1 = T.empty((1,), F32, local)
        oc = T.empty((cols,), F32, local)
        nbytes = rows_local * cols * 4  # per rank; add_row multiplies by world
        for op, axes, out, label in (("sum", [0, 1], o1, "sum() all axes"), ("mean", [0, 1], o1, "mean() all axes"), ("sum", [0], oc, "sum(axis 0) → [16384]")):
            us = _timeit(torch, stream, [lambda: sharded(op, axes, out)], 20)
            add_row("cfg5", f"{label} f32 [262144,16384] sharded over {world} GPU(s)", us, nbytes + 4, rows_local * cols,
                    ("strong scaling: total size fixed; partials exchanged through " +
                     ("peer-mapped mailboxes (one kernel per rank, NVLink stores)" if hb.lib.hptb_comm_uses_peer_memory(comm) else "ncclAllReduce"))
                    if world > 1 else "single GPU, no exchange")
        # parity of the sharded sum against an f64 accumulation of the same data (computed with torch on the device)
        sharded("sum", [0, 1], o1)
        torch.cuda.synchronize()
        got = float(o1.to_cpu().item())
        ref = big.sum(dtype=torch.float64)
        mag = big.abs().sum(dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ref)
            dist.all_reduce(mag)
        rel = abs(got - ref.item()) / mag.item()
        rows.append({"config": "cfg5", "op": "parity: |sum − f64 sum| / Σ|x|", "value": rel, "bound_1e-6_log2n": 1e-6 * 32, "ok": rel <= 1e-6 * 32})
        if comm is not None:
            hb.lib.hptb_comm_destroy(comm)

    for cfg, fn in (("cfg1", rows_cfg1), ("cfg3", rows_cfg3), ("cfg4", rows_cfg4), ("cfg5", rows_cfg5)):
        if want(cfg):
            fn()
    return rows
