"""bench_rows.py — the other BASELINE.json configurations (1–4), reported under "rows" by `bench.py` (every run).

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

    # ---- config 2: f32 [8192,8192] transposed view: sin, exp, max(0), argmax(0) (input 268 MB > L2) -----------------------
    def rows_cfg2():
        X = dev_randn((8192, 8192))
        V = X.t()
        Y = T.empty((8192, 8192), F32, local)
        Mx, Ai = T.empty((8192,), F32, local), T.empty((8192,), I64, local)

        def unary_into(op):
            _ffi.check(hb.lib.hptb_unary(X.ctx.handle, _ffi.UNARY_OPS[op], byref(V._c()), byref(Y._c()), 0.0, 0.0, hb.get_stream()))
        for op in ("sin", "exp"):
            us = _timeit(torch, stream, [lambda op=op: unary_into(op)], 100)
            add_row("cfg2", f"{op} f32 [8192,8192] transposed view → contiguous", us, 536870912, 67108864, "input and output 268 MB each > L2")
        us = _timeit(torch, stream, [lambda: V._reduce("max", [0], out=Mx)], 100)
        add_row("cfg2", "max(axis 0) f32 [8192,8192] transposed view", us, 268468224, 67108864, "")
        us = _timeit(torch, stream, [lambda: V._reduce("argmax", [0], out=Ai)], 100)
        add_row("cfg2", "argmax(axis 0) f32 [8192,8192] transposed view", us, 268500992, 67108864, "")
        del X, V, Y

    for cfg, fn in (("cfg1", rows_cfg1), ("cfg2", rows_cfg2), ("cfg3", rows_cfg3), ("cfg4", rows_cfg4)):
        if want(cfg):
            fn()
    return rows
