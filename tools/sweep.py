#!/usr/bin/env python
"""tools/sweep.py — launch-shape sweep of the hot-path kernels on one B200 (development tool, not the bench).

    HPTB_TUNE=1 [HPTB_LIB_VARIANT=u8] python tools/sweep.py [--out gpurun_out/sweep.jsonl] [--cases a,b,…]

For each BASELINE.json configuration it times the C-ABI call with the policy's own choice and with the launch
shape forced through HPTB_TUNE_G / HPTB_TUNE_S (reduce.cuh), rotating buffer sets where the working set fits the
126 MB L2, and times the same op in torch beside it as a YARDSTICK for what a tuned library kernel reaches at
that size (torch is never on the product path).  One JSON line per measurement.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("HPTB_TUNE", "1")

import torch  # noqa: E402

import hpt_b200 as hb  # noqa: E402
from hpt_b200 import _ffi  # noqa: E402

PEAK = 6650.0
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fns, reps, warm=3):
    stream = torch.cuda.current_stream()
    for i in range(warm * len(fns)):
        fns[i % len(fns)]()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t0 = time.perf_counter()
    for i in range(reps):
        fns[i % len(fns)]()
    timeit.host_us = (time.perf_counter() - t0) * 1e6 / reps  # enqueue cost: ≈ the device time means host-bound
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


timeit.host_us = 0.0


def prep(fn, *args):
    """a zero-argument callable around one C-ABI entry with its ctypes arguments built once"""
    def run():
        rc = fn(*args)
        if rc:
            _ffi.check(rc)
    run.keep = args
    return run


def h():
    return hb.lib, hb.get_stream()


def c_binary(op, a, b, out):
    lib, st = h()
    return prep(lib.hptb_binary, a.ctx.handle, _ffi.BINARY_OPS[op], _ffi.byref(a._c()), _ffi.byref(b._c()), _ffi.byref(out._c()), st)


def c_unary(op, a, out):
    lib, st = h()
    return prep(lib.hptb_unary, a.ctx.handle, _ffi.UNARY_OPS[op], _ffi.byref(a._c()), _ffi.byref(out._c()), _ffi.c_double(0.0), _ffi.c_double(0.0), st)


def c_copy(a, out):
    lib, st = h()
    return prep(lib.hptb_copy, a.ctx.handle, _ffi.byref(a._c()), _ffi.byref(out._c()), st)


def c_reduce(op, a, axes, out):
    lib, st = h()
    ax = (_ffi.c_int32 * len(axes))(*axes)
    return prep(lib.hptb_reduce, a.ctx.handle, _ffi.REDUCE_OPS[op], _ffi.byref(a._c()), ax, len(axes), _ffi.byref(out._c()), 1, st)


def c_softmax(a, axis, out):
    lib, st = h()
    return prep(lib.hptb_softmax, a.ctx.handle, _ffi.byref(a._c()), axis, 0, _ffi.byref(out._c()), st)


def c_meanvar(a, axes, m, v):
    lib, st = h()
    ax = (_ffi.c_int32 * len(axes))(*axes)
    return prep(lib.hptb_mean_var, a.ctx.handle, _ffi.byref(a._c()), ax, len(axes), _ffi.byref(m._c()), _ffi.byref(v._c()), st)


class Sweep:
    def __init__(self, out):
        self.f = open(out, "a")
        self.variant = os.environ.get("HPTB_LIB_VARIANT", "") or "default"

    def emit(self, case, impl, knobs, us, nbytes):
        gbs = nbytes / (us * 1e-6) / 1e9
        rec = {"case": case, "impl": impl, "variant": self.variant, "knobs": knobs, "us": round(us, 2), "gbs": round(gbs, 1),
               "frac": round(gbs / PEAK, 4), "host_us": round(timeit.host_us, 2)}
        self.f.write(json.dumps(rec) + "\n")
        self.f.flush()
        print(f"{case:28s} {impl:8s} {self.variant:8s} {str(knobs):24s} {us:9.2f} us {gbs:8.1f} GB/s {gbs / PEAK:6.3f}  host {timeit.host_us:6.1f} us", flush=True)

    def knob_sweep(self, case, fns, reps, nbytes, G=(), S=(), flags=(), GS=()):
        for k in ("HPTB_TUNE_G", "HPTB_TUNE_S") + tuple(flags):
            os.environ.pop(k, None)
        self.emit(case, "hptb", {}, timeit(fns, reps), nbytes)
        for fl in flags:
            os.environ[fl] = "1"
            self.emit(case, "hptb", {fl[10:]: 1}, timeit(fns, reps), nbytes)
            os.environ.pop(fl, None)
        for g in G:
            os.environ["HPTB_TUNE_G"] = str(g)
            self.emit(case, "hptb", {"G": g}, timeit(fns, reps), nbytes)
        os.environ.pop("HPTB_TUNE_G", None)
        for s in S:
            os.environ["HPTB_TUNE_S"] = str(s)
            self.emit(case, "hptb", {"S": s}, timeit(fns, reps), nbytes)
        os.environ.pop("HPTB_TUNE_S", None)
        for g, s_ in GS:
            os.environ["HPTB_TUNE_G"], os.environ["HPTB_TUNE_S"] = str(g), str(s_)
            self.emit(case, "hptb", {"G": g, "S": s_}, timeit(fns, reps), nbytes)
        os.environ.pop("HPTB_TUNE_G", None)
        os.environ.pop("HPTB_TUNE_S", None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    ap.add_argument("--cases", default="")
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    want = set(c for c in args.cases.split(",") if c)
    sw = Sweep(args.out)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    hb.set_stream(torch.cuda.current_stream().cuda_stream)
    T = hb.Tensor
    F32, F64, I64, BF16 = hb.F32, hb.F64, hb.I64, hb.BF16
    g = torch.Generator(device=dev).manual_seed(1234)
    tdt = {F32: torch.float32, BF16: torch.bfloat16}

    def wrap(t, dt):
        return T.from_device_ptr(t.data_ptr(), dt, tuple(t.shape), device=0, keepalive=t)

    def randn(shape, dt=F32):
        return torch.randn(shape, generator=g, device=dev, dtype=torch.float32).to(tdt[dt])

    def on(name):
        return not want or name in want

    R = 4
    if on("cfg1"):
        a = [randn((4096, 4096)) for _ in range(R)]
        b = [randn((1, 4096)) for _ in range(R)]
        c = [torch.empty((4096, 4096), device=dev) for _ in range(R)]
        s = [torch.empty((4096,), device=dev) for _ in range(R)]
        A, B, C, S_ = [wrap(x, F32) for x in a], [wrap(x, F32) for x in b], [wrap(x, F32) for x in c], [wrap(x, F32) for x in s]
        sw.knob_sweep("cfg1.add_bcast", [c_binary("add", A[i], B[i], C[i]) for i in range(R)], 200, 134234112)
        sw.emit("cfg1.add_bcast", "torch", {}, timeit([lambda i=i: torch.add(a[i], b[i], out=c[i]) for i in range(R)], 200), 134234112)
        sw.knob_sweep("cfg1.add_same", [c_binary("add", A[i], A[(i + 1) % R], C[i]) for i in range(R)], 200, 3 * 67108864)
        sw.emit("cfg1.add_same", "torch", {}, timeit([lambda i=i: torch.add(a[i], a[(i + 1) % R], out=c[i]) for i in range(R)], 200), 3 * 67108864)
        sw.knob_sweep("cfg1.sum_axis1", [c_reduce("sum", A[i], [1], S_[i]) for i in range(R)], 200, 67125248, G=(32, 64, 128, 256), flags=("HPTB_TUNE_NOLEAN",))
        sw.emit("cfg1.sum_axis1", "torch", {}, timeit([lambda i=i: torch.sum(a[i], 1, out=s[i]) for i in range(R)], 200), 67125248)
        sw.knob_sweep("cfg1.sum_axis0", [c_reduce("sum", A[i], [0], S_[i]) for i in range(R)], 200, 67125248, S=(1, 4, 9, 18, 37))
        sw.emit("cfg1.sum_axis0", "torch", {}, timeit([lambda i=i: torch.sum(a[i], 0, out=s[i]) for i in range(R)], 200), 67125248)
        del a, b, c, s, A, B, C, S_
    if on("big"):
        a = [randn((16384, 16384)) for _ in range(2)]
        c = torch.empty((16384, 16384), device=dev)
        A, C = [wrap(x, F32) for x in a], wrap(c, F32)
        sw.knob_sweep("big.add_same_1GiB", [c_binary("add", A[0], A[1], C)], 20, 3 * 16384 * 16384 * 4)
        sw.emit("big.add_same_1GiB", "torch", {}, timeit([lambda: torch.add(a[0], a[1], out=c)], 20), 3 * 16384 * 16384 * 4)
        sw.knob_sweep("big.sin_1GiB", [c_unary("sin", A[0], C)], 20, 2 * 16384 * 16384 * 4)
        sw.emit("big.sin_1GiB", "torch", {}, timeit([lambda: torch.sin(a[0], out=c)], 20), 2 * 16384 * 16384 * 4)
        sw.emit("big.copy_1GiB", "torch", {}, timeit([lambda: c.copy_(a[0])], 20), 2 * 16384 * 16384 * 4)
        del a, c, A, C
    if on("cfg2"):
        n = 8192
        x = randn((n, n))
        y = torch.empty((n, n), device=dev)
        m = torch.empty((n,), device=dev)
        ix = torch.empty((n,), device=dev, dtype=torch.int64)
        X, Y, Mx, Ix = wrap(x, F32), wrap(y, F32), wrap(m, F32), wrap(ix, I64)
        V = X.t()
        sw.knob_sweep("cfg2.sin_T", [c_unary("sin", V, Y)], 50, 536870912, flags=("HPTB_TUNE_NO_SMEMT",))
        sw.emit("cfg2.sin_T", "torch", {}, timeit([lambda: torch.sin(x.t(), out=y)], 50), 536870912)
        sw.knob_sweep("cfg2.exp_T", [c_unary("exp", V, Y)], 50, 536870912)
        sw.knob_sweep("cfg2.copy_T", [c_copy(V, Y)], 50, 536870912, flags=("HPTB_TUNE_NO_SMEMT",))
        sw.emit("cfg2.copy_T", "torch", {}, timeit([lambda: y.copy_(x.t())], 50), 536870912)

        def red(op, out):
            ax = (_ffi.c_int32 * 1)(0)
            _ffi.check(hb.lib.hptb_reduce(V.ctx.handle, _ffi.REDUCE_OPS[op], _ffi.byref(V._c()), ax, 1, _ffi.byref(out._c()), 1, hb.get_stream()))
        sw.knob_sweep("cfg2.max_T0", [c_reduce("max", V, [0], Mx)], 50, 268468224, G=(32, 64, 128, 256), flags=("HPTB_TUNE_NOLEAN",))
        sw.emit("cfg2.max_T0", "torch", {}, timeit([lambda: torch.amax(x.t(), 0, out=m)], 50), 268468224)
        sw.knob_sweep("cfg2.argmax_T0", [c_reduce("argmax", V, [0], Ix)], 50, 268500992, G=(32, 64, 128, 256), flags=("HPTB_TUNE_NOLEAN",))
        sw.emit("cfg2.argmax_T0", "torch", {}, timeit([lambda: torch.argmax(x.t(), 0, out=ix)], 50), 268500992)
        # the same reductions on the contiguous tensor over axis 0 (cols kernel)
        def redc(op, out):
            ax = (_ffi.c_int32 * 1)(0)
            _ffi.check(hb.lib.hptb_reduce(X.ctx.handle, _ffi.REDUCE_OPS[op], _ffi.byref(X._c()), ax, 1, _ffi.byref(out._c()), 1, hb.get_stream()))
        sw.knob_sweep("cfg2.max_C0", [c_reduce("max", X, [0], Mx)], 50, 268468224, S=(1, 4, 9, 18, 37))
        sw.emit("cfg2.max_C0", "torch", {}, timeit([lambda: torch.amax(x, 0, out=m)], 50), 268468224)
        sw.knob_sweep("cfg2.argmax_C0", [c_reduce("argmax", X, [0], Ix)], 50, 268500992, S=(4, 9, 18))
        del x, y, X, Y, V
    if on("cfg3"):
        x = [randn((64, 512, 56, 56), BF16) for _ in range(2)]
        X = [wrap(t, BF16) for t in x]
        V = [t.permute([0, 2, 3, 1]) for t in X]
        o = [T.empty((512,), BF16, 0) for _ in range(2)]
        o2 = [T.empty((512,), BF16, 0) for _ in range(2)]

        def mean(i):
            ax = (_ffi.c_int32 * 3)(0, 1, 2)
            _ffi.check(hb.lib.hptb_reduce(V[i].ctx.handle, _ffi.REDUCE_OPS["mean"], _ffi.byref(V[i]._c()), ax, 3, _ffi.byref(o[i]._c()), 1, hb.get_stream()))

        def meanvar(i):
            ax = (_ffi.c_int32 * 3)(0, 1, 2)
            _ffi.check(hb.lib.hptb_mean_var(V[i].ctx.handle, _ffi.byref(V[i]._c()), ax, 3, _ffi.byref(o[i]._c()), _ffi.byref(o2[i]._c()), hb.get_stream()))
        sw.knob_sweep("cfg3.mean_bf16", [c_reduce("mean", V[i], [0, 1, 2], o[i]) for i in range(2)], 100, 205521920, S=(4, 8, 16),
                      GS=((256, 1), (256, 2), (256, 3), (256, 4), (256, 8), (128, 1), (64, 1)))
        sw.emit("cfg3.mean_bf16", "torch", {}, timeit([lambda i=i: x[i].mean((0, 2, 3)) for i in range(2)], 100), 205521920)
        sw.knob_sweep("cfg3.meanvar_bf16", [c_meanvar(V[i], [0, 1, 2], o[i], o2[i]) for i in range(2)], 100, 205522944, S=(16, 49, 64, 128))
        sw.emit("cfg3.meanvar_bf16", "torch", {}, timeit([lambda i=i: torch.var_mean(x[i], (0, 2, 3), correction=0) for i in range(2)], 100), 205522944)
        del x, X, V
    if on("cfg4"):
        x = [randn((32, 128, 4096)) for _ in range(R)]
        y = [torch.empty((32, 128, 4096), device=dev) for _ in range(R)]
        l = [torch.empty((32, 128), device=dev) for _ in range(R)]
        X, Y, L = [wrap(t, F32) for t in x], [wrap(t, F32) for t in y], [wrap(t, F32) for t in l]

        def softmax(i):
            _ffi.check(hb.lib.hptb_softmax(X[i].ctx.handle, _ffi.byref(X[i]._c()), 2, 0, _ffi.byref(Y[i]._c()), hb.get_stream()))

        def lse(i):
            ax = (_ffi.c_int32 * 1)(2)
            _ffi.check(hb.lib.hptb_reduce(X[i].ctx.handle, _ffi.REDUCE_OPS["logsumexp"], _ffi.byref(X[i]._c()), ax, 1, _ffi.byref(L[i]._c()), 1, hb.get_stream()))
        sw.knob_sweep("cfg4.softmax", [c_softmax(X[i], 2, Y[i]) for i in range(R)], 200, 134217728)
        sw.emit("cfg4.softmax", "torch", {}, timeit([lambda i=i: torch.softmax(x[i], -1, out=y[i]) for i in range(R)], 200), 134217728)
        sw.knob_sweep("cfg4.logsumexp", [c_reduce("logsumexp", X[i], [2], L[i]) for i in range(R)], 200, 67125248, G=(32, 64, 128, 256), flags=("HPTB_TUNE_NOLEAN",))
        sw.emit("cfg4.logsumexp", "torch", {}, timeit([lambda i=i: torch.logsumexp(x[i], -1, out=l[i]) for i in range(R)], 200), 67125248)
        kt = torch.randint(-1000, 1000, (4096,), generator=g, device=dev, dtype=torch.int64)
        K = wrap(kt, I64)
        z = [torch.empty((32, 128, 4096), device=dev, dtype=torch.float64) for _ in range(3)]
        Z = [wrap(t, F64) for t in z]
        sw.knob_sweep("cfg4.add_f32_i64_f64", [c_binary("add", X[i], K, Z[i]) for i in range(3)], 200, 201359360)
        kd = kt.to(torch.float64)
        sw.emit("cfg4.add_f32_i64_f64", "torch", {"note": "f32+f64 bcast"}, timeit([lambda i=i: torch.add(x[i], kd, out=z[i]) for i in range(3)], 200), 201359360)
        del x, y, X, Y, z, Z
    if on("cfg5"):
        rows, cols = 262144, 16384
        big = torch.empty((rows, cols), device=dev, dtype=torch.float32)
        for r0 in range(0, rows, 16384):
            big[r0:r0 + 16384].normal_(generator=g)
        Xs = wrap(big, F32)
        o1 = T.empty((1,), F32, 0)
        oc = T.empty((cols,), F32, 0)
        orow = T.empty((rows,), F32, 0)
        t1 = torch.empty((), device=dev)
        tc = torch.empty((cols,), device=dev)
        nb = rows * cols * 4

        def red(axes, out):
            ax = (_ffi.c_int32 * len(axes))(*axes)
            _ffi.check(hb.lib.hptb_reduce(Xs.ctx.handle, 0, _ffi.byref(Xs._c()), ax, len(axes), _ffi.byref(out._c()), 1, hb.get_stream()))
        sw.knob_sweep("cfg5.sum_all", [c_reduce("sum", Xs, [0, 1], o1)], 10, nb, S=(1184, 2368, 4736, 8192))
        sw.emit("cfg5.sum_all", "torch", {}, timeit([lambda: torch.sum(big)], 10), nb)
        sw.knob_sweep("cfg5.sum_axis0", [c_reduce("sum", Xs, [0], oc)], 10, nb, S=(2, 4, 5, 9, 18, 37, 74))
        sw.emit("cfg5.sum_axis0", "torch", {}, timeit([lambda: torch.sum(big, 0, out=tc)], 10), nb)
        sw.knob_sweep("cfg5.sum_axis1", [c_reduce("sum", Xs, [1], orow)], 10, nb, G=(32, 64, 128, 256))
        del big, Xs


if __name__ == "__main__":
    main()
