#!/usr/bin/env python
"""TMA-staged vs shared-memory-scatter transposing kernel, same calls, same process (HPTB_TUNE_NO_TMA toggles the
path per launch).  Device-resident, CUDA events, 50 launches after 5 warm-ups; f32 [8192,8192] (config 2) plus the
2-byte and binary cases the round-1 verdict lists."""
import os
import sys

os.environ["HPTB_TUNE"] = "1"
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpt_b200 as hb  # noqa: E402
from ctypes import byref  # noqa: E402
from hpt_b200 import _ffi  # noqa: E402

stream = torch.cuda.current_stream()
hb.set_stream(stream.cuda_stream)
PEAK = 6552.0
TD = {hb.F32: torch.float32, hb.F16: torch.float16, hb.BF16: torch.bfloat16, hb.I32: torch.int32}


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def both(fn):
    os.environ.pop("HPTB_TUNE_NO_TMA", None)
    a = min(timeit(fn), timeit(fn))
    os.environ["HPTB_TUNE_NO_TMA"] = "1"
    b = min(timeit(fn), timeit(fn))
    os.environ.pop("HPTB_TUNE_NO_TMA", None)
    return a, b


def dev(shape, dt):
    t = torch.randn(shape, device="cuda", dtype=torch.float32).to(TD[dt])
    return hb.Tensor.from_device_ptr(t.data_ptr(), dt, tuple(shape), keepalive=t)


print(f"{'case':58s} {'TMA us':>9s} {'frac':>6s} {'smem us':>9s} {'frac':>6s}")
for dt, name, n in ((hb.F32, "f32", 8192), (hb.F16, "f16", 8192), (hb.BF16, "bf16", 8192), (hb.F32, "f32", 16384), (hb.BF16, "bf16", 16384)):
    X = dev((n, n), dt)
    V = X.t()
    Y = hb.Tensor.empty((n, n), dt)
    B = dev((n, n), dt)
    esz = _ffi.DTYPE_SIZES[dt]
    for op in ("sin", "exp", "tanh", "gelu"):
        code = _ffi.UNARY_OPS[op]
        fn = lambda: _ffi.check(hb.lib.hptb_unary(X.ctx.handle, code, byref(V._c()), byref(Y._c()), 0.0, 0.0, hb.get_stream()))
        a, b = both(fn)
        nb = 2 * n * n * esz
        print(f"{name} [{n},{n}].t().{op}()".ljust(58) + f" {a:9.1f} {nb / a / 1e3 / PEAK:6.3f} {b:9.1f} {nb / b / 1e3 / PEAK:6.3f}", flush=True)
    fn = lambda: _ffi.check(hb.lib.hptb_copy(X.ctx.handle, byref(V._c()), byref(Y._c()), hb.get_stream()))
    a, b = both(fn)
    nb = 2 * n * n * esz
    print(f"{name} [{n},{n}].t().contiguous()".ljust(58) + f" {a:9.1f} {nb / a / 1e3 / PEAK:6.3f} {b:9.1f} {nb / b / 1e3 / PEAK:6.3f}", flush=True)
    fn = lambda: _ffi.check(hb.lib.hptb_binary(X.ctx.handle, _ffi.BINARY_OPS["add"], byref(V._c()), byref(B._c()), byref(Y._c()), hb.get_stream()))
    a, b = both(fn)
    nb = 3 * n * n * esz
    print(f"{name} [{n},{n}].t() + b".ljust(58) + f" {a:9.1f} {nb / a / 1e3 / PEAK:6.3f} {b:9.1f} {nb / b / 1e3 / PEAK:6.3f}", flush=True)
    del X, V, Y, B
