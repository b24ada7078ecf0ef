#!/usr/bin/env python
"""softmax band kernels: target KB of band per CTA (→ cluster size) sweep on the survey shapes."""
import os, sys
os.environ["HPTB_TUNE"] = "1"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpt_b200 as hb
from ctypes import byref
from hpt_b200 import _ffi
stream = torch.cuda.current_stream(); hb.set_stream(stream.cuda_stream)
def timeit(fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
for shape, axis in (((256, 131072), 1), ((4096, 8192), 0), ((1024, 32768), 1), ((2048, 16384), 0), ((8192, 4096), 0)):
    t = torch.randn(shape, device="cuda")
    X = hb.Tensor.from_device_ptr(t.data_ptr(), hb.F32, shape, keepalive=t)
    Y = hb.Tensor.empty(shape, hb.F32)
    res = []
    for kb in (16, 32, 48, 64, 96):
        os.environ["HPTB_TUNE_BAND_KB"] = str(kb)
        fn = lambda: _ffi.check(hb.lib.hptb_softmax(X.ctx.handle, byref(X._c()), axis, 0, byref(Y._c()), hb.get_stream()))
        res.append(f"{kb}KB:{min(timeit(fn), timeit(fn)):.1f}")
    os.environ.pop("HPTB_TUNE_BAND_KB", None)
    os.environ["HPTB_TUNE_NO_BAND"] = "1"
    res.append(f"noband:{timeit(fn):.1f}")
    os.environ.pop("HPTB_TUNE_NO_BAND", None)
    nb = 2 * t.numel() * 4
    print(f"f32 {shape} softmax(axis {axis})  ideal {nb / 6552e3:.1f} us   " + "  ".join(res), flush=True)
