#!/usr/bin/env python
"""softmax over a strided axis: the streaming two-launch kernels (statistics, then apply) against the cluster band
kernel, over the split count S and the tile width; and the streamed-row kernel against the row band."""
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
def env(**kw):
    for k in ("HPTB_TUNE_NO_BAND", "HPTB_TUNE_SMC_S", "HPTB_TUNE_SMC_TX"): os.environ.pop(k, None)
    for k, v in kw.items(): os.environ[k] = str(v)
cases = [((4096, 8192), 0, torch.float32), ((8192, 4096), 0, torch.float32), ((2048, 16384), 0, torch.float32),
         ((16384, 2048), 0, torch.float32), ((64, 4096, 512), 1, torch.float32), ((4096, 8192), 0, torch.bfloat16),
         ((256, 131072), 1, torch.float32), ((1024, 32768), 1, torch.float32), ((64, 524288), 1, torch.float32)]
DT = {torch.float32: hb.F32, torch.bfloat16: hb.BF16}
for shape, axis, dt in cases:
    t = torch.randn(shape, device="cuda").to(dt)
    X = hb.Tensor.from_device_ptr(t.data_ptr(), DT[dt], shape, keepalive=t)
    Y = hb.Tensor.empty(shape, DT[dt])
    fn = lambda: _ffi.check(hb.lib.hptb_softmax(X.ctx.handle, byref(X._c()), axis, 0, byref(Y._c()), hb.get_stream()))
    res = []
    env(); res.append(f"default:{min(timeit(fn), timeit(fn)):.1f}")
    env(HPTB_TUNE_NO_BAND=1); res.append(f"noband:{min(timeit(fn), timeit(fn)):.1f}")
    if axis != len(shape) - 1:
        for tx in (8, 32):
            for S in (1, 2, 4, 5, 8, 12, 16, 32):
                env(HPTB_TUNE_NO_BAND=1, HPTB_TUNE_SMC_S=S, HPTB_TUNE_SMC_TX=tx)
                res.append(f"tx{tx}/S{S}:{min(timeit(fn), timeit(fn)):.1f}")
    env()
    nb = 2 * t.numel() * t.element_size()
    print(f"{str(dt)[6:]} {shape} softmax(axis {axis})  ideal {nb / 6552e3:.1f} us   " + "  ".join(res), flush=True)
