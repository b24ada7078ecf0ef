#!/usr/bin/env python
"""register softmax: two rows per CTA (G = 128, 4 packs per thread) against one row per CTA (G = 256, 2 packs)"""
import os, sys
os.environ["HPTB_TUNE"] = "1"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpt_b200 as hb
from ctypes import byref
from hpt_b200 import _ffi
stream = torch.cuda.current_stream(); hb.set_stream(stream.cuda_stream)
def timeit(fns, reps=40):
    for i in range(8): fns[i % len(fns)]()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(reps): fns[i % len(fns)]()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
DT = {torch.float32: hb.F32, torch.bfloat16: hb.BF16, torch.float16: hb.F16}
for shape, dt in [((4096, 4096), torch.bfloat16), ((4096, 4096), torch.float16), ((16384, 2048), torch.float32), ((8192, 1024), torch.float32),
                  ((32768, 4096), torch.bfloat16), ((16384, 1536), torch.float32), ((65536, 768), torch.float32)]:
    fns = []
    for _ in range(4):  # rotate buffer sets: the 32 MB cases would otherwise sit in L2
        t = torch.randn(shape, device="cuda").to(dt)
        X = hb.Tensor.from_device_ptr(t.data_ptr(), DT[dt], shape, keepalive=t)
        Y = hb.Tensor.empty(shape, DT[dt])
        fns.append(lambda X=X, Y=Y: _ffi.check(hb.lib.hptb_softmax(X.ctx.handle, byref(X._c()), 1, 0, byref(Y._c()), hb.get_stream())))
    res = []
    for off in ("1", "0"):
        os.environ["HPTB_TUNE_SM_NO_G128"] = off
        res.append(("G256" if off == "1" else "G128") + f":{min(timeit(fns), timeit(fns)):.1f}")
    os.environ.pop("HPTB_TUNE_SM_NO_G128")
    nb = 2 * t.numel() * t.element_size()
    print(f"{str(dt)[6:]} {shape} softmax(1)  ideal {nb / 6552e3:.1f} us   " + "  ".join(res), flush=True)
