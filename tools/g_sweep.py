#!/usr/bin/env python
"""threads-per-output (G) sweep of the lean rows kernel for short rows (HPTB_TUNE_G), f32 [256,512,512] over the last axis."""
import os, sys
os.environ["HPTB_TUNE"] = "1"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpt_b200 as hb
stream = torch.cuda.current_stream(); hb.set_stream(stream.cuda_stream)
def timeit(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
for shape in ((256, 512, 512), (65536, 1024), (1 << 20, 128)):
    t = torch.randn(shape, device="cuda")
    X = hb.Tensor.from_device_ptr(t.data_ptr(), hb.F32, shape, keepalive=t)
    ax = [len(shape) - 1]
    oi = hb.Tensor.empty(shape[:-1], hb.I64); of = hb.Tensor.empty(shape[:-1], hb.F32)
    for op, out in (("argmax", oi), ("max", of), ("sum", of)):
        res = []
        for G in (0, 4, 8, 16, 32, 64, 128):
            if G: os.environ["HPTB_TUNE_G"] = str(G)
            else: os.environ.pop("HPTB_TUNE_G", None)
            res.append(f"G={G or 'auto'}:{min(timeit(lambda: X._reduce(op, ax, out=out)), timeit(lambda: X._reduce(op, ax, out=out))):.1f}")
        os.environ.pop("HPTB_TUNE_G", None)
        print(f"f32 {shape} {op}(-1)  ideal {t.numel() * 4 / 6552e3:.1f} us  " + "  ".join(res), flush=True)
