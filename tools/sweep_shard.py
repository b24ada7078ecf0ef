#!/usr/bin/env python
"""Launch-shape sweep for config 5's shards (HPTB_TUNE=1): splits per output S of reduce_rows_kernel (sum()) and row
slabs per column tile S of reduce_cols_lean_kernel (sum(axis 0)) on [262144/N, 16384] f32, N = 1..8."""
import os
import sys

os.environ["HPTB_TUNE"] = "1"
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpt_b200 as hb  # noqa: E402

stream = torch.cuda.current_stream()
hb.set_stream(stream.cuda_stream)
big = torch.randn((262144, 16384), device="cuda", dtype=torch.float32)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


for n in (8, 4, 2, 1):
    rows = 262144 // n
    X = hb.Tensor.from_device_ptr(big.data_ptr(), hb.F32, (rows, 16384), keepalive=big)
    o1, oc = hb.Tensor.empty((1,), hb.F32), hb.Tensor.empty((16384,), hb.F32)
    ideal = rows * 16384 * 4 / 7.4e6
    for name, fn, ss in (("sum()", lambda: X._reduce("sum", [0, 1], out=o1), (0, 2960, 3552, 4096, 4144, 4440, 4736, 5328, 5920, 6144, 7104, 8192)),
                         ("sum(0)", lambda: X._reduce("sum", [0], out=oc), (0, 32, 64, 128, 256))):
        for al in ((0, 8, 256) if name == "sum()" else (0,)):
            res = []
            if al:
                os.environ["HPTB_TUNE_CPS_ALIGN"] = str(al)
            for S in ss:
                if S:
                    os.environ["HPTB_TUNE_S"] = str(S)
                else:
                    os.environ.pop("HPTB_TUNE_S", None)
                res.append(f"S={S or 'auto'}:{min(timeit(fn), timeit(fn)):.1f}")
            os.environ.pop("HPTB_TUNE_S", None)
            os.environ.pop("HPTB_TUNE_CPS_ALIGN", None)
            print(f"N={n} [{rows},16384] {name:7s} align={al} (7.4 TB/s = {ideal:.0f} us)  " + "  ".join(res), flush=True)
