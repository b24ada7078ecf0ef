#!/usr/bin/env python
"""tools/experiments/snake_order.py — does traversing memory in the opposite direction of the previous kernel pay?

The 126 MB L2 still holds the tail of what the previous kernel streamed; a kernel that starts where that one
ended re-reads those lines from L2 instead of HBM.  Emulated here without kernel changes through reversed views
(negative strides): config 2's max → argmax and sin → exp pairs, forward/forward vs forward/reversed."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import hpt_b200 as hb  # noqa: E402

n = 8192
x = torch.randn((n, n), device="cuda")
stream = torch.cuda.current_stream()
hb.set_stream(stream.cuda_stream)
X = hb.Tensor.from_device_ptr(x.data_ptr(), hb.F32, (n, n), keepalive=x)
V = X.t()
Vr_keep = V[:, ::-1]    # kept dim reversed: outputs (memory rows) visited from the end
Vr_rows = V[::-1, :]
Vr_both = V[::-1, ::-1]


def time_seq(fns, reps=30):
    for f in fns * 3:
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        for f in fns:
            f()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


cases = {
    "max(V); argmax(V)": [lambda: V.max([0]), lambda: V.argmax([0])],
    "max(V); argmax(V[:, ::-1])": [lambda: V.max([0]), lambda: Vr_keep.argmax([0])],
    "max(V[:, ::-1]); argmax(V[:, ::-1])": [lambda: Vr_keep.max([0]), lambda: Vr_keep.argmax([0])],
    "sin(V); exp(V)": [lambda: V.sin(), lambda: V.exp()],
    "sin(V); exp(V[:, ::-1])": [lambda: V.sin(), lambda: Vr_keep.exp()],
    "sin(V); exp(V[::-1, :])": [lambda: V.sin(), lambda: Vr_rows.exp()],
    "sin(V); exp(V[::-1, ::-1])": [lambda: V.sin(), lambda: Vr_both.exp()],
    "step fwd: sin exp max argmax": [lambda: V.sin(), lambda: V.exp(), lambda: V.max([0]), lambda: V.argmax([0])],
    "step snake: sin(V) exp(rev) max(V) argmax(rev)": [lambda: V.sin(), lambda: Vr_both.exp(), lambda: V.max([0]), lambda: Vr_keep.argmax([0])],
}
for name, fns in cases.items():
    print(f"{name:55s} {time_seq(fns):8.1f} us", flush=True)
