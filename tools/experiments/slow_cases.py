#!/usr/bin/env python
"""tools/experiments/slow_cases.py — the layouts tools/layout_survey.py found below 60 %, one call each (for ncu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import hpt_b200 as hb  # noqa: E402
from layout_survey import wrap  # noqa: E402

torch.cuda.set_device(0)
hb.set_stream(torch.cuda.current_stream().cuda_stream)
n = 8192
p = torch.randn(n, n, device="cuda")
P = wrap(p)
for rep in range(2):
    P[:, ::2].exp()
    P[5:8000, 3:8100].exp()
    P[5:8000, 3:8100].sum([1])
    t3 = torch.randn(256, 512, 512, device="cuda")
    T3 = wrap(t3)
    T3.max([0])
    T3.permute([2, 0, 1]).sum([2])
    s = torch.randn(4096, 8192, device="cuda")
    S = wrap(s)
    S.softmax(0)
    S.t().softmax(0)
    lr = torch.randn(256, 131072, device="cuda")
    wrap(lr).softmax(1)
    i8 = torch.randint(-100, 100, (16384, 16384), device="cuda", dtype=torch.int8)
    wrap(i8).t().contiguous()
    col, row = torch.randn(n, 1, device="cuda"), torch.randn(1, n, device="cuda")
    wrap(col) * wrap(row)
    P + wrap(col)
torch.cuda.synchronize()
print("done")
