#!/usr/bin/env python
"""Local (single-GPU, no exchange) time of config 5's three reductions on the shard an N-GPU run owns:
[262144/N, 16384] f32 for N = 1, 2, 4, 8 — the per-rank floor the sharded bench is compared with."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpt_b200 as hb  # noqa: E402

stream = torch.cuda.current_stream()
hb.set_stream(stream.cuda_stream)
big = torch.randn((262144, 16384), device="cuda", dtype=torch.float32)
for n in (1, 2, 4, 8):
    rows = 262144 // n
    X = hb.Tensor.from_device_ptr(big.data_ptr(), hb.F32, (rows, 16384), keepalive=big)
    o1, oc = hb.Tensor.empty((1,), hb.F32), hb.Tensor.empty((16384,), hb.F32)
    for name, fn in (("sum()", lambda: X._reduce("sum", [0, 1], out=o1)), ("mean()", lambda: X._reduce("mean", [0, 1], out=o1)),
                     ("sum(0)", lambda: X._reduce("sum", [0], out=oc))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(20):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 20
        print(f"shard of N={n}: [{rows},16384] {name:7s} {us:9.1f} us  {rows * 16384 * 4 / us / 1e3:8.1f} GB/s")
