#!/usr/bin/env python
"""one exp() of the unaligned window a[5:8000, 3:8100] (map_ragged_kernel), for ncu"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpt_b200 as hb
stream = torch.cuda.current_stream(); hb.set_stream(stream.cuda_stream)
t = torch.randn((8192, 8192), device="cuda")
X = hb.Tensor.from_device_ptr(t.data_ptr(), hb.F32, (8192, 8192), keepalive=t)
for _ in range(3):
    Y = X[5:8000, 3:8100].exp()
torch.cuda.synchronize()
