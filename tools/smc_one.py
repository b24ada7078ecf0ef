#!/usr/bin/env python
"""one softmax call per listed case, for ncu: HPTB_TUNE_* knobs come from the environment."""
import os, sys
os.environ["HPTB_TUNE"] = "1"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpt_b200 as hb
from ctypes import byref
from hpt_b200 import _ffi
stream = torch.cuda.current_stream(); hb.set_stream(stream.cuda_stream)
shape = tuple(int(v) for v in sys.argv[1].split("x")); axis = int(sys.argv[2])
dt = {"f32": (torch.float32, hb.F32), "bf16": (torch.bfloat16, hb.BF16)}[sys.argv[3] if len(sys.argv) > 3 else "f32"]
t = torch.randn(shape, device="cuda").to(dt[0])
X = hb.Tensor.from_device_ptr(t.data_ptr(), dt[1], shape, keepalive=t)
Y = hb.Tensor.empty(shape, dt[1])
for _ in range(3):
    _ffi.check(hb.lib.hptb_softmax(X.ctx.handle, byref(X._c()), axis, 0, byref(Y._c()), hb.get_stream()))
torch.cuda.synchronize()
