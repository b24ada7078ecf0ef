#!/usr/bin/env python
"""tools/prof_cases_r02.py — the round-2 kernels, one launch each, for ncu captures:
config 5's three reductions on the full [262144,16384] f32 tensor (the bench's dominant kernels), the TMA-staged and
the shared-memory-scatter transposing kernel on the same call (config 2 sin, a.t() + b, bf16 exp), the cluster band
softmax kernels.  Development tool; torch only generates the inputs."""
import os
import sys

os.environ["HPTB_TUNE"] = "1"
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ctypes import byref, c_int32  # noqa: E402

import hpt_b200 as hb  # noqa: E402
from hpt_b200 import _ffi  # noqa: E402

hb.set_stream(torch.cuda.current_stream().cuda_stream)
T = hb.Tensor
TD = {hb.F32: torch.float32, hb.BF16: torch.bfloat16}


def dev(shape, dt=hb.F32):
    t = torch.randn(shape, device="cuda", dtype=torch.float32).to(TD[dt])
    return T.from_device_ptr(t.data_ptr(), dt, tuple(shape), keepalive=t)


def red(X, op, axes, out):
    ax = (c_int32 * len(axes))(*axes)
    _ffi.check(hb.lib.hptb_reduce(X.ctx.handle, _ffi.REDUCE_OPS[op], byref(X._c()), ax, len(axes), byref(out._c()), 1, hb.get_stream()))


def un(V, op, Y):
    _ffi.check(hb.lib.hptb_unary(V.ctx.handle, _ffi.UNARY_OPS[op], byref(V._c()), byref(Y._c()), 0.0, 0.0, hb.get_stream()))


cases = sys.argv[1].split(",") if len(sys.argv) > 1 else ["cfg5", "cfg2", "softmax"]
if "cfg5" in cases:
    X = dev((262144, 16384))
    o1, oc = T.empty((1,), hb.F32), T.empty((16384,), hb.F32)
    for _ in range(2):
        red(X, "sum", [0, 1], o1)
        red(X, "mean", [0, 1], o1)
        red(X, "sum", [0], oc)
    torch.cuda.synchronize()
    del X
if "cfg2" in cases:
    X = dev((8192, 8192))
    B = dev((8192, 8192))
    Y = T.empty((8192, 8192), hb.F32)
    Xh, Yh = dev((8192, 8192), hb.BF16), T.empty((8192, 8192), hb.BF16)
    for no_tma in ("0", "1"):
        os.environ["HPTB_TUNE_NO_TMA"] = no_tma
        for _ in range(2):
            un(X.t(), "sin", Y)
            un(X.t(), "exp", Y)
            _ffi.check(hb.lib.hptb_binary(X.ctx.handle, _ffi.BINARY_OPS["add"], byref(X.t()._c()), byref(B._c()), byref(Y._c()), hb.get_stream()))
            un(Xh.t(), "exp", Yh)
    os.environ.pop("HPTB_TUNE_NO_TMA", None)
    torch.cuda.synchronize()
if "softmax" in cases:
    for shape, axis in (((256, 131072), 1), ((4096, 8192), 0)):
        X = dev(shape)
        Y = T.empty(shape, hb.F32)
        for _ in range(2):
            _ffi.check(hb.lib.hptb_softmax(X.ctx.handle, byref(X._c()), axis, 0, byref(Y._c()), hb.get_stream()))
    torch.cuda.synchronize()
