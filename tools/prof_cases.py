#!/usr/bin/env python
"""tools/prof_cases.py — runs each BASELINE configuration's hot-path calls a few times, for ncu captures:

    ncu --set full --clock-control none --import-source on -k regex:'map_|reduce_|softmax_' -o gpurun_out/prof \\
        python tools/prof_cases.py [--reps 2] [--cases cfg1,cfg2,…]

Development tool; torch only generates the synthetic inputs.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import hpt_b200 as hb  # noqa: E402
from sweep import c_binary, c_copy, c_meanvar, c_reduce, c_softmax, c_unary  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--cases", default="")
    a = ap.parse_args()
    want = set(c for c in a.cases.split(",") if c)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    hb.set_stream(torch.cuda.current_stream().cuda_stream)
    T = hb.Tensor
    F32, F64, I64, BF16 = hb.F32, hb.F64, hb.I64, hb.BF16
    g = torch.Generator(device=dev).manual_seed(7)

    def wrap(t, dt):
        return T.from_device_ptr(t.data_ptr(), dt, tuple(t.shape), device=0, keepalive=t)

    def run(fns):
        for _ in range(a.reps):
            for f in fns:
                f()
        torch.cuda.synchronize()

    def on(n):
        return not want or n in want

    if on("cfg1"):
        x = torch.randn((4096, 4096), generator=g, device=dev)
        b = torch.randn((1, 4096), generator=g, device=dev)
        X, B = wrap(x, F32), wrap(b, F32)
        C, S1 = T.empty((4096, 4096), F32, 0), T.empty((4096,), F32, 0)
        run([c_binary("add", X, B, C), c_binary("add", X, X, C), c_reduce("sum", X, [1], S1), c_reduce("sum", X, [0], S1)])
    if on("cfg2"):
        x = torch.randn((8192, 8192), generator=g, device=dev)
        X = wrap(x, F32)
        V = X.t()
        Y, M, I = T.empty((8192, 8192), F32, 0), T.empty((8192,), F32, 0), T.empty((8192,), I64, 0)
        run([c_unary("sin", V, Y), c_unary("exp", V, Y), c_copy(V, Y), c_reduce("max", V, [0], M), c_reduce("argmax", V, [0], I),
             c_reduce("max", X, [0], M), c_reduce("argmax", X, [0], I)])
    if on("cfg3"):
        x = torch.randn((64, 512, 56, 56), generator=g, device=dev).to(torch.bfloat16)
        V = wrap(x, BF16).permute([0, 2, 3, 1])
        o, o2 = T.empty((512,), BF16, 0), T.empty((512,), BF16, 0)
        run([c_reduce("mean", V, [0, 1, 2], o), c_meanvar(V, [0, 1, 2], o, o2)])
    if on("cfg4"):
        x = torch.randn((32, 128, 4096), generator=g, device=dev)
        k = torch.randint(-1000, 1000, (4096,), generator=g, device=dev, dtype=torch.int64)
        X, K = wrap(x, F32), wrap(k, I64)
        Y, L, Z = T.empty((32, 128, 4096), F32, 0), T.empty((32, 128), F32, 0), T.empty((32, 128, 4096), F64, 0)
        run([c_softmax(X, 2, Y), c_reduce("logsumexp", X, [2], L), c_binary("add", X, K, Z)])
    if on("cfg5"):
        x = torch.randn((32768, 16384), generator=g, device=dev)
        X = wrap(x, F32)
        o1, oc = T.empty((1,), F32, 0), T.empty((16384,), F32, 0)
        run([c_reduce("sum", X, [0, 1], o1), c_reduce("sum", X, [0], oc)])


if __name__ == "__main__":
    main()
