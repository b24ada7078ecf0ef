#!/usr/bin/env python
"""tools/layout_survey.py — achieved HBM fraction over a spread of permuted / broadcast / strided layouts and dtypes
(the ≥ 60 % target of BASELINE.json applies to all of them, not only to the five named configurations).

    python tools/layout_survey.py [--out gpurun_out/layout_survey.txt]

One line per case: µs (CUDA events, rotating buffers where the set fits L2), algorithmic GB/s, fraction of the measured
copy peak, and torch's time for the same op as a yardstick.  Development tool; torch is never on the product path."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import hpt_b200 as hb  # noqa: E402
from sweep import PEAK, timeit  # noqa: E402

TD = {torch.float32: hb.F32, torch.float16: hb.F16, torch.bfloat16: hb.BF16, torch.float64: hb.F64, torch.int8: hb.I8,
      torch.int64: hb.I64, torch.int32: hb.I32}


def wrap(t):
    """hb.Tensor over a torch CUDA tensor (any strides)."""
    return hb.Tensor.from_device_ptr(t.data_ptr(), TD[t.dtype], tuple(t.shape), tuple(t.stride()), keepalive=t)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    hb.set_stream(torch.cuda.current_stream().cuda_stream)
    out = open(a.out, "w") if a.out else None
    rows = []

    def case(name, nbytes, ours, theirs=None, reps=50):
        us = timeit([ours], reps)
        tus = timeit([theirs], max(5, reps // 5)) if theirs else float("nan")
        gbs = nbytes / us / 1e3
        line = f"{name:58s} {us:9.1f} us {gbs:8.0f} GB/s {gbs / PEAK:6.3f}   torch {tus:9.1f} us"
        rows.append((gbs / PEAK, name))
        print(line, flush=True)
        if out:
            out.write(line + "\n")
            out.flush()

    g = torch.Generator(device="cuda").manual_seed(7)
    rn = lambda *s, dtype=torch.float32: torch.randn(*s, device="cuda", generator=g).to(dtype)

    # ---- transposes / permutes of one operand, by element size -----------------------------------------
    for dt, n in ((torch.float32, 8192), (torch.float16, 8192), (torch.bfloat16, 8192), (torch.float64, 4096), (torch.int8, 16384)):
        x = rn(n, n, dtype=dt) if dt != torch.int8 else torch.randint(-100, 100, (n, n), device="cuda", dtype=torch.int8)
        X, y = wrap(x), torch.empty_like(x)
        nb = 2 * x.numel() * x.element_size()
        case(f"contiguous() of {str(dt)[6:]} [{n},{n}].t()", nb, lambda X=X: X.t().contiguous(), lambda x=x, y=y: y.copy_(x.t()))
        if dt.is_floating_point:
            case(f"exp of {str(dt)[6:]} [{n},{n}].t()", nb, lambda X=X: X.t().exp(), lambda x=x, y=y: torch.exp(x.t(), out=y))
    x = rn(1024, 1024, 64)
    X, y = wrap(x), torch.empty(64, 1024, 1024, device="cuda")
    case("contiguous() of f32 [1024,1024,64].permute(2,0,1)", 2 * x.numel() * 4, lambda: X.permute([2, 0, 1]).contiguous(), lambda: y.copy_(x.permute(2, 0, 1)))
    y2 = torch.empty(1024, 64, 1024, device="cuda")
    case("contiguous() of f32 [1024,1024,64].permute(0,2,1)", 2 * x.numel() * 4, lambda: X.permute([0, 2, 1]).contiguous(), lambda: y2.copy_(x.permute(0, 2, 1)))
    y3 = torch.empty(1024, 1024, 64, device="cuda")
    case("contiguous() of f32 [1024,1024,64].permute(1,0,2)", 2 * x.numel() * 4, lambda: X.permute([1, 0, 2]).contiguous(), lambda: y3.copy_(x.permute(1, 0, 2)))
    xb = rn(64, 512, 56, 56, dtype=torch.bfloat16)
    Xb, yb = wrap(xb), torch.empty(64, 56, 56, 512, device="cuda", dtype=torch.bfloat16)
    case("contiguous() of bf16 NCHW [64,512,56,56] → NHWC", 2 * xb.numel() * 2, lambda: Xb.permute([0, 2, 3, 1]).contiguous(), lambda: yb.copy_(xb.permute(0, 2, 3, 1)))
    case("relu of bf16 NCHW [64,512,56,56] → NHWC", 2 * xb.numel() * 2, lambda: Xb.permute([0, 2, 3, 1]).relu(), lambda: torch.relu(xb.permute(0, 2, 3, 1)))

    # ---- binary ops: one transposed operand, broadcasts -----------------------------------------------------
    n = 8192
    p, q = rn(n, n), rn(n, n)
    P, Q = wrap(p), wrap(q)
    o = torch.empty(n, n, device="cuda")
    case("f32 a.t() + b [8192,8192]", 3 * n * n * 4, lambda: P.t() + Q, lambda: torch.add(p.t(), q, out=o))
    case("f32 a.t() * b.t() [8192,8192]", 3 * n * n * 4, lambda: P.t() * Q.t(), lambda: torch.mul(p.t(), q.t(), out=o))
    col, row = rn(n, 1), rn(1, n)
    C, R = wrap(col), wrap(row)
    case("f32 a + column [8192,1]", 2 * n * n * 4, lambda: P + C, lambda: torch.add(p, col, out=o))
    case("f32 a + row [1,8192]", 2 * n * n * 4, lambda: P + R, lambda: torch.add(p, row, out=o))
    case("f32 column [8192,1] * row [1,8192] (outer product)", n * n * 4, lambda: C * R, lambda: torch.mul(col, row, out=o))
    z = rn(64, 1, 4096)
    w = rn(1, 512, 1)
    Z, W = wrap(z), wrap(w)
    o3 = torch.empty(64, 512, 4096, device="cuda")
    case("f32 [64,1,4096] + [1,512,1] → [64,512,4096]", 64 * 512 * 4096 * 4, lambda: Z + W, lambda: torch.add(z, w, out=o3))
    # ---- sliced / stepped views ---------------------------------------------------------------------------------
    case("f32 exp of a[:, ::2] (stride-2 inner: half of every sector)", n * n // 2 * 8, lambda: P[:, ::2].exp(), lambda: torch.exp(p[:, ::2]))
    case("f32 exp of a[::2, :] (every other row)", n * n // 2 * 8, lambda: P[::2, :].exp(), lambda: torch.exp(p[::2, :]))
    case("f32 exp of a[5:8000, 3:8100] (unaligned window)", 7995 * 8097 * 8, lambda: P[5:8000, 3:8100].exp(), lambda: torch.exp(p[5:8000, 3:8100]))
    case("f32 sum(1) of a[5:8000, 3:8100]", 7995 * 8097 * 4, lambda: P[5:8000, 3:8100].sum([1]), lambda: torch.sum(p[5:8000, 3:8100], 1))

    # ---- reductions over non-last axes / several axes -----------------------------------------------------------
    t3 = rn(256, 512, 512)
    T3 = wrap(t3)
    nb3 = t3.numel() * 4
    case("f32 [256,512,512] sum(axis 1)", nb3, lambda: T3.sum([1]), lambda: torch.sum(t3, 1))
    case("f32 [256,512,512] sum(axes 0,2)", nb3, lambda: T3.sum([0, 2]), lambda: torch.sum(t3, (0, 2)))
    case("f32 [256,512,512] max(axis 0)", nb3, lambda: T3.max([0]), lambda: torch.amax(t3, 0))
    case("f32 [256,512,512] argmax(axis 1)", nb3, lambda: T3.argmax(1), lambda: torch.argmax(t3, 1))
    case("f32 [256,512,512] argmax(axis 2)", nb3, lambda: T3.argmax(2), lambda: torch.argmax(t3, 2))
    case("f32 [256,512,512].permute(2,0,1) sum(axis 2)", nb3, lambda: T3.permute([2, 0, 1]).sum([2]), lambda: torch.sum(t3.permute(2, 0, 1), 2))
    case("f32 [8192,8192] logsumexp(axis 0)", n * n * 4, lambda: P.logsumexp([0]), lambda: torch.logsumexp(p, 0))
    case("f32 [8192,8192] mean(axis 0)", n * n * 4, lambda: P.mean([0]), lambda: torch.mean(p, 0))
    i64 = torch.randint(-1000, 1000, (4096, 8192), device="cuda", dtype=torch.int64)
    I64 = wrap(i64)
    case("i64 [4096,8192] sum(axis 1)", i64.numel() * 8, lambda: I64.sum([1]), lambda: torch.sum(i64, 1))
    case("i64 [4096,8192] argmin(axis 0)", i64.numel() * 8, lambda: I64.argmin(0), lambda: torch.argmin(i64, 0))

    # ---- softmax / layernorm off the last axis and on long rows -------------------------------------------------
    s = rn(4096, 8192)
    S = wrap(s)
    so = torch.empty_like(s)
    case("f32 [4096,8192] softmax(axis 1)", 2 * s.numel() * 4, lambda: S.softmax(1), lambda: torch.softmax(s, 1, out=so))
    case("f32 [4096,8192] softmax(axis 0)", 2 * s.numel() * 4, lambda: S.softmax(0), lambda: torch.softmax(s, 0, out=so), reps=10)
    case("f32 [4096,8192].t() softmax(axis 0) (contiguous axis)", 2 * s.numel() * 4, lambda: S.t().softmax(0), lambda: torch.softmax(s.t(), 0), reps=10)
    lr = rn(256, 131072)
    LR = wrap(lr)
    lo = torch.empty_like(lr)
    case("f32 [256,131072] softmax(axis 1) (streamed rows)", 2 * lr.numel() * 4, lambda: LR.softmax(1), lambda: torch.softmax(lr, 1, out=lo), reps=10)
    case("f32 [256,131072] log_softmax(axis 1)", 2 * lr.numel() * 4, lambda: LR.log_softmax(1), lambda: torch.log_softmax(lr, 1), reps=10)
    hbf = rn(32 * 128, 4096, dtype=torch.bfloat16)
    HB = wrap(hbf)
    gam, bet = rn(4096, dtype=torch.bfloat16), rn(4096, dtype=torch.bfloat16)
    G_, B_ = wrap(gam), wrap(bet)
    case("bf16 [4096,4096] layernorm(last) + affine", 2 * hbf.numel() * 2, lambda: HB.layernorm([4096], G_, B_), lambda: torch.nn.functional.layer_norm(hbf, (4096,), gam, bet))
    case("bf16 [4096,4096] softmax(axis 1)", 2 * hbf.numel() * 2, lambda: HB.softmax(1), lambda: torch.softmax(hbf, 1))

    rows.sort()
    print("\nlowest fractions:")
    for f, nme in rows[:10]:
        print(f"  {f:6.3f}  {nme}")
    if out:
        out.write("\nlowest fractions:\n" + "".join(f"  {f:6.3f}  {nme}\n" for f, nme in rows[:10]))


if __name__ == "__main__":
    main()
