"""Shared helpers for the parity tests: host containers, random inputs, comparisons."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import hpt_oracle as O  # noqa: E402

DTYPES = O.DTYPES
ENUM = {n: i for i, n in enumerate(DTYPES)}
TORCH = {"bool": torch.bool, "i8": torch.int8, "i16": torch.int16, "i32": torch.int32, "i64": torch.int64,
         "u8": torch.uint8, "u16": torch.uint16, "u32": torch.uint32, "u64": torch.uint64, "f16": torch.float16,
         "bf16": torch.bfloat16, "f32": torch.float32, "f64": torch.float64}


def to_torch(arr, d):
    """oracle array (bf16 carried as f32) → torch CPU tensor of the real dtype."""
    arr = np.ascontiguousarray(arr)
    if d == "bf16":
        return torch.from_numpy(arr.astype(np.float32)).to(torch.bfloat16)
    return torch.from_numpy(arr.astype(O.NP[d]))


def to_numpy(t, d):
    """torch CPU tensor → oracle array."""
    if d == "bf16":
        return t.float().numpy()
    return t.numpy()


def rand(rng, shape, d, lo=None, hi=None):
    """Random oracle array of dtype d.  Floats ~ N(0,1) (representable in d); ints span a useful range."""
    if d == "bool":
        return rng.integers(0, 2, size=shape).astype(np.bool_)
    if d in O.INTS:
        info = np.iinfo(O.NP[d])
        a = info.min if lo is None else max(lo, info.min)
        b = info.max if hi is None else min(hi, info.max)
        return rng.integers(a, b, size=shape, dtype=O.NP[d], endpoint=True)
    x = rng.standard_normal(size=shape)
    if lo is not None:
        x = rng.uniform(lo, hi, size=shape)
    if d == "f64":
        return x
    if d == "f32":
        return x.astype(np.float32)
    if d == "f16":
        return x.astype(np.float16)
    return O.round_bf16_from_f32(x.astype(np.float32))


def assert_exact(got, want, d, what=""):
    g, w = np.asarray(got), np.asarray(want)
    assert g.shape == w.shape, f"{what}: shape {g.shape} != {w.shape}"
    if d in O.FLOATS:
        ok = (g == w) | (np.isnan(g.astype(np.float64)) & np.isnan(w.astype(np.float64)))
    else:
        ok = g == w
    if not ok.all():
        idx = np.argwhere(~ok)[0]
        raise AssertionError(f"{what}: {(~ok).sum()} mismatches of {ok.size}; first at {tuple(idx)}: got {g[tuple(idx)]!r} want {w[tuple(idx)]!r}")


def assert_ulp(got, want, d, max_ulp, what=""):
    g, w = np.asarray(got), np.asarray(want)
    assert g.shape == w.shape, f"{what}: shape {g.shape} != {w.shape}"
    u = O.ulp_diff(g, w, d)
    if (u > max_ulp).any():
        idx = tuple(np.argwhere(u > max_ulp)[0])
        raise AssertionError(f"{what}: max ulp {u.max()} > {max_ulp}; first at {idx}: got {g[idx]!r} want {w[idx]!r}")


def assert_reduce_bar(op, x, d, axes, got, want, od, exact, what=""):
    """The reduction bars of BASELINE.json north_star, shared by the single-GPU and the sharded tests: integers, bools,
    max/min and arg indices bit-exact; sums/means within 1e-6·log2(n) of an f64 accumulation (relative to Σ|x|: a pure
    relative bound on Σx is unattainable under cancellation) or within 1 ulp of the output dtype (f16/bf16 outputs:
    the final rounding alone is 2^-9 / 2^-11); f64 within 1e-13·log2(n)."""
    import math
    if exact:
        assert_exact(got, want, od, what)
        return
    ax = O.process_axes(axes, x.ndim)
    n = max(2, int(np.prod([x.shape[a] for a in ax])))
    ref = O.reduce_f64(op, x, d, axes).reshape(np.asarray(want).shape)
    scale = np.abs(ref)
    if op in ("sum", "mean", "nansum"):
        mag = O.reduce_f64(op, np.abs(O.to_compute(x, d).astype(np.float64)), "f64", axes).reshape(np.asarray(want).shape)
        scale = np.maximum(scale, mag)
    tol = 1e-6 * math.log2(n) * scale
    g64 = np.asarray(got, np.float64)
    err = np.abs(g64 - ref)
    if od == "f64":
        ok = err <= 1e-13 * math.log2(n) * np.maximum(scale, 1e-300)
    else:
        ok = (err <= tol) | (O.ulp_diff(got, want, od) <= 1)
    ok |= np.isnan(ref) & np.isnan(g64)
    ok |= np.isinf(ref) & (g64 == ref)
    assert ok.all(), f"{what}: {np.count_nonzero(~ok)} outside tolerance; max err {np.nanmax(err)} (tol {np.nanmax(tol)})"
