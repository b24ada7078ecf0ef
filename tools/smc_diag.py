#!/usr/bin/env python
"""where the streaming softmax exceeds the 4 + |x − max| ulp bound (development)."""
import os, sys
os.environ["HPTB_TUNE"] = "1"; os.environ["HPTB_TUNE_NO_BAND"] = "1"
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hpt_b200 as hb
from util import O, rand, to_numpy, to_torch
rng = np.random.default_rng(33)
cases = [((6, 131072), 1), ((3, 200000), 1), ((40, 50000), 1), ((2, 400000), 1), ((4096, 256), 0), ((6144, 96), 0), ((3000, 520), 0), ((7000, 64), 0)]
for shape, axis in cases:
    x = rand(rng, shape, "f32"); x = (x * 4).astype(x.dtype)
    x.flat[7] = 30.0; x.flat[x.size // 2] = -np.inf
    X = hb.Tensor.to_cuda(to_torch(x, "f32"))
    got = to_numpy(X.softmax(axis).to_cpu(), "f32")
    want, od = O.softmax(x, "f32", axis, False)
    u = O.ulp_diff(got, want, od).astype(np.float64)
    shift = np.abs(x.astype(np.float64) - np.max(x, axis=axis, keepdims=True)); shift = np.where(np.isfinite(shift), shift, 0)
    exc = u - (4 + np.ceil(shift))
    idx = np.argsort(exc.ravel())[::-1][:4]
    # Σ error per lane: got/want ratio at the maximum element
    am = np.argmax(x, axis=axis)
    print(shape, "max excess", exc.max(), "max ulp", u.max())
    for i in idx:
        j = np.unravel_index(i, x.shape)
        print("   idx", j, "u", u[j], "shift", shift[j], "got", got[j], "want", want[j], "rel", (float(got[j]) - float(want[j])) / float(want[j]) / 2.0 ** -24)
    lane = np.unravel_index(idx[0], x.shape)[1 - axis] if len(shape) == 2 else 0
    sl = (slice(None), lane) if axis == 0 else (lane, slice(None))
    w64 = np.exp(x[sl].astype(np.float64) - x[sl].max()); w64 /= w64.sum()
    r = (got[sl].astype(np.float64) - w64) / w64 / 2.0 ** -24
    fin = np.isfinite(r) & (w64 > 1e-30)
    print("   lane", lane, "rel err (2^-24 units) at max elem", r[np.argmax(x[sl])], "median", np.median(r[fin]), "min", r[fin].min(), "max", r[fin].max())
