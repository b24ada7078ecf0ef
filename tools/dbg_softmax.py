import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
os.environ["HPTB_TUNE"] = "1"
import numpy as np, torch
import hpt_b200 as hb
from util import O, rand, to_torch, to_numpy
rng = np.random.default_rng(33)
for shape, axis in (((3000, 520), 0), ((6, 131072), 1)):
    x = (rand(rng, shape, "f32") * 4).astype(np.float32)
    x.flat[7] = 30.0
    x.flat[x.size // 2] = -np.inf
    for off in ("0", "1"):
        os.environ["HPTB_TUNE_NO_BAND"] = off
        X = hb.Tensor.to_cuda(to_torch(x, "f32"))
        got = X.softmax(axis).to_cpu().numpy()
        want, od = O.softmax(x, "f32", axis, False)
        u = O.ulp_diff(got, want, od)
        xc = x.astype(np.float64)
        shift = np.abs(xc - np.max(xc, axis=axis, keepdims=True)); shift = np.where(np.isfinite(shift), shift, 0)
        viol = u - (4 + np.ceil(shift))
        idx = np.unravel_index(np.argmax(viol), viol.shape)
        print(shape, "NO_BAND", off, "max ulp", u.max(), "worst violation", viol.max(), "at", idx, "x", x[idx], "shift", shift[idx], "got", got[idx], "want", want[idx], "n viol", (viol > 0).sum())
        # per-column relative error of the column sum
        s = got.astype(np.float64).sum(axis=axis)
        print("   sum of outputs: min", s.min(), "max", s.max())
