"""Three-way parity against the REFERENCE's own device code: Hpt's CUDA kernels (oracle/_ref/*.cubin, compiled by
oracle/build_ref.sh from the reference sources), the oracle (oracle/hpt_oracle.py) and this library, on the same
inputs — all bit-exact (indices and copies).  Sizes stay below the reference kernels' ~9.7 M-element limit (SURVEY.md
fact 2).  Only argmax / argmin / strided_copy compile with this image's toolchain (build_ref.sh says why the rest does
not); for them this upgrades "pinned to the reference's test oracle" to "pinned to the reference's own output"."""
import numpy as np
import pytest
import torch

import ref_kernels as R
from util import O, to_torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hb():
    import hpt_b200
    return hpt_b200


@pytest.fixture(scope="module")
def mods():
    if not R.available():
        pytest.skip("oracle/_ref/*.cubin not built (run oracle/build_ref.sh where /root/reference exists)")
    return {n: R.RefModule(n) for n in ("argmax", "argmin", "strided_copy")}


@pytest.mark.parametrize("op", ["argmax", "argmin"])
def test_full_arg_reduction_three_ways(hb, mods, op):
    rng = np.random.default_rng(50)
    for n in (1000, 65536, 1 << 20, (1 << 23) - 7):
        for ties in (False, True):
            x = (rng.integers(0, 5, size=n).astype(np.float32) if ties else rng.standard_normal(n).astype(np.float32))
            want, _, _ = O.reduce(op, x, "f32", [0])
            xt = torch.from_numpy(x).cuda()
            ref = R.ref_arg_flat(mods[op], op, xt).numpy()
            ours = getattr(hb.Tensor.to_cuda(torch.from_numpy(x)), op)(0).to_cpu().numpy()
            assert ref.reshape(-1)[0] == want.reshape(-1)[0], f"oracle vs reference kernel: {op} n={n} ties={ties}"
            assert ours.reshape(-1)[0] == ref.reshape(-1)[0], f"library vs reference kernel: {op} n={n} ties={ties}"


@pytest.mark.parametrize("op", ["argmax", "argmin"])
def test_last_axis_arg_reduction_three_ways(hb, mods, op):
    rng = np.random.default_rng(51)
    for shape in ((4096, 2048), (300, 1000), (17, 4099), (2000, 32)):
        for ties in (False, True):
            x = (rng.integers(0, 4, size=shape).astype(np.float32) if ties else rng.standard_normal(shape).astype(np.float32))
            want, _, _ = O.reduce(op, x, "f32", [1])
            ref = R.ref_arg_rows(mods[op], op, torch.from_numpy(x).cuda()).numpy()
            ours = getattr(hb.Tensor.to_cuda(torch.from_numpy(x)), op)(1).to_cpu().numpy()
            np.testing.assert_array_equal(ref, want, err_msg=f"oracle vs reference kernel: {op} {shape} ties={ties}")
            np.testing.assert_array_equal(ours, ref, err_msg=f"library vs reference kernel: {op} {shape} ties={ties}")


def test_strided_copy_three_ways(hb, mods):
    rng = np.random.default_rng(52)
    x = rng.standard_normal((96, 130, 72)).astype(np.float32)
    xt = torch.from_numpy(x).cuda()
    X = hb.Tensor.to_cuda(torch.from_numpy(x))
    cases = [(lambda t: t.permute(2, 0, 1), lambda t: t.permute([2, 0, 1]), lambda a: np.transpose(a, (2, 0, 1))),
             (lambda t: t.permute(1, 0, 2), lambda t: t.permute([1, 0, 2]), lambda a: np.transpose(a, (1, 0, 2))),
             (lambda t: t[3:90:2, 5:, ::3], lambda t: t[3:90:2, 5:, ::3], lambda a: a[3:90:2, 5:, ::3])]
    for tv, hv, nv in cases:
        want = np.ascontiguousarray(nv(x))
        ref = R.ref_strided_copy(mods["strided_copy"], tv(xt)).numpy()
        ours = hv(X).contiguous().to_cpu().numpy()
        np.testing.assert_array_equal(ref, want)
        np.testing.assert_array_equal(ours, ref)
