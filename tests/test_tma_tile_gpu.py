"""GPU parity for the TMA-staged transposing kernel (hpt_b200/csrc/tma_tile.cuh): a permuted operand of a 2- or 4-byte
type read through a tensor map, `ldmatrix`-transposed, written along the output's unit-stride dim.

Every layout here meets the kernel's preconditions (one staged operand, 16-byte-aligned positive strides, extents that
are multiples of a pack) so the TMA path is the one that runs; edge tiles (extents that are not multiples of the 64 ×
256/512-element CTA tile), batch dims, sliced bases, both operand orders of a binary op and a scalar partner are
covered.  Bars as everywhere: integer results and copies bit-exact, transcendentals ≤ 2 ulp.  Each case is also run
with the tensor-map path switched off (HPTB_TUNE_NO_TMA) and the two results must be bit-identical."""
import os

import numpy as np
import pytest

from util import ENUM, O, assert_exact, assert_ulp, rand, to_numpy, to_torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hb():
    import hpt_b200  # tests/conftest.py sets HPTB_TUNE=1 so that HPTB_TUNE_NO_TMA below is honoured
    return hpt_b200


def _both_paths(fn):
    os.environ.pop("HPTB_TUNE_NO_TMA", None)
    a = fn()
    os.environ["HPTB_TUNE_NO_TMA"] = "1"
    try:
        b = fn()
    finally:
        os.environ.pop("HPTB_TUNE_NO_TMA", None)
    return a, b


SHAPES = [(64, 256), (128, 512), (300, 520), (72, 1032), (1000, 264), (8, 2048), (2048, 8), (1024, 1024)]


@pytest.mark.parametrize("d", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("op", ["sin", "exp"])
def test_transposed_unary(hb, d, op):
    rng = np.random.default_rng(40)
    for shape in SHAPES:
        x = rand(rng, shape, d)
        X = hb.Tensor.to_cuda(to_torch(x, d))
        got, alt = _both_paths(lambda: to_numpy(getattr(X.t(), op)().to_cpu(), d))
        want, od = O.unary(op, x.T, d)
        assert_ulp(got, want, od, 2, f"{op} {d} {shape}.t()")
        assert got.tobytes() == alt.tobytes(), f"{op} {d} {shape}: TMA and shared-memory tile kernels disagree"


@pytest.mark.parametrize("d", ["f32", "i32", "u32", "f16", "bf16", "i16", "u16", "i8", "u8", "bool"])
def test_transposed_copy_is_bit_exact(hb, d):
    rng = np.random.default_rng(41)
    for shape in SHAPES + [(3, 5, 136, 200), (2, 264, 72)]:
        x = rand(rng, shape, d)
        X = hb.Tensor.to_cuda(to_torch(x, d))
        perm = list(range(len(shape)))
        perm[-1], perm[-2] = perm[-2], perm[-1]
        got, alt = _both_paths(lambda: to_numpy(X.permute(perm).contiguous().to_cpu(), d))
        assert_exact(got, np.transpose(x, perm), d, f"contiguous {d} {shape}")
        assert got.tobytes() == alt.tobytes()


def test_batched_and_sliced_views(hb):
    rng = np.random.default_rng(42)
    x = rand(rng, (4, 3, 200, 328), "f32")
    X = hb.Tensor.to_cuda(to_torch(x, "f32"))
    cases = [
        (lambda t: t.permute([0, 1, 3, 2]), lambda a: np.transpose(a, (0, 1, 3, 2))),
        (lambda t: t.permute([1, 0, 3, 2]), lambda a: np.transpose(a, (1, 0, 3, 2))),
        (lambda t: t[:, :, 8:136, 16:272].permute([0, 1, 3, 2]), lambda a: np.transpose(a[:, :, 8:136, 16:272], (0, 1, 3, 2))),
        (lambda t: t[1:3, ::2, :, 4:].permute([0, 1, 3, 2]), lambda a: np.transpose(a[1:3, ::2, :, 4:], (0, 1, 3, 2))),
        (lambda t: t.permute([3, 1, 0, 2]), lambda a: np.transpose(a, (3, 1, 0, 2))),
    ]
    for tv, nv in cases:
        got, alt = _both_paths(lambda: tv(X).exp().to_cpu().numpy())
        want, od = O.unary("exp", nv(x), "f32")
        assert_ulp(got, want, od, 2, "batched / sliced exp")
        assert got.tobytes() == alt.tobytes()


@pytest.mark.parametrize("d", ["f32", "i32", "bf16", "i16", "i8", "u8"])
def test_binary_with_one_permuted_operand(hb, d):
    rng = np.random.default_rng(43)
    for shape in [(256, 512), (300, 520), (5, 136, 72)]:
        perm = list(range(len(shape)))
        perm[-1], perm[-2] = perm[-2], perm[-1]
        tshape = tuple(shape[p] for p in perm)
        a, b = rand(rng, shape, d), rand(rng, tshape, d)
        A, B = hb.Tensor.to_cuda(to_torch(a, d)), hb.Tensor.to_cuda(to_torch(b, d))
        for op in ("add", "sub", "mul"):
            for left in (True, False):  # the permuted operand on either side (sub is not commutative)
                def run():
                    r = A.permute(perm)._binary(op, B) if left else B._binary(op, A.permute(perm))
                    return to_numpy(r.to_cpu(), d)
                got, alt = _both_paths(run)
                at = np.transpose(a, perm)
                want, od = O.binary(op, at, d, b, d) if left else O.binary(op, b, d, at, d)
                assert ENUM[od] == ENUM[d]
                assert_exact(got, want, od, f"{op} {d} {shape} left={left}")
                assert got.tobytes() == alt.tobytes()
    # a scalar partner (stride 0 everywhere)
    a = rand(rng, (264, 136), d)
    s = rand(rng, (1,), d)
    A, S = hb.Tensor.to_cuda(to_torch(a, d)), hb.Tensor.to_cuda(to_torch(s, d))
    out = hb.Tensor.empty((136, 264), ENUM[d])
    got, alt = _both_paths(lambda: to_numpy(A.t().sub_(S, out).to_cpu(), d))
    want, od = O.binary("sub", a.T, d, s, d)
    assert_exact(got, want, od, f"sub scalar {d}")
    assert got.tobytes() == alt.tobytes()


@pytest.mark.parametrize("d", ["i8", "u8"])
def test_one_byte_unary_on_transposed_view(hb, d):
    """1-byte types ride the 2-byte tiles as pairs along b and are pulled apart with byte permutes (tma_tile.cuh)"""
    rng = np.random.default_rng(44)
    for shape in [(256, 512), (264, 1040), (1000, 72), (3, 136, 200)]:
        x = rand(rng, shape, d)
        X = hb.Tensor.to_cuda(to_torch(x, d))
        perm = list(range(len(shape)))
        perm[-1], perm[-2] = perm[-2], perm[-1]
        for op in ("neg", "abs", "square"):
            got, alt = _both_paths(lambda: to_numpy(getattr(X.permute(perm), op)().to_cpu(), d))
            want, od = O.normal_unary(op, np.transpose(x, perm), d)
            assert_exact(got, want, od, f"{op} {d} {shape}")
            assert got.tobytes() == alt.tobytes()
