"""GPU parity: NormalReduce / FloatReduce / IndexReduce through the C ABI vs the oracle.

Bars (BASELINE.json north_star): integers, bools and arg indices bit-exact (ties → lowest index, NaN never
wins); f32/f16/bf16/f64 sums and means within 1e-6·log2(n) relative of an f64 accumulation, checked on the
value before the final rounding for f16/bf16 outputs by allowing 1 output ulp; max/min exact.
Shapes/layouts follow hpt-tests/src/hpt/cuda/reduce.rs:155-420 (all axis subsets of 3-D shapes, permuted,
sliced, sliced-with-step)."""
import itertools
import math

import numpy as np
import pytest

from util import DTYPES, ENUM, O, assert_exact, assert_reduce_bar, assert_ulp, rand, to_numpy, to_torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hb():
    import hpt_b200
    return hpt_b200


def _check(hb, op, x, d, axes, keep=False, view=None):
    X = hb.Tensor.to_cuda(to_torch(x, d))
    if view:
        X, x = view(X), view(x)
    got_t = getattr(X, op)(axes, keep)
    want, od, exact = O.reduce(op, x, d, axes, keep)
    assert got_t.dtype == ENUM[od]
    assert tuple(got_t.shape) == tuple(want.shape), (got_t.shape, want.shape)
    got = to_numpy(got_t.to_cpu(), od)
    assert_reduce_bar(op, x, d, axes, got, want, od, exact, f"{op} {d} shape={x.shape} axes={axes}")


ALL_OPS = ["sum", "mean", "max", "min", "argmax", "argmin", "logsumexp", "sum_square", "prod"]


@pytest.mark.parametrize("op", ALL_OPS)
def test_all_dtypes_all_axis_subsets(hb, op):
    rng = np.random.default_rng(20)
    shape = (9, 12, 20)
    for d in DTYPES:
        lo, hi = (-3, 3) if op in ("prod",) else (None, None)
        x = rand(rng, shape, d, lo, hi) if d in O.INTS else rand(rng, shape, d)
        if op == "prod" and d in O.FLOATS:
            x = (np.sign(x) * (0.8 + 0.4 * np.abs(np.tanh(x)))).astype(x.dtype)
            if d == "bf16":
                x = O.round_bf16_from_f32(x)
        if op == "logsumexp" and d in O.INTS:
            x = rand(rng, shape, d, -5, 5)
        subsets = [[a] for a in range(3)] if op in ("argmax", "argmin") else \
            [list(s) for r in (1, 2, 3) for s in itertools.combinations(range(3), r)]
        for axes in subsets:
            _check(hb, op, x, d, axes)
    _check(hb, op, rand(rng, shape, "f32"), "f32", [-1], keep=True)


@pytest.mark.parametrize("op", ["sum", "max", "argmax", "mean"])
def test_layouts(hb, op):
    rng = np.random.default_rng(21)
    x = rand(rng, (40, 33, 50), "f32")
    views = [lambda t: t.permute([1, 0, 2]) if hasattr(t, "storage") else np.transpose(t, (1, 0, 2)),
             lambda t: t.permute([2, 1, 0]) if hasattr(t, "storage") else np.transpose(t, (2, 1, 0)),
             lambda t: t[3:30, 2:20, 5:45], lambda t: t[1:39:3, ::2, 4:50:5], lambda t: t[::-1, :, ::-3]]
    for v in views:
        for axes in ([0], [1], [2]) + (([0, 1], [1, 2], [0, 2], [0, 1, 2]) if op not in ("argmax",) else ()):
            _check(hb, op, x, "f32", list(axes), view=v)


@pytest.mark.parametrize("shape,axes", [
    ((1,), [0]), ((7,), [0]), ((100000,), [0]), ((3, 1), [1]), ((1, 5), [0]), ((513, 1025), [1]), ((513, 1025), [0]),
    ((513, 1025), [0, 1]), ((2, 300000), [1]), ((300000, 2), [0]), ((300000, 2), [1]), ((70000, 5), [1]), ((5, 70000), [0]),
    ((33, 4096), [1]), ((4096, 33), [0]), ((16, 8, 4, 2, 3, 5), [1, 3, 5]), ((16, 8, 4, 2, 3, 5), [0, 2, 4]),
    ((64, 64, 64), [1]), ((1031, 7, 129), [0, 2])])
def test_shapes_sum_max_argmax_f32(hb, shape, axes):
    rng = np.random.default_rng(22)
    x = rand(rng, shape, "f32")
    _check(hb, "sum", x, "f32", axes)
    _check(hb, "max", x, "f32", axes)
    _check(hb, "mean", x, "f32", axes)
    if len(axes) == 1:
        _check(hb, "argmax", x, "f32", axes)
        _check(hb, "argmin", x, "f32", axes)
    xi = rand(rng, shape, "i64", -1000, 1000)
    _check(hb, "sum", xi, "i64", axes)
    _check(hb, "min", xi, "i64", axes)


def test_argmax_ties_and_nans(hb):
    rng = np.random.default_rng(23)
    # every row/column has ties (values in {0,1,2,3}); first index must win at every level of the tree
    for shape in [(257, 4099), (4099, 257), (3, 100001)]:
        x = rng.integers(0, 4, size=shape).astype(np.float32)
        for ax in (0, 1):
            _check(hb, "argmax", x, "f32", [ax])
            _check(hb, "argmin", x, "f32", [ax])
    x = rng.standard_normal((64, 1000)).astype(np.float32)
    x[:, ::3] = np.nan
    x[5, :] = np.nan        # all-NaN row → 0
    x[6, :] = -np.inf       # all-identity row → 0
    x[7, 10] = np.inf
    _check(hb, "argmax", x, "f32", [1])
    _check(hb, "argmin", x, "f32", [1])
    _check(hb, "max", x, "f32", [1])   # NaN-ignoring
    _check(hb, "min", x, "f32", [1])
    xi = np.full((8, 50), np.iinfo(np.int32).min, dtype=np.int32)
    _check(hb, "argmax", xi, "i32", [1])
    xb = rng.integers(0, 2, size=(33, 65)).astype(np.bool_)
    _check(hb, "argmax", xb, "bool", [1])
    _check(hb, "argmin", xb, "bool", [0])


def test_integer_sums_wrap(hb):
    rng = np.random.default_rng(24)
    for d in ("i8", "u8", "i16", "i32", "u64"):
        x = rand(rng, (300, 500), d)
        for axes in ([0], [1], [0, 1]):
            _check(hb, "sum", x, d, axes)
            _check(hb, "sum_square", x, d, axes)


def test_half_precision_accumulates_in_f32(hb):
    rng = np.random.default_rng(25)
    for d in ("bf16", "f16"):
        x = rand(rng, (4, 32, 14, 14), d) + (1.0 if d == "f16" else 0)  # biased so half-precision accumulation would drift
        if d == "bf16":
            x = O.round_bf16_from_f32(x.astype(np.float32) + 1.0)
        nhwc = lambda t: t.permute([0, 2, 3, 1]) if hasattr(t, "storage") else np.transpose(t, (0, 2, 3, 1))
        _check(hb, "mean", x, d, [0, 1, 2], view=nhwc)   # config 3 layout
        _check(hb, "sum", x, d, [0, 1, 2], view=nhwc)
        _check(hb, "sum", x, d, [0, 1, 2, 3])


def test_sum_out_init_out(hb):
    rng = np.random.default_rng(26)
    x = rand(rng, (37, 53), "f32")
    X = hb.Tensor.to_cuda(to_torch(x, "f32"))
    prev = rand(rng, (37,), "f32")
    out = hb.Tensor.to_cuda(to_torch(prev, "f32"))
    r = X.sum_([1], False, True, out)  # init_out = true → out re-initialised → plain sum
    assert r.ptr == out.ptr
    np.testing.assert_allclose(to_numpy(out.to_cpu(), "f32"), x.astype(np.float64).sum(1), rtol=1e-5)
    out = hb.Tensor.to_cuda(to_torch(prev, "f32"))
    X.sum_([1], False, False, out)     # init_out = false → accumulates into the existing contents
    np.testing.assert_allclose(to_numpy(out.to_cpu(), "f32"), x.astype(np.float64).sum(1) + prev, rtol=1e-5, atol=1e-5)


def test_errors(hb):
    X = hb.Tensor.empty((4, 5), ENUM["f32"])
    with pytest.raises(hb.HptError) as e:
        X.sum([2])
    assert e.value.status == 3 and "Dimension out of range" in str(e.value)
    with pytest.raises(hb.HptError) as e:
        X.sum([0, 0])
    assert e.value.status == 3
    with pytest.raises(hb.HptError):
        X.argmax([0, 1])
    with pytest.raises(hb.HptError) as e:
        X.sum_([1], False, True, hb.Tensor.empty((5,), ENUM["f32"]))
    assert e.value.status == 1


def test_determinism_of_split_reduction(hb):
    rng = np.random.default_rng(27)
    x = rand(rng, (3, 2_000_003), "f32")
    X = hb.Tensor.to_cuda(to_torch(x, "f32"))
    a = to_numpy(X.sum([0, 1]).to_cpu(), "f32")
    for _ in range(5):
        assert (to_numpy(X.sum([0, 1]).to_cpu(), "f32") == a).all()
    _check(hb, "sum", x, "f32", [0, 1])
    _check(hb, "sum", x, "f32", [1])
    _check(hb, "sum", x, "f32", [0])


def test_bench_config_shapes_reduced(hb):
    rng = np.random.default_rng(28)
    # config 1 at 1/4 size, config 2 at 1/16, config 5 slice
    c = rand(rng, (1024, 4096), "f32")
    _check(hb, "sum", c, "f32", [1])
    x = rand(rng, (2048, 2048), "f32")
    tr = lambda t: t.t() if hasattr(t, "storage") else t.T
    _check(hb, "max", x, "f32", [0], view=tr)
    _check(hb, "argmax", x, "f32", [0], view=tr)
    y = rand(rng, (4096, 4096), "f32")
    _check(hb, "sum", y, "f32", [0, 1])
    _check(hb, "mean", y, "f32", [0, 1])
    _check(hb, "sum", y, "f32", [0])
    z = rand(rng, (8, 128, 4096), "f32")
    _check(hb, "logsumexp", z, "f32", [-1])


@pytest.mark.gpu
def test_logsumexp_overflow_and_special_values(hb):
    """naive ln Σ exp as the reference (common_reduce.rs:451-480): large inputs overflow to +inf (not NaN),
    -inf contributes 0, NaN poisons its row; covers the lean, multi-run and split launch shapes."""
    rng = np.random.default_rng(31)
    x = rand(rng, (64, 4096), "f32")
    x[0, 5] = 100.0          # exp overflows f32
    x[1, 7] = np.inf
    x[2, :] = -np.inf        # Σ exp = 0 → ln 0 = -inf
    x[3, 9] = np.nan
    x[4, 100] = 88.9         # just above the overflow threshold of a single exp
    x[5, :] = -200.0         # every exp underflows
    _check(hb, "logsumexp", x, "f32", [1])
    y = rand(rng, (6, 16, 520), "f32")
    y[1, 3, 7] = 95.0
    _check(hb, "logsumexp", y, "f32", [0, 2])      # kept middle dim: runs along one more reduced dim
    _check(hb, "logsumexp", y, "f32", [0, 1, 2])   # one output: split over CTAs
    h = rand(rng, (32, 2048), "bf16")
    _check(hb, "logsumexp", h, "bf16", [1])


@pytest.mark.gpu
def test_multi_run_outputs_match(hb):
    """NCHW channel statistics shape class (config 3 at reduced size): every output is many unit-stride runs."""
    rng = np.random.default_rng(32)
    for d in ("bf16", "f16", "f32", "i32"):
        x = rand(rng, (6, 40, 14, 14), d)
        perm = lambda t: t.permute([0, 2, 3, 1]) if hasattr(t, "storage") else np.transpose(t, (0, 2, 3, 1))
        for op in ("mean", "sum", "max", "sum_square"):
            _check(hb, op, x, d, [0, 1, 2], view=perm)
    x = rand(rng, (3, 7, 1000), "f32")  # runs longer than one CTA pass, odd run counts
    _check(hb, "sum", x, "f32", [0, 2])
    _check(hb, "mean", x, "f32", [0, 2])


def test_rows_off_the_pack_boundary_are_peeled(hb):
    """a[5:R, 3:C].op(1): every row starts off the 16-byte boundary by the same amount → head / aligned body / tail.
    Foldable ops on f32 / f64 / integers fold the pieces through `out` (PEEL); mean, logsumexp, reducel2/3 and the half
    types combine them in a scratch of accumulators finished by one more kernel (PEEL_RAW).  Same bars as everywhere."""
    rng = np.random.default_rng(27)
    for d, shape, sl in (("f32", (700, 2100), (slice(5, 690), slice(3, 2090))), ("bf16", (600, 4200), (slice(2, 590), slice(5, 4190))),
                         ("f16", (600, 4200), (slice(0, 600), slice(1, 4199))), ("i16", (500, 3000), (slice(1, 499), slice(3, 2999))),
                         ("f64", (300, 3000), (slice(0, 300), slice(1, 2999)))):
        x = rand(rng, shape, d)
        view = lambda t: t[sl]  # noqa: E731
        for op in ("sum", "mean", "max", "logsumexp", "reducel2", "sum_square", "prod"):
            if d in O.INTS and op in ("logsumexp",):
                continue
            xs = x
            if op == "prod":
                xs = (np.sign(x) * (0.99 + 0.02 * np.abs(np.tanh(x)))).astype(x.dtype) if d in O.FLOATS else rand(rng, shape, d, -2, 2)
                if d == "bf16":
                    xs = O.round_bf16_from_f32(xs)
            _check(hb, op, xs, d, [1], view=view)
    # 3-D: the misaligned dim is the last of two reduced dims
    x = rand(rng, (40, 50, 1030), "bf16")
    _check(hb, "mean", x, "bf16", [1, 2], view=lambda t: t[:, 2:48, 3:1027])
    _check(hb, "sum", x, "bf16", [2], view=lambda t: t[:, :, 1:1025])
