"""GPU parity: FloatUnaryOps through the C ABI vs the oracle (f64-evaluated formula rounded once).

Tolerance (BASELINE.json north_star): direct transcendentals ≤ 2 ulp of the output dtype.  Composite
activations (sigmoid, gelu, mish, softplus, celu, selu, elu …) evaluate the reference's f32 formula
(hpt-types/src/scalars/_f32.rs:288-323) whose own rounding error against the exact value exceeds 2 ulp
near cancellation points; for those the bound is 2 ulp OR |err| ≤ 4·eps·max(1,|x|) — the reference's tests use
allclose(1e-3) (hpt-tests/src/lib.rs:13-18)."""
import numpy as np
import pytest

from util import DTYPES, ENUM, O, assert_ulp, rand, to_numpy, to_torch

pytestmark = pytest.mark.gpu

DIRECT = ["sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh", "exp", "exp2",
          "exp10", "ln", "log2", "log10", "sqrt", "cbrt", "recip", "erf"]
COMPOSITE = ["sigmoid", "gelu", "selu", "elu", "celu", "mish", "softplus", "softsign", "hard_sigmoid", "hard_swish"]
EPS = {"f16": 2.0 ** -10, "bf16": 2.0 ** -7, "f32": 2.0 ** -23, "f64": 2.0 ** -52}


@pytest.fixture(scope="module")
def hb():
    import hpt_b200
    return hpt_b200


def _domain(rng, op, shape, d):
    if op in ("asin", "acos", "atanh"):
        x = rng.uniform(-0.999, 0.999, size=shape)
    elif op == "acosh":
        x = rng.uniform(1.0, 50.0, size=shape)
    elif op in ("ln", "log2", "log10", "sqrt"):
        x = rng.uniform(1e-3, 100.0, size=shape)
    elif op in ("exp", "exp2", "exp10", "sinh", "cosh"):
        x = rng.uniform(-8.0, 8.0, size=shape)
    else:
        x = rng.standard_normal(size=shape) * 3.0
    if d == "bf16":
        return O.round_bf16_from_f32(x.astype(np.float32))
    return x.astype(O.NP[d])


def _call(hb, X, op):
    if op == "selu":
        return X.selu(), (O.SELU_ALPHA, O.SELU_SCALE)
    if op in ("elu", "celu"):
        return getattr(X, op)(1.3), (1.3, 0.0)
    return getattr(X, op)(), (0.0, 0.0)


def _check(hb, op, x, d, view=None):
    X = hb.Tensor.to_cuda(to_torch(x, d))
    if view:
        X, x = view(X), view(x)
    got_t, (al, be) = _call(hb, X, op)
    want, od = O.unary(op, x, d, al, be)
    assert got_t.dtype == ENUM[od]
    got = to_numpy(got_t.to_cpu(), od)
    if op in DIRECT:
        assert_ulp(got, want, od, 2, f"{op} {d}")
    else:
        u = O.ulp_diff(got, want, od)
        ref64 = O.unary_f64(op, O.to_compute(O.cast(x, d, od), od).astype(np.float64), al, be)
        err = np.abs(np.asarray(got, np.float64) - ref64)
        ok = (u <= 2) | (err <= 4 * EPS[od] * np.maximum(1.0, np.abs(np.asarray(x, np.float64))))
        assert ok.all(), f"{op} {d}: {np.count_nonzero(~ok)} outside tolerance, max ulp {u.max()}"


@pytest.mark.parametrize("op", DIRECT + COMPOSITE)
def test_float_dtypes(hb, op):
    rng = np.random.default_rng(10)
    for d in ("f32", "f64", "f16", "bf16"):
        _check(hb, op, _domain(rng, op, (257, 129), d), d)


@pytest.mark.parametrize("op", ["sin", "exp", "sqrt", "sigmoid", "tanh"])
def test_integer_and_bool_inputs_promote(hb, op):
    rng = np.random.default_rng(11)
    for d in ("bool", "i8", "i16", "i32", "i64", "u8", "u16", "u32", "u64"):
        x = rand(rng, (64, 33), d, 0 if op == "sqrt" else -8, 8)
        _check(hb, op, x, d)


def test_special_values(hb):
    x = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-40, -1e-40, 88.8, -104.0, 1e30], dtype=np.float32)
    for op in ("sin", "exp", "ln", "sqrt", "tanh", "recip", "erf", "atan"):
        _check(hb, op, x, "f32")


@pytest.mark.parametrize("op", ["sin", "exp"])
def test_layouts(hb, op):
    rng = np.random.default_rng(12)
    for d in ("f32", "f16", "f64"):
        x = _domain(rng, op, (130, 258), d)
        tr = lambda t: t.t() if hasattr(t, "storage") else t.T
        _check(hb, op, x, d, tr)  # config 2 layout: transposed view → contiguous logical output
        x3 = _domain(rng, op, (10, 40, 36), d)
        _check(hb, op, x3, d, lambda t: t.permute([1, 0, 2]) if hasattr(t, "storage") else np.transpose(t, (1, 0, 2)))
        _check(hb, op, x3, d, lambda t: t[2:9, 3:30, 1:35])
        _check(hb, op, x3, d, lambda t: t[::3, 1::2, ::5])


def test_unary_out_and_errors(hb):
    rng = np.random.default_rng(13)
    x = _domain(rng, "sin", (40, 40), "f32")
    X = hb.Tensor.to_cuda(to_torch(x, "f32"))
    out = hb.Tensor.empty((40, 40), ENUM["f32"])
    r = X.sin_(out)
    assert r.ptr == out.ptr
    assert_ulp(to_numpy(out.to_cpu(), "f32"), O.unary("sin", x, "f32")[0], "f32", 2)
    with pytest.raises(hb.HptError) as e:
        X.sin_(hb.Tensor.empty((40, 41), ENUM["f32"]))
    assert e.value.status == 1
    with pytest.raises(hb.HptError) as e:
        X.sin_(hb.Tensor.empty((40, 40), ENUM["f16"]))
    assert e.value.status == 2


def test_large_tail_is_written(hb):
    # the reference's strided kernels leave the tail of tensors > grid·block unwritten
    n = 12_000_003
    x = np.linspace(-3, 3, n, dtype=np.float32)
    X = hb.Tensor.to_cuda(to_torch(x, "f32"))
    got = to_numpy(X.exp().to_cpu(), "f32")
    assert_ulp(got[-1000:], O.unary("exp", x[-1000:], "f32")[0], "f32", 2)
    V = X[: 3000 * 4000].reshape((3000, 4000)).t()
    got = to_numpy(V.sin().to_cpu(), "f32")
    want = O.unary("sin", x[: 3000 * 4000].reshape(3000, 4000).T, "f32")[0]
    assert_ulp(got, want, "f32", 2)


@pytest.mark.parametrize("op", ["sin", "cos"])
def test_sin_cos_bit_pattern_sweep(hb, op):
    """ops.cuh evaluates f32 sin/cos with its own Cody–Waite + minimax path up to |x| = 105615 and f64 above:
    sweep every 509th f32 bit pattern of both signs (8.4 M values: every binade, the 105615 switch, subnormals,
    inf, NaN) plus the floats nearest to k·π/2, in the flat, rows and transposing-tile kernels."""
    bits = np.arange(0, 2 ** 32, 509, dtype=np.uint64).astype(np.uint32)
    x = bits.view(np.float32)
    k = np.arange(1, 70000, dtype=np.float64)
    near = (k[:, None] * (np.pi / 2) + np.array([0.0, 1e-4, -1e-4])[None, :]).astype(np.float32).ravel()
    edge = np.array([0.0, -0.0, 105615.0, -105615.0, np.nextafter(np.float32(105615.0), np.float32(np.inf)), 1e30, -3e38,
                     np.inf, -np.inf, np.nan], np.float32)
    x = np.concatenate([x, near, -near, edge])
    x = np.concatenate([x, np.zeros((-x.size) % 4096, np.float32)]).reshape(-1, 4096)
    with np.errstate(invalid="ignore"):
        want, od = O.unary(op, x, "f32")
    X = hb.Tensor.to_cuda(to_torch(x, "f32"))
    for name, got in (("flat", getattr(X, op)()), ("rows", getattr(X[:, 4:4092], op)()), ("tile", getattr(X.t(), op)())):
        g = to_numpy(got.to_cpu(), od)
        w = want if name == "flat" else want[:, 4:4092] if name == "rows" else want.T
        assert_ulp(g, w, od, 2, f"{op} sweep {name}")
        if name == "flat" and op == "sin":  # sin(−0) = −0
            assert np.signbit(g.ravel()[np.flatnonzero(np.signbit(x.ravel()) & (x.ravel() == 0))]).all()
