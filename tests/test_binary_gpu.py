"""GPU parity: NormalBinOps (+ div, maximum, minimum) through the C ABI vs the oracle.

Layouts follow the reference's house rule (docs/dev_guide/test_rules.md:1-9): contiguous, permuted,
sliced, sliced-with-step, plus broadcasts (hpt-tests/src/hpt/cuda/binary.rs:118-…).  Results are
bit-exact for every dtype pair: both sides cast to the promoted dtype and apply one IEEE/wrapping op."""
import numpy as np
import pytest

from util import DTYPES, ENUM, O, assert_exact, assert_ulp, rand, to_numpy, to_torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hb():
    import hpt_b200
    return hpt_b200


def _run(hb, op, x, xd, y, yd, xview=None, yview=None):
    X = hb.Tensor.to_cuda(to_torch(x, xd))
    Y = hb.Tensor.to_cuda(to_torch(y, yd))
    if xview:
        X, x = xview(X), xview(x)
    if yview:
        Y, y = yview(Y), yview(y)
    want, od = O.binary(op, x, xd, y, yd)
    got = getattr(X, "_binary")(op, Y)
    assert got.dtype == ENUM[od]
    assert_exact(to_numpy(got.to_cpu(), od), want, od, f"{op} {xd},{yd}")


@pytest.mark.parametrize("op", ["add", "sub", "mul", "rem", "div", "maximum", "minimum"])
def test_all_dtype_pairs(hb, op):
    rng = np.random.default_rng(1)
    for xd in DTYPES:
        for yd in DTYPES:
            if O.binary_out_dtype(op, xd, yd) is None:
                X = hb.Tensor.to_cuda(to_torch(rand(rng, (4,), xd), xd))
                Y = hb.Tensor.to_cuda(to_torch(rand(rng, (4,), yd), yd))
                with pytest.raises(hb.HptError) as e:
                    X._binary(op, Y)
                assert e.value.status == 2
                continue
            x = rand(rng, (7, 33), xd)
            y = rand(rng, (7, 33), yd)
            if op in ("rem", "div") and yd in O.INTS:
                y[0, :4] = 0  # division by zero is defined (0 for int rem, inf/nan for float div)
            _run(hb, op, x, xd, y, yd)


@pytest.mark.parametrize("shape", [(1,), (5,), (127,), (4096,), (1025, 3), (3, 1000), (64, 64, 17), (2, 3, 4, 5, 6)])
def test_contiguous_shapes_f32_i64(hb, shape):
    rng = np.random.default_rng(2)
    _run(hb, "add", rand(rng, shape, "f32"), "f32", rand(rng, shape, "f32"), "f32")
    _run(hb, "mul", rand(rng, shape, "i64", -1000, 1000), "i64", rand(rng, shape, "i64", -1000, 1000), "i64")


def test_broadcasts(hb):
    rng = np.random.default_rng(3)
    cases = [((64, 96), (1, 96)), ((64, 96), (64, 1)), ((64, 96), (96,)), ((1, 96), (64, 96)), ((5, 1, 7), (1, 6, 1)),
             ((8, 1, 6, 1), (7, 1, 5)), ((3, 4, 5), (1,)), ((1,), (3, 4, 5)), ((10, 100, 1, 100), (1, 1, 30, 1)),
             ((2, 3, 64), (3, 1))]
    for sa, sb in cases:
        for xd, yd in (("f32", "f32"), ("f32", "i64"), ("bf16", "i8"), ("u8", "i16")):
            _run(hb, "add", rand(rng, sa, xd), xd, rand(rng, sb, yd), yd)
            _run(hb, "sub", rand(rng, sa, xd), xd, rand(rng, sb, yd), yd)


def test_scalar_operand(hb):
    rng = np.random.default_rng(4)
    x = rand(rng, (33, 65), "f32")
    X = hb.Tensor.to_cuda(to_torch(x, "f32"))
    got = X * 3  # tensor ⊕ python int → i64 scalar tensor → f64 (Hpt: f32 ⊕ i64 → f64)
    assert got.dtype == ENUM["f64"]
    assert_exact(to_numpy(got.to_cpu(), "f64"), x.astype(np.float64) * 3.0, "f64")
    got = X + 0.5
    assert_exact(to_numpy(got.to_cpu(), "f64"), x.astype(np.float64) + 0.5, "f64")


@pytest.mark.parametrize("xd,yd", [("f32", "f32"), ("f64", "f32"), ("i32", "u8"), ("f16", "bf16"), ("bool", "i64")])
def test_layouts_permuted_sliced_stepped(hb, xd, yd):
    rng = np.random.default_rng(5)
    x = rand(rng, (18, 70, 66), xd)
    y = rand(rng, (18, 70, 66), yd)
    perm = lambda t: t.permute([2, 1, 0]) if hasattr(t, "permute") else np.transpose(t, (2, 1, 0))
    _run(hb, "add", x, xd, y, yd, perm, perm)
    _run(hb, "mul", x, xd, np.ascontiguousarray(np.transpose(y, (2, 1, 0))), yd, perm, None)  # only lhs permuted
    sl = lambda t: t[2:9, 3:60, 1:65]
    _run(hb, "add", x, xd, y, yd, sl, sl)
    st = lambda t: t[1:17:3, ::2, 5:60:7]
    _run(hb, "sub" if O.binary_out_dtype("sub", xd, yd) else "add", x, xd, y, yd, st, st)
    neg = lambda t: t[::-1, :, ::-2] if isinstance(t, np.ndarray) else t[::-1, :, ::-2]
    _run(hb, "add", x, xd, y, yd, neg, neg)
    # transposed 2-D (config 2's layout) with non-multiple-of-tile extents
    x2, y2 = rand(rng, (130, 257), xd), rand(rng, (257, 130), yd)
    tr = lambda t: t.t() if hasattr(t, "t") else t.T
    _run(hb, "add", x2, xd, y2, yd, tr, None)


def test_out_argument_and_errors(hb):
    rng = np.random.default_rng(6)
    x, y = rand(rng, (16, 16), "f32"), rand(rng, (16, 16), "f32")
    X, Y = hb.Tensor.to_cuda(to_torch(x, "f32")), hb.Tensor.to_cuda(to_torch(y, "f32"))
    out = hb.Tensor.empty((16, 16), ENUM["f32"])
    res = X.add_(Y, out)
    assert res.ptr == out.ptr  # `out` is aliased, not copied (binary_normal.rs:546-564)
    assert_exact(to_numpy(out.to_cpu(), "f32"), x + y, "f32")
    # in place on the lhs storage
    X.add_(Y, X)
    assert_exact(to_numpy(X.to_cpu(), "f32"), x + y, "f32")
    with pytest.raises(hb.HptError) as e:
        X.add_(Y, hb.Tensor.empty((16, 15), ENUM["f32"]))
    assert e.value.status == 1
    with pytest.raises(hb.HptError) as e:
        X.add_(Y, hb.Tensor.empty((16, 16), ENUM["f64"]))
    assert e.value.status == 2
    with pytest.raises(hb.HptError) as e:
        X + hb.Tensor.empty((16, 3), ENUM["f32"])
    assert "Broadcasting error" in str(e.value)
    # empty tensors are a no-op
    E = hb.Tensor.empty((0, 5), ENUM["f32"])
    assert (E + E).shape == (0, 5)


def test_config1_shape_and_config4_promotion(hb):
    rng = np.random.default_rng(7)
    a, b = rand(rng, (1024, 4096), "f32"), rand(rng, (1, 4096), "f32")
    _run(hb, "add", a, "f32", b, "f32")
    x, k = rand(rng, (8, 128, 4096), "f32"), rand(rng, (4096,), "i64", -1000, 1000)
    _run(hb, "add", x, "f32", k, "i64")  # → f64


def test_ragged_kernel_contiguous_output(hb):
    """Dense output, inputs whose rows start off a 16-byte boundary or have ragged / tiny / stepped / reversed inner dims
    (map_ragged_kernel): packs that straddle one or several row ends, totals that are not a multiple of the pack, a
    broadcast partner, several outer dims, every element size."""
    rng = np.random.default_rng(77)
    cases = [((40, 70), lambda t: t[1:39, 3:68]),          # unaligned window, 65-wide rows
             ((33, 9), lambda t: t[:, 1:4]),               # rows of 3: a pack spans two or three rows
             ((50, 7), lambda t: t[:, 2:3]),               # rows of 1
             ((7, 6, 35), lambda t: t[1:6, ::2, 2:33]),    # two outer dims, one stepped
             ((64, 41), lambda t: t[:, ::2]),              # stepped inner dim (21 per row)
             ((64, 41), lambda t: t[:, ::-1]),             # reversed inner dim
             ((1, 1003), lambda t: t[:, 1:1000])]          # one row, ragged total
    for xd, yd in (("f32", "f32"), ("bf16", "bf16"), ("i8", "i8"), ("f64", "f64"), ("i16", "f32")):
        for shape, view in cases:
            x, y = rand(rng, shape, xd), rand(rng, shape, yd)
            _run(hb, "add", x, xd, y, yd, view, view)
        # partner broadcast along the inner dim / along the rows
        x = rand(rng, (40, 70), xd)
        v = lambda t: t[1:39, 3:68]
        _run(hb, "mul", x, xd, rand(rng, (38, 1), yd), yd, v, None)
        _run(hb, "mul", x, xd, rand(rng, (1, 65), yd), yd, v, None)
    # unary through the same path, full-size rows of odd length
    x = rand(rng, (300, 8200), "f32")
    X = hb.Tensor.to_cuda(to_torch(x, "f32"))
    got = X[5:290, 3:8100].exp()
    want, od = O.unary("exp", x[5:290, 3:8100], "f32")
    assert_ulp(to_numpy(got.to_cpu(), od), want, od, 2, "exp of an unaligned window")
