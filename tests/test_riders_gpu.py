"""GPU parity for the riders on the same launch classes: FloatBinOps pow/hypot, BitWiseOut,
TensorCmp, NormalUaryOps and the remaining reductions (reducel1/2/3, nansum, nanprod, all, any), through the C ABI
against the oracle.  Integer, bool, compare and NormalUaryOps results are bit-exact; pow/hypot within 2 ulp of the
f64-evaluated value; float sums within 1e-6·log2(n) of an f64 accumulation (relative to Σ|terms|)."""
import math

import numpy as np
import pytest

from util import DTYPES, ENUM, O, assert_exact, assert_ulp, rand, to_numpy, to_torch

pytestmark = pytest.mark.gpu
INTB = O.INTS + ("bool",)


@pytest.fixture(scope="module")
def hb():
    import hpt_b200
    return hpt_b200


def _views():
    perm = lambda t: t.permute([1, 0]) if hasattr(t, "storage") else t.T
    sl = lambda t: t[1:6, 2:30]
    step = lambda t: t[::2, 1::3]
    return [None, perm, sl, step]


@pytest.mark.parametrize("op", ["pow", "hypot"])
def test_float_binary_all_pairs(hb, op):
    rng = np.random.default_rng(41)
    for xd in DTYPES:
        for yd in DTYPES:
            od = O.binary_out_dtype(op, xd, yd)
            x = rand(rng, (5, 37), xd, 0, 6) if xd in O.INTS else (np.abs(rand(rng, (5, 37), xd)) if xd != "bool" else rand(rng, (5, 37), xd))
            y = rand(rng, (5, 37), yd, 0, 3) if yd in O.INTS else rand(rng, (5, 37), yd)
            X, Y = hb.Tensor.to_cuda(to_torch(x, xd)), hb.Tensor.to_cuda(to_torch(y, yd))
            want, _ = O.binary(op, x, xd, y, yd)
            got = X._binary(op, Y)
            assert got.dtype == ENUM[od]
            assert_ulp(to_numpy(got.to_cpu(), od), want, od, 2, f"{op} {xd},{yd}")
    # negative bases with integral exponents, zero, inf
    x = np.array([-2.0, -1.5, 0.0, np.inf, 2.0, -0.0, 1.0, np.nan], dtype=np.float32)
    y = np.array([3.0, 2.0, 0.0, -1.0, 0.5, 3.0, np.nan, 0.0], dtype=np.float32)
    want, _ = O.binary(op, x, "f32", y, "f32")
    got = hb.Tensor.to_cuda(to_torch(x, "f32"))._binary(op, hb.Tensor.to_cuda(to_torch(y, "f32")))
    assert_ulp(to_numpy(got.to_cpu(), "f32"), want, "f32", 2, f"{op} special values")
    # permuted operand (transposing tile kernel) and broadcast
    a, b = np.abs(rand(rng, (96, 80), "f32")) + 0.1, rand(rng, (80, 96), "f32")
    A, B = hb.Tensor.to_cuda(to_torch(a, "f32")), hb.Tensor.to_cuda(to_torch(b, "f32"))
    want, _ = O.binary(op, a, "f32", b.T, "f32")
    assert_ulp(to_numpy(A._binary(op, B.t()).to_cpu(), "f32"), want, "f32", 2, f"{op} permuted")


@pytest.mark.parametrize("op", list(O.BIT_OPS))
def test_bit_ops_all_integer_pairs(hb, op):
    rng = np.random.default_rng(42)
    for xd in DTYPES:
        for yd in DTYPES:
            x, y = rand(rng, (6, 35), xd), rand(rng, (6, 35), yd)
            if op in ("shl", "shr") and yd in O.INTS:
                y = rand(rng, (6, 35), yd, -3, 70)  # counts beyond the bit width and negative counts wrap
            X, Y = hb.Tensor.to_cuda(to_torch(x, xd)), hb.Tensor.to_cuda(to_torch(y, yd))
            if O.binary_out_dtype(op, xd, yd) is None:
                with pytest.raises(hb.HptError) as e:
                    X._binary(op, Y)
                assert e.value.status == 2
                continue
            want, od = O.binary(op, x, xd, y, yd)
            got = X._binary(op, Y)
            assert got.dtype == ENUM[od]
            assert_exact(to_numpy(got.to_cpu(), od), want, od, f"{op} {xd},{yd}")
    x = rand(rng, (40, 24), "i32")
    X = hb.Tensor.to_cuda(to_torch(x, "i32"))
    assert_exact(to_numpy((~X).to_cpu(), "i32"), ~x, "i32", "bitnot")
    assert_exact(to_numpy((X & X.t().contiguous().t()).to_cpu(), "i32"), x, "i32", "and self")


@pytest.mark.parametrize("op", list(O.CMP_OPS))
def test_compare_all_pairs_and_layouts(hb, op):
    rng = np.random.default_rng(43)
    name = {"eq": "tensor_eq", "ne": "tensor_neq", "lt": "tensor_lt", "le": "tensor_le", "gt": "tensor_gt", "ge": "tensor_ge"}[op]
    for xd in DTYPES:
        for yd in DTYPES:
            x = rand(rng, (5, 40), xd, -3, 3) if xd in O.INTS else rand(rng, (5, 40), xd)
            y = rand(rng, (5, 40), yd, -3, 3) if yd in O.INTS else rand(rng, (5, 40), yd)
            if xd in O.FLOATS and yd in O.FLOATS:
                y[:, ::3] = O.cast(x[:, ::3], xd, yd)  # plenty of equal pairs
            if xd in O.FLOATS:
                x[0, 1] = np.nan
            X, Y = hb.Tensor.to_cuda(to_torch(x, xd)), hb.Tensor.to_cuda(to_torch(y, yd))
            want, _ = O.compare(op, x, xd, y, yd)
            got = getattr(X, name)(Y)
            assert got.dtype == ENUM["bool"]
            assert_exact(to_numpy(got.to_cpu(), "bool"), want, "bool", f"{op} {xd},{yd}")
    x, y = rand(rng, (12, 33), "f32"), rand(rng, (12, 33), "f32")
    X, Y = hb.Tensor.to_cuda(to_torch(x, "f32")), hb.Tensor.to_cuda(to_torch(y, "f32"))
    for v in _views():
        xv, yv = (v(x), v(y)) if v else (x, y)
        XV, YV = (v(X), v(Y)) if v else (X, Y)
        assert_exact(to_numpy(getattr(XV, name)(YV).to_cpu(), "bool"), O.compare(op, xv, "f32", yv, "f32")[0], "bool", f"{op} view")
    # broadcast against a row and a scalar-shaped tensor
    r = rand(rng, (1, 33), "i64", -2, 2)
    R = hb.Tensor.to_cuda(to_torch(r, "i64"))
    assert_exact(to_numpy(getattr(X, name)(R).to_cpu(), "bool"), O.compare(op, x, "f32", r, "i64")[0], "bool", f"{op} bcast")


def _nu_call(X, op):
    if op == "leaky_relu":
        return X.leaky_relu(0.25), (0.25, 0.0)
    if op == "clamp":
        return X.clamp(-1.0, 2.0), (-1.0, 2.0)
    if op == "bitnot":
        return ~X, (0.0, 0.0)
    return getattr(X, op)(), (0.0, 0.0)


@pytest.mark.parametrize("op", list(O.NORMAL_UNARY_OPS))
def test_normal_unary_all_dtypes_and_layouts(hb, op):
    rng = np.random.default_rng(44)
    for d in DTYPES:
        x = rand(rng, (9, 64), d, -100, 100) if d in O.INTS else rand(rng, (9, 64), d)
        if d in O.FLOATS:
            x = (x * 3).astype(x.dtype) if d != "bf16" else O.round_bf16_from_f32((x * 3).astype(np.float32))
            x[0, :8] = np.array([0.5, -0.5, 1.5, 2.5, -2.5, np.nan, -0.0, np.inf], dtype=x.dtype)
        if d in ("i8", "i16", "i32", "i64"):
            x[0, 0], x[0, 1] = np.iinfo(O.NP[d]).min, np.iinfo(O.NP[d]).max
        X = hb.Tensor.to_cuda(to_torch(x, d))
        if op == "bitnot" and d in O.FLOATS:
            with pytest.raises(hb.HptError) as e:
                ~X
            assert e.value.status == 2
            continue
        for v in _views():
            xv = v(x) if v else x
            XV = v(X) if v else X
            if op == "clamp" and d == "bool":
                got_t, (al, be) = XV.clamp(0.0, 1.0), (0.0, 1.0)
            elif op == "clamp" and d[0] == "u":
                got_t, (al, be) = XV.clamp(1.0, 50.0), (1.0, 50.0)
            else:
                got_t, (al, be) = _nu_call(XV, op)
            want, od = O.normal_unary(op, xv, d, al, be)
            assert got_t.dtype == ENUM[od] and tuple(got_t.shape) == xv.shape
            assert_exact(to_numpy(got_t.to_cpu(), od), want, od, f"{op} {d}")
    with pytest.raises(hb.HptError):
        hb.Tensor.to_cuda(to_torch(rand(rng, (4,), "f32"), "f32")).clamp(2.0, 1.0)


NEW_RED = ["reducel1", "nansum", "nanprod", "all", "any", "reducel2", "reducel3"]


def _red_check(hb, op, x, d, axes, view=None):
    X = hb.Tensor.to_cuda(to_torch(x, d))
    if view:
        X, x = view(X), view(x)
    got_t = getattr(X, op)(axes)
    want, od, exact = O.reduce(op, x, d, axes)
    assert got_t.dtype == ENUM[od] and tuple(got_t.shape) == tuple(want.shape)
    got = to_numpy(got_t.to_cpu(), od)
    what = f"{op} {d} {x.shape} axes={axes}"
    if exact or od == "bool" or (d in INTB and od == d):
        assert_exact(got, want, od, what)
        return
    ax = O.process_axes(axes, x.ndim)
    n = max(2, int(np.prod([x.shape[a] for a in ax])))
    ref = O.reduce_f64(op, x, d, axes).reshape(want.shape)
    scale = np.abs(ref)
    if op == "nansum":  # a cancelling sum: the bound is relative to Σ|x| (as for sum in test_reduce_gpu.py)
        v = np.nan_to_num(O.to_compute(x, d).astype(np.float64), nan=0.0)
        scale = np.maximum(scale, np.abs(v).sum(axis=tuple(ax)).reshape(want.shape))
    tol = (1e-13 if od == "f64" else 1e-6) * math.log2(n) * np.maximum(scale, 1e-30)
    err = np.abs(np.asarray(got, np.float64) - ref)
    ok = (err <= tol) | (O.ulp_diff(got, want, od) <= 2)
    ok |= np.isnan(ref) & np.isnan(np.asarray(got, np.float64))
    ok |= np.isinf(ref) & (np.asarray(got, np.float64) == ref)
    assert ok.all(), f"{what}: {np.count_nonzero(~ok)} outside tolerance; max err {err.max()}"


@pytest.mark.parametrize("op", NEW_RED)
def test_new_reductions_all_dtypes(hb, op):
    rng = np.random.default_rng(45)
    for d in DTYPES:
        if op in ("nanprod",) and d in O.FLOATS:
            x = rand(rng, (6, 5, 4), d, 0.5, 1.5)
            if d == "bf16":
                x = O.round_bf16_from_f32(np.asarray(x, np.float32))
            else:
                x = x.astype(O.NP[d])
        elif d in O.INTS:
            x = rand(rng, (6, 5, 4), d, -4, 4)
        else:
            x = rand(rng, (6, 5, 4), d)
        if d in O.FLOATS and op in ("nansum", "nanprod", "all", "any"):
            x[1, 2, 3] = np.nan
        if op in ("all", "any"):
            x[0] = 0
            x[1] = 1 if d != "bool" else True
        for axes in ([0], [1], [2], [0, 2], [0, 1, 2]):
            _red_check(hb, op, x, d, axes)


@pytest.mark.parametrize("op", NEW_RED)
def test_new_reductions_layouts_and_sizes(hb, op):
    rng = np.random.default_rng(46)
    x = rand(rng, (48, 40, 36), "f32", 0.9, 1.1).astype(np.float32) if op == "nanprod" else rand(rng, (48, 40, 36), "f32")
    if op in ("nansum", "nanprod"):
        x[3, 4, 5] = np.nan
    perm = lambda t: t.permute([2, 0, 1]) if hasattr(t, "storage") else np.transpose(t, (2, 0, 1))
    sl = lambda t: t[2:40:3, 1:33, ::2]
    for v in (None, perm, sl):
        for axes in ([0], [2], [0, 1], [0, 1, 2]):
            _red_check(hb, op, x, "f32", axes, view=v)
    big = rand(rng, (64, 4096), "f32", 0.99, 1.01).astype(np.float32) if op == "nanprod" else rand(rng, (64, 4096), "f32")
    _red_check(hb, op, big, "f32", [1])
    _red_check(hb, op, big, "f32", [0, 1])


def test_fused_binary_reduce_equals_the_two_calls(hb):
    """hptb_binary_reduce (extension, §8f rank 4): one pass, same semantics as binary followed by reduce — integer
    results bit-exact, float sums within the sum tolerance, max/min exact; fast path (same dtype, unit-stride reduced
    axis, row-broadcast / same-shape / scalar rhs) and the composed fallback (mixed dtypes, other axes, views)."""
    rng = np.random.default_rng(47)

    def check(x, xd, y, yd, bop, rop, axes, xv=None):
        X, Y = hb.Tensor.to_cuda(to_torch(x, xd)), hb.Tensor.to_cuda(to_torch(y, yd))
        if xv:
            X, x = xv(X), xv(x)
        mid, md = O.binary(bop, x, xd, y, yd)
        want, od, exact = O.reduce(rop, mid, md, axes)
        got_t = X.binary_reduce(bop, Y, rop, axes)
        assert got_t.dtype == ENUM[od] and tuple(got_t.shape) == tuple(want.shape)
        got = to_numpy(got_t.to_cpu(), od)
        what = f"{bop}->{rop} {xd},{yd} {x.shape} axes={axes}"
        two = to_numpy(getattr(X._binary(bop, Y), rop)(axes).to_cpu(), od)  # the unfused pair on the device
        if exact or od in INTB:
            assert_exact(got, want, od, what)
            assert_exact(got, two, od, what + " vs unfused")
            return
        n = max(2, int(np.prod([mid.shape[a] for a in O.process_axes(axes, mid.ndim)])))
        ref = O.reduce_f64(rop, mid, md, axes).reshape(want.shape)
        mag = O.reduce_f64("sum", np.abs(O.to_compute(mid, md).astype(np.float64)) ** (2 if rop == "sum_square" else 1), "f64", axes).reshape(want.shape)
        tol = 1e-6 * math.log2(n) * np.maximum(np.abs(ref), mag)
        ok = (np.abs(np.asarray(got, np.float64) - ref) <= tol) | (O.ulp_diff(got, want, od) <= 1)
        assert ok.all(), what

    a, b = rand(rng, (96, 4096), "f32"), rand(rng, (1, 4096), "f32")
    for bop in ("add", "sub", "mul"):
        for rop in ("sum", "max", "min", "sum_square"):
            check(a, "f32", b, "f32", bop, rop, [1])                       # config 1 shape class: row broadcast
    check(a, "f32", rand(rng, (96, 4096), "f32"), "f32", "add", "sum", [1])  # same shape
    check(a, "f32", rand(rng, (1,), "f32"), "f32", "mul", "max", [1])        # scalar rhs
    check(a, "f32", rand(rng, (96, 1), "f32"), "f32", "add", "sum", [1])     # column broadcast: one rhs element per row
    for d in ("i32", "i64", "bf16", "f16", "f64", "u8"):
        x = rand(rng, (33, 512), d, -50, 50) if d in O.INTS else rand(rng, (33, 512), d)
        y = rand(rng, (1, 512), d, -50, 50) if d in O.INTS else rand(rng, (1, 512), d)
        check(x, d, y, d, "add", "sum", [1])
        check(x, d, y, d, "mul", "max", [1])
    # fallbacks: mixed dtypes, reduce over the outer axis, all axes, a sliced view, mean
    check(a, "f32", rand(rng, (4096,), "i64", -5, 5), "i64", "add", "sum", [1])
    check(a, "f32", b, "f32", "add", "sum", [0])
    check(a, "f32", b, "f32", "add", "sum", [0, 1])
    check(a, "f32", rand(rng, (1, 1365), "f32"), "f32", "add", "sum", [1], xv=lambda t: t[::2, 1::3])
    check(a, "f32", b, "f32", "add", "mean", [1])
