"""GPU parity on the edges of the domain: empty and ragged shapes, 8-D views, broadcast (stride-0) and negative
strides as kernel INPUTS, misaligned base pointers, in-place aliasing, 64-bit extents and indices, extreme
integers.  Everything is compared with the oracle, exact unless the op is a float transcendental / sum."""
import numpy as np
import pytest
import torch

from test_binary_gpu import _run as run_binary
from test_reduce_gpu import _check as check_reduce
from test_softmax_misc_gpu import _softmax_check
from test_unary_gpu import _check as check_unary
from util import ENUM, O, assert_exact, rand, to_numpy, to_torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hb():
    import hpt_b200
    return hpt_b200


def test_empty_tensors(hb):
    for shape in [(0,), (0, 5), (5, 0), (3, 0, 7)]:
        E = hb.Tensor.empty(shape, ENUM["f32"])
        assert E.sin().shape == shape and (E * E).shape == shape and E.contiguous().shape == shape
        assert E.astype(ENUM["i32"]).shape == shape
    # reducing an empty axis yields the identity (Hpt: init value; numpy agrees for sum/prod/all/any)
    x = np.zeros((4, 0), np.float32)
    X = hb.Tensor.to_cuda(torch.from_numpy(x))
    assert (X.sum([1]).to_cpu().numpy() == 0).all()
    assert (X.prod([1]).to_cpu().numpy() == 1).all()
    assert (X.max([1]).to_cpu().numpy() == -np.inf).all() and (X.min([1]).to_cpu().numpy() == np.inf).all()
    assert (X.any([1]).to_cpu().numpy() == False).all() and (X.all([1]).to_cpu().numpy() == True).all()  # noqa: E712
    assert (X.argmax(1).to_cpu().numpy() == 0).all()
    assert X.sum([0]).shape == (0,)
    xi = np.zeros((0, 3), np.int32)
    Xi = hb.Tensor.to_cuda(torch.from_numpy(xi))
    assert (Xi.max([0]).to_cpu().numpy() == np.iinfo(np.int32).min).all()
    assert (Xi.min([0]).to_cpu().numpy() == np.iinfo(np.int32).max).all()
    assert X.softmax(1).shape == (4, 0)


def test_eight_dims_and_permutations(hb):
    rng = np.random.default_rng(40)
    shape = (2, 3, 2, 5, 2, 3, 2, 4)
    x, y = rand(rng, shape, "f32"), rand(rng, shape, "i32", -50, 50)
    perm = [7, 0, 6, 1, 5, 2, 4, 3]
    pv = lambda t: t.permute(perm) if hasattr(t, "storage") else np.transpose(t, perm)
    run_binary(hb, "add", x, "f32", y, "i32", pv, pv)
    run_binary(hb, "mul", x, "f32", np.ascontiguousarray(np.transpose(y, perm)), "i32", pv, None)
    check_unary(hb, "exp", x, "f32", view=pv)
    for axes in ([0], [7], [0, 7], [1, 3, 5], [0, 1, 2, 3, 4, 5, 6, 7], [2, 4, 6, 7]):
        check_reduce(hb, "sum", x, "f32", axes, view=pv)
        check_reduce(hb, "max", y, "i32", axes, view=pv)
    check_reduce(hb, "argmin", x, "f32", [4], view=pv)
    _softmax_check(hb, x, "f32", 3, False, pv)


def test_broadcast_and_negative_strides_as_inputs(hb):
    rng = np.random.default_rng(41)
    col = rand(rng, (37, 1), "f32")
    ex = lambda t: t.expand((37, 129)) if hasattr(t, "storage") else np.broadcast_to(t, (37, 129))
    check_unary(hb, "sin", col, "f32", view=ex)
    for axes in ([0], [1], [0, 1]):
        check_reduce(hb, "sum", col, "f32", axes, view=ex)
        check_reduce(hb, "argmax", col, "f32", axes[:1], view=ex)  # every row is one long tie → index 0
    x = rand(rng, (65, 130), "f32")
    flip = lambda t: t[::-1, ::-1]
    check_unary(hb, "exp", x, "f32", view=flip)
    run_binary(hb, "sub", x, "f32", x, "f32", flip, None)
    for op in ("sum", "max", "argmax", "logsumexp"):
        check_reduce(hb, op, x, "f32", [1], view=flip)
        check_reduce(hb, op, x, "f32", [0], view=flip)
    _softmax_check(hb, x, "f32", 1, False, flip)
    _softmax_check(hb, x, "f32", 0, True, flip)


def test_misaligned_base_pointers_and_ragged_tails(hb):
    """Views that start 1–3 elements into an allocation are not 16-byte aligned: the vector kernels must fall back
    (or peel) without touching the neighbours."""
    rng = np.random.default_rng(42)
    for d in ("f32", "bf16", "i8", "f64"):
        base = rand(rng, (5000,), d)
        for off in (1, 2, 3):
            for n in (1, 7, 1023, 4096, 4099):
                v = lambda t, off=off, n=n: t[off:off + n]
                run_binary(hb, "add", base, d, base, d, v, v)
                check_reduce(hb, "sum", base, d, [0], view=v)
                check_reduce(hb, "argmax", base, d, [0], view=v)
        m = rand(rng, (33, 1030), d)
        for v in (lambda t: t[:, 1:1028], lambda t: t[1:, 3:], lambda t: t[:, 2::3]):
            check_reduce(hb, "max", m, d, [1], view=v)
            check_reduce(hb, "sum", m, d, [0], view=v)
            if d in ("f32", "bf16"):
                check_unary(hb, "sqrt", np.abs(m), d, view=v)
                _softmax_check(hb, m, d, 1, False, v)
    # neighbours untouched: write through an out view in the middle of a poisoned buffer
    buf = hb.Tensor.full(7.0, (1000,), ENUM["f32"])
    X = hb.Tensor.to_cuda(torch.arange(100, dtype=torch.float32))
    X.add_(X, buf[3:103])
    got = buf.to_cpu().numpy()
    assert (got[:3] == 7).all() and (got[103:] == 7).all() and (got[3:103] == 2 * np.arange(100)).all()


def test_inplace_aliasing(hb):
    """`out` may alias an input (the reference's `_` ops reuse the lhs storage, binary_normal.rs:546-564)."""
    rng = np.random.default_rng(43)
    x, y = rand(rng, (257, 129), "f32"), rand(rng, (257, 129), "f32")
    X, Y = hb.Tensor.to_cuda(torch.from_numpy(x)), hb.Tensor.to_cuda(torch.from_numpy(y))
    X.add_(Y, X)
    assert (X.to_cpu().numpy() == x + y).all()
    X.mul_(X, X)
    assert (X.to_cpu().numpy() == (x + y) * (x + y)).all()
    Y.exp(out=Y)
    want, _ = O.unary("exp", y, "f32")
    assert O.ulp_diff(Y.to_cpu().numpy(), want, "f32").max() <= 2


def test_extreme_integers_are_exact(hb):
    rng = np.random.default_rng(44)
    for d in ("i64", "u64"):
        info = np.iinfo(O.NP[d])
        x = rng.integers(info.max - 1000, info.max, size=(64, 300), dtype=O.NP[d], endpoint=True)  # all above 2^53
        x[3, 7] = info.max
        x[5, :] = info.min
        for op in ("max", "min", "argmax", "argmin", "sum", "prod"):
            check_reduce(hb, op, x, d, [1])
            check_reduce(hb, op, x, d, [0])
        y = rng.integers(info.min, info.max, size=(64, 300), dtype=O.NP[d], endpoint=True)
        for op in ("add", "sub", "mul", "maximum", "minimum", "rem"):
            run_binary(hb, op, x, d, y, d)
    # float → int casts saturate, NaN → 0 (Rust `as`)
    f = np.array([np.nan, np.inf, -np.inf, 3e38, -3e38, 1e10, -1e10, 255.5, -0.5, 2147483648.0], np.float32)
    F = hb.Tensor.to_cuda(torch.from_numpy(f))
    for d in ("i8", "u8", "i32", "u32", "i64", "u64", "bool"):
        want = O.cast(f, "f32", d)
        assert_exact(to_numpy(F.astype(ENUM[d]).to_cpu(), d), want, d, f"f32 → {d}")


def test_index_beyond_int32(hb):
    """One row of 2^31 + 4099 int8 elements: chunk counters, element offsets and the argmax index need 64 bits
    (the reference's kernels are i32-indexed)."""
    n = 2 ** 31 + 4099
    free, _ = torch.cuda.mem_get_info()
    if free < 8 * 2 ** 30:
        pytest.skip("needs 2 GiB for the tensor")
    x = torch.zeros(n, dtype=torch.int8, device="cuda")
    x[n - 5] = 7
    x[n - 3] = 7      # tie: the lower index wins
    x[12345] = -9
    X = hb.Tensor.from_device_ptr(x.data_ptr(), ENUM["i8"], (n,), keepalive=x)
    assert int(X.argmax(0).to_cpu().numpy()[0]) == n - 5
    assert int(X.argmin(0).to_cpu().numpy()[0]) == 12345
    assert int(X.max([0]).to_cpu().numpy()[0]) == 7
    assert int(X.sum([0]).to_cpu().numpy()[0]) == np.int8((7 + 7 - 9) & 0xFF)
    M = hb.Tensor.from_device_ptr(x.data_ptr(), ENUM["i8"], (2, n // 2), keepalive=x)  # 2-D: rows of 2^30 elements
    got = M.argmax(1).to_cpu().numpy()
    assert got[0] == 0 and got[1] == (n - 5) - n // 2
    Y = X.astype(ENUM["u8"])  # elementwise over > 2^31 elements
    assert int(Y.max([0]).to_cpu().numpy()[0]) == 247  # −9 as u8
    del X, M, Y, x
    torch.cuda.empty_cache()


def _softmax_check_long(hb, x, axis, log, view=None):
    """The ONLINE kernels (rows beyond the register-resident limit of 8192 f32, and every strided axis) keep a running
    (max, Σ) pair: whenever the running max moves, Σ is rescaled by an exp() that is itself good to ≈ 3 ulp, a thread
    sees ≈ log2 of its share of L new maxima and the thread rows / splits merge the same way, so Σ — and with it every
    output — carries ≈ 3·(1 + log2 L) ulp on top of the plain f32 summation error (L/2048 ulp for L/256 sequential
    adds per thread).  BASELINE.json bounds sums by 1e-6·log2(n) relative, which caps the allowance.  Bar: the
    elementwise 4 + |x − max| ulp of _softmax_check plus that allowance; log_softmax: the same as an absolute error
    of ln Σ."""
    X = hb.Tensor.to_cuda(to_torch(x, "f32"))
    if view:
        X, x = view(X), view(x)
    got = (X.log_softmax(axis) if log else X.softmax(axis)).to_cpu().numpy()
    want, od = O.softmax(x, "f32", axis, log)
    L = x.shape[axis]
    extra = min(np.ceil(L / 2048) + 3 * (1 + np.ceil(np.log2(L))), 1e-6 * np.log2(max(L, 2)) / 2.0 ** -23)
    xc = x.astype(np.float64)
    shift = np.abs(xc - xc.max(axis=axis, keepdims=True))
    if log:
        err = np.abs(got.astype(np.float64) - want.astype(np.float64))
        ok = err <= 2.0 ** -23 * ((4 + extra) + 4 * (np.abs(xc).max(axis=axis, keepdims=True) + 1.0))
    else:
        ok = O.ulp_diff(got, want, "f32") <= 4 + np.ceil(shift) + extra
    assert ok.all(), f"softmax log={log} shape={x.shape} axis={axis}: {np.count_nonzero(~ok)} outside the bar"
    if not log:
        assert np.abs(got.astype(np.float64).sum(axis=axis) - 1.0).max() <= 1e-6 * np.log2(L)


def test_softmax_row_lengths_around_the_register_limit(hb):
    rng = np.random.default_rng(45)
    for n in (1, 2, 31, 33, 255, 257, 8191, 8192):
        x = rand(rng, (3, n), "f32") * 4
        _softmax_check(hb, x, "f32", 1, False)
        _softmax_check(hb, x, "f32", 1, True)
    for n in (8193, 8200, 16384, 65536 + 3, 300001):
        x = rand(rng, (3, n), "f32") * 4
        _softmax_check_long(hb, x, 1, False)
        _softmax_check_long(hb, x, 1, True)
    h = rand(rng, (5, 8000), "bf16")
    _softmax_check(hb, h, "bf16", 1, False)
    x = rand(rng, (70000, 3), "f32")  # strided axis: one thread walks a column
    _softmax_check_long(hb, x, 0, False)
    _softmax_check_long(hb, x, 0, True)


def test_stepped_reversed_and_ragged_rows_scalar_path(hb):
    """map_rows_kernel<VEC = 1>: stepped / reversed inner dims and rows off the pack boundary, every element size,
    unary and binary (operands with DIFFERENT inner strides), large enough for several CTAs."""
    rng = np.random.default_rng(46)
    for d in ("f32", "bf16", "f64", "i8", "i64"):
        x = rand(rng, (300, 1030), d, -50, 50) if d in O.INTS else rand(rng, (300, 1030), d)
        y = rand(rng, (300, 1030), d, -50, 50) if d in O.INTS else rand(rng, (300, 1030), d)
        for v in (lambda t: t[:, ::2], lambda t: t[:, ::-1], lambda t: t[:, 1::3], lambda t: t[7:, 5:1028], lambda t: t[::-2, 3:1000:7]):
            run_binary(hb, "add", x, d, y, d, v, v)
            run_binary(hb, "mul", x, d, np.ascontiguousarray(v(y)), d, v, None)   # stepped ⊕ contiguous
            if d in ("f32", "bf16", "f64"):
                check_unary(hb, "exp", x, d, view=v)
    z = rand(rng, (2_000_000,), "f32")  # one stepped dim, > 2^20 elements: the 1-D form of a[:, ::2]
    check_unary(hb, "sin", z, "f32", view=lambda t: t[::2])
    check_unary(hb, "sin", z, "f32", view=lambda t: t[::-1])
    run_binary(hb, "sub", z, "f32", z, "f32", lambda t: t[::2], lambda t: t[1::2])


@pytest.mark.parametrize("op", ["sum", "max", "argmax", "mean", "logsumexp", "prod"])
def test_transposing_reductions_two_step(hb, op):
    """The output's fastest dim is not the input's fastest kept dim (api_reduce.cpp reduces into an input-ordered
    scratch and gathers): permuted 3-D and 4-D views, every op class, i64 exactness."""
    rng = np.random.default_rng(47)
    x = rand(rng, (40, 24, 64), "f32") * (0.05 if op == "prod" else 1.0) + (1.0 if op == "prod" else 0.0)
    p201 = lambda t: t.permute([2, 0, 1]) if hasattr(t, "storage") else np.transpose(t, (2, 0, 1))
    p120 = lambda t: t.permute([1, 2, 0]) if hasattr(t, "storage") else np.transpose(t, (1, 2, 0))
    check_reduce(hb, op, x, "f32", [2], view=p201)   # kept (orig 2, orig 0): out fastest = orig 0, in fastest = orig 2
    check_reduce(hb, op, x, "f32", [0], view=p120)   # reduce orig 1; kept (orig 2, orig 0) again
    if op not in ("argmax",):
        w = rand(rng, (12, 9, 10, 48), "f32") * (0.05 if op == "prod" else 1.0) + (1.0 if op == "prod" else 0.0)
        p3012 = lambda t: t.permute([3, 0, 1, 2]) if hasattr(t, "storage") else np.transpose(t, (3, 0, 1, 2))
        check_reduce(hb, op, w, "f32", [1, 3], view=p3012)
    if op in ("sum", "max", "argmax"):
        xi = rand(rng, (40, 24, 64), "i64", -1000, 1000)
        check_reduce(hb, op, xi, "i64", [2], view=p201)


def test_softmax_strided_axis_tiled_and_two_step(hb):
    """softmax_cols_tiled (one launch and the split-axis two-phase form, packs and scalars, 32 and 8 lanes) and the
    two-step path for an axis that is contiguous in the input but not in the output."""
    rng = np.random.default_rng(48)
    for shape in [(5000, 40), (5000, 37), (300, 4096), (64, 8), (9000, 256), (33, 7, 130)]:
        for log in (False, True):
            _softmax_check_long(hb, rand(rng, shape, "f32") * 3, 0, log)
            _softmax_check(hb, rand(rng, shape, "bf16"), "bf16", 0, log)  # one bf16 ulp dwarfs the Σ allowance
    x3 = rand(rng, (33, 700, 130), "f32")
    _softmax_check_long(hb, x3, 1, False)        # middle axis: outer dims on both sides of the column tiles
    t = rand(rng, (64, 1100), "f32") * 3
    tv = lambda a: a.t() if hasattr(a, "storage") else a.T
    _softmax_check(hb, t, "f32", 0, False, tv)   # axis contiguous in the input, output contiguous the other way:
    _softmax_check(hb, t, "f32", 0, True, tv)    # the register-resident row kernel + a transposing copy
    t3 = rand(rng, (40, 50, 64), "f32")
    pv = lambda a: a.permute([2, 0, 1]) if hasattr(a, "storage") else np.transpose(a, (2, 0, 1))
    _softmax_check(hb, t3, "f32", 0, False, pv)


def test_misaligned_rows_head_body_tail(hb):
    """api_reduce.cpp peels rows that start off the 16-byte boundary into head + aligned body + tail and folds the
    three partial results into `out` (init_out = 0): every foldable op, exact-accumulator dtypes, 2-D and 3-D."""
    rng = np.random.default_rng(49)
    for d in ("f32", "f64", "i8", "i32", "i64", "bool"):
        m = rand(rng, (300, 1030), d, -3, 3) if d in O.INTS else rand(rng, (300, 1030), d)
        for v in (lambda t: t[:, 1:1028], lambda t: t[2:, 3:], lambda t: t[:, 5:1029]):
            for op in ("sum", "max", "min", "sum_square", "reducel1", "nansum", "all", "any"):
                if d == "bool" and op in ("sum_square", "reducel1"):
                    continue
                mm = np.abs(m) if (op == "nansum" and d in ("f32", "f64")) else m  # the checker's nansum bound is relative to |Σ|
                check_reduce(hb, op, mm, d, [1], view=v)
                check_reduce(hb, op, mm, d, [0, 1], view=v)
        c = rand(rng, (12, 40, 517), d, -3, 3) if d in O.INTS else rand(rng, (12, 40, 517), d)
        check_reduce(hb, "sum", c, d, [2], view=lambda t: t[:, :, 3:515])   # row stride 517: NOT pack-aligned → scalar path
        c = rand(rng, (12, 40, 528), d, -3, 3) if d in O.INTS else rand(rng, (12, 40, 528), d)  # 528 = 33·16: every dtype peels
        for op in ("sum", "max", "min", "sum_square", "all", "any"):
            if d == "bool" and op == "sum_square":
                continue
            check_reduce(hb, op, c, d, [2], view=lambda t: t[:, :, 3:515])
            check_reduce(hb, op, c, d, [0, 2], view=lambda t: t[:, 1:, 1:519])
            check_reduce(hb, op, c, d, [0, 1, 2], view=lambda t: t[:, :, 5:])
    x = rand(rng, (300, 1030), "f32") * 0.01 + 1.0
    check_reduce(hb, "prod", x, "f32", [1], view=lambda t: t[:, 1:200])
    xn = np.abs(rand(rng, (300, 1030), "f32"))
    xn[::7, ::5] = np.nan
    check_reduce(hb, "nansum", xn, "f32", [1], view=lambda t: t[:, 3:1029])
    check_reduce(hb, "nanprod", xn * 0.01 + 1.0, "f32", [1], view=lambda t: t[:, 3:300])
