"""CPU-only: pin the oracle (oracle/hpt_oracle.py, oracle/oracle_cpu.cpp) before it is trusted.

Sources of truth, in order: (1) the reference's own promotion tables (golden JSON extracted from its Rust
files); (2) the known-answer style checks of hpt-tests/src/hpt_types/tests.rs (f32↔f16/bf16 conversions incl.
±0, ±inf, NaN, MAX/MIN at :823-877; integer add/sub/mul == wrapping_* at :880-1078; bool add = OR); (3) the
reference's test oracle, libtorch (hpt-tests/src/hpt/cpu/{binary,unary,reduce,softmax}.rs compare against
`tch`), reproduced here with torch CPU on the same shapes / layout transforms."""
import ctypes
import itertools
import os

import numpy as np
import pytest
import torch

from util import DTYPES, O, ROOT, TORCH, rand, to_numpy, to_torch


def test_f32_to_half_conversions_match_half_crate_semantics():
    # hpt_types/tests.rs:823-845: 1000 random values over the whole range + inf, -inf, ±0, MAX, MIN, NaN
    rng = np.random.default_rng(0)
    vals = np.concatenate([
        (rng.standard_normal(1000) * 10.0 ** rng.uniform(-8, 8, 1000)).astype(np.float32),
        np.array([np.inf, -np.inf, -0.0, 0.0, np.finfo(np.float32).max, np.finfo(np.float32).min, np.nan, 65504.0, 65520.0,
                  1.0 + 2.0 ** -8, 1.0 + 2.0 ** -9, 1.0 + 3 * 2.0 ** -9, 3.3895313892515355e38], dtype=np.float32)])
    got_bf = O.cast(vals, "f32", "bf16")
    want_bf = torch.from_numpy(vals).to(torch.bfloat16).float().numpy()   # torch uses the same RNE conversion as `half`
    assert ((got_bf == want_bf) | (np.isnan(got_bf) & np.isnan(want_bf))).all()
    assert np.signbit(got_bf[1002]) and got_bf[1002] == 0  # -0 survives
    got_h = O.cast(vals, "f32", "f16")
    want_h = torch.from_numpy(vals).to(torch.float16).numpy()
    assert ((got_h == want_h) | (np.isnan(got_h) & np.isnan(want_h))).all()
    # and back (tests.rs:846-877): exact widening
    assert (O.cast(got_h[~np.isnan(got_h)], "f16", "f32") == got_h[~np.isnan(got_h)].astype(np.float32)).all()


def test_f64_to_bf16_single_rounding():
    # 1 + 2^-8 + 2^-30 is above the bf16 midpoint 1 + 2^-8: one rounding goes up, double rounding (via f32) ties to even (down)
    x = np.array([1.0 + 2.0 ** -8 + 2.0 ** -30], dtype=np.float64)
    assert O.round_bf16_from_f64(x)[0] == np.float32(1.0 + 2.0 ** -7)
    assert O.round_bf16_from_f32(x.astype(np.float32))[0] == np.float32(1.0)
    # i64 → bf16 goes through f64 (scalar_convert.rs: I64 → from_f64): 2^24·(1+2^-8)+1 must round up
    v = np.array([(1 << 24) + (1 << 16) + 1], dtype=np.int64)
    assert O.cast(v, "i64", "bf16")[0] == np.float32((1 << 24) + (1 << 17))
    # i32 → bf16 goes through f32 (double rounding is the reference's behaviour): the +1 is lost first
    assert O.cast(v.astype(np.int32), "i32", "bf16")[0] == np.float32(1 << 24)


def test_rust_as_casts_known_answers():
    f = np.array([np.nan, np.inf, -np.inf, 300.7, -300.7, -0.9, 2147483648.0, -2147483649.0, 1e20, 255.5], dtype=np.float64)
    assert O.cast(f, "f64", "i8").tolist() == [0, 127, -128, 127, -128, 0, 127, -128, 127, 127]
    assert O.cast(f, "f64", "u8").tolist() == [0, 255, 0, 255, 0, 0, 255, 0, 255, 255]
    assert O.cast(f, "f64", "i32").tolist() == [0, 2147483647, -2147483648, 300, -300, 0, 2147483647, -2147483648, 2147483647, 255]
    assert O.cast(f, "f64", "u64").tolist()[:3] == [0, 18446744073709551615, 0]
    assert O.cast(f, "f64", "i64").tolist()[8] == 9223372036854775807
    i = np.array([-1, 256, 65535, -32769, 2 ** 31], dtype=np.int64)
    assert O.cast(i, "i64", "u8").tolist() == [255, 0, 255, 255, 0]
    assert O.cast(i, "i64", "i16").tolist() == [-1, 256, -1, 32767, 0]
    assert O.cast(i, "i64", "u32").tolist() == [4294967295, 256, 65535, 4294934527, 2147483648]
    assert O.cast(np.array([0.0, -0.0, np.nan, 1e-30]), "f64", "bool").tolist() == [False, False, True, True]
    assert O.cast(np.array([True, False]), "bool", "f16").tolist() == [1.0, 0.0]


def test_integer_ops_are_wrapping_like_rust():
    # tests.rs:880-1078 pin __add/__sub/__mul to wrapping_add/sub/mul over the whole range
    rng = np.random.default_rng(1)
    for d in O.INTS:
        a, b = rand(rng, (1000,), d), rand(rng, (1000,), d)
        bits = np.dtype(O.NP[d]).itemsize * 8
        mod = 1 << bits
        signed = d.startswith("i")

        def wrap(v):
            v = [int(t) % mod for t in v]
            return [t - mod if signed and t >= mod // 2 else t for t in v]
        ai, bi = [int(t) for t in a], [int(t) for t in b]
        assert O.binary("add", a, d, b, d)[0].tolist() == wrap([x + y for x, y in zip(ai, bi)])
        assert O.binary("sub", a, d, b, d)[0].tolist() == wrap([x - y for x, y in zip(ai, bi)])
        assert O.binary("mul", a, d, b, d)[0].tolist() == wrap([x * y for x, y in zip(ai, bi)])
        b[b == 0] = 1
        rem = [abs(x) % abs(y) * (1 if x >= 0 else -1) for x, y in zip(ai, [int(t) for t in b])]  # truncated remainder
        assert O.binary("rem", a, d, b, d)[0].tolist() == wrap(rem)
    assert O.binary("add", np.array([True, False]), "bool", np.array([False, False]), "bool")[0].tolist() == [True, False]
    assert O.binary("mul", np.array([True, True]), "bool", np.array([False, True]), "bool")[0].tolist() == [False, True]
    m = np.array([np.iinfo(np.int32).min], np.int32)
    assert O.binary("rem", m, "i32", np.array([-1], np.int32), "i32")[0].tolist() == [0]  # wrapping_rem(MIN, -1) = 0


def test_promotion_rules_in_binary():
    rng = np.random.default_rng(2)
    x, k = rand(rng, (4, 8), "f32"), rand(rng, (8,), "i64", -1000, 1000)
    r, od = O.binary("add", x, "f32", k, "i64")
    assert od == "f64" and r.dtype == np.float64 and (r == x.astype(np.float64) + k.astype(np.float64)).all()
    r, od = O.binary("add", rand(rng, (8,), "i32"), "i32", rand(rng, (8,), "f16"), "f16")
    assert od == "f16"
    r, od = O.binary("add", rand(rng, (8,), "f16"), "f16", rand(rng, (8,), "i32"), "i32")
    assert od == "f32"
    assert O.binary_out_dtype("div", "i32", "i32") == "f32" and O.binary_out_dtype("sub", "bool", "bool") is None


def test_same_dtype_binary_matches_torch_on_reference_layouts():
    # hpt-tests/src/hpt/cpu/binary.rs: f32 randn / i64 arange; same shape, broadcast dim0 / dim1 = 1, permute([2,1,0]), slices with step
    rng = np.random.default_rng(3)
    for d in ("f32", "f64", "i64", "i32", "f16", "bf16"):
        a = rand(rng, (13, 10, 8), d, -1000, 1000) if d in O.INTS else rand(rng, (13, 10, 8), d)
        for bshape in [(13, 10, 8), (1, 10, 8), (13, 1, 8), (8,), (1,)]:
            b = rand(rng, bshape, d, 1, 1000) if d in O.INTS else rand(rng, bshape, d)
            for op, tf in (("add", torch.add), ("sub", torch.sub), ("mul", torch.mul)):
                got, od = O.binary(op, a, d, b, d)
                want = to_numpy(tf(to_torch(a, d), to_torch(b, d)), d)
                assert od == d and (got == want).all(), (d, bshape, op)
        pa = np.transpose(a, (2, 1, 0))
        got, _ = O.binary("mul", pa, d, pa, d)
        assert (got == to_numpy(to_torch(a, d).permute(2, 1, 0) * to_torch(a, d).permute(2, 1, 0), d)).all()
        sa = a[1:12:3, ::2, 2:7]
        got, _ = O.binary("add", sa, d, sa, d)
        assert (got == to_numpy(to_torch(a, d)[1:12:3, ::2, 2:7] * 2, d)).all()


def test_unary_matches_torch_within_reference_tolerance():
    # hpt-tests/src/hpt/cpu/unary.rs:101-176: randn inputs, allclose(1e-3) against tch; here ≤ 1 ulp of torch's f64 result
    rng = np.random.default_rng(4)
    x = rng.standard_normal(2000)
    tmap = {"sin": torch.sin, "cos": torch.cos, "tan": torch.tan, "atan": torch.atan, "sinh": torch.sinh, "cosh": torch.cosh,
            "tanh": torch.tanh, "asinh": torch.asinh, "exp": torch.exp, "exp2": torch.exp2, "erf": torch.erf,
            "sigmoid": torch.sigmoid, "softplus": torch.nn.functional.softplus, "softsign": torch.nn.functional.softsign,
            "mish": torch.nn.functional.mish, "gelu": torch.nn.functional.gelu, "hard_sigmoid": torch.nn.functional.hardsigmoid,
            "hard_swish": torch.nn.functional.hardswish, "selu": torch.selu, "recip": torch.reciprocal}
    for op, tf in tmap.items():
        al, be = (O.SELU_ALPHA, O.SELU_SCALE) if op == "selu" else (0.0, 0.0)
        got = O.unary_f64(op, x, al, be)
        want = tf(torch.from_numpy(x)).numpy()
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14, err_msg=op)
    xp = np.abs(x) + 0.1
    for op, tf in {"ln": torch.log, "log2": torch.log2, "log10": torch.log10, "sqrt": torch.sqrt}.items():
        np.testing.assert_allclose(O.unary_f64(op, xp), tf(torch.from_numpy(xp)).numpy(), rtol=1e-13, err_msg=op)
    np.testing.assert_allclose(O.unary_f64("elu", x, 1.3), torch.nn.functional.elu(torch.from_numpy(x), 1.3).numpy(), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(O.unary_f64("celu", x, 1.0), torch.celu(torch.from_numpy(x), 1.0).numpy(), rtol=1e-12, atol=1e-14)
    # dtype promotion of the output (FloatOutUnaryPromote)
    assert O.unary("sin", np.arange(4, dtype=np.int64), "i64")[1] == "f64"
    assert O.unary("sin", np.arange(4, dtype=np.int8), "i8")[1] == "f16"
    assert O.unary("sin", np.arange(4, dtype=np.uint32), "u32")[1] == "f32"


def test_reductions_match_torch_all_axis_subsets():
    # hpt-tests/src/hpt/cpu/reduce.rs:62-130: every axis subset of random ≤3-D shapes, contiguous and permuted
    rng = np.random.default_rng(5)
    x = rand(rng, (7, 9, 11), "f64")
    t = torch.from_numpy(x)
    for r in (1, 2, 3):
        for axes in itertools.combinations(range(3), r):
            for keep in (False, True):
                got = O.reduce("sum", x, "f64", list(axes), keep)[0]
                want = t.sum(dim=axes, keepdim=keep).numpy().reshape(got.shape)
                np.testing.assert_allclose(got, want, rtol=1e-12)
                np.testing.assert_allclose(O.reduce("mean", x, "f64", list(axes), keep)[0],
                                           t.mean(dim=axes, keepdim=keep).numpy().reshape(got.shape), rtol=1e-12)
                np.testing.assert_allclose(O.reduce("logsumexp", x, "f64", list(axes), keep)[0],
                                           t.logsumexp(dim=axes, keepdim=keep).numpy().reshape(got.shape), rtol=1e-12)
                assert (O.reduce("max", x, "f64", list(axes), keep)[0] == t.amax(dim=axes, keepdim=keep).numpy().reshape(got.shape)).all()
                assert (O.reduce("min", x, "f64", list(axes), keep)[0] == t.amin(dim=axes, keepdim=keep).numpy().reshape(got.shape)).all()
    for ax in range(3):
        assert (O.reduce("argmax", x, "f64", [ax])[0] == t.argmax(dim=ax).numpy()).all()
        assert (O.reduce("argmin", x, "f64", [ax])[0] == t.argmin(dim=ax).numpy()).all()
    xi = rand(rng, (7, 9, 11), "i64", -1000, 1000)
    assert (O.reduce("sum", xi, "i64", [0, 2])[0] == torch.from_numpy(xi).sum(dim=(0, 2)).numpy()).all()
    assert (O.reduce("prod", xi % 3 + 1, "i64", [1])[0] == torch.from_numpy(xi % 3 + 1).prod(dim=1).numpy()).all()
    assert O.reduce("sum", x, "f64", [0, 1, 2])[0].shape == (1,)  # all axes → shape [1], not []


def test_argmax_tie_and_nan_rules():
    x = np.array([[1.0, 3.0, 3.0, 2.0], [np.nan, np.nan, np.nan, np.nan], [-np.inf, -np.inf, -np.inf, -np.inf],
                  [np.nan, 2.0, np.nan, 2.0]], dtype=np.float32)
    assert O.reduce("argmax", x, "f32", [1])[0].tolist() == [1, 0, 0, 1]
    assert O.reduce("argmin", x, "f32", [1])[0].tolist() == [0, 0, 0, 1]
    # the scalar loop of argreduce_kernels.rs:13-21, literally
    rng = np.random.default_rng(6)
    y = rng.integers(0, 3, size=(50, 40)).astype(np.float32)
    y[rng.random(y.shape) < 0.2] = np.nan
    ref = []
    for row in y:
        best, idx = -np.inf, 0
        for i, v in enumerate(row):
            if v > best:
                best, idx = v, i
        ref.append(idx)
    assert O.reduce("argmax", y, "f32", [1])[0].tolist() == ref


def test_softmax_matches_torch_on_reference_shapes():
    # hpt-tests/src/hpt/cpu/softmax.rs:31-54 and cuda/normalization.rs:33-53: arange inputs, both axes
    for shape in [(1, 13), (2, 1024), (3, 1123), (3, 4096), (3, 5551)]:
        x = np.arange(shape[0] * shape[1], dtype=np.float64).reshape(shape)
        for axis in (0, 1):
            np.testing.assert_allclose(O.softmax(x, "f64", axis)[0], torch.softmax(torch.from_numpy(x), axis).numpy(), rtol=1e-12, atol=1e-300)
            np.testing.assert_allclose(O.softmax(x, "f64", axis, True)[0], torch.log_softmax(torch.from_numpy(x), axis).numpy(), rtol=1e-12, atol=1e-9)


def test_shape_helpers():
    assert O.reduce_shape([4, 5, 6], [1], False) == [4, 6]
    assert O.reduce_shape([4, 5, 6], [0, 1, 2], False) == [1]
    assert O.process_axes([-1], 2) == [1]
    with pytest.raises(IndexError):
        O.process_axes([10], 2)
    with pytest.raises(ValueError):
        O.process_axes([1, 1], 3)


def test_cpu_port_agrees_with_numpy_oracle():
    """oracle/oracle_cpu.cpp (the timed CPU baseline) against oracle/hpt_oracle.py on small inputs."""
    path = os.path.join(ROOT, "oracle", "_build", "liboracle_cpu.so")
    if not os.path.exists(path):
        import build
        build.build_oracle()
    L = ctypes.CDLL(path)
    vp, i64, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    rng = np.random.default_rng(7)
    n = 96
    x = rng.standard_normal((n, n)).astype(np.float32)
    s = np.empty((n, n), np.float32)
    L.orc_unary_f32_strided2d.argtypes = [ci, vp, i64, i64, i64, i64, vp]
    for op, name in ((0, "sin"), (1, "exp")):
        L.orc_unary_f32_strided2d(op, x.ctypes.data, n, n, 1, n, s.ctypes.data)
        assert O.ulp_diff(s, O.unary(name, x.T, "f32")[0], "f32").max() <= 4  # libmvec: ≤ 4 ulp
    m, a = np.empty(n, np.float32), np.empty(n, np.int64)
    L.orc_max_f32_axis0.argtypes = [vp, i64, i64, i64, i64, vp]
    L.orc_argmax_f32_axis0.argtypes = [vp, i64, i64, i64, i64, vp]
    L.orc_max_f32_axis0(x.ctypes.data, n, n, 1, n, m.ctypes.data)
    L.orc_argmax_f32_axis0(x.ctypes.data, n, n, 1, n, a.ctypes.data)
    assert (m == O.reduce("max", x.T, "f32", [0])[0]).all() and (a == O.reduce("argmax", x.T, "f32", [0])[0]).all()
    b = rng.standard_normal((1, n)).astype(np.float32)
    c = np.empty((n, n), np.float32)
    L.orc_add_f32_bcast_row.argtypes = [vp, vp, vp, i64, i64]
    L.orc_add_f32_bcast_row(x.ctypes.data, b.ctypes.data, c.ctypes.data, n, n)
    assert (c == O.binary("add", x, "f32", b, "f32")[0]).all()
    k = rng.integers(-1000, 1000, size=n).astype(np.int64)
    d = np.empty((n, n), np.float64)
    L.orc_add_f32_i64_bcast_row.argtypes = [vp, vp, vp, i64, i64]
    L.orc_add_f32_i64_bcast_row(x.ctypes.data, k.ctypes.data, d.ctypes.data, n, n)
    assert (d == O.binary("add", x, "f32", k, "i64")[0]).all()
    r = np.empty(n, np.float32)
    L.orc_sum_f32_rows.argtypes = [vp, i64, i64, vp]
    L.orc_sum_f32_rows(x.ctypes.data, n, n, r.ctypes.data)
    np.testing.assert_allclose(r, x.astype(np.float64).sum(1), rtol=1e-5, atol=1e-5)
    L.orc_softmax_f32_rows.argtypes = [vp, i64, i64, vp]
    L.orc_softmax_f32_rows(x.ctypes.data, n, n, s.ctypes.data)
    np.testing.assert_allclose(s, O.softmax(x, "f32", 1)[0], rtol=1e-5)
    L.orc_logsumexp_f32_rows.argtypes = [vp, i64, i64, vp]
    L.orc_logsumexp_f32_rows(x.ctypes.data, n, n, r.ctypes.data)
    np.testing.assert_allclose(r, O.reduce("logsumexp", x, "f32", [1])[0], rtol=1e-5)
