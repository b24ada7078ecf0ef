"""GPU parity: softmax / log_softmax, copy / contiguous / astype, fill, allocator, transfers."""
import numpy as np
import pytest

from util import DTYPES, ENUM, O, assert_exact, assert_ulp, rand, to_numpy, to_torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hb():
    import hpt_b200
    return hpt_b200


def _softmax_check(hb, x, d, axis, log, view=None):
    X = hb.Tensor.to_cuda(to_torch(x, d))
    if view:
        X, x = view(X), view(x)
    got_t = X.log_softmax(axis) if log else X.softmax(axis)
    want, od = O.softmax(x, d, axis, log)
    assert got_t.dtype == ENUM[od] and tuple(got_t.shape) == tuple(x.shape)
    got = to_numpy(got_t.to_cpu(), od)
    u = O.ulp_diff(got, want, od)
    # softmax = exp(x − max) ÷ Σ is a composite, not an elementwise transcendental.  Error budget against the
    # f64-evaluated oracle, for the f32 formula the reference's CPU kernel uses (cpu/kernels/softmax.rs:204-310):
    # exp ≤ 2 ulp, Σ and the scaling ≤ 2 ulp, and the ROUNDING OF x − max, an absolute error of ≤ ulp(x − max)/2
    # that exp turns into a relative error of |x − max|·2^-24, i.e. |x − max|/2 ulp of the result.  Bound:
    # 4 + |x − max| ulp (the reference's own tests use allclose 1e-3).  log_softmax subtracts two O(|x|)
    # numbers: absolute bound 4·eps·max|x| as well.
    eps = {"f16": 2.0 ** -10, "bf16": 2.0 ** -7, "f32": 2.0 ** -23, "f64": 2.0 ** -52}[od]
    err = np.abs(np.asarray(got, np.float64) - np.asarray(want, np.float64))
    xc = np.asarray(O.to_compute(O.cast(x, d, od), od), np.float64)
    shift = np.abs(xc - np.max(xc, axis=axis, keepdims=True))
    shift = np.where(np.isfinite(shift), shift, 0.0)
    scale = np.max(np.abs(xc), axis=axis, keepdims=True) + 1.0
    ulp_ok = u <= 4 + np.ceil(shift)
    ok = ulp_ok | (err <= 4 * eps * scale if log else err <= 4 * eps * np.maximum(np.asarray(want, np.float64), 1e-30) + 1e-45)
    assert ok.all(), f"softmax log={log} {d} shape={x.shape} axis={axis}: max ulp {u.max()}"


@pytest.mark.parametrize("log", [False, True])
def test_reference_shapes_arange(hb, log):
    # hpt-tests/src/hpt/cuda/normalization.rs:33-53: arange inputs, both axes
    for shape in [(1, 13), (2, 1024), (3, 1123), (3, 4096), (3, 5551)]:
        n = shape[0] * shape[1]
        for d in ("f32", "f64", "f16"):
            x = np.arange(n, dtype=np.float64).reshape(shape)
            if d == "f16":
                x = x % 1000  # keep arange representable
            x = x.astype(O.NP[d])
            for axis in (0, 1):
                _softmax_check(hb, x, d, axis, log)


@pytest.mark.parametrize("log", [False, True])
def test_dtypes_shapes_layouts(hb, log):
    rng = np.random.default_rng(30)
    for d in DTYPES:
        x = rand(rng, (5, 7, 130), d, -6, 6) if d in O.INTS else rand(rng, (5, 7, 130), d)
        for axis in (0, 1, 2, -1):
            _softmax_check(hb, x, d, axis, log)
    for shape in [(4, 1), (1, 4), (33, 127), (9, 1024), (9, 1028), (5, 4100), (3, 8192), (2, 9000), (2, 40000), (300, 31)]:
        x = rand(rng, shape, "f32") * 3
        for axis in (0, 1):
            _softmax_check(hb, x, "f32", axis, log)
    x = rand(rng, (12, 34, 56), "f32")
    _softmax_check(hb, x, "f32", 2, log, lambda t: t.permute([2, 0, 1]) if hasattr(t, "storage") else np.transpose(t, (2, 0, 1)))
    _softmax_check(hb, x, "f32", 0, log, lambda t: t.permute([1, 0, 2]) if hasattr(t, "storage") else np.transpose(t, (1, 0, 2)))
    _softmax_check(hb, x, "f32", 1, log, lambda t: t[::2, 3:30, ::3])
    # config 4 row length
    _softmax_check(hb, rand(rng, (64, 4096), "f32"), "f32", -1, log)


def test_softmax_errors(hb):
    X = hb.Tensor.empty((4, 5), ENUM["f32"])
    with pytest.raises(hb.HptError) as e:
        X.softmax(2)
    assert e.value.status == 3


def test_contiguous_and_astype(hb):
    rng = np.random.default_rng(31)
    for d in DTYPES:
        x = rand(rng, (20, 30, 12), d)
        X = hb.Tensor.to_cuda(to_torch(x, d))
        for v in (lambda t: t.permute([2, 0, 1]) if hasattr(t, "storage") else np.transpose(t, (2, 0, 1)),
                  lambda t: t[1:19:2, ::3, 2:11]):
            got = v(X).contiguous()
            assert got.is_contiguous()
            assert_exact(to_numpy(got.to_cpu(), d), np.ascontiguousarray(v(x)), d, f"contiguous {d}")
            assert_exact(to_numpy(v(X).to_cpu(), d), np.ascontiguousarray(v(x)), d, f"to_cpu of view {d}")
    # astype = Rust `as` for every pair, including saturation / NaN→0 / wrap
    special = np.array([0.0, -0.0, 1.5, -1.5, 127.6, -128.9, 255.5, 3e9, -3e9, 1e19, -1e19, 65504.0, 70000.0, np.inf, -np.inf,
                        np.nan, 0.1, 1e-8, 16777217.0, 2.0 ** 63], dtype=np.float64)
    for src in DTYPES:
        if src in O.FLOATS:
            base = O.cast(special, "f64", src)
        else:
            info = None if src == "bool" else np.iinfo(O.NP[src])
            base = rand(rng, (40,), src)
            if info is not None:
                base[:4] = [info.min, info.max, 0, 1]
        X = hb.Tensor.to_cuda(to_torch(base, src))
        for dst in DTYPES:
            want = O.cast(base, src, dst)
            got = to_numpy(X.astype(ENUM[dst]).to_cpu(), dst)
            assert_exact(got, want, dst, f"astype {src}->{dst}")


def test_fill(hb):
    for d, v in (("f32", 1.5), ("i64", -7), ("u8", 200), ("bf16", 0.5), ("f64", 2.25), ("bool", True), ("i16", -3)):
        T = hb.Tensor.empty((37, 41), ENUM[d])
        T.fill_(v)
        got = to_numpy(T.to_cpu(), d)
        assert (np.asarray(got, np.float64) == float(v)).all()
        # strided view: only the view is written
        T.fill_(0)
        T[::2, 1::3].fill_(v)
        got = np.asarray(to_numpy(T.to_cpu(), d), np.float64)
        want = np.zeros((37, 41))
        want[::2, 1::3] = float(v)
        assert (got == want).all()


def test_allocator_caches_and_transfers(hb):
    ctx = hb.context(0)
    ctx.empty_cache()
    s0 = ctx.alloc_stats()
    a = hb.Tensor.empty((1000, 1000), ENUM["f32"])
    p = a.ptr
    del a
    b = hb.Tensor.empty((1000, 999), ENUM["f32"])  # same size class → same block, no device malloc
    s1 = ctx.alloc_stats()
    assert b.ptr == p
    assert s1["n_cache_hit"] >= s0["n_cache_hit"] + 1
    assert s1["n_device_malloc"] == s0["n_device_malloc"] + 1
    del b
    ctx.empty_cache()
    assert ctx.alloc_stats()["bytes_cached"] == 0
    # steady-state op loop performs no device mallocs
    x = hb.Tensor.to_cuda(to_torch(np.ones((256, 256), np.float32), "f32"))
    y = x + x
    z = y.sum([1])
    del y, z
    m0 = ctx.alloc_stats()["n_device_malloc"]
    for _ in range(20):
        y = x + x
        z = y.sum([1])
        del y, z
    assert ctx.alloc_stats()["n_device_malloc"] == m0
    with pytest.raises(hb.HptError) as e:
        hb.Tensor.empty((1 << 20, 1 << 20), ENUM["f64"])  # 8 TiB
    assert e.value.status == 6


def test_mean_var_extension(hb):
    rng = np.random.default_rng(32)
    for d in ("bf16", "f16", "f32"):
        x = rand(rng, (4, 16, 14, 14), d)
        if d != "bf16":
            x = (x + 3).astype(x.dtype)
        X = hb.Tensor.to_cuda(to_torch(x, d)).permute([0, 2, 3, 1])
        m, v = X.mean_var([0, 1, 2])
        wm, wv, od = O.mean_var(np.transpose(x, (0, 2, 3, 1)), d, [0, 1, 2])
        gm, gv = to_numpy(m.to_cpu(), od).astype(np.float64), to_numpy(v.to_cpu(), od).astype(np.float64)
        eps = {"f16": 2.0 ** -10, "bf16": 2.0 ** -7, "f32": 2.0 ** -23}[od]
        assert np.all(np.abs(gm - wm) <= 2 * eps * np.abs(wm) + 1e-6)
        assert np.all(np.abs(gv - wv) <= 2 * eps * np.abs(wv) + 1e-6)
