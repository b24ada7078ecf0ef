"""GPU parity: softmax / log_softmax, copy / contiguous / astype, fill, allocator, transfers."""
import numpy as np
import pytest

from util import DTYPES, ENUM, O, assert_exact, assert_ulp, rand, to_numpy, to_torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hb():
    import hpt_b200
    return hpt_b200


def _softmax_check(hb, x, d, axis, log, view=None, streaming=False):
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
    # streaming=True — the two-sweep kernels for lanes that fit neither registers nor a cluster's shared memory: their
    # reference is r_K = fl(K·ln2), not the maximum itself, so the LARGEST term is exp(max − r_K) through ex2 (≤ 2 ulp)
    # instead of exactly 1, and that error reaches every output of the lane through Σ: 6 + |x − max| ulp.
    eps = {"f16": 2.0 ** -10, "bf16": 2.0 ** -7, "f32": 2.0 ** -23, "f64": 2.0 ** -52}[od]
    err = np.abs(np.asarray(got, np.float64) - np.asarray(want, np.float64))
    xc = np.asarray(O.to_compute(O.cast(x, d, od), od), np.float64)
    shift = np.abs(xc - np.max(xc, axis=axis, keepdims=True))
    shift = np.where(np.isfinite(shift), shift, 0.0)
    scale = np.max(np.abs(xc), axis=axis, keepdims=True) + 1.0
    ulp_ok = u <= (6 if streaming else 4) + np.ceil(shift)
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


@pytest.mark.parametrize("log", [False, True])
def test_cluster_band_and_streaming_kernels(hb, log):
    """Rows beyond the register limit and strided axes: the cluster-resident band kernels (softmax_band.cuh: rows of up to
    8 × 96 KB, 128-byte column bands of up to 6144 rows), and past those the streaming kernels with the exact power-of-two
    rescale.  The band kernels hold the register kernel's bound (4 + |x − max| ulp); each shape also runs with them switched
    off, i.e. through the streaming kernels, whose bound is 6 + |x − max| ulp (_softmax_check): no allowance that grows
    with the lane length any more."""
    import os
    rng = np.random.default_rng(33)
    cases = [((6, 131072), 1, "f32"), ((3, 200000), 1, "f32"), ((40, 50000), 1, "f32"), ((2, 400000), 1, "f32"),   # rows; 400000 > 8 × 96 KB
             ((5, 65536), 1, "bf16"), ((3, 300000), 1, "f16"),
             ((4096, 256), 0, "f32"), ((6144, 96), 0, "f32"), ((3000, 520), 0, "f32"), ((7000, 64), 0, "f32"),      # cols; 7000 rows > 6144
             ((2, 1000, 72), 1, "f32"), ((4096, 512), 0, "bf16"), ((777, 1032), 0, "f16"), ((64, 40), 0, "f32")]
    for shape, axis, d in cases:
        x = rand(rng, shape, d)
        x = (x * 4).astype(x.dtype) if d == "f32" else x
        if d == "f32":
            x.flat[7] = 30.0          # a late, dominant maximum: every running reference moves
            x.flat[x.size // 2] = -np.inf
        beyond_band = shape in ((2, 400000), (3, 300000), (7000, 64))
        for off in ("0", "1"):
            os.environ["HPTB_TUNE_NO_BAND"] = off
            try:
                _softmax_check(hb, x, d, axis, log, streaming=(off == "1" or beyond_band))
            finally:
                os.environ.pop("HPTB_TUNE_NO_BAND", None)
    # increasing rows: the running maximum moves at every element
    x = np.tile(np.linspace(-40, 40, 32768, dtype=np.float32), (3, 1))
    for off in ("0", "1"):
        os.environ["HPTB_TUNE_NO_BAND"] = off
        try:
            _softmax_check(hb, x, "f32", 1, log, streaming=(off == "1"))
            _softmax_check(hb, np.ascontiguousarray(x.T[:4096]), "f32", 0, log, streaming=(off == "1"))
        finally:
            os.environ.pop("HPTB_TUNE_NO_BAND", None)


@pytest.mark.parametrize("log", [False, True])
def test_large_offsets_and_special_lanes(hb, log):
    """Every softmax kernel (register rows, cluster bands, streaming rows / columns) on data far from zero — the streaming
    kernels' frame K is then in the thousands to millions, its reference r_K = fl(K·ln2) inexact — and on lanes holding
    NaN, +inf, −inf among finite values, or nothing but −inf (exp(−inf − (−inf)) = NaN, as the reference computes)."""
    import os
    rng = np.random.default_rng(35)
    cases = [((64, 4096), 1), ((5, 131072), 1), ((5, 400000), 1), ((4096, 64), 0), ((7000, 64), 0)]
    for shape, axis in cases:
        for offset in (5000.0, -20000.0, 3.0e5, 1.0e6):
            x = (rand(rng, shape, "f32") * 3 + np.float32(offset)).astype(np.float32)
            for off in ("0", "1"):
                os.environ["HPTB_TUNE_NO_BAND"] = off
                try:
                    _softmax_check(hb, x, "f32", axis, log, streaming=(off == "1" or shape in ((5, 400000), (7000, 64))))
                finally:
                    os.environ.pop("HPTB_TUNE_NO_BAND", None)
        x = (rand(rng, shape, "f32") * 3).astype(np.float32)
        lane = (lambda i: (slice(None), i)) if axis == 0 else (lambda i: (i, slice(None)))
        n = shape[axis]
        x[lane(0)][n // 3] = np.nan
        x[lane(1)][n // 2] = np.inf
        x[lane(2)][:] = -np.inf
        x[lane(3)][: n - 5] = -np.inf      # a handful of finite values at the very end
        x[lane(4)][5:] = -np.inf           # ... or at the very start
        for off in ("0", "1"):
            os.environ["HPTB_TUNE_NO_BAND"] = off
            try:
                with np.errstate(all="ignore"):
                    _softmax_check(hb, x, "f32", axis, log, streaming=(off == "1" or shape in ((5, 400000), (7000, 64))))
            finally:
                os.environ.pop("HPTB_TUNE_NO_BAND", None)


@pytest.mark.parametrize("log", [False, True])
def test_few_very_long_rows_split_into_slabs(hb, log):
    """Fewer rows than CTA slots (a 1-D softmax is ONE row): the streaming rows kernel splits every row into slabs, two
    launches (statistics per slab, then merge + apply).  Ragged slab ends, a late maximum, −inf and a NaN row."""
    rng = np.random.default_rng(36)
    for shape in ((3000000,), (1, 4 * 1000003), (3, 1300000), (20, 500000)):
        x = (rand(rng, shape, "f32") * 4).astype(np.float32)
        x.flat[x.size - 3] = 25.0
        x.flat[x.size // 2] = -np.inf
        _softmax_check(hb, x, "f32", len(shape) - 1, log, streaming=True)
    x = (rand(rng, (2, 2000000), "f32")).astype(np.float32)
    x[1, 1234567] = np.nan
    with np.errstate(all="ignore"):
        _softmax_check(hb, x, "f32", 1, log, streaming=True)
    x = rand(rng, (2, 1 << 21), "bf16")
    _softmax_check(hb, x, "bf16", 1, log, streaming=True)


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


@pytest.mark.gpu
def test_layernorm(hb):
    """NormalizationOps::layernorm vs the oracle (reference test: hpt-tests/src/hpt/cpu/softmax.rs:56-80, shapes
    [.., normalized] with gamma and beta against torch at 1e-5): register-resident rows, long rows, every dtype,
    sliced rows, optional gamma / beta."""
    rng = np.random.default_rng(51)
    EPS = {"f16": 2.0 ** -10, "bf16": 2.0 ** -7, "f32": 2.0 ** -23, "f64": 2.0 ** -52}

    def check(x, d, ns, use_g, use_b, view=None, eps=1e-5):
        od = O.float_out_binary(d, d)
        L = tuple(x.shape[len(x.shape) - ns:])
        g = rand(rng, L, od) if use_g else None
        b = rand(rng, L, od) if use_b else None
        X = hb.Tensor.to_cuda(to_torch(x, d))
        xv = x
        if view:
            X, xv = view(X), view(x)
        G = hb.Tensor.to_cuda(to_torch(g, od)) if use_g else None
        B = hb.Tensor.to_cuda(to_torch(b, od)) if use_b else None
        got_t = X.layernorm(xv.shape[xv.ndim - ns:], G, B, eps)
        assert got_t.dtype == ENUM[od] and tuple(got_t.shape) == tuple(xv.shape)
        ref, _ = O.layernorm(xv, d, ns, g, b, eps)
        got = np.asarray(to_numpy(got_t.to_cpu(), od), np.float64)
        tol = 4 * EPS[od] * np.maximum(1.0, np.abs(ref)) + (2e-6 if od == "f32" else 0.0)
        assert (np.abs(got - ref) <= tol).all(), f"layernorm {d} {xv.shape} ns={ns}: max err {np.abs(got - ref).max()}"

    for d in DTYPES:
        x = rand(rng, (6, 5, 64), d, 0, 20) if d in O.INTS else rand(rng, (6, 5, 64), d)
        check(x, d, 1, True, True)
    for shape, ns in (((3, 7, 13), 1), ((33, 1024), 1), ((5, 4096), 1), ((2, 3, 8192), 1), ((3, 20000), 1), ((4, 6, 32, 48), 2),
                      ((2, 3, 4, 5), 3), ((7, 129), 1)):
        x = rand(rng, shape, "f32")
        check(x, "f32", ns, True, True)
        check(x, "f32", ns, False, True)
        check(x, "f32", ns, True, False)
        check(x, "f32", ns, False, False)
    x = rand(rng, (40, 12, 256), "f32")
    check(x, "f32", 1, True, True, view=lambda t: t[3:30:2, 1:9])   # strided kept dims, contiguous rows
    check(rand(rng, (16, 512), "bf16"), "bf16", 1, True, True)
    with pytest.raises(hb.HptError):
        hb.Tensor.to_cuda(to_torch(x, "f32")).layernorm((128,))
    # C ABI: `in` and `out` sharing a PERMUTED layout over the normalized dims (in-place layernorm of x.transpose(-1, -2))
    # would merge into one unit-stride run in MEMORY order and put gamma on the wrong elements — it must be refused
    from ctypes import byref
    from hpt_b200 import _ffi
    xs = rand(rng, (4, 24, 16), "f32")
    X = hb.Tensor.to_cuda(to_torch(xs, "f32"))
    V = X.transpose(-1, -2)  # logical [4, 16, 24], strides [384, 1, 16]
    g = rand(rng, (16, 24), "f32")
    G = hb.Tensor.to_cuda(to_torch(g, "f32"))
    st = hb.lib.hptb_layernorm(X.ctx.handle, byref(V._c()), 2, byref(G._c()), None, 1e-5, byref(V._c()), hb.get_stream())
    assert st == 8, f"permuted normalized dims: expected HPTB_ERR_UNSUPPORTED, got status {st}"
    # … and through a contiguous copy the result is the oracle's
    got = V.contiguous().layernorm((16, 24), G, None, 1e-5).to_cpu().numpy()
    ref, _ = O.layernorm(np.transpose(xs, (0, 2, 1)), "f32", 2, g, None, 1e-5)
    assert (np.abs(got - ref) <= 2e-5 * (1 + np.abs(ref))).all()


@pytest.mark.gpu
def test_creation_ops(hb):
    """TensorCreator (normal_creation.rs:34-234): zeros / ones / full / arange / arange_step / linspace / eye, bit-exact
    against the reference formula `start + T(i)·step` evaluated with a rounding per step in T."""
    from hpt_b200 import _ffi

    def host_arange(n, start, step, d):
        i = O.cast(np.arange(n, dtype=np.uint64), "u64", d)
        if d == "bool":
            return np.logical_or(np.bool_(start), np.logical_and(i, np.bool_(step)))
        st = O.cast(np.array([step], dtype=np.float64 if d in O.FLOATS else np.int64), "f64" if d in O.FLOATS else "i64", d)
        s0 = O.cast(np.array([start], dtype=np.float64 if d in O.FLOATS else np.int64), "f64" if d in O.FLOATS else "i64", d)
        prod, _ = O.binary("mul", i, d, st, d)
        r, _ = O.binary("add", s0, d, prod, d)
        return r

    for d in DTYPES:
        e = ENUM[d]
        z = hb.Tensor.zeros((3, 5), e).to_cpu()
        o = hb.Tensor.ones((3, 5), e).to_cpu()
        assert (to_numpy(z, d) == 0).all() and (to_numpy(o, d) == 1).all()
        if d == "bool":
            continue
        lo, hi = (2, 40)
        got = to_numpy(hb.Tensor.arange(lo, hi, e).to_cpu(), d)
        assert_exact(got, host_arange(hi - lo, lo, 1, d), d, f"arange {d}")
        step = 3 if d in O.INTS else 0.37
        n = int(np.floor((hi - lo) / step)) + 1
        got = to_numpy(hb.Tensor.arange_step(lo, hi, step, e).to_cpu(), d)
        assert got.shape == (n,)
        assert_exact(got, host_arange(n, lo, step, d), d, f"arange_step {d}")
        f = to_numpy(hb.Tensor.full(7, (4, 9), e).to_cpu(), d)
        assert (f == 7).all()
    assert hb.Tensor.arange(5, 5, hb.I64).shape == (0,)
    for d in ("f32", "f64", "f16", "bf16"):
        for inc in (True, False):
            num = 57
            got = to_numpy(hb.Tensor.linspace(-1.5, 4.0, num, inc, ENUM[d]).to_cpu(), d)
            step = (4.0 + 1.5) / ((num - 1.0) if inc else num)
            want = host_arange(num, -1.5, step, d)
            if inc:
                want[-1] = O.cast(np.array([4.0]), "f64", d)[0]
            assert_exact(got, want, d, f"linspace {d} inc={inc}")
    for d in ("f32", "i64", "bool", "bf16"):
        for n, m, k in ((5, 7, 0), (6, 4, 1), (3, 3, 2)):
            got = to_numpy(hb.Tensor.eye(n, m, k, ENUM[d]).to_cpu(), d)
            want = np.eye(n, m, k)
            assert (got.astype(np.float64) == want).all(), f"eye {d} {n},{m},{k}"
    assert (hb.Tensor.identity(4, hb.F32).to_cpu().numpy() == np.eye(4)).all()
    big = hb.Tensor.arange(0, 1 << 20, hb.F32).to_cpu().numpy()
    assert (big == np.arange(1 << 20, dtype=np.float32)).all()
