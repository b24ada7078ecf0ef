"""CPU-only checks of the C ABI surface: the library loads, exports every symbol include/hpt_b200.h
declares, and its pure-host helpers (promotion, broadcasting, axes, collapse, allocator logic) agree
with the reference's golden data.  No compute calls: there is no GPU here."""
import ctypes
import json
import os
import re
from ctypes import byref, c_int, c_int32, c_int64, c_uint8, POINTER

import pytest

from util import DTYPES, ENUM, ROOT

from hpt_b200 import _ffi
from hpt_b200._ffi import HptbCollapsePlan, HptbTensor, HptError, check, lib, make_tensor


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "hpt_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(hptb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 35
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"symbols declared in the header but not exported: {missing}"
    # and the ctypes table covers exactly the header
    assert set(_ffi.SIGNATURES) == declared, set(_ffi.SIGNATURES) ^ declared
    assert not _ffi.MISSING


def test_rust_sys_crate_binds_exactly_the_header():
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "hpt_b200.h")).read(), flags=re.S)
    declared = set(re.findall(r"\b(hptb_[a-z0-9_]+)\s*\(", hdr))
    rs = set(re.findall(r"pub fn (hptb_[a-z0-9_]+)", open(os.path.join(ROOT, "rust", "hpt-b200-sys", "src", "lib.rs")).read()))
    assert declared == rs, declared ^ rs


def test_rust_sys_crate_is_generated_from_the_header_and_ctypes_agrees():
    """One source of truth: rust/hpt-b200-sys/src/lib.rs is exactly what tools/gen_rust_sys.py emits for the current
    header, and the ctypes mirror carries the same enum VALUES and struct FIELD LISTS (names, order, types, array extents)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_rust_sys as g
    assert g.generate() == open(g.OUT).read(), "rust/hpt-b200-sys/src/lib.rs is stale: run python tools/gen_rust_sys.py"
    h = g.parse_header()
    enums = {k: dict(v) for k, v in h["enums"].items()}

    def by_name(enum, prefix, table, count):
        for name, val in table.items():
            assert enums[enum][prefix + name.upper()] == val, (enum, name)
        assert enums[enum][count] == len(table), (enum, "count")
    by_name("hptb_binary_op", "HPTB_", _ffi.BINARY_OPS, "HPTB_BINARY_COUNT")
    by_name("hptb_cmp_op", "HPTB_", _ffi.CMP_OPS, "HPTB_CMP_COUNT")
    by_name("hptb_unary_op", "HPTB_", _ffi.UNARY_OPS, "HPTB_UNARY_COUNT")
    by_name("hptb_reduce_op", "HPTB_", _ffi.REDUCE_OPS, "HPTB_REDUCE_COUNT")
    assert enums["hptb_unary_op"]["HPTB_FLOOR"] == _ffi.FLOAT_UNARY_COUNT
    for i, n in enumerate(_ffi.DTYPE_NAMES):
        assert enums["hptb_dtype"]["HPTB_" + n.upper()] == i
    for code, n in _ffi.STATUS_NAMES.items():
        assert enums["hptb_status"]["HPTB_" + ("OK" if n == "OK" else "ERR_" + n)] == code
    for i, n in enumerate(_ffi.ROUTES):
        assert enums["hptb_route_kind"]["HPTB_ROUTE_" + n.upper()] == i
    for i, n in enumerate(_ffi.COLLECTIVES):
        assert enums["hptb_collective"]["HPTB_COLL_" + n.upper()] == i
    assert (_ffi.PROMOTE_NORMAL, _ffi.PROMOTE_FLOAT_BINARY, _ffi.PROMOTE_FLOAT_UNARY) == tuple(
        enums["hptb_promote_kind"][k] for k in ("HPTB_PROMOTE_NORMAL", "HPTB_PROMOTE_FLOAT_BINARY", "HPTB_PROMOTE_FLOAT_UNARY"))
    cmap = {"void*": ctypes.c_void_p, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64,
            "uint8_t": ctypes.c_uint8}
    for cname, cls in (("hptb_tensor", _ffi.HptbTensor), ("hptb_alloc_stats", _ffi.HptbAllocStats),
                       ("hptb_collapse_plan", _ffi.HptbCollapsePlan), ("hptb_reduce_route_t", _ffi.HptbReduceRoute),
                       ("hptb_shard_plan", _ffi.HptbShardPlan)):
        want = []
        for fname, ctype, dims in h["structs"][cname]:
            t = cmap[ctype]
            for d in reversed(dims):
                t = t * d
            want.append((fname, t))
        got = [(n, t) for n, t in cls._fields_]
        assert [n for n, _ in got] == [n for n, _ in want], cname
        for (n, tg), (_, tw) in zip(got, want):
            assert ctypes.sizeof(tg) == ctypes.sizeof(tw) and tg._type_ == tw._type_ if hasattr(tw, "_length_") else tg is tw, (cname, n)
        assert ctypes.sizeof(cls) == sum(ctypes.sizeof(t) for _, t in want) or cname in ("hptb_collapse_plan",), cname


def test_rust_shim_op_tables_match_the_header():
    """rust/hpt-b200-shim maps the reference's op-name strings to the library's enums: the unary table is positional,
    the binary / reduce tables name the constants — both must agree with the header."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_rust_sys as g
    enums = {k: dict(v) for k, v in g.parse_header()["enums"].items()}
    shim = os.path.join(ROOT, "rust", "hpt-b200-shim", "src")
    un = open(os.path.join(shim, "unary.rs")).read()
    names = re.findall(r'"(\w+)"', re.search(r"const NAMES: \[&str; (\d+)\] = \[(.*?)\];", un, flags=re.S).group(2))
    assert len(names) == enums["hptb_unary_op"]["HPTB_UNARY_COUNT"] == int(re.search(r"const NAMES: \[&str; (\d+)\]", un).group(1))
    for i, n in enumerate(names):
        assert enums["hptb_unary_op"]["HPTB_" + n.upper()] == i, n
    for fname, enum in (("binary.rs", ("hptb_binary_op", "hptb_cmp_op")), ("reduce.rs", ("hptb_reduce_op",))):
        src = open(os.path.join(shim, fname)).read()
        pairs = re.findall(r'"(\w+)" => \(?(?:true, |false, )?sys::(HPTB_\w+)', src)
        assert len(pairs) >= 16
        for name, const in pairs:
            assert any(const in enums[e] for e in enum), const
            want = {"max": "MAXIMUM", "min": "MINIMUM"}.get(name, name.upper()) if fname == "binary.rs" else name.upper()
            assert const == "HPTB_" + want, (name, const)


def test_version_and_dtype_sizes():
    assert lib.hptb_version() == 100
    for i, n in enumerate(DTYPES):
        assert lib.hptb_dtype_name(i).decode() == n
        assert lib.hptb_dtype_size(i) == _ffi.DTYPE_SIZES[i]


def test_promotion_tables_match_reference_golden():
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "promotion.json")))
    for a in DTYPES:
        assert DTYPES[lib.hptb_promote(ENUM[a], 0, 2)] == g["float_out_unary"][a]
        for b in DTYPES:
            assert DTYPES[lib.hptb_promote(ENUM[a], ENUM[b], 0)] == g["normal_out"][a][b], (a, b)
            assert DTYPES[lib.hptb_promote(ENUM[a], ENUM[b], 1)] == g["float_out_binary"][a][b], (a, b)
    # the asymmetry called out in SURVEY.md §8c
    assert DTYPES[lib.hptb_promote(ENUM["i32"], ENUM["f16"], 0)] == "f16"
    assert DTYPES[lib.hptb_promote(ENUM["f16"], ENUM["i32"], 0)] == "f32"
    assert DTYPES[lib.hptb_promote(ENUM["f32"], ENUM["i64"], 0)] == "f64"  # config 4
    assert lib.hptb_promote(13, 0, 0) == -1


def test_out_dtype_helpers():
    assert lib.hptb_binary_out_dtype(_ffi.BINARY_OPS["add"], ENUM["bool"], ENUM["bool"]) == ENUM["bool"]
    assert lib.hptb_binary_out_dtype(_ffi.BINARY_OPS["sub"], ENUM["bool"], ENUM["bool"]) == -1  # _bool.rs:31 panics
    assert lib.hptb_binary_out_dtype(_ffi.BINARY_OPS["div"], ENUM["i32"], ENUM["i32"]) == ENUM["f32"]
    assert lib.hptb_unary_out_dtype(_ffi.UNARY_OPS["sin"], ENUM["i64"]) == ENUM["f64"]
    assert lib.hptb_reduce_out_dtype(_ffi.REDUCE_OPS["mean"], ENUM["bf16"]) == ENUM["bf16"]
    assert lib.hptb_reduce_out_dtype(_ffi.REDUCE_OPS["mean"], ENUM["i32"]) == ENUM["f32"]
    assert lib.hptb_reduce_out_dtype(_ffi.REDUCE_OPS["argmax"], ENUM["f32"]) == ENUM["i64"]
    assert lib.hptb_reduce_out_dtype(_ffi.REDUCE_OPS["sum"], ENUM["u8"]) == ENUM["u8"]


def _bshape(a, b):
    out = (c_int64 * 8)()
    n = c_int()
    check(lib.hptb_broadcast_shape((c_int64 * max(len(a), 1))(*a), len(a), (c_int64 * max(len(b), 1))(*b), len(b), out, byref(n)))
    return [out[i] for i in range(n.value)]


def test_broadcast_shape_reference_vectors():
    # hpt-tests/src/hpt_common/layout.rs:23-40
    assert _bshape([5, 2, 10], [5, 1, 10]) == [5, 2, 10]
    with pytest.raises(HptError) as e:
        _bshape([5, 2, 10], [5, 1, 11])
    assert "Broadcasting error: broadcast failed at index 2, lhs shape: [5, 2, 10], rhs shape: [5, 1, 11]" in str(e.value)
    assert e.value.status == 1
    assert _bshape([4096, 4096], [1, 4096]) == [4096, 4096]
    assert _bshape([32, 128, 4096], [4096]) == [32, 128, 4096]
    assert _bshape([1, 100, 1, 100], [100, 1, 100, 1]) == [100, 100, 100, 100]  # docs/benchmarks/binary.md
    assert _bshape([], [3]) == [3]


def _axes(axes, ndim):
    out = (c_int32 * max(len(axes), 1))()
    check(lib.hptb_process_axes((c_int64 * max(len(axes), 1))(*axes), len(axes), ndim, out))
    return [out[i] for i in range(len(axes))]


def test_process_axes_reference_vectors():
    # hpt-tests/src/hpt_common/axis.rs:6-50
    assert _axes([-1], 2) == [1]
    with pytest.raises(HptError) as e:
        _axes([10], 2)
    assert "Dimension out of range: expected in 0..2, got 10" in str(e.value)
    with pytest.raises(HptError) as e:
        _axes([-3], 2)
    assert "Dimension out of range: expected in 0..2, got -1" in str(e.value)
    with pytest.raises(HptError) as e:
        _axes([1, 1], 3)
    assert "Axis 1 is duplicated" in str(e.value) and e.value.status == 3
    assert _axes([0, -1], 3) == [0, 2]


def test_reduce_shape():
    def rs(shape, axes, keep):
        out = (c_int64 * 8)()
        n = c_int()
        check(lib.hptb_reduce_shape((c_int64 * len(shape))(*shape), len(shape), (c_int32 * max(len(axes), 1))(*axes), len(axes),
                                    keep, out, byref(n)))
        return [out[i] for i in range(n.value)]
    assert rs([4, 5, 6], [1], 0) == [4, 6]
    assert rs([4, 5, 6], [1], 1) == [4, 1, 6]
    assert rs([4, 5, 6], [0, 1, 2], 0) == [1]  # layout_utils.rs:342-347: never rank 0
    assert rs([4, 5, 6], [0, 1, 2], 1) == [1, 1, 1]


def _collapse(ops, mask=None):
    ts = [make_tensor(0x1000, ENUM["f32"], s, st) for s, st in ops]
    arr = (POINTER(HptbTensor) * len(ts))(*[ctypes.pointer(t) for t in ts])
    plan = HptbCollapsePlan()
    m = (c_uint8 * 8)(*mask) if mask is not None else None
    check(lib.hptb_collapse(arr, len(ts), m, byref(plan)))
    nd = plan.ndim
    return (plan.launch_class, [plan.shape[i] for i in range(nd)],
            [[plan.strides[o][i] for i in range(nd)] for o in range(len(ts))], [plan.reduced[i] for i in range(nd)])


def test_collapse_elementwise_classes():
    C, I, S = 0, 1, 2
    # same-shape contiguous: one dim
    cls, shape, st, _ = _collapse([((4096, 4096), (4096, 1))] * 3)
    assert (cls, shape, st) == (C, [4096 * 4096], [[1], [1], [1]])
    # config 1: [4096,4096] + [1,4096] → inner-contiguous, rhs outer stride 0
    cls, shape, st, _ = _collapse([((4096, 4096), (4096, 1)), ((4096, 4096), (4096, 1)), ((1, 4096), (4096, 1))])
    assert (cls, shape, st) == (I, [4096, 4096], [[4096, 1], [4096, 1], [0, 1]])
    # scalar operand: everything folds to one dim, stride 0
    cls, shape, st, _ = _collapse([((8, 16), (16, 1)), ((8, 16), (16, 1)), ((1,), (1,))])
    assert (cls, shape, st) == (C, [128], [[1], [1], [0]])
    # config 4: [32,128,4096] + [4096] → outer dims merge
    cls, shape, st, _ = _collapse([((32, 128, 4096), (524288, 4096, 1)), ((32, 128, 4096), (524288, 4096, 1)), ((4096,), (1,))])
    assert (cls, shape, st) == (I, [4096, 4096], [[4096, 1], [4096, 1], [0, 1]])
    # config 2: transposed input → strided
    cls, shape, st, _ = _collapse([((8192, 8192), (8192, 1)), ((8192, 8192), (1, 8192))])
    assert (cls, shape, st) == (S, [8192, 8192], [[8192, 1], [1, 8192]])
    # the reference's broadcast benchmark [1,100,1,100] + [100,1,100,1]
    cls, shape, st, _ = _collapse([((100, 100, 100, 100), (1000000, 10000, 100, 1)), ((1, 100, 1, 100), (10000, 100, 100, 1)),
                                   ((100, 1, 100, 1), (100, 100, 1, 1))])
    assert cls == I and shape == [100, 100, 100, 100]  # inner dim: lhs stride 1, rhs broadcast (stride 0)
    assert st[1] == [0, 100, 0, 1] and st[2] == [100, 0, 1, 0]
    # size-1 dims vanish, sliced-with-step inner dim is strided
    cls, shape, st, _ = _collapse([((1, 6, 1, 5), (30, 5, 5, 1)), ((1, 6, 1, 5), (120, 20, 20, 2))])
    assert (cls, shape, st) == (S, [6, 5], [[5, 1], [20, 2]])
    # Layout::coalesce_dims example: fully contiguous 3-D → 1 dim
    cls, shape, st, _ = _collapse([((2, 5, 10), (50, 10, 1)), ((2, 5, 10), (50, 10, 1))])
    assert (cls, shape) == (C, [100])


def test_collapse_reduce():
    # config 3: NCHW viewed NHWC, reduce view axes (0,1,2) → kept C, reduced (N, HW merged)
    _, shape, st, red = _collapse([((512,), (1,)), ((64, 56, 56, 512), (1605632, 56, 1, 3136))], mask=[1, 1, 1, 0])
    assert shape == [512, 64, 3136] and red == [0, 1, 1]
    assert st[1] == [3136, 1605632, 1] and st[0] == [1, 0, 0]
    # config 2: transposed view, reduce view axis 0 (the unit-stride one)
    _, shape, st, red = _collapse([((8192,), (1,)), ((8192, 8192), (1, 8192))], mask=[1, 0])
    assert shape == [8192, 8192] and red == [0, 1] and st[1] == [8192, 1]
    # full reduce of a contiguous tensor: one reduced dim
    _, shape, st, red = _collapse([((1,), (1,)), ((262144, 16384), (16384, 1))], mask=[1, 1])
    assert shape == [262144 * 16384] and red == [1]
    # axis-0 reduce keeps the contiguous dim
    _, shape, st, red = _collapse([((16384,), (1,)), ((262144, 16384), (16384, 1))], mask=[1, 0])
    assert shape == [16384, 262144] and red == [0, 1] and st[1] == [1, 16384]
    # two kept dims around a reduced one do not merge; 1024×1024×80 axis 1 (docs/benchmarks/reduce.md)
    _, shape, st, red = _collapse([((1024, 80), (80, 1)), ((1024, 1024, 80), (81920, 80, 1))], mask=[0, 1, 0])
    assert shape == [1024, 80, 1024] and red == [0, 0, 1]


def test_alloc_selftest_fake_device():
    check(lib.hptb_alloc_selftest())


def test_errors_without_gpu():
    h = ctypes.c_void_p()
    st = lib.hptb_ctx_create(0, byref(h))
    import torch
    if not torch.cuda.is_available():
        assert st == 5 and b"cuda" in lib.hptb_last_error().lower()  # DeviceError, not a crash
    t = make_tensor(0, ENUM["f32"], (2, 2), (2, 1))
    assert lib.hptb_binary(None, 0, byref(t), byref(t), byref(t), None) == 4
    assert b"null ctx" in lib.hptb_last_error()


def test_rust_sys_crate_declares_every_abi_entry():
    """rust/hpt-b200-sys is delivered as source (no cargo in the image): keep its extern block in step with the header."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "hpt_b200.h")).read()
    rust = open(os.path.join(root, "rust", "hpt-b200-sys", "src", "lib.rs")).read()
    declared = set(re.findall(r"\b(hptb_[a-z0-9_]+)\s*\(", header)) - {"hptb_status"}
    bound = set(re.findall(r"pub fn (hptb_[a-z0-9_]+)", rust))
    assert declared == bound, (sorted(declared - bound), sorted(bound - declared))
    # struct images that cross the boundary by value or pointer
    assert "pub post_root: i32" in rust and "post_root" in header


def test_reduce_route_is_pure_host_logic():
    """hptb_reduce_route: how hptb_reduce will run a reduction — decided from shapes, strides, dtypes and the alignment of
    the input pointer alone (api_reduce.cpp plan_reduce_route), so it is testable without a GPU."""
    from hpt_b200._ffi import REDUCE_OPS, ROUTES, HptbReduceRoute

    def route(op, dtype, shape, strides, axes, out_shape, out_strides, ptr=0x10000, init_out=1):
        t_in = make_tensor(ptr, ENUM[dtype], shape, strides)
        odt = lib.hptb_reduce_out_dtype(REDUCE_OPS[op], ENUM[dtype])
        t_out = make_tensor(0x900000, odt, out_shape, out_strides)
        r = HptbReduceRoute()
        ax = (ctypes.c_int32 * len(axes))(*axes)
        check(lib.hptb_reduce_route(REDUCE_OPS[op], byref(t_in), ax, len(axes), byref(t_out), init_out, byref(r)))
        return ROUTES[r.kind], (r.head, r.body, r.tail), [r.scratch_strides[i] for i in range(len(out_shape))]

    # aligned rows: direct
    assert route("sum", "f32", (7995, 8097), (8192, 1), [1], (7995,), (1,))[0] == "direct"
    # a[5:8000, 3:8100]: base 12 bytes past a 16-byte boundary, row stride 8192 elements → head 1, body 8096, tail 0
    kind, hbt, _ = route("sum", "f32", (7995, 8097), (8192, 1), [1], (7995,), (1,), ptr=0x10000 + 3 * 4)
    assert (kind, hbt) == ("peel", (1, 8096, 0))
    kind, hbt, _ = route("max", "i8", (300, 1030), (1040, 1), [1], (300,), (1,), ptr=0x10000 + 5)
    assert (kind, hbt) == ("peel", (11, 1008, 11))
    # rows whose stride is not a multiple of a pack have no common misalignment: no peeling
    assert route("sum", "f32", (7995, 8097), (8191, 1), [1], (7995,), (1,), ptr=0x10000 + 12)[0] == "direct"
    # half types (f32 accumulator, half output) and ops whose output is not their accumulator peel through a scratch of
    # accumulators: bf16 base 6 bytes past a boundary → head (16 − 6) / 2 = 5 elements
    kind, hbt, _ = route("sum", "bf16", (7995, 8097), (8192, 1), [1], (7995,), (1,), ptr=0x10000 + 6)
    assert (kind, hbt) == ("peel_raw", (5, 8088, 4))
    kind, hbt, _ = route("mean", "f32", (7995, 8097), (8192, 1), [1], (7995,), (1,), ptr=0x10000 + 12)
    assert (kind, hbt) == ("peel_raw", (1, 8096, 0))
    assert route("logsumexp", "f32", (7995, 8097), (8192, 1), [1], (7995,), (1,), ptr=0x10000 + 12)[0] == "peel_raw"
    # arg reductions carry indices: not peeled
    assert route("argmax", "f32", (7995, 8097), (8192, 1), [1], (7995,), (1,), ptr=0x10000 + 12)[0] == "direct"
    assert route("sum", "f32", (7995, 8097), (8192, 1), [1], (7995,), (1,), ptr=0x10000 + 12, init_out=0)[0] == "direct"
    # x[256,512,512].permute(2,0,1).sum(2): view shape (512,256,512), strides (1,262144,512); out (512,256) contiguous.
    # kept dims: view 0 (input stride 1) and view 1 (input stride 262144) → scratch strides (1, 512): input order
    kind, _, ss = route("sum", "f32", (512, 256, 512), (1, 262144, 512), [2], (512, 256), (256, 1))
    assert (kind, ss) == ("two_step", [1, 512])
    # the same reduction into an output that already has the input's order needs no second step
    assert route("sum", "f32", (512, 256, 512), (1, 262144, 512), [2], (512, 256), (1, 512))[0] == "direct"
    # config 2 (transposed view, axis 0) and config 3 (NHWC view) are single-kept-dim reductions: direct
    assert route("max", "f32", (8192, 8192), (1, 8192), [0], (8192,), (1,))[0] == "direct"
    assert route("mean", "bf16", (64, 56, 56, 512), (1605632, 56, 1, 3136), [0, 1, 2], (512,), (1,))[0] == "direct"
    # short reduced extents are not worth a second launch
    assert route("sum", "f32", (512, 256, 4), (1, 2048, 512), [2], (512, 256), (256, 1))[0] == "direct"
