"""CPU oracle: a numpy restatement of Hpt's CPU semantics for the hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under hpt_b200/ imports this; only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may.  It is the checker, never the product.

What it restates (paths relative to the Hpt repository):
  promotion      hpt-types/src/promotion/normal_promote/_*.rs          (tables read from tests/golden/promotion.json,
                                                                         which gen_promotion_golden.py extracts from those files)
  casts          hpt-macros/src/scalar_convert.rs:39-255               (Rust `as`: int→int wraps, float→int saturates, NaN→0,
                                                                         →f16/bf16 via f32 for ≤32-bit ints except u32, via f64 for u32/i64/u64)
  binary ops     hpt-macros/src/normal_out.rs:34-135                   (cast both operands to Output, then the scalar op)
                 hpt-types/src/scalars/impls.rs:29-70                  (wrapping_add/sub/mul/rem, max/min)
                 hpt-types/src/scalars/_bool.rs:25-63                  (add = OR, mul = AND, max = OR, min = AND)
                 hpt-types/src/scalars/_bf16.rs:28-66                  (half types: f32 arithmetic, one rounding)
  broadcasting   hpt-common/src/shape/shape_utils.rs:370-400
  unary ops      hpt-types/src/scalars/_f32.rs:182-330                 (std / libm formulas)
  normal unary   hpt-types/src/scalars/_f32.rs:70-130, impls.rs:71-131, _bool.rs (NormalOutUnary2: T → T)
  pow / hypot    hpt-types/src/scalars/_f32.rs:14-20                   (f32::powf / hypot in FloatOutBinaryPromote)
  bit ops        hpt-macros/src/lib.rs:384-480, impls.rs:133-163       (cast to NormalOutPromote; wrapping shifts)
  compares       hpt-macros/src/lib.rs:570-650                         (cast to NormalOutPromote, compare, bool)
  reductions     hpt/src/backends/cpu/tensor_internal/common_reduce.rs:32-168, :352-380, :451-480
  argmax/argmin  hpt/src/backends/cpu/kernels/argreduce_kernels.rs:13-21,49-57
  softmax        hpt/src/backends/cpu/kernels/softmax.rs:204-310
  output shapes  hpt-common/src/layout/layout_utils.rs:310-349, hpt-common/src/axis/axis.rs:38-72

Pinning: the promotion tables are the reference's own (golden JSON); casts, wrapping arithmetic and
f16/bf16 conversions are pinned against the known-answer values of hpt-tests/src/hpt_types/tests.rs
(tests/test_oracle.py); reductions / softmax / unary follow the reference's own test oracle, libtorch
(hpt-tests/src/hpt/cpu/reduce.rs:14), which tests/test_oracle.py cross-checks with torch CPU.
Cross-dtype promotion at tensor level is pinned by NO reference test; for the four NormalBinOps
it is pinned instead to the OUTPUT of the reference's own CUDA kernels (oracle/_ref/binary_*.cubin, every dtype pair,
tests/test_reference_kernels_gpu.py), as are argmax / argmin and strided copies.  For the unary and reduce paths,
whose reference kernels do not build here, it stays "parity unpinned": the tables are the only spec.

Representation: numpy arrays; f16 is np.float16; bf16 values are carried as np.float32 arrays holding
bf16-representable numbers together with the dtype name "bf16"; bool is np.bool_.
"""
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_GOLDEN = os.path.join(os.path.dirname(_HERE), "tests", "golden", "promotion.json")
_P = json.load(open(_GOLDEN))

DTYPES = _P["dtypes"]  # bool i8 i16 i32 i64 u8 u16 u32 u64 f16 bf16 f32 f64
NP = {"bool": np.bool_, "i8": np.int8, "i16": np.int16, "i32": np.int32, "i64": np.int64, "u8": np.uint8,
      "u16": np.uint16, "u32": np.uint32, "u64": np.uint64, "f16": np.float16, "bf16": np.float32,
      "f32": np.float32, "f64": np.float64}
INTS = ("i8", "i16", "i32", "i64", "u8", "u16", "u32", "u64")
FLOATS = ("f16", "bf16", "f32", "f64")


def normal_out(a, b):
    return _P["normal_out"][a][b]


def float_out_binary(a, b):
    return _P["float_out_binary"][a][b]


def float_out_unary(a):
    return _P["float_out_unary"][a]


def compute_dtype(d):
    """`Intermediate`: f32 for f16/bf16 outputs, else the output type."""
    return "f32" if d in ("f16", "bf16") else d


# ---- bf16 helpers ----------------------------------------------------------------------------------
def round_bf16_from_f32(x):
    """f32 → bf16 (round to nearest even, half::bf16::from_f32), returned as f32."""
    x = np.asarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    nan = np.isnan(x)
    lsb = (u >> np.uint64(16)) & np.uint64(1)
    r = ((u + np.uint64(0x7FFF) + lsb) >> np.uint64(16)) << np.uint64(16)
    out = r.astype(np.uint32).view(np.float32).copy()
    out[nan] = np.float32(np.nan)
    return out


def _f64_to_f32_round_odd(x):
    x = np.asarray(x, dtype=np.float64)
    with np.errstate(over="ignore", invalid="ignore"):
        f = x.astype(np.float32)
    exact = (f.astype(np.float64) == x) | np.isnan(x) | np.isinf(f)
    # truncate toward zero, then force the last bit to 1 (sticky)
    toward0 = np.where(np.abs(f.astype(np.float64)) > np.abs(x), np.nextafter(f, np.float32(0)), f).astype(np.float32)
    odd = (toward0.view(np.uint32) | np.uint32(1)).view(np.float32)
    return np.where(exact, f, odd).astype(np.float32)


def round_bf16_from_f64(x):
    """f64 → bf16 with ONE rounding (half::bf16::from_f64): round-to-odd to f32, then RNE to bf16."""
    x = np.asarray(x, dtype=np.float64)
    with np.errstate(over="ignore"):
        f = x.astype(np.float32)
    # values that overflow f32 go to inf either way; round-to-odd is only needed for finite results
    ro = _f64_to_f32_round_odd(x)
    ro = np.where(np.isinf(f), f, ro)
    return round_bf16_from_f32(ro)


# ---- casts (Rust `as`) -----------------------------------------------------------------------------
def _float_to_int(x, to):
    info = np.iinfo(NP[to])
    x = np.asarray(x)
    xf = x.astype(np.float64)  # exact for f16/f32/f64 inputs
    out = np.zeros(x.shape, dtype=NP[to])
    nan = np.isnan(xf)
    lo = xf <= float(info.min)
    hi = xf >= float(info.max)
    mid = ~(nan | lo | hi)
    out[lo] = info.min
    out[hi] = info.max
    # truncation toward zero; float(info.max) may round up to 2^63/2^64, those inputs are in `hi`
    t = np.trunc(xf[mid])
    if to == "u64":
        out[mid] = t.astype(np.uint64)
    else:
        out[mid] = t.astype(np.int64).astype(NP[to])
    out[nan] = 0
    return out


def cast(x, frm, to):
    """Cast array `x` holding dtype `frm` to dtype `to` with Hpt's `Cast` semantics."""
    x = np.asarray(x, dtype=NP[frm])
    if frm == to:
        return x.copy()
    if to == "bool":
        return x != 0  # NaN != 0 → true, as in Rust
    if frm == "bool":
        if to == "bf16":
            return x.astype(np.float32)
        return x.astype(NP[to])
    if to in INTS:
        if frm in INTS:
            return x.astype(NP[to])  # two's complement wrap
        return _float_to_int(x, to)  # f16/bf16 go through f32 first: exact, same result
    # float targets
    if to == "f64":
        return x.astype(np.float64)
    if to == "f32":
        return x.astype(np.float32)  # ints round to nearest even; f64 → f32 RNE; halves exact
    via_f64 = frm in ("u32", "i64", "u64", "f64")
    if to == "f16":
        with np.errstate(over="ignore"):
            if frm == "bf16":
                return x.astype(np.float32).astype(np.float16)
            if via_f64:
                return x.astype(np.float64).astype(np.float16)  # numpy converts double→half with one rounding
            return x.astype(np.float32).astype(np.float16)
    if to == "bf16":
        if frm == "f16":
            return round_bf16_from_f32(x.astype(np.float32))
        if via_f64:
            return round_bf16_from_f64(x.astype(np.float64))
        return round_bf16_from_f32(x.astype(np.float32))
    raise ValueError((frm, to))


def to_compute(x, d):
    """widen dtype-d values to the compute type (f32 for halves)."""
    if d in ("f16", "bf16"):
        return np.asarray(x).astype(np.float32)
    return np.asarray(x, dtype=NP[d])


def from_compute(x, d):
    if d == "f16":
        with np.errstate(over="ignore"):
            return np.asarray(x, dtype=np.float32).astype(np.float16)
    if d == "bf16":
        return round_bf16_from_f32(np.asarray(x, dtype=np.float32))
    return np.asarray(x).astype(NP[d])


# ---- binary -----------------------------------------------------------------------------------------
BIT_OPS = ("bitand", "bitor", "bitxor", "shl", "shr")
CMP_OPS = ("eq", "ne", "lt", "le", "gt", "ge")


def binary_out_dtype(op, a, b):
    if op in BIT_OPS and not (a in INTS + ("bool",) and b in INTS + ("bool",)):
        return None
    o = float_out_binary(a, b) if op in ("div", "pow", "hypot") else normal_out(a, b)
    if o == "bool" and op in ("sub", "rem", "div"):
        return None
    return o


def _int_rem(a, b):
    # Rust wrapping_rem: truncated remainder (sign of the dividend); MIN % -1 = 0.  Division by zero panics in
    # the reference; the device library defines it as 0 and so does the oracle.
    a64 = a.astype(np.int64) if a.dtype != np.uint64 else a
    b64 = b.astype(np.int64) if b.dtype != np.uint64 else b
    zero = b64 == 0
    safe_b = np.where(zero, 1, b64).astype(a64.dtype)
    if a.dtype == np.uint64:
        r = a64 % safe_b
    else:
        r = np.fmod(a64, safe_b)  # C semantics for integers: truncated
    r = np.where(zero, 0, r)
    return r.astype(a.dtype)


def binary(op, x, xd, y, yd):
    """Returns (result array, result dtype name) for broadcast `x op y`."""
    od = binary_out_dtype(op, xd, yd)
    if od is None:
        raise TypeError(f"{op} unsupported for ({xd}, {yd})")
    a = to_compute(cast(x, xd, od), od)
    b = to_compute(cast(y, yd, od), od)
    a, b = np.broadcast_arrays(a, b)
    if od == "bool":
        r = {"add": a | b, "mul": a & b, "maximum": a | b, "minimum": a & b, "bitand": a & b, "bitor": a | b,
             "bitxor": a ^ b, "shl": a, "shr": a}[op]  # _bool.rs:133-163: shifts leave a bool unchanged
        return r, od
    with np.errstate(all="ignore"):
        if od in INTS:
            if op == "add":
                r = a + b
            elif op == "sub":
                r = a - b
            elif op == "mul":
                r = a * b
            elif op == "rem":
                r = _int_rem(a, b)
            elif op == "maximum":
                r = np.maximum(a, b)
            elif op == "minimum":
                r = np.minimum(a, b)
            elif op == "bitand":
                r = a & b
            elif op == "bitor":
                r = a | b
            elif op == "bitxor":
                r = a ^ b
            elif op in ("shl", "shr"):
                # wrapping_shl / wrapping_shr of `rhs as u32`: the count is taken modulo the bit width
                bits = np.dtype(NP[od]).itemsize * 8
                n = (b.astype(np.int64) if b.dtype != np.uint64 else b).astype(np.uint64) & np.uint64(bits - 1)
                n = n.astype(NP[od])
                r = np.left_shift(a, n) if op == "shl" else np.right_shift(a, n)  # >> is arithmetic for signed
            else:
                raise ValueError(op)
            return r.astype(NP[od]), od
        if op == "add":
            r = a + b
        elif op == "sub":
            r = a - b
        elif op == "mul":
            r = a * b
        elif op == "div":
            r = a / b
        elif op == "rem":
            r = np.fmod(a, b)
        elif op == "maximum":
            r = np.fmax(a, b)  # f32::max ignores NaN
        elif op == "minimum":
            r = np.fmin(a, b)
        elif op in ("pow", "hypot"):
            # correctly rounded value: evaluated in f64, rounded once (the device is allowed 2 ulp around it)
            a64, b64 = a.astype(np.float64), b.astype(np.float64)
            r = np.power(a64, b64) if op == "pow" else np.hypot(a64, b64)
            if od != "f64":
                r = r.astype(np.float32)
        else:
            raise ValueError(op)
    return from_compute(r, od), od


def compare(op, x, xd, y, yd):
    """TensorCmp: cast both sides to NormalOutPromote<L,R>, compare there; bool result (NaN is unordered)."""
    pd = normal_out(xd, yd)
    a = to_compute(cast(x, xd, pd), pd)
    b = to_compute(cast(y, yd, pd), pd)
    a, b = np.broadcast_arrays(a, b)
    with np.errstate(all="ignore"):
        r = {"eq": a == b, "ne": a != b, "lt": a < b, "le": a <= b, "gt": a > b, "ge": a >= b}[op]
    return np.asarray(r, dtype=np.bool_), "bool"


NORMAL_UNARY_OPS = ("floor", "ceil", "round", "trunc", "abs", "neg", "sign", "square", "relu", "relu6", "leaky_relu",
                    "clamp", "bitnot")


def normal_unary(op, x, xd, alpha=0.0, beta=0.0):
    """NormalUaryOps (+ bitnot): T → T, exact.  Floats follow Rust std (round half away from zero, signum(±0) = ±1,
    relu = f32::max(x, 0) so NaN → 0, clamp keeps NaN); integers wrap; unsigned neg/abs/sign and every bool op
    except neg/bitnot are the identity."""
    if op == "bitnot" and xd in FLOATS:
        raise TypeError("bitnot is integer/bool only")
    x = np.asarray(x, dtype=NP[xd])
    if xd == "bool":
        return (~x if op in ("neg", "bitnot") else x.copy()), xd
    with np.errstate(all="ignore"):
        if xd in INTS:
            t = NP[xd]
            signed = xd[0] == "i"
            al, be = _float_to_int(np.float64(alpha), xd), _float_to_int(np.float64(beta), xd)
            zero = t(0)
            if op in ("floor", "ceil", "round", "trunc"):
                r = x.copy()
            elif op == "square":
                r = x * x
            elif op == "abs":
                r = np.where(x < 0, zero - x, x) if signed else x.copy()
            elif op == "neg":
                r = (zero - x) if signed else x.copy()
            elif op == "sign":
                r = np.sign(x) if signed else x.copy()
            elif op == "relu":
                r = np.maximum(x, zero)
            elif op == "relu6":
                r = np.maximum(np.minimum(x, t(6)), zero)
            elif op == "leaky_relu":
                r = np.maximum(x, zero) + t(al) * np.minimum(x, zero)
            elif op == "clamp":
                r = np.where(x < t(al), t(al), np.where(x > t(be), t(be), x))
            elif op == "bitnot":
                r = ~x
            else:
                raise ValueError(op)
            return np.asarray(r).astype(t), xd
        v = to_compute(x, xd)
        c = v.dtype.type
        al, be = c(alpha), c(beta)
        if op == "floor":
            r = np.floor(v)
        elif op == "ceil":
            r = np.ceil(v)
        elif op == "round":
            r = np.where(np.isfinite(v), np.copysign(np.floor(np.abs(v.astype(np.float64)) + 0.5), v), v).astype(v.dtype)
        elif op == "trunc":
            r = np.trunc(v)
        elif op == "abs":
            r = np.abs(v)
        elif op == "neg":
            r = -v
        elif op == "sign":
            r = np.where(np.isnan(v), v, np.copysign(c(1), v))
        elif op == "square":
            r = v * v
        elif op == "relu":
            r = np.fmax(v, c(0))
        elif op == "relu6":
            r = np.fmin(np.fmax(v, c(0)), c(6))
        elif op == "leaky_relu":
            r = np.fmax(v, c(0)) + al * np.fmin(v, c(0))
        elif op == "clamp":
            r = np.where(v < al, al, np.where(v > be, be, v))
        else:
            raise ValueError(op)
    return from_compute(r.astype(v.dtype), xd), xd


# ---- unary ---------------------------------------------------------------------------------------------
SELU_ALPHA = 1.6732632423543772848170429916717
SELU_SCALE = 1.0507009873554804934193349852946


def _erf(x):
    from math import erf
    return np.vectorize(erf, otypes=[np.float64])(x)


def unary_f64(op, x, alpha=0.0, beta=0.0):
    """The formula of hpt-types/src/scalars/_f64.rs evaluated in float64 (the ≤2 ulp reference)."""
    x = np.asarray(x, dtype=np.float64)
    with np.errstate(all="ignore"):
        f = {
            "sin": np.sin, "cos": np.cos, "tan": np.tan, "asin": np.arcsin, "acos": np.arccos, "atan": np.arctan,
            "sinh": np.sinh, "cosh": np.cosh, "tanh": np.tanh, "asinh": np.arcsinh, "acosh": np.arccosh,
            "atanh": np.arctanh, "exp": np.exp, "exp2": np.exp2, "exp10": lambda v: np.power(10.0, v), "ln": np.log,
            "log2": np.log2, "log10": np.log10, "sqrt": np.sqrt, "cbrt": np.cbrt, "recip": lambda v: 1.0 / v,
            "erf": _erf,
            "sigmoid": lambda v: 1.0 / (1.0 + np.exp(-v)),
            "gelu": lambda v: 0.5 * v * (_erf(v * 0.7071067811865476) + 1.0),
            "elu": lambda v: np.fmax(v, 0.0) + alpha * np.fmin(np.expm1(v), 0.0),
            "selu": lambda v: beta * (np.fmax(v, 0.0) + alpha * np.fmin(np.expm1(v), 0.0)),
            "celu": lambda v: np.where(v > 0, 1.0, 0.0) * v + (1.0 - np.where(v > 0, 1.0, 0.0)) * (alpha * (np.exp(v) - 1.0)),
            "mish": lambda v: v * np.tanh(np.log(1.0 + np.exp(v))),
            "softplus": lambda v: np.log(1.0 + np.exp(v)),
            "softsign": lambda v: v / (1.0 + np.abs(v)),
            "hard_sigmoid": lambda v: np.fmax(np.fmin(v * (1.0 / 6.0) + 0.5, 1.0), 0.0),
            "hard_swish": lambda v: v * (np.fmin(np.fmax(v + 3.0, 0.0), 6.0) / 6.0),
        }[op]
        return f(x)


def unary(op, x, xd, alpha=0.0, beta=0.0):
    """(result, dtype): cast to FloatOutUnaryPromote<T>, evaluate, round once to the output dtype.
    The value is the correctly rounded one; the device is allowed 2 ulp around it."""
    od = float_out_unary(xd)
    v = to_compute(cast(x, xd, od), od).astype(np.float64)
    r = unary_f64(op, v, alpha, beta)
    with np.errstate(all="ignore"):
        if od == "f64":
            return r, od
        return from_compute(r.astype(np.float32), od), od


# ---- reductions ---------------------------------------------------------------------------------------
def process_axes(axes, ndim):
    if isinstance(axes, int):
        axes = [axes]
    out, seen = [], set()
    for a in axes:
        if a in seen:
            raise ValueError(f"axis {a} duplicated")
        seen.add(a)
        v = a + ndim if a < 0 else a
        if ndim > 0 and not 0 <= v < ndim:
            raise IndexError(f"axis {a} out of range")
        out.append(v)
    return out


def reduce_shape(shape, axes, keep_dims):
    if keep_dims:
        s = [1 if i in axes else d for i, d in enumerate(shape)]
    else:
        s = [d for i, d in enumerate(shape) if i not in axes]
    return s if s else [1]


def reduce_out_dtype(op, d):
    if op in ("sum", "max", "min", "prod", "sum_square", "reducel1", "nansum", "nanprod"):
        return d
    if op in ("all", "any"):
        return "bool"
    if op in ("mean", "logsumexp", "reducel2", "reducel3"):
        return float_out_binary(d, d)
    if op in ("argmax", "argmin"):
        return "i64"
    raise ValueError(op)


def reduce(op, x, xd, axes, keep_dims=False):
    """(result, dtype, exact).  Integer/bool/arg results are exact.  Float results are the f64-accumulated
    reference (BASELINE.json north_star) rounded once to the output dtype."""
    x = np.asarray(x, dtype=NP[xd])
    axes = tuple(process_axes(axes, x.ndim))
    od = reduce_out_dtype(op, xd)
    oshape = reduce_shape(x.shape, axes, keep_dims)
    with np.errstate(all="ignore"):
        if op in ("argmax", "argmin"):
            (ax,) = axes
            v = to_compute(x, xd)
            if xd == "bool":
                v = v.astype(np.int8)
            if v.dtype.kind == "f":
                # strict compare from ∓inf with index 0: NaN never wins; all-NaN / all-identity → 0
                fillv = -np.inf if op == "argmax" else np.inf
                v = np.where(np.isnan(v), fillv, v)
            r = (np.argmax(v, axis=ax) if op == "argmax" else np.argmin(v, axis=ax)).astype(np.int64)
            return r.reshape(oshape), od, True
        if op in ("all", "any"):
            # `x != 0` (NaN is true), AND / OR (common_reduce.rs:200-242)
            t = to_compute(x, xd) != 0
            r = np.all(t, axis=axes) if op == "all" else np.any(t, axis=axes)
            return np.asarray(r, dtype=np.bool_).reshape(oshape), od, True
        if xd == "bool":
            if op in ("sum", "max", "reducel1", "nansum"):
                r = np.any(x, axis=axes)
            elif op in ("prod", "min", "sum_square", "nanprod"):
                r = np.all(x, axis=axes) if op != "sum_square" else np.any(x, axis=axes)
            else:
                r = None
            if r is not None:
                return np.asarray(r).reshape(oshape), od, True
        if xd in INTS and op in ("reducel1", "nansum", "nanprod"):
            if op == "reducel1":
                ax_ = np.where(x < 0, NP[xd](0) - x, x).astype(NP[xd]) if xd[0] == "i" else x
                r = np.add.reduce(ax_, axis=axes, dtype=NP[xd])
            elif op == "nansum":
                r = np.add.reduce(x, axis=axes, dtype=NP[xd])
            else:
                r = np.multiply.reduce(x, axis=axes, dtype=NP[xd])
            return np.asarray(r, dtype=NP[xd]).reshape(oshape), od, True
        if xd in INTS and op in ("sum", "prod", "max", "min", "sum_square"):
            if op == "sum":
                r = np.add.reduce(x, axis=axes, dtype=NP[xd])
            elif op == "prod":
                r = np.multiply.reduce(x, axis=axes, dtype=NP[xd])
            elif op == "sum_square":
                r = np.add.reduce((x * x).astype(NP[xd]), axis=axes, dtype=NP[xd])
            elif op == "max":
                r = np.max(x, axis=axes) if x.size else np.full(oshape, np.iinfo(NP[xd]).min, NP[xd])
            else:
                r = np.min(x, axis=axes) if x.size else np.full(oshape, np.iinfo(NP[xd]).max, NP[xd])
            return np.asarray(r, dtype=NP[xd]).reshape(oshape), od, True
        # float-valued results: f64 accumulation
        if op in ("mean", "logsumexp", "reducel2"):
            v = to_compute(cast(x, xd, od), od).astype(np.float64)
        elif op == "reducel3":
            ax_, _ = normal_unary("abs", x, xd)  # |x| in T (wraps for the integer minimum), then cast
            v = to_compute(cast(ax_, xd, od), od).astype(np.float64)
        else:
            v = to_compute(x, xd).astype(np.float64)
        n = 1
        for a in axes:
            n *= x.shape[a]
        if op == "sum":
            r = v.sum(axis=axes)
        elif op == "prod":
            r = v.prod(axis=axes)
        elif op == "sum_square":
            r = (v * v).sum(axis=axes)
        elif op == "reducel1":
            r = np.abs(v).sum(axis=axes)
        elif op == "nansum":
            r = np.where(np.isnan(v), 0.0, v).sum(axis=axes)
        elif op == "nanprod":
            r = np.where(np.isnan(v), 1.0, v).prod(axis=axes)
        elif op == "reducel2":
            r = np.sqrt((v * v).sum(axis=axes))
        elif op == "reducel3":
            third = np.float64(to_compute(cast(np.array([1.0 / 3.0]), "f64", od), od)[0])  # `(1.0 / 3.0).cast()` to the output dtype
            r = np.power((v * v * v).sum(axis=axes), third)
        elif op == "mean":
            r = v.sum(axis=axes) / n
        elif op == "logsumexp":
            cd = np.float32 if compute_dtype(od) == "f32" else np.float64
            r = np.log(np.exp(v.astype(cd)).astype(np.float64).sum(axis=axes))
        elif op == "max":
            r = np.fmax.reduce(v, axis=axes, initial=-np.inf)
        elif op == "min":
            r = np.fmin.reduce(v, axis=axes, initial=np.inf)
        else:
            raise ValueError(op)
        r = np.asarray(r, dtype=np.float64).reshape(oshape)
        exact = op in ("max", "min")
        return (r if od == "f64" else from_compute(r.astype(np.float32), od)), od, exact


def reduce_f64(op, x, xd, axes):
    """Unrounded f64 reference for the tolerance check of sums (1e-6·log2 n relative)."""
    x = np.asarray(x, dtype=NP[xd])
    axes = tuple(process_axes(axes, x.ndim))
    od = reduce_out_dtype(op, xd)
    if op in ("mean", "logsumexp", "reducel2"):
        v = to_compute(cast(x, xd, od), od).astype(np.float64)
    elif op == "reducel3":
        ax_, _ = normal_unary("abs", x, xd)
        v = to_compute(cast(ax_, xd, od), od).astype(np.float64)
    else:
        v = to_compute(x, xd).astype(np.float64)
    n = 1
    for a in axes:
        n *= x.shape[a]
    with np.errstate(all="ignore"):
        if op == "sum":
            return v.sum(axis=axes)
        if op == "mean":
            return v.sum(axis=axes) / n
        if op == "sum_square":
            return (v * v).sum(axis=axes)
        if op == "prod":
            return v.prod(axis=axes)
        if op == "logsumexp":
            return np.log(np.exp(v).sum(axis=axes))
        if op == "reducel1":
            return np.abs(v).sum(axis=axes)
        if op == "nansum":
            return np.where(np.isnan(v), 0.0, v).sum(axis=axes)
        if op == "nanprod":
            return np.where(np.isnan(v), 1.0, v).prod(axis=axes)
        if op == "reducel2":
            return np.sqrt((v * v).sum(axis=axes))
        if op == "reducel3":
            third = np.float64(to_compute(cast(np.array([1.0 / 3.0]), "f64", od), od)[0])
            return np.power((v * v * v).sum(axis=axes), third)
    raise ValueError(op)


def softmax(x, xd, axis, log=False):
    """(result, dtype): max-shifted softmax evaluated in f64, rounded once to FloatOutUnaryPromote<T>."""
    od = float_out_unary(xd)
    v = to_compute(cast(x, xd, od), od).astype(np.float64)
    with np.errstate(all="ignore"):
        m = np.max(v, axis=axis, keepdims=True)
        sh = v - m
        e = np.exp(sh)
        s = e.sum(axis=axis, keepdims=True)
        r = sh - np.log(s) if log else e / s
        if od == "f64":
            return r, od
        return from_compute(r.astype(np.float32), od), od


def layernorm(x, xd, n_normalized_dims, gamma=None, beta=None, eps=1e-5):
    """(f64 reference, output dtype): hpt/src/backends/cpu/tensor_internal/normalization.rs:49-200 — over the last
    n dims, (x − mean) / sqrt(var + eps) with the population variance, then gamma·y + beta; output
    FloatOutBinaryPromote<T,T>.  Returned unrounded (f64) for a tolerance check."""
    od = float_out_binary(xd, xd)
    v = to_compute(cast(x, xd, od), od).astype(np.float64)
    axes = tuple(range(v.ndim - n_normalized_dims, v.ndim))
    mean = v.mean(axis=axes, keepdims=True)
    var = ((v - mean) ** 2).mean(axis=axes, keepdims=True)
    y = (v - mean) / np.sqrt(var + eps)
    if gamma is not None:
        y = y * np.asarray(gamma, dtype=np.float64)
    if beta is not None:
        y = y + np.asarray(beta, dtype=np.float64)
    return y, od


def mean_var(x, xd, axes):
    x = np.asarray(x, dtype=NP[xd])
    axes = tuple(process_axes(axes, x.ndim))
    od = float_out_binary(xd, xd)
    v = to_compute(cast(x, xd, od), od).astype(np.float64)
    return v.mean(axis=axes), v.var(axis=axes), od


# ---- ulp distance ----------------------------------------------------------------------------------------
def ulp_diff(a, b, d):
    """|a − b| in units of the last place of dtype d (bf16 carried as f32).  NaN==NaN and equal infs are 0."""
    if d == "bf16":
        ia = np.asarray(a, np.float32).view(np.int32).astype(np.int64) >> 16
        ib = np.asarray(b, np.float32).view(np.int32).astype(np.int64) >> 16
        bits = 16
    else:
        it = {"f16": np.int16, "f32": np.int32, "f64": np.int64}[d]
        ia = np.asarray(a, NP[d]).view(it).astype(np.int64)
        ib = np.asarray(b, NP[d]).view(it).astype(np.int64)
        bits = np.dtype(it).itemsize * 8
    sign = np.int64(1) << np.int64(bits - 1)
    ia = np.where(ia < 0, -(ia & (sign - 1)), ia)  # sign-magnitude → monotone integer line
    ib = np.where(ib < 0, -(ib & (sign - 1)), ib)
    d_ = np.abs(ia - ib)
    both_nan = np.isnan(np.asarray(a, np.float64)) & np.isnan(np.asarray(b, np.float64))
    return np.where(both_nan, 0, d_)
