"""ctypes binding of the C ABI in include/hpt_b200.h — the same symbols the Rust shim binds.

There is no fallback: if libhpt_b200.so is missing the import fails loudly (build it with
`python build.py`).  Nothing in this package computes on the CPU.
"""
import ctypes
import os
from ctypes import POINTER, Structure, byref, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_uint8, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# HPTB_LIB_VARIANT selects a tuning build (python build.py --variant NAME …), used only by tools/sweep.py
_VARIANT = os.environ.get("HPTB_LIB_VARIANT", "")
LIB_PATH = os.path.join(_HERE, "lib", f"libhpt_b200_{_VARIANT}.so" if _VARIANT else "libhpt_b200.so")
MAX_DIMS = 8

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: the CUDA extension is required (run `python build.py`); "
        "hpt_b200 has no CPU fallback")
lib = ctypes.CDLL(LIB_PATH)

# dtype enum (include/hpt_b200.h hptb_dtype)
BOOL, I8, I16, I32, I64, U8, U16, U32, U64, F16, BF16, F32, F64 = range(13)
DTYPE_NAMES = ["bool", "i8", "i16", "i32", "i64", "u8", "u16", "u32", "u64", "f16", "bf16", "f32", "f64"]
DTYPE_SIZES = [1, 1, 2, 4, 8, 1, 2, 4, 8, 2, 2, 4, 8]

BINARY_OPS = {"add": 0, "sub": 1, "mul": 2, "rem": 3, "div": 4, "maximum": 5, "minimum": 6, "pow": 7, "hypot": 8,
              "bitand": 9, "bitor": 10, "bitxor": 11, "shl": 12, "shr": 13}
CMP_OPS = {"eq": 0, "ne": 1, "lt": 2, "le": 3, "gt": 4, "ge": 5}
UNARY_OPS = {n: i for i, n in enumerate([
    "sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh",
    "exp", "exp2", "exp10", "ln", "log2", "log10", "sqrt", "cbrt", "recip", "erf", "sigmoid", "gelu",
    "selu", "elu", "celu", "mish", "softplus", "softsign", "hard_sigmoid", "hard_swish",
    # NormalUaryOps (+ bitnot): output dtype = input dtype
    "floor", "ceil", "round", "trunc", "abs", "neg", "sign", "square", "relu", "relu6", "leaky_relu", "clamp", "bitnot"])}
FLOAT_UNARY_COUNT = UNARY_OPS["floor"]
REDUCE_OPS = {"sum": 0, "mean": 1, "max": 2, "min": 3, "argmax": 4, "argmin": 5, "logsumexp": 6,
              "sum_square": 7, "prod": 8, "reducel1": 9, "nansum": 10, "nanprod": 11, "all": 12, "any": 13,
              "reducel2": 14, "reducel3": 15}
PROMOTE_NORMAL, PROMOTE_FLOAT_BINARY, PROMOTE_FLOAT_UNARY = 0, 1, 2

STATUS_NAMES = {0: "OK", 1: "SHAPE", 2: "DTYPE", 3: "AXIS", 4: "INVALID", 5: "CUDA", 6: "OOM", 7: "NCCL",
                8: "UNSUPPORTED"}


class HptbTensor(Structure):
    _fields_ = [("data", c_void_p), ("dtype", c_int32), ("ndim", c_int32),
                ("shape", c_int64 * MAX_DIMS), ("strides", c_int64 * MAX_DIMS)]


class HptbAllocStats(Structure):
    _fields_ = [(n, c_uint64) for n in ("bytes_in_use", "bytes_cached", "bytes_reserved_peak", "n_alloc",
                                        "n_cache_hit", "n_device_malloc", "n_device_free")]


class HptbCollapsePlan(Structure):
    _fields_ = [("ndim", c_int32), ("launch_class", c_int32), ("n_operands", c_int32), ("reserved", c_int32),
                ("shape", c_int64 * MAX_DIMS), ("strides", (c_int64 * MAX_DIMS) * 4), ("reduced", c_uint8 * MAX_DIMS)]


class HptbReduceRoute(Structure):
    _fields_ = [("kind", c_int32), ("reserved", c_int32), ("head", c_int64), ("body", c_int64), ("tail", c_int64),
                ("scratch_strides", c_int64 * MAX_DIMS)]


ROUTES = ["direct", "peel", "two_step", "peel_raw"]


class HptbShardPlan(Structure):
    _fields_ = [("crosses", c_int32), ("collective", c_int32), ("pre_exp", c_int32), ("post_ln", c_int32),
                ("global_count", c_int32), ("post_root", c_int32)]


COLLECTIVES = ["none", "allreduce_sum", "allreduce_prod", "allreduce_max", "allreduce_min", "allgather_arg"]


class HptError(RuntimeError):
    """A non-zero hptb_status.  `.status` is the code; shape/axis/dtype errors mirror Hpt's
    TensorError::{Shape, Param, Kernel} (hpt-common/src/error/*.rs)."""

    def __init__(self, status, msg):
        super().__init__(f"[HPTB_ERR_{STATUS_NAMES.get(status, status)}] {msg}")
        self.status = status


# every symbol declared in include/hpt_b200.h, with its signature
_T = POINTER(HptbTensor)
SIGNATURES = {
    "hptb_version": (c_int, []),
    "hptb_kernel_launches": (c_uint64, []),
    "hptb_last_error": (c_char_p, []),
    "hptb_dtype_size": (c_size_t, [c_int]),
    "hptb_dtype_name": (c_char_p, [c_int]),
    "hptb_ctx_create": (c_int, [c_int, POINTER(c_void_p)]),
    "hptb_ctx_destroy": (c_int, [c_void_p]),
    "hptb_ctx_device": (c_int, [c_void_p, POINTER(c_int)]),
    "hptb_ctx_sm_count": (c_int, [c_void_p, POINTER(c_int)]),
    "hptb_stream_sync": (c_int, [c_void_p, c_void_p]),
    "hptb_alloc": (c_int, [c_void_p, c_size_t, POINTER(c_void_p), c_void_p]),
    "hptb_free": (c_int, [c_void_p, c_void_p, c_void_p]),
    "hptb_record_stream": (c_int, [c_void_p, c_void_p, c_void_p]),
    "hptb_empty_cache": (c_int, [c_void_p]),
    "hptb_alloc_get_stats": (c_int, [c_void_p, POINTER(HptbAllocStats)]),
    "hptb_alloc_selftest": (c_int, []),
    "hptb_memcpy_h2d": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "hptb_memcpy_d2h": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "hptb_memcpy_d2d": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "hptb_memcpy_d2h_async": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "hptb_stream_wait_stream": (c_int, [c_void_p, c_void_p, c_void_p]),
    "hptb_host_alloc_pinned": (c_int, [c_size_t, POINTER(c_void_p)]),
    "hptb_host_free_pinned": (c_int, [c_void_p]),
    "hptb_promote": (c_int, [c_int, c_int, c_int]),
    "hptb_binary_out_dtype": (c_int, [c_int, c_int, c_int]),
    "hptb_unary_out_dtype": (c_int, [c_int, c_int]),
    "hptb_reduce_out_dtype": (c_int, [c_int, c_int]),
    "hptb_broadcast_shape": (c_int, [POINTER(c_int64), c_int, POINTER(c_int64), c_int, POINTER(c_int64), POINTER(c_int)]),
    "hptb_process_axes": (c_int, [POINTER(c_int64), c_int, c_int, POINTER(c_int32)]),
    "hptb_reduce_shape": (c_int, [POINTER(c_int64), c_int, POINTER(c_int32), c_int, c_int, POINTER(c_int64), POINTER(c_int)]),
    "hptb_collapse": (c_int, [POINTER(_T), c_int, POINTER(c_uint8), POINTER(HptbCollapsePlan)]),
    "hptb_reduce_route": (c_int, [c_int, _T, POINTER(c_int32), c_int, _T, c_int, POINTER(HptbReduceRoute)]),
    "hptb_binary": (c_int, [c_void_p, c_int, _T, _T, _T, c_void_p]),
    "hptb_compare": (c_int, [c_void_p, c_int, _T, _T, _T, c_void_p]),
    "hptb_unary": (c_int, [c_void_p, c_int, _T, _T, c_double, c_double, c_void_p]),
    "hptb_reduce": (c_int, [c_void_p, c_int, _T, POINTER(c_int32), c_int, _T, c_int, c_void_p]),
    "hptb_binary_reduce": (c_int, [c_void_p, c_int, c_int, _T, _T, POINTER(c_int32), c_int, _T, c_int, c_void_p]),
    "hptb_mean_var": (c_int, [c_void_p, _T, POINTER(c_int32), c_int, _T, _T, c_void_p]),
    "hptb_softmax": (c_int, [c_void_p, _T, c_int, c_int, _T, c_void_p]),
    "hptb_layernorm": (c_int, [c_void_p, _T, c_int, _T, _T, c_double, _T, c_void_p]),
    "hptb_copy": (c_int, [c_void_p, _T, _T, c_void_p]),
    "hptb_fill": (c_int, [c_void_p, _T, c_void_p, c_void_p]),
    "hptb_arange": (c_int, [c_void_p, _T, c_void_p, c_void_p, c_void_p]),
    "hptb_eye": (c_int, [c_void_p, _T, c_int64, c_void_p]),
    "hptb_comm_unique_id": (c_int, [c_void_p]),
    "hptb_comm_init_rank": (c_int, [c_void_p, c_int, c_int, c_void_p, POINTER(c_void_p)]),
    "hptb_comm_init_local_group": (c_int, [c_void_p, c_int, POINTER(c_void_p)]),
    "hptb_comm_destroy": (c_int, [c_void_p]),
    "hptb_comm_uses_peer_memory": (c_int, [c_void_p]),
    "hptb_shard_bounds": (c_int, [c_int64, c_int, c_int, POINTER(c_int64), POINTER(c_int64)]),
    "hptb_shard_plan_reduce": (c_int, [c_int, POINTER(c_int32), c_int, c_int, c_int, POINTER(HptbShardPlan)]),
    "hptb_allreduce": (c_int, [c_void_p, c_int, _T, c_void_p]),
    "hptb_reduce_sharded": (c_int, [c_void_p, c_int, _T, POINTER(c_int32), c_int, c_int, c_int64, c_int64, _T, c_void_p]),
}

MISSING = []
for _name, (_res, _args) in SIGNATURES.items():
    try:
        _f = getattr(lib, _name)
    except AttributeError:
        MISSING.append(_name)
        continue
    _f.restype = _res
    _f.argtypes = _args


def check(status):
    if status != 0:
        raise HptError(status, lib.hptb_last_error().decode("utf-8", "replace"))


def make_tensor(ptr, dtype, shape, strides):
    t = HptbTensor()
    t.data = ptr
    t.dtype = dtype
    t.ndim = len(shape)
    if len(shape) > MAX_DIMS:
        raise HptError(4, f"ndim {len(shape)} exceeds {MAX_DIMS}")
    for i, (s, st) in enumerate(zip(shape, strides)):
        t.shape[i] = s
        t.strides[i] = st
    return t
