"""Host-side mirror of Hpt's `Tensor<T, Cuda, DEVICE>` for the hot path, over the C ABI.

The Rust toolchain is absent from the build image, so this Python class plays the role of the Rust
trait impls in hpt/src/backends/cuda/tensor_{external,internal}/*.rs: same method names, argument
meaning and error behaviour (`add_`, `sin_`, `sum(axes, keep_dims)`, `sum_(axes, keep_dims, init_out, out)`,
`argmax(axis, keep_dims)`, `softmax(axis)`, `to_cuda`, `to_cpu`, `permute`, `t`, `slice`, `contiguous`,
`astype` …).  It only does what the Rust host code does around a kernel launch — argument checks,
output-shape computation, allocation of `out` — and calls the same C entry points the Rust shim
binds.  All arithmetic happens in libhpt_b200.so on the GPU; torch CPU tensors are used purely as the
host-side container for `to_cuda` / `to_cpu` (numpy has no bf16).
"""
import ctypes
from ctypes import byref, c_int, c_int32, c_int64, c_void_p

import torch

from . import _ffi
from ._ffi import HptError, check, lib, make_tensor

_TORCH_DTYPES = [torch.bool, torch.int8, torch.int16, torch.int32, torch.int64, torch.uint8, torch.uint16,
                 torch.uint32, torch.uint64, torch.float16, torch.bfloat16, torch.float32, torch.float64]
_TORCH_TO_ENUM = {d: i for i, d in enumerate(_TORCH_DTYPES)}

_contexts = {}
_default_stream = [None]


def set_stream(stream):
    """Set the cudaStream_t (int handle or None = legacy default stream) used when a call gets no explicit
    `stream`.  The Rust shim passes its device's stream the same way (one stream per CudaDevice)."""
    _default_stream[0] = stream


def get_stream():
    return _default_stream[0]


def _s(stream):
    return _default_stream[0] if stream is None else stream


class Context:
    """One per device: wraps hptb_ctx (stream-ordered pool, device properties)."""

    def __init__(self, device):
        h = c_void_p()
        check(lib.hptb_ctx_create(int(device), byref(h)))
        self.handle = h
        self.device = int(device)

    @property
    def sm_count(self):
        n = c_int()
        check(lib.hptb_ctx_sm_count(self.handle, byref(n)))
        return n.value

    def synchronize(self, stream=None):
        check(lib.hptb_stream_sync(self.handle, _s(stream)))

    def empty_cache(self):
        check(lib.hptb_empty_cache(self.handle))

    def alloc_stats(self):
        s = _ffi.HptbAllocStats()
        check(lib.hptb_alloc_get_stats(self.handle, byref(s)))
        return {n: getattr(s, n) for n, _ in s._fields_}


def context(device=0):
    ctx = _contexts.get(device)
    if ctx is None:
        ctx = _contexts[device] = Context(device)
    return ctx


class _Storage:
    """Device allocation owned by the pool; dropped → hptb_free (mirrors Drop for _Tensor,
    hpt/src/tensor_base.rs:30-44 → Allocator::deallocate)."""

    def __init__(self, ctx, nbytes, stream=None):
        self.ctx = ctx
        p = c_void_p()
        check(lib.hptb_alloc(ctx.handle, max(int(nbytes), 1), byref(p), _s(stream)))
        self.ptr = p.value
        self.nbytes = nbytes
        self.stream = _s(stream)

    def used_on(self, stream):
        """The block is consumed by work enqueued on `stream`: if that is not the stream it will be freed on, tell the
        allocator (hptb_record_stream) so the block is not reused before that work has completed."""
        if stream != self.stream and self.ptr:
            check(lib.hptb_record_stream(self.ctx.handle, c_void_p(self.ptr), stream))

    def __del__(self):
        try:
            if self.ptr and lib is not None:
                lib.hptb_free(self.ctx.handle, c_void_p(self.ptr), self.stream)
        except Exception:
            pass


class _Borrowed:
    """Memory owned by someone else (e.g. a torch CUDA tensor used by the benchmark plumbing)."""

    def __init__(self, ctx, ptr, keepalive=None):
        self.ctx, self.ptr, self.keepalive, self.stream = ctx, ptr, keepalive, None

    def used_on(self, stream):
        pass


def _on(stream, *tensors):
    """Resolve the stream of a call and note it on every tensor the call touches (cross-stream reuse safety)."""
    s = _s(stream)
    for t in tensors:
        if t is not None:
            t.storage.used_on(s)
    return s


def _contig_strides(shape):
    st, acc = [], 1
    for d in reversed(shape):
        st.append(acc)
        acc *= d
    return tuple(reversed(st))


def _numel(shape):
    n = 1
    for d in shape:
        n *= d
    return n


def _axes_list(axes):
    if isinstance(axes, int):
        return [axes]
    return [int(a) for a in axes]


class Tensor:
    def __init__(self, storage, ptr, dtype, shape, strides):
        self.storage = storage
        self.ptr = ptr
        self.dtype = int(dtype)
        self.shape = tuple(int(s) for s in shape)
        self.strides = tuple(int(s) for s in strides)
        self._cstruct = None  # hptb_tensor image, built once (a Tensor's pointer, shape and strides never change)

    # ---- creation / transfer -----------------------------------------------------------------
    @staticmethod
    def empty(shape, dtype, device=0, stream=None):
        ctx = context(device)
        shape = tuple(int(s) for s in shape)
        if any(s < 0 for s in shape):
            raise HptError(1, "negative extent")
        st = _Storage(ctx, _numel(shape) * _ffi.DTYPE_SIZES[dtype], stream)
        return Tensor(st, st.ptr, dtype, shape, _contig_strides(shape))

    # ---- TensorCreator (hpt-traits/src/ops/creation.rs; CPU semantics normal_creation.rs:34-234) ------------------
    @staticmethod
    def _scalar(value, dtype):
        """one host element of `dtype` (Rust `as` from the Python number)"""
        if dtype in (_ffi.U16, _ffi.U32, _ffi.U64):
            return torch.tensor([int(value)], dtype=torch.int64).to(_TORCH_DTYPES[dtype])
        if dtype in (_ffi.F16, _ffi.BF16, _ffi.F32, _ffi.F64):
            return torch.tensor([float(value)], dtype=torch.float64).to(_TORCH_DTYPES[dtype])
        return torch.tensor([value]).to(_TORCH_DTYPES[dtype])

    @staticmethod
    def full(val, shape, dtype, device=0, stream=None):
        return Tensor.empty(shape, dtype, device, stream).fill_(val, stream)

    @staticmethod
    def zeros(shape, dtype, device=0, stream=None): return Tensor.full(0, shape, dtype, device, stream)

    @staticmethod
    def ones(shape, dtype, device=0, stream=None): return Tensor.full(1, shape, dtype, device, stream)

    def empty_like(self): return Tensor.empty(self.shape, self.dtype, self.ctx.device)
    def zeros_like(self): return Tensor.zeros(self.shape, self.dtype, self.ctx.device)
    def ones_like(self): return Tensor.ones(self.shape, self.dtype, self.ctx.device)
    def full_like(self, val): return Tensor.full(val, self.shape, self.dtype, self.ctx.device)

    @staticmethod
    def _arange_into(n, start, step, dtype, device, stream):
        out = Tensor.empty((max(int(n), 0),), dtype, device, stream)
        a, b = Tensor._scalar(start, dtype), Tensor._scalar(step, dtype)
        check(lib.hptb_arange(out.ctx.handle, byref(out._c()), c_void_p(a.data_ptr()), c_void_p(b.data_ptr()), _s(stream)))
        return out

    @staticmethod
    def arange(start, end, dtype, device=0, stream=None):
        """`Tensor::<T>::arange(start, end)`: size = end as i64 − start as i64 (empty if ≤ 0), x[i] = start + i."""
        return Tensor._arange_into(int(end) - int(start), start, 1, dtype, device, stream)

    @staticmethod
    def arange_step(start, end, step, dtype, device=0, stream=None):
        """size = floor((end − start) / step) + 1 in f64, as normal_creation.rs:163-185 computes it."""
        import math
        s, e, st = float(start), float(end), float(step)
        n = math.floor((e - s) / st) + 1 if st > 0 else math.floor((s - e) / (-st)) + 1
        return Tensor._arange_into(n, start, step, dtype, device, stream)

    @staticmethod
    def linspace(start, end, num, include_end, dtype, device=0, stream=None):
        """step = (end − start) / (num − 1 | num) in f64, cast to T; x[i] = start + T(i)·step, last = end if included."""
        n = int(num)
        step = (float(end) - float(start)) / ((n - 1.0) if include_end else float(n)) if n > (1 if include_end else 0) else 0.0
        out = Tensor._arange_into(n, start, step, dtype, device, stream)
        if include_end and n > 0:
            out[n - 1:n].fill_(end, stream)
        return out

    @staticmethod
    def eye(n, m, k, dtype, device=0, stream=None):
        out = Tensor.empty((int(n), int(m)), dtype, device, stream)
        check(lib.hptb_eye(out.ctx.handle, byref(out._c()), int(k), _s(stream)))
        return out

    @staticmethod
    def identity(n, dtype, device=0, stream=None): return Tensor.eye(n, n, 0, dtype, device, stream)

    @staticmethod
    def to_cuda(host, device=0, stream=None, sync=True):
        """`cpu_tensor.to_cuda::<DEVICE>()` (hpt/src/backends/cpu/tensor_impls.rs:295-313).
        sync=False (pinned sources only): the upload is only ordered on `stream`."""
        if not isinstance(host, torch.Tensor):
            host = torch.as_tensor(host)
        if host.dtype not in _TORCH_TO_ENUM:
            raise HptError(2, f"unsupported host dtype {host.dtype}")
        host = host.contiguous()
        t = Tensor.empty(tuple(host.shape), _TORCH_TO_ENUM[host.dtype], device, stream)
        nbytes = host.numel() * host.element_size()
        if nbytes:
            check(lib.hptb_memcpy_h2d(t.ctx.handle, c_void_p(t.ptr), c_void_p(host.data_ptr()), nbytes, _s(stream)))
            # pageable source: make the staging complete before the host tensor can be mutated
            if sync or not host.is_pinned():
                t.ctx.synchronize(stream)
        return t

    @staticmethod
    def from_device_ptr(ptr, dtype, shape, strides=None, device=0, keepalive=None):
        shape = tuple(shape)
        return Tensor(_Borrowed(context(device), ptr, keepalive), ptr, dtype, shape,
                      strides if strides is not None else _contig_strides(shape))

    def to_cpu(self, stream=None, out=None, sync=True):
        """`to_cpu::<0>()` (hpt/src/backends/cuda/tensor_impls.rs:142-169): views are gathered first.
        `out` may be a preallocated (e.g. pinned) contiguous host tensor of the right shape and dtype."""
        src = self if self.is_contiguous() else self.contiguous(stream)
        if out is not None:
            if tuple(out.shape) != self.shape or out.dtype != _TORCH_DTYPES[self.dtype] or not out.is_contiguous():
                raise HptError(1, "to_cpu: out must be a contiguous host tensor of the same shape and dtype")
            host = out
        else:
            host = torch.empty(self.shape, dtype=_TORCH_DTYPES[self.dtype])
        nbytes = host.numel() * host.element_size()
        if nbytes:
            if sync or out is None or not host.is_pinned():
                check(lib.hptb_memcpy_d2h(self.ctx.handle, c_void_p(host.data_ptr()), c_void_p(src.ptr), nbytes, _on(stream, src)))
            else:  # pinned destination, caller synchronises the stream before reading `out`
                check(lib.hptb_memcpy_d2h_async(self.ctx.handle, c_void_p(host.data_ptr()), c_void_p(src.ptr), nbytes, _on(stream, src)))
                host._hptb_src = src  # keep the device buffer alive until the caller has synchronised
        return host

    # ---- metadata ---------------------------------------------------------------------------------
    @property
    def ctx(self):
        return self.storage.ctx

    @property
    def ndim(self):
        return len(self.shape)

    def size(self):
        return _numel(self.shape)

    def is_contiguous(self):
        # Layout::is_contiguous (hpt-common/src/layout/layout_utils.rs:363-375)
        exp = 1
        for d, s in zip(reversed(self.shape), reversed(self.strides)):
            if d == 0:
                continue
            if s != exp:
                return False
            exp *= d
        return True

    def _c(self):
        c = self._cstruct
        if c is None:
            c = self._cstruct = make_tensor(self.ptr, self.dtype, self.shape, self.strides)
        return c

    def __repr__(self):
        return f"Tensor<{_ffi.DTYPE_NAMES[self.dtype]}, Cuda, {self.ctx.device}>(shape={self.shape}, strides={self.strides})"

    # ---- views (pure host metadata, as in hpt/src/backends/common/shape_manipulate.rs) ---------------
    def _view(self, shape, strides, offset=0):
        return Tensor(self.storage, self.ptr + offset * _ffi.DTYPE_SIZES[self.dtype], self.dtype, shape, strides)

    def permute(self, axes):
        axes = [a + self.ndim if a < 0 else a for a in _axes_list(axes)]
        if sorted(axes) != list(range(self.ndim)):
            raise HptError(3, f"permute axes {axes} are not a permutation of 0..{self.ndim}")
        return self._view([self.shape[a] for a in axes], [self.strides[a] for a in axes])

    def transpose(self, a, b):
        ax = list(range(self.ndim))
        ax[a], ax[b] = ax[b], ax[a]
        return self.permute(ax)

    def t(self):
        if self.ndim < 2:
            return self
        return self.transpose(-2 % self.ndim, -1 % self.ndim)

    def reshape(self, shape):
        shape = list(shape)
        if -1 in shape:
            i = shape.index(-1)
            rest = _numel([s for s in shape if s != -1])
            shape[i] = self.size() // max(rest, 1)
        if _numel(shape) != self.size():
            raise HptError(1, f"cannot reshape {self.shape} to {tuple(shape)}")
        src = self if self.is_contiguous() else self.contiguous()
        return Tensor(src.storage, src.ptr, self.dtype, shape, _contig_strides(shape))

    def __getitem__(self, idx):
        """`slice!(a[lo:hi:step, …])`: python slices / ints per dim."""
        if not isinstance(idx, tuple):
            idx = (idx,)
        shape, strides, off = [], [], 0
        for d in range(self.ndim):
            if d < len(idx):
                ix = idx[d]
                if isinstance(ix, int):
                    if ix < 0:
                        ix += self.shape[d]
                    if not 0 <= ix < self.shape[d]:
                        raise HptError(1, f"index {ix} out of range for dim {d}")
                    off += ix * self.strides[d]
                    continue
                lo, hi, step = ix.indices(self.shape[d])
                n = len(range(lo, hi, step))
                off += lo * self.strides[d]
                shape.append(n)
                strides.append(self.strides[d] * step)
            else:
                shape.append(self.shape[d])
                strides.append(self.strides[d])
        return self._view(shape, strides, off)

    def expand(self, shape):
        shape = tuple(shape)
        nd = len(shape)
        strides = []
        for i in range(nd):
            j = i - (nd - self.ndim)
            if j < 0 or (self.shape[j] == 1 and shape[i] != 1):
                strides.append(0)
            elif self.shape[j] == shape[i]:
                strides.append(self.strides[j])
            else:
                raise HptError(1, f"cannot expand {self.shape} to {shape}")
        return self._view(shape, strides)

    # ---- copy / cast ------------------------------------------------------------------------------
    def contiguous(self, stream=None):
        out = Tensor.empty(self.shape, self.dtype, self.ctx.device, stream)
        check(lib.hptb_copy(self.ctx.handle, byref(self._c()), byref(out._c()), _on(stream, self, out)))
        return out

    def astype(self, dtype, stream=None):
        out = Tensor.empty(self.shape, dtype, self.ctx.device, stream)
        check(lib.hptb_copy(self.ctx.handle, byref(self._c()), byref(out._c()), _on(stream, self, out)))
        return out

    def fill_(self, value, stream=None):
        host = Tensor._scalar(value, self.dtype)
        check(lib.hptb_fill(self.ctx.handle, byref(self._c()), c_void_p(host.data_ptr()), _on(stream, self)))
        return self

    # ---- NormalBinOps / std::ops (hpt/src/backends/cuda/std_ops.rs, tensor_external/binary.rs) ----------
    def _binary(self, name, rhs, out=None, stream=None):
        op = _ffi.BINARY_OPS[name]
        if not isinstance(rhs, Tensor):
            # tensor ⊕ scalar: the reference wraps the scalar in a 1-element tensor (std_ops.rs:204-224)
            if isinstance(rhs, bool):
                host = torch.tensor([rhs])
            elif isinstance(rhs, int):
                host = torch.tensor([rhs], dtype=torch.int64)
            else:
                host = torch.tensor([rhs], dtype=torch.float64)
            rhs = Tensor.to_cuda(host, self.ctx.device, stream)
        odt = lib.hptb_binary_out_dtype(op, self.dtype, rhs.dtype)
        if odt < 0:
            raise HptError(2, f"{name} is not supported for ({_ffi.DTYPE_NAMES[self.dtype]}, {_ffi.DTYPE_NAMES[rhs.dtype]})")
        bshape = (c_int64 * _ffi.MAX_DIMS)()
        bn = c_int()
        check(lib.hptb_broadcast_shape((c_int64 * max(self.ndim, 1))(*self.shape), self.ndim,
                                       (c_int64 * max(rhs.ndim, 1))(*rhs.shape), rhs.ndim, bshape, byref(bn)))
        oshape = tuple(bshape[i] for i in range(bn.value))
        if out is None:
            out = Tensor.empty(oshape, odt, self.ctx.device, stream)
        check(lib.hptb_binary(self.ctx.handle, op, byref(self._c()), byref(rhs._c()), byref(out._c()), _on(stream, self, rhs, out)))
        return out

    def add_(self, rhs, out, stream=None): return self._binary("add", rhs, out, stream)
    def sub_(self, rhs, out, stream=None): return self._binary("sub", rhs, out, stream)
    def mul_(self, rhs, out, stream=None): return self._binary("mul", rhs, out, stream)
    def rem_(self, rhs, out, stream=None): return self._binary("rem", rhs, out, stream)
    def div_(self, rhs, out, stream=None): return self._binary("div", rhs, out, stream)
    def __add__(self, rhs): return self._binary("add", rhs)
    def __sub__(self, rhs): return self._binary("sub", rhs)
    def __mul__(self, rhs): return self._binary("mul", rhs)
    def __mod__(self, rhs): return self._binary("rem", rhs)
    def __truediv__(self, rhs): return self._binary("div", rhs)
    def maximum(self, rhs): return self._binary("maximum", rhs)
    def minimum(self, rhs): return self._binary("minimum", rhs)
    # FloatBinOps (hpt-traits/src/ops/binary.rs:94-188)
    def pow(self, rhs): return self._binary("pow", rhs)
    def pow_(self, rhs, out, stream=None): return self._binary("pow", rhs, out, stream)
    def hypot(self, rhs): return self._binary("hypot", rhs)
    def hypot_(self, rhs, out, stream=None): return self._binary("hypot", rhs, out, stream)
    # BitWiseOut / std::ops::{BitAnd, BitOr, BitXor, Shl, Shr, Not} (hpt/src/backends/cuda/std_ops.rs)
    def __and__(self, rhs): return self._binary("bitand", rhs)
    def __or__(self, rhs): return self._binary("bitor", rhs)
    def __xor__(self, rhs): return self._binary("bitxor", rhs)
    def __lshift__(self, rhs): return self._binary("shl", rhs)
    def __rshift__(self, rhs): return self._binary("shr", rhs)
    def __invert__(self): return self._unary("bitnot")
    def __neg__(self): return self._unary("neg")

    # ---- TensorCmp (hpt/src/backends/cuda/tensor_external/cmp.rs) ------------------------------------------
    def _compare(self, name, rhs, stream=None):
        op = _ffi.CMP_OPS[name]
        if not isinstance(rhs, Tensor):
            raise HptError(4, "tensor_<cmp> takes a tensor on the right-hand side")
        bshape = (c_int64 * _ffi.MAX_DIMS)()
        bn = c_int()
        check(lib.hptb_broadcast_shape((c_int64 * max(self.ndim, 1))(*self.shape), self.ndim,
                                       (c_int64 * max(rhs.ndim, 1))(*rhs.shape), rhs.ndim, bshape, byref(bn)))
        out = Tensor.empty(tuple(bshape[i] for i in range(bn.value)), _ffi.BOOL, self.ctx.device, stream)
        check(lib.hptb_compare(self.ctx.handle, op, byref(self._c()), byref(rhs._c()), byref(out._c()), _on(stream, self, rhs, out)))
        return out

    def tensor_eq(self, rhs): return self._compare("eq", rhs)
    def tensor_neq(self, rhs): return self._compare("ne", rhs)
    def tensor_lt(self, rhs): return self._compare("lt", rhs)
    def tensor_le(self, rhs): return self._compare("le", rhs)
    def tensor_gt(self, rhs): return self._compare("gt", rhs)
    def tensor_ge(self, rhs): return self._compare("ge", rhs)

    # ---- FloatUnaryOps (hpt/src/backends/cuda/tensor_internal/float_out_unary.rs) ---------------------
    def _unary(self, name, out=None, alpha=0.0, beta=0.0, stream=None):
        op = _ffi.UNARY_OPS[name]
        odt = lib.hptb_unary_out_dtype(op, self.dtype)
        if odt < 0:
            raise HptError(2, f"{name} is not supported for {_ffi.DTYPE_NAMES[self.dtype]}")
        if out is None:
            out = Tensor.empty(self.shape, odt, self.ctx.device, stream)
        check(lib.hptb_unary(self.ctx.handle, op, byref(self._c()), byref(out._c()), float(alpha), float(beta), _on(stream, self, out)))
        return out

    def selu(self, out=None, stream=None):
        # constants of float_out_unary.rs:442-453
        return self._unary("selu", out, 1.6732632423543772848170429916717, 1.0507009873554804934193349852946, stream)

    def elu(self, alpha, out=None, stream=None): return self._unary("elu", out, alpha, 0.0, stream)
    # NormalUaryOps with parameters (hpt-traits/src/ops/unary.rs:720-826)
    def leaky_relu(self, alpha, out=None, stream=None): return self._unary("leaky_relu", out, alpha, 0.0, stream)
    def clamp(self, min, max, out=None, stream=None): return self._unary("clamp", out, min, max, stream)
    def celu(self, alpha, out=None, stream=None): return self._unary("celu", out, alpha, 0.0, stream)

    # ---- reductions (hpt/src/backends/cuda/tensor_internal/{common_reduce,arg_reduce}.rs) ---------------
    def _reduce(self, name, axes, keep_dims=False, init_out=True, out=None, stream=None):
        op = _ffi.REDUCE_OPS[name]
        ax_in = _axes_list(axes)
        ax = (c_int32 * max(len(ax_in), 1))()
        check(lib.hptb_process_axes((c_int64 * max(len(ax_in), 1))(*ax_in), len(ax_in), self.ndim, ax))
        if name in ("argmax", "argmin") and len(ax_in) != 1:
            raise HptError(3, f"{name} takes exactly one axis")
        odt = lib.hptb_reduce_out_dtype(op, self.dtype)
        if odt < 0:
            raise HptError(2, f"{name} is not supported for {_ffi.DTYPE_NAMES[self.dtype]}")
        oshape = (c_int64 * _ffi.MAX_DIMS)()
        on = c_int()
        shp = (c_int64 * max(self.ndim, 1))(*self.shape)
        check(lib.hptb_reduce_shape(shp, self.ndim, ax, len(ax_in), 0, oshape, byref(on)))
        red_shape = tuple(oshape[i] for i in range(on.value))
        if out is None:
            res = Tensor.empty(red_shape, odt, self.ctx.device, stream)
        else:
            if _numel(out.shape) != _numel(red_shape) or out.dtype != odt:
                raise HptError(1, f"out has shape {out.shape}/{_ffi.DTYPE_NAMES[out.dtype]}, expected {red_shape}/{_ffi.DTYPE_NAMES[odt]}")
            res = out if out.shape == red_shape else Tensor(out.storage, out.ptr, out.dtype, red_shape, _contig_strides(red_shape))
        check(lib.hptb_reduce(self.ctx.handle, op, byref(self._c()), ax, len(ax_in), byref(res._c()),
                              1 if init_out else 0, _on(stream, self, res)))
        if keep_dims:
            check(lib.hptb_reduce_shape(shp, self.ndim, ax, len(ax_in), 1, oshape, byref(on)))
            ks = tuple(oshape[i] for i in range(on.value))
            res = Tensor(res.storage, res.ptr, res.dtype, ks, _contig_strides(ks))
        return res

    def sum(self, axes, keep_dims=False): return self._reduce("sum", axes, keep_dims)
    def sum_(self, axes, keep_dims, init_out, out): return self._reduce("sum", axes, keep_dims, init_out, out)
    def prod(self, axes, keep_dims=False): return self._reduce("prod", axes, keep_dims)
    def mean(self, axes, keep_dims=False): return self._reduce("mean", axes, keep_dims)
    def max(self, axes, keep_dims=False): return self._reduce("max", axes, keep_dims)
    def min(self, axes, keep_dims=False): return self._reduce("min", axes, keep_dims)
    def argmax(self, axis, keep_dims=False): return self._reduce("argmax", axis, keep_dims)
    def argmin(self, axis, keep_dims=False): return self._reduce("argmin", axis, keep_dims)
    def logsumexp(self, axes, keep_dims=False): return self._reduce("logsumexp", axes, keep_dims)
    def sum_square(self, axes, keep_dims=False): return self._reduce("sum_square", axes, keep_dims)
    def reducel1(self, axes, keep_dims=False): return self._reduce("reducel1", axes, keep_dims)
    def reducel2(self, axes, keep_dims=False): return self._reduce("reducel2", axes, keep_dims)
    def reducel3(self, axes, keep_dims=False): return self._reduce("reducel3", axes, keep_dims)
    def nansum(self, axes, keep_dims=False): return self._reduce("nansum", axes, keep_dims)
    def nansum_(self, axes, keep_dims, init_out, out): return self._reduce("nansum", axes, keep_dims, init_out, out)
    def nanprod(self, axes, keep_dims=False): return self._reduce("nanprod", axes, keep_dims)
    def all(self, axes, keep_dims=False): return self._reduce("all", axes, keep_dims)
    def any(self, axes, keep_dims=False): return self._reduce("any", axes, keep_dims)

    def binary_reduce(self, bin_name, rhs, red_name, axes, keep_dims=False, out=None, stream=None):
        """Extension: `reduce(self ⊕ rhs, axes)` in one pass (hptb_binary_reduce); same result as `(self ⊕ rhs).red(axes)`."""
        bop, rop = _ffi.BINARY_OPS[bin_name], _ffi.REDUCE_OPS[red_name]
        mid = lib.hptb_binary_out_dtype(bop, self.dtype, rhs.dtype)
        if mid < 0:
            raise HptError(2, f"{bin_name} is not supported for ({_ffi.DTYPE_NAMES[self.dtype]}, {_ffi.DTYPE_NAMES[rhs.dtype]})")
        odt = lib.hptb_reduce_out_dtype(rop, mid)
        bshape = (c_int64 * _ffi.MAX_DIMS)()
        bn = c_int()
        check(lib.hptb_broadcast_shape((c_int64 * max(self.ndim, 1))(*self.shape), self.ndim,
                                       (c_int64 * max(rhs.ndim, 1))(*rhs.shape), rhs.ndim, bshape, byref(bn)))
        ax_in = _axes_list(axes)
        ax = (c_int32 * max(len(ax_in), 1))()
        check(lib.hptb_process_axes((c_int64 * max(len(ax_in), 1))(*ax_in), len(ax_in), bn.value, ax))
        oshape = (c_int64 * _ffi.MAX_DIMS)()
        on = c_int()
        check(lib.hptb_reduce_shape(bshape, bn.value, ax, len(ax_in), 0, oshape, byref(on)))
        red_shape = tuple(oshape[i] for i in range(on.value))
        res = out if out is not None else Tensor.empty(red_shape, odt, self.ctx.device, stream)
        check(lib.hptb_binary_reduce(self.ctx.handle, bop, rop, byref(self._c()), byref(rhs._c()), ax, len(ax_in),
                                     byref(res._c()), 1, _on(stream, self, rhs, res)))
        if keep_dims:
            check(lib.hptb_reduce_shape(bshape, bn.value, ax, len(ax_in), 1, oshape, byref(on)))
            ks = tuple(oshape[i] for i in range(on.value))
            res = Tensor(res.storage, res.ptr, res.dtype, ks, _contig_strides(ks))
        return res

    def mean_var(self, axes, stream=None):
        """Extension (Hpt has no `var`): fused single-read population mean and variance."""
        ax_in = _axes_list(axes)
        ax = (c_int32 * max(len(ax_in), 1))()
        check(lib.hptb_process_axes((c_int64 * max(len(ax_in), 1))(*ax_in), len(ax_in), self.ndim, ax))
        odt = lib.hptb_reduce_out_dtype(_ffi.REDUCE_OPS["mean"], self.dtype)
        oshape = (c_int64 * _ffi.MAX_DIMS)()
        on = c_int()
        check(lib.hptb_reduce_shape((c_int64 * max(self.ndim, 1))(*self.shape), self.ndim, ax, len(ax_in), 0, oshape, byref(on)))
        red_shape = tuple(oshape[i] for i in range(on.value))
        m = Tensor.empty(red_shape, odt, self.ctx.device, stream)
        v = Tensor.empty(red_shape, odt, self.ctx.device, stream)
        check(lib.hptb_mean_var(self.ctx.handle, byref(self._c()), ax, len(ax_in), byref(m._c()), byref(v._c()), _on(stream, self, m, v)))
        return m, v

    # ---- NormalizationOps (hpt/src/backends/cuda/tensor_internal/softmax.rs) ------------------------------
    def _softmax(self, axis, log, stream=None):
        axis = int(axis)
        if axis < 0:
            axis += self.ndim
        if not 0 <= axis < max(self.ndim, 1):
            raise HptError(3, f"axis {axis} out of range for ndim {self.ndim}")
        odt = lib.hptb_unary_out_dtype(0, self.dtype)
        out = Tensor.empty(self.shape, odt, self.ctx.device, stream)
        check(lib.hptb_softmax(self.ctx.handle, byref(self._c()), axis, log, byref(out._c()), _on(stream, self, out)))
        return out

    def layernorm(self, normalized_shape, gamma=None, beta=None, eps=1e-5, stream=None):
        """`x.layernorm(&normalized_shape, Some(&gamma), Some(&beta), eps)` (normalization.rs:30-38)."""
        ns = tuple(int(v) for v in normalized_shape)
        if len(ns) < 1 or len(ns) > self.ndim or tuple(self.shape[self.ndim - len(ns):]) != ns:
            raise HptError(1, f"normalized dims must match last dims of input tensor, shape: {self.shape}, normalized_shape: {ns}")
        odt = lib.hptb_promote(self.dtype, self.dtype, _ffi.PROMOTE_FLOAT_BINARY)
        out = Tensor.empty(self.shape, odt, self.ctx.device, stream)
        g = byref(gamma._c()) if gamma is not None else None
        b = byref(beta._c()) if beta is not None else None
        check(lib.hptb_layernorm(self.ctx.handle, byref(self._c()), len(ns), g, b, float(eps), byref(out._c()), _on(stream, self, gamma, beta, out)))
        return out

    def softmax(self, axis): return self._softmax(axis, 0)
    def log_softmax(self, axis): return self._softmax(axis, 1)


def _make_unary(name):
    def fn(self, out=None, stream=None):
        return self._unary(name, out, 0.0, 0.0, stream)
    fn.__name__ = name
    return fn


for _n in _ffi.UNARY_OPS:
    if _n not in ("selu", "elu", "celu", "leaky_relu", "clamp", "bitnot"):
        setattr(Tensor, _n, _make_unary(_n))
        setattr(Tensor, _n + "_", lambda self, out, _n=_n: self._unary(_n, out))
