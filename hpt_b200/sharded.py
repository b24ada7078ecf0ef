"""Outer-axis sharding of the hot path over the GPUs of one node.

New functionality: the reference has no multi-GPU support (its device is a const generic, hpt/src/tensor.rs:32, and
there are no collectives).  One process per GPU; a `ShardedTensor` is this rank's row block of a tensor split along
one axis (k `Tensor<T, Cuda, i>` in Hpt terms).  Elementwise ops and reductions that keep the shard axis are purely
local; a reduction that crosses it reduces locally to one ACCUMULATOR per output and the k accumulators are combined in
rank order inside `hptb_reduce_sharded` — over NVLink peer memory, in the reduce kernel's own epilogue (xchg.cuh), or
over ncclAllGather when peer memory is unavailable (comm.cpp).  torch.distributed is only the out-of-band channel for the
NCCL unique id.
"""
from ctypes import byref, c_char_p, c_int, c_int32, c_int64, c_void_p, create_string_buffer

from . import _ffi
from ._ffi import HptError, check, lib
from .tensor import Tensor, _axes_list, _contig_strides, _s


def shard_bounds(n, world, rank):
    """(offset, length) of rank's contiguous block of an axis of length n (hptb_shard_bounds)."""
    off, ln = c_int64(), c_int64()
    check(lib.hptb_shard_bounds(int(n), int(world), int(rank), byref(off), byref(ln)))
    return off.value, ln.value


def shard_plan(op, axes, shard_axis, world):
    """What hptb_reduce_sharded does for (op, axes, shard_axis): the library's own plan, as a dict."""
    ax = (c_int32 * max(len(axes), 1))(*axes)
    p = _ffi.HptbShardPlan()
    check(lib.hptb_shard_plan_reduce(_ffi.REDUCE_OPS[op], ax, len(axes), int(shard_axis), int(world), byref(p)))
    return {"crosses": bool(p.crosses), "collective": _ffi.COLLECTIVES[p.collective], "pre_exp": bool(p.pre_exp),
            "post_ln": bool(p.post_ln), "global_count": bool(p.global_count),
            "post_root": int(p.post_root)}


class Comm:
    """One NCCL rank per process (hptb_comm)."""

    def __init__(self, ctx, world, rank, unique_id):
        self.ctx, self.world, self.rank = ctx, int(world), int(rank)
        h = c_void_p()
        check(lib.hptb_comm_init_rank(ctx.handle, self.world, self.rank, c_char_p(unique_id), byref(h)))
        self.handle = h

    @staticmethod
    def unique_id():
        buf = create_string_buffer(128)
        check(lib.hptb_comm_unique_id(buf))
        return buf.raw

    @staticmethod
    def from_torch_distributed(ctx):
        """Rank 0 creates the NCCL id, torch.distributed (any backend) carries it to the other ranks."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(), dist.get_rank()
        obj = [Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        return Comm(ctx, world, rank, obj[0])

    @staticmethod
    def local_group(ctx, nranks):
        """`nranks` virtual ranks on ctx's device (hptb_comm_init_local_group): the exchange protocol without NCCL or
        IPC, for single-GPU validation.  Give every rank its own stream."""
        arr = (c_void_p * int(nranks))()
        check(lib.hptb_comm_init_local_group(ctx.handle, int(nranks), arr))
        out = []
        for r in range(int(nranks)):
            c = Comm.__new__(Comm)
            c.ctx, c.world, c.rank, c.handle = ctx, int(nranks), r, c_void_p(arr[r])
            out.append(c)
        return out

    def destroy(self):
        if self.handle:
            lib.hptb_comm_destroy(self.handle)
            self.handle = None


class ShardedTensor:
    """This rank's block `local` of a tensor of extent `global_len` along `shard_axis`, starting at `offset`."""

    def __init__(self, local, comm, shard_axis, global_len, offset):
        self.local, self.comm = local, comm
        self.shard_axis, self.global_len, self.offset = int(shard_axis), int(global_len), int(offset)

    @staticmethod
    def scatter_from_host(host, comm, shard_axis=0, device=None, stream=None):
        """Every rank holds (or can produce) the full host tensor and uploads only its own block."""
        n = host.shape[shard_axis]
        off, ln = shard_bounds(n, comm.world, comm.rank)
        block = host.narrow(shard_axis, off, ln)
        local = Tensor.to_cuda(block, comm.ctx.device if device is None else device, stream)
        return ShardedTensor(local, comm, shard_axis, n, off)

    @property
    def shape(self):
        s = list(self.local.shape)
        s[self.shard_axis] = self.global_len
        return tuple(s)

    # ---- elementwise: no exchange ------------------------------------------------------------------------
    def _wrap(self, local):
        return ShardedTensor(local, self.comm, self.shard_axis, self.global_len, self.offset)

    def _binary(self, name, rhs):
        if isinstance(rhs, ShardedTensor):
            if (rhs.shard_axis, rhs.global_len, rhs.offset) != (self.shard_axis, self.global_len, self.offset) or \
                    rhs.local.ndim != self.local.ndim:
                raise HptError(1, "sharded operands must share the partition")
            rhs = rhs.local
        elif isinstance(rhs, Tensor):
            # a replicated operand must broadcast along the shard axis (e.g. [1, C] against [N/k, C])
            ax = self.shard_axis - (self.local.ndim - rhs.ndim)
            if ax >= 0 and rhs.shape[ax] != 1:
                raise HptError(1, "a replicated operand must have extent 1 along the shard axis")
        return self._wrap(self.local._binary(name, rhs))

    def __add__(self, rhs): return self._binary("add", rhs)
    def __sub__(self, rhs): return self._binary("sub", rhs)
    def __mul__(self, rhs): return self._binary("mul", rhs)
    def __truediv__(self, rhs): return self._binary("div", rhs)
    def __mod__(self, rhs): return self._binary("rem", rhs)
    def maximum(self, rhs): return self._binary("maximum", rhs)
    def minimum(self, rhs): return self._binary("minimum", rhs)

    def unary(self, name, **kw):
        return self._wrap(getattr(self.local, name)(**kw))

    def softmax(self, axis):
        axis = axis + self.local.ndim if axis < 0 else axis
        if axis == self.shard_axis:
            raise HptError(8, "softmax along the shard axis is not supported")
        return self._wrap(self.local.softmax(axis))

    # ---- reductions -----------------------------------------------------------------------------------------
    def _reduce(self, name, axes, stream=None):
        op = _ffi.REDUCE_OPS[name]
        ax_in = _axes_list(axes)
        nd = self.local.ndim
        ax = (c_int32 * max(len(ax_in), 1))()
        check(lib.hptb_process_axes((c_int64 * max(len(ax_in), 1))(*ax_in), len(ax_in), nd, ax))
        odt = lib.hptb_reduce_out_dtype(op, self.local.dtype)
        if odt < 0:
            raise HptError(2, f"{name} is not supported for {_ffi.DTYPE_NAMES[self.local.dtype]}")
        oshape = (c_int64 * _ffi.MAX_DIMS)()
        on = c_int()
        check(lib.hptb_reduce_shape((c_int64 * max(nd, 1))(*self.local.shape), nd, ax, len(ax_in), 0, oshape, byref(on)))
        red_shape = tuple(oshape[i] for i in range(on.value))
        out = Tensor.empty(red_shape, odt, self.local.ctx.device, stream)
        check(lib.hptb_reduce_sharded(self.comm.handle, op, byref(self.local._c()), ax, len(ax_in), self.shard_axis,
                                      self.offset, self.global_len, byref(out._c()), _s(stream)))
        axes_set = set(ax[i] for i in range(len(ax_in)))
        if self.shard_axis in axes_set:
            return out  # replicated: every rank holds the full result
        new_axis = self.shard_axis - sum(1 for a in axes_set if a < self.shard_axis)
        return ShardedTensor(out, self.comm, new_axis, self.global_len, self.offset)

    def sum(self, axes): return self._reduce("sum", axes)
    def mean(self, axes): return self._reduce("mean", axes)
    def max(self, axes): return self._reduce("max", axes)
    def min(self, axes): return self._reduce("min", axes)
    def prod(self, axes): return self._reduce("prod", axes)
    def logsumexp(self, axes): return self._reduce("logsumexp", axes)
    def sum_square(self, axes): return self._reduce("sum_square", axes)
    def reducel1(self, axes): return self._reduce("reducel1", axes)
    def reducel2(self, axes): return self._reduce("reducel2", axes)
    def reducel3(self, axes): return self._reduce("reducel3", axes)
    def nansum(self, axes): return self._reduce("nansum", axes)
    def nanprod(self, axes): return self._reduce("nanprod", axes)
    def all(self, axes): return self._reduce("all", axes)
    def any(self, axes): return self._reduce("any", axes)
    def argmax(self, axis): return self._reduce("argmax", axis)
    def argmin(self, axis): return self._reduce("argmin", axis)
