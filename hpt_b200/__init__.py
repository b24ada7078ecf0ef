"""hpt_b200 — B200-native backend for Hpt's strided/broadcast elementwise + axis-reduction hot path.

`hpt_b200.Tensor` mirrors `hpt::Tensor<T, Cuda, DEVICE>` for that path over the C ABI of
libhpt_b200.so (include/hpt_b200.h).  Importing this package requires the built extension.
"""
from . import _ffi
from ._ffi import (BF16, BOOL, F16, F32, F64, I8, I16, I32, I64, U8, U16, U32, U64, DTYPE_NAMES, HptError, lib)
from .tensor import Context, Tensor, context, get_stream, set_stream
from .sharded import Comm, ShardedTensor, shard_bounds, shard_plan

__all__ = ["Tensor", "Context", "Comm", "ShardedTensor", "shard_bounds", "shard_plan", "context", "set_stream", "get_stream", "HptError", "lib", "DTYPE_NAMES",
           "BOOL", "I8", "I16", "I32", "I64", "U8", "U16", "U32", "U64", "F16", "BF16", "F32", "F64"]
