//! Raw bindings to include/hpt_b200.h.  UNVERIFIED SOURCE: the build image has no Rust toolchain, so this
//! file has not been compiled; the same symbols are bound and exercised from Python (hpt_b200/_ffi.py).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_int, c_void};

pub const HPTB_MAX_DIMS: usize = 8;
pub const HPTB_NCCL_ID_BYTES: usize = 128;

pub type hptb_status = c_int;
pub const HPTB_OK: hptb_status = 0;
pub const HPTB_ERR_SHAPE: hptb_status = 1;
pub const HPTB_ERR_DTYPE: hptb_status = 2;
pub const HPTB_ERR_AXIS: hptb_status = 3;
pub const HPTB_ERR_INVALID: hptb_status = 4;
pub const HPTB_ERR_CUDA: hptb_status = 5;
pub const HPTB_ERR_OOM: hptb_status = 6;
pub const HPTB_ERR_NCCL: hptb_status = 7;
pub const HPTB_ERR_UNSUPPORTED: hptb_status = 8;

/// hptb_dtype, in the order of hpt-types/src/dtype.rs
#[repr(i32)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum hptb_dtype { Bool = 0, I8, I16, I32, I64, U8, U16, U32, U64, F16, BF16, F32, F64 }

#[repr(i32)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum hptb_binary_op { Add = 0, Sub, Mul, Rem, Div, Maximum, Minimum }

#[repr(i32)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum hptb_unary_op {
    Sin = 0, Cos, Tan, Asin, Acos, Atan, Sinh, Cosh, Tanh, Asinh, Acosh, Atanh, Exp, Exp2, Exp10, Ln, Log2, Log10,
    Sqrt, Cbrt, Recip, Erf, Sigmoid, Gelu, Selu, Elu, Celu, Mish, Softplus, Softsign, HardSigmoid, HardSwish,
}

#[repr(i32)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum hptb_reduce_op { Sum = 0, Mean, Max, Min, Argmax, Argmin, Logsumexp, SumSquare, Prod }

#[repr(C)]
#[derive(Clone, Copy)]
pub struct hptb_tensor {
    pub data: *mut c_void,
    pub dtype: i32,
    pub ndim: i32,
    pub shape: [i64; HPTB_MAX_DIMS],
    pub strides: [i64; HPTB_MAX_DIMS],
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct hptb_alloc_stats {
    pub bytes_in_use: u64, pub bytes_cached: u64, pub bytes_reserved_peak: u64,
    pub n_alloc: u64, pub n_cache_hit: u64, pub n_device_malloc: u64, pub n_device_free: u64,
}

#[repr(C)]
pub struct hptb_collapse_plan {
    pub ndim: i32, pub launch_class: i32, pub n_operands: i32, pub reserved: i32,
    pub shape: [i64; HPTB_MAX_DIMS],
    pub strides: [[i64; HPTB_MAX_DIMS]; 4],
    pub reduced: [u8; HPTB_MAX_DIMS],
}

/// hptb_route_kind: 0 direct, 1 peel (head / aligned body / tail along the last axis), 2 two-step (input-ordered scratch)
#[repr(C)] #[derive(Clone, Copy, Debug, Default)]
pub struct hptb_reduce_route_t {
    pub kind: i32,
    pub reserved: i32,
    pub head: i64,
    pub body: i64,
    pub tail: i64,
    pub scratch_strides: [i64; HPTB_MAX_DIMS],
}

#[repr(C)] pub struct hptb_ctx { _private: [u8; 0] }
#[repr(C)] pub struct hptb_comm { _private: [u8; 0] }
#[repr(C)] #[derive(Clone, Copy, Debug, Default)]
pub struct hptb_shard_plan { pub crosses: i32, pub collective: i32, pub pre_exp: i32, pub post_ln: i32, pub global_count: i32, pub post_root: i32 }

extern "C" {
    pub fn hptb_version() -> c_int;
    pub fn hptb_kernel_launches() -> u64;
    pub fn hptb_last_error() -> *const c_char;
    pub fn hptb_dtype_size(dtype: c_int) -> usize;
    pub fn hptb_dtype_name(dtype: c_int) -> *const c_char;
    pub fn hptb_ctx_create(device: c_int, out: *mut *mut hptb_ctx) -> hptb_status;
    pub fn hptb_ctx_destroy(ctx: *mut hptb_ctx) -> hptb_status;
    pub fn hptb_ctx_device(ctx: *const hptb_ctx, device: *mut c_int) -> hptb_status;
    pub fn hptb_ctx_sm_count(ctx: *const hptb_ctx, sms: *mut c_int) -> hptb_status;
    pub fn hptb_stream_sync(ctx: *mut hptb_ctx, stream: *mut c_void) -> hptb_status;
    pub fn hptb_alloc(ctx: *mut hptb_ctx, bytes: usize, ptr: *mut *mut c_void, stream: *mut c_void) -> hptb_status;
    pub fn hptb_free(ctx: *mut hptb_ctx, ptr: *mut c_void, stream: *mut c_void) -> hptb_status;
    pub fn hptb_empty_cache(ctx: *mut hptb_ctx) -> hptb_status;
    pub fn hptb_alloc_get_stats(ctx: *mut hptb_ctx, out: *mut hptb_alloc_stats) -> hptb_status;
    pub fn hptb_alloc_selftest() -> hptb_status;
    pub fn hptb_memcpy_h2d(ctx: *mut hptb_ctx, dst: *mut c_void, src: *const c_void, bytes: usize, stream: *mut c_void) -> hptb_status;
    pub fn hptb_memcpy_d2h(ctx: *mut hptb_ctx, dst: *mut c_void, src: *const c_void, bytes: usize, stream: *mut c_void) -> hptb_status;
    pub fn hptb_memcpy_d2d(ctx: *mut hptb_ctx, dst: *mut c_void, src: *const c_void, bytes: usize, stream: *mut c_void) -> hptb_status;
    pub fn hptb_memcpy_d2h_async(ctx: *mut hptb_ctx, dst_pinned: *mut c_void, src: *const c_void, bytes: usize, stream: *mut c_void) -> hptb_status;
    pub fn hptb_stream_wait_stream(ctx: *mut hptb_ctx, stream: *mut c_void, other: *mut c_void) -> hptb_status;
    pub fn hptb_host_alloc_pinned(bytes: usize, ptr: *mut *mut c_void) -> hptb_status;
    pub fn hptb_host_free_pinned(ptr: *mut c_void) -> hptb_status;
    pub fn hptb_promote(lhs: c_int, rhs: c_int, kind: c_int) -> c_int;
    pub fn hptb_binary_out_dtype(op: c_int, lhs: c_int, rhs: c_int) -> c_int;
    pub fn hptb_unary_out_dtype(op: c_int, inp: c_int) -> c_int;
    pub fn hptb_reduce_out_dtype(op: c_int, inp: c_int) -> c_int;
    pub fn hptb_broadcast_shape(a: *const i64, na: c_int, b: *const i64, nb: c_int, out: *mut i64, nout: *mut c_int) -> hptb_status;
    pub fn hptb_process_axes(axes: *const i64, naxes: c_int, ndim: c_int, out: *mut i32) -> hptb_status;
    pub fn hptb_reduce_shape(shape: *const i64, ndim: c_int, axes: *const i32, naxes: c_int, keep_dims: c_int,
                             out_shape: *mut i64, out_ndim: *mut c_int) -> hptb_status;
    pub fn hptb_collapse(operands: *const *const hptb_tensor, n_operands: c_int, reduce_mask: *const u8,
                         plan: *mut hptb_collapse_plan) -> hptb_status;
    pub fn hptb_reduce_route(op: c_int, input: *const hptb_tensor, axes: *const i32, naxes: c_int, out: *const hptb_tensor,
                             init_out: c_int, route: *mut hptb_reduce_route_t) -> hptb_status;
    pub fn hptb_binary(ctx: *mut hptb_ctx, op: c_int, lhs: *const hptb_tensor, rhs: *const hptb_tensor,
                       out: *mut hptb_tensor, stream: *mut c_void) -> hptb_status;
    pub fn hptb_compare(ctx: *mut hptb_ctx, op: c_int, lhs: *const hptb_tensor, rhs: *const hptb_tensor,
                        out: *mut hptb_tensor, stream: *mut c_void) -> hptb_status;
    pub fn hptb_unary(ctx: *mut hptb_ctx, op: c_int, inp: *const hptb_tensor, out: *mut hptb_tensor,
                      alpha: c_double, beta: c_double, stream: *mut c_void) -> hptb_status;
    pub fn hptb_reduce(ctx: *mut hptb_ctx, op: c_int, inp: *const hptb_tensor, axes: *const i32, naxes: c_int,
                       out: *mut hptb_tensor, init_out: c_int, stream: *mut c_void) -> hptb_status;
    pub fn hptb_binary_reduce(ctx: *mut hptb_ctx, bin_op: c_int, red_op: c_int, lhs: *const hptb_tensor, rhs: *const hptb_tensor,
                              axes: *const i32, naxes: c_int, out: *mut hptb_tensor, init_out: c_int, stream: *mut c_void) -> hptb_status;
    pub fn hptb_mean_var(ctx: *mut hptb_ctx, inp: *const hptb_tensor, axes: *const i32, naxes: c_int,
                         mean_out: *mut hptb_tensor, var_out: *mut hptb_tensor, stream: *mut c_void) -> hptb_status;
    pub fn hptb_softmax(ctx: *mut hptb_ctx, inp: *const hptb_tensor, axis: c_int, log: c_int,
                        out: *mut hptb_tensor, stream: *mut c_void) -> hptb_status;
    pub fn hptb_layernorm(ctx: *mut hptb_ctx, inp: *const hptb_tensor, n_normalized_dims: c_int, gamma: *const hptb_tensor,
                          beta: *const hptb_tensor, eps: c_double, out: *mut hptb_tensor, stream: *mut c_void) -> hptb_status;
    pub fn hptb_copy(ctx: *mut hptb_ctx, inp: *const hptb_tensor, out: *mut hptb_tensor, stream: *mut c_void) -> hptb_status;
    pub fn hptb_fill(ctx: *mut hptb_ctx, out: *mut hptb_tensor, scalar: *const c_void, stream: *mut c_void) -> hptb_status;
    pub fn hptb_arange(ctx: *mut hptb_ctx, out: *mut hptb_tensor, start: *const c_void, step: *const c_void, stream: *mut c_void) -> hptb_status;
    pub fn hptb_eye(ctx: *mut hptb_ctx, out: *mut hptb_tensor, k: i64, stream: *mut c_void) -> hptb_status;
    pub fn hptb_comm_unique_id(id128: *mut c_void) -> hptb_status;
    pub fn hptb_comm_init_rank(ctx: *mut hptb_ctx, nranks: c_int, rank: c_int, id128: *const c_void,
                               out: *mut *mut hptb_comm) -> hptb_status;
    pub fn hptb_comm_destroy(comm: *mut hptb_comm) -> hptb_status;
    pub fn hptb_comm_uses_peer_memory(comm: *const hptb_comm) -> c_int;
    pub fn hptb_shard_bounds(n: i64, world: c_int, rank: c_int, offset: *mut i64, len: *mut i64) -> hptb_status;
    pub fn hptb_shard_plan_reduce(op: c_int, axes: *const i32, naxes: c_int, shard_axis: c_int, world: c_int,
                                  plan: *mut hptb_shard_plan) -> hptb_status;
    pub fn hptb_allreduce(comm: *mut hptb_comm, op: c_int, inout: *mut hptb_tensor, stream: *mut c_void) -> hptb_status;
    pub fn hptb_reduce_sharded(comm: *mut hptb_comm, op: c_int, shard: *const hptb_tensor, axes: *const i32, naxes: c_int,
                               shard_axis: c_int, shard_offset: i64, global_axis_len: i64, out: *mut hptb_tensor,
                               stream: *mut c_void) -> hptb_status;
}
