/*
 * hpt_b200.h — C ABI of libhpt_b200: a B200 (sm_100a) backend for Hpt's data-parallel hot path.
 *
 * This is the drop-in boundary.  Every entry point replaces one place where the
 * reference's Rust host code looks up a PTX kernel by string name and launches it through cudarc
 * (`load_ptx_and_get_data` + `CudaFunction::launch`, hpt/src/backends/cuda/cuda_utils.rs:71-129).
 * The Rust shim (rust/hpt-b200-sys, delivered as source) binds these symbols 1:1; the Python
 * host mirror (hpt_b200/) binds them through ctypes.  Reference file:line for each entry is
 * given next to its declaration.  All paths below are relative to the Hpt repository root.
 *
 * Conventions
 *  - plain C: pointers, sizes, enums; no C++ types and no exceptions cross this boundary.
 *  - every call returns hptb_status (0 = ok); the message for the last failure on the calling
 *    thread is available from hptb_last_error().  Nothing aborts or panics.
 *  - every compute call is asynchronous and ordered on `stream` (a cudaStream_t passed as
 *    void*; NULL = the legacy default stream, which is what cudarc 0.13's CudaDevice uses).
 *  - the caller owns every tensor buffer, including `out` (mirrors `extract_out`,
 *    hpt/src/backends/cuda/utils/binary/binary_normal.rs:546-564); scratch is library-owned and
 *    comes from the context's stream-ordered pool.
 *  - strides are in ELEMENTS (as in hpt-common Layout), may be 0 (broadcast) or negative.
 *  - sizes and indices are 64-bit throughout (the reference is i32-limited,
 *    hpt/src/backends/cuda/tensor_internal/normal_creation.rs:41-43).
 */
#ifndef HPT_B200_H
#define HPT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define HPTB_MAX_DIMS 8
#define HPTB_VERSION 100

/* Status codes.  Map to hpt-common/src/error/{shape,param,kernel,device,memory}.rs on the Rust side. */
typedef enum {
  HPTB_OK = 0,
  HPTB_ERR_SHAPE = 1,       /* ShapeError: broadcast mismatch, out layout invalid, size mismatch */
  HPTB_ERR_DTYPE = 2,       /* unsupported dtype / op combination (reference: kernel lookup miss → KernelError) */
  HPTB_ERR_AXIS = 3,        /* ParamError::AxisDuplicated / ShapeError::DimOutOfRange (axis.rs:38-72) */
  HPTB_ERR_INVALID = 4,     /* null pointers, bad enum values, ndim > HPTB_MAX_DIMS */
  HPTB_ERR_CUDA = 5,        /* DeviceError: a CUDA runtime call failed (message carries the code) */
  HPTB_ERR_OOM = 6,         /* MemoryError: device allocation failed even after emptying the cache */
  HPTB_ERR_NCCL = 7,        /* collective failed / NCCL unavailable */
  HPTB_ERR_UNSUPPORTED = 8  /* valid request the library does not implement yet */
} hptb_status;

/* Dtype order follows hpt-types/src/dtype.rs (bool, i8..u64, f16, bf16, f32, f64). */
typedef enum {
  HPTB_BOOL = 0, HPTB_I8 = 1, HPTB_I16 = 2, HPTB_I32 = 3, HPTB_I64 = 4,
  HPTB_U8 = 5, HPTB_U16 = 6, HPTB_U32 = 7, HPTB_U64 = 8,
  HPTB_F16 = 9, HPTB_BF16 = 10, HPTB_F32 = 11, HPTB_F64 = 12,
  HPTB_DTYPE_COUNT = 13
} hptb_dtype;

/* A strided view of device memory: the C image of `_Tensor{data, layout}` (hpt/src/tensor_base.rs:17-29).
 * `data` already points at the first logical element (as Hpt's `Pointer<T>` does after slicing). */
typedef struct {
  void* data;
  int32_t dtype;                  /* hptb_dtype */
  int32_t ndim;                   /* 0..HPTB_MAX_DIMS; ndim 0 = one element */
  int64_t shape[HPTB_MAX_DIMS];
  int64_t strides[HPTB_MAX_DIMS]; /* elements */
} hptb_tensor;

/* Binary ops.  ADD..REM replace `{add,sub,mul,rem}_<L>_<R>_{contiguous,uncontiguous,*_scalar}`
 * (hpt-cudakernels/src/binary/binary_template.cuh:107-138) called from binary_fn_precompiled
 * (hpt/src/backends/cuda/utils/binary/binary_normal.rs:371-544).  Output dtype NormalOutPromote<L,R>.
 * DIV is FloatBinOps::div_ (hpt-traits/src/ops/binary.rs:149), output FloatOutBinaryPromote<L,R>.
 * MAX/MIN are NormalOut::_max/_min (hpt-macros/src/normal_out.rs:121-133).
 * POW/HYPOT are FloatBinOps::{pow,hypot} (hpt-traits/src/ops/binary.rs:114-183), output FloatOutBinaryPromote<L,R>.
 * BITAND..SHR are BitWiseOut (hpt-macros/src/lib.rs:384-480): bool and integer dtypes only, both sides cast to
 * NormalOutPromote<L,R>, shifts wrap the count to the bit width (wrapping_shl/shr, hpt-types/src/scalars/impls.rs:155-162). */
typedef enum {
  HPTB_ADD = 0, HPTB_SUB = 1, HPTB_MUL = 2, HPTB_REM = 3, HPTB_DIV = 4, HPTB_MAXIMUM = 5, HPTB_MINIMUM = 6,
  HPTB_POW = 7, HPTB_HYPOT = 8,
  HPTB_BITAND = 9, HPTB_BITOR = 10, HPTB_BITXOR = 11, HPTB_SHL = 12, HPTB_SHR = 13,
  HPTB_BINARY_COUNT = 14
} hptb_binary_op;

/* TensorCmp (hpt-traits/src/ops/cmp.rs; device: hpt-cudakernels/src/binary/cmp.cu): both sides are cast to
 * NormalOutPromote<L,R> and compared there (hpt-macros/src/lib.rs:570-650); the output is bool. */
typedef enum {
  HPTB_EQ = 0, HPTB_NE = 1, HPTB_LT = 2, HPTB_LE = 3, HPTB_GT = 4, HPTB_GE = 5,
  HPTB_CMP_COUNT = 6
} hptb_cmp_op;

/* FloatUnaryOps (hpt-traits/src/ops/unary.rs:8-608); replaces `<op>_<T>_{contiguous,uncontiguous}`
 * (hpt-cudakernels/src/unary/unary_template.cuh:65-75) called from uary_fn_precompiled
 * (hpt/src/backends/cuda/utils/unary/unary.rs:122-196).  Output dtype FloatOutUnaryPromote<T>. */
typedef enum {
  HPTB_SIN = 0, HPTB_COS, HPTB_TAN, HPTB_ASIN, HPTB_ACOS, HPTB_ATAN,
  HPTB_SINH, HPTB_COSH, HPTB_TANH, HPTB_ASINH, HPTB_ACOSH, HPTB_ATANH,
  HPTB_EXP, HPTB_EXP2, HPTB_EXP10, HPTB_LN, HPTB_LOG2, HPTB_LOG10,
  HPTB_SQRT, HPTB_CBRT, HPTB_RECIP, HPTB_ERF,
  HPTB_SIGMOID, HPTB_GELU, HPTB_SELU /* alpha, beta=scale */, HPTB_ELU /* alpha */, HPTB_CELU /* alpha */,
  HPTB_MISH, HPTB_SOFTPLUS, HPTB_SOFTSIGN, HPTB_HARD_SIGMOID, HPTB_HARD_SWISH,
  /* NormalUaryOps (hpt-traits/src/ops/unary.rs:611-836; scalar semantics hpt-types/src/scalars/{_f32,impls,_bool}.rs
   * NormalOutUnary2): output dtype = input dtype, every dtype.  LEAKY_RELU takes alpha; CLAMP takes alpha = min,
   * beta = max (doubles: integer bounds beyond 2^53 are not representable).  BITNOT is BitWiseOut::_not (bool and
   * integer dtypes only). */
  HPTB_FLOOR, HPTB_CEIL, HPTB_ROUND, HPTB_TRUNC, HPTB_ABS, HPTB_NEG, HPTB_SIGN, HPTB_SQUARE,
  HPTB_RELU, HPTB_RELU6, HPTB_LEAKY_RELU, HPTB_CLAMP, HPTB_BITNOT,
  HPTB_UNARY_COUNT
} hptb_unary_op;
#define HPTB_FLOAT_UNARY_COUNT HPTB_FLOOR /* ops below this value promote with FloatOutUnaryPromote */

/* Reductions: NormalReduce / FloatReduce / IndexReduce (hpt-traits/src/ops/reduce.rs:6-358); replaces
 * the 8-kernel-per-op families of hpt-cudakernels/src/reduce/declare_macros.cuh:48-326 chosen by
 * hpt/src/backends/cuda/utils/reduce/reduce.rs:62-838.
 * Output dtypes: SUM/MAX/MIN/PROD/SUM_SQUARE = T; MEAN/LOGSUMEXP = FloatOutBinaryPromote<T,T>;
 * ARGMAX/ARGMIN = i64 (exactly one axis).
 * REDUCEL1 = Σ|x| (T); NANSUM / NANPROD treat NaN as 0 / 1 (T); ALL / ANY = bool (x != 0; NaN is true);
 * REDUCEL2 = sqrt Σ x², REDUCEL3 = (Σ |x|³)^(1/3), both FloatOutBinaryPromote<T,T>
 * (hpt/src/backends/cpu/tensor_internal/common_reduce.rs:147-167,200-324,384-450). */
typedef enum {
  HPTB_SUM = 0, HPTB_MEAN = 1, HPTB_MAX = 2, HPTB_MIN = 3, HPTB_ARGMAX = 4, HPTB_ARGMIN = 5,
  HPTB_LOGSUMEXP = 6, HPTB_SUM_SQUARE = 7, HPTB_PROD = 8,
  HPTB_REDUCEL1 = 9, HPTB_NANSUM = 10, HPTB_NANPROD = 11, HPTB_ALL = 12, HPTB_ANY = 13,
  HPTB_REDUCEL2 = 14, HPTB_REDUCEL3 = 15,
  HPTB_REDUCE_COUNT = 16
} hptb_reduce_op;

typedef enum {
  HPTB_PROMOTE_NORMAL = 0,        /* NormalOutPromote<L,R>::Output        (hpt-types/src/promotion/normal_promote/_*.rs) */
  HPTB_PROMOTE_FLOAT_BINARY = 1,  /* FloatOutBinaryPromote<L,R>::Output */
  HPTB_PROMOTE_FLOAT_UNARY = 2    /* FloatOutUnaryPromote<L>::Output (rhs ignored) */
} hptb_promote_kind;

typedef struct hptb_ctx hptb_ctx;    /* one per (process, device): stream-ordered pool + scratch */
typedef struct hptb_comm hptb_comm;  /* one NCCL rank */

/* ---- library / context --------------------------------------------------------------------- */
int hptb_version(void);
/* kernels launched by this library in this process so far (bench evidence; monotonically increasing) */
uint64_t hptb_kernel_launches(void);
const char* hptb_last_error(void);   /* thread-local, valid until the next failing call on this thread */
size_t hptb_dtype_size(int dtype);
const char* hptb_dtype_name(int dtype);

/* Replaces `CudaDevice::new(id)` + per-device LRU creation (hpt-allocator/src/allocators/cuda.rs:50-69). */
hptb_status hptb_ctx_create(int device, hptb_ctx** out);
hptb_status hptb_ctx_destroy(hptb_ctx* ctx);
hptb_status hptb_ctx_device(const hptb_ctx* ctx, int* device);
hptb_status hptb_ctx_sm_count(const hptb_ctx* ctx, int* sms);
hptb_status hptb_stream_sync(hptb_ctx* ctx, void* stream);

/* ---- allocator: replaces HptAllocator<Cuda> (hpt-allocator/src/allocators/cuda.rs:41-199,
 *      utils/allocate.rs:66-124, utils/deallocate.rs:11-32) with a stream-ordered caching pool ------- */
hptb_status hptb_alloc(hptb_ctx* ctx, size_t bytes, void** ptr, void* stream);
hptb_status hptb_free(hptb_ctx* ctx, void* ptr, void* stream);
/* `ptr` (from hptb_alloc, still live) is consumed by work enqueued on `stream`, which is not the stream it will be
 * freed on: the block is not handed out again before that work has completed (one event per recorded stream at free
 * time).  Call it whenever a tensor is used on a stream other than its allocation / free stream — without it a block
 * dropped right after enqueueing cross-stream work could be reused while that work still reads it. */
hptb_status hptb_record_stream(hptb_ctx* ctx, void* ptr, void* stream);
hptb_status hptb_empty_cache(hptb_ctx* ctx);              /* resize_cuda_lru_cache(0) analogue, hpt/src/lib.rs:441-511 */
typedef struct {
  uint64_t bytes_in_use, bytes_cached, bytes_reserved_peak;
  uint64_t n_alloc, n_cache_hit, n_device_malloc, n_device_free;
} hptb_alloc_stats;
hptb_status hptb_alloc_get_stats(hptb_ctx* ctx, hptb_alloc_stats* out);
/* host-only self test of the caching logic against a fake device (runs without a GPU). */
hptb_status hptb_alloc_selftest(void);

/* ---- transfers: to_cuda / to_cpu (hpt/src/backends/cpu/tensor_impls.rs:295-313,
 *      hpt/src/backends/cuda/tensor_impls.rs:142-169) ------------------------------------------------ */
hptb_status hptb_memcpy_h2d(hptb_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes, void* stream);
hptb_status hptb_memcpy_d2h(hptb_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes, void* stream);
hptb_status hptb_memcpy_d2d(hptb_ctx* ctx, void* dst_dev, const void* src_dev, size_t bytes, void* stream);
/* Asynchronous read-back into PINNED host memory: ordered on `stream`, returns immediately; the host buffer is
 * valid after hptb_stream_sync (or an event) on that stream.  An addition over the reference's blocking to_cpu
 * that lets a caller overlap the read-back of one result with the next kernels and the next upload. */
hptb_status hptb_memcpy_d2h_async(hptb_ctx* ctx, void* dst_pinned_host, const void* src_dev, size_t bytes, void* stream);
/* Make `stream` wait for everything enqueued so far on `other` (event record + stream wait; no host sync). */
hptb_status hptb_stream_wait_stream(hptb_ctx* ctx, void* stream, void* other);
hptb_status hptb_host_alloc_pinned(size_t bytes, void** ptr);
hptb_status hptb_host_free_pinned(void* ptr);

/* ---- type promotion and layout helpers (pure host; usable without a GPU) ------------------------------ */
/* Returns the promoted dtype or -1.  Tables restated from hpt-types/src/promotion/normal_promote/_*.rs. */
int hptb_promote(int lhs, int rhs, int kind);
int hptb_binary_out_dtype(int op, int lhs, int rhs);
int hptb_unary_out_dtype(int op, int in);
int hptb_reduce_out_dtype(int op, int in);
/* numpy-style broadcast of two shapes (hpt-common/src/shape/shape_utils.rs:370-400). */
hptb_status hptb_broadcast_shape(const int64_t* a, int na, const int64_t* b, int nb, int64_t* out, int* nout);
/* process_axes (hpt-common/src/axis/axis.rs:38-72): normalises negatives, rejects duplicates / out of range. */
hptb_status hptb_process_axes(const int64_t* axes, int naxes, int ndim, int32_t* out);
/* Layout::reduce (hpt-common/src/layout/layout_utils.rs:310-349): output shape; all axes reduced → [1]. */
hptb_status hptb_reduce_shape(const int64_t* shape, int ndim, const int32_t* axes, int naxes, int keep_dims,
                              int64_t* out_shape, int* out_ndim);

/* The shape/stride collapse + broadcast-folding pass, exposed for inspection and tests.
 * operands[0] is the output.  `reduce_mask` (may be NULL) flags reduced dims.  On return `plan` holds
 * the collapsed dims (outermost first) with per-operand strides and the launch class. */
typedef enum { HPTB_CLASS_CONTIGUOUS = 0, HPTB_CLASS_INNER_CONTIGUOUS = 1, HPTB_CLASS_STRIDED = 2 } hptb_launch_class;
typedef struct {
  int32_t ndim;
  int32_t launch_class;
  int32_t n_operands;
  int32_t reserved;
  int64_t shape[HPTB_MAX_DIMS];
  int64_t strides[4][HPTB_MAX_DIMS];
  uint8_t reduced[HPTB_MAX_DIMS];
} hptb_collapse_plan;
hptb_status hptb_collapse(const hptb_tensor* const* operands, int n_operands, const uint8_t* reduce_mask,
                          hptb_collapse_plan* plan);
/* Host-side routing of a reduction (pure; no launch): how hptb_reduce will run it.  New functionality — the
 * reference's planner (reduce.rs:641-838) has neither case.  DIRECT: one kernel.  PEEL: rows start off the 16-byte
 * boundary with a common misalignment → head / aligned body / tail along the last axis, folded into `out`
 * (PEEL_RAW: into a scratch of accumulators finished by one more kernel — mean, logsumexp, reducel2/3, f16 / bf16).
 * TWO_STEP: the output's fastest dim is not the input's fastest kept dim → reduce into a scratch with
 * `scratch_strides` (input dim order), then gather into `out`. */
typedef enum hptb_route_kind { HPTB_ROUTE_DIRECT = 0, HPTB_ROUTE_PEEL = 1, HPTB_ROUTE_TWO_STEP = 2, HPTB_ROUTE_PEEL_RAW = 3 } hptb_route_kind;
typedef struct hptb_reduce_route_t {
  int32_t kind;      /* hptb_route_kind */
  int32_t reserved;
  int64_t head, body, tail;                 /* PEEL: extents along the last axis */
  int64_t scratch_strides[HPTB_MAX_DIMS];   /* TWO_STEP: element strides of the scratch, per `out` dim */
} hptb_reduce_route_t;
hptb_status hptb_reduce_route(int op, const hptb_tensor* in, const int32_t* axes, int naxes, const hptb_tensor* out,
                              int init_out, hptb_reduce_route_t* route);

/* ---- compute entry points ------------------------------------------------------------------------------- */
/* out = op(cast(lhs), cast(rhs)) with numpy broadcasting.  `out->dtype` must equal
 * hptb_binary_out_dtype(op, lhs, rhs) and `out->shape` the broadcast shape; any strides. */
hptb_status hptb_binary(hptb_ctx* ctx, int op, const hptb_tensor* lhs, const hptb_tensor* rhs,
                        hptb_tensor* out, void* stream);
/* out = lhs <op> rhs (hptb_cmp_op) with numpy broadcasting; `out` is bool with the broadcast shape. */
hptb_status hptb_compare(hptb_ctx* ctx, int op, const hptb_tensor* lhs, const hptb_tensor* rhs,
                         hptb_tensor* out, void* stream);
/* out = op(cast(in)); alpha/beta only for ELU/CELU/LEAKY_RELU (alpha), SELU (alpha, scale), CLAMP (min, max). */
hptb_status hptb_unary(hptb_ctx* ctx, int op, const hptb_tensor* in, hptb_tensor* out,
                       double alpha, double beta, void* stream);
/* Reduce `in` over `axes` (already normalised, unique).  `out` has the keep_dims=false shape
 * (or [1] when every axis is reduced) and may carry any strides.  init_out follows reduce_prepare
 * (hpt/src/backends/cpu/utils/reduce/reduce_utils.rs:46-50): non-zero = `out` is initialised to the
 * identity first, i.e. it receives the plain result; zero = the previous contents of `out` are folded
 * into the result with the op's combine (sum_ accumulating into an existing buffer).  Pass 1 for a
 * freshly allocated output. */
hptb_status hptb_reduce(hptb_ctx* ctx, int op, const hptb_tensor* in, const int32_t* axes, int naxes,
                        hptb_tensor* out, int init_out, void* stream);
/* Elementwise → reduce fusion (conceptual hook: hpt-codegen/src/fuse/*): out = reduce_<red_op>(
 * lhs <bin_op> rhs, axes), with exactly the semantics of hptb_binary followed by hptb_reduce but — for same-dtype
 * add / sub / mul feeding sum / max / min / sum_square over a unit-stride axis — computed in ONE pass that never
 * writes the elementwise result (config 1: (A + B).sum(1) reads 67 MB instead of moving 201 MB).  Other
 * combinations run the two kernels through a pooled temporary. */
hptb_status hptb_binary_reduce(hptb_ctx* ctx, int bin_op, int red_op, const hptb_tensor* lhs, const hptb_tensor* rhs,
                               const int32_t* axes, int naxes, hptb_tensor* out, int init_out, void* stream);
/* Fused single-read mean + variance (population, ddof = 0) — an EXTENSION: Hpt has no `var`
 * (SURVEY.md §8a row a7); both outputs have dtype FloatOutBinaryPromote<T,T>. */
hptb_status hptb_mean_var(hptb_ctx* ctx, const hptb_tensor* in, const int32_t* axes, int naxes,
                          hptb_tensor* mean_out, hptb_tensor* var_out, void* stream);
/* softmax / log_softmax along one axis (NormalizationOps, hpt-traits/src/ops/normalization.rs:51-64);
 * replaces `<T>_{softmax,logsoftmax}_{warp,block,block_large}` (hpt-cudakernels/src/normalization/softmax.cu).
 * Computed in f32 (f64 for 64-bit types) as exp(x - max) / sum; against an f64 evaluation the result is within
 * 4 + |x - max| ulp of the output type when the lane fits registers or a cluster's shared memory, 6 + |x - max| ulp
 * through the two-sweep kernels for longer lanes.  Any layout; axis < 0 counts from the end. */
hptb_status hptb_softmax(hptb_ctx* ctx, const hptb_tensor* in, int axis, int log, hptb_tensor* out, void* stream);
/* NormalizationOps::layernorm over the LAST n_normalized_dims dims (hpt-traits/src/ops/normalization.rs:12-38;
 * hpt/src/backends/cuda/tensor_internal/layernorm.rs:47-…, kernels layernorm.cu + layernorm_post.cu):
 * y = (x − mean) / sqrt(var + eps) · gamma + beta, population variance; gamma / beta (either may be NULL) are
 * contiguous tensors of the normalized shape; every tensor but `in` has dtype FloatOutBinaryPromote<T,T>. */
hptb_status hptb_layernorm(hptb_ctx* ctx, const hptb_tensor* in, int n_normalized_dims, const hptb_tensor* gamma,
                           const hptb_tensor* beta, double eps, hptb_tensor* out, void* stream);
/* Gather a view into another layout with optional dtype conversion (Rust `as` semantics):
 * replaces strided_copy_<T> (hpt-cudakernels/src/strided_copy.cu) used by contiguous()/to_cpu. */
hptb_status hptb_copy(hptb_ctx* ctx, const hptb_tensor* in, hptb_tensor* out, void* stream);
/* out[...] = *scalar (scalar has out's dtype, host memory): replaces set_val_<T> / fill_<T>. */
hptb_status hptb_fill(hptb_ctx* ctx, hptb_tensor* out, const void* scalar, void* stream);
/* TensorCreator (hpt-traits/src/ops/creation.rs; replaces the creation kernels of hpt-cudakernels/src/creation.cu and
 * the NVRTC arange of tensor_internal/normal_creation.rs).  zeros / ones / full are hptb_fill.
 * hptb_arange: out[i] = start + T(i)·step for a 1-D `out`, every step rounded in T as the reference's
 * `start._add(i.cast()._mul(step))` (arange, arange_step, linspace; scalars have out's dtype, host memory).
 * hptb_eye: 2-D `out`, 1 where col == row + k (eye, identity). */
hptb_status hptb_arange(hptb_ctx* ctx, hptb_tensor* out, const void* start, const void* step, void* stream);
hptb_status hptb_eye(hptb_ctx* ctx, hptb_tensor* out, int64_t k, void* stream);

/* ---- multi-GPU (new: the reference has no collectives) ---------------------------------- */
#define HPTB_NCCL_ID_BYTES 128
hptb_status hptb_comm_unique_id(void* id128);                          /* rank 0, then broadcast out of band */
hptb_status hptb_comm_init_rank(hptb_ctx* ctx, int nranks, int rank, const void* id128, hptb_comm** out);
/* `nranks` VIRTUAL ranks on one device, no NCCL and no IPC: comms[r] behaves like rank r of an nranks-rank communicator
 * whose mailboxes are plain allocations of ctx's device.  For tests and single-GPU validation of the exchange protocol
 * (the kernels of one collective call spin on each other: give every virtual rank its own stream and keep the shards
 * small enough for their kernels to be resident together).  At most 65536 outputs per call. */
hptb_status hptb_comm_init_local_group(hptb_ctx* ctx, int nranks, hptb_comm** comms);
hptb_status hptb_comm_destroy(hptb_comm* comm);
/* 1 if accumulators are exchanged through peer-mapped mailboxes (CUDA IPC over NVLink, inside the reduce kernel),
 * 0 if every exchange goes through NCCL (IPC unavailable, or HPTB_NO_P2P=1 on any rank). */
int hptb_comm_uses_peer_memory(const hptb_comm* comm);
/* Outer-axis sharding helpers (pure host code, usable without a GPU).
 * hptb_shard_bounds: rank r of `world` owns rows [offset, offset+len) of an axis of length n — contiguous
 * blocks, the first n % world ranks one row longer.
 * hptb_shard_plan_reduce: what hptb_reduce_sharded does for (op, axes, shard_axis): whether the reduction crosses
 * the shard axis, which collective combines the per-rank partials and which post-op follows. */
typedef enum hptb_collective {
  HPTB_COLL_NONE = 0,           /* the reduced axes do not include the shard axis: purely local */
  /* how the k per-rank ACCUMULATORS combine (always in rank order, on every rank; mathematically the named all-reduce): */
  HPTB_COLL_ALLREDUCE_SUM = 1,  /* sum, sum_square, reducel1, nansum; mean (Σ, then ÷ GLOBAL count); logsumexp (Σexp, then ln);
                                   reducel2/3 (Σ|x|^p, then the root) */
  HPTB_COLL_ALLREDUCE_PROD = 2,
  HPTB_COLL_ALLREDUCE_MAX = 3,
  HPTB_COLL_ALLREDUCE_MIN = 4,
  HPTB_COLL_ALLGATHER_ARG = 5   /* argmax/argmin: (extreme value, global index) pairs, rank-ordered strict combine */
} hptb_collective;
typedef struct hptb_shard_plan {
  int32_t crosses;      /* 1 if shard_axis is among the reduced axes */
  int32_t collective;   /* hptb_collective */
  int32_t pre_exp;      /* 1: the accumulator is Σ exp(x) (logsumexp) */
  int32_t post_ln;      /* 1: ln() after the combine (logsumexp) */
  int32_t global_count; /* 1: the post step divides by the GLOBAL element count (mean) */
  int32_t post_root;    /* p = 2 or 3: the accumulator is Σ|x|^p, the p-th root follows the combine (reducel2/3) */
} hptb_shard_plan;
hptb_status hptb_shard_bounds(int64_t n, int world, int rank, int64_t* offset, int64_t* len);
hptb_status hptb_shard_plan_reduce(int op, const int32_t* axes, int naxes, int shard_axis, int world,
                                   hptb_shard_plan* plan);
/* Combine per-rank partials in place.  op = HPTB_SUM / HPTB_MAX / HPTB_MIN / HPTB_PROD.  One stream per comm. */
hptb_status hptb_allreduce(hptb_comm* comm, int op, hptb_tensor* inout, void* stream);
/* Reduction of a tensor sharded along axis `shard_axis` (each rank passes its own shard).  Reduces locally, then —
 * only if shard_axis is among `axes` — combines the per-rank ACCUMULATORS (f32 for f16/bf16/f32 inputs, (value, global
 * index) pairs for ARGMAX/ARGMIN, Σexp for LOGSUMEXP, the unrooted power sum for REDUCEL2/3) in rank order and applies
 * the op's post step once (÷ GLOBAL count for MEAN, ln, root): the result is rounded exactly like the single-GPU one
 * and is bit-identical on every rank.  With peer memory the exchange runs inside the reduce kernel's epilogue (one
 * launch per rank) or in one small follow-up kernel; otherwise over ncclAllGather.  `global_axis_len` is the full
 * length of the sharded axis, `shard_offset` this rank's start.  Collective: every rank of the comm must make the same
 * sequence of calls, all on ONE stream per comm. */
hptb_status hptb_reduce_sharded(hptb_comm* comm, int op, const hptb_tensor* shard, const int32_t* axes,
                                int naxes, int shard_axis, int64_t shard_offset, int64_t global_axis_len,
                                hptb_tensor* out, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* HPT_B200_H */
