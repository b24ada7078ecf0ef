// misc.cu — fill (`set_val_<T>` / `fill_<T>`, hpt-cudakernels/src/set_val.cu, creation.cu) for any output layout.
#include "context.h"
#include "dtypes_x.h"
#include "layout.h"
#include "reduce.cuh"
#include "scalar.cuh"

namespace hptb {
namespace {

template <typename U>
__global__ void __launch_bounds__(256) fill_contig_kernel(U* __restrict__ out, U v, int64_t n) {
  constexpr int VEC = 16 / sizeof(U);
  const int64_t nvec = n / VEC;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  Pack<U, VEC> pk;
#pragma unroll
  for (int k = 0; k < VEC; ++k) pk.v[k] = v;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) store_pack<U, VEC>(out + i * VEC, pk);
  for (int64_t i = nvec * VEC + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = v;
}

template <typename U>
__global__ void __launch_bounds__(256) fill_strided_kernel(U* __restrict__ out, U v, int64_t n, DimWalk w, int use64) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int64_t oa = 0, ob = 0;
    walk2(i, w, use64, oa, ob);
    out[oa] = v;
  }
}

template <typename U>
hptb_status fill_impl(hptb_ctx* ctx, const Collapsed& c, void* out, const void* scalar, cudaStream_t stream) {
  U v;
  memcpy(&v, scalar, sizeof(U));
  const int64_t n = c.numel;
  if (n == 0) return HPTB_OK;
  int64_t blocks = (n + 256 * 16 - 1) / (256 * 16);
  const int64_t cap = (int64_t)ctx->sm_count * 16;
  if (blocks > cap) blocks = cap;
  const bool contig = c.ndim == 0 || (c.ndim == 1 && c.strides[0][0] == 1);
  if (contig && reinterpret_cast<uintptr_t>(out) % 16 == 0) {
    fill_contig_kernel<U><<<(unsigned)blocks, 256, 0, stream>>>(static_cast<U*>(out), v, n);
  } else {
    DimWalk w;
    memset(&w, 0, sizeof(w));
    bool big = false;
    int dims[kRedMaxDims];
    for (int i = 0; i < c.ndim; ++i) dims[i] = c.ndim - 1 - i;
    w.n = c.ndim;
    for (int i = 0; i < c.ndim; ++i) {
      int d = dims[i];
      if (!red_fits_u32(c.shape[d])) big = true;
      w.shape[i] = (uint32_t)c.shape[d];
      w.div[i] = FastDiv((uint32_t)c.shape[d]);
      w.stride_a[i] = c.strides[0][d];
    }
    if (!red_fits_u32(n)) big = true;
    fill_strided_kernel<U><<<(unsigned)blocks, 256, 0, stream>>>(static_cast<U*>(out), v, n, w, big ? 1 : 0);
  }
  HPTB_CUDA_CHECK(cudaGetLastError());
  count_launches(1);
  return HPTB_OK;
}

// ---- creation ops (TensorCreator, hpt-traits/src/ops/creation.rs; CPU semantics
// hpt/src/backends/cpu/tensor_internal/normal_creation.rs:140-234) ------------------------------------------------
// arange / arange_step / linspace: out[i] = start._add(i.cast()._mul(step)) — every step rounds in T (wrapping for
// integers; half types: cast i to T, multiply in f32 and round, add in f32 and round; bool: OR of AND).
template <typename T>
__device__ __forceinline__ T arange_value(T start, T step, int64_t i) {
  typedef compute_t<T> C;
  const T ti = cast<T>((uint64_t)i);  // `usize as T`
  if constexpr (is_bool_t<T>::value) {
    return b8{(uint8_t)(start.v | (ti.v & step.v))};
  } else if constexpr (std::is_integral<T>::value) {
    typedef typename std::make_unsigned<T>::type U;
    U prod;
    if constexpr (sizeof(T) < 4) prod = (U)((uint32_t)(U)ti * (uint32_t)(U)step);
    else prod = (U)ti * (U)step;
    return (T)(U)((U)start + prod);
  } else if constexpr (std::is_same<C, float>::value) {
    const T prod = from_compute<T>(__fmul_rn(to_compute<T>(ti), to_compute<T>(step)));  // no FMA contraction
    return from_compute<T>(__fadd_rn(to_compute<T>(start), to_compute<T>(prod)));
  } else {
    return __dadd_rn(start, __dmul_rn(ti, step));
  }
}

template <typename T>
__global__ void __launch_bounds__(256) arange_kernel(T* __restrict__ out, T start, T step, int64_t n, int64_t stride) {
  pdl_prologue();
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) out[i * stride] = arange_value<T>(start, step, i);
}

// eye(n, m, k): 1 where col == row + k (normal_creation.rs:187-203)
template <typename T>
__global__ void __launch_bounds__(256) eye_kernel(T* __restrict__ out, int64_t n, int64_t m, int64_t k, int64_t s0, int64_t s1) {
  pdl_prologue();
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * m; i += gs) {
    const int64_t r = i / m, c = i - r * m;
    out[r * s0 + c * s1] = cast<T>(b8{(uint8_t)(c == r + k)});
  }
}

template <typename T>
hptb_status arange_impl(hptb_ctx* ctx, hptb_tensor* out, const void* start, const void* step, cudaStream_t stream) {
  T a, b;
  memcpy(&a, start, sizeof(T));
  memcpy(&b, step, sizeof(T));
  const int64_t n = out->shape[0];
  if (n == 0) return HPTB_OK;
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)ctx->sm_count * 16;
  if (blocks > cap) blocks = cap;
  HPTB_CUDA_CHECK(launch_kernel(arange_kernel<T>, dim3((unsigned)blocks), dim3(256), 0, stream, static_cast<T*>(out->data), a, b, n, out->strides[0]));
  count_launches(1);
  return HPTB_OK;
}

template <typename T>
hptb_status eye_impl(hptb_ctx* ctx, hptb_tensor* out, int64_t k, cudaStream_t stream) {
  const int64_t n = out->shape[0], m = out->shape[1];
  if (n * m == 0) return HPTB_OK;
  int64_t blocks = (n * m + 255) / 256;
  const int64_t cap = (int64_t)ctx->sm_count * 16;
  if (blocks > cap) blocks = cap;
  HPTB_CUDA_CHECK(launch_kernel(eye_kernel<T>, dim3((unsigned)blocks), dim3(256), 0, stream, static_cast<T*>(out->data), n, m, k, out->strides[0],
                                out->strides[1]));
  count_launches(1);
  return HPTB_OK;
}

}  // namespace
}  // namespace hptb

using namespace hptb;

extern "C" hptb_status hptb_arange(hptb_ctx* ctx, hptb_tensor* out, const void* start, const void* step, void* stream) {
  if (!ctx || !start || !step) return fail(HPTB_ERR_INVALID, "arange: null argument");
  HPTB_TRY(validate_tensor(out, "arange out"));
  if (out->ndim != 1) return fail(HPTB_ERR_SHAPE, "arange: out must be one-dimensional");
  DeviceGuard g(ctx->device);
  switch (out->dtype) {
#define X(T, N, E) \
  case E: return arange_impl<T>(ctx, out, start, step, (cudaStream_t)stream);
    HPTB_FOR_DTYPES(X)
#undef X
    default: return fail(HPTB_ERR_DTYPE, "arange: bad dtype");
  }
}

extern "C" hptb_status hptb_eye(hptb_ctx* ctx, hptb_tensor* out, int64_t k, void* stream) {
  if (!ctx) return fail(HPTB_ERR_INVALID, "eye: null ctx");
  HPTB_TRY(validate_tensor(out, "eye out"));
  if (out->ndim != 2) return fail(HPTB_ERR_SHAPE, "eye: out must be two-dimensional");
  DeviceGuard g(ctx->device);
  switch (out->dtype) {
#define X(T, N, E) \
  case E: return eye_impl<T>(ctx, out, k, (cudaStream_t)stream);
    HPTB_FOR_DTYPES(X)
#undef X
    default: return fail(HPTB_ERR_DTYPE, "eye: bad dtype");
  }
}

extern "C" hptb_status hptb_fill(hptb_ctx* ctx, hptb_tensor* out, const void* scalar, void* stream) {
  if (!ctx || !scalar) return fail(HPTB_ERR_INVALID, "fill: null argument");
  HPTB_TRY(validate_tensor(out, "fill out"));
  int64_t strides[kMaxOperands][HPTB_MAX_DIMS] = {{0}};
  for (int i = 0; i < out->ndim; ++i) strides[0][i] = out->strides[i];
  Collapsed c;
  collapse(out->ndim, out->shape, 1, strides, nullptr, &c);
  DeviceGuard g(ctx->device);
  switch (dtype_size(out->dtype)) {
    case 1: return fill_impl<uint8_t>(ctx, c, out->data, scalar, (cudaStream_t)stream);
    case 2: return fill_impl<uint16_t>(ctx, c, out->data, scalar, (cudaStream_t)stream);
    case 4: return fill_impl<uint32_t>(ctx, c, out->data, scalar, (cudaStream_t)stream);
    case 8: return fill_impl<uint64_t>(ctx, c, out->data, scalar, (cudaStream_t)stream);
    default: return fail(HPTB_ERR_DTYPE, "fill: bad dtype");
  }
}
