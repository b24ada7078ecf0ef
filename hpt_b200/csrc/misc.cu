// misc.cu — fill (`set_val_<T>` / `fill_<T>`, hpt-cudakernels/src/set_val.cu, creation.cu) for any output layout.
#include "context.h"
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

}  // namespace
}  // namespace hptb

using namespace hptb;

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
