// sharded.cu — device helpers for reductions that cross the shard axis of an outer-axis-sharded tensor.
//
// argmax / argmin cannot be expressed as an allreduce: every rank reduces its shard to (extreme value, local
// index), turns the index into a GLOBAL one (+ the shard's offset along the axis), the k (value, index) arrays
// are all-gathered and combined here with the reference's rule — strict "better" from the identity with index 0,
// scanning shards in rank order, so ties resolve to the lowest global index, NaN never wins and an all-NaN /
// all-identity row yields 0 (hpt/src/backends/cpu/kernels/argreduce_kernels.rs:13-21,49-57).
#include "dtypes_x.h"
#include "reduce.cuh"

namespace hptb {
namespace {

template <typename T, bool IS_MAX>
__global__ void __launch_bounds__(256) arg_combine_kernel(const T* __restrict__ vals, const int64_t* __restrict__ idx, int k,
                                                          int64_t M, int64_t* __restrict__ out) {
  pdl_prologue();
  typedef ArgOp<T, IS_MAX> Op;
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  typename Op::V best = Op::identity().val;
  int64_t bi = 0;
  for (int r = 0; r < k; ++r) {
    const typename Op::V v = to_compute<T>(vals[(int64_t)r * M + m]);
    if (Op::better(v, best)) {
      best = v;
      bi = idx[(int64_t)r * M + m];
    }
  }
  out[m] = bi;
}

__global__ void __launch_bounds__(256) add_offset_kernel(int64_t* __restrict__ p, int64_t off, int64_t n) {
  pdl_prologue();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] += off;
}

template <typename T>
hptb_status combine_t(bool is_max, const void* vals, const int64_t* idx, int k, int64_t M, int64_t* out, cudaStream_t s) {
  const unsigned grid = (unsigned)((M + 255) / 256);
  if (is_max)
    HPTB_CUDA_CHECK(launch_kernel(arg_combine_kernel<T, true>, dim3(grid), dim3(256), 0, s, (const T*)vals, idx, k, M, out));
  else
    HPTB_CUDA_CHECK(launch_kernel(arg_combine_kernel<T, false>, dim3(grid), dim3(256), 0, s, (const T*)vals, idx, k, M, out));
  return HPTB_OK;
}

}  // namespace

hptb_status arg_combine(int dtype, bool is_max, const void* vals, const int64_t* idx, int k, int64_t M, int64_t* out, cudaStream_t s) {
  if (M <= 0) return HPTB_OK;
  if (M > (int64_t)0x7fffffff * 256) return fail(HPTB_ERR_UNSUPPORTED, "arg_combine: too many outputs");
  switch (dtype) {
#define X(T, N, E) \
  case E: return combine_t<T>(is_max, vals, idx, k, M, out, s);
    HPTB_FOR_DTYPES(X)
#undef X
    default: return fail(HPTB_ERR_DTYPE, "arg_combine: bad dtype");
  }
}

hptb_status add_offset_i64(int64_t* p, int64_t off, int64_t n, cudaStream_t s) {
  if (n <= 0 || off == 0) return HPTB_OK;
  HPTB_CUDA_CHECK(launch_kernel(add_offset_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, p, off, n));
  return HPTB_OK;
}

}  // namespace hptb
