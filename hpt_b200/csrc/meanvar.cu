// meanvar.cu — fused single-read mean + population variance over a set of axes.
//
// EXTENSION: Hpt has no `var` (grep `fn var` finds nothing).  BASELINE.json config 3
// ("mean/var reduction over (0,2,3) with f32 accumulate") is served either by composing `mean` and
// `sum_square` (two reads) or by this kernel (one read).  Both outputs have dtype
// FloatOutBinaryPromote<T,T>, like `mean` (hpt/src/backends/cpu/tensor_internal/common_reduce.rs:352-380).
// Accumulation is (Σx, Σx²) in f64 so that E[x²] − E[x]² does not cancel catastrophically; the two
// f64 adds per element are far below the HBM-bound budget.  Runs on the same two reduction skeletons.
#include "dtypes_x.h"
#include "reduce.cuh"

namespace hptb {
hptb_status build_reduce_plan(hptb_ctx* ctx, const hptb_tensor* in, const int32_t* axes, int naxes,
                              const hptb_tensor* out, ReducePlan* plan);  // api_reduce.cpp
namespace {

struct SQ {
  double s, q;
};

// Two outputs (mean → out, variance → out2, same layout) written by the reduction kernel itself: one launch.
template <typename T> struct MeanVarOp : PlainLocal<MeanVarOp<T>, T, SQ> {
  typedef typename type_of_dtype<promote_ct(dtype_of<T>::value, dtype_of<T>::value, 1)>::type O;
  typedef O Out;
  typedef SQ Acc;
  static constexpr bool kIndexed = false;
  static constexpr bool kTwoOutputs = true;
  static __device__ __forceinline__ Acc identity() { return SQ{0.0, 0.0}; }
  static __device__ __forceinline__ Acc pre(T x, int64_t) {
    const double v = (double)to_compute<O>(cast<O>(x));
    return SQ{v, v * v};
  }
  static __device__ __forceinline__ Acc combine(Acc a, Acc b) { return SQ{a.s + b.s, a.q + b.q}; }
  // Half-precision inputs: x and x² are exact in f32 (8 / 11-bit significands), so the Σx and Σx² of ONE pack are
  // formed with f32 adds (≤ 7 roundings of 2^-24, far below the inputs' own precision) and only the two pack
  // sums go through the f64 pipe: 0.5 instead of 3 f64 instructions per element (ncu, profiles/r01c: the
  // all-f64 version ran at 0.40 of HBM peak against 0.70 for the plain mean).  f32 / f64 / integer inputs keep
  // per-element f64 accumulation — their squares are not exact in f32.
  template <int VEC>
  static __device__ __forceinline__ void accumulate_pack(SQ (&acc)[VEC], const Pack<T, VEC>& v, int32_t) {
    if constexpr (is_half<T>::value && VEC >= 2) {
      float s = 0.0f, q = 0.0f;
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const float x = to_compute<T>(v.v[k]);
        s += x;
        q = fmaf(x, x, q);
      }
      acc[0].s += (double)s;
      acc[0].q += (double)q;
    } else {
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[k] = combine(acc[k], pre(v.v[k], 0));
    }
  }
  static __device__ __forceinline__ void store2(Out* mean, Out* var, int64_t off, Acc a, double n) {
    const double m = a.s / n;
    double v = a.q / n - m * m;
    if (v < 0.0) v = 0.0;
    typedef compute_t<O> C;
    mean[off] = from_compute<O>((C)m);
    var[off] = from_compute<O>((C)v);
  }
  // unused by a two-output op (red_store never folds), present for the common interface
  static __device__ __forceinline__ Out post(Acc a, double n) { return from_compute<O>((compute_t<O>)(a.s / n)); }
  static __device__ __forceinline__ Acc from_out(Out) { return identity(); }
};

template <typename T>
hptb_status run(const ReducePlan& plan, cudaStream_t s) {
  return launch_reduce<MeanVarOp<T>, T>(plan, s);
}

}  // namespace
}  // namespace hptb

using namespace hptb;

extern "C" hptb_status hptb_mean_var(hptb_ctx* ctx, const hptb_tensor* in, const int32_t* axes, int naxes,
                                     hptb_tensor* mean_out, hptb_tensor* var_out, void* stream) {
  if (!ctx) return fail(HPTB_ERR_INVALID, "mean_var: null ctx");
  if (!axes && naxes) return fail(HPTB_ERR_INVALID, "mean_var: null axes");
  HPTB_TRY(validate_tensor(in, "mean_var in"));
  HPTB_TRY(validate_tensor(mean_out, "mean_var mean_out"));
  HPTB_TRY(validate_tensor(var_out, "mean_var var_out"));
  const int odt = kFloatOutBinary[in->dtype][in->dtype];
  if (mean_out->dtype != odt || var_out->dtype != odt)
    return fail(HPTB_ERR_DTYPE, "mean_var: outputs must be %s", dtype_name(odt));
  bool same = mean_out->ndim == var_out->ndim;
  for (int i = 0; same && i < mean_out->ndim; ++i) same = mean_out->shape[i] == var_out->shape[i];
  if (!same) return fail(HPTB_ERR_SHAPE, "mean_var: mean_out and var_out shapes differ");
  pass_direction(ctx, in->data, 0, false);  // a forward streaming pass (snake order, context.h)
  bool same_layout = true;
  for (int i = 0; i < mean_out->ndim; ++i)
    if (mean_out->shape[i] != 1 && mean_out->strides[i] != var_out->strides[i]) same_layout = false;
  // the kernel writes both results at the same element offset; a var_out with other strides is filled through
  // a temporary that has mean_out's layout, then gathered with hptb_copy
  hptb_tensor var_t = *var_out;
  Scratch tmp;
  if (!same_layout) {
    int64_t span = 1;
    for (int i = 0; i < mean_out->ndim; ++i) {
      if (mean_out->shape[i] == 0) return HPTB_OK;
      if (mean_out->strides[i] < 0) return fail(HPTB_ERR_UNSUPPORTED, "mean_var: negative strides in mean_out with differently laid out var_out");
      span += (mean_out->shape[i] - 1) * mean_out->strides[i];
    }
    HPTB_TRY(tmp.get(ctx, (size_t)span * dtype_size(odt), stream));
    var_t = *mean_out;
    var_t.data = tmp.ptr;
  }
  ReducePlan plan;
  HPTB_TRY(build_reduce_plan(ctx, in, axes, naxes, mean_out, &plan));
  plan.out2 = var_t.data;
  plan.fold_out = 0;
  hptb_status st = HPTB_ERR_DTYPE;
  {
    DeviceGuard g(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    switch (in->dtype) {
#define X(T, N, E) \
  case E: st = run<T>(plan, s); break;
      HPTB_FOR_DTYPES(X)
#undef X
      default: break;
    }
  }
  HPTB_TRY(st);
  count_launches(1);
  if (!same_layout) HPTB_TRY(hptb_copy(ctx, &var_t, var_out, stream));
  return HPTB_OK;
}
