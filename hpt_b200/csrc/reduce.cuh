// reduce.cuh — axis reductions: op traits, the two kernel skeletons and their launcher.
//
// Replaces the eight-kernels-per-op families of hpt-cudakernels/src/reduce/reduce_template.cuh and
// arg_template.cuh (`contiguous_<op>`, `<op>_fast_dim_include`, `{contiguous,uncontiguous}_<op>_fast_dim_only`,
// `*_small_fast_dim_only`, `<op>_fast_dim_no_include`) and the host planner that picks among them
// (hpt/src/backends/cuda/utils/reduce/reduce.rs:231-838).  After the collapse pass (layout.cpp) a
// reduction is a list of kept dims and reduced dims; two skeletons cover every case:
//
//   reduce_rows_kernel  the innermost reduced dim is walked by adjacent lanes (unit stride when the
//                       reduced axis is the contiguous one: full reduce, last-axis reduce, NCHW→C
//                       statistics, and the transposed-view axis-0 reduce of config 2).  128-bit loads,
//                       per-thread vector accumulators, warp shuffle + shared-memory tree.  A group
//                       of G = 1..32 lanes (small rows) or a whole CTA (long rows) owns one output;
//                       when there are too few outputs to fill the GPU each output is split over S
//                       CTAs whose partials are combined by the last CTA to arrive (fixed order →
//                       run-to-run deterministic, no float atomics) in the same launch.
//   reduce_cols_kernel  the contiguous dim is kept: lanes run along it with 128-bit loads and each
//                       thread strides over the reduced space; shared-memory tree over the CTA's
//                       rows, the same single-launch split/combine over CTAs.
//
// Accumulation: f32 for f16/bf16/f32 (the reference sums halves in half precision,
// reduce_classes.cuh:58-61), f64 for f64, wrapping integer arithmetic in T for ints, OR/AND for bool.
#pragma once
#include <cstdlib>

#include "common.h"
#include "context.h"
#include "launch.cuh"
#include "layout.h"
#include "promote.h"
#include "reduce_plan.h"
#include "scalar.cuh"
#include "xchg.cuh"

namespace hptb {

constexpr int kRedThreads = 256;
#ifndef HPTB_RED_UNROLL
#define HPTB_RED_UNROLL 4
#endif
#ifndef HPTB_LEAN_MINB
#define HPTB_LEAN_MINB 6
#endif
constexpr int kRedMaxDims = HPTB_MAX_DIMS;

// ---- op traits ---------------------------------------------------------------------------------------
template <typename V>
struct ArgPair {
  V val;
  int64_t idx;
};

template <typename T> struct is_argpair : std::false_type {};
template <typename V> struct is_argpair<ArgPair<V>> : std::true_type {};

template <int OP, typename T, typename Enable = void> struct ReduceOp;

// helpers in the compute type
template <typename C> __device__ __forceinline__ C red_add(C a, C b) {
  if constexpr (is_bool_t<C>::value) return b8{(uint8_t)(a.v | b.v)};
  else if constexpr (std::is_integral<C>::value) {
    typedef typename std::make_unsigned<C>::type U;
    return (C)(U)((U)a + (U)b);
  } else return a + b;
}
template <typename C> __device__ __forceinline__ C red_mul(C a, C b) {
  if constexpr (is_bool_t<C>::value) return b8{(uint8_t)(a.v & b.v)};
  else if constexpr (std::is_integral<C>::value) {
    typedef typename std::make_unsigned<C>::type U;
    if constexpr (sizeof(C) < 4) return (C)(U)((uint32_t)(U)a * (uint32_t)(U)b);
    else return (C)((U)a * (U)b);
  } else return a * b;
}
template <typename C> __device__ __forceinline__ C red_max(C a, C b) {
  if constexpr (is_bool_t<C>::value) return b8{(uint8_t)(a.v | b.v)};
  else if constexpr (std::is_same<C, float>::value) return fmaxf(a, b);
  else if constexpr (std::is_same<C, double>::value) return fmax(a, b);
  else return a > b ? a : b;
}
template <typename C> __device__ __forceinline__ C red_min(C a, C b) {
  if constexpr (is_bool_t<C>::value) return b8{(uint8_t)(a.v & b.v)};
  else if constexpr (std::is_same<C, float>::value) return fminf(a, b);
  else if constexpr (std::is_same<C, double>::value) return fmin(a, b);
  else return a < b ? a : b;
}
template <typename C> __device__ __forceinline__ C red_zero() {
  if constexpr (is_bool_t<C>::value) return b8{0};
  else return (C)0;
}
template <typename C> __device__ __forceinline__ C red_one() {
  if constexpr (is_bool_t<C>::value) return b8{1};
  else return (C)1;
}

// Per-thread running state.  Ordinary reductions keep an Acc; arg reductions keep (value, iteration number) and
// rebuild the 64-bit element index once, after the loop (PlainLocal / ArgOp below).
template <typename D, typename T, typename AccT>
struct PlainLocal {
  typedef AccT Local;
  static constexpr bool kTwoOutputs = false;
  static __device__ __forceinline__ Local local_identity() { return D::identity(); }
  static __device__ __forceinline__ void accumulate(Local& l, T x, int32_t) { l = D::combine(l, D::pre(x, 0)); }
  static __device__ __forceinline__ AccT finish(Local l, int64_t, int64_t, int, int) { return l; }
  // one loaded pack into the VEC per-slot accumulators (ops may override to pre-combine a pack more cheaply)
  template <int VEC>
  static __device__ __forceinline__ void accumulate_pack(Local (&acc)[VEC], const Pack<T, VEC>& v, int32_t it) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) D::accumulate(acc[k], v.v[k], it);
  }
};

// SUM: common_reduce.rs:32-52 (identity ZERO, combine _add), output dtype T
template <typename T> struct ReduceOp<HPTB_SUM, T> : PlainLocal<ReduceOp<HPTB_SUM, T>, T, compute_t<T>> {
  typedef T Out;
  typedef compute_t<T> Acc;
  static constexpr bool kIndexed = false;
  static __device__ __forceinline__ Acc identity() { return red_zero<Acc>(); }
  static __device__ __forceinline__ Acc pre(T x, int64_t) { return to_compute<T>(x); }
  static __device__ __forceinline__ Acc combine(Acc a, Acc b) { return red_add<Acc>(a, b); }
  static __device__ __forceinline__ Out post(Acc a, double) { return from_compute<T>(a); }
  static __device__ __forceinline__ Acc from_out(Out o) { return to_compute<T>(o); }
};
// PROD: common_reduce.rs:80-99
template <typename T> struct ReduceOp<HPTB_PROD, T> : PlainLocal<ReduceOp<HPTB_PROD, T>, T, compute_t<T>> {
  typedef T Out;
  typedef compute_t<T> Acc;
  static constexpr bool kIndexed = false;
  static __device__ __forceinline__ Acc identity() { return red_one<Acc>(); }
  static __device__ __forceinline__ Acc pre(T x, int64_t) { return to_compute<T>(x); }
  static __device__ __forceinline__ Acc combine(Acc a, Acc b) { return red_mul<Acc>(a, b); }
  static __device__ __forceinline__ Out post(Acc a, double) { return from_compute<T>(a); }
  static __device__ __forceinline__ Acc from_out(Out o) { return to_compute<T>(o); }
};
// SUM_SQUARE: hpt-traits/src/ops/reduce.rs:171-175 (Σ x², in T)
template <typename T> struct ReduceOp<HPTB_SUM_SQUARE, T> : PlainLocal<ReduceOp<HPTB_SUM_SQUARE, T>, T, compute_t<T>> {
  typedef T Out;
  typedef compute_t<T> Acc;
  static constexpr bool kIndexed = false;
  static __device__ __forceinline__ Acc identity() { return red_zero<Acc>(); }
  static __device__ __forceinline__ Acc pre(T x, int64_t) { Acc c = to_compute<T>(x); return red_mul<Acc>(c, c); }
  static __device__ __forceinline__ Acc combine(Acc a, Acc b) { return red_add<Acc>(a, b); }
  static __device__ __forceinline__ Out post(Acc a, double) { return from_compute<T>(a); }
  static __device__ __forceinline__ Acc from_out(Out o) { return to_compute<T>(o); }
};
// |x| in the compute type: wrapping for signed integers, identity for unsigned / bool (impls.rs:77-80,185-196)
template <typename C> __device__ __forceinline__ C red_abs(C a) {
  if constexpr (is_bool_t<C>::value) return a;
  else if constexpr (std::is_integral<C>::value) {
    if constexpr (std::is_signed<C>::value) {
      typedef typename std::make_unsigned<C>::type U;
      return a < 0 ? (C)(U)((U)0 - (U)a) : a;
    } else return a;
  } else if constexpr (std::is_same<C, float>::value) return fabsf(a);
  else return fabs(a);
}
template <typename C> __device__ __forceinline__ bool red_isnan(C a) {
  if constexpr (std::is_floating_point<C>::value) return a != a;
  else return false;
}
// REDUCEL1: Σ|x| in T (common_reduce.rs:147-167)
template <typename T> struct ReduceOp<HPTB_REDUCEL1, T> : PlainLocal<ReduceOp<HPTB_REDUCEL1, T>, T, compute_t<T>> {
  typedef T Out;
  typedef compute_t<T> Acc;
  static constexpr bool kIndexed = false;
  static __device__ __forceinline__ Acc identity() { return red_zero<Acc>(); }
  static __device__ __forceinline__ Acc pre(T x, int64_t) { return red_abs<Acc>(to_compute<T>(x)); }
  static __device__ __forceinline__ Acc combine(Acc a, Acc b) { return red_add<Acc>(a, b); }
  static __device__ __forceinline__ Out post(Acc a, double) { return from_compute<T>(a); }
  static __device__ __forceinline__ Acc from_out(Out o) { return to_compute<T>(o); }
};
// NANSUM / NANPROD: NaN counts as 0 / 1 (common_reduce.rs:255-324)
template <typename T> struct ReduceOp<HPTB_NANSUM, T> : PlainLocal<ReduceOp<HPTB_NANSUM, T>, T, compute_t<T>> {
  typedef T Out;
  typedef compute_t<T> Acc;
  static constexpr bool kIndexed = false;
  static __device__ __forceinline__ Acc identity() { return red_zero<Acc>(); }
  static __device__ __forceinline__ Acc pre(T x, int64_t) { const Acc c = to_compute<T>(x); return red_isnan<Acc>(c) ? red_zero<Acc>() : c; }
  static __device__ __forceinline__ Acc combine(Acc a, Acc b) { return red_add<Acc>(a, b); }
  static __device__ __forceinline__ Out post(Acc a, double) { return from_compute<T>(a); }
  static __device__ __forceinline__ Acc from_out(Out o) { return to_compute<T>(o); }
};
template <typename T> struct ReduceOp<HPTB_NANPROD, T> : PlainLocal<ReduceOp<HPTB_NANPROD, T>, T, compute_t<T>> {
  typedef T Out;
  typedef compute_t<T> Acc;
  static constexpr bool kIndexed = false;
  static __device__ __forceinline__ Acc identity() { return red_one<Acc>(); }
  static __device__ __forceinline__ Acc pre(T x, int64_t) { const Acc c = to_compute<T>(x); return red_isnan<Acc>(c) ? red_one<Acc>() : c; }
  static __device__ __forceinline__ Acc combine(Acc a, Acc b) { return red_mul<Acc>(a, b); }
  static __device__ __forceinline__ Out post(Acc a, double) { return from_compute<T>(a); }
  static __device__ __forceinline__ Acc from_out(Out o) { return to_compute<T>(o); }
};
// ALL / ANY: AND / OR of `x != 0` (NaN is true), bool output (common_reduce.rs:200-242)
template <typename T> struct ReduceOp<HPTB_ALL, T> : PlainLocal<ReduceOp<HPTB_ALL, T>, T, b8> {
  typedef b8 Out;
  typedef b8 Acc;
  static constexpr bool kIndexed = false;
  static __device__ __forceinline__ Acc identity() { return b8{1}; }
  static __device__ __forceinline__ Acc pre(T x, int64_t) { return cast<b8>(x); }
  static __device__ __forceinline__ Acc combine(Acc a, Acc b) { return b8{(uint8_t)(a.v & b.v)}; }
  static __device__ __forceinline__ Out post(Acc a, double) { return a; }
  static __device__ __forceinline__ Acc from_out(Out o) { return b8{(uint8_t)(o.v != 0)}; }
};
template <typename T> struct ReduceOp<HPTB_ANY, T> : PlainLocal<ReduceOp<HPTB_ANY, T>, T, b8> {
  typedef b8 Out;
  typedef b8 Acc;
  static constexpr bool kIndexed = false;
  static __device__ __forceinline__ Acc identity() { return b8{0}; }
  static __device__ __forceinline__ Acc pre(T x, int64_t) { return cast<b8>(x); }
  static __device__ __forceinline__ Acc combine(Acc a, Acc b) { return b8{(uint8_t)(a.v | b.v)}; }
  static __device__ __forceinline__ Out post(Acc a, double) { return a; }
  static __device__ __forceinline__ Acc from_out(Out o) { return b8{(uint8_t)(o.v != 0)}; }
};
// `count` value that asks REDUCEL2 / REDUCEL3 for the bare power sum (hptb_reduce_sharded exchanges Σ|x|^p, comm.cpp)
constexpr double kPartialPowerSum = -2.0;
// REDUCEL2 = sqrt Σ x², REDUCEL3 = (Σ |x|³)^(1/3), in FloatOutBinaryPromote<T,T> (common_reduce.rs:384-450; the
// exponent 1/3 is rounded to the output dtype first, as the reference's `(1.0 / 3.0).cast()` does)
template <typename T> struct ReduceOp<HPTB_REDUCEL2, T>
    : PlainLocal<ReduceOp<HPTB_REDUCEL2, T>, T, compute_t<typename type_of_dtype<promote_ct(dtype_of<T>::value, dtype_of<T>::value, 1)>::type>> {
  typedef typename type_of_dtype<promote_ct(dtype_of<T>::value, dtype_of<T>::value, 1)>::type Out;
  typedef compute_t<Out> Acc;
  static constexpr bool kIndexed = false;
  static __device__ __forceinline__ Acc identity() { return (Acc)0; }
  static __device__ __forceinline__ Acc pre(T x, int64_t) { const Acc c = to_compute<Out>(cast<Out>(x)); return c * c; }
  static __device__ __forceinline__ Acc combine(Acc a, Acc b) { return a + b; }
  static __device__ __forceinline__ Out post(Acc a, double n) {
    if (n == kPartialPowerSum) return from_compute<Out>(a);  // sharded: Σ x² leaves unrooted, the root follows the exchange
    if constexpr (std::is_same<Acc, float>::value) return from_compute<Out>(sqrtf(a));
    else return from_compute<Out>(sqrt(a));
  }
  static __device__ __forceinline__ Acc from_out(Out o) { const Acc c = to_compute<Out>(o); return c * c; }
};
template <typename T> struct ReduceOp<HPTB_REDUCEL3, T>
    : PlainLocal<ReduceOp<HPTB_REDUCEL3, T>, T, compute_t<typename type_of_dtype<promote_ct(dtype_of<T>::value, dtype_of<T>::value, 1)>::type>> {
  typedef typename type_of_dtype<promote_ct(dtype_of<T>::value, dtype_of<T>::value, 1)>::type Out;
  typedef compute_t<Out> Acc;
  static constexpr bool kIndexed = false;
  static __device__ __forceinline__ Acc identity() { return (Acc)0; }
  static __device__ __forceinline__ Acc pre(T x, int64_t) {
    const T ax = from_compute<T>(red_abs<compute_t<T>>(to_compute<T>(x)));  // |x| in T (wraps for the integer minimum), then cast
    const Acc c = to_compute<Out>(cast<Out>(ax));
    return c * c * c;
  }
  static __device__ __forceinline__ Acc combine(Acc a, Acc b) { return a + b; }
  static __device__ __forceinline__ Out post(Acc a, double n) {
    if (n == kPartialPowerSum) return from_compute<Out>(a);
    const Acc third = to_compute<Out>(cast<Out>(1.0 / 3.0));
    if constexpr (std::is_same<Acc, float>::value) return from_compute<Out>((float)pow((double)a, (double)third));
    else return from_compute<Out>(pow(a, third));
  }
  static __device__ __forceinline__ Acc from_out(Out o) { const Acc c = to_compute<Out>(o); return c * c * c; }
};
// MAX / MIN: common_reduce.rs:101-168 (identity NEG_INF / INF; f32::max/min ignore NaN)
template <typename T> struct ReduceOp<HPTB_MAX, T> : PlainLocal<ReduceOp<HPTB_MAX, T>, T, compute_t<T>> {
  typedef T Out;
  typedef compute_t<T> Acc;
  static constexpr bool kIndexed = false;
  static __device__ __forceinline__ Acc identity() { return Limits<Acc>::lowest(); }
  static __device__ __forceinline__ Acc pre(T x, int64_t) { return to_compute<T>(x); }
  static __device__ __forceinline__ Acc combine(Acc a, Acc b) { return red_max<Acc>(a, b); }
  static __device__ __forceinline__ Out post(Acc a, double) { return from_compute<T>(a); }
  static __device__ __forceinline__ Acc from_out(Out o) { return to_compute<T>(o); }
};
template <typename T> struct ReduceOp<HPTB_MIN, T> : PlainLocal<ReduceOp<HPTB_MIN, T>, T, compute_t<T>> {
  typedef T Out;
  typedef compute_t<T> Acc;
  static constexpr bool kIndexed = false;
  static __device__ __forceinline__ Acc identity() { return Limits<Acc>::highest(); }
  static __device__ __forceinline__ Acc pre(T x, int64_t) { return to_compute<T>(x); }
  static __device__ __forceinline__ Acc combine(Acc a, Acc b) { return red_min<Acc>(a, b); }
  static __device__ __forceinline__ Out post(Acc a, double) { return from_compute<T>(a); }
  static __device__ __forceinline__ Acc from_out(Out o) { return to_compute<T>(o); }
};
// MEAN: common_reduce.rs:352-380 — cast to FloatOutBinaryPromote<T,T>, Σ, ÷ n.  (The reference rounds n to
// the output dtype before dividing; here the division uses the exact count in the compute type.)
template <typename T> struct ReduceOp<HPTB_MEAN, T>
    : PlainLocal<ReduceOp<HPTB_MEAN, T>, T, compute_t<typename type_of_dtype<promote_ct(dtype_of<T>::value, dtype_of<T>::value, 1)>::type>> {
  typedef typename type_of_dtype<promote_ct(dtype_of<T>::value, dtype_of<T>::value, 1)>::type Out;
  typedef compute_t<Out> Acc;
  static constexpr bool kIndexed = false;
  static __device__ __forceinline__ Acc identity() { return (Acc)0; }
  static __device__ __forceinline__ Acc pre(T x, int64_t) { return to_compute<Out>(cast<Out>(x)); }
  static __device__ __forceinline__ Acc combine(Acc a, Acc b) { return a + b; }
  static __device__ __forceinline__ Out post(Acc a, double n) { return from_compute<Out>(a / (Acc)n); }
  static __device__ __forceinline__ Acc from_out(Out o) { return to_compute<Out>(o); }
};
// LOGSUMEXP: common_reduce.rs:451-480 — ln Σ exp(x), no max shift (as the reference; overflows to +inf alike)
// LONG (reductions of ≥ kLogSumExpLongMin elements, picked in api_reduce.cpp): exp is the bare ex2(x·log2e), 2
// instructions instead of 11.  Dropping the product's rounding error adds ≤ |x|·2^-24 relative error per term
// (|x| < 89 before overflow), i.e. ≤ 5e-6 ABSOLUTE on the logarithm — far inside the 1e-6·log2(n) sum bound.
constexpr int64_t kLogSumExpLongMin = 512;
template <typename T, bool LONG> struct LogSumExpOp
    : PlainLocal<LogSumExpOp<T, LONG>, T, compute_t<typename type_of_dtype<promote_ct(dtype_of<T>::value, dtype_of<T>::value, 1)>::type>> {
  typedef typename type_of_dtype<promote_ct(dtype_of<T>::value, dtype_of<T>::value, 1)>::type Out;
  typedef compute_t<Out> Acc;
  static constexpr bool kIndexed = false;
  static __device__ __forceinline__ Acc identity() { return (Acc)0; }
  static __device__ __forceinline__ Acc pre(T x, int64_t) {
    Acc c = to_compute<Out>(cast<Out>(x));
    if constexpr (std::is_same<Acc, float>::value) {
      if constexpr (LONG) {
        float e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(c * 1.4426950408889634f));  // −inf → 0, overflow → +inf, NaN → NaN
        return e;
      } else {
        return fast_expf_ovf(c);  // ≤ 2 ulp, 9 instructions (scalar.cuh)
      }
    } else {
      return exp(c);
    }
  }
  static __device__ __forceinline__ Acc combine(Acc a, Acc b) { return a + b; }
  static __device__ __forceinline__ Out post(Acc a, double) {
    if constexpr (std::is_same<Acc, float>::value) return from_compute<Out>(logf(a));
    else return from_compute<Out>(log(a));
  }
  static __device__ __forceinline__ Acc from_out(Out o) {  // fold a previous logsumexp value: exp it back
    Acc c = to_compute<Out>(o);
    if constexpr (std::is_same<Acc, float>::value) return expf(c);
    else return exp(c);
  }
};
template <typename T> struct ReduceOp<HPTB_LOGSUMEXP, T> : LogSumExpOp<T, false> {};
// ARGMAX / ARGMIN: cpu/kernels/argreduce_kernels.rs:13-21,49-57 — strict compare from NEG_INF / INF with
// index 0 as the start, so ties resolve to the lowest index, NaN never wins and an all-NaN (or all-identity)
// row yields 0.  A (value, index) pair with "better value, else lower index" is the associative form.
template <typename T, bool IS_MAX> struct ArgOp {
  typedef int64_t Out;
  typedef ArgPair<compute_t<T>> Acc;
  typedef compute_t<T> V;
  static constexpr bool kIndexed = true;
  static __device__ __forceinline__ Acc identity() {
    return Acc{IS_MAX ? Limits<V>::lowest() : Limits<V>::highest(), 0};
  }
  static __device__ __forceinline__ Acc pre(T x, int64_t i) { return Acc{to_compute<T>(x), i}; }
  static __device__ __forceinline__ bool better(V a, V b) {  // a strictly better than b
    if constexpr (is_bool_t<V>::value) return IS_MAX ? (a.v > b.v) : (a.v < b.v);
    else return IS_MAX ? (a > b) : (a < b);
  }
  static __device__ __forceinline__ bool equal(V a, V b) {
    if constexpr (is_bool_t<V>::value) return a.v == b.v;
    else return a == b;
  }
  static __device__ __forceinline__ Acc combine(Acc a, Acc b) {
    if (better(b.val, a.val) || (equal(b.val, a.val) && b.idx < a.idx)) return b;
    return a;
  }
  static __device__ __forceinline__ Out post(Acc a, double) { return a.idx; }
  static __device__ __forceinline__ Acc from_out(Out) { return identity(); }
  // running state of one thread: every accumulator sees strictly increasing indices, so a strict "better" test
  // keeps the first occurrence; only the 32-bit iteration number is tracked per element and the 64-bit index
  // (c0 + it·stride)·vec + k is rebuilt once.  it < 0: nothing beat the identity (all NaN / all -inf) → index 0.
  struct Local {
    V val;
    int32_t it;
  };
  static constexpr bool kTwoOutputs = false;
  static __device__ __forceinline__ Local local_identity() {
    return Local{IS_MAX ? Limits<V>::lowest() : Limits<V>::highest(), -1};
  }
  static __device__ __forceinline__ void accumulate(Local& l, T x, int32_t it) {
    const V v = to_compute<T>(x);
    if (better(v, l.val)) { l.val = v; l.it = it; }
  }
  static __device__ __forceinline__ Acc finish(Local l, int64_t c0, int64_t stride, int vec, int k) {
    if (l.it < 0) return identity();
    return Acc{l.val, (c0 + (int64_t)l.it * stride) * vec + k};
  }
  template <int VEC>
  static __device__ __forceinline__ void accumulate_pack(Local (&acc)[VEC], const Pack<T, VEC>& v, int32_t it) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) accumulate(acc[k], v.v[k], it);
  }
};
template <typename T> struct ReduceOp<HPTB_ARGMAX, T> : ArgOp<T, true> {};
template <typename T> struct ReduceOp<HPTB_ARGMIN, T> : ArgOp<T, false> {};

// ---- launch parameters ---------------------------------------------------------------------------------
struct DimWalk {
  int32_t n;
  uint32_t shape[kRedMaxDims];  // innermost first
  FastDiv div[kRedMaxDims];
  int64_t stride_a[kRedMaxDims];
  int64_t stride_b[kRedMaxDims];
};

struct RowsRedParams {
  DimWalk kept;        // outputs: stride_a = input strides, stride_b = output strides
  DimWalk outer;       // reduced dims other than the innermost one: stride_a = input strides
  int64_t M;           // outputs
  int64_t L;           // innermost reduced extent
  int64_t inner_stride;
  int64_t cpr;         // chunks per inner run
  int64_t chunks;      // chunks per output = (prod outer) * cpr
  int64_t S;           // splits per output: CTAs (G == kRedThreads) or warps (G == 32) that share one output
  int64_t chunks_per_split;
  double count;        // elements reduced per output
  FastDiv cpr_div;     // chunk → (outer index, column)
  FastDiv S_div;       // virtual row → (output, split) in warp mode
  int32_t G;           // threads per output: a power of two, 1..kRedThreads
  int32_t logG;
  int32_t use64;
  int32_t fold_out;    // 1: combine with the previous contents of out; 2 (kFoldRaw): store the bare accumulator
  XchgParams xchg;     // sharded reductions: exchange fused into the epilogue (G == kRedThreads only)
};

struct ColsRedParams {
  DimWalk kept;        // kept dims other than the contiguous one: stride_a input, stride_b output
  DimWalk red;         // reduced dims: stride_a input
  int64_t C;           // extent of the contiguous kept dim (unit stride in input and output)
  int64_t R;           // reduced elements per output
  int64_t K;           // prod(kept.shape)
  int64_t col_tiles;
  int64_t S, rows_per_split;
  double count;
  int32_t TX;          // lanes along C (power of two ≤ 32)
  int32_t use64;
  int32_t fold_out;
  int32_t pad;
};

__device__ __forceinline__ void walk2(int64_t idx, const DimWalk& w, int use64, int64_t& oa, int64_t& ob) {
  if (!use64) {
    uint32_t r = (uint32_t)idx;
#pragma unroll 1
    for (int i = 0; i < w.n; ++i) {
      uint32_t q = w.div[i].div(r);
      uint32_t rem = r - q * w.shape[i];
      oa += (int64_t)rem * w.stride_a[i];
      ob += (int64_t)rem * w.stride_b[i];
      r = q;
    }
  } else {
#pragma unroll 1
    for (int i = 0; i < w.n; ++i) {
      int64_t q = idx / (int64_t)w.shape[i];
      int64_t rem = idx - q * (int64_t)w.shape[i];
      oa += rem * w.stride_a[i];
      ob += rem * w.stride_b[i];
      idx = q;
    }
  }
}

template <typename V>
__device__ __forceinline__ V shfl_xor(V v, int off) {
  if constexpr (is_bool_t<V>::value) {
    return b8{(uint8_t)__shfl_xor_sync(0xffffffffu, (int)v.v, off)};
  } else if constexpr (std::is_arithmetic<V>::value && sizeof(V) < 4) {
    return (V)__shfl_xor_sync(0xffffffffu, (int)v, off);
  } else if constexpr (std::is_arithmetic<V>::value && sizeof(V) == 4) {
    return __shfl_xor_sync(0xffffffffu, v, off);
  } else {
    // 8-byte scalars and accumulator structs (ArgPair, mean/var pair): word-wise
    static_assert(sizeof(V) % 4 == 0, "accumulator size must be a multiple of 4 bytes");
    V o;
    uint32_t w[sizeof(V) / 4];
    memcpy(w, &v, sizeof(V));
#pragma unroll
    for (int i = 0; i < (int)(sizeof(V) / 4); ++i) w[i] = __shfl_xor_sync(0xffffffffu, w[i], off);
    memcpy(&o, w, sizeof(V));
    return o;
  }
}

// partials written by other CTAs are read through L2 (ld.global.cg): L1 is not coherent across SMs
template <typename A>
__device__ __forceinline__ A load_cg(const A* p) {
  typedef typename VecBytes<sizeof(A)>::type V;
  V raw = __ldcg(reinterpret_cast<const V*>(p));
  A r;
  memcpy(&r, &raw, sizeof(A));
  return r;
}

template <typename Op, typename Acc>
__device__ __forceinline__ Acc warp_reduce(Acc v, int width) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    if (off < width) {
      Acc o = shfl_xor<Acc>(v, off);
      // keep operand order independent of the lane so the result is identical on both sides
      const bool lower = (threadIdx.x & off) == 0;
      v = lower ? Op::combine(v, o) : Op::combine(o, v);
    }
  }
  return v;
}

// the last CTA (of S) to take a ticket finishes the output group; the counter resets itself for the next launch
__device__ __forceinline__ bool take_ticket(uint32_t* ticket, uint32_t S) {
  __shared__ uint32_t s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t old = atomicAdd(ticket, 1u);
    s_last = (old == S - 1);
    if (old == S - 1) *ticket = 0;
  }
  __syncthreads();
  bool last = s_last != 0;
  if (last) __threadfence();
  return last;
}

// final write of one output (or of the (mean, var) pair of the two-output extension op)
// fold = kFoldRaw: `out` is an array of ACCUMULATORS (sharded reductions whose kernel shape cannot fuse the exchange:
// the bare local accumulator goes to a scratch, xchg_combine_kernel exchanges, combines and applies post)
constexpr int kFoldRaw = 2;
// fold = kFoldRawAcc: combine into the accumulator already there (the head / tail pieces of a peeled reduction)
constexpr int kFoldRawAcc = 3;
template <typename Op>
__device__ __forceinline__ void red_store(typename Op::Out* out, typename Op::Out* out2, int64_t off, typename Op::Acc a,
                                          double count, int fold) {
  if constexpr (Op::kTwoOutputs) {
    Op::store2(out, out2, off, a, count);
  } else {
    if (fold >= kFoldRaw) {
      typename Op::Acc* r = reinterpret_cast<typename Op::Acc*>(out);
      r[off] = fold == kFoldRawAcc ? Op::combine(r[off], a) : a;
      return;
    }
    if (fold) a = Op::combine(Op::from_out(out[off]), a);
    out[off] = Op::post(a, count);
  }
}

// ---- rows kernel ---------------------------------------------------------------------------------------
// A group of G threads (G = 1..256, power of two; 256/G outputs per CTA) owns one output and walks its
// chunks with stride G; with S > 1 (few outputs, long rows; G = 256) CTA (m, s) reduces split s of output m
// and the last CTA to finish combines the S partials.  VEC elements per load when the inner run is
// unit-stride and aligned, else VEC = 1.
template <typename Op, typename T, int VEC>
__global__ void __launch_bounds__(kRedThreads)
reduce_rows_kernel(const T* __restrict__ in, typename Op::Out* __restrict__ out, typename Op::Out* __restrict__ out2,
                   typename Op::Acc* __restrict__ scratch, uint32_t* __restrict__ tickets, RowsRedParams p) {
  pdl_prologue();
  typedef typename Op::Acc Acc;
  typedef typename Op::Local Local;
  constexpr int UNROLL = HPTB_RED_UNROLL;
  __shared__ Acc s_part[kRedThreads / 32];
  const int tid = threadIdx.x;
  const int G = p.G;
  int64_t m, c_begin, c_end, split = 0;
  int g;
  if (p.S > 1 && G == kRedThreads) {
    m = blockIdx.x / p.S;  // 64-bit division once per CTA
    split = blockIdx.x - m * p.S;
    g = tid;
  } else {
    // virtual row v = (output, split): one group of G threads each (splits only with G == 32)
    const int64_t v = (((int64_t)blockIdx.x * kRedThreads) >> p.logG) + (tid >> p.logG);
    m = v;
    if (p.S > 1) {
      if (!p.use64) m = p.S_div.div((uint32_t)v);
      else m = v / p.S;
      split = v - m * p.S;
    }
    g = tid & (G - 1);
  }
  c_begin = split * p.chunks_per_split;
  c_end = c_begin + p.chunks_per_split;
  if (c_end > p.chunks) c_end = p.chunks;
  const bool active = m < p.M;
  int64_t in_off = 0, out_off = 0;
  if (active) walk2(m, p.kept, p.use64, in_off, out_off);

  Local acc[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) acc[k] = Op::local_identity();

  if (active) {
    // Hot loop.  It must stay far below the issue budget of an HBM-bound kernel (ncu, profiles/r01c: 86 warp
    // instructions per 16-byte pack with a divide per pack; the bf16 NCHW mean sat at 62 % issue utilisation),
    // so the (run, column) position of a thread is carried incrementally: one add and one compare per pack, the
    // run offset recomputed only when a run boundary is crossed.  No partial packs (the host picks VEC > 1 only
    // when the run length is a multiple of VEC); 32-bit local counters (a split never holds 2^31 chunks).
    const T* base = in + in_off;
    const uint32_t n_local = (uint32_t)(c_end - c_begin);
    const int64_t estride = VEC > 1 ? (int64_t)VEC : p.inner_stride;  // elements between consecutive chunks
    auto run_offset = [&](int64_t r) -> int64_t {
      if (p.outer.n == 1) return r * p.outer.stride_a[0];
      int64_t roff = 0, dummy = 0;
      walk2(r, p.outer, p.use64, roff, dummy);
      return roff;
    };
    uint32_t col, cpr_eff;
    int64_t r = 0;
    const T* runp;
    if (p.outer.n == 0) {  // one run per output: the split itself is the run, it never wraps
      col = (uint32_t)g;
      cpr_eff = 0xffffffffu;
      runp = base + c_begin * estride;
    } else {
      const int64_t cu0 = c_begin + g;
      r = cu0 / p.cpr;  // once per thread
      col = (uint32_t)(cu0 - r * p.cpr);
      cpr_eff = (uint32_t)p.cpr;  // the host guarantees cpr < 2^31 whenever outer dims exist
      runp = base + run_offset(r);
    }
    int32_t it = 0;
    for (uint32_t lc = (uint32_t)g; lc < n_local; lc += (uint32_t)G * UNROLL, it += UNROLL) {
      Pack<T, VEC> v[UNROLL];
      bool ok[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        ok[u] = lc + (uint32_t)u * (uint32_t)G < n_local;
        if (ok[u]) {
          const T* src = runp + (int64_t)col * estride;
          if constexpr (VEC > 1) load_pack<T, VEC>(v[u], src);
          else v[u].v[0] = load_one(src);
        }
        col += (uint32_t)G;
        if (col >= cpr_eff) {
          do { col -= cpr_eff; ++r; } while (col >= cpr_eff);
          runp = base + run_offset(r);
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        if (!ok[u]) continue;
        Op::template accumulate_pack<VEC>(acc, v[u], it + u);
      }
    }
  }
  // thread total; for arg reductions the element index of (iteration it, slot k) is ((c_begin + g) + it·G)·VEC + k
  // (exactly one reduced dim, so chunk number == column)
  Acc a = Op::finish(acc[0], c_begin + g, G, VEC, 0);
#pragma unroll
  for (int k = 1; k < VEC; ++k) a = Op::combine(a, Op::finish(acc[k], c_begin + g, G, VEC, k));

  if (G <= 32) {
    a = warp_reduce<Op, Acc>(a, G);
    if (p.S == 1) {
      if (active && g == 0) red_store<Op>(out, out2, out_off, a, p.count, p.fold_out);
      return;
    }
    // warp mode with splits (G == 32: the warp is uniform in m): fixed-slot partial, warp-level ticket, the
    // last warp to arrive combines the S partials in slot order → deterministic
    if (!active) return;
    const int lane = tid & 31;
    uint32_t last = 0;
    if (lane == 0) {
      scratch[m * p.S + split] = a;
      __threadfence();
      const uint32_t old = atomicAdd(tickets + m, 1u);
      last = old == (uint32_t)p.S - 1;
      if (last) tickets[m] = 0;  // self-reset for the next launch on this stream
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    __threadfence();
    Acc r = Op::identity();
    for (int64_t sidx = lane; sidx < p.S; sidx += 32) r = Op::combine(r, load_cg(scratch + m * p.S + sidx));
    r = warp_reduce<Op, Acc>(r, 32);
    if (lane == 0) red_store<Op>(out, out2, out_off, r, p.count, p.fold_out);
    return;
  }
  // G = 64 / 128 / 256: warps_per_out partials per output through shared memory
  a = warp_reduce<Op, Acc>(a, 32);
  if ((tid & 31) == 0) s_part[tid >> 5] = a;
  __syncthreads();
  if (g == 0) {
    const int w0 = tid >> 5, nw = G >> 5;
    a = s_part[w0];
    for (int w = 1; w < nw; ++w) a = Op::combine(a, s_part[w0 + w]);
  }
  if (p.S == 1) {
    if constexpr (!Op::kTwoOutputs) {
      if (p.xchg.enabled) {  // host: only with G == kRedThreads — one output per CTA, `active` is uniform
        if (active && tid < 32) {
          Acc t = s_part[0];
          for (int w = 1; w < kRedThreads / 32; ++w) t = Op::combine(t, s_part[w]);
          t = xchg_finish_warp<Op>(p.xchg, out_off, t);  // lane r ↔ rank r
          if (tid == 0) red_store<Op>(out, out2, out_off, t, p.count, p.fold_out);
        }
        return;
      }
    }
    if (active && g == 0) red_store<Op>(out, out2, out_off, a, p.count, p.fold_out);
    return;
  }
  // split outputs (G == kRedThreads, one output per CTA): fixed-slot partials + last-CTA combine → deterministic
  if (tid == 0) scratch[m * p.S + split] = a;
  if (take_ticket(tickets + m, (uint32_t)p.S)) {
    Acc r = Op::identity();
    for (int64_t s = tid; s < p.S; s += kRedThreads) r = Op::combine(r, load_cg(scratch + m * p.S + s));
    r = warp_reduce<Op, Acc>(r, 32);
    __syncthreads();
    if ((tid & 31) == 0) s_part[tid >> 5] = r;
    __syncthreads();
    if (tid < 32) {  // every lane of warp 0 forms the CTA total (the same value): lane r then talks to rank r
      r = s_part[0];
      for (int w = 1; w < kRedThreads / 32; ++w) r = Op::combine(r, s_part[w]);
      // sharded: this rank's accumulator goes to every peer, the k accumulators are combined in rank order (xchg.cuh)
      if constexpr (!Op::kTwoOutputs) {
        if (p.xchg.enabled) r = xchg_finish_warp<Op>(p.xchg, out_off, r);
      }
      if (tid == 0) red_store<Op>(out, out2, out_off, r, p.count, p.fold_out);
    }
  }
}


// ---- lean rows kernel ------------------------------------------------------------------------------------
// The common case — every output is ONE unit-stride, 16-byte-aligned run (last-axis reductions, the axis-0
// reduction of a transposed view, softmax statistics) and there are enough outputs to fill the GPU without
// splitting them — gets a kernel stripped to what that case needs: 32-bit counters, no run tracking, no split /
// ticket code.  tools/microbench/stream_reduce.cu (profiles/r01c_microbench_stream_reduce.txt) showed what
// matters for these 10–40 µs jobs: registers.  At ≤ 32 registers 8 CTAs (2048 threads) are resident per SM, so
// twice the bytes are in flight per SM compared with the 62-register general kernel, and short-lived CTAs that
// issue ALL of a thread's loads before the first use beat long-lived warps that walk a row in dependent round
// trips (f32 [4096,4096] row sums: 12.3 µs CTA-per-row vs 14.4 µs warp-per-row; [16384,16384]: 7.28 vs 6.54 TB/s).
struct LeanRowsParams {
  int64_t M;           // outputs
  int64_t in_stride;   // elements between the first runs of consecutive outputs
  int64_t out_stride;
  int64_t run_stride;  // MULTI: elements between consecutive runs of one output (the second reduced dim)
  double count;
  uint32_t cpr;        // 16-byte chunks per run
  uint32_t chunks;     // chunks per output (= cpr unless MULTI)
  int32_t logG;        // log2 of the threads per output (8 = a whole CTA)
  int32_t fold_out;
  int32_t reverse;     // walk the outputs from the last CTA's to the first (snake order, context.h pass_direction)
};

// resident CTAs per SM the register allocation must allow: 6 (≤ 40 registers) for plain 4-byte accumulators — 8
// (≤ 32 registers) spills 8–48 bytes per thread and measured no faster (profiles/r01d_sweep.txt, variant mb6)
template <typename Op, bool MULTI>
constexpr int lean_min_blocks() {
  return (MULTI || sizeof(typename Op::Local) > 8) ? 4 : (Op::kIndexed || sizeof(typename Op::Local) > 4) ? 5 : HPTB_LEAN_MINB;
}

// MULTI: an output is `chunks / cpr` runs of cpr chunks, run_stride apart (NCHW channel statistics: 64 runs of
// 3136 contiguous elements per channel).  A thread walks the flattened (run, column) chunk space with stride G and
// carries its position incrementally: one add and one compare per chunk.  Such outputs are long and usually few
// (512 channels → 3.5 CTAs per SM), so the MULTI kernel keeps 2·UNROLL loads in flight per thread instead of
// relying on resident CTAs for memory-level parallelism.
template <typename Op, typename T, int VEC, bool MULTI>
__global__ void __launch_bounds__(kRedThreads, lean_min_blocks<Op, MULTI>())
reduce_rows_lean_kernel(const T* __restrict__ in, typename Op::Out* __restrict__ out, typename Op::Out* __restrict__ out2,
                        LeanRowsParams p) {
  pdl_prologue();
  typedef typename Op::Acc Acc;
  typedef typename Op::Local Local;
  constexpr int UNROLL = MULTI ? 2 * HPTB_RED_UNROLL : HPTB_RED_UNROLL;
  __shared__ Acc s_part[kRedThreads / 32];
  const uint32_t tid = threadIdx.x;
  const uint32_t G = 1u << p.logG;
  const uint32_t g = tid & (G - 1);
  const uint32_t bx = p.reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const int64_t m = (((int64_t)bx * kRedThreads) >> p.logG) + (tid >> p.logG);
  const bool active = m < p.M;
  Local acc[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) acc[k] = Op::local_identity();
  if (active) {
    const T* row = in + m * p.in_stride;
    const uint32_t n = MULTI ? p.chunks : p.cpr;
    int32_t it = 0;
    if constexpr (!MULTI) {
      for (uint32_t c = g; c < n; c += G * UNROLL, it += UNROLL) {
        Pack<T, VEC> v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
          if (c + (uint32_t)u * G < n) load_pack<T, VEC>(v[u], row + (size_t)(c + (uint32_t)u * G) * VEC);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
          if (c + (uint32_t)u * G < n) Op::template accumulate_pack<VEC>(acc, v[u], it + u);
      }
    } else {
      const uint32_t cpr = p.cpr;
      uint32_t col = g % cpr;
      const T* runp = row + (int64_t)(g / cpr) * p.run_stride;
      for (uint32_t c = g; c < n; c += G * UNROLL) {
        Pack<T, VEC> v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          if (c + (uint32_t)u * G < n) load_pack<T, VEC>(v[u], runp + (size_t)col * VEC);
          col += G;
          while (col >= cpr) { col -= cpr; runp += p.run_stride; }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
          if (c + (uint32_t)u * G < n) Op::template accumulate_pack<VEC>(acc, v[u], 0);
      }
    }
  }
  Acc a = Op::finish(acc[0], g, G, VEC, 0);
#pragma unroll
  for (int k = 1; k < VEC; ++k) a = Op::combine(a, Op::finish(acc[k], g, G, VEC, k));
  if (G <= 32) {
    a = warp_reduce<Op, Acc>(a, (int)G);
    if (active && g == 0) red_store<Op>(out, out2, m * p.out_stride, a, p.count, p.fold_out);
    return;
  }
  a = warp_reduce<Op, Acc>(a, 32);
  if ((tid & 31) == 0) s_part[tid >> 5] = a;
  __syncthreads();
  if (g == 0 && active) {
    const int w0 = tid >> 5, nw = G >> 5;
    a = s_part[w0];
    for (int w = 1; w < nw; ++w) a = Op::combine(a, s_part[w0 + w]);
    red_store<Op>(out, out2, m * p.out_stride, a, p.count, p.fold_out);
  }
}

// ---- cols kernel ---------------------------------------------------------------------------------------
// The contiguous dim is kept: TX lanes run along it (VEC elements each), TY = 256/TX thread rows stride over
// the reduced space; shared-memory tree over the thread rows; S row-splits per column tile combined by the
// last CTA.
template <typename Op, typename T, int VEC>
__global__ void __launch_bounds__(kRedThreads)
reduce_cols_kernel(const T* __restrict__ in, typename Op::Out* __restrict__ out, typename Op::Out* __restrict__ out2,
                   typename Op::Acc* __restrict__ scratch, uint32_t* __restrict__ tickets, ColsRedParams p) {
  pdl_prologue();
  typedef typename Op::Acc Acc;
  typedef typename Op::Local Local;
  constexpr int UNROLL = HPTB_RED_UNROLL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Acc* sm = reinterpret_cast<Acc*>(smem_raw);  // [TY][TX*VEC]
  const int tid = threadIdx.x;
  const int TX = p.TX, TY = kRedThreads / TX;
  const int tx = tid & (TX - 1), ty = tid / TX;
  // blockIdx.x = (split * K + k) * col_tiles + tile: CTAs that run together read ADJACENT column segments of the
  // same rows (whole rows stream from DRAM page by page) rather than the same columns of far-apart row ranges
  int64_t b = blockIdx.x;
  const int64_t tile = b % p.col_tiles;
  b /= p.col_tiles;
  const int64_t k = b % p.K;
  const int64_t split = b / p.K;
  const int64_t col = (tile * TX + tx) * VEC;
  const bool col_ok = col < p.C;
  int32_t ncol = 0;
  if (col_ok) ncol = (p.C - col) >= VEC ? VEC : (int32_t)(p.C - col);
  int64_t in_off = 0, out_off = 0;
  walk2(k, p.kept, p.use64, in_off, out_off);
  const int64_t r_begin = split * p.rows_per_split;
  int64_t r_end = r_begin + p.rows_per_split;
  if (r_end > p.R) r_end = p.R;

  Local acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) acc[j] = Op::local_identity();
  if (col_ok) {
    const T* base = in + in_off + col;
    int32_t it = 0;
    for (int64_t r = r_begin + ty; r < r_end; r += (int64_t)TY * UNROLL, it += UNROLL) {
      Pack<T, VEC> v[UNROLL];
      bool ok[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int64_t ru = r + (int64_t)u * TY;
        ok[u] = ru < r_end;
        if (ok[u]) {
          int64_t roff = 0, dummy = 0;
          if (p.red.n == 1) roff = ru * p.red.stride_a[0];
          else walk2(ru, p.red, p.use64, roff, dummy);
          if (VEC > 1 && ncol == VEC) load_pack<T, VEC>(v[u], base + roff);
          else {
#pragma unroll
            for (int j = 0; j < VEC; ++j)
              if (j < ncol) v[u].v[j] = load_one(base + roff + j);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        if (!ok[u]) continue;
#pragma unroll
        for (int j = 0; j < VEC; ++j)
          if (j < ncol) Op::accumulate(acc[j], v[u].v[j], it + u);
      }
    }
  }
  // tree over ty in shared memory (arg reductions: row index of iteration it is (r_begin + ty) + it·TY)
  const int W = TX * VEC;
#pragma unroll
  for (int j = 0; j < VEC; ++j) sm[ty * W + tx * VEC + j] = Op::finish(acc[j], r_begin + ty, TY, 1, 0);
  __syncthreads();
  for (int h = TY >> 1; h > 0; h >>= 1) {
    if (ty < h) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        Acc x = sm[ty * W + tx * VEC + j], y = sm[(ty + h) * W + tx * VEC + j];
        sm[ty * W + tx * VEC + j] = Op::combine(x, y);
      }
    }
    __syncthreads();
  }
  if (p.S == 1) {
    if (ty == 0 && col_ok) {
      for (int j = 0; j < ncol; ++j) red_store<Op>(out, out2, out_off + col + j, sm[tx * VEC + j], p.count, p.fold_out);
    }
    return;
  }
  const int64_t group = k * p.col_tiles + tile;  // outputs sharing one ticket
  Acc* my = scratch + (group * p.S + split) * W;
  if (ty == 0) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) my[tx * VEC + j] = sm[tx * VEC + j];
  }
  if (take_ticket(tickets + group, (uint32_t)p.S)) {
    // thread (tx, ty): partials s = ty, ty+TY, … of its columns, then the same shared-memory tree
    Acc part[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) part[j] = Op::identity();
    for (int64_t s = ty; s < p.S; s += TY) {
      const Acc* src = scratch + (group * p.S + s) * W;
#pragma unroll
      for (int j = 0; j < VEC; ++j) part[j] = Op::combine(part[j], load_cg(src + tx * VEC + j));
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < VEC; ++j) sm[ty * W + tx * VEC + j] = part[j];
    __syncthreads();
    for (int h = TY >> 1; h > 0; h >>= 1) {
      if (ty < h) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          Acc x = sm[ty * W + tx * VEC + j], y = sm[(ty + h) * W + tx * VEC + j];
          sm[ty * W + tx * VEC + j] = Op::combine(x, y);
        }
      }
      __syncthreads();
    }
    if (ty == 0 && col_ok) {
      for (int j = 0; j < ncol; ++j) red_store<Op>(out, out2, out_off + col + j, sm[tx * VEC + j], p.count, p.fold_out);
    }
  }
}

// ---- lean cols kernel ------------------------------------------------------------------------------------
// The common cols case — ONE reduced dim (row stride `rs`), the kept contiguous dim a multiple of the pack width,
// 32-bit row counters — stripped to what it needs, for the same reason as the lean rows kernel: at ≤ 40 registers
// six CTAs are resident per SM, and a CTA walks only a short slab of rows (TY·UNROLL·4 ≈ 128), so the hardware
// scheduler keeps every SM streaming.  TX = 32 lanes × VEC columns (512 contiguous bytes of every row), TY = 8
// thread rows; shared-memory tree over the thread rows; row slabs are combined by the last CTA of a column tile
// (S ≤ 64 partials, read back from L2 in fixed order → deterministic).
struct LeanColsParams {
  DimWalk kept;          // kept dims other than the contiguous one: stride_a input, stride_b output
  int64_t rs;            // input stride of the reduced dim
  int64_t C;             // contiguous kept extent
  int64_t K;             // prod(kept.shape)
  double count;
  uint32_t R;            // reduced extent
  uint32_t rows_per_split;
  uint32_t S;
  uint32_t col_tiles;
  int32_t use64;
  int32_t fold_out;
  XchgParams xchg;       // sharded reductions: exchange fused into the epilogue
};

// VEC (value, index) pairs or 16-byte accumulators per thread need 64 registers: 4 CTAs/SM for those; 8-byte (value, index)
// pairs stay at 3 CTAs/SM — they fit 64 registers without spills, but i64 [4096,8192] argmin(0) then takes 74.9 µs instead of 63
template <typename Op, int VEC>
constexpr int lean_cols_min_blocks() {
  return sizeof(typename Op::Local) > 8 ? 3 : (Op::kIndexed || sizeof(typename Op::Local) * VEC > 16) ? 4 : HPTB_LEAN_MINB;
}

// Final write of a column tile, called by the WHOLE finishing CTA after a barrier; sm row 0 ([j][tx]) holds the tile's
// accumulators.  Sharded: thread row ty pushes the tile to rank ty and polls rank ty's entries — the k pushes and the k
// polls are in flight together (xchg.cuh) — the polled accumulators meet in shared memory ([rank][j][tx]) and thread
// row 0 combines them in rank order.
template <typename Op, int VEC>
__device__ __forceinline__ void lean_cols_finish(typename Op::Out* out, typename Op::Out* out2, int64_t off0, typename Op::Acc* sm, bool col_ok,
                                                 const LeanColsParams& p) {
  typedef typename Op::Acc Acc;
  constexpr int TX = 32, TY = kRedThreads / TX;
  const uint32_t tx = threadIdx.x & (TX - 1), ty = threadIdx.x / TX;
  if constexpr (!Op::kTwoOutputs) {
    if (p.xchg.enabled) {
      Acc mine[VEC], acc[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) mine[j] = sm[j * TX + tx];
      __syncthreads();  // row 0 has been read by everybody: the rows are free for the polled accumulators
      for (int r0 = 0; r0 < p.xchg.nranks; r0 += TY) {
        const int r = r0 + (int)ty;
        if (r < p.xchg.nranks && col_ok) {
#pragma unroll
          for (int j = 0; j < VEC; ++j) xchg_push_to<Op>(p.xchg, r, off0 + j, mine[j]);
#pragma unroll
          for (int j = 0; j < VEC; ++j) sm[(ty * VEC + j) * TX + tx] = xchg_poll_from<Op>(p.xchg, r, off0 + j);
        }
        __syncthreads();
        if (ty == 0 && col_ok) {
          const int nr = p.xchg.nranks - r0 < TY ? p.xchg.nranks - r0 : TY;
          for (int q = 0; q < nr; ++q) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
              const Acc pq = sm[(q * VEC + j) * TX + tx];
              acc[j] = (r0 == 0 && q == 0) ? pq : Op::combine(acc[j], pq);
            }
          }
        }
        __syncthreads();
      }
      if (ty == 0 && col_ok) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) red_store<Op>(out, out2, off0 + j, acc[j], p.count, p.fold_out);
      }
      return;
    }
  }
  if (ty == 0 && col_ok) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) red_store<Op>(out, out2, off0 + j, sm[j * TX + tx], p.count, p.fold_out);
  }
}

template <typename Op, typename T, int VEC>
__global__ void __launch_bounds__(kRedThreads, lean_cols_min_blocks<Op, VEC>())
reduce_cols_lean_kernel(const T* __restrict__ in, typename Op::Out* __restrict__ out, typename Op::Out* __restrict__ out2,
                        typename Op::Acc* __restrict__ scratch, uint32_t* __restrict__ tickets, LeanColsParams p) {
  pdl_prologue();
  typedef typename Op::Acc Acc;
  typedef typename Op::Local Local;
  // 8-byte (value, index) pairs run at 3 CTAs/SM (85 registers): twice the loads in flight per thread make up for the
  // missing CTAs
  constexpr int UNROLL = (Op::kIndexed && sizeof(Local) > 8 ? 2 : 1) * HPTB_RED_UNROLL;
  constexpr int TX = 32, TY = kRedThreads / TX, W = TX * VEC;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Acc* sm = reinterpret_cast<Acc*>(smem_raw);  // [TY][VEC][TX]: lanes are adjacent, so 8- and 16-byte accumulators are conflict-free
  const uint32_t tid = threadIdx.x, tx = tid & (TX - 1), ty = tid / TX;
  // blockIdx.x = (split · K + k) · col_tiles + tile: CTAs that run together read adjacent column segments of the same rows
  uint32_t b = blockIdx.x;
  const uint32_t tile = b % p.col_tiles;
  b /= p.col_tiles;
  const uint32_t k = b % (uint32_t)p.K;
  const uint32_t split = b / (uint32_t)p.K;
  const int64_t col = ((int64_t)tile * TX + tx) * VEC;
  const bool col_ok = col < p.C;
  int64_t in_off = 0, out_off = 0;
  if (p.kept.n) walk2(k, p.kept, p.use64, in_off, out_off);
  const uint32_t r_begin = split * p.rows_per_split;
  const uint32_t r_end = min(p.R, r_begin + p.rows_per_split);
  Local acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) acc[j] = Op::local_identity();
  if (col_ok) {
    const T* base = in + in_off + col;
    int32_t it = 0;
    for (uint32_t r = r_begin + ty; r < r_end; r += TY * UNROLL, it += UNROLL) {
      Pack<T, VEC> v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        if (r + (uint32_t)u * TY < r_end) load_pack<T, VEC>(v[u], base + (int64_t)(r + (uint32_t)u * TY) * p.rs);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        if (r + (uint32_t)u * TY < r_end) {
#pragma unroll
          for (int j = 0; j < VEC; ++j) Op::accumulate(acc[j], v[u].v[j], it + u);
        }
    }
  }
#pragma unroll
  for (int j = 0; j < VEC; ++j) sm[(ty * VEC + j) * TX + tx] = Op::finish(acc[j], r_begin + ty, TY, 1, 0);
  __syncthreads();
#pragma unroll
  for (int h = TY >> 1; h > 0; h >>= 1) {
    if (ty < h) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) sm[(ty * VEC + j) * TX + tx] = Op::combine(sm[(ty * VEC + j) * TX + tx], sm[((ty + h) * VEC + j) * TX + tx]);
    }
    __syncthreads();
  }
  if (p.S == 1) {
    lean_cols_finish<Op, VEC>(out, out2, out_off + col, sm, col_ok, p);
    return;
  }
  const uint32_t group = k * p.col_tiles + tile;
  Acc* my = scratch + ((size_t)group * p.S + split) * W;
  if (ty == 0) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) my[tx * VEC + j] = sm[j * TX + tx];
  }
  if (take_ticket(tickets + group, p.S)) {
    Acc part[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) part[j] = Op::identity();
    for (uint32_t sidx = ty; sidx < p.S; sidx += TY) {
      const Acc* src = scratch + ((size_t)group * p.S + sidx) * W;
#pragma unroll
      for (int j = 0; j < VEC; ++j) part[j] = Op::combine(part[j], load_cg(src + tx * VEC + j));
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < VEC; ++j) sm[(ty * VEC + j) * TX + tx] = part[j];
    __syncthreads();
#pragma unroll
    for (int h = TY >> 1; h > 0; h >>= 1) {
      if (ty < h) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) sm[(ty * VEC + j) * TX + tx] = Op::combine(sm[(ty * VEC + j) * TX + tx], sm[((ty + h) * VEC + j) * TX + tx]);
      }
      __syncthreads();
    }
    lean_cols_finish<Op, VEC>(out, out2, out_off + col, sm, col_ok, p);
  }
}

// ---- standalone exchange + combine ---------------------------------------------------------------------------
// The unfused half of a sharded reduction (xchg.cuh): `part` holds this rank's M bare accumulators (kFoldRaw), or —
// without peer memory — the k·M accumulators of every rank as all-gathered by NCCL ([rank][M]).  Push everything
// first, then collect: a push never waits, so no thread of any rank can block another's progress.  Rank-ordered
// combine → the result is bit-identical on every rank and equal to what the fused epilogue produces.
struct CombineParams {
  int64_t M;
  int32_t nk, gathered;
  int64_t shape[kRedMaxDims], stride[kRedMaxDims];  // `out` dims, innermost first (M counts them row-major)
  double count;
  XchgParams xchg;
};
template <typename Op>
__global__ void __launch_bounds__(256) xchg_combine_kernel(const typename Op::Acc* __restrict__ part, typename Op::Out* __restrict__ out,
                                                           CombineParams p) {
  pdl_prologue();
  typedef typename Op::Acc Acc;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, step = (int64_t)gridDim.x * blockDim.x;
  if (!p.gathered)
    for (int64_t m = i0; m < p.M; m += step) xchg_push<Op>(p.xchg, m, part[m]);
  for (int64_t m = i0; m < p.M; m += step) {
    Acc a;
    if (p.gathered) {
      a = part[m];
      for (int r = 1; r < p.gathered; ++r) a = Op::combine(a, part[(int64_t)r * p.M + m]);
    } else {
      a = xchg_collect<Op>(p.xchg, m);
    }
    int64_t rest = m, off = 0;
    for (int i = 0; i < p.nk; ++i) {
      const int64_t q = rest / p.shape[i];
      off += (rest - q * p.shape[i]) * p.stride[i];
      rest = q;
    }
    out[off] = Op::post(a, p.count);
  }
}

template <typename Op>
hptb_status launch_combine(const ReducePlan& plan, cudaStream_t stream) {
  if constexpr (Op::kTwoOutputs) {
    return fail(HPTB_ERR_UNSUPPORTED, "sharded exchange of a two-output reduction");
  } else {
    if (plan.comb_M <= 0) return HPTB_OK;
    CombineParams p;
    memset(&p, 0, sizeof(p));
    p.M = plan.comb_M;
    p.nk = plan.comb_nk;
    p.gathered = plan.gathered;
    for (int i = 0; i < plan.comb_nk; ++i) { p.shape[i] = plan.comb_shape[i]; p.stride[i] = plan.comb_stride[i]; }
    p.count = plan.count;
    if (!plan.gathered) {
      if (!plan.xchg) return fail(HPTB_ERR_INVALID, "combine: no exchange parameters");
      p.xchg = *plan.xchg;
      p.xchg.enabled = 1;
    }
    int64_t blocks = (p.M + 255) / 256;
    const int64_t cap = (int64_t)plan.ctx->sm_count * 4;
    if (blocks > cap) blocks = cap;
    HPTB_CUDA_CHECK(launch_kernel(xchg_combine_kernel<Op>, dim3((unsigned)blocks), dim3(256), 0, stream,
                                  static_cast<const typename Op::Acc*>(plan.in), static_cast<typename Op::Out*>(plan.out), p));
    return HPTB_OK;
  }
}

// ---- host launcher -------------------------------------------------------------------------------------------
inline bool red_fits_u32(int64_t v) { return v >= 0 && v < (int64_t(1) << 31); }

// Launch-shape overrides for tuning sweeps (tools/sweep.py): HPTB_TUNE_G = threads per output of the rows
// kernel, HPTB_TUNE_S = CTAs per output / per column tile.  Read per launch, only when HPTB_TUNE=1 at load.
inline int64_t tune_knob(const char* name) {
  static const bool on = [] { const char* e = getenv("HPTB_TUNE"); return e && e[0] == '1'; }();
  if (!on) return 0;
  const char* e = getenv(name);
  return e ? atoll(e) : 0;
}

inline void fill_walk(DimWalk& w, const Collapsed& c, const int* dims, int n, bool with_out, bool& big) {
  w.n = n;
  for (int i = 0; i < n; ++i) {
    int d = dims[i];
    if (!red_fits_u32(c.shape[d])) big = true;
    w.shape[i] = (uint32_t)c.shape[d];
    w.div[i] = FastDiv((uint32_t)c.shape[d]);
    w.stride_a[i] = c.strides[1][d];
    w.stride_b[i] = with_out ? c.strides[0][d] : 0;
  }
}

// An empty reduced extent (some reduced dim has length 0): every output is the op's value on no elements — the
// identity through post() (sum 0, prod 1, max NEG_INF, all true, argmax 0, mean 0/0 = NaN), as the reference's
// init value left untouched by an empty loop (cpu/tensor_internal/common_reduce.rs:32-168).
struct EmptyRedParams {
  int64_t M;
  int32_t nk, fold_out;
  int64_t shape[kRedMaxDims], stride[kRedMaxDims];  // kept dims, innermost first
  double count;
};
template <typename Op>
__global__ void reduce_empty_kernel(typename Op::Out* __restrict__ out, typename Op::Out* __restrict__ out2, EmptyRedParams p) {
  pdl_prologue();
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < p.M; m += (int64_t)gridDim.x * blockDim.x) {
    int64_t rest = m, off = 0;
    for (int i = 0; i < p.nk; ++i) {
      const int64_t q = rest / p.shape[i];
      off += (rest - q * p.shape[i]) * p.stride[i];
      rest = q;
    }
    red_store<Op>(out, out2, off, Op::identity(), p.count, p.fold_out);
  }
}

// resident CTAs per SM of one kernel instantiation (registers / shared memory decide), asked once
template <typename K>
inline int ctas_per_sm(K kernel, size_t smem) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kRedThreads, smem) != cudaSuccess || n < 1) {
    cudaGetLastError();
    n = 1;
  }
  return n;
}

// Launch-shape policy.  These kernels are HBM-bound, so what matters is that every SM streams from the first
// cycle to the last: a grid slightly larger than the number of resident CTA slots runs a second, mostly empty
// wave (1.3 waves = 65 % efficiency).  So: either ONE balanced wave of fat threads (grid ≤ slots) or many
// (≥ 4) waves of small CTAs, never in between.
template <typename Op, typename T>
hptb_status launch_reduce(const ReducePlan& plan, cudaStream_t stream) {
  typedef typename Op::Acc Acc;
  typedef typename Op::Out Out;
  if (plan.acc_bytes) *plan.acc_bytes = sizeof(Acc);
  if (plan.mode == kPlanCombine) return launch_combine<Op>(plan, stream);
  const Collapsed& c = plan.c;
  const T* in = static_cast<const T*>(plan.in);
  Out* out = static_cast<Out*>(plan.out);
  Out* out2 = static_cast<Out*>(plan.out2);
  constexpr int VECMAX = 16 / sizeof(T) > 8 ? 8 : 16 / sizeof(T);
  const int sms = plan.ctx->sm_count;
  // Sharded reductions (plan.raw_out set; plan.xchg = peer mailboxes, if any): a kernel shape whose final accumulators sit in at most half of the resident CTA
  // slots exchanges them itself (xchg.cuh: one launch per rank); any other shape stores bare accumulators to
  // plan.raw_out (same element offsets as `out`) and xchg_combine_kernel follows.  `settle(fuse)` is called by
  // every path right before its launch.
  const bool want_x = plan.raw_out != nullptr && !Op::kTwoOutputs;
  int fold = plan.fold_out;
  XchgParams xp;
  memset(&xp, 0, sizeof(xp));
  if (plan.fused) *plan.fused = false;
  auto settle = [&](bool fuse) {
    if (!want_x) return;
    if (fuse && plan.out && plan.xchg) {
      xp = *plan.xchg;
      xp.enabled = 1;
      if (plan.fused) *plan.fused = true;
    } else {
      out = reinterpret_cast<Out*>(plan.raw_out);
      fold = plan.fold_out == kFoldRawAcc ? kFoldRawAcc : kFoldRaw;
    }
  };

  // split dims (innermost first lists)
  int kept[kRedMaxDims], red[kRedMaxDims], nk = 0, nr = 0;
  for (int d = c.ndim - 1; d >= 0; --d) {
    if (c.reduced[d]) red[nr++] = d; else kept[nk++] = d;
  }
  int64_t M = 1, R = 1;
  for (int i = 0; i < nk; ++i) M *= c.shape[kept[i]];
  for (int i = 0; i < nr; ++i) R *= c.shape[red[i]];
  if (M == 0) return HPTB_OK;  // no outputs
  if (R == 0) {
    EmptyRedParams e;
    memset(&e, 0, sizeof(e));
    e.M = M;
    e.nk = nk;
    settle(false);
    e.fold_out = fold;
    e.count = plan.count;
    for (int i = 0; i < nk; ++i) { e.shape[i] = c.shape[kept[i]]; e.stride[i] = c.strides[0][kept[i]]; }
    const int64_t blocks = (M + 255) / 256;
    HPTB_CUDA_CHECK(launch_kernel(reduce_empty_kernel<Op>, dim3((unsigned)(blocks > 4096 ? 4096 : blocks)), dim3(256), 0, stream, out, out2, e));
    return HPTB_OK;
  }

  // ---- cols: the unit-stride dim is kept (and is the output's unit-stride dim) ---------------------------
  int cdim = -1;
  for (int i = 0; i < nk; ++i)
    if (c.strides[1][kept[i]] == 1 && c.strides[0][kept[i]] == 1) { cdim = kept[i]; break; }
  bool inner_red_unit = nr > 0 && c.strides[1][red[0]] == 1;
  if (cdim >= 0 && !inner_red_unit && nr > 0 && c.shape[cdim] >= 4) {
    ColsRedParams p;
    memset(&p, 0, sizeof(p));
    bool big = false;
    int kd[kRedMaxDims], nkd = 0;
    for (int i = 0; i < nk; ++i) if (kept[i] != cdim) kd[nkd++] = kept[i];
    fill_walk(p.kept, c, kd, nkd, true, big);
    fill_walk(p.red, c, red, nr, false, big);
    p.C = c.shape[cdim];
    p.R = R;
    p.K = M / p.C;
    p.count = plan.count;
    p.fold_out = plan.fold_out;
    // vector width: alignment of base pointers and of every stride that moves the row start
    int vec = VECMAX;
    auto aligned = [&](int v) {
      if (v == 1) return true;
      size_t ain = sizeof(T) * v > 16 ? 16 : sizeof(T) * v;
      if (reinterpret_cast<uintptr_t>(in) % ain) return false;
      if (p.C % v) return false;
      for (int d = 0; d < c.ndim; ++d) {
        if (d == cdim) continue;
        if ((uint64_t)(std::llabs(c.strides[1][d]) * (int64_t)sizeof(T)) % ain) return false;
      }
      return true;
    };
    if (!aligned(vec) || p.C < vec) vec = 1;
    // lean path: one reduced dim, full-width tiles of 32 lanes × VECMAX columns, 32-bit counters
    if constexpr (VECMAX > 1) {
      if (vec == VECMAX && nr == 1 && p.C >= 16 * VECMAX && red_fits_u32(R) && red_fits_u32(p.K) && !big &&
          !tune_knob("HPTB_TUNE_NOLEAN")) {
        constexpr int LTX = 32, LTY = kRedThreads / LTX;
        LeanColsParams q;
        memset(&q, 0, sizeof(q));
        q.kept = p.kept;
        q.rs = c.strides[1][red[0]];
        q.C = p.C;
        q.K = p.K;
        q.count = plan.count;
        q.R = (uint32_t)R;
        q.fold_out = plan.fold_out;
        const int64_t ctiles = (p.C + (int64_t)LTX * VECMAX - 1) / ((int64_t)LTX * VECMAX);
        const int64_t lgroups = p.K * ctiles;
        // Row slabs: about ONE wave of CTAs at 4 per SM (each CTA pays a shared-memory tree, a partial write and a
        // ticket, so more, thinner slabs lose: f32 [8192,8192] max(0): 576 CTAs → 43.9 µs, 4096 CTAs → 47.8 µs;
        // [4096,4096] sum(0): 576 CTAs → 15.4 µs, 1024 → 18.0 µs), but never more than 4096 rows per slab
        // ([262144,16384] sum(0): 64 slabs → 7.28 TB/s, 5 slabs → 6.9 TB/s); at most 64 partials per column tile.
        // Long columns (R ≥ 32768, the shards of config 5): MANY thin slabs of ≥ 512 rows — 9+ waves of CTAs, so the
        // last wave is a few percent of the run ([32768,16384] sum(0): 8 slabs = 1.15 waves → 338 µs, 64 slabs → 297 µs;
        // [262144,16384]: 64 → 2367 µs, 128 → 2347 µs; profiles/r02b_sweep_shard.txt).
        int64_t S = ((int64_t)sms * 4 + lgroups / 2) / lgroups;
        const int64_t by_rows = R >= 32768 ? R / 512 : (R + 4095) / 4096;
        // kernels whose registers allow fewer than four CTAs per SM ((value, index) pairs of 8-byte types: three): keep
        // the grid inside ONE wave of what is really resident (i64 [4096,8192] argmin(0): 640 CTAs on 444 slots, 73.6 µs)
        static const int occ_s = ctas_per_sm(reduce_cols_lean_kernel<Op, T, VECMAX>, (size_t)kRedThreads * VECMAX * sizeof(Acc));
        if (occ_s < 4 && S > 1 && lgroups * S > (int64_t)sms * occ_s) S = ((int64_t)sms * occ_s) / lgroups > 0 ? ((int64_t)sms * occ_s) / lgroups : 1;
        if (S < by_rows) S = by_rows;
        if (S > (R >= 32768 ? 128 : 64)) S = R >= 32768 ? 128 : 64;
        const int64_t max_s = (R + (int64_t)LTY * HPTB_RED_UNROLL - 1) / ((int64_t)LTY * HPTB_RED_UNROLL);  // ≥ one batch per thread row
        if (S > max_s) S = max_s;
        if (S < 1) S = 1;
        if (int64_t t = tune_knob("HPTB_TUNE_S")) S = t > R ? R : t;
        int64_t rps = (R + S - 1) / S;
        S = (R + rps - 1) / rps;
        if (lgroups * S <= 0x7fffffffLL && ctiles <= 0x7fffffffLL) {
          q.rows_per_split = (uint32_t)rps;
          q.S = (uint32_t)S;
          q.col_tiles = (uint32_t)ctiles;
          q.use64 = 0;
          Scratch scratch;
          uint32_t* tickets = nullptr;
          if (S > 1) {
            HPTB_TRY(scratch.get(plan.ctx, (size_t)(lgroups * S * LTX * VECMAX) * sizeof(Acc), stream));
            tickets = ctx_tickets(plan.ctx, stream, (size_t)lgroups);
            if (!tickets) return fail(HPTB_ERR_OOM, "reduce: ticket buffer allocation failed");
          }
          const size_t lsmem = (size_t)kRedThreads * VECMAX * sizeof(Acc);
          static const int occ_l = ctas_per_sm(reduce_cols_lean_kernel<Op, T, VECMAX>, (size_t)kRedThreads * VECMAX * sizeof(Acc));
          settle(lgroups * 2 <= (int64_t)sms * occ_l);  // one finishing CTA per column tile
          q.fold_out = fold;
          q.xchg = xp;
          HPTB_CUDA_CHECK(launch_kernel(reduce_cols_lean_kernel<Op, T, VECMAX>, dim3((unsigned)(lgroups * S)), dim3(kRedThreads), lsmem, stream, in, out,
                                        out2, (Acc*)scratch.ptr, tickets, q));
          return HPTB_OK;
        }
      }
    }
    int TX = 32;
    while (TX > 1 && (int64_t)(TX / 2) * vec >= p.C) TX >>= 1;
    p.TX = TX;
    const int TY = kRedThreads / TX;
    p.col_tiles = (p.C + (int64_t)TX * vec - 1) / ((int64_t)TX * vec);
    const int64_t groups = p.K * p.col_tiles;
    const size_t smem = (size_t)kRedThreads * vec * sizeof(Acc);
    static const int occ_v = ctas_per_sm(reduce_cols_kernel<Op, T, (VECMAX > 1 ? VECMAX : 1)>, (size_t)kRedThreads * VECMAX * sizeof(Acc));
    static const int occ_1 = ctas_per_sm(reduce_cols_kernel<Op, T, 1>, (size_t)kRedThreads * sizeof(Acc));
    const int64_t slots = (int64_t)sms * (vec > 1 ? occ_v : occ_1);
    // row splits: at least one full wave when the column tiles alone cannot fill it, and no more than ~512 rows
    // per thread row (measured: 17 GB sum(axis 0) reaches 0.92 of peak with 37–74 splits vs 0.80 with 4)
    int64_t S = 1;
    if (groups < slots) S = slots / groups;
    {
      int64_t byrows = p.R / ((int64_t)TY * 512);
      if (byrows > S) S = byrows;
      int64_t maxS = (p.R + (int64_t)TY * 8 - 1) / ((int64_t)TY * 8);
      if (S > maxS) S = maxS;
      if (S > 4096) S = 4096;
      if (S < 1) S = 1;
    }
    if (int64_t t = tune_knob("HPTB_TUNE_S")) S = t > p.R ? p.R : t;
    p.rows_per_split = (p.R + S - 1) / S;
    S = (p.R + p.rows_per_split - 1) / p.rows_per_split;
    if (S < 1) S = 1;
    p.S = S;
    if (groups * S > 0x7fffffffLL) return fail(HPTB_ERR_UNSUPPORTED, "reduce: grid too large");
    if (groups * S >= (int64_t(1) << 31) || !red_fits_u32(p.R)) big = true;
    p.use64 = big ? 1 : 0;
    Scratch scratch;
    uint32_t* tickets = nullptr;
    if (S > 1) {
      HPTB_TRY(scratch.get(plan.ctx, (size_t)(groups * S * TX * vec) * sizeof(Acc), stream));
      tickets = ctx_tickets(plan.ctx, stream, (size_t)groups);
      if (!tickets) return fail(HPTB_ERR_OOM, "reduce: ticket buffer allocation failed");
    }
    unsigned grid = (unsigned)(groups * S);
    settle(false);
    p.fold_out = fold;
    if (vec == VECMAX && VECMAX > 1)
      HPTB_CUDA_CHECK(launch_kernel(reduce_cols_kernel<Op, T, (VECMAX > 1 ? VECMAX : 1)>, dim3(grid), dim3(kRedThreads), smem, stream, in, out, out2, (Acc*)scratch.ptr, tickets, p));
    else
      HPTB_CUDA_CHECK(launch_kernel(reduce_cols_kernel<Op, T, 1>, dim3(grid), dim3(kRedThreads), smem, stream, in, out, out2, (Acc*)scratch.ptr, tickets, p));
    return HPTB_OK;
  }

  // ---- rows ---------------------------------------------------------------------------------------------
  RowsRedParams p;
  memset(&p, 0, sizeof(p));
  bool big = false;
  fill_walk(p.kept, c, kept, nk, true, big);
  p.M = M;
  p.count = plan.count;
  p.fold_out = plan.fold_out;
  if (nr == 0) { p.L = 1; p.inner_stride = 1; }
  else { p.L = c.shape[red[0]]; p.inner_stride = c.strides[1][red[0]]; }
  if (nr > 1) fill_walk(p.outer, c, red + 1, nr - 1, false, big);
  int vec = 1;
  if (p.inner_stride == 1) {
    vec = VECMAX;
    auto aligned = [&](int v) {
      if (v == 1) return true;
      size_t ain = sizeof(T) * v > 16 ? 16 : sizeof(T) * v;
      if (reinterpret_cast<uintptr_t>(in) % ain) return false;
      if (p.L % v) return false;  // no partial packs in the vector kernel
      for (int d = 0; d < c.ndim; ++d) {
        if (nr > 0 && d == red[0]) continue;
        if ((uint64_t)(std::llabs(c.strides[1][d]) * (int64_t)sizeof(T)) % ain) return false;
      }
      return true;
    };
    if (!aligned(vec) || p.L < vec) vec = 1;
  }
  p.cpr = (p.L + vec - 1) / vec;
  int64_t router = 1;
  for (int i = 1; i < nr; ++i) router *= c.shape[red[i]];
  p.chunks = router * p.cpr;
  if (nr > 1 && !red_fits_u32(p.cpr)) return fail(HPTB_ERR_UNSUPPORTED, "reduce: a reduced run of more than 2^31 chunks next to other reduced dims");
  if (!red_fits_u32(p.cpr) || p.chunks >= (int64_t(1) << 32) || !red_fits_u32(M)) big = true;
  p.use64 = big ? 1 : 0;
  p.cpr_div = FastDiv(big ? 1u : (uint32_t)p.cpr);

  static const int occ_v = ctas_per_sm(reduce_rows_kernel<Op, T, (VECMAX > 1 ? VECMAX : 1)>, 0);
  static const int occ_1 = ctas_per_sm(reduce_rows_kernel<Op, T, 1>, 0);
  const int64_t cta_slots = (int64_t)sms * (vec > 1 ? occ_v : occ_1);
  const int64_t thread_slots = cta_slots * kRedThreads;
  // Launch shape (profiles/r01d_sweep.txt, tools/microbench/stream_reduce.cu).  What these HBM-bound kernels need
  // is every thread issuing one batch of UNROLL independent 16-byte loads as early as possible and the hardware
  // CTA scheduler — not a static partition — balancing the SMs: so G = the largest power of two that still gives
  // every thread a full batch (a whole CTA per output from 1024 chunks up), and outputs are split over S CTAs
  // only when the outputs alone cannot fill the resident CTA slots AND every split keeps ≥ 64 chunks per thread
  // (bf16 NCHW channel mean, 512 outputs × 25088 chunks: S = 1 → 38 µs, S = 2 → 42 µs, S = 10 → 55 µs; the full
  // sum of 17 GB: 8 waves of CTAs → 7.3 TB/s).  Arg reductions carry a (value, index) pair through the cross-warp
  // combine and prefer G = 64 (transposed f32 [8192,8192] argmax: G = 64 → 41.4 µs, G = 256 → 46.6 µs).
  // (value, index) pairs pay 16-byte shuffles per combine step: half the lanes per output, twice the batches per thread
  // (f32 [256,512,512] argmax(2): G = 32 → 59.2 µs, G = 8 → 38.2 µs; [65536,1024]: 59.9 → 35.7 µs; tools/g_sweep.py)
  const int64_t min_batches = Op::kIndexed ? 8 : 2;
  int64_t G = 1;
  while (G < kRedThreads && G * min_batches * HPTB_RED_UNROLL <= p.chunks) G <<= 1;
  int64_t S = 1;
  if (G == kRedThreads && M < cta_slots) {
    S = (8 * cta_slots + M - 1) / M;
    const int64_t maxC = p.chunks / (kRedThreads * 64);
    if (S > maxC) S = maxC;
    if (S > 8192) S = 8192;
    if (S < 1) S = 1;
    // a power of two: for the power-of-two extents that dominate in practice every split then starts on a large
    // power-of-two boundary, and ≈ 7 waves of CTAs keep the last wave short (profiles/r02c_sweep_shard.txt: the
    // full sum of [32768,16384] f32 305 µs at S = 4096 vs 315–318 at 4736 (8.0 waves) and 313 at 8192)
    while (S & (S - 1)) S &= S - 1;
  }
  if (Op::kIndexed && S == 1 && G > 64 && M * 64 >= thread_slots) G = 64;
  (void)thread_slots;
  if (int64_t t = tune_knob("HPTB_TUNE_G")) {
    G = 1;
    while (G < t && G < kRedThreads) G <<= 1;
    S = 1;
  }
  if (int64_t t = tune_knob("HPTB_TUNE_S")) {
    if (G == kRedThreads || G == 32) S = t > p.chunks ? p.chunks : t;
  }
  // lean path: unsplit outputs made of aligned unit-stride runs (one run, or runs along ONE more reduced dim),
  // ≤ 1 kept dim, 32-bit counters
  // sharded: one CTA per output and at most half of the resident slots spinning → exchange in the epilogue
  const bool fuse_rows = want_x && plan.out && plan.xchg && G == kRedThreads && M * 2 <= cta_slots;
  if constexpr (VECMAX > 1) {
    const bool multi = nr == 2;
    if (S == 1 && vec == VECMAX && nr >= 1 && nr <= 2 && nk <= 1 && !big && !(multi && Op::kIndexed) && p.chunks < (int64_t(1) << 31) &&
        !fuse_rows && !tune_knob("HPTB_TUNE_NOLEAN")) {
      int logG = 0;
      while ((1 << logG) < G) ++logG;
      const int64_t lean_blocks = (M + (kRedThreads >> logG) - 1) / (kRedThreads >> logG);
      if (lean_blocks <= 0x7fffffffLL) {
        LeanRowsParams q;
        memset(&q, 0, sizeof(q));
        q.M = M;
        q.in_stride = nk ? c.strides[1][kept[0]] : 0;
        q.out_stride = nk ? c.strides[0][kept[0]] : 0;
        q.run_stride = multi ? c.strides[1][red[1]] : 0;
        q.count = plan.count;
        q.cpr = (uint32_t)p.cpr;
        q.chunks = (uint32_t)p.chunks;
        q.logG = logG;
        settle(false);
        q.fold_out = fold;
        q.reverse = plan.reverse;
        if (multi)
          HPTB_CUDA_CHECK(launch_kernel(reduce_rows_lean_kernel<Op, T, VECMAX, true>, dim3((unsigned)lean_blocks), dim3(kRedThreads), 0, stream, in, out, out2, q));
        else
          HPTB_CUDA_CHECK(launch_kernel(reduce_rows_lean_kernel<Op, T, VECMAX, false>, dim3((unsigned)lean_blocks), dim3(kRedThreads), 0, stream, in, out, out2, q));
        return HPTB_OK;
      }
    }
  }
  p.chunks_per_split = (p.chunks + S - 1) / S;
  // split boundaries on 4 KB (256 chunks of 16 bytes): a boundary inside a 128-byte line makes two CTAs fetch it, and
  // odd split sizes were erratic (17 GB sum, S = 2960: 2603 µs unaligned, 2381 µs aligned; profiles/r02c_sweep_shard.txt)
  if (S > 1 && p.chunks_per_split >= 4096) p.chunks_per_split = (p.chunks_per_split + 255) / 256 * 256;
  if (int64_t al = tune_knob("HPTB_TUNE_CPS_ALIGN")) p.chunks_per_split = (p.chunks_per_split + al - 1) / al * al;
  S = (p.chunks + p.chunks_per_split - 1) / p.chunks_per_split;
  if (S < 1) S = 1;
  if (p.chunks_per_split >= (int64_t(1) << 31))
    return fail(HPTB_ERR_UNSUPPORTED, "reduce: more than 2^31 chunks per split (reduce fewer elements per output)");
  p.S = S;
  p.G = (int32_t)G;
  p.logG = 0;
  while ((1 << p.logG) < G) ++p.logG;
  const int64_t per_cta = kRedThreads / G;
  const int64_t vrows = M * S;  // virtual rows
  const int64_t blocks = G == kRedThreads ? vrows : (vrows + per_cta - 1) / per_cta;
  if (blocks > 0x7fffffffLL) return fail(HPTB_ERR_UNSUPPORTED, "reduce: grid too large");
  if (vrows >= (int64_t(1) << 32)) p.use64 = 1;
  p.S_div = FastDiv(p.use64 ? 1u : (uint32_t)S);
  Scratch scratch;
  uint32_t* tickets = nullptr;
  if (S > 1) {
    HPTB_TRY(scratch.get(plan.ctx, (size_t)(M * S) * sizeof(Acc), stream));
    tickets = ctx_tickets(plan.ctx, stream, (size_t)M);
    if (!tickets) return fail(HPTB_ERR_OOM, "reduce: ticket buffer allocation failed");
  }
  const unsigned grid = (unsigned)blocks;
  settle(fuse_rows);
  p.fold_out = fold;
  p.xchg = xp;
  if (vec == VECMAX && VECMAX > 1)
    HPTB_CUDA_CHECK(launch_kernel(reduce_rows_kernel<Op, T, (VECMAX > 1 ? VECMAX : 1)>, dim3(grid), dim3(kRedThreads), 0, stream, in, out, out2, (Acc*)scratch.ptr, tickets, p));
  else
    HPTB_CUDA_CHECK(launch_kernel(reduce_rows_kernel<Op, T, 1>, dim3(grid), dim3(kRedThreads), 0, stream, in, out, out2, (Acc*)scratch.ptr, tickets, p));
  return HPTB_OK;
}

}  // namespace hptb
