// softmax.cu — softmax / log_softmax along one axis (NormalizationOps::{softmax, log_softmax},
// hpt-traits/src/ops/normalization.rs:51-64).
//
// Replaces `<T>_{softmax,logsoftmax}_{warp,block,block_large}[_uncontiguous]`
// (hpt-cudakernels/src/normalization/softmax.cu:19-318, f16/f32/f64 only there) and the host code in
// hpt/src/backends/cuda/tensor_internal/softmax.rs:85-….  Semantics follow the CPU kernel
// (hpt/src/backends/cpu/kernels/softmax.rs:204-310): y = exp(x − max) / Σ exp(x − max), computed in the
// Intermediate type of FloatOutUnaryPromote<T> (f32 for f16/bf16/ints ≤ 32 bit, f64 for 64-bit types).
//
// Kernels, picked after the collapse pass (launch_softmax):
//   softmax_rows_reg[_loop]   axis has unit stride and the row fits in registers (≤ 8 packs per thread): one warp
//                             (short rows), half a CTA (4-byte types, ≤ 512 packs) or one 256-thread CTA per row; ONE
//                             read and ONE write of the row with 128-bit accesses (the reference's block variant
//                             stages the row in shared memory and makes three passes over it); 8-pack rows persistent.
//   softmax_band_rows / _cols (softmax_band.cuh) rows of up to 8 × 96 KB, 128-byte column bands of up to 6144 positions:
//                             the band stays in the shared memory of a thread-block cluster — one read, exact maximum.
//   softmax_rows_stream_vec   longer unit-stride rows: statistics sweep + apply sweep (power-of-two frame, SmRun); few
//                             long rows (a 1-D softmax) are split into slabs over two launches.
//   softmax_cols_tiled        strided axis, another dim contiguous: column tiles × axis splits, the same two sweeps.
//   softmax_rows_stream, softmax_cols   scalar fallbacks (any strides, unaligned shapes, 64-bit compute type).
// A unit-stride axis whose OUTPUT axis is strided (x.t().softmax(0) into a fresh tensor) goes through a scratch in the
// input's order and the transposing copy (hptb_softmax below).
#include <cooperative_groups.h>
#include <array>

#include <algorithm>
#include "context.h"
#include "dtypes_x.h"
#include "layout.h"
#include "promote.h"
#include "reduce.cuh"
#include "scalar.cuh"

namespace hptb {
namespace {

constexpr int kSmThreads = 256;
constexpr int kSmChunks = 8;

struct SoftmaxParams {
  DimWalk kept;  // stride_a = input, stride_b = output
  int64_t M;     // rows
  int64_t L;     // axis length
  int64_t sa_in, sa_out;
  int32_t log;
  int32_t use64;
  int32_t nchunks;  // packs per thread (rows_reg)
  int32_t G;        // threads per row (32 or 256)
  // cols: the contiguous kept dim is kept.shape[0] with unit strides
  DimWalk outer;    // cols_tiled: the kept dims other than the contiguous one
  int64_t C;        // cols_tiled: extent of the contiguous kept dim
  int64_t ctiles;   // cols_tiled: column tiles per outer index
  int64_t rps;      // cols_tiled: axis positions per split (gridDim.y splits)
};

template <typename C> __device__ __forceinline__ C sm_exp(C x) {
  if constexpr (std::is_same<C, float>::value) return expf(x); else return exp(x);
}
// exp(x − m) with the rounding error of the subtraction folded back in.  x − m rounds to sh with an absolute
// error of up to ulp(sh)/2, which exp() turns into a RELATIVE error of |sh|·2^-24 (f32) — 18 ulp at sh = −36.
// TwoSum recovers that error exactly (lo) and exp(sh + lo) = exp(sh)·(1 + lo) to first order.
template <typename C> struct ShExp { C sh, ex; };
template <typename C> __device__ __forceinline__ ShExp<C> sm_shift_exp(C x, C m) {
  const C b = -m;
  const C sh = x + b;
  if (!(sh > Limits<C>::lowest())) return ShExp<C>{sh, sm_exp<C>(sh)};  // −inf / NaN: nothing to compensate
  const C xv = sh - b, bv = sh - xv;
  const C lo = (x - xv) + (b - bv);
  const C e = sm_exp<C>(sh);
  if constexpr (std::is_same<C, float>::value) return ShExp<C>{sh, fmaf(e, lo, e)};
  else return ShExp<C>{sh, fma(e, lo, e)};
}
template <typename C> __device__ __forceinline__ C sm_log(C x) {
  if constexpr (std::is_same<C, float>::value) return logf(x); else return log(x);
}
template <typename C> __device__ __forceinline__ C sm_max(C a, C b) {
  if constexpr (std::is_same<C, float>::value) return fmaxf(a, b); else return fmax(a, b);
}

// exp for the register-resident kernel: fast_expf (scalar.cuh)
__device__ __forceinline__ float sm_exp_fast(float x) { return fast_expf(x); }
__device__ __forceinline__ double sm_exp_fast(double x) { return exp(x); }
// exp for a result that is rounded to a 16-BIT output type (and summed in f32): one multiply and one ex2.  The rounding
// of x·log2e costs |x|·2^-24 relative — under 1e-5 for any argument whose exp is not zero — against an output ulp of
// 2^-11 (f16) or 2^-8 (bf16); the compensated fast_expf is six more instructions per element, and the 16-bit kernels
// (8 elements per 16-byte pack) are bound by exactly that instruction stream (bf16 [32768,4096] softmax: 113 µs
// against 82 µs of HBM time).  −inf → 0, NaN propagates.
__device__ __forceinline__ float sm_exp_half(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
  return e;
}
// by OUTPUT type
template <typename O> __device__ __forceinline__ float sm_exp_o(float x) {
  if constexpr (sizeof(O) == 2) return sm_exp_half(x); else return fast_expf(x);
}
template <typename O> __device__ __forceinline__ double sm_exp_o(double x) { return exp(x); }

struct MaxOp {
  template <typename C> static __device__ __forceinline__ C combine(C a, C b) { return sm_max<C>(a, b); }
};
struct AddOp {
  template <typename C> static __device__ __forceinline__ C combine(C a, C b) { return a + b; }
};
template <typename OpT, typename C>
struct WrapOp {
  static __device__ __forceinline__ C combine(C a, C b) { return OpT::template combine<C>(a, b); }
};

// reduce over the G threads that own a row (G = 32: warp shuffle; G = 128 / 256: + shared memory)
template <typename OpT, typename C, int G>
__device__ __forceinline__ C group_reduce(C v, C* s_buf, C ident) {
  v = warp_reduce<WrapOp<OpT, C>, C>(v, 32);
  if constexpr (G > 32) {
    const int tid = threadIdx.x;
    __syncthreads();  // s_buf reuse
    if ((tid & 31) == 0) s_buf[tid >> 5] = v;
    __syncthreads();
    const int wbase = (tid / G) * (G / 32);  // first warp slot of this thread's group (G = 128: two groups per CTA)
    C r = (tid & 31) < G / 32 ? s_buf[wbase + (tid & 31)] : ident;
    v = warp_reduce<WrapOp<OpT, C>, C>(r, 32);  // full butterfly: every lane ends with the row total
  }
  return v;
}

// NCH = packs per thread the row needs (a template parameter, so that a 4-pack row does not carry the registers
// and predicated instructions of an 8-pack one: 4096-element f32 rows went from 48 to ≤ 40 registers, 5 → 6 CTAs/SM)
template <typename T, int VEC, int G, bool LOG, int NCH>
__global__ void __launch_bounds__(kSmThreads)
softmax_rows_reg(const T* __restrict__ in, typename type_of_dtype<promote_ct(dtype_of<T>::value, 0, 2)>::type* __restrict__ out,
                 SoftmaxParams p) {
  pdl_prologue();
  typedef typename type_of_dtype<promote_ct(dtype_of<T>::value, 0, 2)>::type O;
  typedef compute_t<O> C;
  __shared__ C s_buf[kSmThreads / 32];
  const int tid = threadIdx.x;
  const int lane = tid & (G - 1);
  const int64_t row = (int64_t)blockIdx.x * (kSmThreads / G) + tid / G;
  const bool active = row < p.M;  // whole groups are active or not (G divides the CTA)
  int64_t in_off = 0, out_off = 0;
  if (active) walk2(row, p.kept, p.use64, in_off, out_off);
  const C neg_inf = Limits<C>::lowest();
  C x[NCH][VEC];
  C mx = neg_inf;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int e = (i * G + lane) * VEC;  // a register-resident row has at most 8·256·VEC elements
    if (i < p.nchunks && active && e < (int)p.L) {
      Pack<T, VEC> v;
      load_pack<T, VEC>(v, in + in_off + e);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        x[i][k] = to_compute<O>(cast<O>(v.v[k]));
        mx = sm_max<C>(mx, x[i][k]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < VEC; ++k) x[i][k] = neg_inf;
    }
  }
  mx = group_reduce<MaxOp, C, G>(mx, s_buf, neg_inf);
  C sum = (C)0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    if (i < p.nchunks) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        // plain f32 exp(x − max), as the reference's CPU kernel computes it (cpu/kernels/softmax.rs:204-310): the
        // rounding of x − max costs up to |x − max|/2 ulp of the result, which the parity bound accounts for.
        // (The two-pass kernels below fold that error back in — they have instruction slack; this one does not.)
        const C sh = x[i][k] - mx;
        const C ex = sm_exp_o<O>(sh);
        sum += ex;  // padding lanes: exp(-inf - mx) = 0 (or NaN only if mx itself is -inf/NaN, handled by row)
        x[i][k] = LOG ? sh : ex;
      }
    }
  }
  sum = group_reduce<AddOp, C, G>(sum, s_buf, (C)0);
  const C lg = sm_log<C>(sum);
  const C inv = (C)1 / sum;  // one division per row; the per-element multiply adds ≤ 0.5 ulp over a division
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int e = (i * G + lane) * VEC;
    if (i < p.nchunks && active && e < (int)p.L) {
      Pack<O, VEC> o;
#pragma unroll
      for (int k = 0; k < VEC; ++k) o.v[k] = from_compute<O>(LOG ? x[i][k] - lg : x[i][k] * inv);
      store_pack<O, VEC>(out + out_off + e, o);
    }
  }
}

// The same for rows of more than 4 packs per thread, PERSISTENT: one wave of resident CTAs, each walking row groups
// blockIdx.x, + gridDim.x, … (f32 [4096,8192]: 48.8 → 44.1 µs).  Shorter rows keep the one-group-per-CTA kernel above:
// there a fresh CTA from the block scheduler overlaps its neighbours better than a loop does (cfg 4 [4096,4096]:
// 22.9 µs against 24.7 µs persistent, 25.7 µs persistent with the next group's packs prefetched — the extra registers
// cost two CTAs per SM).
template <typename T, int VEC, int G, bool LOG, int NCH>
__global__ void __launch_bounds__(kSmThreads)
softmax_rows_reg_loop(const T* __restrict__ in, typename type_of_dtype<promote_ct(dtype_of<T>::value, 0, 2)>::type* __restrict__ out,
                 SoftmaxParams p) {
  pdl_prologue();
  typedef typename type_of_dtype<promote_ct(dtype_of<T>::value, 0, 2)>::type O;
  typedef compute_t<O> C;
  constexpr int kRows = kSmThreads / G;
  __shared__ C s_buf[kSmThreads / 32];
  const int tid = threadIdx.x;
  const int lane = tid & (G - 1);
  const C neg_inf = Limits<C>::lowest();
  const int64_t nblk = (p.M + kRows - 1) / kRows;
  // raw packs of a row group; lanes past the row's end (and inactive groups) hold nothing and are refilled below
  auto fetch = [&](int64_t blk, Pack<T, VEC> (&v)[NCH], int64_t& out_off, bool& active) {
    const int64_t row = blk * kRows + tid / G;
    active = row < p.M;  // whole groups are active or not (G divides the CTA)
    int64_t in_off = 0;
    out_off = 0;
    if (active) walk2(row, p.kept, p.use64, in_off, out_off);
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int e = (i * G + lane) * VEC;  // a register-resident row has at most 8·256·VEC elements
      if (i < p.nchunks && active && e < (int)p.L) load_pack<T, VEC>(v[i], in + in_off + e);
    }
  };
  Pack<T, VEC> cur[NCH];
  int64_t out_off = 0;
  bool active = false;
  if ((int64_t)blockIdx.x < nblk) fetch(blockIdx.x, cur, out_off, active);
  for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    C x[NCH][VEC];
    C mx = neg_inf;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int e = (i * G + lane) * VEC;
      if (i < p.nchunks && active && e < (int)p.L) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          x[i][k] = to_compute<O>(cast<O>(cur[i].v[k]));
          mx = sm_max<C>(mx, x[i][k]);
        }
      } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) x[i][k] = neg_inf;
      }
    }
    mx = group_reduce<MaxOp, C, G>(mx, s_buf, neg_inf);
    C sum = (C)0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      if (i < p.nchunks) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          // plain f32 exp(x − max), as the reference's CPU kernel computes it (cpu/kernels/softmax.rs:204-310): the
          // rounding of x − max costs up to |x − max|/2 ulp of the result, which the parity bound accounts for.
          const C sh = x[i][k] - mx;
          const C ex = sm_exp_o<O>(sh);
          sum += ex;  // padding lanes: exp(-inf - mx) = 0 (or NaN only if mx itself is -inf/NaN, handled by row)
          x[i][k] = LOG ? sh : ex;
        }
      }
    }
    sum = group_reduce<AddOp, C, G>(sum, s_buf, (C)0);
    const C lg = sm_log<C>(sum);
    const C inv = (C)1 / sum;  // one division per row; the per-element multiply adds ≤ 0.5 ulp over a division
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int e = (i * G + lane) * VEC;
      if (i < p.nchunks && active && e < (int)p.L) {
        Pack<O, VEC> o;
#pragma unroll
        for (int k = 0; k < VEC; ++k) o.v[k] = from_compute<O>(LOG ? x[i][k] - lg : x[i][k] * inv);
        store_pack<O, VEC>(out + out_off + e, o);
      }
    }
    if (blk + gridDim.x < nblk) fetch(blk + gridDim.x, cur, out_off, active);
  }
}

// online (max, Σ exp(x − max)) pair
template <typename C> struct MS {
  C m, s;
};
template <typename C>
__device__ __forceinline__ MS<C> ms_combine(MS<C> a, MS<C> b) {
  if (a.s == (C)0) return b;  // identity (also avoids (-inf) - (-inf))
  if (b.s == (C)0) return a;
  const C m = sm_max<C>(a.m, b.m);
  return MS<C>{m, a.s * sm_shift_exp<C>(a.m, m).ex + b.s * sm_shift_exp<C>(b.m, m).ex};
}

// fold one element into a running (max, Σ) pair: one exp per element
template <typename C>
__device__ __forceinline__ void ms_push(MS<C>& a, C x) {
  if (a.s == (C)0) { a.m = x; a.s = (C)1; return; }
  if (x <= a.m) a.s += sm_shift_exp<C>(x, a.m).ex;
  else if (x > a.m) { a.s = a.s * sm_shift_exp<C>(a.m, x).ex + (C)1; a.m = x; }
  else a.s += x;  // NaN element: poison the row as exp(NaN) would
}

// ---- streaming statistics with a power-of-two frame (f32 compute type) --------------------------------------------------
// The vectorised streaming kernels keep, instead of (max, Σ exp(x − max)), a pair (K, T): K an INTEGER no smaller than
// any x·log2e seen so far, T = Σ exp(x) · 2^-K.  Moving from K to K' multiplies T by 2^(K − K') — an exponent shift,
// exact — so however often the running maximum moves, and however many partial pairs are merged (threads, warps, CTAs,
// axis splits), T carries only the rounding of its additions.
// Terms are formed in the x domain, against the float r_K = fl(K·ln2): exp(x − r_K) is the register kernel's term — one
// rounded subtraction, compensated ex2, 8 instructions — and differs from exp(x)·2^-K by the constant
// c_K = exp(r_K − K·ln2), 1 to within the rounding of r_K, which is applied ONCE per stretch of equal K (a thread's raw sum S joins T as fma(S, c_K, T)), not per element.  Round 2 first
// shipped terms formed in the log2 domain (2^f · 2^(n − K) with n = rint(x·log2e)): exact whatever the distance to the
// maximum, but 15–24 instructions per element including three conversion-pipe slots, which left f32 AND bf16
// [4096,8192] axis 0 at 97 µs (XU-bound), 86 µs with the conversions replaced by magic-number adds (issue-bound).
constexpr float kLog2eHi = 1.4426950216293335f;
constexpr float kLn2Hi = 0.693145751953125f, kLn2Lo = 1.42860682030941723212e-6f;  // ln2 = hi + lo, hi with 9 trailing zero bits
__device__ __forceinline__ float sm_exp2i(float d) {  // 2^d for an integer-valued d ≤ 0 (−inf → 0)
  return d < -126.0f ? 0.0f : __int_as_float((127 + (int)d) << 23);
}
// r_K = fl(K·ln2), ln c_K = r_K − K·ln2 (a rounding residue: ≤ ulp(r_K)/2), and c_K itself — a cubic while |ln c_K| < 2^-6
// (remainder < 2^-28).  __fmul_rn: the product must not be contracted into a neighbouring add, every user of r_K has
// to see the same float.
__device__ __forceinline__ float sm_ref(float K) { return fmaf(K, kLn2Lo, __fmul_rn(K, kLn2Hi)); }
__device__ __forceinline__ float sm_frame_ln(float K) {
  const float p = __fmul_rn(K, kLn2Hi);
  const float r = fmaf(K, kLn2Lo, p);
  const float e1 = fmaf(K, kLn2Hi, -p);  // K·ln2_hi − p, exact
  return fmaf(-K, kLn2Lo, r - p) - e1;   // r − p is exact
}
__device__ __forceinline__ float sm_frame(float K) {
  const float z = sm_frame_ln(K);
  if (fabsf(z) < 0.015625f) return fmaf(z, fmaf(z, fmaf(z, 1.0f / 6.0f, 0.5f), 1.0f), 1.0f);
  return fast_expf(z);
}
// a thread's running statistics of one softmax lane
template <typename C> struct SmRun;
// While a thread sweeps, its frame may LAG the running maximum by up to kSmLag binary orders: terms up to 2^kSmLag are
// as accurate as terms below 1 (what matters is |x − r_K|, and a reference below the maximum is closer to everything
// under it), Σ has the exponent range to spare, and the frame then moves a few times per sweep instead of whenever some
// lane of the warp sees a new integer part of max·log2e — which, with 32 independent maxima per warp, was in most
// blocks: 22 instructions per element instead of 13 (ncu, f32 [4096,8192]: 23.3 M warp instructions, 39 µs).  done()
// moves the pair to the frame of the TRUE maximum (an exponent shift), so merges and the apply sweep never see the lag.
constexpr float kSmLag = 8.0f;
template <> struct SmRun<float> {
  float K, r, T, S, comp, ymax;  // frame, its reference r_K, the sum in frame units, the raw sum of the current stretch (+ its lost bits), max of x·log2e
  __device__ __forceinline__ void init() { K = Limits<float>::lowest(); r = 0.0f; T = 0.0f; S = 0.0f; comp = 0.0f; ymax = K; }
  __device__ __forceinline__ void close() {
    if (S != 0.0f) T = fmaf(S, sm_frame(K), T);  // also when S is NaN
    S = 0.0f;
    comp = 0.0f;
  }
  // N values at once: ONE frame update for the block (its maximum decides), then N plain terms
  template <bool HALF, int N>
  __device__ __forceinline__ void push(const float (&x)[N]) {
    float bm = x[0];
#pragma unroll
    for (int i = 1; i < N; ++i) bm = fmaxf(bm, x[i]);  // NaNs drop out here and poison S through their term
    const float y = bm * kLog2eHi;
    ymax = fmaxf(ymax, y);
    if (y > K + kSmLag) {  // false for NaN and for −inf against the initial −inf
      const float Kn = ceilf(y);  // +inf input: K = r = +inf, every term NaN or 0 as exp(x − inf) is
      close();
      T *= sm_exp2i(K - Kn);
      K = Kn;
      r = sm_ref(Kn);
    }
    // A thread adds a hundred or more terms to S, most of them far below the largest: added one by one, terms under
    // half an ulp of S vanish — a one-sided loss, measured at 3–6 · 2^-24 of the lane's Σ on N(0, 4²) columns of 4096.
    // The block is summed on its own first (pairwise), and its sum joins S with the rounding error carried along.
    float t[N];
#pragma unroll
    for (int i = 0; i < N; ++i) t[i] = HALF ? sm_exp_half(x[i] - r) : fast_expf(x[i] - r);  // −inf → 0 (r = 0 until the first finite value)
#pragma unroll
    for (int w = 1; w < N; w <<= 1)
#pragma unroll
      for (int i = 0; i + w < N; i += 2 * w) t[i] += t[i + w];
    const float v = t[0] - comp, s2 = S + v;
    comp = (s2 - S) - v;
    S = s2;
  }
  __device__ __forceinline__ MS<float> done() {
    close();
    const float Kt = ceilf(ymax);  // the frame of the true maximum
    if (Kt > K) {
      T *= sm_exp2i(K - Kt);
      K = Kt;
    }
    return MS<float>{K, T};
  }
};
// 64-bit compute type: the plain online pair
__device__ __forceinline__ void ms_push_fast(MS<double>& a, double x) {
  if (a.s == 0.0) { a.m = x; a.s = 1.0; return; }
  if (x <= a.m) a.s += exp(x - a.m);
  else if (x > a.m) { a.s = a.s * exp(a.m - x) + 1.0; a.m = x; }
  else a.s += x;
}
template <> struct SmRun<double> {
  MS<double> a;
  __device__ __forceinline__ void init() { a = MS<double>{Limits<double>::lowest(), 0.0}; }
  template <bool HALF, int N>
  __device__ __forceinline__ void push(const double (&x)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i)
      if (x[i] > Limits<double>::lowest() || x[i] != x[i]) ms_push_fast(a, x[i]);  // −inf (a real one or tail padding) adds nothing
  }
  __device__ __forceinline__ MS<double> done() { return a; }
};
// merge two (K, T) pairs: exact scalings, one rounded addition
__device__ __forceinline__ MS<float> ms_merge_fast(MS<float> a, MS<float> b) {
  if (a.s == 0.0f) return b;
  if (b.s == 0.0f) return a;
  const float K = fmaxf(a.m, b.m);
  return MS<float>{K, a.s * sm_exp2i(a.m - K) + b.s * sm_exp2i(b.m - K)};
}
__device__ __forceinline__ MS<double> ms_merge_fast(MS<double> a, MS<double> b) { return ms_combine<double>(a, b); }
// a RUN of merges into one accumulator (a row's warps, a column's axis splits): typically one partial holds the maximum
// and the others arrive one at a time at half an ulp of it or less, where plain additions round the same way every
// time (f32 [6144,96] axis 0 in 24 splits: Σ low by 5 · 2^-24).  TwoSum keeps what each addition drops; ms_run_done
// folds it back.
__device__ __forceinline__ void ms_run_add(MS<float>& a, float& comp, MS<float> b) {
  if (b.s == 0.0f) return;
  if (a.s == 0.0f) { a = b; comp = 0.0f; return; }
  const float K = fmaxf(a.m, b.m);
  const float fa = sm_exp2i(a.m - K), sa = a.s * fa, sb = b.s * sm_exp2i(b.m - K);
  const float t = sa + sb, bb = t - sa;
  comp = comp * fa + ((sa - (t - bb)) + (sb - bb));
  a = MS<float>{K, t};
}
__device__ __forceinline__ void ms_run_done(MS<float>& a, float comp) { a.s += comp; }
__device__ __forceinline__ void ms_run_add(MS<double>& a, double&, MS<double> b) { a = ms_combine<double>(a, b); }
__device__ __forceinline__ void ms_run_done(MS<double>&, double) {}
// the lane's normalised output from its (K, T): softmax = exp(x − r_K) · c_K / T, log_softmax = (x − r_K) − (ln T − ln c_K)
struct SmFinal {
  float r, q, lgq;
};
__device__ __forceinline__ SmFinal sm_final(MS<float> a) {
  return SmFinal{sm_ref(a.m), sm_frame(a.m) / a.s, logf(a.s) - sm_frame_ln(a.m)};
}
template <bool HALF>
__device__ __forceinline__ float sm_out(float x, const SmFinal& f, int log) {
  const float sh = x - f.r;
  return log ? sh - f.lgq : (HALF ? sm_exp_half(sh) : fast_expf(sh)) * f.q;
}
struct SmFinalD {
  double m, inv, lg;
};
__device__ __forceinline__ SmFinalD sm_final(MS<double> r) { return SmFinalD{r.m, 1.0 / r.s, log(r.s)}; }
template <bool HALF>
__device__ __forceinline__ double sm_out(double x, const SmFinalD& f, int log) {
  const double sh = x - f.m;
  return log ? sh - f.lg : exp(sh) * f.inv;
}

template <typename T>
__global__ void __launch_bounds__(kSmThreads)
softmax_rows_stream(const T* __restrict__ in, typename type_of_dtype<promote_ct(dtype_of<T>::value, 0, 2)>::type* __restrict__ out,
                    SoftmaxParams p) {
  pdl_prologue();
  typedef typename type_of_dtype<promote_ct(dtype_of<T>::value, 0, 2)>::type O;
  typedef compute_t<O> C;
  __shared__ C s_m[kSmThreads / 32], s_s[kSmThreads / 32];
  const int tid = threadIdx.x;
  for (int64_t row = blockIdx.x; row < p.M; row += gridDim.x) {
    int64_t in_off = 0, out_off = 0;
    walk2(row, p.kept, p.use64, in_off, out_off);
    const T* src = in + in_off;
    O* dst = out + out_off;
    MS<C> a{Limits<C>::lowest(), (C)0};
    for (int64_t e = tid; e < p.L; e += kSmThreads) {
      const C x = to_compute<O>(cast<O>(load_one(src + e * p.sa_in)));
      ms_push<C>(a, x);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      MS<C> b{shfl_xor<C>(a.m, off), shfl_xor<C>(a.s, off)};
      a = (tid & off) == 0 ? ms_combine<C>(a, b) : ms_combine<C>(b, a);
    }
    __syncthreads();
    if ((tid & 31) == 0) { s_m[tid >> 5] = a.m; s_s[tid >> 5] = a.s; }
    __syncthreads();
    MS<C> r{Limits<C>::lowest(), (C)0};
    for (int w = 0; w < kSmThreads / 32; ++w) r = ms_combine<C>(r, MS<C>{s_m[w], s_s[w]});
    const C ssum = r.s, lg = sm_log<C>(r.s);
    for (int64_t e = tid; e < p.L; e += kSmThreads) {
      const C x = to_compute<O>(cast<O>(src[e * p.sa_in]));
      const ShExp<C> se = sm_shift_exp<C>(x, r.m);
      dst[e * p.sa_out] = from_compute<O>(p.log ? se.sh - lg : se.ex / ssum);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kSmThreads)
softmax_cols(const T* __restrict__ in, typename type_of_dtype<promote_ct(dtype_of<T>::value, 0, 2)>::type* __restrict__ out,
             SoftmaxParams p) {
  pdl_prologue();
  typedef typename type_of_dtype<promote_ct(dtype_of<T>::value, 0, 2)>::type O;
  typedef compute_t<O> C;
  const int64_t col = (int64_t)blockIdx.x * kSmThreads + threadIdx.x;
  if (col >= p.M) return;
  int64_t in_off = 0, out_off = 0;
  walk2(col, p.kept, p.use64, in_off, out_off);
  const T* src = in + in_off;
  O* dst = out + out_off;
  // blocks of 256 elements, combined pairwise-style: a plain running Σ over a long column (70000 rows) would carry
  // ~sqrt(L)/2 ulp of accumulated rounding, at the edge of the 1e-6·log2(n) sum bound
  MS<C> a{Limits<C>::lowest(), (C)0};
  for (int64_t e0 = 0; e0 < p.L; e0 += 256) {
    MS<C> blk{Limits<C>::lowest(), (C)0};
    const int64_t e1 = e0 + 256 < p.L ? e0 + 256 : p.L;
    for (int64_t e = e0; e < e1; ++e) {
      const C x = to_compute<O>(cast<O>(load_one(src + e * p.sa_in)));
      ms_push<C>(blk, x);
    }
    a = ms_combine<C>(a, blk);
  }
  const C ssum = a.s, lg = sm_log<C>(a.s);
  for (int64_t e = 0; e < p.L; ++e) {
    const C x = to_compute<O>(cast<O>(src[e * p.sa_in]));
    const ShExp<C> se = sm_shift_exp<C>(x, a.m);
    dst[e * p.sa_out] = from_compute<O>(p.log ? se.sh - lg : se.ex / ssum);
  }
}

// Long unit-stride rows (more than 8 packs per thread): one CTA per row, 16-byte loads with four in flight per
// thread, a statistics sweep and an apply sweep whose reads come back from L2 (a row is ≤ a few MB).  The scalar
// softmax_rows_stream above stays for strided axes: f32 [256,131072] softmax(1) 551 µs through it.
// NT = 1024 threads for rows of ≥ 8192 packs: a grid of a few hundred rows otherwise leaves each SM with one or two
// 256-thread CTAs (f32 [256,131072]: 118 µs with 256 threads per row).
// PHASE 0: the whole job in one launch.  FEWER rows than CTA slots (a 1-D softmax: one row) would leave most of the GPU
// idle — f32 [64,524288]: 64 CTAs on 148 SMs, 94 µs — so each row is split into gridDim.x slabs and the job becomes two
// launches, as for the column kernel below: PHASE 1 leaves one (K, T) pair per (row, slab) in `part`, PHASE 2 merges a
// row's pairs in slab order and applies to its own slab.
template <typename T, int VEC, int NT, int PHASE>
__global__ void __launch_bounds__(NT)
softmax_rows_stream_vec(const T* __restrict__ in, typename type_of_dtype<promote_ct(dtype_of<T>::value, 0, 2)>::type* __restrict__ out,
                        compute_t<typename type_of_dtype<promote_ct(dtype_of<T>::value, 0, 2)>::type>* __restrict__ part, SoftmaxParams p) {
  pdl_prologue();
  typedef typename type_of_dtype<promote_ct(dtype_of<T>::value, 0, 2)>::type O;
  typedef compute_t<O> C;
  constexpr int UN = 4;
  __shared__ C s_m[NT / 32], s_s[NT / 32];
  const int tid = threadIdx.x;
  const int64_t packs = p.L / VEC;  // the host picks this kernel only when L is a multiple of VEC
  const int64_t row0 = PHASE == 0 ? blockIdx.x : blockIdx.y, row_step = PHASE == 0 ? gridDim.x : p.M;
  const int64_t c_begin = PHASE == 0 ? 0 : (int64_t)blockIdx.x * p.rps;
  const int64_t c_end = PHASE == 0 ? packs : (c_begin + p.rps < packs ? c_begin + p.rps : packs);
  for (int64_t row = row0; row < p.M; row += row_step) {
    int64_t in_off = 0, out_off = 0;
    walk2(row, p.kept, p.use64, in_off, out_off);
    const T* src = in + in_off;
    O* dst = out + out_off;
    MS<C> r{Limits<C>::lowest(), (C)0};
    C rcomp = (C)0;
    if constexpr (PHASE != 2) {
      SmRun<C> run;
      run.init();
      int64_t c = c_begin + tid;
      for (; c + (int64_t)(UN - 1) * NT < c_end; c += (int64_t)NT * UN) {  // whole batches: no predicates
        Pack<T, VEC> v[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) load_pack<T, VEC>(v[u], src + (c + (int64_t)u * NT) * VEC);
        C xs[UN * VEC];
#pragma unroll
        for (int u = 0; u < UN; ++u)
#pragma unroll
          for (int k = 0; k < VEC; ++k) xs[u * VEC + k] = to_compute<O>(cast<O>(v[u].v[k]));
        run.template push<sizeof(O) == 2>(xs);
      }
      for (; c < c_end; c += NT) {  // ragged tail, a pack at a time
        Pack<T, VEC> v;
        load_pack<T, VEC>(v, src + c * VEC);
        C xs[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) xs[k] = to_compute<O>(cast<O>(v.v[k]));
        run.template push<sizeof(O) == 2>(xs);
      }
      MS<C> a = run.done();
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        MS<C> b{shfl_xor<C>(a.m, off), shfl_xor<C>(a.s, off)};
        a = (tid & off) == 0 ? ms_merge_fast(a, b) : ms_merge_fast(b, a);
      }
      __syncthreads();
      if ((tid & 31) == 0) { s_m[tid >> 5] = a.m; s_s[tid >> 5] = a.s; }
      __syncthreads();
      for (int w = 0; w < NT / 32; ++w) ms_run_add(r, rcomp, MS<C>{s_m[w], s_s[w]});
      ms_run_done(r, rcomp);
      if constexpr (PHASE == 1) {
        if (tid == 0) {
          C* pm = part + (row * gridDim.x + blockIdx.x) * 2;
          pm[0] = r.m;
          pm[1] = r.s;
        }
        continue;
      }
    } else {
      // the lanes of a warp take the slabs round-robin, a shuffle tree merges the 32 pairs: every warp of every slab's
      // CTA computes the same bits
      const C* pm = part + row * gridDim.x * 2;
      for (int sp = tid & 31; sp < (int)gridDim.x; sp += 32) ms_run_add(r, rcomp, MS<C>{pm[2 * sp], pm[2 * sp + 1]});
      ms_run_done(r, rcomp);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        MS<C> b{shfl_xor<C>(r.m, off), shfl_xor<C>(r.s, off)};
        r = (tid & off) == 0 ? ms_merge_fast(r, b) : ms_merge_fast(b, r);
      }
    }
    const auto fin = sm_final(r);
    int64_t c = c_begin + tid;
    for (; c + (int64_t)(UN - 1) * NT < c_end; c += (int64_t)NT * UN) {
      Pack<T, VEC> v[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) load_pack_cached<T, VEC>(v[u], src + (c + (int64_t)u * NT) * VEC);
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        Pack<O, VEC> o;
#pragma unroll
        for (int k = 0; k < VEC; ++k) o.v[k] = from_compute<O>(sm_out<sizeof(O) == 2>(to_compute<O>(cast<O>(v[u].v[k])), fin, p.log));
        store_pack<O, VEC>(dst + (c + (int64_t)u * NT) * VEC, o);
      }
    }
    for (; c < c_end; c += NT) {
      Pack<T, VEC> v;
      load_pack_cached<T, VEC>(v, src + c * VEC);
      Pack<O, VEC> o;
#pragma unroll
      for (int k = 0; k < VEC; ++k) o.v[k] = from_compute<O>(sm_out<sizeof(O) == 2>(to_compute<O>(cast<O>(v.v[k])), fin, p.log));
      store_pack<O, VEC>(dst + c * VEC, o);
    }
  }
}

// Strided axis, contiguous kept dim (softmax over axis 0 of a row-major matrix): a CTA owns TX·VEC adjacent columns,
// lanes run along them with 16-byte loads (a warp reads 32/TX full rows of TX·VEC·sizeof(T) bytes per instruction),
// the TY = 256/TX thread rows stride down the axis keeping one online (max, Σ) pair per column, a shared-memory
// tree merges the thread rows, then a second sweep writes.  2 reads + 1 write; the column-per-thread kernel this
// replaces for wide tensors ran 4096 serial steps on 32 CTAs (f32 [4096,8192] axis 0: 4050 µs).
// PHASE 0: the whole job in one launch.  When the column tiles alone cannot fill the GPU the axis is split over
// gridDim.y CTAs and the job becomes two launches: PHASE 1 leaves one (max, Σ) pair per (split, column) in `part`,
// PHASE 2 merges a column's pairs (in split order: deterministic) and writes its own slab of rows.
template <typename T, int VEC, int TX, int PHASE>
__global__ void __launch_bounds__(kSmThreads)
softmax_cols_tiled(const T* __restrict__ in, typename type_of_dtype<promote_ct(dtype_of<T>::value, 0, 2)>::type* __restrict__ out,
                   compute_t<typename type_of_dtype<promote_ct(dtype_of<T>::value, 0, 2)>::type>* __restrict__ part, SoftmaxParams p) {
  pdl_prologue();
  typedef typename type_of_dtype<promote_ct(dtype_of<T>::value, 0, 2)>::type O;
  typedef compute_t<O> C;
  constexpr int TY = kSmThreads / TX, W = TX * VEC, UN = 4;
  __shared__ C s_m[PHASE == 2 ? 1 : TY][W], s_s[PHASE == 2 ? 1 : TY][W];
  const int lane = threadIdx.x % TX, ty = threadIdx.x / TX;
  // PHASE 2 walks the grid backwards: the slabs PHASE 1 read last are still in L2
  const unsigned bx = PHASE == 2 ? gridDim.x - 1 - blockIdx.x : blockIdx.x, by = PHASE == 2 ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
  const int64_t e_begin = (int64_t)by * p.rps;
  const int64_t e_end = e_begin + p.rps < p.L ? e_begin + p.rps : p.L;
  const int64_t ncols = (int64_t)gridDim.x * W;  // row length of the `part` arrays (padded to whole tiles)
  const int64_t outer = (int64_t)bx / p.ctiles, tile = (int64_t)bx - outer * p.ctiles;
  const int64_t col0 = tile * W + (int64_t)lane * VEC;
  int64_t in_off = 0, out_off = 0;
  if (p.outer.n > 0) walk2(outer, p.outer, p.use64, in_off, out_off);
  const bool active = col0 < p.C;  // C is a multiple of VEC when VEC > 1
  const T* src = in + in_off + col0;
  O* dst = out + out_off + col0;
  MS<C> a[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) a[k] = MS<C>{Limits<C>::lowest(), (C)0};
  if constexpr (PHASE != 2) {
    SmRun<C> run[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) run[k].init();
    if (active) {
      // rows per batch of the statistics sweep = 16-byte loads in flight per thread.  f32: 8 rows need 76 registers (three
      // CTAs per SM) and measured 85 µs on [4096,8192] against 80 µs for 4 rows at 63 registers (82 vs 76 µs with the lagging
      // frame); bf16 (8 lanes per pack)
      // is at two CTAs per SM either way and prefers 8 rows (64 vs 79 µs)
      constexpr int US = VEC > 4 ? 8 : 4;
      int64_t e = e_begin + ty;
      for (; e + (int64_t)(US - 1) * TY < e_end; e += (int64_t)TY * US) {  // whole batches: no predicates
        Pack<T, VEC> v[US];
#pragma unroll
        for (int u = 0; u < US; ++u) load_pack<T, VEC>(v[u], src + (e + (int64_t)u * TY) * p.sa_in);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          C xs[US];
#pragma unroll
          for (int u = 0; u < US; ++u) xs[u] = to_compute<O>(cast<O>(v[u].v[k]));
          run[k].template push<sizeof(O) == 2>(xs);
        }
      }
      for (; e < e_end; e += TY) {  // ragged tail, a row at a time
        Pack<T, VEC> v;
        load_pack<T, VEC>(v, src + e * p.sa_in);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          C xs[1] = {to_compute<O>(cast<O>(v.v[k]))};
          run[k].template push<sizeof(O) == 2>(xs);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) a[k] = run[k].done();
#pragma unroll
    for (int k = 0; k < VEC; ++k) { s_m[ty][lane * VEC + k] = a[k].m; s_s[ty][lane * VEC + k] = a[k].s; }
    __syncthreads();
#pragma unroll 1
    for (int off = TY / 2; off > 0; off >>= 1) {
      if (ty < off) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          a[k] = ms_merge_fast(a[k], MS<C>{s_m[ty + off][lane * VEC + k], s_s[ty + off][lane * VEC + k]});
          s_m[ty][lane * VEC + k] = a[k].m;
          s_s[ty][lane * VEC + k] = a[k].s;
        }
      }
      __syncthreads();
    }
    if constexpr (PHASE == 1) {  // one pair per (split, column)
      if (ty == 0) {
        C* pm = part + ((int64_t)by * 2) * ncols + (int64_t)bx * W + lane * VEC;
#pragma unroll
        for (int k = 0; k < VEC; ++k) { pm[k] = a[k].m; pm[ncols + k] = a[k].s; }
      }
      return;
    }
  } else {
    // merge the column's per-split pairs in split order; every thread row needs them, thread row 0 fetches
    if (ty == 0) {
      C acomp[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) { a[k] = MS<C>{Limits<C>::lowest(), (C)0}; acomp[k] = (C)0; }
      for (int sp = 0; sp < (int)gridDim.y; ++sp) {
        const C* pm = part + ((int64_t)sp * 2) * ncols + (int64_t)bx * W + lane * VEC;
#pragma unroll
        for (int k = 0; k < VEC; ++k) ms_run_add(a[k], acomp[k], MS<C>{pm[k], pm[ncols + k]});
      }
#pragma unroll
      for (int k = 0; k < VEC; ++k) ms_run_done(a[k], acomp[k]);
#pragma unroll
      for (int k = 0; k < VEC; ++k) { s_m[0][lane * VEC + k] = a[k].m; s_s[0][lane * VEC + k] = a[k].s; }
    }
    __syncthreads();
  }
  if (!active) return;
  decltype(sm_final(MS<C>{})) fin[VEC];  // one division per column; the per-element multiply adds ≤ 0.5 ulp
#pragma unroll
  for (int k = 0; k < VEC; ++k) fin[k] = sm_final(MS<C>{s_m[0][lane * VEC + k], s_s[0][lane * VEC + k]});
  int64_t e = e_begin + ty;
  for (; e + (int64_t)(UN - 1) * TY < e_end; e += (int64_t)TY * UN) {
    Pack<T, VEC> v[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) load_pack<T, VEC>(v[u], src + (e + (int64_t)u * TY) * p.sa_in);
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      Pack<O, VEC> o;
#pragma unroll
      for (int k = 0; k < VEC; ++k) o.v[k] = from_compute<O>(sm_out<sizeof(O) == 2>(to_compute<O>(cast<O>(v[u].v[k])), fin[k], p.log));
      store_pack<O, VEC>(dst + (e + (int64_t)u * TY) * p.sa_out, o);
    }
  }
  for (; e < e_end; e += TY) {
    Pack<T, VEC> v;
    load_pack<T, VEC>(v, src + e * p.sa_in);
    Pack<O, VEC> o;
#pragma unroll
    for (int k = 0; k < VEC; ++k) o.v[k] = from_compute<O>(sm_out<sizeof(O) == 2>(to_compute<O>(cast<O>(v.v[k])), fin[k], p.log));
    store_pack<O, VEC>(dst + e * p.sa_out, o);
  }
}

#include "softmax_band.cuh"

// CTAs per cluster for a band of `band_bytes`: the smallest power of two that brings a CTA's share under 64 KB (three
// CTAs per SM), grown while the grid would leave SMs idle and every CTA keeps ≥ min_bytes; 0 = the band does not fit
// (> 8 × 96 KB).
inline int band_cluster(int64_t band_bytes, int64_t nbands, int sms, int64_t min_bytes) {
  int cl = 1;
  static const bool tune = [] { const char* e = getenv("HPTB_TUNE"); return e && e[0] == '1'; }();
  const char* tb = tune ? getenv("HPTB_TUNE_BAND_KB") : nullptr;  // development: target KB per CTA (tools/band_sweep.py)
  const int64_t target = tb ? atoll(tb) * 1024 : 65536;
  while (cl < kBandMaxCl && band_bytes > (int64_t)cl * target) cl <<= 1;
  if (band_bytes > (int64_t)cl * 98304) return 0;
  while (cl < kBandMaxCl && nbands * cl < (int64_t)sms * 3 && band_bytes / (cl * 2) >= min_bytes) cl <<= 1;
  return cl;
}
inline bool band_tune_on() {
  static const bool on = [] { const char* e = getenv("HPTB_TUNE"); return e && e[0] == '1'; }();
  return on;
}
inline bool band_disabled() {
  if (!band_tune_on()) return false;
  const char* e = getenv("HPTB_TUNE_NO_BAND");
  return e && e[0] == '1';
}

inline bool sm_g128_off() {
  if (!band_tune_on()) return false;
  const char* e = getenv("HPTB_TUNE_SM_NO_G128");
  return e && e[0] == '1';
}

template <typename T>
hptb_status launch_softmax(hptb_ctx* ctx, const Collapsed& c, const void* in_v, void* out_v, int log, cudaStream_t stream) {
  typedef typename type_of_dtype<promote_ct(dtype_of<T>::value, 0, 2)>::type O;
  const T* in = static_cast<const T*>(in_v);
  O* out = static_cast<O*>(out_v);
  SoftmaxParams p;
  memset(&p, 0, sizeof(p));
  int ad = -1, kept[kRedMaxDims], nk = 0;
  for (int d = c.ndim - 1; d >= 0; --d) {
    if (c.reduced[d]) ad = d; else kept[nk++] = d;
  }
  bool big = false;
  p.log = log;
  if (ad < 0) { p.L = 1; p.sa_in = 1; p.sa_out = 1; }  // axis of extent 1: softmax = 1, log_softmax = 0
  else { p.L = c.shape[ad]; p.sa_in = c.strides[1][ad]; p.sa_out = c.strides[0][ad]; }
  // kept dims: innermost first = smallest |out stride| first (they come sorted by out stride desc)
  fill_walk(p.kept, c, kept, nk, true, big);
  int64_t M = 1;
  for (int i = 0; i < nk; ++i) M *= c.shape[kept[i]];
  p.M = M;
  if (M == 0 || p.L == 0) return HPTB_OK;
  if (!red_fits_u32(M)) big = true;
  p.use64 = big ? 1 : 0;
  constexpr int kMinSz = sizeof(T) < sizeof(O) ? sizeof(T) : sizeof(O);
  constexpr int VECMAX = 16 / kMinSz > 8 ? 8 : 16 / kMinSz;
  if (p.sa_in == 1 && p.sa_out == 1) {
    // vector path needs every row start 16 B (pack) aligned in both tensors
    int vec = VECMAX;
    auto aligned = [&](int v) {
      if (v == 1) return true;
      size_t ai = sizeof(T) * v > 16 ? 16 : sizeof(T) * v, ao = sizeof(O) * v > 16 ? 16 : sizeof(O) * v;
      if (reinterpret_cast<uintptr_t>(in) % ai || reinterpret_cast<uintptr_t>(out) % ao) return false;
      if (p.L % v) return false;
      for (int i = 0; i < nk; ++i) {
        if ((uint64_t)(std::llabs(c.strides[1][kept[i]]) * (int64_t)sizeof(T)) % ai) return false;
        if ((uint64_t)(std::llabs(c.strides[0][kept[i]]) * (int64_t)sizeof(O)) % ao) return false;
      }
      return true;
    };
    if (!aligned(vec)) vec = 1;
    const int64_t packs = (p.L + vec - 1) / vec;
    int G = 0;
    if (packs <= 32 * 4) G = 32;                       // short rows: a warp each, ≤ 4 packs per lane
    // two rows per CTA, 4 packs per thread: one row's block reductions hide under the other's loads (f32 [16384,1536]
    // 40.1 → 31.1 µs, [65536,768] 118 → 88 µs, [16384,2048] 43.8 → 41.1 µs); the 16-bit types, whose packs carry 8
    // elements of arithmetic, prefer one row per CTA (bf16 [4096,4096] 14.3 vs 15.5 µs)
    else if (vec > 1 && VECMAX <= 4 && packs <= 128 * 4 && !sm_g128_off()) G = 128;
    else if (packs <= (int64_t)kSmThreads * kSmChunks) G = kSmThreads;
    if (G) {
      p.G = G;
      p.nchunks = (int)((packs + G - 1) / G);
      const int64_t nblk = (M + (kSmThreads / G) - 1) / (kSmThreads / G);
      if (nblk > 0x7fffffffLL) return fail(HPTB_ERR_UNSUPPORTED, "softmax: grid too large");
#define HPTB_SM_LAUNCH3(KERN, PERSIST)                                                                            \
  do {                                                                                                            \
    static const int occ = [] {                                                                                   \
      int o = 0;                                                                                                  \
      return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, KERN, kSmThreads, 0) == cudaSuccess && o > 0 ? o : 4; \
    }();                                                                                                          \
    const int64_t wave = (int64_t)ctx->sm_count * occ;                                                            \
    const int64_t blocks = (PERSIST) && wave < nblk ? wave : nblk;                                                \
    HPTB_CUDA_CHECK(launch_kernel(KERN, dim3((unsigned)blocks), dim3(kSmThreads), 0, stream, in, out, p));         \
  } while (0)
#define HPTB_SM_LAUNCH2(V, GG, N)                                                                               \
  do {                                                                                                            \
    if constexpr ((N) > 4) {                                                                                      \
      if (log) HPTB_SM_LAUNCH3((softmax_rows_reg_loop<T, V, GG, true, N>), true);                                 \
      else HPTB_SM_LAUNCH3((softmax_rows_reg_loop<T, V, GG, false, N>), true);                                    \
    } else {                                                                                                      \
      if (log) HPTB_SM_LAUNCH3((softmax_rows_reg<T, V, GG, true, N>), false);                                     \
      else HPTB_SM_LAUNCH3((softmax_rows_reg<T, V, GG, false, N>), false);                                        \
    }                                                                                                             \
  } while (0)
#define HPTB_SM_LAUNCH(V, GG)                              \
  do {                                                     \
    if (GG == 32 || p.nchunks <= 2) HPTB_SM_LAUNCH2(V, GG, (GG == 32 ? 4 : 2)); \
    else if (p.nchunks <= 4) HPTB_SM_LAUNCH2(V, GG, 4);    \
    else HPTB_SM_LAUNCH2(V, GG, kSmChunks);                \
  } while (0)
      if (vec > 1) {
        if (G == 32) HPTB_SM_LAUNCH(VECMAX, 32);
        else if (G == 128) HPTB_SM_LAUNCH2(VECMAX, 128, 4);
        else HPTB_SM_LAUNCH(VECMAX, kSmThreads);
      } else {
        if (G == 32) HPTB_SM_LAUNCH(1, 32);
        else HPTB_SM_LAUNCH(1, kSmThreads);
      }
#undef HPTB_SM_LAUNCH
#undef HPTB_SM_LAUNCH2
#undef HPTB_SM_LAUNCH3
      HPTB_CUDA_CHECK(cudaGetLastError());
      count_launches(1);
      return HPTB_OK;
    }
  }
  if constexpr (VECMAX > 1) {
    // long unit-stride rows with pack-aligned starts: the vectorised streaming kernel
    if (p.sa_in == 1 && p.sa_out == 1 && p.L % VECMAX == 0) {
      size_t ai = sizeof(T) * VECMAX > 16 ? 16 : sizeof(T) * VECMAX, ao = sizeof(O) * VECMAX > 16 ? 16 : sizeof(O) * VECMAX;
      bool ok = !(reinterpret_cast<uintptr_t>(in) % ai) && !(reinterpret_cast<uintptr_t>(out) % ao);
      for (int i = 0; ok && i < nk; ++i) {
        if ((uint64_t)(std::llabs(c.strides[1][kept[i]]) * (int64_t)sizeof(T)) % ai) ok = false;
        if ((uint64_t)(std::llabs(c.strides[0][kept[i]]) * (int64_t)sizeof(O)) % ao) ok = false;
      }
      if constexpr (std::is_same<T, O>::value && std::is_same<compute_t<O>, float>::value) {
        // the row fits the shared memory of a cluster: one HBM read, exact maximum before the first exp (softmax_band.cuh)
        const int cl = ok && !band_disabled() ? band_cluster(p.L * (int64_t)sizeof(T), M, ctx->sm_count, 4096) : 0;
        if (cl > 0 && M * cl <= 0x7fffffffLL) {
          BandParams q;
          memset(&q, 0, sizeof(q));
          q.kept = p.kept;
          q.L = p.L;
          const int64_t packs = p.L / VECMAX;
          q.per_cta = (packs + cl - 1) / cl;
          q.cl = cl;
          q.log = log;
          q.use64 = p.use64;
          const size_t smem = (size_t)q.per_cta * 16;
          auto kern = softmax_band_rows<T, VECMAX>;
          static const cudaError_t attr = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
          if (attr != cudaSuccess) return fail(HPTB_ERR_CUDA, "cudaFuncSetAttribute(softmax_band_rows) failed: %s", cudaGetErrorString(attr));
          HPTB_CUDA_CHECK(launch_kernel_cluster(kern, dim3((unsigned)(M * cl)), dim3(kSmThreads), smem, stream, (unsigned)cl, in, out, q));
          count_launches(1);
          return HPTB_OK;
        }
      }
      if (ok) {
        typedef compute_t<O> CT;
        CT* none = nullptr;
        int64_t blocks = M < (int64_t)ctx->sm_count * 16 ? M : (int64_t)ctx->sm_count * 16;
        const int64_t packs = p.L / VECMAX;
        if (packs >= 8192 && M < (int64_t)ctx->sm_count * 8) {
          // fewer rows than 1024-thread CTA slots: split every row into slabs of ≥ 8 batches per thread
          const int64_t slots = (int64_t)ctx->sm_count * 2;
          int64_t S = slots / M;
          if (S > packs / (1024 * 4 * 8)) S = packs / (1024 * 4 * 8);
          if (S > 512) S = 512;
          if (S >= 2 && M <= 65535) {
            p.rps = (packs + S - 1) / S;
            S = (packs + p.rps - 1) / p.rps;
            Scratch part;
            HPTB_TRY(part.get(ctx, (size_t)M * S * 2 * sizeof(CT), stream));
            CT* pp = static_cast<CT*>(part.ptr);
            HPTB_CUDA_CHECK(launch_kernel(softmax_rows_stream_vec<T, VECMAX, 1024, 1>, dim3((unsigned)S, (unsigned)M), dim3(1024), 0, stream, in, out, pp, p));
            HPTB_CUDA_CHECK(launch_kernel(softmax_rows_stream_vec<T, VECMAX, 1024, 2>, dim3((unsigned)S, (unsigned)M), dim3(1024), 0, stream, in, out, pp, p));
            HPTB_CUDA_CHECK(cudaGetLastError());
            count_launches(2);
            return HPTB_OK;
          }
          HPTB_CUDA_CHECK(launch_kernel(softmax_rows_stream_vec<T, VECMAX, 1024, 0>, dim3((unsigned)blocks), dim3(1024), 0, stream, in, out, none, p));
        } else {
          HPTB_CUDA_CHECK(launch_kernel(softmax_rows_stream_vec<T, VECMAX, kSmThreads, 0>, dim3((unsigned)blocks), dim3(kSmThreads), 0, stream, in, out, none, p));
        }
        HPTB_CUDA_CHECK(cudaGetLastError());
        count_launches(1);
        return HPTB_OK;
      }
    }
  }
  const bool cols_ok = nk > 0 && c.strides[1][kept[0]] == 1 && c.strides[0][kept[0]] == 1 && !(p.sa_in == 1 && p.sa_out == 1) &&
                       M >= 32;
  if (cols_ok && c.shape[kept[0]] >= 8) {
    // tiled columns: 16-byte packs when every row start is pack-aligned in both tensors, else one element per lane
    bool big2 = false;
    fill_walk(p.outer, c, kept + 1, nk - 1, true, big2);
    p.C = c.shape[kept[0]];
    const int64_t outer_n = M / p.C;
    if (!red_fits_u32(outer_n)) big2 = true;
    p.use64 = (big || big2) ? 1 : 0;
    int vec = VECMAX;
    {
      size_t ai = sizeof(T) * vec > 16 ? 16 : sizeof(T) * vec, ao = sizeof(O) * vec > 16 ? 16 : sizeof(O) * vec;
      bool ok = !(reinterpret_cast<uintptr_t>(in) % ai) && !(reinterpret_cast<uintptr_t>(out) % ao) && p.C % vec == 0;
      for (int d = 0; ok && d < c.ndim; ++d) {
        if (d == kept[0]) continue;
        if ((uint64_t)(std::llabs(c.strides[1][d]) * (int64_t)sizeof(T)) % ai) ok = false;
        if ((uint64_t)(std::llabs(c.strides[0][d]) * (int64_t)sizeof(O)) % ao) ok = false;
      }
      if (!ok) vec = 1;
    }
    if constexpr (std::is_same<T, O>::value && std::is_same<compute_t<O>, float>::value && VECMAX > 1) {
      // a 128-byte-wide column band fits the shared memory of a cluster (softmax_band.cuh)
      constexpr int BW = 8 * VECMAX;
      const int64_t bctiles = (p.C + BW - 1) / BW;
      const int cl = vec == VECMAX && p.L >= 64 && !band_disabled() ? band_cluster(p.L * 128, outer_n * bctiles, ctx->sm_count, 8192) : 0;
      // Clusters of 8 that fill the GPU for only one to six rounds lose to the streaming kernels (their two cluster
      // barriers and the quantised last round weigh most there): f32 [4096,8192] axis 0 — 4.6 rounds — 79 µs in bands,
      // 76 µs streamed; bf16 61 vs 56 µs.  With many rounds ([64,4096,512] axis 1: 18 rounds, 268 vs 285 µs) or less
      // than one wave the bands win.
      const int64_t band_ctas = outer_n * bctiles * cl, band_wave = (int64_t)ctx->sm_count * 3;
      const bool few_rounds = cl == kBandMaxCl && band_ctas >= band_wave && band_ctas < 6 * band_wave;
      if (cl > 0 && !few_rounds && outer_n * bctiles * cl <= 0x7fffffffLL && std::llabs(p.sa_in) < (int64_t(1) << 40)) {
        BandParams q;
        memset(&q, 0, sizeof(q));
        q.kept = p.outer;
        q.L = p.L;
        q.per_cta = (p.L + cl - 1) / cl;
        q.sa_in = p.sa_in;
        q.sa_out = p.sa_out;
        q.C = p.C;
        q.ctiles = bctiles;
        q.cl = cl;
        q.log = log;
        q.use64 = p.use64;
        // enough bands for ≥ 4 rounds per resident cluster: the persistent, double-buffered kernel (one 1024-thread CTA per SM)
        const size_t smem = (size_t)q.per_cta * 128;
        auto kern = softmax_band_cols<T, VECMAX>;
        static const cudaError_t attr = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
        if (attr != cudaSuccess) return fail(HPTB_ERR_CUDA, "cudaFuncSetAttribute(softmax_band_cols) failed: %s", cudaGetErrorString(attr));
        HPTB_CUDA_CHECK(launch_kernel_cluster(kern, dim3((unsigned)(outer_n * bctiles * cl)), dim3(kSmThreads), smem, stream, (unsigned)cl, in, out, q));
        count_launches(1);
        return HPTB_OK;
      }
    }
    // 32 lanes per row segment unless that leaves the GPU short of CTAs and 8 lanes still fill whole sectors
    int tx = 32;
    if (outer_n * ((p.C + 32 * vec - 1) / (32 * vec)) < (int64_t)ctx->sm_count * 2 && vec > 1) tx = 8;
    if (band_tune_on() && vec > 1) {
      if (const char* e = getenv("HPTB_TUNE_SMC_TX")) tx = atoi(e) == 8 ? 8 : 32;
    }
    p.ctiles = (p.C + (int64_t)tx * vec - 1) / ((int64_t)tx * vec);
    const int64_t blocks = outer_n * p.ctiles;
    if (blocks > 0x7fffffffLL) return fail(HPTB_ERR_UNSUPPORTED, "softmax: grid too large");
    // splits of the axis: ONE wave of resident CTAs (f32 [4096,8192]: 256 column tiles × 2 splits 75 µs, × 5 splits —
    // 2.2 waves — 87 µs), each thread row keeping ≥ 2 batches of loads
    const int ty = kSmThreads / tx;
    static const std::array<int, 3> occs = [] {  // resident CTAs per SM of the three statistics kernels
      const void* ks[3] = {(const void*)softmax_cols_tiled<T, 1, 32, 1>, (const void*)softmax_cols_tiled<T, VECMAX, 8, 1>,
                           (const void*)softmax_cols_tiled<T, VECMAX, 32, 1>};
      std::array<int, 3> o{};
      for (int i = 0; i < 3; ++i)
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o[i], ks[i], kSmThreads, 0) != cudaSuccess || o[i] < 1) o[i] = 2;
      return o;
    }();
    const int occ = occs[vec > 1 ? (tx == 32 ? 2 : 1) : 0];
    // the largest split count that still fits one wave; a lone wave filling under 70 % of the slots is split in two anyway
    // (256 tiles on 444 slots: 101 µs unsplit, 85 µs as 1.15 waves), but 128 tiles × 5 on 592 slots — 1.08 waves — cost
    // 108 µs against 82 µs for × 4
    const int64_t slots = (int64_t)ctx->sm_count * occ;
    int64_t S = slots / blocks;
    if (S == 1 && blocks * 10 < slots * 7) S = 2;
    const int64_t max_s = p.L / ((int64_t)ty * 4 * 2);
    if (S > max_s) S = max_s;
    if (S > 64) S = 64;
    if (S < 1) S = 1;
    if (band_tune_on()) {  // development (tools/band_sweep.py)
      if (const char* e = getenv("HPTB_TUNE_SMC_S")) S = atoll(e) < 1 ? 1 : (atoll(e) > max_s && max_s >= 1 ? max_s : atoll(e));
    }
    p.rps = (p.L + S - 1) / S;
    S = (p.L + p.rps - 1) / p.rps;
    typedef compute_t<O> CT;
    Scratch part;
    CT* pp = nullptr;
    if (S > 1) {
      HPTB_TRY(part.get(ctx, (size_t)S * 2 * (size_t)blocks * tx * vec * sizeof(CT), stream));
      pp = static_cast<CT*>(part.ptr);
    }
#define HPTB_SMC(V, X, PH) HPTB_CUDA_CHECK(launch_kernel(softmax_cols_tiled<T, V, X, PH>, dim3((unsigned)blocks, (unsigned)(PH == 0 ? 1 : S)), dim3(kSmThreads), 0, stream, in, out, pp, p))
#define HPTB_SMC_ALL(PH)                                  \
  do {                                                    \
    if (vec > 1 && tx == 32) HPTB_SMC(VECMAX, 32, PH);    \
    else if (vec > 1) HPTB_SMC(VECMAX, 8, PH);            \
    else HPTB_SMC(1, 32, PH);                             \
  } while (0)
    if (S == 1) HPTB_SMC_ALL(0);
    else { HPTB_SMC_ALL(1); HPTB_SMC_ALL(2); count_launches(1); }
#undef HPTB_SMC_ALL
#undef HPTB_SMC
  } else if (cols_ok) {
    int64_t blocks = (M + kSmThreads - 1) / kSmThreads;
    if (blocks > 0x7fffffffLL) return fail(HPTB_ERR_UNSUPPORTED, "softmax: grid too large");
    HPTB_CUDA_CHECK(launch_kernel(softmax_cols<T>, dim3((unsigned)blocks), dim3(kSmThreads), 0, stream, in, out, p));
  } else {
    int64_t blocks = M < (int64_t)ctx->sm_count * 16 ? M : (int64_t)ctx->sm_count * 16;
    HPTB_CUDA_CHECK(launch_kernel(softmax_rows_stream<T>, dim3((unsigned)blocks), dim3(kSmThreads), 0, stream, in, out, p));
  }
  HPTB_CUDA_CHECK(cudaGetLastError());
  count_launches(1);
  return HPTB_OK;
}


// ---- layernorm ---------------------------------------------------------------------------------------------
// NormalizationOps::layernorm (hpt-traits/src/ops/normalization.rs:12-38; CPU semantics
// hpt/src/backends/cpu/tensor_internal/normalization.rs:49-200; device: hpt-cudakernels/src/normalization/
// {layernorm,layernorm_post}.cu): over the last k dims, y = (x − mean) / sqrt(var + eps) with the population
// variance Σ(x − mean)²/n (two passes over the register-resident row, as the reference computes it), then
// γ·y + β when given — fused here into the same kernel (the reference launches layernorm_post or a binary op).
// Output dtype O = FloatOutBinaryPromote<T,T>; statistics in the compute type of O.
struct LayerNormParams {
  DimWalk kept;  // stride_a = input, stride_b = output
  int64_t M, L;
  double eps;
  int32_t use64;
  int32_t nchunks;
  int32_t has_gamma, has_beta;
};

template <typename T, int VEC, int G, int NCH>
__global__ void __launch_bounds__(kSmThreads)
layernorm_rows_reg(const T* __restrict__ in, typename type_of_dtype<promote_ct(dtype_of<T>::value, dtype_of<T>::value, 1)>::type* __restrict__ out,
                   const typename type_of_dtype<promote_ct(dtype_of<T>::value, dtype_of<T>::value, 1)>::type* __restrict__ gamma,
                   const typename type_of_dtype<promote_ct(dtype_of<T>::value, dtype_of<T>::value, 1)>::type* __restrict__ beta,
                   LayerNormParams p) {
  pdl_prologue();
  typedef typename type_of_dtype<promote_ct(dtype_of<T>::value, dtype_of<T>::value, 1)>::type O;
  typedef compute_t<O> C;
  __shared__ C s_buf[kSmThreads / 32];
  const int tid = threadIdx.x;
  const int lane = tid & (G - 1);
  const int64_t row = (int64_t)blockIdx.x * (kSmThreads / G) + tid / G;
  const bool active = row < p.M;
  int64_t in_off = 0, out_off = 0;
  if (active) walk2(row, p.kept, p.use64, in_off, out_off);
  C x[NCH][VEC];
  C sum = (C)0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int e = (i * G + lane) * VEC;
    if (i < p.nchunks && active && e < (int)p.L) {
      Pack<T, VEC> v;
      load_pack<T, VEC>(v, in + in_off + e);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        x[i][k] = to_compute<O>(cast<O>(v.v[k]));
        sum += x[i][k];
      }
    } else {
#pragma unroll
      for (int k = 0; k < VEC; ++k) x[i][k] = (C)0;
    }
  }
  sum = group_reduce<AddOp, C, G>(sum, s_buf, (C)0);
  const C mean = sum / (C)p.L;
  C sq = (C)0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int e = (i * G + lane) * VEC;
    if (i < p.nchunks && e < (int)p.L) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        x[i][k] -= mean;
        sq += x[i][k] * x[i][k];
      }
    }
  }
  sq = group_reduce<AddOp, C, G>(sq, s_buf, (C)0);
  C rstd;
  if constexpr (std::is_same<C, float>::value) rstd = 1.0f / sqrtf(sq / (C)p.L + (C)p.eps);
  else rstd = 1.0 / sqrt(sq / (C)p.L + (C)p.eps);
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int e = (i * G + lane) * VEC;
    if (i < p.nchunks && active && e < (int)p.L) {
      Pack<O, VEC> g, b, o;
      if (p.has_gamma) load_pack_cached<O, VEC>(g, gamma + e);
      if (p.has_beta) load_pack_cached<O, VEC>(b, beta + e);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        // the reference rounds y to O before γ·y + β (two kernels); here the affine step stays in the compute type
        C y = x[i][k] * rstd;
        if (p.has_gamma) y *= to_compute<O>(g.v[k]);
        if (p.has_beta) y += to_compute<O>(b.v[k]);
        o.v[k] = from_compute<O>(y);
      }
      store_pack<O, VEC>(out + out_off + e, o);
    }
  }
}

// any row length: one CTA per row, three streaming passes (Σx, Σ(x−mean)², write); rows come back from L2
template <typename T>
__global__ void __launch_bounds__(kSmThreads)
layernorm_rows_stream(const T* __restrict__ in, typename type_of_dtype<promote_ct(dtype_of<T>::value, dtype_of<T>::value, 1)>::type* __restrict__ out,
                      const typename type_of_dtype<promote_ct(dtype_of<T>::value, dtype_of<T>::value, 1)>::type* __restrict__ gamma,
                      const typename type_of_dtype<promote_ct(dtype_of<T>::value, dtype_of<T>::value, 1)>::type* __restrict__ beta,
                      LayerNormParams p) {
  pdl_prologue();
  typedef typename type_of_dtype<promote_ct(dtype_of<T>::value, dtype_of<T>::value, 1)>::type O;
  typedef compute_t<O> C;
  __shared__ C s_buf[kSmThreads / 32];
  const int tid = threadIdx.x;
  for (int64_t row = blockIdx.x; row < p.M; row += gridDim.x) {
    int64_t in_off = 0, out_off = 0;
    walk2(row, p.kept, p.use64, in_off, out_off);
    const T* src = in + in_off;
    O* dst = out + out_off;
    C sum = (C)0;
    for (int64_t e = tid; e < p.L; e += kSmThreads) sum += to_compute<O>(cast<O>(load_one(src + e)));
    sum = group_reduce<AddOp, C, kSmThreads>(sum, s_buf, (C)0);
    const C mean = sum / (C)p.L;
    C sq = (C)0;
    for (int64_t e = tid; e < p.L; e += kSmThreads) {
      const C d = to_compute<O>(cast<O>(src[e])) - mean;
      sq += d * d;
    }
    sq = group_reduce<AddOp, C, kSmThreads>(sq, s_buf, (C)0);
    C rstd;
    if constexpr (std::is_same<C, float>::value) rstd = 1.0f / sqrtf(sq / (C)p.L + (C)p.eps);
    else rstd = 1.0 / sqrt(sq / (C)p.L + (C)p.eps);
    for (int64_t e = tid; e < p.L; e += kSmThreads) {
      C y = (to_compute<O>(cast<O>(src[e])) - mean) * rstd;
      if (p.has_gamma) y *= to_compute<O>(gamma[e]);
      if (p.has_beta) y += to_compute<O>(beta[e]);
      dst[e] = from_compute<O>(y);
    }
    __syncthreads();
  }
}

template <typename T>
hptb_status launch_layernorm(hptb_ctx* ctx, const Collapsed& c, const void* in_v, void* out_v, const void* gamma_v, const void* beta_v,
                             double eps, cudaStream_t stream) {
  typedef typename type_of_dtype<promote_ct(dtype_of<T>::value, dtype_of<T>::value, 1)>::type O;
  const T* in = static_cast<const T*>(in_v);
  O* out = static_cast<O*>(out_v);
  const O* gamma = static_cast<const O*>(gamma_v);
  const O* beta = static_cast<const O*>(beta_v);
  LayerNormParams p;
  memset(&p, 0, sizeof(p));
  int ad = -1, kept[kRedMaxDims], nk = 0, nred = 0;
  for (int d = c.ndim - 1; d >= 0; --d) {
    if (c.reduced[d]) { ad = d; ++nred; } else kept[nk++] = d;
  }
  if (nred > 1 || (ad >= 0 && (c.strides[1][ad] != 1 || c.strides[0][ad] != 1)))
    return fail(HPTB_ERR_UNSUPPORTED, "layernorm: the normalized dims must be contiguous in the input and the output");
  bool big = false;
  p.L = ad < 0 ? 1 : c.shape[ad];
  p.eps = eps;
  p.has_gamma = gamma != nullptr;
  p.has_beta = beta != nullptr;
  fill_walk(p.kept, c, kept, nk, true, big);
  int64_t M = 1;
  for (int i = 0; i < nk; ++i) M *= c.shape[kept[i]];
  p.M = M;
  if (M == 0 || p.L == 0) return HPTB_OK;
  if (!red_fits_u32(M)) big = true;
  p.use64 = big ? 1 : 0;
  constexpr int kMinSz = sizeof(T) < sizeof(O) ? sizeof(T) : sizeof(O);
  constexpr int VECMAX = 16 / kMinSz > 8 ? 8 : 16 / kMinSz;
  int vec = VECMAX;
  auto aligned = [&](int v) {
    if (v == 1) return true;
    size_t ai = sizeof(T) * v > 16 ? 16 : sizeof(T) * v, ao = sizeof(O) * v > 16 ? 16 : sizeof(O) * v;
    if (reinterpret_cast<uintptr_t>(in) % ai || reinterpret_cast<uintptr_t>(out) % ao) return false;
    if ((gamma && reinterpret_cast<uintptr_t>(gamma) % ao) || (beta && reinterpret_cast<uintptr_t>(beta) % ao)) return false;
    if (p.L % v) return false;
    for (int i = 0; i < nk; ++i) {
      if ((uint64_t)(std::llabs(c.strides[1][kept[i]]) * (int64_t)sizeof(T)) % ai) return false;
      if ((uint64_t)(std::llabs(c.strides[0][kept[i]]) * (int64_t)sizeof(O)) % ao) return false;
    }
    return true;
  };
  if (!aligned(vec)) vec = 1;
  const int64_t packs = (p.L + vec - 1) / vec;
  int G = 0;
  if (packs <= 32 * 4) G = 32;
  else if (packs <= (int64_t)kSmThreads * kSmChunks) G = kSmThreads;
  if (G) {
    p.nchunks = (int)((packs + G - 1) / G);
    int64_t blocks = (M + (kSmThreads / G) - 1) / (kSmThreads / G);
    if (blocks > 0x7fffffffLL) return fail(HPTB_ERR_UNSUPPORTED, "layernorm: grid too large");
#define HPTB_LN_LAUNCH2(V, GG, N) \
  HPTB_CUDA_CHECK(launch_kernel(layernorm_rows_reg<T, V, GG, N>, dim3((unsigned)blocks), dim3(kSmThreads), 0, stream, in, out, gamma, beta, p))
#define HPTB_LN_LAUNCH(V, GG)                                                 \
  do {                                                                        \
    if (GG == 32 || p.nchunks <= 2) HPTB_LN_LAUNCH2(V, GG, (GG == 32 ? 4 : 2)); \
    else if (p.nchunks <= 4) HPTB_LN_LAUNCH2(V, GG, 4);                       \
    else HPTB_LN_LAUNCH2(V, GG, kSmChunks);                                   \
  } while (0)
    if (vec > 1) {
      if (G == 32) HPTB_LN_LAUNCH(VECMAX, 32);
      else HPTB_LN_LAUNCH(VECMAX, kSmThreads);
    } else {
      if (G == 32) HPTB_LN_LAUNCH(1, 32);
      else HPTB_LN_LAUNCH(1, kSmThreads);
    }
#undef HPTB_LN_LAUNCH
#undef HPTB_LN_LAUNCH2
  } else {
    int64_t blocks = M < (int64_t)ctx->sm_count * 16 ? M : (int64_t)ctx->sm_count * 16;
    HPTB_CUDA_CHECK(launch_kernel(layernorm_rows_stream<T>, dim3((unsigned)blocks), dim3(kSmThreads), 0, stream, in, out, gamma, beta, p));
  }
  count_launches(1);
  return HPTB_OK;
}

}  // namespace
}  // namespace hptb

using namespace hptb;

extern "C" hptb_status hptb_softmax(hptb_ctx* ctx, const hptb_tensor* in, int axis, int log, hptb_tensor* out, void* stream) {
  if (!ctx) return fail(HPTB_ERR_INVALID, "softmax: null ctx");
  HPTB_TRY(validate_tensor(in, "softmax in"));
  HPTB_TRY(validate_tensor(out, "softmax out"));
  if (in->ndim == 0) return fail(HPTB_ERR_AXIS, "softmax: input has no dims");
  if (axis < 0) axis += in->ndim;
  if (axis < 0 || axis >= in->ndim) return fail(HPTB_ERR_AXIS, "softmax: axis %d out of range for ndim %d", axis, in->ndim);
  int odt = kFloatOutUnary[in->dtype];
  if (out->dtype != odt) return fail(HPTB_ERR_DTYPE, "softmax: out dtype is %s, expected %s", dtype_name(out->dtype), dtype_name(odt));
  bool same = in->ndim == out->ndim;
  for (int i = 0; same && i < in->ndim; ++i) same = in->shape[i] == out->shape[i];
  if (!same) return fail(HPTB_ERR_SHAPE, "softmax: out shape differs from the input shape");
  pass_direction(ctx, in->data, 0, false);  // a forward streaming pass (snake order, context.h)
  // The axis is the input's unit-stride dim but the output runs along another dim (x.t().softmax(0) into a fresh
  // contiguous tensor): every row kernel would scatter 4-byte writes over separate lines (f32 [8192,4096]ᵀ: 319 µs).
  // Normalise into a scratch laid out in the INPUT's dim order (register-resident row kernel), then let the
  // transposing copy kernel move it: two coalesced passes (≈ 95 µs) instead of one scattered one.
  if (in->strides[axis] == 1 && out->strides[axis] != 1 && in->shape[axis] >= 32 && numel(*in) >= (1 << 16)) {
    hptb_tensor tmp = *out;
    int order[HPTB_MAX_DIMS];
    for (int i = 0; i < in->ndim; ++i) order[i] = i;
    std::sort(order, order + in->ndim, [&](int a, int b) { return std::llabs(in->strides[a]) > std::llabs(in->strides[b]); });
    int64_t st = 1;
    for (int j = in->ndim - 1; j >= 0; --j) { tmp.strides[order[j]] = st; st *= in->shape[order[j]]; }
    if (tmp.strides[axis] == 1) {
      Scratch sc;
      HPTB_TRY(sc.get(ctx, (size_t)numel(*out) * dtype_size(out->dtype), stream));
      tmp.data = sc.ptr;
      HPTB_TRY(hptb_softmax(ctx, in, axis, log, &tmp, stream));
      return hptb_copy(ctx, &tmp, out, stream);
    }
  }
  uint8_t mask[HPTB_MAX_DIMS] = {0};
  mask[axis] = 1;
  int64_t strides[kMaxOperands][HPTB_MAX_DIMS] = {{0}};
  for (int i = 0; i < in->ndim; ++i) { strides[0][i] = out->strides[i]; strides[1][i] = in->strides[i]; }
  Collapsed c;
  collapse(in->ndim, in->shape, 2, strides, mask, &c);
  DeviceGuard g(ctx->device);
  switch (in->dtype) {
#define X(T, N, E) \
  case E: return launch_softmax<T>(ctx, c, in->data, out->data, log ? 1 : 0, (cudaStream_t)stream);
    HPTB_FOR_DTYPES(X)
#undef X
    default: return fail(HPTB_ERR_DTYPE, "softmax: bad dtype");
  }
}

extern "C" hptb_status hptb_layernorm(hptb_ctx* ctx, const hptb_tensor* in, int n_normalized_dims, const hptb_tensor* gamma,
                                      const hptb_tensor* beta, double eps, hptb_tensor* out, void* stream) {
  if (!ctx) return fail(HPTB_ERR_INVALID, "layernorm: null ctx");
  HPTB_TRY(validate_tensor(in, "layernorm in"));
  HPTB_TRY(validate_tensor(out, "layernorm out"));
  if (n_normalized_dims < 1 || n_normalized_dims > in->ndim)
    return fail(HPTB_ERR_SHAPE, "layernorm: normalized_shape has %d dims, the input %d", n_normalized_dims, in->ndim);
  const int odt = kFloatOutBinary[in->dtype][in->dtype];
  if (out->dtype != odt) return fail(HPTB_ERR_DTYPE, "layernorm: out dtype is %s, expected %s", dtype_name(out->dtype), dtype_name(odt));
  bool same = in->ndim == out->ndim;
  for (int i = 0; same && i < in->ndim; ++i) same = in->shape[i] == out->shape[i];
  if (!same) return fail(HPTB_ERR_SHAPE, "layernorm: out shape differs from the input shape");
  const int first = in->ndim - n_normalized_dims;
  const hptb_tensor* gb[2] = {gamma, beta};
  for (int t = 0; t < 2; ++t) {
    if (!gb[t]) continue;
    HPTB_TRY(validate_tensor(gb[t], t ? "layernorm beta" : "layernorm gamma"));
    if (gb[t]->dtype != odt) return fail(HPTB_ERR_DTYPE, "layernorm: gamma / beta must be %s", dtype_name(odt));
    bool ok = gb[t]->ndim == n_normalized_dims;
    int64_t exp = 1;
    for (int i = n_normalized_dims - 1; ok && i >= 0; --i) {
      ok = gb[t]->shape[i] == in->shape[first + i] && (gb[t]->shape[i] == 1 || gb[t]->strides[i] == exp);
      exp *= gb[t]->shape[i];
    }
    if (!ok) return fail(HPTB_ERR_SHAPE, "layernorm: gamma / beta must be contiguous tensors of the normalized shape");
  }
  // The kernels index gamma / beta by position in the normalized run, so that run must be the ROW-MAJOR order of the
  // normalized dims in `in` and in `out`.  The collapse pass below orders reduced dims by stride and would happily
  // merge a PERMUTED pair (in-place layernorm of x.transpose(-1, -2)) into one unit-stride run in memory order —
  // gamma would then land on the wrong elements without any error.  Checked here, before the collapse.
  for (const hptb_tensor* t : {in, static_cast<const hptb_tensor*>(out)}) {
    int64_t exp = 1;
    for (int i = in->ndim - 1; i >= first; --i) {
      if (t->shape[i] != 1 && t->strides[i] != exp)
        return fail(HPTB_ERR_UNSUPPORTED, "layernorm: the normalized dims must be dense and row-major in the input and the output "
                                          "(dim %d has stride %lld, expected %lld); call contiguous() first",
                    i, (long long)t->strides[i], (long long)exp);
      exp *= t->shape[i];
    }
  }
  uint8_t mask[HPTB_MAX_DIMS] = {0};
  for (int i = first; i < in->ndim; ++i) mask[i] = 1;
  int64_t strides[kMaxOperands][HPTB_MAX_DIMS] = {{0}};
  for (int i = 0; i < in->ndim; ++i) { strides[0][i] = out->strides[i]; strides[1][i] = in->strides[i]; }
  Collapsed c;
  collapse(in->ndim, in->shape, 2, strides, mask, &c);
  DeviceGuard g(ctx->device);
  switch (in->dtype) {
#define X(T, N, E) \
  case E: return launch_layernorm<T>(ctx, c, in->data, out->data, gamma ? gamma->data : nullptr, beta ? beta->data : nullptr, eps, (cudaStream_t)stream);
    HPTB_FOR_DTYPES(X)
#undef X
    default: return fail(HPTB_ERR_DTYPE, "layernorm: bad dtype");
  }
}
