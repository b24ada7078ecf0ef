// fused_inst.cu — elementwise → reduce fusion: out = reduce(lhs ⊕ rhs, axes) in ONE pass,
// without materialising lhs ⊕ rhs.  One translation unit per binary op (-DHPTB_OP / -DHPTB_OPNAME); the reduction
// (sum, max, min, sum_square) is a launch-time choice among instantiations.
//
// Semantics are exactly those of the two calls it replaces — NormalBinOps::{add_,sub_,mul_} followed by
// NormalReduce — including the rounding of the elementwise result to T before it is accumulated (f16 / bf16), so
// fused and unfused results agree bit for bit whenever the accumulation order does.  The reference has no fused
// kernels; its conceptual hook is the `fuse` macro of hpt-codegen (hpt-codegen/src/fuse/*).  Config 1
// ((A + B).sum(1), B a broadcast row) reads 67 MB instead of moving 201 MB.
//
// Fast path: both operands have dtype T, lhs is made of one aligned unit-stride run per output with ≤ 1 kept dim
// (the lean rows layout), rhs has the same layout, is a row broadcast over the outputs, or is a scalar.  Everything
// else reports HPTB_FALLBACK and the API layer composes the two ordinary kernels through a temporary.
#include "dtypes_x.h"
#include "map_plan.h"
#include "ops.cuh"
#include "reduce.cuh"

namespace hptb {

namespace {

struct FusedParams {
  int64_t M;
  int64_t a_stride, b_stride, out_stride;  // between consecutive outputs (b_stride 0: rhs row is shared)
  double count;
  uint32_t cpr;
  int32_t logG;
  int32_t b_inner;  // 1: rhs walks the run, 0: one rhs element per output (scalar / column broadcast)
  int32_t fold_out;
};

template <typename Op, typename T, int VEC>
__global__ void __launch_bounds__(kRedThreads, lean_min_blocks<Op, true>())
fused_rows_kernel(const T* __restrict__ a, const T* __restrict__ b, typename Op::Out* __restrict__ out, FusedParams p) {
  pdl_prologue();
  typedef typename Op::Acc Acc;
  typedef typename Op::Local Local;
  typedef BinaryFn<HPTB_OP, T, T, T> F;
  constexpr int UNROLL = HPTB_RED_UNROLL;
  __shared__ Acc s_part[kRedThreads / 32];
  const uint32_t tid = threadIdx.x;
  const uint32_t G = 1u << p.logG;
  const uint32_t g = tid & (G - 1);
  const int64_t m = (((int64_t)blockIdx.x * kRedThreads) >> p.logG) + (tid >> p.logG);
  const bool active = m < p.M;
  Local acc[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) acc[k] = Op::local_identity();
  if (active) {
    const T* ra = a + m * p.a_stride;
    const T* rb = b + m * p.b_stride;
    const uint32_t n = p.cpr;
    const F f{};
    int32_t it = 0;
    if (p.b_inner) {
      for (uint32_t c = g; c < n; c += G * UNROLL, it += UNROLL) {
        Pack<T, VEC> va[UNROLL], vb[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
          if (c + (uint32_t)u * G < n) {
            load_pack<T, VEC>(va[u], ra + (size_t)(c + (uint32_t)u * G) * VEC);
            if (p.b_stride == 0) load_pack_cached<T, VEC>(vb[u], rb + (size_t)(c + (uint32_t)u * G) * VEC);  // shared row: keep it in L1
            else load_pack<T, VEC>(vb[u], rb + (size_t)(c + (uint32_t)u * G) * VEC);
          }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
          if (c + (uint32_t)u * G < n) {
            Pack<T, VEC> r;
#pragma unroll
            for (int k = 0; k < VEC; ++k) r.v[k] = f(va[u].v[k], vb[u].v[k]);
            Op::template accumulate_pack<VEC>(acc, r, it + u);
          }
      }
    } else {
      const T sb = load_one(rb);
      for (uint32_t c = g; c < n; c += G * UNROLL, it += UNROLL) {
        Pack<T, VEC> va[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
          if (c + (uint32_t)u * G < n) load_pack<T, VEC>(va[u], ra + (size_t)(c + (uint32_t)u * G) * VEC);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
          if (c + (uint32_t)u * G < n) {
            Pack<T, VEC> r;
#pragma unroll
            for (int k = 0; k < VEC; ++k) r.v[k] = f(va[u].v[k], sb);
            Op::template accumulate_pack<VEC>(acc, r, it + u);
          }
      }
    }
  }
  Acc acc_t = Op::finish(acc[0], g, G, VEC, 0);
#pragma unroll
  for (int k = 1; k < VEC; ++k) acc_t = Op::combine(acc_t, Op::finish(acc[k], g, G, VEC, k));
  if (G <= 32) {
    acc_t = warp_reduce<Op, Acc>(acc_t, (int)G);
    if (active && g == 0) red_store<Op>(out, out, m * p.out_stride, acc_t, p.count, p.fold_out);
    return;
  }
  acc_t = warp_reduce<Op, Acc>(acc_t, 32);
  if ((tid & 31) == 0) s_part[tid >> 5] = acc_t;
  __syncthreads();
  if (g == 0 && active) {
    const int w0 = tid >> 5, nw = G >> 5;
    acc_t = s_part[w0];
    for (int w = 1; w < nw; ++w) acc_t = Op::combine(acc_t, s_part[w0 + w]);
    red_store<Op>(out, out, m * p.out_stride, acc_t, p.count, p.fold_out);
  }
}

// The shared-row case of config 1 — (A + row).sum(1), one 256-thread CTA per output, ≤ UNROLL packs per thread: a CTA
// walks several outputs and keeps ITS packs of the row in registers, so the row is read once per CTA instead of once
// per output (through L1, but as many load instructions as A itself: 13.7 µs against 10.9 µs for the plain sum(1)).
template <typename Op, typename T, int VEC>
__global__ void __launch_bounds__(kRedThreads, 5)
fused_rows_shared_kernel(const T* __restrict__ a, const T* __restrict__ b, typename Op::Out* __restrict__ out, FusedParams p) {
  pdl_prologue();
  typedef typename Op::Acc Acc;
  typedef typename Op::Local Local;
  typedef BinaryFn<HPTB_OP, T, T, T> F;
  constexpr int UNROLL = HPTB_RED_UNROLL;
  __shared__ Acc s_part[kRedThreads / 32];
  const uint32_t tid = threadIdx.x, n = p.cpr;
  const F f{};
  Pack<T, VEC> vb[UNROLL];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u)
    if (tid + (uint32_t)u * kRedThreads < n) load_pack_cached<T, VEC>(vb[u], b + (size_t)(tid + (uint32_t)u * kRedThreads) * VEC);
  // the next output's packs of A are requested before the current output is reduced: a CTA streams without a gap
  Pack<T, VEC> va[UNROLL], vn[UNROLL];
  auto fetch = [&](int64_t m, Pack<T, VEC> (&dst)[UNROLL]) {
    if (m >= p.M) return;
    const T* ra = a + m * p.a_stride;
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      if (tid + (uint32_t)u * kRedThreads < n) load_pack<T, VEC>(dst[u], ra + (size_t)(tid + (uint32_t)u * kRedThreads) * VEC);
  };
  fetch(blockIdx.x, vn);
  for (int64_t m = blockIdx.x; m < p.M; m += gridDim.x) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) va[u] = vn[u];
    fetch(m + gridDim.x, vn);
    Local acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = Op::local_identity();
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      if (tid + (uint32_t)u * kRedThreads < n) {
        Pack<T, VEC> r;
#pragma unroll
        for (int k = 0; k < VEC; ++k) r.v[k] = f(va[u].v[k], vb[u].v[k]);
        Op::template accumulate_pack<VEC>(acc, r, u);
      }
    Acc t = Op::finish(acc[0], tid, kRedThreads, VEC, 0);
#pragma unroll
    for (int k = 1; k < VEC; ++k) t = Op::combine(t, Op::finish(acc[k], tid, kRedThreads, VEC, k));
    t = warp_reduce<Op, Acc>(t, 32);
    __syncthreads();  // s_part of the previous output has been read
    if ((tid & 31) == 0) s_part[tid >> 5] = t;
    __syncthreads();
    if (tid == 0) {
      t = s_part[0];
      for (int w = 1; w < kRedThreads / 32; ++w) t = Op::combine(t, s_part[w]);
      red_store<Op>(out, out, m * p.out_stride, t, p.count, p.fold_out);
    }
  }
}

template <typename Op, typename T>
hptb_status launch_fused(const FusedPlan& plan, cudaStream_t stream) {
  const Collapsed& c = plan.c;
  constexpr int VECMAX = 16 / sizeof(T) > 8 ? 8 : 16 / sizeof(T);
  if constexpr (VECMAX < 2) return HPTB_FALLBACK;
  else {
    int kept[kRedMaxDims], red[kRedMaxDims], nk = 0, nr = 0;
    for (int d = c.ndim - 1; d >= 0; --d) {
      if (c.reduced[d]) red[nr++] = d; else kept[nk++] = d;
    }
    if (nr != 1 || nk > 1) return HPTB_FALLBACK;
    const int rd = red[0];
    const int64_t L = c.shape[rd];
    const int64_t M = nk ? c.shape[kept[0]] : 1;
    if (M == 0 || L == 0) return HPTB_FALLBACK;
    if (c.strides[1][rd] != 1 || L % VECMAX || L / VECMAX >= (int64_t(1) << 31)) return HPTB_FALLBACK;
    const int64_t bs_inner = c.strides[2][rd];
    if (bs_inner != 1 && bs_inner != 0) return HPTB_FALLBACK;
    const int64_t a_stride = nk ? c.strides[1][kept[0]] : 0, b_stride = nk ? c.strides[2][kept[0]] : 0;
    auto al16 = [](const void* p, int64_t stride_elems) {
      return reinterpret_cast<uintptr_t>(p) % 16 == 0 && (uint64_t)(std::llabs(stride_elems) * (int64_t)sizeof(T)) % 16 == 0;
    };
    if (!al16(plan.lhs, a_stride)) return HPTB_FALLBACK;
    if (bs_inner == 1 && !al16(plan.rhs, b_stride)) return HPTB_FALLBACK;
    FusedParams p;
    memset(&p, 0, sizeof(p));
    p.M = M;
    p.a_stride = a_stride;
    p.b_stride = b_stride;
    p.out_stride = nk ? c.strides[0][kept[0]] : 0;
    p.count = plan.count;
    p.cpr = (uint32_t)(L / VECMAX);
    p.b_inner = bs_inner == 1;
    p.fold_out = plan.fold_out;
    int logG = 8;
    while (logG > 0 && ((int64_t)1 << logG) * HPTB_RED_UNROLL > p.cpr) --logG;
    p.logG = logG;
    const int64_t blocks = (M + (kRedThreads >> logG) - 1) / (kRedThreads >> logG);
    // few, very long outputs would need the split machinery of the general kernel: leave them to the unfused path
    if (blocks > 0x7fffffffLL || (blocks < plan.ctx->sm_count && p.cpr > 16 * kRedThreads * HPTB_RED_UNROLL)) return HPTB_FALLBACK;
    if (p.b_inner && b_stride == 0 && logG == 8 && p.cpr <= (uint32_t)kRedThreads * HPTB_RED_UNROLL && M >= (int64_t)plan.ctx->sm_count * 6 &&
        !Op::kIndexed) {
      HPTB_CUDA_CHECK(launch_kernel(fused_rows_shared_kernel<Op, T, VECMAX>, dim3((unsigned)(plan.ctx->sm_count * 5)), dim3(kRedThreads), 0, stream,
                                    static_cast<const T*>(plan.lhs), static_cast<const T*>(plan.rhs), static_cast<typename Op::Out*>(plan.out), p));
      return HPTB_OK;
    }
    HPTB_CUDA_CHECK(launch_kernel(fused_rows_kernel<Op, T, VECMAX>, dim3((unsigned)blocks), dim3(kRedThreads), 0, stream,
                                  static_cast<const T*>(plan.lhs), static_cast<const T*>(plan.rhs), static_cast<typename Op::Out*>(plan.out), p));
    return HPTB_OK;
  }
}

template <typename T>
hptb_status dispatch(const FusedPlan& plan, cudaStream_t s) {
  switch (plan.red_op) {
    case HPTB_SUM: return launch_fused<ReduceOp<HPTB_SUM, T>, T>(plan, s);
    case HPTB_MAX: return launch_fused<ReduceOp<HPTB_MAX, T>, T>(plan, s);
    case HPTB_MIN: return launch_fused<ReduceOp<HPTB_MIN, T>, T>(plan, s);
    case HPTB_SUM_SQUARE: return launch_fused<ReduceOp<HPTB_SUM_SQUARE, T>, T>(plan, s);
    default: return HPTB_FALLBACK;
  }
}

}  // namespace
}  // namespace hptb

extern "C" hptb::FusedLauncher HPTB_CAT(hptb_fused_, HPTB_OPNAME)(int dt) {
  using namespace hptb;
  switch (dt) {
#define X(T, N, E) \
  case E: return &dispatch<T>;
    HPTB_FOR_DTYPES(X)
#undef X
    default: return nullptr;
  }
}
