// elementwise.cuh — the two elementwise ("map") kernel skeletons and their launcher.
//
// Every binary / unary / copy call of the hot path ends here after the collapse pass
// (layout.cpp).  Launch classes:
//   * contiguous + inner-contiguous  → map_rows_kernel: each thread moves VEC consecutive inner
//     elements with 128-bit accesses; outer dims (≤7 after collapse) are walked with by-value
//     fast-divmods, broadcast operands have inner stride 0 (one scalar load per row chunk).
//     This replaces `*_contiguous`, `*_contiguous_{lhs,rhs}_scalar` and the broadcast use of
//     `*_uncontiguous` (hpt-cudakernels/src/binary/binary_template.cuh:5-105).
//   * strided → map_tiled_kernel: tiles over (a = the output's unit-stride dim, b = the dim in which
//     a permuted input has unit stride).  Every thread owns an MA×MB register micro-tile: permuted
//     inputs are read with 128-bit loads ALONG b (8 lanes = one 128 B line), transposed in registers
//     (compile-time indexing, no shared memory, no barrier), and the result is written with 128-bit
//     stores ALONG a — both global sides stay sector-efficient and each thread keeps MA 16-byte
//     loads in flight.  Replaces `*_uncontiguous` (binary_template.cuh:48-61,
//     unary_template.cuh:48-62) whose lane-adjacent reads of a transposed operand are a full
//     stride apart.
// All indices are 64-bit capable; a 32-bit fast-divmod path is taken when the counts fit.
#pragma once
#include <cstdlib>

#include "common.h"
#include "launch.cuh"
#include "layout.h"
#include "map_plan.h"
#include "scalar.cuh"

namespace hptb {

constexpr int kMapThreads = 256;
#ifndef HPTB_MAP_UNROLL_SCALE
#define HPTB_MAP_UNROLL_SCALE 1
#endif
constexpr int kMaxOuter = HPTB_MAX_DIMS - 1;
constexpr int kScalarUnroll = 8;  // elements per thread of the scalar (VEC = 1) rows kernel

struct RowsParams {
  int64_t inner;         // elements in the inner dim
  int64_t cpr;           // chunks per row
  int64_t total_chunks;
  int32_t nouter;
  int32_t use64;         // 1: counts exceed 32 bits, use 64-bit division
  int32_t inner_stride[3];
  int32_t reuse[3];      // operand is re-read across rows (a broadcast outer dim): keep it in L1
  uint32_t outer_shape[kMaxOuter];  // innermost outer dim first
  FastDiv outer_div[kMaxOuter];
  FastDiv cpr_div;
  int64_t outer_stride[3][kMaxOuter];
};

struct TileParams {
  int64_t A, B;            // extents of the tile dims
  int64_t sa[3], sb[3];    // strides of each operand along a and b
  int64_t tiles_a, tiles_b, ntiles;
  int32_t nbatch;
  int32_t use64;
  int32_t mode[3];         // 0 = scalar access, 1 = vector along b (permuted operand), 2 = vector along a
  uint32_t batch_shape[kMaxOuter];
  FastDiv batch_div[kMaxOuter];
  FastDiv tiles_a_div, tiles_b_div;
  int64_t batch_stride[3][kMaxOuter];
};

template <int NOPS>
__device__ __forceinline__ void walk_outer(int64_t row, int nouter, int use64, const uint32_t* shape,
                                           const FastDiv* div, const int64_t (*stride)[kMaxOuter],
                                           int64_t (&off)[NOPS]) {
  if (!use64) {
    uint32_t r = (uint32_t)row;
#pragma unroll 1
    for (int i = 0; i < nouter; ++i) {
      uint32_t q = div[i].div(r);
      uint32_t rem = r - q * shape[i];
#pragma unroll
      for (int o = 0; o < NOPS; ++o) off[o] += (int64_t)rem * stride[o][i];
      r = q;
    }
  } else {
#pragma unroll 1
    for (int i = 0; i < nouter; ++i) {
      int64_t q = row / (int64_t)shape[i];
      int64_t rem = row - q * (int64_t)shape[i];
#pragma unroll
      for (int o = 0; o < NOPS; ++o) off[o] += rem * stride[o][i];
      row = q;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// flat kernel: every operand is one contiguous run (or a single broadcast scalar)
// ------------------------------------------------------------------------------------------------
// These kernels are issue-bound before they are HBM-bound if the per-chunk index arithmetic is not kept to a
// handful of instructions (ncu, profiles/r01b: the first general version spent 181 warp instructions per
// 3×512-byte chunk and sat at 75 % issue utilisation, 70 % of HBM peak).  So: the flat case does no index
// arithmetic at all, and the rows case below uses 32-bit offsets and one magic-number division per chunk.
struct FlatParams {
  int64_t n;             // elements
  int32_t stride[3];     // 1, or 0 for a broadcast scalar input
};

// Pack-level application of a unary functor.  Functors with a rare out-of-line path (sin/cos: large-argument
// reduction, ops.cuh) expose guard()/fast(): the range test becomes one predicate accumulated over the pack and
// the call never sits between two elements, so constants stay in registers and the fast path has no
// convergence barriers (26 → 19 instructions per element for sinf).
template <typename F, typename = void> struct has_guard : std::false_type {};
template <typename F> struct has_guard<F, std::enable_if_t<F::kGuarded>> : std::true_type {};
template <typename F, typename O, typename A, int E>
__device__ __forceinline__ void apply_pack(const F& f, Pack<O, E>& r, const Pack<A, E>& x) {
  if constexpr (has_guard<F>::value) {
    bool slow = false;
#pragma unroll
    for (int k = 0; k < E; ++k) slow |= f.guard(x.v[k]);
    if (!slow) {
#pragma unroll
      for (int k = 0; k < E; ++k) r.v[k] = f.fast(x.v[k]);
      return;
    }
  }
#pragma unroll
  for (int k = 0; k < E; ++k) r.v[k] = f(x.v[k]);
}

}  // namespace hptb
#include "tma_tile.cuh"  // needs apply_pack
namespace hptb {

template <int NIN, int VEC, int UNROLL, typename F, typename O, typename A, typename B>
__global__ void __launch_bounds__(kMapThreads)
map_flat_kernel(O* __restrict__ out, const A* __restrict__ a, const B* __restrict__ b, FlatParams p, F f) {
  pdl_prologue();
  constexpr int64_t kPerCta = (int64_t)kMapThreads * UNROLL * VEC;
  const int64_t e0 = (int64_t)blockIdx.x * kPerCta + (int64_t)threadIdx.x * VEC;
  Pack<A, VEC> pa[UNROLL];
  Pack<B, VEC> pb[UNROLL];
  // broadcast scalars are read once
  A sa = load_one(a);
  B sb = load_one(b);
  if ((int64_t)(blockIdx.x + 1) * kPerCta <= p.n) {  // full CTA: no bounds checks
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t e = e0 + (int64_t)u * (kMapThreads * VEC);
      if (p.stride[1] != 0) load_pack<A, VEC>(pa[u], a + e);
      if (NIN == 2 && p.stride[2] != 0) load_pack<B, VEC>(pb[u], b + e);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t e = e0 + (int64_t)u * (kMapThreads * VEC);
      Pack<O, VEC> po;
      if constexpr (NIN == 1) {
        if (p.stride[1] == 0) {
#pragma unroll
          for (int k = 0; k < VEC; ++k) pa[u].v[k] = sa;
        }
        apply_pack<F, O, A, VEC>(f, po, pa[u]);
      } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const A x = p.stride[1] != 0 ? pa[u].v[k] : sa;
          if constexpr (NIN == 2) po.v[k] = f(x, p.stride[2] != 0 ? pb[u].v[k] : sb);
        }
      }
      store_pack<O, VEC>(out + e, po);
    }
    return;
  }
  // last CTA: element-wise bounds
#pragma unroll 1
  for (int u = 0; u < UNROLL; ++u) {
    const int64_t e = e0 + (int64_t)u * (kMapThreads * VEC);
#pragma unroll 1
    for (int k = 0; k < VEC; ++k) {
      if (e + k >= p.n) break;
      const A x = p.stride[1] != 0 ? load_one(a + e + k) : sa;
      if constexpr (NIN == 2) out[e + k] = f(x, p.stride[2] != 0 ? load_one(b + e + k) : sb);
      else out[e + k] = f(x);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// rows kernel: inner-contiguous layouts (row / column broadcasts, row-strided views), 32-bit offsets
// ------------------------------------------------------------------------------------------------
struct Rows32Params {
  uint32_t total_chunks;
  uint32_t total_elems;             // ragged kernel only
  uint32_t cpr;                     // chunks per row (ragged kernel: elements per row)
  FastDiv cpr_div;
  int32_t nouter;
  int32_t inner_stride[3];          // 1 or 0
  int32_t reuse[3];                 // operand re-read across rows (broadcast outer dim): keep it in L1
  uint32_t outer_shape[kMaxOuter];  // innermost outer dim first
  FastDiv outer_div[kMaxOuter];
  int32_t outer_stride[3][kMaxOuter];
};

template <int NIN, int VEC, int UNROLL, typename F, typename O, typename A, typename B>
__global__ void __launch_bounds__(kMapThreads)
map_rows_kernel(O* __restrict__ out, const A* __restrict__ a, const B* __restrict__ b, Rows32Params p, F f) {
  pdl_prologue();
  // VEC == 1: one element per lane and chunk, any inner stride per operand (stepped slices, rows that do not start
  // on a 16-byte boundary) — consecutive lanes still touch consecutive elements, so a warp's access is one run
  const uint32_t c0 = blockIdx.x * (uint32_t)(kMapThreads * UNROLL) + threadIdx.x;
  Pack<A, VEC> pa[UNROLL];
  Pack<B, VEC> pb[UNROLL];
  int32_t oo[UNROLL];
  bool ok[UNROLL];
  const int last = p.nouter - 1;
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    const uint32_t c = c0 + (uint32_t)u * kMapThreads;
    ok[u] = c < p.total_chunks;
    if (ok[u]) {
      uint32_t r = p.cpr_div.div(c);
      const int32_t e = (int32_t)((c - r * p.cpr) * VEC);
      int32_t o0 = e, o1 = e * p.inner_stride[1], o2 = NIN == 2 ? e * p.inner_stride[2] : 0;
#pragma unroll 1
      for (int i = 0; i < last; ++i) {  // all but the outermost dim (usually none)
        const uint32_t q = p.outer_div[i].div(r);
        const int32_t rem = (int32_t)(r - q * p.outer_shape[i]);
        o0 += rem * p.outer_stride[0][i];
        o1 += rem * p.outer_stride[1][i];
        if (NIN == 2) o2 += rem * p.outer_stride[2][i];
        r = q;
      }
      o0 += (int32_t)r * p.outer_stride[0][last];
      o1 += (int32_t)r * p.outer_stride[1][last];
      if (NIN == 2) o2 += (int32_t)r * p.outer_stride[2][last];
      oo[u] = o0;
      if (p.inner_stride[1] == 0) {
        // a column operand ([N,1]): every chunk of a row re-reads the same element — let it live in L1 (the no-allocate
        // load crossed to L2 per chunk: f32 a + column [8192,1] 104.7 µs against 80.0 µs for a + row)
        Pack<A, 1> s1;
        load_pack_cached<A, 1>(s1, a + o1);
        const A s = s1.v[0];
#pragma unroll
        for (int k = 0; k < VEC; ++k) pa[u].v[k] = s;
      } else if (p.reuse[1]) load_pack_cached<A, VEC>(pa[u], a + o1);
      else load_pack<A, VEC>(pa[u], a + o1);
      if constexpr (NIN == 2) {
        if (p.inner_stride[2] == 0) {
          Pack<B, 1> s1;
          load_pack_cached<B, 1>(s1, b + o2);
          const B s = s1.v[0];
#pragma unroll
          for (int k = 0; k < VEC; ++k) pb[u].v[k] = s;
        } else if (p.reuse[2]) load_pack_cached<B, VEC>(pb[u], b + o2);
        else load_pack<B, VEC>(pb[u], b + o2);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    if (!ok[u]) continue;
    Pack<O, VEC> po;
    if constexpr (NIN == 2) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) po.v[k] = f(pa[u].v[k], pb[u].v[k]);
    } else {
      apply_pack<F, O, A, VEC>(f, po, pa[u]);
    }
    store_pack<O, VEC>(out + oo[u], po);
  }
}

// ------------------------------------------------------------------------------------------------
// ragged kernel: CONTIGUOUS output; inputs whose rows start off a pack boundary, or with a ragged or stepped inner dim
// ------------------------------------------------------------------------------------------------
// The output is walked as ONE flat run in aligned 16-byte packs.  A pack's elements come from one input row, or from the
// end of one and the start of the next, through element-sized L1-allocating loads: the VEC loads of a warp touch the
// same sectors, so one of them goes to L2 and the rest hit L1.  One index computation and one 128-bit store per VEC
// elements, where the scalar rows kernel pays both per element (f32 exp of a[5:8000, 3:8100]: 116 → 95 µs; 77 µs —
// the copy roofline — with the one-row fast path below: ncu had the first version at 39 instructions per element).
template <int NIN, int VEC, int UNROLL, typename F, typename O, typename A, typename B>
__global__ void __launch_bounds__(kMapThreads)
map_ragged_kernel(O* __restrict__ out, const A* __restrict__ a, const B* __restrict__ b, Rows32Params p, F f) {
  pdl_prologue();
  const uint32_t c0 = blockIdx.x * (uint32_t)(kMapThreads * UNROLL) + threadIdx.x;
  const int last = p.nouter - 1;
  auto row_base = [&](uint32_t r, int32_t& o1, int32_t& o2) {
    o1 = 0;
    o2 = 0;
#pragma unroll 1
    for (int i = 0; i < last; ++i) {  // all but the outermost dim (usually none)
      const uint32_t q = p.outer_div[i].div(r);
      const int32_t rem = (int32_t)(r - q * p.outer_shape[i]);
      o1 += rem * p.outer_stride[1][i];
      if (NIN == 2) o2 += rem * p.outer_stride[2][i];
      r = q;
    }
    o1 += (int32_t)r * p.outer_stride[1][last];
    if (NIN == 2) o2 += (int32_t)r * p.outer_stride[2][last];
  };
  Pack<A, VEC> pa[UNROLL];
  Pack<B, VEC> pb[UNROLL];
  int32_t cnt[UNROLL];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    const uint32_t c = c0 + (uint32_t)u * kMapThreads;
    cnt[u] = 0;
    if (c < p.total_chunks) {
      const uint32_t f0 = c * VEC;
      uint32_t r = p.cpr_div.div(f0);
      int32_t e = (int32_t)(f0 - r * p.cpr);
      int32_t o1, o2;
      row_base(r, o1, o2);
      const uint32_t left = p.total_elems - f0;
      cnt[u] = left < (uint32_t)VEC ? (int32_t)left : VEC;
      if (cnt[u] == VEC && e + VEC <= (int32_t)p.cpr) {  // the whole pack lies in one row: no per-element bookkeeping
        const A* pa1 = a + o1 + e * p.inner_stride[1];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          Pack<A, 1> s1;
          load_pack_cached<A, 1>(s1, pa1 + k * p.inner_stride[1]);
          pa[u].v[k] = s1.v[0];
        }
        if constexpr (NIN == 2) {
          const B* pb1 = b + o2 + e * p.inner_stride[2];
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            Pack<B, 1> s2;
            load_pack_cached<B, 1>(s2, pb1 + k * p.inner_stride[2]);
            pb[u].v[k] = s2.v[0];
          }
        }
        continue;
      }
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        if (k < cnt[u]) {
          if (e == (int32_t)p.cpr) {  // the pack runs into the next row (rows shorter than a pack: again and again)
            e = 0;
            row_base(++r, o1, o2);
          }
          Pack<A, 1> s1;
          load_pack_cached<A, 1>(s1, a + o1 + e * p.inner_stride[1]);
          pa[u].v[k] = s1.v[0];
          if constexpr (NIN == 2) {
            Pack<B, 1> s2;
            load_pack_cached<B, 1>(s2, b + o2 + e * p.inner_stride[2]);
            pb[u].v[k] = s2.v[0];
          }
          ++e;
        } else {
          pa[u].v[k] = pa[u].v[0];
          if constexpr (NIN == 2) pb[u].v[k] = pb[u].v[0];
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    if (cnt[u] == 0) continue;
    Pack<O, VEC> po;
    if constexpr (NIN == 2) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) po.v[k] = f(pa[u].v[k], pb[u].v[k]);
    } else {
      apply_pack<F, O, A, VEC>(f, po, pa[u]);
    }
    O* dst = out + (size_t)(c0 + (uint32_t)u * kMapThreads) * VEC;
    if (cnt[u] == VEC) store_pack<O, VEC>(dst, po);
    else
      for (int k = 0; k < cnt[u]; ++k) dst[k] = po.v[k];
  }
}

// ------------------------------------------------------------------------------------------------
// tiled kernel: permuted / strided layouts (register micro-tile transpose)
// ------------------------------------------------------------------------------------------------
constexpr int kTileAG = 4, kTileBG = 8;    // lanes of a warp: 4 micro-tiles along a × 8 along b
constexpr int kTileWA = 4, kTileWB = 2;    // warps of a CTA: 4 along a × 2 along b

template <typename O, typename A, typename B>
struct TileGeom {
  static constexpr int szmin_in = sizeof(A) < sizeof(B) ? sizeof(A) : sizeof(B);
  static constexpr int ma0 = 16 / sizeof(O), mb0 = 16 / szmin_in;
  static constexpr int MA = ma0 > 8 ? 8 : ma0;  // elements along a per thread (one 16 B store for ≥2-byte outputs)
  static constexpr int MB = mb0 > 8 ? 8 : mb0;  // elements along b per thread (one 16 B load for ≥2-byte inputs)
  static constexpr int TA = kTileWA * kTileAG * MA;
  static constexpr int TB = kTileWB * kTileBG * MB;
};

// load an MA×MB micro-tile of operand X at (a0, b0) into v[ai][bj]
template <typename X, int MA, int MB>
__device__ __forceinline__ void load_micro(X (&v)[MA][MB], const X* __restrict__ base, int64_t sa, int64_t sb, int mode,
                                           bool full, int na, int nb) {
  if (mode == 1 && full) {  // rows along b
#pragma unroll
    for (int i = 0; i < MA; ++i) {
      Pack<X, MB> pk;
      load_pack<X, MB>(pk, base + (int64_t)i * sa);
#pragma unroll
      for (int j = 0; j < MB; ++j) v[i][j] = pk.v[j];
    }
  } else if (mode == 2 && full) {  // columns along a
#pragma unroll
    for (int j = 0; j < MB; ++j) {
      Pack<X, MA> pk;
      load_pack<X, MA>(pk, base + (int64_t)j * sb);
#pragma unroll
      for (int i = 0; i < MA; ++i) v[i][j] = pk.v[i];
    }
  } else {
#pragma unroll
    for (int i = 0; i < MA; ++i)
#pragma unroll
      for (int j = 0; j < MB; ++j)
        if (i < na && j < nb) v[i][j] = load_one(base + (int64_t)i * sa + (int64_t)j * sb);
  }
}

template <int NIN, typename F, typename O, typename A, typename B>
__global__ void __launch_bounds__(kMapThreads)
map_tiled_kernel(O* __restrict__ out, const A* __restrict__ a, const B* __restrict__ b, TileParams p, F f) {
  pdl_prologue();
  typedef TileGeom<O, A, B> G;
  constexpr int MA = G::MA, MB = G::MB;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // micro-tile coordinates inside the CTA tile
  const int ua = ((warp % kTileWA) * kTileAG + (lane / kTileBG)) * MA;
  const int ub = ((warp / kTileWA) * kTileBG + (lane % kTileBG)) * MB;

  for (int64_t t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
    int64_t ta, tb, batch;
    if (!p.use64) {
      uint32_t q = p.tiles_a_div.div((uint32_t)t);
      ta = (uint32_t)t - q * (uint32_t)p.tiles_a;
      uint32_t q2 = p.tiles_b_div.div(q);
      tb = q - q2 * (uint32_t)p.tiles_b;
      batch = q2;
    } else {
      int64_t q = t / p.tiles_a;
      ta = t - q * p.tiles_a;
      int64_t q2 = q / p.tiles_b;
      tb = q - q2 * p.tiles_b;
      batch = q2;
    }
    const int64_t a0 = ta * G::TA + ua, b0 = tb * G::TB + ub;
    if (a0 >= p.A || b0 >= p.B) continue;
    int64_t off[3] = {0, 0, 0};
    walk_outer<3>(batch, p.nbatch, p.use64, p.batch_shape, p.batch_div, p.batch_stride, off);
    const int na = (int)((p.A - a0) < MA ? (p.A - a0) : MA);
    const int nb = (int)((p.B - b0) < MB ? (p.B - b0) : MB);
    const bool full = na == MA && nb == MB;

    A va[MA][MB];
    B vb[MA][MB];
    load_micro<A, MA, MB>(va, a + off[1] + a0 * p.sa[1] + b0 * p.sb[1], p.sa[1], p.sb[1], p.mode[1], full, na, nb);
    if constexpr (NIN == 2)
      load_micro<B, MA, MB>(vb, b + off[2] + a0 * p.sa[2] + b0 * p.sb[2], p.sa[2], p.sb[2], p.mode[2], full, na, nb);

    O* dst = out + off[0] + a0 * p.sa[0] + b0 * p.sb[0];
    if (p.mode[0] == 2 && full) {
#pragma unroll
      for (int j = 0; j < MB; ++j) {
        Pack<O, MA> pk;
#pragma unroll
        for (int i = 0; i < MA; ++i) {
          if constexpr (NIN == 2) pk.v[i] = f(va[i][j], vb[i][j]);
          else pk.v[i] = f(va[i][j]);
        }
        store_pack<O, MA>(dst + (int64_t)j * p.sb[0], pk);
      }
    } else {
#pragma unroll
      for (int j = 0; j < MB; ++j)
#pragma unroll
        for (int i = 0; i < MA; ++i)
          if (i < na && j < nb) {
            O r;
            if constexpr (NIN == 2) r = f(va[i][j], vb[i][j]);
            else r = f(va[i][j]);
            dst[(int64_t)i * p.sa[0] + (int64_t)j * p.sb[0]] = r;
          }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// transposing tile kernel: permuted operands staged through shared memory (same-size element types)
// ------------------------------------------------------------------------------------------------
// ncu on the register micro-tile kernel above (profiles/r01b): l1tex throughput 81 %, DRAM 57 % — its 16-byte
// stores leave the lanes of a quarter-warp 32 KB apart, so one store instruction costs 32 L1 wavefronts instead
// of 4 and the LSU pipe, not HBM, bounds it.  Here both global sides are quarter-warp contiguous (8 lanes × 16 B
// = one 128-byte line): a permuted operand is read with 16-byte loads along its own unit-stride dim b,
// scattered into a [b][a] shared-memory tile with element-sized stores, and read back with 16-byte loads along a
// — the output's unit-stride dim — for the 16-byte global stores.  16-byte chunk c of tile row b lives at chunk
// c ^ ((b / E) & 7), which makes the scatter and the gather bank-conflict free.  16 L1 wavefronts per 512 bytes
// in + out instead of 36.
// (kSmemModeStaged / Direct / Scalar: tma_tile.cuh)

// one tile; pointers are already at the tile origin, in-tile offsets are 32-bit (the host checks the strides)
template <int NIN, typename F, typename T, bool FULL>
__device__ __forceinline__ void smem_tile_body(T* __restrict__ dst, const T* __restrict__ in0, const T* __restrict__ in1,
                                               const int32_t (&sa)[3], const int32_t (&sb)[3], const int32_t (&mode)[3],
                                               int rem_a, int rem_b, T* __restrict__ sm0, F f) {
  constexpr int E = 16 / sizeof(T);
  constexpr int TA = 16 * E;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = lane >> 3, l8 = lane & 7;
  const T* in[2] = {in0, in1};
  // phase 1: staged operands — 16-byte loads along b, element scatter into sm[b][a]
  {
    Pack<T, E> v[2][E];
    bool ok[E];
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const int g = i * 8 + warp;  // group of 4 consecutive a-rows × one half of the b extent
      const int ra = (g % (4 * E)) * 4 + q;
      const int pb = (g / (4 * E)) * 8 + l8;
      ok[i] = FULL || (ra < rem_a && pb * E < rem_b);
      if (ok[i]) {
#pragma unroll
        for (int o = 0; o < NIN; ++o)
          if (NIN == 1 || mode[o + 1] == kSmemModeStaged) load_pack<T, E>(v[o][i], in[o] + (int64_t)ra * sa[o + 1] + pb * E);
      }
    }
    int nst = 0;
#pragma unroll
    for (int o = 0; o < NIN; ++o) {
      if (NIN == 2 && mode[o + 1] != kSmemModeStaged) continue;
      T* sm = sm0 + nst * (TA * TA);
      ++nst;
#pragma unroll
      for (int i = 0; i < E; ++i) {
        if (!ok[i]) continue;
        const int g = i * 8 + warp;
        const int ra = (g % (4 * E)) * 4 + q;
        const int pb = (g / (4 * E)) * 8 + l8;
        T* d = sm + (pb * E) * TA + (((ra / E) ^ (pb & 7)) * E) + (ra % E);
#pragma unroll
        for (int k = 0; k < E; ++k) d[k * TA] = v[o][i].v[k];
      }
    }
  }
  __syncthreads();
  // phase 2: 16-byte gathers along a, compute, 16-byte stores along a
#pragma unroll
  for (int i = 0; i < E; ++i) {
    const int g = i * 8 + warp;
    const int rb = (g % (4 * E)) * 4 + q;
    const int ca = (g / (4 * E)) * 8 + l8;
    if (!FULL && (rb >= rem_b || ca * E >= rem_a)) continue;
    Pack<T, E> x[2];
    int nst = 0;
#pragma unroll
    for (int o = 0; o < NIN; ++o) {
      if (NIN == 1 || mode[o + 1] == kSmemModeStaged) {
        const T* sm = sm0 + nst * (TA * TA);
        ++nst;
        x[o] = *reinterpret_cast<const Pack<T, E>*>(sm + rb * TA + ((ca ^ ((rb / E) & 7)) * E));
      } else if (mode[o + 1] == kSmemModeDirect) {
        load_pack<T, E>(x[o], in[o] + (int64_t)rb * sb[o + 1] + ca * E);
      } else {
        const T s = load_one(in[o]);
#pragma unroll
        for (int k = 0; k < E; ++k) x[o].v[k] = s;
      }
    }
    Pack<T, E> r;
    if constexpr (NIN == 2) {
#pragma unroll
      for (int k = 0; k < E; ++k) r.v[k] = f(x[0].v[k], x[1].v[k]);
    } else {
      apply_pack<F, T, T, E>(f, r, x[0]);
    }
    store_pack<T, E>(dst + (int64_t)rb * sb[0] + ca * E, r);
  }
}

struct SmemTileParams {
  int64_t A, B;
  int64_t tiles_a, tiles_b;
  int32_t sa[3], sb[3];  // element strides along a and b (|stride| < 2^31, checked on the host)
  int32_t mode[3];
  int32_t nbatch, use64;
  uint32_t batch_shape[kMaxOuter];
  FastDiv batch_div[kMaxOuter];
  FastDiv tiles_a_div, tiles_b_div;
  int64_t batch_stride[3][kMaxOuter];
};

// guarded functors carry an out-of-line call whose ABI spill area would cost the sixth resident CTA (48 vs 40 registers)
// 0 = unspecified: a literal 1 lets ptxas spend up to 255 registers (exp went 40 -> 56 registers, 88 -> 98 us)
template <typename F> constexpr int tiled_min_ctas() { return has_guard<F>::value ? 6 : 0; }
template <int NIN, typename F, typename T>
__global__ void __launch_bounds__(kMapThreads, tiled_min_ctas<F>())
map_tiled_smem_kernel(T* __restrict__ out, const T* __restrict__ a, const T* __restrict__ b, SmemTileParams p, F f) {
  pdl_prologue();
  constexpr int E = 16 / sizeof(T);  // elements per 16-byte pack
  constexpr int TA = 16 * E, TB = 16 * E;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm0 = reinterpret_cast<T*>(smem_raw);  // staged operand tiles, [TB][TA] each
  // one tile per CTA.  Usual case: a 3-D grid (tile along a, tile along b, batch) — no division; grids whose
  // y/z extents exceed 65535 fall back to a linear index
  int64_t ta, tb, batch;
  if (gridDim.y > 1 || gridDim.z > 1 || p.use64 == 2) {
    ta = blockIdx.x;
    tb = blockIdx.y;
    batch = blockIdx.z;
  } else {
    const int64_t t = blockIdx.x;
    if (!p.use64) {
      uint32_t qq = p.tiles_a_div.div((uint32_t)t);
      ta = (uint32_t)t - qq * (uint32_t)p.tiles_a;
      uint32_t q2 = p.tiles_b_div.div(qq);
      tb = qq - q2 * (uint32_t)p.tiles_b;
      batch = q2;
    } else {
      int64_t qq = t / p.tiles_a;
      ta = t - qq * p.tiles_a;
      int64_t q2 = qq / p.tiles_b;
      tb = qq - q2 * p.tiles_b;
      batch = q2;
    }
  }
  const int64_t a0 = ta * TA, b0 = tb * TB;
  int64_t off[3] = {0, 0, 0};
  if (p.nbatch > 0) walk_outer<3>(batch, p.nbatch, p.use64 == 1, p.batch_shape, p.batch_div, p.batch_stride, off);
  T* dst = out + off[0] + a0 * p.sa[0] + b0 * p.sb[0];
  const T* in0 = a + off[1] + a0 * p.sa[1] + b0 * p.sb[1];
  const T* in1 = b + off[2] + a0 * p.sa[2] + b0 * p.sb[2];
  const int64_t ra64 = p.A - a0, rb64 = p.B - b0;
  if (ra64 >= TA && rb64 >= TB) smem_tile_body<NIN, F, T, true>(dst, in0, in1, p.sa, p.sb, p.mode, TA, TB, sm0, f);
  else smem_tile_body<NIN, F, T, false>(dst, in0, in1, p.sa, p.sb, p.mode, (int)(ra64 < TA ? ra64 : TA), (int)(rb64 < TB ? rb64 : TB), sm0, f);
}

// ------------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------------
// elements per thread-chunk: the WIDEST operand moves 16 bytes per access (narrower operands 8/4/2 — still
// lane-contiguous).  A wider pack would make the widest operand issue two 16-byte accesses 32 bytes apart per
// lane, i.e. half-filled sectors per instruction (measured: f32 ⊕ i64 → f64 at 52 % vs 16-byte stores).
template <typename O, typename A, typename B>
constexpr int map_vec_width() {
  int mx = sizeof(O) > sizeof(A) ? sizeof(O) : sizeof(A);
  mx = mx > (int)sizeof(B) ? mx : (int)sizeof(B);
  return 16 / mx;
}

inline bool fits_u32(int64_t v) { return v >= 0 && v < (int64_t(1) << 31); }

// development switches for tools/sweep.py (only consulted when HPTB_TUNE=1 at load)
inline bool tune_flag(const char* name) {
  static const bool on = [] { const char* e = getenv("HPTB_TUNE"); return e && e[0] == '1'; }();
  if (!on) return false;
  const char* e = getenv(name);
  return e && e[0] == '1';
}

// rows launch: the innermost collapsed dim is walked by consecutive lanes (VEC elements each), the other dims by one
// magic-number division per chunk; 32-bit element offsets (larger tensors take the runtime-typed kernel, which is
// 64-bit).  VEC == 1 accepts any inner stride per operand.
template <int NIN, int VEC, int UNROLL, typename F, typename O, typename A, typename B>
hptb_status launch_rows(const Collapsed& c, O* out, const A* a, const B* b, F f, cudaStream_t stream) {
  const int nd = c.ndim;
  if (nd < 1) return HPTB_FALLBACK;
  Rows32Params p;
  memset(&p, 0, sizeof(p));
  const int64_t inner = c.shape[nd - 1];
  if (inner % VEC) return HPTB_FALLBACK;
  p.nouter = nd - 1;
  int64_t rows = 1;
  for (int o = 0; o <= NIN; ++o) {
    if (std::llabs(c.strides[o][nd - 1]) > 0x7fffffffLL) return HPTB_FALLBACK;
    p.inner_stride[o] = (int32_t)c.strides[o][nd - 1];
    int64_t span = (inner - 1) * std::llabs(c.strides[o][nd - 1]);  // largest |offset| this operand can reach
    for (int i = 0; i < p.nouter; ++i) {
      const int d = nd - 2 - i;
      span += (c.shape[d] - 1) * std::llabs(c.strides[o][d]);
      if (std::llabs(c.strides[o][d]) > 0x7fffffffLL) return HPTB_FALLBACK;
      p.outer_stride[o][i] = (int32_t)c.strides[o][d];
      if (o > 0 && c.strides[o][d] == 0) p.reuse[o] = 1;
    }
    if (span > 0x7fffffffLL - 64) return HPTB_FALLBACK;
  }
  for (int i = 0; i < p.nouter; ++i) {
    const int d = nd - 2 - i;
    rows *= c.shape[d];
    p.outer_shape[i] = (uint32_t)c.shape[d];
    p.outer_div[i] = FastDiv((uint32_t)c.shape[d]);
  }
  if (p.nouter == 0) {  // a single dim: one row (the kernel always folds an outermost dim)
    p.nouter = 1;
    p.outer_shape[0] = 1;
    p.outer_div[0] = FastDiv(1u);
  }
  const int64_t cpr = inner / VEC;
  const int64_t total = rows * cpr;
  if (total >= (int64_t(1) << 31)) return HPTB_FALLBACK;
  p.cpr = (uint32_t)cpr;
  p.cpr_div = FastDiv((uint32_t)cpr);
  p.total_chunks = (uint32_t)total;
  int64_t blocks = (total + kMapThreads * UNROLL - 1) / (kMapThreads * UNROLL);
  HPTB_CUDA_CHECK(launch_kernel(map_rows_kernel<NIN, VEC, UNROLL, F, O, A, B>, dim3((unsigned)blocks), dim3(kMapThreads), 0, stream, out, a, b, p, f));
  return HPTB_OK;
}

// The ragged kernel serves a dense row-major output of < 2^31 elements whose operands fit 32-bit offsets; otherwise
// HPTB_FALLBACK and the caller takes the scalar rows kernel.
template <int NIN, int VEC, typename F, typename O, typename A, typename B>
hptb_status launch_ragged(const Collapsed& c, O* out, const A* a, const B* b, F f, cudaStream_t stream) {
  const int nd = c.ndim;
  if (nd < 1 || nd - 1 > kMaxOuter) return HPTB_FALLBACK;
  static const bool off = [] { const char* t = getenv("HPTB_TUNE"); const char* e = getenv("HPTB_TUNE_NO_RAGGED"); return t && t[0] == '1' && e && e[0] == '1'; }();
  if (off) return HPTB_FALLBACK;
  if (reinterpret_cast<uintptr_t>(out) % (sizeof(O) * VEC > 16 ? 16 : sizeof(O) * VEC)) return HPTB_FALLBACK;
  int64_t dense = 1;
  for (int d = nd - 1; d >= 0; --d) {
    if (c.strides[0][d] != dense) return HPTB_FALLBACK;
    dense *= c.shape[d];
  }
  if (dense >= (int64_t(1) << 31) - VEC) return HPTB_FALLBACK;
  Rows32Params p;
  memset(&p, 0, sizeof(p));
  p.nouter = nd - 1;
  for (int o = 1; o <= NIN; ++o) {
    if (std::llabs(c.strides[o][nd - 1]) > 0x7fffffffLL) return HPTB_FALLBACK;
    p.inner_stride[o] = (int32_t)c.strides[o][nd - 1];
    int64_t span = (c.shape[nd - 1] - 1) * std::llabs(c.strides[o][nd - 1]);
    for (int i = 0; i < p.nouter; ++i) {
      const int d = nd - 2 - i;
      span += (c.shape[d] - 1) * std::llabs(c.strides[o][d]);
      if (std::llabs(c.strides[o][d]) > 0x7fffffffLL) return HPTB_FALLBACK;
      p.outer_stride[o][i] = (int32_t)c.strides[o][d];
    }
    if (span > 0x7fffffffLL - 64) return HPTB_FALLBACK;
  }
  for (int i = 0; i < p.nouter; ++i) {
    const int d = nd - 2 - i;
    p.outer_shape[i] = (uint32_t)c.shape[d];
    p.outer_div[i] = FastDiv((uint32_t)c.shape[d]);
  }
  if (p.nouter == 0) {
    p.nouter = 1;
    p.outer_shape[0] = 1;
    p.outer_div[0] = FastDiv(1u);
  }
  p.cpr = (uint32_t)c.shape[nd - 1];
  p.cpr_div = FastDiv(p.cpr);
  p.total_elems = (uint32_t)dense;
  p.total_chunks = (uint32_t)((dense + VEC - 1) / VEC);
  constexpr int UNROLL = VEC >= 8 ? 1 : 2;  // 8 element loads in flight per thread; 16 lose their L1 hits (f32 window 107 vs 95 µs)
  const int64_t blocks = ((int64_t)p.total_chunks + kMapThreads * UNROLL - 1) / (kMapThreads * UNROLL);
  HPTB_CUDA_CHECK(launch_kernel(map_ragged_kernel<NIN, VEC, UNROLL, F, O, A, B>, dim3((unsigned)blocks), dim3(kMapThreads), 0, stream, out, a, b, p, f));
  return HPTB_OK;
}

// TMA-staged transposing launch (tma_tile.cuh) for a tile plan with exactly ONE operand read along b (mode 1), the output
// and any partner operand along a (mode 2) or scalar.  HPTB_FALLBACK: the layout is outside what the tensor map or the
// kernel's grid can describe — the caller takes its other kernels.
template <int NIN, typename F, typename O>
hptb_status launch_tma_tile(const TileParams& p, int64_t batch, bool big, O* out, const O* a, const O* b, F f, cudaStream_t stream) {
  typedef TmaGeom<(sizeof(O) == 1 ? 2 : sizeof(O))> TG;
  constexpr int BWE = TG::BW * (sizeof(O) == 1 ? 2 : 1);
  if (p.nbatch > 3 || batch > 65535 || big || tma_disabled() || p.A * p.B * batch < 4096 || tune_flag("HPTB_TUNE_NO_TMA")) return HPTB_FALLBACK;
  int so = 0;  // the staged operand
  for (int o = 1; o <= NIN; ++o)
    if (p.mode[o] == 1) so = o;
  if (!so) return HPTB_FALLBACK;
  for (int o = 0; o <= NIN; ++o)
    if (std::llabs(p.sa[o]) > 0x7fffffffLL || std::llabs(p.sb[o]) > 0x7fffffffLL) return HPTB_FALLBACK;
  const int64_t tiles_a = (p.A + kTmaTA - 1) / kTmaTA, tiles_b = (p.B + (int64_t)kTmaSub * BWE - 1) / ((int64_t)kTmaSub * BWE);
  if (tiles_a > 0x7fffffffLL || tiles_b > 65535) return HPTB_FALLBACK;
  CUtensorMap tmap;
  const O* staged = so == 1 ? a : b;
  if (!tma_make_map<O>(&tmap, staged, p.A, p.B, p.sa[so], p.nbatch, p.batch_shape, p.batch_stride[so])) return HPTB_FALLBACK;
  TmaTileParams q;
  memset(&q, 0, sizeof(q));
  q.A = p.A;
  q.B = p.B;
  q.out_sb = p.sb[0];
  q.nbatch = p.nbatch;
  const int other = NIN == 2 ? 3 - so : 0;
  const O* in1 = other == 1 ? a : b;
  if (other) {
    q.in1_sb = p.sb[other];
    q.in1_mode = p.mode[other] == 2 ? kSmemModeDirect : kSmemModeScalar;
    q.swap = other == 1 ? 1 : 0;
  }
  for (int i = 0; i < 3; ++i) {
    q.batch_shape[i] = i < p.nbatch ? p.batch_shape[i] : 1u;
    q.batch_out[i] = i < p.nbatch ? p.batch_stride[0][i] : 0;
    q.batch_in1[i] = (other && i < p.nbatch) ? p.batch_stride[other][i] : 0;
  }
  auto kern = map_tma_tile_kernel<NIN, F, O>;
  constexpr int smem_bytes = kTmaSub * TG::kSubStride + 64;
  static const cudaError_t attr = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (attr != cudaSuccess) return fail(HPTB_ERR_CUDA, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed: %s", cudaGetErrorString(attr));
  HPTB_CUDA_CHECK(launch_kernel(kern, dim3((unsigned)tiles_a, (unsigned)tiles_b, (unsigned)batch), dim3(kTmaThreads), smem_bytes, stream, out, in1, tmap, q, f));
  return HPTB_OK;
}

template <int NIN, typename F, typename O, typename A, typename B>
hptb_status launch_map(const MapPlan& plan, F f, cudaStream_t stream) {
  const Collapsed& c = plan.c;
  if (c.numel == 0) return HPTB_OK;
  O* out = static_cast<O*>(plan.ptr[0]);
  const A* a = static_cast<const A*>(plan.ptr[1]);
  const B* b = NIN == 2 ? static_cast<const B*>(plan.ptr[2]) : reinterpret_cast<const B*>(plan.ptr[1]);
  constexpr int VEC = map_vec_width<O, A, B>();
  const size_t esz[3] = {sizeof(O), sizeof(A), sizeof(B)};

  if (c.launch_class != HPTB_CLASS_STRIDED) {
    const int nd = c.ndim;
    constexpr int UNROLL = (VEC >= 16 ? 1 : VEC >= 8 ? 2 : 4) * HPTB_MAP_UNROLL_SCALE;  // ≥ 64 bytes of the widest operand per thread
    // every unit-stride operand must keep 16 B (or pack-size) alignment at the start of every row
    for (int o = 0; o <= NIN; ++o) {
      if (nd && c.strides[o][nd - 1] == 0) continue;
      size_t align = esz[o] * VEC > 16 ? 16 : esz[o] * VEC;
      bool misaligned = reinterpret_cast<uintptr_t>(plan.ptr[o]) % align;
      for (int d = 0; d + 1 < nd; ++d)
        if ((uint64_t)(std::llabs(c.strides[o][d]) * (int64_t)esz[o]) % align) misaligned = true;
      // rows that do not start on a pack boundary (a[5:8000, 3:8100]): one element per lane, still one run per warp
      if (misaligned) {
        if (!nd) return HPTB_FALLBACK;
        if constexpr (VEC > 1) {
          const hptb_status st = launch_ragged<NIN, VEC, F, O, A, B>(c, out, a, b, f, stream);
          if (st != HPTB_FALLBACK) return st;
        }
        return launch_rows<NIN, 1, kScalarUnroll, F, O, A, B>(c, out, a, b, f, stream);
      }
    }
    if (nd >= 2 && c.shape[nd - 1] % VEC) {  // ragged rows
      if constexpr (VEC > 1) {
        const hptb_status st = launch_ragged<NIN, VEC, F, O, A, B>(c, out, a, b, f, stream);
        if (st != HPTB_FALLBACK) return st;
      }
      return launch_rows<NIN, 1, kScalarUnroll, F, O, A, B>(c, out, a, b, f, stream);
    }
    if (nd <= 1) {  // flat: one contiguous run per operand (or a broadcast scalar)
      FlatParams p;
      p.n = nd ? c.shape[0] : 1;
      for (int o = 0; o <= NIN; ++o) p.stride[o] = nd ? (int32_t)c.strides[o][0] : 1;
      p.stride[0] = 1;
      if (NIN == 1) p.stride[2] = p.stride[1];
      constexpr int64_t per_cta = (int64_t)kMapThreads * UNROLL * VEC;
      int64_t blocks = (p.n + per_cta - 1) / per_cta;
      if (blocks > 0x7fffffffLL) return fail(HPTB_ERR_UNSUPPORTED, "elementwise: tensor too large for one launch");
      HPTB_CUDA_CHECK(launch_kernel(map_flat_kernel<NIN, VEC, UNROLL, F, O, A, B>, dim3((unsigned)blocks), dim3(kMapThreads), 0, stream, out, a, b, p, f));
      return HPTB_OK;
    }
    return launch_rows<NIN, VEC, UNROLL, F, O, A, B>(c, out, a, b, f, stream);
  }

  // strided: choose tile dims
  const int nd = c.ndim;
  int da = -1;
  {  // a = dim with the smallest |out stride| (unit stride when the output is contiguous)
    int64_t best = 0;
    for (int d = 0; d < nd; ++d) {
      int64_t s = std::llabs(c.strides[0][d]);
      if (da < 0 || s < best) { da = d; best = s; }
    }
  }
  int db = -1;
  for (int o = 1; o <= NIN && db < 0; ++o) {
    if (c.strides[o][da] == 1 || c.strides[o][da] == 0) continue;  // already coalesced along a
    for (int d = 0; d < nd; ++d)
      if (d != da && c.strides[o][d] == 1) { db = d; break; }
  }
  if (db < 0) {
    // No operand has its unit stride on another dim: nothing to transpose, the inner dim is merely stepped
    // (a[:, ::2]) or reversed.  Lanes along the output's fastest dim with scalar accesses are as coalesced as such a
    // layout allows (the tile kernel would run 64-element tiles here: f32 a[:, ::2].exp() 1146 µs).
    if (da == nd - 1) {
      if constexpr (VEC > 1) {
        const hptb_status sr = launch_ragged<NIN, VEC, F, O, A, B>(c, out, a, b, f, stream);
        if (sr != HPTB_FALLBACK) return sr;
      }
      hptb_status st = launch_rows<NIN, 1, kScalarUnroll, F, O, A, B>(c, out, a, b, f, stream);
      if (st != HPTB_FALLBACK) return st;
    }
  }
  if (db < 0) {  // take the innermost remaining dim (largest extent wins ties)
    for (int d = nd - 1; d >= 0; --d)
      if (d != da) { db = d; break; }
  }
  typedef TileGeom<O, A, B> G;
  TileParams p;
  memset(&p, 0, sizeof(p));
  p.A = c.shape[da];
  p.B = db >= 0 ? c.shape[db] : 1;
  for (int o = 0; o <= NIN; ++o) {
    p.sa[o] = c.strides[o][da];
    p.sb[o] = db >= 0 ? c.strides[o][db] : 0;
    // vector access needs a unit stride along the vector dim and 16 B (pack) alignment of every micro-tile row
    const int along_b = G::MB, along_a = G::MA;
    auto aligned = [&](int vec, int unit_dim) {
      size_t al = esz[o] * vec > 16 ? 16 : esz[o] * vec;
      if (al <= esz[o]) return true;
      if (reinterpret_cast<uintptr_t>(plan.ptr[o]) % al) return false;
      for (int d = 0; d < nd; ++d) {
        if (d == unit_dim) continue;
        if ((uint64_t)(std::llabs(c.strides[o][d]) * (int64_t)esz[o]) % al) return false;
      }
      return true;
    };
    p.mode[o] = 0;
    if (o > 0 && db >= 0 && p.sb[o] == 1 && aligned(along_b, db)) p.mode[o] = 1;
    else if (p.sa[o] == 1 && aligned(along_a, da)) p.mode[o] = 2;
  }
  bool big = false;
  int nb = 0;  // innermost batch dim first
  for (int d = nd - 1; d >= 0; --d) {
    if (d == da || d == db) continue;
    if (!fits_u32(c.shape[d])) big = true;
    p.batch_shape[nb] = (uint32_t)c.shape[d];
    p.batch_div[nb] = FastDiv((uint32_t)c.shape[d]);
    for (int o = 0; o <= NIN; ++o) p.batch_stride[o][nb] = c.strides[o][d];
    ++nb;
  }
  p.nbatch = nb;
  p.tiles_a = (p.A + G::TA - 1) / G::TA;
  p.tiles_b = (p.B + G::TB - 1) / G::TB;
  int64_t batch = 1;
  for (int i = 0; i < nb; ++i) batch *= p.batch_shape[i];
  p.ntiles = p.tiles_a * p.tiles_b * batch;
  if (p.ntiles >= (int64_t(1) << 32) || !fits_u32(p.tiles_a) || !fits_u32(p.tiles_b)) big = true;
  p.use64 = big ? 1 : 0;
  p.tiles_a_div = FastDiv(big ? 1u : (uint32_t)p.tiles_a);
  p.tiles_b_div = FastDiv(big ? 1u : (uint32_t)p.tiles_b);
  // same-size element types with a permuted operand: the shared-memory transposing kernel
  if constexpr (std::is_same<O, A>::value && std::is_same<O, B>::value && sizeof(O) >= 2) {
    constexpr int E = 16 / (int)sizeof(O);
    bool ok = db >= 0 && p.mode[0] == 2 && p.A % E == 0 && p.B % E == 0 && !tune_flag("HPTB_TUNE_NO_SMEMT");
    int nstaged = 0;
    TileParams ps = p;
    for (int o = 1; o <= NIN && ok; ++o) {
      if (p.mode[o] == 1) { ps.mode[o] = kSmemModeStaged; ++nstaged; }
      else if (p.mode[o] == 2) ps.mode[o] = kSmemModeDirect;
      else if (p.sa[o] == 0 && p.sb[o] == 0) ps.mode[o] = kSmemModeScalar;
      else ok = false;
    }
    for (int o = 0; o <= NIN && ok; ++o)
      if (std::llabs(p.sa[o]) > 0x7fffffffLL || std::llabs(p.sb[o]) > 0x7fffffffLL) ok = false;
    // one staged operand whose layout a tensor map can describe: TMA-staged tiles (tma_tile.cuh)
    if constexpr (sizeof(O) == 2 || sizeof(O) == 4) {
      if (ok && nstaged == 1) {
        const hptb_status st = launch_tma_tile<NIN, F, O>(p, batch, big, out, reinterpret_cast<const O*>(a), reinterpret_cast<const O*>(b), f, stream);
        if (st != HPTB_FALLBACK) return st;
      }
    }
    if (ok && nstaged > 0) {
      constexpr int T2 = 16 * E;
      SmemTileParams q;
      memset(&q, 0, sizeof(q));
      q.A = p.A;
      q.B = p.B;
      for (int o = 0; o < 3; ++o) { q.sa[o] = (int32_t)p.sa[o]; q.sb[o] = (int32_t)p.sb[o]; q.mode[o] = ps.mode[o]; }
      q.nbatch = p.nbatch;
      for (int i = 0; i < kMaxOuter; ++i) {
        q.batch_shape[i] = p.batch_shape[i];
        q.batch_div[i] = p.batch_div[i];
        for (int o = 0; o < 3; ++o) q.batch_stride[o][i] = p.batch_stride[o][i];
      }
      q.tiles_a = (p.A + T2 - 1) / T2;
      q.tiles_b = (p.B + T2 - 1) / T2;
      const int64_t ntiles = q.tiles_a * q.tiles_b * batch;
      const bool big2 = big || ntiles >= (int64_t(1) << 32);
      q.use64 = big2 ? 1 : 0;
      q.tiles_a_div = FastDiv(big2 ? 1u : (uint32_t)q.tiles_a);
      q.tiles_b_div = FastDiv(big2 ? 1u : (uint32_t)q.tiles_b);
      if (ntiles > 0x7fffffffLL) return fail(HPTB_ERR_UNSUPPORTED, "elementwise: tensor too large for one launch");
      const size_t smem = (size_t)nstaged * T2 * T2 * sizeof(O);
      auto kern = map_tiled_smem_kernel<NIN, F, O>;
      if (smem > 48 * 1024) {
        static const cudaError_t attr = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * T2 * T2 * (int)sizeof(O));
        if (attr != cudaSuccess) return fail(HPTB_ERR_CUDA, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed: %s", cudaGetErrorString(attr));
      }
      dim3 grid((unsigned)ntiles, 1, 1);
      if (q.tiles_a <= 0x7fffffffLL && q.tiles_b <= 65535 && batch <= 65535) {
        grid = dim3((unsigned)q.tiles_a, (unsigned)q.tiles_b, (unsigned)batch);
        q.use64 = 2;  // 3-D grid: blockIdx is the tile coordinate (walk_outer then takes the 32-bit path: batch < 65536)
      }
      HPTB_CUDA_CHECK(launch_kernel(kern, grid, dim3(kMapThreads), smem, stream, out, reinterpret_cast<const O*>(a), reinterpret_cast<const O*>(b), q, f));
      return HPTB_OK;
    }
  }
  // 1-byte types (i8 / u8 / bool): the same TMA tiles on a u16 view of the staged operand, bytes pulled apart with prmt
  // (the register micro-tile kernel below runs a transposed i8 copy at 0.35 of peak)
  if constexpr (std::is_same<O, A>::value && std::is_same<O, B>::value && sizeof(O) == 1) {
    bool ok1 = db >= 0 && p.mode[0] == 2 && p.A % 8 == 0 && p.B % 2 == 0;
    int nstaged = 0;
    for (int o = 1; o <= NIN && ok1; ++o) {
      if (p.mode[o] == 1) ++nstaged;
      else if (p.mode[o] == 2) {}
      else if (p.sa[o] == 0 && p.sb[o] == 0) {}
      else ok1 = false;
    }
    if (ok1 && nstaged == 1) {
      const hptb_status st = launch_tma_tile<NIN, F, O>(p, batch, big, out, reinterpret_cast<const O*>(a), reinterpret_cast<const O*>(b), f, stream);
      if (st != HPTB_FALLBACK) return st;
    }
  }
  int64_t blocks = p.ntiles < 0x7fffffffLL ? p.ntiles : 0x7fffffffLL;
  HPTB_CUDA_CHECK(launch_kernel(map_tiled_kernel<NIN, F, O, A, B>, dim3((unsigned)blocks), dim3(kMapThreads), 0, stream, out, a, b, p, f));
  return HPTB_OK;
}

}  // namespace hptb
