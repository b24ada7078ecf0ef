// elementwise.cuh — the two elementwise ("map") kernel skeletons and their launcher.
//
// Every binary / unary / copy call of the hot path ends here after the collapse pass
// (layout.cpp).  Launch classes:
//   * contiguous + inner-contiguous  → map_rows_kernel: each thread moves VEC consecutive inner
//     elements with 128-bit accesses; outer dims (≤7 after collapse) are walked with by-value
//     fast-divmods, broadcast operands have inner stride 0 (one scalar load per row chunk).
//     This replaces `*_contiguous`, `*_contiguous_{lhs,rhs}_scalar` and the broadcast use of
//     `*_uncontiguous` (hpt-cudakernels/src/binary/binary_template.cuh:5-105).
//   * strided → map_tiled_kernel: 64×64 tiles over (a = the output's unit-stride dim,
//     b = the dim in which a permuted input has unit stride); permuted inputs are read along b
//     (coalesced) into padded shared memory and re-read along a, so both global sides stay
//     coalesced.  Replaces `*_uncontiguous` (binary_template.cuh:48-61, unary_template.cuh:48-62)
//     whose lane-adjacent reads of a transposed operand are a full stride apart.
// All indices are 64-bit capable; a 32-bit fast-divmod path is taken when the counts fit.
#pragma once
#include "common.h"
#include "layout.h"
#include "map_plan.h"
#include "scalar.cuh"

namespace hptb {

constexpr int kMapThreads = 256;
constexpr int kMaxOuter = HPTB_MAX_DIMS - 1;
constexpr int kTile = 64;
constexpr int kTilePitch = kTile + 1;

struct RowsParams {
  int64_t inner;         // elements in the inner dim
  int64_t cpr;           // chunks per row
  int64_t total_chunks;
  int32_t nouter;
  int32_t use64;         // 1: counts exceed 32 bits, use 64-bit division
  int32_t inner_stride[3];
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
  int32_t transposed[3];   // operand is staged through shared memory
  int32_t smem_off[3];     // byte offset of the operand's tile in dynamic shared memory
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
// rows kernel: contiguous and inner-contiguous (broadcast) layouts
// ------------------------------------------------------------------------------------------------
template <int NIN, int VEC, int UNROLL, typename F, typename O, typename A, typename B>
__global__ void __launch_bounds__(kMapThreads)
map_rows_kernel(O* __restrict__ out, const A* __restrict__ a, const B* __restrict__ b, RowsParams p, F f) {
  const int64_t c0 = (int64_t)blockIdx.x * (kMapThreads * UNROLL) + threadIdx.x;
  Pack<A, VEC> pa[UNROLL];
  Pack<B, VEC> pb[UNROLL];
  int64_t oo[UNROLL];
  int32_t cnt[UNROLL];  // valid elements of the chunk (0 = chunk out of range)
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    const int64_t c = c0 + (int64_t)u * kMapThreads;
    cnt[u] = 0;
    if (c < p.total_chunks) {
      int64_t row = 0, col = c;
      int64_t off[3] = {0, 0, 0};
      if (p.nouter > 0) {
        if (!p.use64) { row = p.cpr_div.div((uint32_t)c); col = c - row * p.cpr; }
        else { row = c / p.cpr; col = c - row * p.cpr; }
        walk_outer<3>(row, p.nouter, p.use64, p.outer_shape, p.outer_div, p.outer_stride, off);
      }
      const int64_t e = col * VEC;
      const int64_t left = p.inner - e;
      cnt[u] = left >= VEC ? VEC : (int32_t)left;
      oo[u] = off[0] + e;
      const A* ap = a + off[1] + e * p.inner_stride[1];
      if (p.inner_stride[1] == 0) {
        A s = load_one(ap);
#pragma unroll
        for (int k = 0; k < VEC; ++k) pa[u].v[k] = s;
      } else if (cnt[u] == VEC) {
        load_pack<A, VEC>(pa[u], ap);
      } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k)
          if (k < cnt[u]) pa[u].v[k] = ap[k];
      }
      if constexpr (NIN == 2) {
        const B* bp = b + off[2] + e * p.inner_stride[2];
        if (p.inner_stride[2] == 0) {
          B s = load_one(bp);
#pragma unroll
          for (int k = 0; k < VEC; ++k) pb[u].v[k] = s;
        } else if (cnt[u] == VEC) {
          load_pack<B, VEC>(pb[u], bp);
        } else {
#pragma unroll
          for (int k = 0; k < VEC; ++k)
            if (k < cnt[u]) pb[u].v[k] = bp[k];
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    if (cnt[u] == 0) continue;
    Pack<O, VEC> po;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      if constexpr (NIN == 2) po.v[k] = f(pa[u].v[k], pb[u].v[k]);
      else po.v[k] = f(pa[u].v[k]);
    }
    if (cnt[u] == VEC) store_pack<O, VEC>(out + oo[u], po);
    else {
#pragma unroll
      for (int k = 0; k < VEC; ++k)
        if (k < cnt[u]) out[oo[u] + k] = po.v[k];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// tiled kernel: permuted / strided layouts
// ------------------------------------------------------------------------------------------------
template <int NIN, typename F, typename O, typename A, typename B>
__global__ void __launch_bounds__(kMapThreads)
map_tiled_kernel(O* __restrict__ out, const A* __restrict__ a, const B* __restrict__ b, TileParams p, F f) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  A* sm_a = reinterpret_cast<A*>(smem_raw + p.smem_off[1]);
  B* sm_b = reinterpret_cast<B*>(smem_raw + p.smem_off[2]);
  const int tx = threadIdx.x & (kTile - 1);
  const int ty = threadIdx.x >> 6;  // 0..3
  constexpr int kPasses = kTile / (kMapThreads / kTile);  // 16

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
    int64_t off[3] = {0, 0, 0};
    walk_outer<3>(batch, p.nbatch, p.use64, p.batch_shape, p.batch_div, p.batch_stride, off);
    const int64_t a0 = ta * kTile, b0 = tb * kTile;
    const int na = (int)((p.A - a0) < kTile ? (p.A - a0) : kTile);
    const int nb = (int)((p.B - b0) < kTile ? (p.B - b0) : kTile);

    // phase 1: stage permuted inputs (read along b, coalesced)
    if (p.transposed[1]) {
      A r[kPasses];
      const A* src = a + off[1] + a0 * p.sa[1] + (b0 + tx) * p.sb[1];
#pragma unroll
      for (int k = 0; k < kPasses; ++k) {
        const int al = ty + 4 * k;
        if (tx < nb && al < na) r[k] = load_one(src + (int64_t)al * p.sa[1]);
      }
#pragma unroll
      for (int k = 0; k < kPasses; ++k) {
        const int al = ty + 4 * k;
        if (tx < nb && al < na) sm_a[al * kTilePitch + tx] = r[k];
      }
    }
    if constexpr (NIN == 2) if (p.transposed[2]) {
      B r[kPasses];
      const B* src = b + off[2] + a0 * p.sa[2] + (b0 + tx) * p.sb[2];
#pragma unroll
      for (int k = 0; k < kPasses; ++k) {
        const int al = ty + 4 * k;
        if (tx < nb && al < na) r[k] = load_one(src + (int64_t)al * p.sa[2]);
      }
#pragma unroll
      for (int k = 0; k < kPasses; ++k) {
        const int al = ty + 4 * k;
        if (tx < nb && al < na) sm_b[al * kTilePitch + tx] = r[k];
      }
    }
    __syncthreads();

    // phase 2: compute and write along a (coalesced)
    {
      A ra[kPasses];
      B rb[kPasses];
      const A* adir = a + off[1] + (a0 + tx) * p.sa[1] + b0 * p.sb[1];
      const B* bdir = b + off[2] + (a0 + tx) * p.sa[2] + b0 * p.sb[2];
#pragma unroll
      for (int k = 0; k < kPasses; ++k) {
        const int bl = ty + 4 * k;
        if (tx < na && bl < nb) {
          ra[k] = p.transposed[1] ? sm_a[tx * kTilePitch + bl] : load_one(adir + (int64_t)bl * p.sb[1]);
          if constexpr (NIN == 2)
            rb[k] = p.transposed[2] ? sm_b[tx * kTilePitch + bl] : load_one(bdir + (int64_t)bl * p.sb[2]);
        }
      }
      O* dst = out + off[0] + (a0 + tx) * p.sa[0] + b0 * p.sb[0];
#pragma unroll
      for (int k = 0; k < kPasses; ++k) {
        const int bl = ty + 4 * k;
        if (tx < na && bl < nb) {
          O v;
          if constexpr (NIN == 2) v = f(ra[k], rb[k]);
          else v = f(ra[k]);
          dst[(int64_t)bl * p.sb[0]] = v;
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------------
template <typename O, typename A, typename B>
constexpr int map_vec_width() {
  int mn = sizeof(O) < sizeof(A) ? sizeof(O) : sizeof(A);
  mn = mn < (int)sizeof(B) ? mn : (int)sizeof(B);
  int mx = sizeof(O) > sizeof(A) ? sizeof(O) : sizeof(A);
  mx = mx > (int)sizeof(B) ? mx : (int)sizeof(B);
  int v = 16 / mn;
  while (v * mx > 64) v /= 2;
  return v;
}

inline bool fits_u32(int64_t v) { return v >= 0 && v < (int64_t(1) << 31); }

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
    RowsParams p;
    memset(&p, 0, sizeof(p));
    const int nd = c.ndim;
    p.inner = nd ? c.shape[nd - 1] : 1;
    p.nouter = nd ? nd - 1 : 0;
    bool big = false;
    for (int o = 0; o <= NIN; ++o) p.inner_stride[o] = nd ? (int32_t)c.strides[o][nd - 1] : 1;
    for (int i = 0; i < p.nouter; ++i) {
      int d = nd - 2 - i;
      if (!fits_u32(c.shape[d])) big = true;
      p.outer_shape[i] = (uint32_t)c.shape[d];
      p.outer_div[i] = FastDiv((uint32_t)c.shape[d]);
      for (int o = 0; o <= NIN; ++o) p.outer_stride[o][i] = c.strides[o][d];
    }
    // vector path: every unit-stride operand must keep 16 B (or pack-size) alignment on every row
    bool vec_ok = VEC > 1;
    if (vec_ok) {
      for (int o = 0; o <= NIN && vec_ok; ++o) {
        if (p.inner_stride[o] == 0) continue;
        size_t align = esz[o] * VEC > 16 ? 16 : esz[o] * VEC;
        if (reinterpret_cast<uintptr_t>(plan.ptr[o]) % align) vec_ok = false;
        for (int i = 0; i < p.nouter; ++i)
          if ((uint64_t)(std::llabs(p.outer_stride[o][i]) * (int64_t)esz[o]) % align) vec_ok = false;
      }
      if (p.nouter > 0 && p.inner % VEC) vec_ok = false;
    }
    const int vec = vec_ok ? VEC : 1;
    p.cpr = (p.inner + vec - 1) / vec;
    int64_t rows = 1;
    for (int i = 0; i < p.nouter; ++i) rows *= c.shape[nd - 2 - i];
    p.total_chunks = rows * p.cpr;
    if (!fits_u32(p.cpr) || p.total_chunks >= (int64_t(1) << 32)) big = true;
    p.use64 = big ? 1 : 0;
    p.cpr_div = FastDiv(big ? 1u : (uint32_t)p.cpr);
    constexpr int UNROLL = VEC >= 16 ? 1 : VEC >= 8 ? 2 : 4;  // ~16 elements per thread
    int64_t blocks = (p.total_chunks + kMapThreads * UNROLL - 1) / (kMapThreads * UNROLL);
    if (blocks > 0x7fffffffLL) return fail(HPTB_ERR_UNSUPPORTED, "elementwise: tensor too large for one launch");
    if (vec_ok)
      map_rows_kernel<NIN, VEC, UNROLL, F, O, A, B><<<(unsigned)blocks, kMapThreads, 0, stream>>>(out, a, b, p, f);
    else
      map_rows_kernel<NIN, 1, UNROLL, F, O, A, B><<<(unsigned)blocks, kMapThreads, 0, stream>>>(out, a, b, p, f);
    HPTB_CUDA_CHECK(cudaGetLastError());
    return HPTB_OK;
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
  if (db < 0) {  // no permuted unit-stride dim: take the innermost remaining dim (largest extent wins ties)
    for (int d = nd - 1; d >= 0; --d)
      if (d != da) { db = d; break; }
  }
  TileParams p;
  memset(&p, 0, sizeof(p));
  p.A = c.shape[da];
  p.B = db >= 0 ? c.shape[db] : 1;
  size_t smem = 0;
  for (int o = 0; o <= NIN; ++o) {
    p.sa[o] = c.strides[o][da];
    p.sb[o] = db >= 0 ? c.strides[o][db] : 0;
    p.transposed[o] = (o > 0 && db >= 0 && p.sb[o] == 1 && p.sa[o] != 1 && p.sa[o] != 0) ? 1 : 0;
    if (p.transposed[o]) {
      smem = (smem + 15) / 16 * 16;
      p.smem_off[o] = (int32_t)smem;
      smem += (size_t)kTile * kTilePitch * esz[o];
    }
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
  p.tiles_a = (p.A + kTile - 1) / kTile;
  p.tiles_b = (p.B + kTile - 1) / kTile;
  int64_t batch = 1;
  for (int i = 0; i < nb; ++i) batch *= p.batch_shape[i];
  p.ntiles = p.tiles_a * p.tiles_b * batch;
  if (p.ntiles >= (int64_t(1) << 32) || !fits_u32(p.tiles_a) || !fits_u32(p.tiles_b)) big = true;
  p.use64 = big ? 1 : 0;
  p.tiles_a_div = FastDiv(big ? 1u : (uint32_t)p.tiles_a);
  p.tiles_b_div = FastDiv(big ? 1u : (uint32_t)p.tiles_b);
  int64_t blocks = p.ntiles < 0x7fffffffLL ? p.ntiles : 0x7fffffffLL;
  auto kern = map_tiled_kernel<NIN, F, O, A, B>;
  if (smem > 48 * 1024) {
    static bool opted_in = false;  // per instantiation
    if (!opted_in) {
      HPTB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kTile * kTilePitch * 8));
      opted_in = true;
    }
  }
  kern<<<(unsigned)blocks, kMapThreads, smem, stream>>>(out, a, b, p, f);
  HPTB_CUDA_CHECK(cudaGetLastError());
  return HPTB_OK;
}

}  // namespace hptb
