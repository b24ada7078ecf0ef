// tma_tile.cuh — transposing elementwise kernel with TMA-staged tiles (1-, 2- and 4-byte element types).
//
// The case: a permuted operand whose unit-stride dim (b) is not the output's (a) — BASELINE config 2 `V = X.t();
// V.sin()`.  map_tiled_smem_kernel (elementwise.cuh) stages such tiles with per-thread 16-byte loads and an
// element-sized STS scatter: 15 of its 35 issued instructions per element are addressing and tile traffic, and the
// kernel is issue-bound for anything heavier than exp (ncu, profiles/r01final_ncu_full_key_metrics.csv).  Here the
// TMA engine does the addressing: the operand is described once per launch by a tensor map (cuTensorMapEncodeTiled,
// dims b, a, ≤ 3 batch dims) and each thread's tile work is ONE ldmatrix per 16-byte output pack.
//
// Layout trick.  TMA deposits a box with b contiguous (128-byte rows, SWIZZLE_128B); an output pack needs E = 16 /
// sizeof(T) CONSECUTIVE a at one b, i.e. E different rows.  ldmatrix hands thread `lane` 4 bytes of row lane/4 of
// each of its four 8×16-byte matrices, so matrix j must hold rows a ≡ j (mod 4) (4-byte types; for 2-byte types
// ldmatrix.trans and a ≡ 2j, 2j+1 (mod 8)).  With a plain box those rows would all sit at a ≡ const (mod 4), two
// swizzle phases, 4-way bank conflicts.  So the tile is loaded as E boxes with a TRAVERSAL STRIDE of E along a
// (`elementStrides`): box r gathers rows a0 + r, a0 + r + E, … into its own region, where CONSECUTIVE region rows —
// eight distinct swizzle phases — are what one ldmatrix matrix reads: conflict-free, and the global side is still
// full 128-byte lines.  (2-byte types: a matrix takes 4 rows from each of two regions; the odd regions sit 512
// bytes off the 1 KB swizzle period, which shifts their phase by 4.)
//
// A CTA owns 64 (a) × kTmaSub = 4 sub-tiles of 128 bytes (b): all sub-tile loads (32 KB in flight per CTA, four CTAs per
// SM) are issued before the first is consumed; one single-use mbarrier per sub-tile, no ring, no producer warp.
// Out-of-range parts of edge tiles are zero-filled by TMA and masked at the store.
#pragma once
#include <cuda.h>  // CUtensorMap and its enums only — the encoder is resolved at run time (no libcuda link dependency)

#include "common.h"
#include "launch.cuh"
#include "scalar.cuh"

namespace hptb {

constexpr int kSmemModeStaged = 1, kSmemModeDirect = 2, kSmemModeScalar = 3;  // how an operand reaches a tile (elementwise.cuh)
// sub-tiles per CTA.  4 (32 KB of tiles, 4 CTAs per SM, 8192 CTAs for config 2) beat 8 (64 KB, 3 CTAs per SM) on every
// case of tools/tma_compare.py (profiles/r02e_tma_compare{,_sub4}.txt): f32 a.t() + b 155 → 121 µs, f16 85 → 64 µs,
// f32 sin 87.3 → 85.5 µs — shorter-lived CTAs, a shorter last wave, and the partner prefetch covers half the CTA.
#ifndef HPTB_TMA_SUB
#define HPTB_TMA_SUB 4
#endif
constexpr int kTmaSub = HPTB_TMA_SUB;
constexpr int kTmaMinCtas = kTmaSub >= 8 ? 3 : 4;
constexpr int kTmaTA = 64;           // tile extent along a
constexpr int kTmaThreads = 256;
template <int ESZ> struct TmaGeom;
template <> struct TmaGeom<4> {
  static constexpr int E = 4, BW = 32;             // elements per pack / per 128-byte sub-tile row
  static constexpr int kBoxes = 4, kBoxBytes = 2048, kSubStride = 8192;
  static __host__ __device__ constexpr int region_off(int r) { return r * 2048; }
};
template <> struct TmaGeom<2> {
  static constexpr int E = 8, BW = 64;
  static constexpr int kBoxes = 8, kBoxBytes = 1024, kSubStride = 9216;
  // even regions back to back from 0, odd regions from 4 KB + 512 B: half a swizzle period off
  static __host__ __device__ constexpr int region_off(int r) { return (r & 1) ? 4608 + (r >> 1) * 1024 : (r >> 1) * 1024; }
};

struct TmaTileParams {
  int64_t A, B;
  int64_t out_sb;                       // output stride along b (elements; 1 along a)
  int64_t in1_sb;                       // second operand (NIN == 2), read where the output is written
  int32_t in1_mode;                     // kSmemModeDirect (unit stride along a) or kSmemModeScalar
  int32_t swap;                         // 1: f(other, staged) instead of f(staged, other)
  int32_t nbatch;
  uint32_t batch_shape[3];              // innermost batch dim first (tensor-map dims 2..4)
  int64_t batch_out[3], batch_in1[3];   // element strides
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NIN, typename F, typename T>
__global__ void __launch_bounds__(kTmaThreads, kTmaMinCtas)
map_tma_tile_kernel(T* __restrict__ out, const T* __restrict__ in1, const __grid_constant__ CUtensorMap tmap, TmaTileParams p, F f) {
  // 1-byte types ride the 2-byte geometry: the tile is loaded and ldmatrix-transposed as u16 PAIRS along b, and the two
  // bytes of every pair — rows 2·b2 and 2·b2 + 1 of the output — are pulled apart with four byte permutes per thread
  constexpr bool kBytes = sizeof(T) == 1;
  typedef TmaGeom<(kBytes ? 2 : sizeof(T))> G;
  constexpr int E = G::E;  // elements per 16-byte pack of the transposed unit (8 for the byte variant: two 8-byte rows)
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTmaSub * G::kSubStride);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t a0 = (int64_t)blockIdx.x * kTmaTA;
  constexpr int BWE = G::BW * (kBytes ? 2 : 1);  // b-elements per sub-tile
  const int64_t b0 = (int64_t)blockIdx.y * (kTmaSub * BWE);
  int nsub = (int)((p.B - b0 + BWE - 1) / BWE);
  if (nsub > kTmaSub) nsub = kTmaSub;
  // batch coordinates (once per CTA)
  int32_t bc3[3] = {0, 0, 0};
  int64_t off_out = 0, off_in1 = 0;
  {
    uint32_t rest = blockIdx.z;
#pragma unroll
    for (int i = 0; i < 3; ++i) {  // unused batch dims have extent 1
      const uint32_t q = rest / p.batch_shape[i];
      const uint32_t c = rest - q * p.batch_shape[i];
      bc3[i] = (int32_t)c;
      off_out += (int64_t)c * p.batch_out[i];
      off_in1 += (int64_t)c * p.batch_in1[i];
      rest = q;
    }
  }
  if (tid == 0) {
    for (int s = 0; s < kTmaSub; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bars + s)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_prologue();  // the operand may be the previous kernel's output: no global read before this point
  if (tid < nsub)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bars + tid)), "r"(G::kBoxes * G::kBoxBytes) : "memory");
  if (tid < nsub * G::kBoxes) {
    const int s = tid / G::kBoxes, r = tid % G::kBoxes;
    const uint32_t dst = smem_addr(smem + s * G::kSubStride + G::region_off(r));
    const int32_t c0 = (int32_t)((b0 + (int64_t)s * BWE) / (kBytes ? 2 : 1)), c1 = (int32_t)(a0 + r);
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
        "l"(&tmap), "r"(c0), "r"(c1), "r"(bc3[0]), "r"(bc3[1]), "r"(bc3[2]), "r"(smem_addr(bars + s))
        : "memory");
  }
  T* dst_base = out + off_out + a0;
  // this thread's two units of every sub-tile: 32 a × one 16-byte chunk of b each (ah = unit, chunk = warp)
  int a_loc[2], b_loc[2];
  uint32_t ld_off[2];  // ldmatrix row address of this lane within a sub-tile
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int ah = i, bc = warp;
    if constexpr (sizeof(T) == 4) {
      const int j = lane >> 3, row = lane & 7;
      ld_off[i] = G::region_off(j) + (ah * 8 + row) * 128 + ((bc ^ row) << 4);
      a_loc[i] = ah * 32 + (lane >> 2) * 4;
      b_loc[i] = bc * 4 + (lane & 3);
    } else if constexpr (kBytes) {
      const int j = lane >> 3, row = lane & 7, t = row >> 1, reg = 2 * j + (row & 1), rl = ah * 4 + t;
      ld_off[i] = G::region_off(reg) + rl * 128 + ((bc ^ ((rl + 4 * (reg & 1)) & 7)) << 4);
      a_loc[i] = ah * 32 + (lane & 3) * 8;
      b_loc[i] = 2 * (bc * 8 + (lane >> 2));  // the EVEN row of the pair
    } else {
      const int j = lane >> 3, row = lane & 7, t = row >> 1, reg = 2 * j + (row & 1), rl = ah * 4 + t;
      ld_off[i] = G::region_off(reg) + rl * 128 + ((bc ^ ((rl + 4 * (reg & 1)) & 7)) << 4);
      a_loc[i] = ah * 32 + (lane & 3) * 8;
      b_loc[i] = bc * 8 + (lane >> 2);
    }
  }
  auto wait_sub = [&](int s) {  // phase 0 of a single-use barrier
    const uint32_t bar = smem_addr(bars + s);
    uint32_t done;
    do {
      asm volatile(
          "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
          : "=r"(done)
          : "r"(bar)
          : "memory");
    } while (!done);
  };
  // 16 bytes of the transposed unit: E consecutive a at one b — for 1-byte types 8 consecutive a of the EVEN row
  // (bytes 0..7) followed by the same 8 a of the ODD row (bytes 8..15)
  auto load_unit = [&](int s, int i, Pack<T, (kBytes ? 16 : E)>& x) {
    uint32_t r[4];
    const uint32_t addr = smem_addr(smem + s * G::kSubStride) + ld_off[i];
    if constexpr (sizeof(T) == 4)
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
    else
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
    if constexpr (kBytes) {
      const uint32_t q[4] = {__byte_perm(r[0], r[1], 0x6420), __byte_perm(r[2], r[3], 0x6420), __byte_perm(r[0], r[1], 0x7531),
                             __byte_perm(r[2], r[3], 0x7531)};
      memcpy(&x, q, 16);
    } else {
      memcpy(&x, r, 16);
    }
  };
  if constexpr (NIN == 1) {
    for (int s = 0; s < nsub; ++s) {
      wait_sub(s);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        Pack<T, (kBytes ? 16 : E)> x, y;
        load_unit(s, i, x);
        const int64_t bg = b0 + (int64_t)s * BWE + b_loc[i];
        if (bg >= p.B || a0 + a_loc[i] >= p.A) continue;  // A is a multiple of E (B of 2): packs are all-in or all-out
        apply_pack<F, T, T, (kBytes ? 16 : E)>(f, y, x);
        if constexpr (kBytes) {
          Pack<T, 8> lo, hi;
          memcpy(&lo, &y, 8);
          memcpy(&hi, reinterpret_cast<const char*>(&y) + 8, 8);
          store_pack<T, 8>(dst_base + bg * p.out_sb + a_loc[i], lo);
          store_pack<T, 8>(dst_base + (bg + 1) * p.out_sb + a_loc[i], hi);
        } else {
          store_pack<T, E>(dst_base + bg * p.out_sb + a_loc[i], y);
        }
      }
    }
  } else {
    // The partner operand is read where the output is written (unit stride along a).  Its loads run kPre sub-tiles
    // ahead of the tile that consumes them: issued per sub-tile after the mbarrier wait they formed a chain of
    // kTmaSub dependent DRAM round trips per CTA (f32 a.t() + b: 0.67 of peak, behind the shared-memory kernel's 0.75).
    constexpr int kPre = 2;
    const bool direct = p.in1_mode == kSmemModeDirect;
    const T* in1_base = in1 + off_in1 + (direct ? a0 : 0);
    T scalar1{};
    if (!direct) scalar1 = load_one(in1 + off_in1);
    constexpr int PE = kBytes ? 16 : E;
    Pack<T, PE> ring[kPre + 1][2];
    auto fetch = [&](int s, Pack<T, PE> (&dst)[2]) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int64_t bg = b0 + (int64_t)s * BWE + b_loc[i];
        if (direct && s < nsub && bg < p.B && a0 + a_loc[i] < p.A) {
          if constexpr (kBytes) {  // the even and the odd row of the pair, 8 bytes each
            Pack<T, 8> lo, hi;
            load_pack<T, 8>(lo, in1_base + bg * p.in1_sb + a_loc[i]);
            load_pack<T, 8>(hi, in1_base + (bg + 1) * p.in1_sb + a_loc[i]);
            memcpy(&dst[i], &lo, 8);
            memcpy(reinterpret_cast<char*>(&dst[i]) + 8, &hi, 8);
          } else {
            load_pack<T, E>(dst[i], in1_base + bg * p.in1_sb + a_loc[i]);
          }
        }
      }
    };
#pragma unroll
    for (int d = 0; d < kPre; ++d) fetch(d, ring[d]);
#pragma unroll
    for (int s = 0; s < kTmaSub; ++s) {
      if (s >= nsub) break;
      if (s + kPre < kTmaSub) fetch(s + kPre, ring[(s + kPre) % (kPre + 1)]);
      wait_sub(s);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        Pack<T, PE> x, y;
        load_unit(s, i, x);
        const int64_t bg = b0 + (int64_t)s * BWE + b_loc[i];
        if (bg >= p.B || a0 + a_loc[i] >= p.A) continue;
        Pack<T, PE> o = ring[s % (kPre + 1)][i];
        if (!direct) {
#pragma unroll
          for (int k = 0; k < PE; ++k) o.v[k] = scalar1;
        }
        if (p.swap) {
#pragma unroll
          for (int k = 0; k < PE; ++k) y.v[k] = f(o.v[k], x.v[k]);
        } else {
#pragma unroll
          for (int k = 0; k < PE; ++k) y.v[k] = f(x.v[k], o.v[k]);
        }
        if constexpr (kBytes) {
          Pack<T, 8> lo, hi;
          memcpy(&lo, &y, 8);
          memcpy(&hi, reinterpret_cast<const char*>(&y) + 8, 8);
          store_pack<T, 8>(dst_base + bg * p.out_sb + a_loc[i], lo);
          store_pack<T, 8>(dst_base + (bg + 1) * p.out_sb + a_loc[i], hi);
        } else {
          store_pack<T, E>(dst_base + bg * p.out_sb + a_loc[i], y);
        }
      }
    }
  }
}

// ---- host side -----------------------------------------------------------------------------------------------------
typedef CUresult (*TmaEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline TmaEncodeTiledFn tma_encoder() {
  static const TmaEncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      p = nullptr;
    }
    return reinterpret_cast<TmaEncodeTiledFn>(p);
  }();
  return fn;
}
inline bool tma_disabled() {
  static const bool off = [] { const char* e = getenv("HPTB_NO_TMA"); return e && e[0] == '1'; }();
  return off;
}

// The staged operand as a rank-5 tensor map: dim 0 = b (unit stride), dim 1 = a, dims 2..4 = batch (extent 1 when
// unused).  Returns false when the layout is outside what a tiled map can describe (negative / unaligned strides, …):
// the caller then takes map_tiled_smem_kernel.
template <typename T>
inline bool tma_make_map(CUtensorMap* m, const T* base, int64_t A, int64_t B, int64_t sa, int nbatch, const uint32_t* bshape,
                         const int64_t* bstride) {
  typedef TmaGeom<(sizeof(T) == 1 ? 2 : sizeof(T))> G;
  TmaEncodeTiledFn enc = tma_encoder();
  if (!enc) return false;
  if (reinterpret_cast<uintptr_t>(base) % 16 || A <= 0 || B <= 0 || A >= (int64_t(1) << 31) || B >= (int64_t(1) << 31)) return false;
  if (sizeof(T) == 1 && (B & 1)) return false;  // 1-byte types travel as pairs along b
  cuuint64_t dims[5] = {(cuuint64_t)(sizeof(T) == 1 ? B / 2 : B), (cuuint64_t)A, 1, 1, 1};
  cuuint64_t strides[4] = {0, 0, 0, 0};  // bytes, dims 1..4
  auto ok_stride = [](int64_t s_bytes) { return s_bytes > 0 && s_bytes % 16 == 0 && s_bytes < (int64_t(1) << 40); };
  if (!ok_stride(sa * (int64_t)sizeof(T))) return false;
  strides[0] = (cuuint64_t)(sa * (int64_t)sizeof(T));
  cuuint64_t last = strides[0];
  for (int i = 0; i < 3; ++i) {
    if (i < nbatch) {
      if (!ok_stride(bstride[i] * (int64_t)sizeof(T))) return false;
      dims[2 + i] = bshape[i];
      strides[1 + i] = (cuuint64_t)(bstride[i] * (int64_t)sizeof(T));
      last = strides[1 + i];
    } else {
      strides[1 + i] = last;  // extent 1: never stepped
    }
  }
  const cuuint32_t box[5] = {(cuuint32_t)G::BW, (cuuint32_t)kTmaTA, 1, 1, 1};
  const cuuint32_t estr[5] = {1, (cuuint32_t)G::E, 1, 1, 1};
  const CUtensorMapDataType dt = sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16;  // 1-byte: u16 pairs
  const CUresult rc = enc(m, dt, 5, const_cast<T*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS;
}

}  // namespace hptb
