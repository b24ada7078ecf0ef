// softmax_band.cuh — softmax over a band that stays RESIDENT in the shared memory of a thread-block cluster.
// (Included by softmax.cu inside its anonymous namespace.)
//
// The online kernels (softmax_rows_stream_vec, softmax_cols_tiled) read the input twice and carry a running
// (max, Σ) pair whose every move of the maximum rescales Σ inexactly; the round-1 survey had them at 0.33–0.43 of
// peak.  For the shapes that matter the band a softmax needs at once is small: a ROW of ≤ 131072 f32, or — softmax
// over a strided axis — a 128-byte-wide COLUMN BAND of ≤ 6144 rows.  A cluster of up to 8 CTAs holds such a band in
// its shared memory (≤ 64 KB per CTA where that is enough: three CTAs per SM; at most 96 KB), so the input is read from HBM ONCE, the maximum is known exactly before the
// first exp (no online rescale at all: exp(x − max) exactly as the register kernel computes it), and the two
// statistics cross the cluster through distributed shared memory: every CTA PUSHES its partial into a slot of every
// peer's array, one cluster barrier, then everybody reads its own copy (nobody reads a peer that may have exited).
// One launch; ranks are combined in rank order → deterministic.
// (<cooperative_groups.h> is included by softmax.cu at global scope)
namespace cg = cooperative_groups;

constexpr int kBandMaxCl = 8;  // the portable cluster size; 16 (non-portable) measured no faster for rows and slower for column bands

// 16 bytes global → shared without a register stop (L2 only: the band is read once)
__device__ __forceinline__ void band_cp16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void band_cp_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// A CTA may write a peer's shared memory only once that peer is known to have STARTED (compute-sanitizer: "block that
// might not have entered yet").  Split barrier: every CTA arrives as its first instruction and waits right before its
// first remote store — by then the band has been copied and reduced, so the wait is free.
__device__ __forceinline__ void band_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void band_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

struct BandParams {
  DimWalk kept;            // rows: every kept dim; cols: the kept dims other than the contiguous one (stride_a in, stride_b out)
  int64_t L;               // axis length (elements)
  int64_t per_cta;         // rows: 16-byte packs per CTA; cols: axis positions per CTA
  int64_t sa_in, sa_out;   // cols: element strides of the axis
  int64_t C, ctiles;       // cols: extent of the contiguous kept dim, column tiles per outer index
  int32_t cl, log, use64, pad;
};

// ---- rows: the axis has unit stride; one cluster per row, CTA r owns packs [r·per_cta, (r+1)·per_cta) ---------------
template <typename T, int VEC>
__global__ void __launch_bounds__(kSmThreads) softmax_band_rows(const T* __restrict__ in, T* __restrict__ out, BandParams p) {
  pdl_prologue();
  typedef float C;
  extern __shared__ __align__(16) unsigned char band_raw[];
  T* tile = reinterpret_cast<T*>(band_raw);
  __shared__ C s_buf[kSmThreads / 32];
  __shared__ C s_max[kBandMaxCl], s_sum[kBandMaxCl];
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x;
  const unsigned rank = p.cl > 1 ? cluster.block_rank() : 0;
  if (p.cl > 1) band_cluster_arrive();
  const int64_t row = (int64_t)blockIdx.x / p.cl;
  int64_t in_off = 0, out_off = 0;
  walk2(row, p.kept, p.use64, in_off, out_off);
  const int64_t packs = p.L / VEC, c0 = (int64_t)rank * p.per_cta;
  int64_t nmy = packs - c0;
  nmy = nmy < 0 ? 0 : (nmy > p.per_cta ? p.per_cta : nmy);
  const int n = (int)nmy;  // ≤ 96 KB / 16
  const T* src = in + in_off + c0 * VEC;
  T* dst = out + out_off + c0 * VEC;
  const C neg_inf = Limits<C>::lowest();
  // pass A: HBM → shared memory with cp.async — ALL of a thread's 16-byte copies are in flight at once (up to 24; through
  // registers, four at a time, the pass was a chain of DRAM round trips: f32 [256,131072] 73 µs), then the maximum
  for (int c = tid; c < n; c += kSmThreads) band_cp16(tile + (size_t)c * VEC, src + (int64_t)c * VEC);
  band_cp_wait();
  C mx = neg_inf;
  for (int c = tid; c < n; c += kSmThreads) {
    const Pack<T, VEC> v = *reinterpret_cast<const Pack<T, VEC>*>(tile + (size_t)c * VEC);
#pragma unroll
    for (int k = 0; k < VEC; ++k) mx = sm_max<C>(mx, to_compute<T>(v.v[k]));
  }
  mx = group_reduce<MaxOp, C, kSmThreads>(mx, s_buf, neg_inf);
  if (p.cl > 1) {
    band_cluster_wait();  // every peer is running
    if (tid < p.cl) *cluster.map_shared_rank(&s_max[rank], tid) = mx;
    cluster.sync();
    mx = s_max[0];
    for (int r = 1; r < p.cl; ++r) mx = sm_max<C>(mx, s_max[r]);
  }
  // pass B: Σ exp(x − max) from shared memory; each thread re-reads the packs it wrote, so no barrier is needed.  The
  // formula is the register kernel's (and the reference CPU kernel's): same error class by construction.  f32 softmax
  // keeps the exponentials in place of the inputs, pass C is then one multiply per element.
  constexpr bool kInPlace = std::is_same<T, float>::value;
  C sum = (C)0;
  for (int c = tid; c < n; c += kSmThreads) {
    Pack<T, VEC>* slot = reinterpret_cast<Pack<T, VEC>*>(tile + (size_t)c * VEC);
    Pack<T, VEC> v = *slot;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const C e = sm_exp_o<T>(to_compute<T>(v.v[k]) - mx);
      sum += e;
      if constexpr (kInPlace) v.v[k] = e;
    }
    if constexpr (kInPlace) {
      if (!p.log) *slot = v;
    }
  }
  sum = group_reduce<AddOp, C, kSmThreads>(sum, s_buf, (C)0);
  if (p.cl > 1) {
    if (tid < p.cl) *cluster.map_shared_rank(&s_sum[rank], tid) = sum;
    cluster.sync();
    sum = s_sum[0];
    for (int r = 1; r < p.cl; ++r) sum += s_sum[r];  // rank order: the same total on every CTA
  }
  const C lg = sm_log<C>(sum), inv = (C)1 / sum;
  // pass C: shared memory → HBM
  for (int c = tid; c < n; c += kSmThreads) {
    const Pack<T, VEC> v = *reinterpret_cast<const Pack<T, VEC>*>(tile + (size_t)c * VEC);
    Pack<T, VEC> o;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      if (kInPlace && !p.log) {
        o.v[k] = from_compute<T>(to_compute<T>(v.v[k]) * inv);
      } else {
        const C sh = to_compute<T>(v.v[k]) - mx;
        o.v[k] = from_compute<T>(p.log ? sh - lg : sm_exp_o<T>(sh) * inv);
      }
    }
    store_pack<T, VEC>(dst + (int64_t)c * VEC, o);
  }
}

// ---- cols: the axis is strided, another dim is contiguous; a cluster owns a 128-byte-wide column band, CTA r the axis
// positions [r·per_cta, (r+1)·per_cta).  8 lanes × one 16-byte pack span the band's width, 32 thread rows walk the axis.
template <typename T, int VEC>
__global__ void __launch_bounds__(kSmThreads) softmax_band_cols(const T* __restrict__ in, T* __restrict__ out, BandParams p) {
  pdl_prologue();
  typedef float C;
  constexpr int TX = 8, TY = kSmThreads / TX, W = TX * VEC, NW = kSmThreads / 32;
  extern __shared__ __align__(16) unsigned char band_raw[];
  T* tile = reinterpret_cast<T*>(band_raw);  // [per_cta][W]
  __shared__ C s_col[NW][W];
  __shared__ C s_xm[kBandMaxCl][W], s_xs[kBandMaxCl][W];
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, lane = tid % TX, ty = tid / TX, warp = tid >> 5;
  const unsigned rank = p.cl > 1 ? cluster.block_rank() : 0;
  if (p.cl > 1) band_cluster_arrive();
  const int64_t band = (int64_t)blockIdx.x / p.cl;
  const int64_t outer = band / p.ctiles, ct = band - outer * p.ctiles;
  const int64_t col0 = ct * W + (int64_t)lane * VEC;
  const bool active = col0 < p.C;  // C is a multiple of VEC
  int64_t in_off = 0, out_off = 0;
  if (p.kept.n > 0) walk2(outer, p.kept, p.use64, in_off, out_off);
  const int64_t e0 = (int64_t)rank * p.per_cta;
  int64_t nmy = p.L - e0;
  nmy = nmy < 0 ? 0 : (nmy > p.per_cta ? p.per_cta : nmy);
  const int n = (int)nmy;
  const T* src = in + in_off + e0 * p.sa_in + col0;
  T* dst = out + out_off + e0 * p.sa_out + col0;
  const C neg_inf = Limits<C>::lowest();
  // column statistics of this CTA → every CTA of the cluster (thread rows → warps by shuffle, warps through s_col)
  auto across = [&](C (&v)[VEC], bool is_max, C (*xslot)[W]) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
#pragma unroll
      for (int off = 8; off <= 16; off <<= 1) {
        const C o = __shfl_xor_sync(0xffffffffu, v[k], off);
        v[k] = is_max ? sm_max<C>(v[k], o) : v[k] + o;
      }
    }
    __syncthreads();  // s_col reuse
    if ((tid & 31) < TX) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) s_col[warp][lane * VEC + k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      C r = s_col[0][lane * VEC + k];
      for (int w = 1; w < NW; ++w) r = is_max ? sm_max<C>(r, s_col[w][lane * VEC + k]) : r + s_col[w][lane * VEC + k];
      v[k] = r;
    }
    if (p.cl > 1) {
      if (ty < p.cl) {  // thread row ty pushes this CTA's W values to rank ty
        C* peer = cluster.map_shared_rank(&xslot[rank][0], ty);
#pragma unroll
        for (int k = 0; k < VEC; ++k) peer[lane * VEC + k] = v[k];
      }
      cluster.sync();
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        C r = xslot[0][lane * VEC + k];
        for (int q = 1; q < p.cl; ++q) r = is_max ? sm_max<C>(r, xslot[q][lane * VEC + k]) : r + xslot[q][lane * VEC + k];
        v[k] = r;
      }
    }
  };
  C mx[VEC], sum[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) { mx[k] = neg_inf; sum[k] = (C)0; }
  if (active) {
    for (int e = ty; e < n; e += TY) band_cp16(tile + (size_t)e * W + lane * VEC, src + (int64_t)e * p.sa_in);
    band_cp_wait();
    for (int e = ty; e < n; e += TY) {
      const Pack<T, VEC> v = *reinterpret_cast<const Pack<T, VEC>*>(tile + (size_t)e * W + lane * VEC);
#pragma unroll
      for (int k = 0; k < VEC; ++k) mx[k] = sm_max<C>(mx[k], to_compute<T>(v.v[k]));
    }
  }
  if (p.cl > 1) band_cluster_wait();  // every peer is running: remote stores are allowed from here on
  across(mx, true, s_xm);
  constexpr bool kInPlace = std::is_same<T, float>::value;
  if (active) {
    for (int e = ty; e < n; e += TY) {
      Pack<T, VEC>* slot = reinterpret_cast<Pack<T, VEC>*>(tile + (size_t)e * W + lane * VEC);
      Pack<T, VEC> v = *slot;
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const C t = sm_exp_o<T>(to_compute<T>(v.v[k]) - mx[k]);
        sum[k] += t;
        if constexpr (kInPlace) v.v[k] = t;
      }
      if constexpr (kInPlace) {
        if (!p.log) *slot = v;
      }
    }
  }
  across(sum, false, s_xs);
  if (!active) return;
  C lg[VEC], inv[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) { lg[k] = sm_log<C>(sum[k]); inv[k] = (C)1 / sum[k]; }
  for (int e = ty; e < n; e += TY) {
    const Pack<T, VEC> v = *reinterpret_cast<const Pack<T, VEC>*>(tile + (size_t)e * W + lane * VEC);
    Pack<T, VEC> o;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      if (kInPlace && !p.log) {
        o.v[k] = from_compute<T>(to_compute<T>(v.v[k]) * inv[k]);
      } else {
        const C sh = to_compute<T>(v.v[k]) - mx[k];
        o.v[k] = from_compute<T>(p.log ? sh - lg[k] : sm_exp_o<T>(sh) * inv[k]);
      }
    }
    store_pack<T, VEC>(dst + (int64_t)e * p.sa_out, o);
  }
}

// Tried and dropped (profiles/r02r_band_pipe.txt): a persistent variant — ONE 1024-thread CTA per SM, two shared-memory
// buffers, the cp.async copies of band i+1 in flight while band i goes through max / exp / store, grid sized by
// cudaOccupancyMaxActiveClusters (a cluster of 8 lives inside one GPC: 16 clusters are co-resident, not 18).  f32
// [4096,8192] axis 0: 123.5 µs against 78.7 µs for the kernel above (181 µs with 18 clusters launched = two waves);
// [2048,16384]: 110.9 vs 70.4 µs.  One CTA per SM cannot hide its own barriers (two cluster syncs + six CTA barriers per
// band), whatever is in flight behind them.
