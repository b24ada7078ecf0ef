// tools/microbench/stream_reduce.cu — development microbenchmark (not product code): which kernel STRUCTURE streams a
// row-sum of an [M, L] f32 matrix fastest on B200 when the whole job is only 10–40 µs long?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/stream_reduce tools/microbench/stream_reduce.cu
//   tools/microbench/stream_reduce [M L]...
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// K0: one warp per row, UNROLL 16-byte loads per lane per iteration (the product kernel's structure)
template <int UNROLL>
__global__ void __launch_bounds__(256) k_warp_row(const float* __restrict__ x, float* __restrict__ out, int M, int L) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float4* p = reinterpret_cast<const float4*>(x + (size_t)row * L);
  const int n = L / 4;
  float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int c = lane; c < n; c += 32 * UNROLL) {
    float4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) if (c + u * 32 < n) v[u] = ldg_stream(p + c + u * 32);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) if (c + u * 32 < n) { a0 += v[u].x; a1 += v[u].y; a2 += v[u].z; a3 += v[u].w; }
  }
  float s = warp_sum((a0 + a1) + (a2 + a3));
  if (lane == 0) out[row] = s;
}

// K1: persistent warps (grid = SMs × occ), rows strided over all warps
template <int UNROLL>
__global__ void __launch_bounds__(256) k_warp_row_persist(const float* __restrict__ x, float* __restrict__ out, int M, int L) {
  const int lane = threadIdx.x & 31;
  const int W = gridDim.x * 8;
  const int n = L / 4;
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < M; row += W) {
    const float4* p = reinterpret_cast<const float4*>(x + (size_t)row * L);
    float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int c = lane; c < n; c += 32 * UNROLL) {
      float4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) if (c + u * 32 < n) v[u] = ldg_stream(p + c + u * 32);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) if (c + u * 32 < n) { a0 += v[u].x; a1 += v[u].y; a2 += v[u].z; a3 += v[u].w; }
    }
    float s = warp_sum((a0 + a1) + (a2 + a3));
    if (lane == 0) out[row] = s;
  }
}

// K2: CTA per row (256 threads × UNROLL loads issued at once, block reduce)
template <int UNROLL>
__global__ void __launch_bounds__(256) k_cta_row(const float* __restrict__ x, float* __restrict__ out, int M, int L) {
  __shared__ float sp[8];
  const int row = blockIdx.x;
  const float4* p = reinterpret_cast<const float4*>(x + (size_t)row * L);
  const int n = L / 4;
  float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int c = threadIdx.x; c < n; c += 256 * UNROLL) {
    float4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) if (c + u * 256 < n) v[u] = ldg_stream(p + c + u * 256);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) if (c + u * 256 < n) { a0 += v[u].x; a1 += v[u].y; a2 += v[u].z; a3 += v[u].w; }
  }
  float s = warp_sum((a0 + a1) + (a2 + a3));
  if ((threadIdx.x & 31) == 0) sp[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { float t = 0; for (int w = 0; w < 8; ++w) t += sp[w]; out[row] = t; }
}

// ---- K3: persistent CTA per SM, bulk-async (TMA) ring per warp --------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// The CTA owns rows [r0, r1): ONE contiguous span of (r1-r0)·L floats, cut into stages of SB bytes.  Warp w takes
// stages w, w+NW, … through its private ring of NS smem buffers (lane 0 issues the bulk copies).  Per-(row, warp)
// partials go to smem and are combined in warp order at the end → deterministic.
template <int NW, int NS, int SB>
__global__ void __launch_bounds__(NW * 32) k_tma_ring(const float* __restrict__ x, float* __restrict__ out, int M, int L, int rows_per_cta) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* ring = smem;                                            // [NW][NS][SB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)NW * NS * SB);  // [NW][NS]
  float* part = reinterpret_cast<float*>(bars + NW * NS);               // [rows_per_cta][NW]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(M, r0 + rows_per_cta);
  const int nrows = r1 - r0;
  if (nrows <= 0) return;
  for (int i = threadIdx.x; i < rows_per_cta * NW; i += NW * 32) part[i] = 0.f;
  if (lane == 0)
    for (int s = 0; s < NS; ++s) mbar_init(&bars[warp * NS + s], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const size_t row_bytes = (size_t)L * 4;
  const int spr = (int)(row_bytes / SB);                   // stages per row (host guarantees divisibility)
  const long long nst = (long long)nrows * spr;            // stages of this CTA
  const unsigned char* gbase = reinterpret_cast<const unsigned char*>(x + (size_t)r0 * L);
  unsigned char* myring = ring + (size_t)warp * NS * SB;
  uint64_t* mybar = bars + warp * NS;
  // my stages: t = warp + k·NW, k = 0..
  const long long mine = nst > warp ? (nst - warp + NW - 1) / NW : 0;
  if (lane == 0) {
    for (int k = 0; k < NS && k < mine; ++k) {
      mbar_expect_tx(&mybar[k], SB);
      bulk_g2s(myring + (size_t)k * SB, gbase + (size_t)(warp + (long long)k * NW) * SB, SB, &mybar[k]);
    }
  }
  float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  int cur_row = -1;
  for (long long k = 0; k < mine; ++k) {
    const int slot = (int)(k % NS);
    const uint32_t parity = (uint32_t)((k / NS) & 1);
    const long long t = warp + k * NW;
    const int row = (int)(t / spr);
    if (row != cur_row) {
      if (cur_row >= 0) {
        float s = warp_sum((a0 + a1) + (a2 + a3));
        if (lane == 0) part[cur_row * NW + warp] = s;
      }
      a0 = a1 = a2 = a3 = 0;
      cur_row = row;
    }
    mbar_wait(&mybar[slot], parity);
    const float4* sp = reinterpret_cast<const float4*>(myring + (size_t)slot * SB);
#pragma unroll
    for (int j = 0; j < SB / 512; ++j) {
      float4 v = sp[j * 32 + lane];
      a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
    }
    __syncwarp();
    if (lane == 0 && k + NS < mine) {
      mbar_expect_tx(&mybar[slot], SB);
      bulk_g2s(myring + (size_t)slot * SB, gbase + (size_t)(warp + (k + NS) * NW) * SB, SB, &mybar[slot]);
    }
  }
  if (cur_row >= 0) {
    float s = warp_sum((a0 + a1) + (a2 + a3));
    if (lane == 0) part[cur_row * NW + warp] = s;
  }
  __syncthreads();
  for (int r = threadIdx.x; r < nrows; r += NW * 32) {
    float t = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) t += part[r * NW + w];
    out[r0 + r] = t;
  }
}

// K4: persistent CTA, plain LDG, same span partition as K3 (isolates the TMA effect from the partition effect)
template <int NW, int UNROLL>
__global__ void __launch_bounds__(NW * 32) k_span_ldg(const float* __restrict__ x, float* __restrict__ out, int M, int L, int rows_per_cta) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* part = reinterpret_cast<float*>(smem);  // [rows_per_cta][NW]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(M, r0 + rows_per_cta);
  const int nrows = r1 - r0;
  if (nrows <= 0) return;
  // warp w takes 512·UNROLL-byte pieces w, w+NW, … of every row
  const int n4 = L / 4;                 // float4 per row
  const int piece = 32 * UNROLL;        // float4 per piece
  for (int r = 0; r < nrows; ++r) {
    const float4* p = reinterpret_cast<const float4*>(x + (size_t)(r0 + r) * L);
    float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int c = warp * piece; c < n4; c += NW * piece) {
      float4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) if (c + u * 32 + lane < n4) v[u] = ldg_stream(p + c + u * 32 + lane);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) if (c + u * 32 + lane < n4) { a0 += v[u].x; a1 += v[u].y; a2 += v[u].z; a3 += v[u].w; }
    }
    float s = warp_sum((a0 + a1) + (a2 + a3));
    if (lane == 0) part[r * NW + warp] = s;
  }
  __syncthreads();
  for (int r = threadIdx.x; r < nrows; r += NW * 32) {
    float t = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) t += part[r * NW + w];
    out[r0 + r] = t;
  }
}

struct Bufs { std::vector<float*> x; std::vector<float*> o; };

template <typename F>
float time_it(F launch, int reps, int nb) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3 * nb; ++i) launch(i % nb);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) launch(i % nb);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaGetLastError());
  return ms * 1e3f / reps;
}

int main(int argc, char** argv) {
  std::vector<std::pair<int, int>> shapes;
  for (int i = 1; i + 1 < argc; i += 2) shapes.push_back({atoi(argv[i]), atoi(argv[i + 1])});
  if (shapes.empty()) shapes = {{4096, 4096}, {8192, 8192}, {32768, 1568}, {16384, 16384}};
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  for (auto sh : shapes) {
    const int M = sh.first, L = sh.second;
    const size_t bytes = (size_t)M * L * 4;
    const int nb = bytes < (size_t)200 << 20 ? 4 : 2;
    Bufs b;
    for (int i = 0; i < nb; ++i) {
      float *x, *o;
      CK(cudaMalloc(&x, bytes)); CK(cudaMalloc(&o, (size_t)M * 4));
      std::vector<float> h((size_t)L);
      for (int j = 0; j < L; ++j) h[j] = (float)((j * 37 + i) % 17) - 8.f;
      for (int r = 0; r < M; ++r) CK(cudaMemcpyAsync(x + (size_t)r * L, h.data(), (size_t)L * 4, cudaMemcpyHostToDevice));
      CK(cudaDeviceSynchronize());
      b.x.push_back(x); b.o.push_back(o);
    }
    double want = 0; for (int j = 0; j < L; ++j) want += (double)((float)((j * 37 + 0) % 17) - 8.f);
    const int reps = 200;
    auto report = [&](const char* name, float us, int bi = 0) {
      std::vector<float> ho(M);
      CK(cudaMemcpy(ho.data(), b.o[0], (size_t)M * 4, cudaMemcpyDeviceToHost));
      bool ok = true;
      for (int r = 0; r < M; r += (M / 64 > 0 ? M / 64 : 1)) if (fabs(ho[r] - want) > 1e-3 * (fabs(want) + 1)) ok = false;
      if (fabs(ho[M - 1] - want) > 1e-3 * (fabs(want) + 1)) ok = false;
      printf("[%6d x %6d] %-34s %8.2f us  %7.1f GB/s  %s\n", M, L, name, us, bytes / (us * 1e-6) / 1e9, ok ? "ok" : "WRONG");
      CK(cudaMemset(b.o[0], 0, (size_t)M * 4));
      fflush(stdout);
    };
    report("warp_row<4>", time_it([&](int i) { k_warp_row<4><<<(M + 7) / 8, 256>>>(b.x[i], b.o[i], M, L); }, reps, nb));
    report("warp_row<8>", time_it([&](int i) { k_warp_row<8><<<(M + 7) / 8, 256>>>(b.x[i], b.o[i], M, L); }, reps, nb));
    for (int occ : {2, 4, 6, 8}) {
      char nm[64]; snprintf(nm, sizeof nm, "warp_row_persist<4> occ=%d", occ);
      report(nm, time_it([&](int i) { k_warp_row_persist<4><<<sms * occ, 256>>>(b.x[i], b.o[i], M, L); }, reps, nb));
      snprintf(nm, sizeof nm, "warp_row_persist<8> occ=%d", occ);
      report(nm, time_it([&](int i) { k_warp_row_persist<8><<<sms * occ, 256>>>(b.x[i], b.o[i], M, L); }, reps, nb));
    }
    report("cta_row<4>", time_it([&](int i) { k_cta_row<4><<<M, 256>>>(b.x[i], b.o[i], M, L); }, reps, nb));
    report("cta_row<8>", time_it([&](int i) { k_cta_row<8><<<M, 256>>>(b.x[i], b.o[i], M, L); }, reps, nb));
    {
      const int rpc = (M + sms - 1) / sms;
#define RUN_TMA(NW, NS, SB)                                                                                      \
  if (((size_t)L * 4) % SB == 0) {                                                                               \
    size_t sm = (size_t)NW * NS * SB + (size_t)NW * NS * 8 + (size_t)rpc * NW * 4;                              \
    if (sm <= 227 * 1024) {                                                                                      \
      CK(cudaFuncSetAttribute(k_tma_ring<NW, NS, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));    \
      char nm[64]; snprintf(nm, sizeof nm, "tma_ring<NW=%d,NS=%d,SB=%d>", NW, NS, SB);                           \
      report(nm, time_it([&](int i) { k_tma_ring<NW, NS, SB><<<sms, NW * 32, sm>>>(b.x[i], b.o[i], M, L, rpc); }, reps, nb)); \
    }                                                                                                            \
  }
      RUN_TMA(8, 4, 2048) RUN_TMA(8, 8, 2048) RUN_TMA(8, 12, 2048) RUN_TMA(8, 4, 4096) RUN_TMA(8, 6, 4096)
      RUN_TMA(4, 8, 4096) RUN_TMA(4, 6, 8192) RUN_TMA(16, 6, 2048) RUN_TMA(16, 3, 4096) RUN_TMA(8, 3, 8192)
      RUN_TMA(8, 16, 1024) RUN_TMA(16, 12, 1024) RUN_TMA(8, 8, 1568 * 2) RUN_TMA(8, 16, 1568)
#define RUN_SPAN(NW, U)                                                                                          \
  {                                                                                                              \
    size_t sm = (size_t)rpc * NW * 4;                                                                            \
    if (sm > 48 * 1024) CK(cudaFuncSetAttribute(k_span_ldg<NW, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
    char nm[64]; snprintf(nm, sizeof nm, "span_ldg<NW=%d,U=%d> 1 CTA/SM", NW, U);                                \
    report(nm, time_it([&](int i) { k_span_ldg<NW, U><<<sms, NW * 32, sm>>>(b.x[i], b.o[i], M, L, rpc); }, reps, nb)); \
  }
      RUN_SPAN(8, 4) RUN_SPAN(16, 4) RUN_SPAN(32, 4) RUN_SPAN(32, 8) RUN_SPAN(16, 8)
    }
    for (int i = 0; i < nb; ++i) { cudaFree(b.x[i]); cudaFree(b.o[i]); }
  }
  return 0;
}
