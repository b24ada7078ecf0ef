// sharded.cu — device helpers of the multi-GPU path that are not part of a reduce kernel.
//
// The exchange of reduction accumulators lives in xchg.cuh (fused into the reduce kernels' epilogue, or run by
// xchg_combine_kernel in reduce.cuh).  Here: the index offset for NCCL-gathered (value, index) accumulators, and the
// plain small-message allreduce behind hptb_allreduce.
#include "dtypes_x.h"
#include "reduce.cuh"

namespace hptb {
namespace {

// argmax/argmin accumulators are ArgPair{value (≤ 8 bytes, padded to 8), int64 index}: 16 bytes, index in the second
// half.  Turns the local index into a GLOBAL one before NCCL gathers the pairs (the mailbox path does it while pushing).
__global__ void __launch_bounds__(256) add_offset_pairs_kernel(int64_t* __restrict__ pairs, int64_t off, int64_t n) {
  pdl_prologue();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pairs[2 * i + 1] += off;
}
static_assert(sizeof(ArgPair<b8>) == 16 && sizeof(ArgPair<double>) == 16 && offsetof(ArgPair<float>, idx) == 8,
              "add_offset_pairs_kernel assumes 16-byte (value, index) accumulators");

// ---- peer-memory allreduce of small partials -------------------------------------------------------------------
// The exchange step of a sharded reduction moves 4 B – 64 KB per rank; a library allreduce costs 20–35 µs of
// launch + protocol latency for that, a visible slice of the 330 µs local pass at 8 GPUs (config 5).  Here every
// rank owns a mailbox that its peers map through CUDA IPC (comm.cpp); ONE small kernel per rank
//   1. stores its partial into slot [my rank] of every peer's mailbox with plain NVLink peer stores,
//   2. fences and publishes a per-CTA sequence flag in the peer's mailbox,
//   3. waits until the same flag arrived from every peer in its own mailbox,
//   4. combines the k slots in RANK ORDER (deterministic, identical on every rank) and writes the result in place.
// Two buffers alternate by call parity: a rank can run at most one call ahead of its slowest peer (it blocks in
// step 3 of call n+1 until that peer has entered call n+1, i.e. finished reading call n), so buffer (n & 1) is
// never overwritten while someone still reads it.  CTA c only ever waits for CTA c of the peers: no CTA of one
// GPU depends on another CTA of the same GPU, so no co-residency assumption is needed.
constexpr int kP2PMaxRanks = 16;
constexpr int kP2PMaxCtas = 8;
constexpr int kP2PThreads = 256;

struct P2PParams {
  unsigned char* box[kP2PMaxRanks];  // mailbox of every rank as mapped in THIS process (box[rank] = own)
  int64_t n;                         // elements
  uint64_t slot_bytes;               // payload capacity of one slot
  uint32_t seq;                      // call number (≥ 1)
  int32_t nranks, rank, op;          // op: hptb_reduce_op family (SUM / PROD / MAX / MIN)
};

// mailbox layout: [2 buffers][nranks slots][slot_bytes] payload, then [2][nranks][kP2PMaxCtas] flags (32 B apart)
__host__ __device__ inline size_t p2p_payload_off(int buf, int slot, int nranks, size_t slot_bytes) {
  return ((size_t)buf * nranks + slot) * slot_bytes;
}
__host__ __device__ inline size_t p2p_flag_off(int buf, int slot, int cta, int nranks, size_t slot_bytes) {
  return (size_t)2 * nranks * slot_bytes + (((size_t)buf * nranks + slot) * kP2PMaxCtas + cta) * 32;
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <typename T>
__device__ __forceinline__ compute_t<T> p2p_combine(int op, compute_t<T> a, compute_t<T> b) {
  typedef compute_t<T> C;
  switch (op) {
    case HPTB_PROD: return red_mul<C>(a, b);
    case HPTB_MAX: return red_max<C>(a, b);
    case HPTB_MIN: return red_min<C>(a, b);
    default: return red_add<C>(a, b);
  }
}

template <typename T>
__global__ void __launch_bounds__(kP2PThreads) p2p_allreduce_kernel(T* __restrict__ inout, P2PParams p) {
  pdl_prologue();
  const int buf = p.seq & 1;
  const int cta = blockIdx.x, nctas = gridDim.x;
  // this CTA's element range
  const int64_t per = (p.n + nctas - 1) / nctas;
  const int64_t e0 = (int64_t)cta * per;
  const int64_t e1 = e0 + per < p.n ? e0 + per : p.n;
  // 1. my partial → slot [rank] of every mailbox (peers over NVLink, my own locally)
  for (int r = 0; r < p.nranks; ++r) {
    T* dst = reinterpret_cast<T*>(p.box[r] + p2p_payload_off(buf, p.rank, p.nranks, p.slot_bytes));
    for (int64_t e = e0 + threadIdx.x; e < e1; e += kP2PThreads) dst[e] = inout[e];
  }
  __threadfence_system();
  __syncthreads();
  // 2. publish
  if (threadIdx.x < p.nranks) {
    const int r = threadIdx.x;
    st_release_sys(reinterpret_cast<uint32_t*>(p.box[r] + p2p_flag_off(buf, p.rank, cta, p.nranks, p.slot_bytes)), p.seq);
  }
  // 3. wait for every rank's flag in my mailbox
  if (threadIdx.x < p.nranks) {
    const uint32_t* f = reinterpret_cast<const uint32_t*>(p.box[p.rank] + p2p_flag_off(buf, threadIdx.x, cta, p.nranks, p.slot_bytes));
    while (ld_acquire_sys(f) != p.seq) {}
  }
  __syncthreads();
  // 4. rank-ordered combine (slots are read through L2: peers wrote them behind this SM's L1)
  typedef compute_t<T> C;
  for (int64_t e = e0 + threadIdx.x; e < e1; e += kP2PThreads) {
    const T* s0 = reinterpret_cast<const T*>(p.box[p.rank] + p2p_payload_off(buf, 0, p.nranks, p.slot_bytes));
    C acc = to_compute<T>(load_cg(s0 + e));
    for (int r = 1; r < p.nranks; ++r) {
      const T* sr = reinterpret_cast<const T*>(p.box[p.rank] + p2p_payload_off(buf, r, p.nranks, p.slot_bytes));
      acc = p2p_combine<T>(p.op, acc, to_compute<T>(load_cg(sr + e)));
    }
    inout[e] = from_compute<T>(acc);
  }
}

template <typename T>
hptb_status p2p_launch(T* inout, const P2PParams& p, cudaStream_t s) {
  int64_t bytes = p.n * (int64_t)sizeof(T);
  int nctas = (int)((bytes + 8191) / 8192);
  if (nctas < 1) nctas = 1;
  if (nctas > kP2PMaxCtas) nctas = kP2PMaxCtas;
  HPTB_CUDA_CHECK(launch_kernel(p2p_allreduce_kernel<T>, dim3(nctas), dim3(kP2PThreads), 0, s, inout, p));
  return HPTB_OK;
}

}  // namespace

hptb_status add_offset_pairs(void* pairs, int64_t off, int64_t n, cudaStream_t s) {
  if (n <= 0 || off == 0) return HPTB_OK;
  HPTB_CUDA_CHECK(launch_kernel(add_offset_pairs_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, static_cast<int64_t*>(pairs), off, n));
  return HPTB_OK;
}

size_t p2p_mailbox_bytes(int nranks, size_t slot_bytes) {
  return (size_t)2 * nranks * slot_bytes + (size_t)2 * nranks * kP2PMaxCtas * 32;
}

hptb_status p2p_allreduce(int dtype, int op, void* inout, int64_t n, void* const* boxes, int nranks, int rank, size_t slot_bytes,
                          uint32_t seq, cudaStream_t s) {
  if (nranks > kP2PMaxRanks) return fail(HPTB_ERR_UNSUPPORTED, "p2p_allreduce: more than %d ranks", kP2PMaxRanks);
  P2PParams p;
  memset(&p, 0, sizeof(p));
  for (int r = 0; r < nranks; ++r) p.box[r] = static_cast<unsigned char*>(boxes[r]);
  p.n = n;
  p.slot_bytes = slot_bytes;
  p.seq = seq;
  p.nranks = nranks;
  p.rank = rank;
  p.op = op;
  switch (dtype) {
#define X(T, N, E) \
  case E: return p2p_launch<T>(static_cast<T*>(inout), p, s);
    HPTB_FOR_DTYPES(X)
#undef X
    default: return fail(HPTB_ERR_DTYPE, "p2p_allreduce: bad dtype");
  }
}

}  // namespace hptb
