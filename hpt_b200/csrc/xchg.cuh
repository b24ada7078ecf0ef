// xchg.cuh — peer-memory exchange of reduction partials, fused into the reduce kernels' epilogue.
//
// New functionality (the reference has no multi-GPU path).  A reduction that crosses the shard
// axis of an outer-axis-sharded tensor (BASELINE config 5) leaves one ACCUMULATOR per output on every rank; the
// k accumulators must be combined and the op's post step (÷ n, ln, root, index) applied once.  Here that exchange
// is part of the reduce kernel itself: the thread that holds the final local accumulator of output m
//   1. stores it into entry m of slot [my rank] of EVERY rank's mailbox (peer-mapped memory, plain NVLink stores),
//   2. spins on entry m of the k slots of its OWN mailbox,
//   3. combines the k accumulators in RANK ORDER (deterministic, bit-identical on every rank), applies post, stores.
// The protocol is flag-in-data ("LL"): an entry is ceil(sizeof(Acc)/4) 8-byte words {payload word, call number};
// an aligned 8-byte store is one transaction on NVLink, so a word is valid exactly when its call number matches —
// no fences, no separate flags, and — because entries are addressed by OUTPUT INDEX — nothing depends on which
// kernel shape (grid, split count, vector width) each rank picked: ranks with different alignments or shard
// lengths interoperate, as does the standalone kernel (xchg_combine_kernel) that serves unfused shapes.
// Accumulators travel in the accumulation type: f32 for f16/bf16/f32 inputs (SURVEY.md §8e "accumulate/allreduce
// in f32"), (value, global index) pairs for argmax/argmin, Σexp for logsumexp, the unrooted power sum for
// reducel2/3 — so the sharded result is rounded ONCE, exactly like the single-GPU kernel's.
// Two buffers alternate by call parity.  A rank can run at most one call ahead of its slowest peer (it spins in
// call n+1 until that peer has pushed call n+1, which the peer does only after finishing call n on the same
// stream), so buffer (n & 1) is never overwritten while somebody still reads call n from it.  One stream per comm.
// Progress: a spinning thread waits only for REMOTE pushes, and a push never waits; the host fuses the exchange
// only when the spinning CTAs can hold at most half of the resident CTA slots, so the rest of the grid (and with
// it every push) always gets scheduled, whatever the peers are doing.
#pragma once
#include <cstdint>
#include <cstring>

namespace hptb {

constexpr int kXchgMaxRanks = 16;

struct XchgParams {
  unsigned char* box[kXchgMaxRanks];  // every rank's mailbox as mapped in THIS process (box[rank] = own)
  uint64_t slot_bytes;                // capacity of one (buffer, source rank) slot
  int64_t idx_offset;                 // argmax/argmin: this shard's offset along the reduced axis
  uint32_t seq;                       // call number, ≥ 1
  int32_t nranks, rank;
  int32_t enabled;                    // 0: plain single-GPU epilogue
};

inline size_t xchg_mailbox_bytes(int nranks, size_t slot_bytes) { return (size_t)2 * nranks * slot_bytes; }

#ifdef __CUDACC__
template <typename Acc>
struct XchgWords {
  static constexpr int n = (int)((sizeof(Acc) + 3) / 4);
};

__device__ __forceinline__ void xchg_store(unsigned char* p, uint32_t w, uint32_t seq) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(w), "r"(seq) : "memory");
}
__device__ __forceinline__ uint2 xchg_load(const unsigned char* p) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}

// entry m of slot [src] of the current buffer, relative to a mailbox base
template <typename Acc>
__device__ __forceinline__ size_t xchg_entry_off(const XchgParams& x, int src, int64_t m) {
  return ((size_t)(x.seq & 1u) * (size_t)x.nranks + (size_t)src) * x.slot_bytes + (size_t)m * (8 * XchgWords<Acc>::n);
}

template <typename Op>
__device__ __forceinline__ void xchg_push(const XchgParams& x, int64_t m, typename Op::Acc a) {
  typedef typename Op::Acc Acc;
  constexpr int W = XchgWords<Acc>::n;
  if constexpr (Op::kIndexed) a.idx += x.idx_offset;  // global index along the sharded axis
  uint32_t w[W];
#pragma unroll
  for (int i = 0; i < W; ++i) w[i] = 0;
  memcpy(w, &a, sizeof(Acc));
  const size_t off = xchg_entry_off<Acc>(x, x.rank, m);
  for (int r = 0; r < x.nranks; ++r) {
    unsigned char* dst = x.box[r] + off;
#pragma unroll
    for (int i = 0; i < W; ++i) xchg_store(dst + 8 * i, w[i], x.seq);
  }
}

template <typename Op>
__device__ __forceinline__ typename Op::Acc xchg_collect(const XchgParams& x, int64_t m) {
  typedef typename Op::Acc Acc;
  constexpr int W = XchgWords<Acc>::n;
  Acc acc = Op::identity();
  const unsigned char* mine = x.box[x.rank];
  for (int r = 0; r < x.nranks; ++r) {
    const unsigned char* src = mine + xchg_entry_off<Acc>(x, r, m);
    uint32_t w[W];
#pragma unroll
    for (int i = 0; i < W; ++i) {
      uint2 v;
      do { v = xchg_load(src + 8 * i); } while (v.y != x.seq);
      w[i] = v.x;
    }
    Acc part;
    memcpy(&part, w, sizeof(Acc));
    acc = r == 0 ? part : Op::combine(acc, part);
  }
  return acc;
}

// push + collect for ONE output held by the calling thread
template <typename Op>
__device__ __forceinline__ typename Op::Acc xchg_finish(const XchgParams& x, int64_t m, typename Op::Acc a) {
  xchg_push<Op>(x, m, a);
  return xchg_collect<Op>(x, m);
}

// ---- the same spread over threads: one (destination / source) rank per thread -------------------------------------------
// A single thread pushing to k mailboxes and then polling k entries one after the other pays k dependent round trips
// through L2 (≈ 5 µs at 8 ranks for one output, ≈ 16 µs for the 4 × 8 entries a thread of the column kernel owned);
// with one rank per lane / thread row the k stores and the k polls are in flight together.
template <typename Op>
__device__ __forceinline__ void xchg_push_to(const XchgParams& x, int dst_rank, int64_t m, typename Op::Acc a) {
  typedef typename Op::Acc Acc;
  constexpr int W = XchgWords<Acc>::n;
  if constexpr (Op::kIndexed) a.idx += x.idx_offset;
  uint32_t w[W];
#pragma unroll
  for (int i = 0; i < W; ++i) w[i] = 0;
  memcpy(w, &a, sizeof(Acc));
  unsigned char* dst = x.box[dst_rank] + xchg_entry_off<Acc>(x, x.rank, m);
#pragma unroll
  for (int i = 0; i < W; ++i) xchg_store(dst + 8 * i, w[i], x.seq);
}
template <typename Op>
__device__ __forceinline__ typename Op::Acc xchg_poll_from(const XchgParams& x, int src_rank, int64_t m) {
  typedef typename Op::Acc Acc;
  constexpr int W = XchgWords<Acc>::n;
  const unsigned char* src = x.box[x.rank] + xchg_entry_off<Acc>(x, src_rank, m);
  uint32_t w[W];
#pragma unroll
  for (int i = 0; i < W; ++i) {
    uint2 v;
    do { v = xchg_load(src + 8 * i); } while (v.y != x.seq);
    w[i] = v.x;
  }
  Acc part;
  memcpy(&part, w, sizeof(Acc));
  return part;
}
template <typename Acc>
__device__ __forceinline__ Acc xchg_shfl(Acc v, int src_lane) {
  constexpr int W = XchgWords<Acc>::n;
  uint32_t w[W];
#pragma unroll
  for (int i = 0; i < W; ++i) w[i] = 0;
  memcpy(w, &v, sizeof(Acc));
#pragma unroll
  for (int i = 0; i < W; ++i) w[i] = __shfl_sync(0xffffffffu, w[i], src_lane);
  Acc o;
  memcpy(&o, w, sizeof(Acc));
  return o;
}
// ONE output, called by all 32 lanes of a warp with the same accumulator: lane r talks to rank r; every lane returns the
// rank-ordered combination
template <typename Op>
__device__ __forceinline__ typename Op::Acc xchg_finish_warp(const XchgParams& x, int64_t m, typename Op::Acc a) {
  typedef typename Op::Acc Acc;
  const int lane = threadIdx.x & 31;
  Acc acc = Op::identity();
  for (int r0 = 0; r0 < x.nranks; r0 += 32) {  // nranks ≤ 16: one round
    const int r = r0 + lane;
    Acc part = Op::identity();
    if (r < x.nranks) {
      xchg_push_to<Op>(x, r, m, a);
      part = xchg_poll_from<Op>(x, r, m);
    }
    const int nr = x.nranks - r0 < 32 ? x.nranks - r0 : 32;
    for (int q = 0; q < nr; ++q) {
      const Acc pq = xchg_shfl<Acc>(part, q);
      acc = (r0 == 0 && q == 0) ? pq : Op::combine(acc, pq);
    }
  }
  return acc;
}
#endif  // __CUDACC__

}  // namespace hptb
