// comm.cpp — multi-GPU: one NCCL rank per process/GPU, sharded reductions.
//
// New functionality: the reference has no collectives and no multi-GPU support (SURVEY.md fact 4;
// `cudarc` is built without its nccl feature, Cargo.toml:75-87; the device is a const generic,
// hpt/src/tensor.rs:32).  A tensor sharded along its outermost axis is k per-GPU tensors; elementwise
// ops and reductions over other axes need no exchange, reductions that cross the shard axis exchange
// one accumulator per output and rank (4 B – 64 KB for BASELINE config 5): through peer-mapped mailboxes written
// by the reduce kernel itself (xchg.cuh), or — without peer memory — with ncclAllGather over NVLink.
//
// NCCL is resolved at run time with dlopen (the process usually already holds torch's bundled
// libnccl.so.2; the same soname resolves to it), so libhpt_b200.so has no link-time NCCL dependency
// and single-GPU users never load it.
#include <dlfcn.h>

#include <mutex>
#include <vector>

#include "context.h"
#include "xchg.cuh"

extern "C" hptb_status hptb_reduce(hptb_ctx*, int, const hptb_tensor*, const int32_t*, int, hptb_tensor*, int, void*);
namespace hptb {
// api_reduce.cpp
hptb_status reduce_for_exchange(hptb_ctx* ctx, int op, const hptb_tensor* in, const int32_t* axes, int naxes, const hptb_tensor* lay,
                                double count, const XchgParams* x, void* raw, bool* fused, size_t* acc_bytes, void* stream);
hptb_status reduce_combine(hptb_ctx* ctx, int op, int in_dtype, double count, const void* partials, int gathered, const XchgParams* x,
                           const hptb_tensor* out, void* stream);
hptb_status add_offset_pairs(void* pairs, int64_t off, int64_t n, cudaStream_t s);  // sharded.cu
size_t p2p_mailbox_bytes(int nranks, size_t slot_bytes);
hptb_status p2p_allreduce(int dtype, int op, void* inout, int64_t n, void* const* boxes, int nranks, int rank, size_t slot_bytes,
                          uint32_t seq, cudaStream_t s);
}

namespace hptb {
namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5, ncclFloat16 = 6,
       ncclFloat32 = 7, ncclFloat64 = 8, ncclBfloat16 = 9 };
enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 };

struct Nccl {
  void* handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

Nccl& nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      n.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (n.handle) break;
    }
    if (!n.handle) return;
    n.GetUniqueId = (int (*)(ncclUniqueId*))dlsym(n.handle, "ncclGetUniqueId");
    n.CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(n.handle, "ncclCommInitRank");
    n.CommDestroy = (int (*)(ncclComm_t))dlsym(n.handle, "ncclCommDestroy");
    n.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(n.handle, "ncclAllReduce");
    n.AllGather = (int (*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t))dlsym(n.handle, "ncclAllGather");
    n.GroupStart = (int (*)())dlsym(n.handle, "ncclGroupStart");
    n.GroupEnd = (int (*)())dlsym(n.handle, "ncclGroupEnd");
    n.GetErrorString = (const char* (*)(int))dlsym(n.handle, "ncclGetErrorString");
    n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllReduce && n.AllGather && n.GroupStart && n.GroupEnd &&
           n.GetErrorString;
  });
  return n;
}

hptb_status nccl_fail(const char* what, int rc) {
  return fail(HPTB_ERR_NCCL, "%s failed: %s (%d)", what, nccl().GetErrorString ? nccl().GetErrorString(rc) : "?", rc);
}

int nccl_dtype(int dt) {
  switch (dt) {
    case HPTB_BOOL: case HPTB_U8: return ncclUint8;
    case HPTB_I8: return ncclInt8;
    case HPTB_I32: return ncclInt32;
    case HPTB_U32: return ncclUint32;
    case HPTB_I64: return ncclInt64;
    case HPTB_U64: return ncclUint64;
    case HPTB_F16: return ncclFloat16;
    case HPTB_BF16: return ncclBfloat16;
    case HPTB_F32: return ncclFloat32;
    case HPTB_F64: return ncclFloat64;
    default: return -1;  // i16/u16: NCCL has no 16-bit integer type
  }
}

bool is_contiguous(const hptb_tensor& t) {
  int64_t exp = 1;
  for (int i = t.ndim - 1; i >= 0; --i) {
    if (t.shape[i] == 1) continue;
    if (t.strides[i] != exp) return false;
    exp *= t.shape[i];
  }
  return true;
}

}  // namespace
}  // namespace hptb

struct hptb_comm {
  hptb_ctx* ctx = nullptr;
  hptb::ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  // Peer-mapped mailboxes (CUDA IPC over NVLink): box[r] = rank r's mailbox as mapped in this process.  Two regions:
  // [0, ll_bytes) the flag-in-data exchange of reduction accumulators (xchg.cuh), then the slots + flags of the plain
  // small-message allreduce (sharded.cu p2p_allreduce_kernel).
  bool p2p = false;
  void* box[16] = {nullptr};
  void* box_ar[16] = {nullptr};  // box[r] + ll_bytes
  size_t ll_slot_bytes = 0, ll_bytes = 0;
  size_t slot_bytes = 0;
  uint32_t seq = 0;     // allreduce call number
  uint32_t ll_seq = 0;  // sharded-reduction call number
  bool local = false;   // virtual rank of hptb_comm_init_local_group: no NCCL, mailboxes are plain allocations of one device
};

namespace hptb {
namespace {
constexpr size_t kSlotBytes = 256 * 1024;        // allreduce region: per-rank payload capacity
constexpr size_t kLLSlotBytes = 2 * 1024 * 1024; // exchange region: 65536 outputs of the widest (16-byte) accumulator per call
constexpr size_t kLLEntryMax = 32;               // bytes per output of the widest accumulator (4 words × {word, call number})

// Map every rank's mailbox into this process: cudaMalloc + CUDA IPC handles exchanged with ncclAllGather.  Any
// failure (no peer access, IPC unavailable in the container, HPTB_NO_P2P=1 on ANY rank) leaves p2p off and NCCL carries
// the data.  Every rank takes part in every collective below whatever happened to its own setup — the tiny device
// buffers those collectives need are allocated before anything that may fail, and a rank that cannot even get those
// reports an error from hptb_comm_init_rank.
hptb_status setup_p2p(hptb_comm* c) {
  if (c->nranks < 2 || c->nranks > 16) return HPTB_OK;  // same on every rank
  const int k = c->nranks;
  struct Rec {
    cudaIpcMemHandle_t h;
    int ok;
  };
  Rec* dev = nullptr;
  if (cudaMalloc((void**)&dev, sizeof(Rec) * k) != cudaSuccess) {
    cudaGetLastError();
    return fail(HPTB_ERR_OOM, "comm_init_rank: cannot allocate the %zu-byte handle exchange buffer", sizeof(Rec) * k);
  }
  std::vector<Rec> recs(k);
  memset(recs.data(), 0, sizeof(Rec) * k);
  const char* off = getenv("HPTB_NO_P2P");
  int ok = !(off && off[0] == '1');
  const size_t ll_bytes = xchg_mailbox_bytes(k, kLLSlotBytes);
  const size_t bytes = ll_bytes + p2p_mailbox_bytes(k, kSlotBytes);
  void* mine = nullptr;
  if (ok && cudaMalloc(&mine, bytes) != cudaSuccess) { ok = 0; mine = nullptr; }
  if (ok && cudaMemset(mine, 0, bytes) != cudaSuccess) ok = 0;
  if (ok && cudaIpcGetMemHandle(&recs[c->rank].h, mine) != cudaSuccess) ok = 0;
  cudaGetLastError();
  auto gather = [&](int my_ok) -> bool {  // all-gather {handle, ok}; false = the collective itself failed
    recs[c->rank].ok = my_ok;
    if (cudaMemcpy(dev + c->rank, &recs[c->rank], sizeof(Rec), cudaMemcpyHostToDevice) != cudaSuccess) return false;
    cudaDeviceSynchronize();
    if (nccl().AllGather(dev + c->rank, dev, sizeof(Rec), ncclInt8, c->comm, 0) != ncclSuccess) return false;
    if (cudaStreamSynchronize(0) != cudaSuccess) return false;
    return cudaMemcpy(recs.data(), dev, sizeof(Rec) * k, cudaMemcpyDeviceToHost) == cudaSuccess;
  };
  bool coll_ok = gather(ok);
  for (int r = 0; r < k && coll_ok && ok; ++r) ok = recs[r].ok;
  if (coll_ok && ok) {
    for (int r = 0; r < k && ok; ++r) {
      if (r == c->rank) { c->box[r] = mine; continue; }
      if (cudaIpcOpenMemHandle(&c->box[r], recs[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; c->box[r] = nullptr; }
    }
  }
  // agree on the outcome: p2p is used only if EVERY rank mapped every mailbox (a mixed choice would deadlock)
  int all_ok = coll_ok && ok;
  if (coll_ok) {
    coll_ok = gather(all_ok);
    for (int r = 0; r < k && coll_ok; ++r) all_ok = all_ok && recs[r].ok;
  }
  if (!coll_ok) all_ok = 0;
  cudaGetLastError();
  cudaFree(dev);
  if (all_ok) {
    c->p2p = true;
    c->slot_bytes = kSlotBytes;
    c->ll_slot_bytes = kLLSlotBytes;
    c->ll_bytes = ll_bytes;
    for (int r = 0; r < k; ++r) c->box_ar[r] = static_cast<char*>(c->box[r]) + ll_bytes;
  } else {
    for (int r = 0; r < k; ++r)
      if (r != c->rank && c->box[r]) cudaIpcCloseMemHandle(c->box[r]);
    if (mine) cudaFree(mine);
    memset(c->box, 0, sizeof(c->box));
    cudaGetLastError();
  }
  return HPTB_OK;
}
}  // namespace
}  // namespace hptb

using namespace hptb;

extern "C" {

hptb_status hptb_comm_unique_id(void* id128) {
  if (!id128) return fail(HPTB_ERR_INVALID, "comm_unique_id: null buffer");
  if (!nccl().ok) return fail(HPTB_ERR_NCCL, "NCCL is not available (dlopen libnccl.so.2 failed)");
  ncclUniqueId id;
  int rc = nccl().GetUniqueId(&id);
  if (rc != ncclSuccess) return nccl_fail("ncclGetUniqueId", rc);
  memcpy(id128, &id, sizeof(id));
  return HPTB_OK;
}

hptb_status hptb_comm_init_rank(hptb_ctx* ctx, int nranks, int rank, const void* id128, hptb_comm** out) {
  if (!ctx || !id128 || !out) return fail(HPTB_ERR_INVALID, "comm_init_rank: null argument");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(HPTB_ERR_INVALID, "comm_init_rank: bad rank %d of %d", rank, nranks);
  if (!nccl().ok) return fail(HPTB_ERR_NCCL, "NCCL is not available (dlopen libnccl.so.2 failed)");
  DeviceGuard g(ctx->device);
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  hptb_comm* c = new hptb_comm();
  c->ctx = ctx;
  c->nranks = nranks;
  c->rank = rank;
  int rc = nccl().CommInitRank(&c->comm, nranks, id, rank);
  if (rc != ncclSuccess) { delete c; return nccl_fail("ncclCommInitRank", rc); }
  hptb_status st = setup_p2p(c);
  if (st != HPTB_OK) { nccl().CommDestroy(c->comm); delete c; return st; }
  *out = c;
  return HPTB_OK;
}

hptb_status hptb_comm_init_local_group(hptb_ctx* ctx, int nranks, hptb_comm** comms) {
  if (!ctx || !comms) return fail(HPTB_ERR_INVALID, "comm_init_local_group: null argument");
  if (nranks < 1 || nranks > 16) return fail(HPTB_ERR_INVALID, "comm_init_local_group: %d ranks (1..16)", nranks);
  DeviceGuard g(ctx->device);
  const size_t ll_bytes = xchg_mailbox_bytes(nranks, kLLSlotBytes);
  const size_t bytes = ll_bytes + p2p_mailbox_bytes(nranks, kSlotBytes);
  void* boxes[16] = {nullptr};
  for (int r = 0; r < nranks; ++r) {
    if (cudaMalloc(&boxes[r], bytes) != cudaSuccess || cudaMemset(boxes[r], 0, bytes) != cudaSuccess) {
      cudaGetLastError();
      for (int q = 0; q <= r; ++q)
        if (boxes[q]) cudaFree(boxes[q]);
      return fail(HPTB_ERR_OOM, "comm_init_local_group: mailbox allocation failed");
    }
  }
  cudaDeviceSynchronize();
  for (int r = 0; r < nranks; ++r) {
    hptb_comm* c = new hptb_comm();
    c->ctx = ctx;
    c->nranks = nranks;
    c->rank = r;
    c->local = true;
    c->p2p = nranks > 1;
    c->slot_bytes = kSlotBytes;
    c->ll_slot_bytes = kLLSlotBytes;
    c->ll_bytes = ll_bytes;
    for (int q = 0; q < nranks; ++q) {
      c->box[q] = boxes[q];
      c->box_ar[q] = static_cast<char*>(boxes[q]) + ll_bytes;
    }
    comms[r] = c;
  }
  return HPTB_OK;
}

hptb_status hptb_comm_destroy(hptb_comm* comm) {
  if (!comm) return HPTB_OK;
  if (comm->local) {  // every virtual rank frees its own mailbox
    DeviceGuard g(comm->ctx->device);
    cudaDeviceSynchronize();
    if (comm->box[comm->rank]) cudaFree(comm->box[comm->rank]);
    delete comm;
    return HPTB_OK;
  }
  if (comm->p2p) {
    DeviceGuard g(comm->ctx->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < comm->nranks; ++r) {
      if (!comm->box[r]) continue;
      if (r == comm->rank) cudaFree(comm->box[r]);
      else cudaIpcCloseMemHandle(comm->box[r]);
    }
  }
  if (comm->comm) nccl().CommDestroy(comm->comm);
  delete comm;
  return HPTB_OK;
}

int hptb_comm_uses_peer_memory(const hptb_comm* comm) { return comm && comm->p2p ? 1 : 0; }

hptb_status hptb_shard_bounds(int64_t n, int world, int rank, int64_t* offset, int64_t* len) {
  if (!offset || !len) return fail(HPTB_ERR_INVALID, "shard_bounds: null argument");
  if (n < 0 || world < 1 || rank < 0 || rank >= world) return fail(HPTB_ERR_INVALID, "shard_bounds: bad rank %d of %d", rank, world);
  const int64_t base = n / world, rem = n % world;
  *offset = base * rank + (rank < rem ? rank : rem);
  *len = base + (rank < rem ? 1 : 0);
  return HPTB_OK;
}

hptb_status hptb_shard_plan_reduce(int op, const int32_t* axes, int naxes, int shard_axis, int world, hptb_shard_plan* plan) {
  if (!plan || (!axes && naxes)) return fail(HPTB_ERR_INVALID, "shard_plan_reduce: null argument");
  if (op < 0 || op >= HPTB_REDUCE_COUNT) return fail(HPTB_ERR_INVALID, "shard_plan_reduce: bad op %d", op);
  memset(plan, 0, sizeof(*plan));
  for (int i = 0; i < naxes; ++i) plan->crosses |= axes[i] == shard_axis;
  if (!plan->crosses || world <= 1) return HPTB_OK;
  switch (op) {
    case HPTB_SUM: case HPTB_SUM_SQUARE: case HPTB_REDUCEL1: case HPTB_NANSUM: plan->collective = HPTB_COLL_ALLREDUCE_SUM; break;
    case HPTB_NANPROD: plan->collective = HPTB_COLL_ALLREDUCE_PROD; break;
    case HPTB_ANY: plan->collective = HPTB_COLL_ALLREDUCE_MAX; break;  // OR of 0/1 bytes
    case HPTB_ALL: plan->collective = HPTB_COLL_ALLREDUCE_MIN; break;  // AND of 0/1 bytes
    case HPTB_MEAN: plan->collective = HPTB_COLL_ALLREDUCE_SUM; plan->global_count = 1; break;
    case HPTB_LOGSUMEXP: plan->collective = HPTB_COLL_ALLREDUCE_SUM; plan->pre_exp = 1; plan->post_ln = 1; break;
    case HPTB_PROD: plan->collective = HPTB_COLL_ALLREDUCE_PROD; break;
    case HPTB_MAX: plan->collective = HPTB_COLL_ALLREDUCE_MAX; break;
    case HPTB_MIN: plan->collective = HPTB_COLL_ALLREDUCE_MIN; break;
    case HPTB_ARGMAX: case HPTB_ARGMIN: plan->collective = HPTB_COLL_ALLGATHER_ARG; break;
    case HPTB_REDUCEL2: case HPTB_REDUCEL3: plan->collective = HPTB_COLL_ALLREDUCE_SUM; plan->post_root = op == HPTB_REDUCEL2 ? 2 : 3; break;
    default: return fail(HPTB_ERR_UNSUPPORTED, "shard_plan_reduce: op %d across the shard axis is not implemented (reduce locally and combine)", op);
  }
  return HPTB_OK;
}

hptb_status hptb_allreduce(hptb_comm* comm, int op, hptb_tensor* t, void* stream) {
  if (!comm) return fail(HPTB_ERR_INVALID, "allreduce: null comm");
  HPTB_TRY(validate_tensor(t, "allreduce tensor"));
  if (!is_contiguous(*t)) return fail(HPTB_ERR_SHAPE, "allreduce: the partial must be contiguous");
  int ndt = nccl_dtype(t->dtype);
  int nop;
  switch (op) {
    case HPTB_SUM: case HPTB_SUM_SQUARE: case HPTB_MEAN: case HPTB_REDUCEL1: case HPTB_NANSUM:
      nop = t->dtype == HPTB_BOOL ? ncclMax : ncclSum; break;  // bool add = OR
    case HPTB_PROD: case HPTB_NANPROD: nop = t->dtype == HPTB_BOOL ? ncclMin : ncclProd; break;  // bool mul = AND
    case HPTB_MAX: case HPTB_ANY: nop = ncclMax; break;
    case HPTB_MIN: case HPTB_ALL: nop = ncclMin; break;
    default: return fail(HPTB_ERR_INVALID, "allreduce: op %d has no collective form", op);
  }
  if (comm->nranks == 1) return HPTB_OK;
  DeviceGuard g(comm->ctx->device);
  const int64_t n = numel(*t);
  // Peer-memory path for the tiny partials (≤ 16 KB: a full reduction exchanges 4–8 B per rank).  Measured on 8 B200s,
  // config 5: sum() 328 µs vs 334 µs through ncclAllReduce; for the 64 KB partial of sum(axis 0) pushing 7 copies
  // with a handful of CTAs is SLOWER than NCCL (375 vs 366 µs), so larger partials stay on NCCL.
  if (comm->p2p && n > 0 && (size_t)n * dtype_size(t->dtype) <= 16 * 1024 && (size_t)n * dtype_size(t->dtype) <= comm->slot_bytes) {
    const int pop = nop == ncclProd ? HPTB_PROD : nop == ncclMax ? HPTB_MAX : nop == ncclMin ? HPTB_MIN : HPTB_SUM;
    const int pdt = (t->dtype == HPTB_BOOL) ? HPTB_U8 : t->dtype;  // bool OR / AND = max / min of 0/1 bytes
    HPTB_TRY(p2p_allreduce(pdt, pop, t->data, n, comm->box_ar, comm->nranks, comm->rank, comm->slot_bytes, ++comm->seq, (cudaStream_t)stream));
    count_launches(1);
    return HPTB_OK;
  }
  if (comm->local) return fail(HPTB_ERR_UNSUPPORTED, "allreduce: a local group carries at most 16 KB per call (no NCCL path)");
  if (ndt < 0) return fail(HPTB_ERR_DTYPE, "allreduce: NCCL has no type for %s", dtype_name(t->dtype));
  int rc = nccl().AllReduce(t->data, t->data, (size_t)numel(*t), ndt, nop, comm->comm, (cudaStream_t)stream);
  if (rc != ncclSuccess) return nccl_fail("ncclAllReduce", rc);
  return HPTB_OK;
}

hptb_status hptb_reduce_sharded(hptb_comm* comm, int op, const hptb_tensor* shard, const int32_t* axes, int naxes,
                                int shard_axis, int64_t shard_offset, int64_t global_axis_len, hptb_tensor* out, void* stream) {
  if (!comm) return fail(HPTB_ERR_INVALID, "reduce_sharded: null comm");
  HPTB_TRY(validate_tensor(shard, "reduce_sharded shard"));
  HPTB_TRY(validate_tensor(out, "reduce_sharded out"));
  if (shard_axis < 0 || shard_axis >= shard->ndim) return fail(HPTB_ERR_AXIS, "reduce_sharded: shard axis %d out of range", shard_axis);
  if (op < 0 || op >= HPTB_REDUCE_COUNT) return fail(HPTB_ERR_INVALID, "reduce_sharded: bad op %d", op);
  if (!axes && naxes) return fail(HPTB_ERR_INVALID, "reduce_sharded: null axes");
  hptb_shard_plan sp;
  HPTB_TRY(hptb_shard_plan_reduce(op, axes, naxes, shard_axis, comm->nranks, &sp));
  const bool crosses = sp.crosses != 0;
  const bool is_arg = op == HPTB_ARGMAX || op == HPTB_ARGMIN;
  const bool offset_only = comm->nranks == 1 && crosses && is_arg && shard_offset != 0;  // a lone shard that does not start at 0
  if ((!crosses || comm->nranks == 1) && !offset_only) return hptb_reduce(comm->ctx, op, shard, axes, naxes, out, 1, stream);
  // ---- the reduction crosses the shard axis: every rank reduces its shard to one ACCUMULATOR per output (f32 for
  // f16/bf16/f32, (value, global index) for argmax/argmin, Σexp for logsumexp, the power sum for reducel2/3), the k
  // accumulators are combined in rank order and the op's post step (÷ GLOBAL count, ln, root, index) runs once — the
  // sharded result is rounded exactly like the single-GPU one.  With peer memory the exchange is part of the reduce
  // kernel (xchg.cuh: one launch per rank) or, for launch shapes that cannot host it, of one small follow-up kernel;
  // without it NCCL all-gathers the accumulators.
  if (is_arg && naxes != 1) return fail(HPTB_ERR_AXIS, "argmax/argmin take exactly one axis (got %d)", naxes);
  const int odt = hptb_reduce_out_dtype(op, shard->dtype);
  if (out->dtype != odt) return fail(HPTB_ERR_DTYPE, "reduce_sharded: out dtype is %s, expected %s", dtype_name(out->dtype), dtype_name(odt));
  const int64_t M = numel(*out);
  if (M == 0) return HPTB_OK;
  double count = (double)global_axis_len;
  for (int i = 0; i < naxes; ++i)
    if (axes[i] != shard_axis) count *= (double)shard->shape[axes[i]];
  hptb_tensor lay = *out;  // out's shape, row-major
  int64_t st = 1;
  for (int i = lay.ndim - 1; i >= 0; --i) { lay.strides[i] = st; st *= lay.shape[i]; }
  lay.data = is_contiguous(*out) ? out->data : nullptr;
  const int k = comm->nranks;
  const bool use_ll = comm->p2p && k > 1 && (uint64_t)M * kLLEntryMax <= comm->ll_slot_bytes;  // the same on every rank
  XchgParams x;
  memset(&x, 0, sizeof(x));
  if (use_ll) {
    for (int r = 0; r < k; ++r) x.box[r] = static_cast<unsigned char*>(comm->box[r]);
    x.slot_bytes = comm->ll_slot_bytes;
    x.idx_offset = is_arg ? shard_offset : 0;
    x.seq = ++comm->ll_seq;
    x.nranks = k;
    x.rank = comm->rank;
  }
  Scratch raw;
  HPTB_TRY(raw.get(comm->ctx, (size_t)M * 16, stream));
  bool fused = false;
  size_t accb = 0;
  HPTB_TRY(reduce_for_exchange(comm->ctx, op, shard, axes, naxes, &lay, count, use_ll ? &x : nullptr, raw.ptr, &fused, &accb, stream));
  if (fused) return HPTB_OK;
  if (use_ll) return reduce_combine(comm->ctx, op, shard->dtype, count, raw.ptr, 0, &x, out, stream);
  if (comm->local && k > 1)
    return fail(HPTB_ERR_UNSUPPORTED, "reduce_sharded: %lld outputs exceed the mailboxes of a local group (no NCCL path)", (long long)M);
  DeviceGuard g(comm->ctx->device);
  cudaStream_t s = (cudaStream_t)stream;
  if (is_arg && shard_offset != 0) {
    HPTB_TRY(add_offset_pairs(raw.ptr, shard_offset, M, s));
    count_launches(1);
  }
  if (k == 1) return reduce_combine(comm->ctx, op, shard->dtype, count, raw.ptr, 1, nullptr, out, stream);
  Scratch all;
  HPTB_TRY(all.get(comm->ctx, (size_t)M * accb * k, stream));
  int rc = nccl().AllGather(raw.ptr, all.ptr, (size_t)M * accb, ncclInt8, comm->comm, s);
  if (rc != ncclSuccess) return nccl_fail("ncclAllGather", rc);
  return reduce_combine(comm->ctx, op, shard->dtype, count, all.ptr, k, nullptr, out, stream);
}

}  // extern "C"
