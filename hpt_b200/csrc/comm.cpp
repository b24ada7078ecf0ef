// comm.cpp — multi-GPU: one NCCL rank per process/GPU, sharded reductions.
//
// New functionality: the reference has no collectives and no multi-GPU support (SURVEY.md fact 4;
// `cudarc` is built without its nccl feature, Cargo.toml:75-87; the device is a const generic,
// hpt/src/tensor.rs:32).  A tensor sharded along its outermost axis is k per-GPU tensors; elementwise
// ops and reductions over other axes need no exchange, reductions that cross the shard axis exchange
// one small partial per rank (4 B – 64 KB for BASELINE config 5) with ncclAllReduce over NVLink.
//
// NCCL is resolved at run time with dlopen (the process usually already holds torch's bundled
// libnccl.so.2; the same soname resolves to it), so libhpt_b200.so has no link-time NCCL dependency
// and single-GPU users never load it.
#include <dlfcn.h>

#include <mutex>
#include <vector>

#include "context.h"

extern "C" hptb_status hptb_reduce(hptb_ctx*, int, const hptb_tensor*, const int32_t*, int, hptb_tensor*, int, void*);
extern "C" hptb_status hptb_unary(hptb_ctx*, int, const hptb_tensor*, hptb_tensor*, double, double, void*);
extern "C" hptb_status hptb_binary(hptb_ctx*, int, const hptb_tensor*, const hptb_tensor*, hptb_tensor*, void*);
extern "C" hptb_status hptb_fill(hptb_ctx*, hptb_tensor*, const void*, void*);
extern "C" hptb_status hptb_copy(hptb_ctx*, const hptb_tensor*, hptb_tensor*, void*);
namespace hptb {
hptb_status reduce_impl(hptb_ctx* ctx, int op, const hptb_tensor* in, const int32_t* axes, int naxes, hptb_tensor* out,
                        int init_out, double count_override, void* stream);  // api_reduce.cpp
hptb_status arg_combine(int dtype, bool is_max, const void* vals, const int64_t* idx, int k, int64_t M, int64_t* out,
                        cudaStream_t s);                                      // sharded.cu
hptb_status add_offset_i64(int64_t* p, int64_t off, int64_t n, cudaStream_t s);
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
  // peer-memory mailboxes for small partials (sharded.cu): box[r] = rank r's mailbox as mapped in this process
  bool p2p = false;
  void* box[16] = {nullptr};
  size_t slot_bytes = 0;
  uint32_t seq = 0;
};

namespace hptb {
namespace {
constexpr size_t kSlotBytes = 256 * 1024;  // per-rank payload capacity (config 5's sum(0) partial is 64 KB)

// Map every rank's mailbox into this process: cudaMalloc + CUDA IPC handles exchanged with ncclAllGather.  Any
// failure (no peer access, IPC unavailable in the container, HPTB_NO_P2P=1) leaves p2p off and NCCL carries the data.
void setup_p2p(hptb_comm* c) {
  const char* off = getenv("HPTB_NO_P2P");
  if ((off && off[0] == '1') || c->nranks < 2 || c->nranks > 16) return;
  const size_t bytes = p2p_mailbox_bytes(c->nranks, kSlotBytes);
  void* mine = nullptr;
  cudaIpcMemHandle_t* dev_handles = nullptr;
  std::vector<cudaIpcMemHandle_t> handles(c->nranks);
  int ok = 1;
  if (cudaMalloc(&mine, bytes) != cudaSuccess) ok = 0;
  if (ok && cudaMemset(mine, 0, bytes) != cudaSuccess) ok = 0;
  if (ok && cudaIpcGetMemHandle(&handles[c->rank], mine) != cudaSuccess) ok = 0;
  if (cudaMalloc((void**)&dev_handles, sizeof(cudaIpcMemHandle_t) * c->nranks) != cudaSuccess) { ok = 0; dev_handles = nullptr; }
  // every rank takes part in the two collectives below even if its own setup failed, so nobody hangs
  int* dev_ok = nullptr;
  std::vector<int> oks(c->nranks, 0);
  if (dev_handles && cudaMalloc((void**)&dev_ok, sizeof(int) * c->nranks) == cudaSuccess) {
    cudaMemcpy((char*)dev_handles + sizeof(cudaIpcMemHandle_t) * c->rank, &handles[c->rank], sizeof(cudaIpcMemHandle_t), cudaMemcpyHostToDevice);
    cudaMemcpy(dev_ok + c->rank, &ok, sizeof(int), cudaMemcpyHostToDevice);
    cudaDeviceSynchronize();
    int rc = nccl().GroupStart();
    if (rc == ncclSuccess) rc = nccl().AllGather((char*)dev_handles + sizeof(cudaIpcMemHandle_t) * c->rank, dev_handles, sizeof(cudaIpcMemHandle_t), ncclInt8, c->comm, 0);
    if (rc == ncclSuccess) rc = nccl().AllGather(dev_ok + c->rank, dev_ok, sizeof(int), ncclInt8, c->comm, 0);
    int rc2 = nccl().GroupEnd();
    if (rc != ncclSuccess || rc2 != ncclSuccess) ok = 0;
    if (cudaStreamSynchronize(0) != cudaSuccess) ok = 0;
    cudaMemcpy(handles.data(), dev_handles, sizeof(cudaIpcMemHandle_t) * c->nranks, cudaMemcpyDeviceToHost);
    cudaMemcpy(oks.data(), dev_ok, sizeof(int) * c->nranks, cudaMemcpyDeviceToHost);
  } else {
    ok = 0;
  }
  for (int r = 0; r < c->nranks && ok; ++r) ok = oks[r];
  if (ok) {
    for (int r = 0; r < c->nranks && ok; ++r) {
      if (r == c->rank) { c->box[r] = mine; continue; }
      if (cudaIpcOpenMemHandle(&c->box[r], handles[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = 0;
    }
  }
  // agree on the outcome: p2p is used only if EVERY rank mapped every mailbox (a mixed choice would deadlock)
  int all_ok = ok;
  if (dev_ok) {
    cudaMemcpy(dev_ok + c->rank, &ok, sizeof(int), cudaMemcpyHostToDevice);
    cudaDeviceSynchronize();
    if (nccl().AllGather(dev_ok + c->rank, dev_ok, sizeof(int), ncclInt8, c->comm, 0) == ncclSuccess && cudaStreamSynchronize(0) == cudaSuccess) {
      cudaMemcpy(oks.data(), dev_ok, sizeof(int) * c->nranks, cudaMemcpyDeviceToHost);
      for (int r = 0; r < c->nranks; ++r) all_ok = all_ok && oks[r];
    } else {
      all_ok = 0;
    }
  }
  cudaGetLastError();
  if (dev_handles) cudaFree(dev_handles);
  if (dev_ok) cudaFree(dev_ok);
  if (all_ok) {
    c->p2p = true;
    c->slot_bytes = kSlotBytes;
  } else {
    for (int r = 0; r < c->nranks; ++r)
      if (r != c->rank && c->box[r]) cudaIpcCloseMemHandle(c->box[r]);
    if (mine) cudaFree(mine);
    memset(c->box, 0, sizeof(c->box));
  }
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
  setup_p2p(c);
  *out = c;
  return HPTB_OK;
}

hptb_status hptb_comm_destroy(hptb_comm* comm) {
  if (!comm) return HPTB_OK;
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
    HPTB_TRY(p2p_allreduce(pdt, pop, t->data, n, comm->box, comm->nranks, comm->rank, comm->slot_bytes, ++comm->seq, (cudaStream_t)stream));
    count_launches(1);
    return HPTB_OK;
  }
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
  if (!crosses || comm->nranks == 1) {
    HPTB_TRY(hptb_reduce(comm->ctx, op, shard, axes, naxes, out, 1, stream));
    // a lone rank may still hold a shard that does not start at 0 (indices are GLOBAL along the shard axis)
    if (is_arg && crosses && shard_offset != 0 && is_contiguous(*out)) {
      DeviceGuard g(comm->ctx->device);
      count_launches(1);
      return add_offset_i64((int64_t*)out->data, shard_offset, numel(*out), (cudaStream_t)stream);
    }
    return HPTB_OK;
  }
  if (!is_contiguous(*out)) return fail(HPTB_ERR_SHAPE, "reduce_sharded: out must be contiguous when partials are exchanged");
  if (is_arg) {
    // (extreme value, global index) per rank → all-gather → rank-ordered strict combine (sharded.cu)
    if (naxes != 1) return fail(HPTB_ERR_AXIS, "argmax/argmin take exactly one axis (got %d)", naxes);
    if (out->dtype != HPTB_I64) return fail(HPTB_ERR_DTYPE, "reduce_sharded: out dtype is %s, expected i64", dtype_name(out->dtype));
    const int64_t M = numel(*out);
    if (M == 0) return HPTB_OK;
    const size_t esz = dtype_size(shard->dtype);
    const int k = comm->nranks;
    Scratch sv, sva, sia;
    HPTB_TRY(sv.get(comm->ctx, (size_t)M * esz, stream));
    HPTB_TRY(sva.get(comm->ctx, (size_t)M * esz * k, stream));
    HPTB_TRY(sia.get(comm->ctx, (size_t)M * sizeof(int64_t) * k, stream));
    hptb_tensor val = *out;
    val.data = sv.ptr;
    val.dtype = shard->dtype;
    HPTB_TRY(hptb_reduce(comm->ctx, op == HPTB_ARGMAX ? HPTB_MAX : HPTB_MIN, shard, axes, naxes, &val, 1, stream));
    HPTB_TRY(hptb_reduce(comm->ctx, op, shard, axes, naxes, out, 1, stream));
    DeviceGuard g(comm->ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    HPTB_TRY(add_offset_i64((int64_t*)out->data, shard_offset, M, s));
    int rc = nccl().GroupStart();
    if (rc != ncclSuccess) return nccl_fail("ncclGroupStart", rc);
    rc = nccl().AllGather(sv.ptr, sva.ptr, (size_t)M * esz, ncclInt8, comm->comm, s);
    if (rc == ncclSuccess) rc = nccl().AllGather(out->data, sia.ptr, (size_t)M * sizeof(int64_t), ncclInt8, comm->comm, s);
    int rc2 = nccl().GroupEnd();
    if (rc != ncclSuccess) return nccl_fail("ncclAllGather", rc);
    if (rc2 != ncclSuccess) return nccl_fail("ncclGroupEnd", rc2);
    HPTB_TRY(arg_combine(shard->dtype, op == HPTB_ARGMAX, sva.ptr, (const int64_t*)sia.ptr, k, M, (int64_t*)out->data, s));
    count_launches(shard_offset != 0 ? 2 : 1);
    return HPTB_OK;
  }
  if (sp.global_count) {
    // each rank computes Σ_local / n_GLOBAL, the allreduce-sum of those is the global mean
    double count = (double)global_axis_len;
    for (int i = 0; i < naxes; ++i)
      if (axes[i] != shard_axis) count *= (double)shard->shape[axes[i]];
    HPTB_TRY(reduce_impl(comm->ctx, HPTB_MEAN, shard, axes, naxes, out, 1, count, stream));
    return hptb_allreduce(comm, HPTB_SUM, out, stream);
  }
  if (sp.pre_exp) {
    HPTB_TRY(hptb_reduce(comm->ctx, HPTB_LOGSUMEXP, shard, axes, naxes, out, 1, stream));
    HPTB_TRY(hptb_unary(comm->ctx, HPTB_EXP, out, out, 0, 0, stream));  // back to Σ exp (naive domain, as the reference)
    HPTB_TRY(hptb_allreduce(comm, HPTB_SUM, out, stream));
    return hptb_unary(comm->ctx, HPTB_LN, out, out, 0, 0, stream);
  }
  if (sp.post_root) {
    // Σ|x|^p is exchanged, then the reference's root: sqrt, or pow with the exponent 1/3 rounded to the OUTPUT dtype
    // (reduce.cuh ReduceOp<HPTB_REDUCEL3>).  f32/f64 outputs: the local kernel leaves the power sum unrooted in `out`.
    // f16/bf16 outputs (8- and 16-bit inputs): a power sum does not fit half precision, so the rooted local result
    // is widened to an f32 scratch, raised to p again, exchanged and rooted there, and rounded to `out` once.
    const bool half = out->dtype == HPTB_F16 || out->dtype == HPTB_BF16;
    const int64_t M = numel(*out);
    Scratch s64, sod, swide, sthird;
    HPTB_TRY(s64.get(comm->ctx, 8, stream));
    HPTB_TRY(sod.get(comm->ctx, 8, stream));
    HPTB_TRY(sthird.get(comm->ctx, 8, stream));
    hptb_tensor acc = *out;  // where the exchange happens
    if (half) {
      HPTB_TRY(swide.get(comm->ctx, (size_t)(M > 0 ? M : 1) * sizeof(float), stream));
      HPTB_TRY(hptb_reduce(comm->ctx, op, shard, axes, naxes, out, 1, stream));
      acc.data = swide.ptr;
      acc.dtype = HPTB_F32;
      HPTB_TRY(hptb_copy(comm->ctx, out, &acc, stream));
      hptb_tensor base = acc;
      Scratch sbase;
      if (sp.post_root == 3) {  // acc = base³ needs base kept
        HPTB_TRY(sbase.get(comm->ctx, (size_t)(M > 0 ? M : 1) * sizeof(float), stream));
        base.data = sbase.ptr;
        HPTB_TRY(hptb_copy(comm->ctx, &acc, &base, stream));
        HPTB_TRY(hptb_binary(comm->ctx, HPTB_MUL, &acc, &base, &acc, stream));
      }
      HPTB_TRY(hptb_binary(comm->ctx, HPTB_MUL, &acc, &base, &acc, stream));
    } else {
      HPTB_TRY(reduce_impl(comm->ctx, op, shard, axes, naxes, out, 1, -2.0, stream));
    }
    HPTB_TRY(hptb_allreduce(comm, HPTB_SUM, &acc, stream));
    if (sp.post_root == 2) {
      HPTB_TRY(hptb_unary(comm->ctx, HPTB_SQRT, &acc, &acc, 0, 0, stream));
    } else {
      const double third = 1.0 / 3.0;
      hptb_tensor t64;
      memset(&t64, 0, sizeof(t64));
      t64.data = s64.ptr; t64.dtype = HPTB_F64; t64.ndim = 1; t64.shape[0] = 1; t64.strides[0] = 1;
      HPTB_TRY(hptb_fill(comm->ctx, &t64, &third, stream));
      hptb_tensor tod = t64, tex = t64;
      tod.data = sod.ptr; tod.dtype = out->dtype;
      HPTB_TRY(hptb_copy(comm->ctx, &t64, &tod, stream));      // 1/3 rounded to the output dtype
      tex.data = sthird.ptr; tex.dtype = acc.dtype;
      HPTB_TRY(hptb_copy(comm->ctx, &tod, &tex, stream));      // … carried in the exchange dtype
      hptb_tensor ex = acc;
      ex.data = sthird.ptr;
      for (int i = 0; i < ex.ndim; ++i) ex.strides[i] = 0;     // broadcast the exponent over the partials
      HPTB_TRY(hptb_binary(comm->ctx, HPTB_POW, &acc, &ex, &acc, stream));
    }
    return half ? hptb_copy(comm->ctx, &acc, out, stream) : HPTB_OK;
  }
  HPTB_TRY(hptb_reduce(comm->ctx, op, shard, axes, naxes, out, 1, stream));
  return hptb_allreduce(comm, op, out, stream);
}

}  // extern "C"
