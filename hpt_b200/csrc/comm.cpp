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

#include "context.h"

extern "C" hptb_status hptb_reduce(hptb_ctx*, int, const hptb_tensor*, const int32_t*, int, hptb_tensor*, int, void*);
extern "C" hptb_status hptb_unary(hptb_ctx*, int, const hptb_tensor*, hptb_tensor*, double, double, void*);
namespace hptb {
hptb_status reduce_impl(hptb_ctx* ctx, int op, const hptb_tensor* in, const int32_t* axes, int naxes, hptb_tensor* out,
                        int init_out, double count_override, void* stream);  // api_reduce.cpp
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
    n.GetErrorString = (const char* (*)(int))dlsym(n.handle, "ncclGetErrorString");
    n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllReduce && n.GetErrorString;
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
};

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
  *out = c;
  return HPTB_OK;
}

hptb_status hptb_comm_destroy(hptb_comm* comm) {
  if (!comm) return HPTB_OK;
  if (comm->comm) nccl().CommDestroy(comm->comm);
  delete comm;
  return HPTB_OK;
}

hptb_status hptb_allreduce(hptb_comm* comm, int op, hptb_tensor* t, void* stream) {
  if (!comm) return fail(HPTB_ERR_INVALID, "allreduce: null comm");
  HPTB_TRY(validate_tensor(t, "allreduce tensor"));
  if (!is_contiguous(*t)) return fail(HPTB_ERR_SHAPE, "allreduce: the partial must be contiguous");
  int ndt = nccl_dtype(t->dtype);
  if (ndt < 0) return fail(HPTB_ERR_DTYPE, "allreduce: NCCL has no type for %s", dtype_name(t->dtype));
  int nop;
  switch (op) {
    case HPTB_SUM: case HPTB_SUM_SQUARE: case HPTB_MEAN: nop = t->dtype == HPTB_BOOL ? ncclMax : ncclSum; break;  // bool add = OR
    case HPTB_PROD: nop = t->dtype == HPTB_BOOL ? ncclMin : ncclProd; break;                                     // bool mul = AND
    case HPTB_MAX: nop = ncclMax; break;
    case HPTB_MIN: nop = ncclMin; break;
    default: return fail(HPTB_ERR_INVALID, "allreduce: op %d has no collective form", op);
  }
  if (comm->nranks == 1) return HPTB_OK;
  DeviceGuard g(comm->ctx->device);
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
  (void)shard_offset;
  bool crosses = false;
  for (int i = 0; i < naxes; ++i) crosses |= axes[i] == shard_axis;
  if (!crosses || comm->nranks == 1) return hptb_reduce(comm->ctx, op, shard, axes, naxes, out, 1, stream);
  if (op == HPTB_ARGMAX || op == HPTB_ARGMIN)
    return fail(HPTB_ERR_UNSUPPORTED, "reduce_sharded: arg reductions across the shard axis are not implemented yet");
  if (!is_contiguous(*out)) return fail(HPTB_ERR_SHAPE, "reduce_sharded: out must be contiguous when partials are exchanged");
  if (op == HPTB_MEAN) {
    // each rank computes Σ_local / n_GLOBAL, the allreduce-sum of those is the global mean
    double count = (double)global_axis_len;
    for (int i = 0; i < naxes; ++i)
      if (axes[i] != shard_axis) count *= (double)shard->shape[axes[i]];
    HPTB_TRY(reduce_impl(comm->ctx, HPTB_MEAN, shard, axes, naxes, out, 1, count, stream));
    return hptb_allreduce(comm, HPTB_SUM, out, stream);
  }
  if (op == HPTB_LOGSUMEXP) {
    HPTB_TRY(hptb_reduce(comm->ctx, HPTB_LOGSUMEXP, shard, axes, naxes, out, 1, stream));
    HPTB_TRY(hptb_unary(comm->ctx, HPTB_EXP, out, out, 0, 0, stream));  // back to Σ exp (naive domain, as the reference)
    HPTB_TRY(hptb_allreduce(comm, HPTB_SUM, out, stream));
    return hptb_unary(comm->ctx, HPTB_LN, out, out, 0, 0, stream);
  }
  HPTB_TRY(hptb_reduce(comm->ctx, op, shard, axes, naxes, out, 1, stream));
  return hptb_allreduce(comm, op, out, stream);
}

}  // extern "C"
