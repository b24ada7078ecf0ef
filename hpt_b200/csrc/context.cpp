#include "context.h"
#include "reduce_plan.h"

#include <atomic>
#include <cstdlib>
#include <new>

namespace hptb {

uint64_t launches();
static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
hptb_status fail(hptb_status st, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return st;
}
const char* last_error() { return g_err; }

static std::atomic<uint64_t> g_launches{0};
void count_launches(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
uint64_t launches() { return g_launches.load(std::memory_order_relaxed); }

const char* dtype_name(int dt) {
  static const char* names[] = {"bool", "i8", "i16", "i32", "i64", "u8", "u16", "u32", "u64", "f16", "bf16", "f32", "f64"};
  return dtype_valid(dt) ? names[dt] : "invalid";
}

hptb_status validate_tensor(const hptb_tensor* t, const char* what) {
  if (!t) return fail(HPTB_ERR_INVALID, "%s: null tensor", what);
  if (!dtype_valid(t->dtype)) return fail(HPTB_ERR_INVALID, "%s: bad dtype %d", what, t->dtype);
  if (t->ndim < 0 || t->ndim > HPTB_MAX_DIMS) return fail(HPTB_ERR_INVALID, "%s: ndim %d out of range", what, t->ndim);
  int64_t n = 1;
  for (int i = 0; i < t->ndim; ++i) {
    if (t->shape[i] < 0) return fail(HPTB_ERR_SHAPE, "%s: negative extent at dim %d", what, i);
    n *= t->shape[i];
  }
  if (n > 0 && !t->data) return fail(HPTB_ERR_INVALID, "%s: null data pointer", what);
  return HPTB_OK;
}

bool pass_direction(hptb_ctx* ctx, const void* in, size_t bytes, bool can_reverse) {
  static const bool off = [] { const char* e = getenv("HPTB_NO_SNAKE"); return e && e[0] == '1'; }();
  if (!ctx) return false;
  const void* prev = ctx->last_pass_in.exchange(in, std::memory_order_relaxed);
  bool rev = false;
  // only worth it when the tensor does not fit L2 anyway (a forward re-read of a resident tensor already hits)
  if (!off && can_reverse && prev == in && bytes >= (size_t)96 << 20) rev = !ctx->last_pass_rev.load(std::memory_order_relaxed);
  ctx->last_pass_rev.store(rev ? 1 : 0, std::memory_order_relaxed);
  return rev;
}

DeviceGuard::DeviceGuard(int dev) {
  if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
  if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
}
DeviceGuard::~DeviceGuard() {
  int cur = -1;
  if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
}

hptb_status Scratch::get(hptb_ctx* c, size_t bytes, void* s) {
  ctx = c;
  stream = s;
  int rc = c->alloc->allocate(bytes, s, &ptr);
  if (rc == 2) return fail(HPTB_ERR_OOM, "scratch allocation of %zu bytes failed", bytes);
  if (rc) return fail(HPTB_ERR_CUDA, "scratch allocation failed (cuda error %d)", rc);
  return HPTB_OK;
}
Scratch::~Scratch() {
  if (ctx && ptr) ctx->alloc->release(ptr, stream);
}

namespace {
struct CudaApi : DeviceApi {
  int device;
  explicit CudaApi(int d) : device(d) {}
  int malloc(void** p, size_t bytes) override {
    DeviceGuard g(device);
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return 2; }
    return e == cudaSuccess ? 0 : (int)e;
  }
  int free(void* p) override { DeviceGuard g(device); return (int)cudaFree(p); }
  int event_create(void** ev) override {
    DeviceGuard g(device);
    cudaEvent_t e;
    cudaError_t rc = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    *ev = e;
    return (int)rc;
  }
  int event_destroy(void* ev) override { return (int)cudaEventDestroy((cudaEvent_t)ev); }
  int event_record(void* ev, void* stream) override {
    DeviceGuard g(device);
    return (int)cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)stream);
  }
  int event_done(void* ev, bool* done) override {
    if (!ev) { *done = false; return 0; }
    cudaError_t e = cudaEventQuery((cudaEvent_t)ev);
    *done = e == cudaSuccess;
    if (e == cudaErrorNotReady) { cudaGetLastError(); return 0; }
    return e == cudaSuccess ? 0 : (int)e;
  }
  int device_sync() override { DeviceGuard g(device); return (int)cudaDeviceSynchronize(); }
};

// fake device for hptb_alloc_selftest: counts calls, events complete only when "synchronised"
struct FakeApi : DeviceApi {
  uintptr_t next = 0x10000;
  size_t cap, used = 0;
  int mallocs = 0, frees = 0, syncs = 0;
  std::unordered_map<void*, size_t> sizes;
  std::unordered_map<void*, bool> ev_done;
  uintptr_t next_ev = 1;
  explicit FakeApi(size_t c) : cap(c) {}
  int malloc(void** p, size_t bytes) override {
    if (used + bytes > cap) return 2;
    *p = (void*)next; next += (bytes + 0xfff) & ~size_t(0xfff);
    sizes[*p] = bytes; used += bytes; ++mallocs;
    return 0;
  }
  int free(void* p) override { used -= sizes[p]; sizes.erase(p); ++frees; return 0; }
  int event_create(void** ev) override { *ev = (void*)(next_ev++); ev_done[*ev] = false; return 0; }
  int event_destroy(void* ev) override { ev_done.erase(ev); return 0; }
  int event_record(void* ev, void*) override { ev_done[ev] = false; return 0; }
  int event_done(void* ev, bool* done) override { *done = ev && ev_done[ev]; return 0; }
  int device_sync() override { for (auto& kv : ev_done) kv.second = true; ++syncs; return 0; }
};
}  // namespace
}  // namespace hptb

using namespace hptb;

extern "C" {

int hptb_version(void) { return HPTB_VERSION; }
uint64_t hptb_kernel_launches(void) { return launches(); }
const char* hptb_last_error(void) { return last_error(); }
size_t hptb_dtype_size(int dtype) { return dtype_size(dtype); }
const char* hptb_dtype_name(int dtype) { return dtype_name(dtype); }

hptb_status hptb_ctx_create(int device, hptb_ctx** out) {
  if (!out) return fail(HPTB_ERR_INVALID, "ctx_create: null out");
  int ndev = 0;
  HPTB_CUDA_CHECK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(HPTB_ERR_INVALID, "ctx_create: device %d out of range (%d visible)", device, ndev);
  DeviceGuard g(device);
  cudaDeviceProp prop;
  HPTB_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(HPTB_ERR_UNSUPPORTED, "device %d is sm_%d%d; libhpt_b200 is built for sm_100a only", device, prop.major, prop.minor);
  hptb_ctx* c = new (std::nothrow) hptb_ctx();
  if (!c) return fail(HPTB_ERR_OOM, "ctx_create: host allocation failed");
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  c->api = new CudaApi(device);
  c->alloc = new CachingAllocator(c->api);
  *out = c;
  return HPTB_OK;
}

hptb_status hptb_ctx_destroy(hptb_ctx* ctx) {
  if (!ctx) return HPTB_OK;
  { DeviceGuard g(ctx->device); ctx_tickets_destroy(ctx); }
  delete ctx->alloc;
  delete ctx->api;
  delete ctx;
  return HPTB_OK;
}

hptb_status hptb_ctx_device(const hptb_ctx* ctx, int* device) {
  if (!ctx || !device) return fail(HPTB_ERR_INVALID, "ctx_device: null argument");
  *device = ctx->device;
  return HPTB_OK;
}
hptb_status hptb_ctx_sm_count(const hptb_ctx* ctx, int* sms) {
  if (!ctx || !sms) return fail(HPTB_ERR_INVALID, "ctx_sm_count: null argument");
  *sms = ctx->sm_count;
  return HPTB_OK;
}
hptb_status hptb_stream_sync(hptb_ctx* ctx, void* stream) {
  if (!ctx) return fail(HPTB_ERR_INVALID, "stream_sync: null ctx");
  DeviceGuard g(ctx->device);
  HPTB_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  return HPTB_OK;
}

hptb_status hptb_alloc(hptb_ctx* ctx, size_t bytes, void** ptr, void* stream) {
  if (!ctx || !ptr) return fail(HPTB_ERR_INVALID, "alloc: null argument");
  int rc = ctx->alloc->allocate(bytes, stream, ptr);
  if (rc == 2) return fail(HPTB_ERR_OOM, "device allocation of %zu bytes failed (after emptying the cache)", bytes);
  if (rc) return fail(HPTB_ERR_CUDA, "device allocation failed: %s (%d)", cudaGetErrorString((cudaError_t)rc), rc);
  return HPTB_OK;
}
hptb_status hptb_free(hptb_ctx* ctx, void* ptr, void* stream) {
  if (!ctx) return fail(HPTB_ERR_INVALID, "free: null ctx");
  if (ctx->alloc->release(ptr, stream)) return fail(HPTB_ERR_INVALID, "free: pointer %p was not allocated by this context", ptr);
  return HPTB_OK;
}
hptb_status hptb_record_stream(hptb_ctx* ctx, void* ptr, void* stream) {
  if (!ctx) return fail(HPTB_ERR_INVALID, "record_stream: null ctx");
  if (ctx->alloc->record_stream(ptr, stream)) return fail(HPTB_ERR_INVALID, "record_stream: pointer %p is not a live allocation of this context", ptr);
  return HPTB_OK;
}
hptb_status hptb_empty_cache(hptb_ctx* ctx) {
  if (!ctx) return fail(HPTB_ERR_INVALID, "empty_cache: null ctx");
  ctx->alloc->empty_cache();
  return HPTB_OK;
}
hptb_status hptb_alloc_get_stats(hptb_ctx* ctx, hptb_alloc_stats* out) {
  if (!ctx || !out) return fail(HPTB_ERR_INVALID, "alloc_get_stats: null argument");
  AllocStats s = ctx->alloc->stats();
  out->bytes_in_use = s.bytes_in_use; out->bytes_cached = s.bytes_cached; out->bytes_reserved_peak = s.bytes_reserved_peak;
  out->n_alloc = s.n_alloc; out->n_cache_hit = s.n_cache_hit; out->n_device_malloc = s.n_device_malloc;
  out->n_device_free = s.n_device_free;
  return HPTB_OK;
}

#define SELFTEST(cond)                                                                   \
  do {                                                                                   \
    if (!(cond)) return fail(HPTB_ERR_INVALID, "alloc selftest failed: %s (line %d)", #cond, __LINE__); \
  } while (0)

hptb_status hptb_alloc_selftest(void) {
  FakeApi api(size_t(64) << 20);
  {
    CachingAllocator a(&api);
    void *s1 = (void*)1, *s2 = (void*)2;
    void *p1, *p2, *p3;
    SELFTEST(CachingAllocator::round_size(1) == 512);
    SELFTEST(CachingAllocator::round_size(513) == 1024);
    SELFTEST(CachingAllocator::round_size((1 << 20) + 1) == (size_t(2) << 20));
    SELFTEST(a.allocate(1000, s1, &p1) == 0 && api.mallocs == 1);
    SELFTEST(a.release(p1, s1) == 0);
    // same stream: reuse at once, no device call, no wait
    SELFTEST(a.allocate(900, s1, &p2) == 0 && p2 == p1 && api.mallocs == 1);
    SELFTEST(a.release(p2, s1) == 0);
    // other stream: the event has not completed → must not reuse, falls to a device malloc
    SELFTEST(a.allocate(1000, s2, &p3) == 0 && p3 != p1 && api.mallocs == 2);
    api.device_sync();  // events complete
    void* p4;
    SELFTEST(a.allocate(1000, s2, &p4) == 0 && p4 == p1 && api.mallocs == 2);
    SELFTEST(a.release((void*)0xdead, s1) == 1);
    AllocStats st = a.stats();
    SELFTEST(st.n_alloc == 4 && st.n_cache_hit == 2 && st.n_device_malloc == 2);
    SELFTEST(st.bytes_in_use == 2048 && st.bytes_cached == 0);
    a.release(p3, s2);
    a.release(p4, s2);
    api.device_sync();
    // cross-stream use: allocated and freed on s1, consumed on s2 in between (record_stream) → no immediate same-stream
    // reuse while s2's work is pending; reuse once it has completed
    void *q1, *q2, *q3;
    SELFTEST(a.allocate(5000, s1, &q1) == 0);
    SELFTEST(a.record_stream(q1, s2) == 0 && a.record_stream(q1, s2) == 0);
    SELFTEST(a.record_stream((void*)0xdead, s2) == 1);
    SELFTEST(a.release(q1, s1) == 0);
    SELFTEST(a.allocate(5000, s1, &q2) == 0 && q2 != q1);  // s2 may still be reading q1
    api.device_sync();
    SELFTEST(a.allocate(5000, s1, &q3) == 0 && q3 == q1);  // both events completed
    // … and a block used on ONE stream only keeps its immediate same-stream reuse
    a.release(q3, s1);
    void* q4;
    SELFTEST(a.allocate(5000, s1, &q4) == 0 && q4 == q1);
    a.release(q2, s1);
    a.release(q4, s1);
    // OOM: cache is emptied and the allocation retried
    void* big;
    SELFTEST(a.allocate(size_t(63) << 20, s1, &big) == 0);
    SELFTEST(api.frees == 4);  // the four cached blocks (two sizes) went back to the device
    void* too_big;
    SELFTEST(a.allocate(size_t(32) << 20, s1, &too_big) == 2);
    a.release(big, s1);
    SELFTEST(a.stats().bytes_cached == (size_t(64) << 20));
    a.empty_cache();
    SELFTEST(a.stats().bytes_cached == 0 && api.used == 0);
  }
  // FastDiv exhaustive-ish check
  for (uint32_t d : {1u, 2u, 3u, 5u, 7u, 10u, 56u, 64u, 100u, 392u, 3136u, 4096u, 65535u, 65536u, 1000003u, 0x7fffffffu}) {
    FastDiv fd(d);
    const uint64_t probes[] = {0ull, 1ull, 2ull, (uint64_t)d - 1, (uint64_t)d, (uint64_t)d + 1, 2ull * d - 1, 2ull * d, 123456789ull,
                               0x7fffffffull, 0x80000000ull, 0xfffffffeull, 0xffffffffull};
    for (uint64_t n : probes) {
      if (n > 0xffffffffull) continue;
      SELFTEST(fd.div((uint32_t)n) == (uint32_t)(n / d));
    }
    for (uint32_t n = 0; n < 200000; n += 7) SELFTEST(fd.div(n) == n / d);
  }
  return HPTB_OK;
}

hptb_status hptb_memcpy_h2d(hptb_ctx* ctx, void* dst, const void* src, size_t bytes, void* stream) {
  if (!ctx) return fail(HPTB_ERR_INVALID, "memcpy_h2d: null ctx");
  DeviceGuard g(ctx->device);
  HPTB_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return HPTB_OK;
}
hptb_status hptb_memcpy_d2h(hptb_ctx* ctx, void* dst, const void* src, size_t bytes, void* stream) {
  if (!ctx) return fail(HPTB_ERR_INVALID, "memcpy_d2h: null ctx");
  DeviceGuard g(ctx->device);
  HPTB_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  HPTB_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  return HPTB_OK;
}
hptb_status hptb_memcpy_d2h_async(hptb_ctx* ctx, void* dst, const void* src, size_t bytes, void* stream) {
  if (!ctx) return fail(HPTB_ERR_INVALID, "memcpy_d2h_async: null ctx");
  DeviceGuard g(ctx->device);
  HPTB_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return HPTB_OK;
}
hptb_status hptb_stream_wait_stream(hptb_ctx* ctx, void* stream, void* other) {
  if (!ctx) return fail(HPTB_ERR_INVALID, "stream_wait_stream: null ctx");
  DeviceGuard g(ctx->device);
  cudaEvent_t ev;
  HPTB_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  cudaError_t e = cudaEventRecord(ev, (cudaStream_t)other);
  if (e == cudaSuccess) e = cudaStreamWaitEvent((cudaStream_t)stream, ev, 0);
  cudaEventDestroy(ev);  // released by the runtime once the wait has been satisfied
  if (e != cudaSuccess) return fail(HPTB_ERR_CUDA, "stream_wait_stream failed: %s", cudaGetErrorString(e));
  return HPTB_OK;
}
hptb_status hptb_memcpy_d2d(hptb_ctx* ctx, void* dst, const void* src, size_t bytes, void* stream) {
  if (!ctx) return fail(HPTB_ERR_INVALID, "memcpy_d2d: null ctx");
  DeviceGuard g(ctx->device);
  HPTB_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return HPTB_OK;
}
hptb_status hptb_host_alloc_pinned(size_t bytes, void** ptr) {
  if (!ptr) return fail(HPTB_ERR_INVALID, "host_alloc_pinned: null out");
  HPTB_CUDA_CHECK(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
  return HPTB_OK;
}
hptb_status hptb_host_free_pinned(void* ptr) {
  HPTB_CUDA_CHECK(cudaFreeHost(ptr));
  return HPTB_OK;
}

}  // extern "C"
