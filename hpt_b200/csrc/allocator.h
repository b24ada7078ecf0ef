// allocator.h — stream-ordered caching device allocator.
//
// Replaces HptAllocator<Cuda> (hpt-allocator/src/allocators/cuda.rs:41-199): the reference caches
// freed blocks in an LRU keyed by the exact `Layout{size,align}` (utils/allocate.rs:66-124), holds one
// size class by default (cuda.rs:26), calls cuMemAlloc on every miss and is only safe because every
// launch sits on a single stream and frees are host-synchronous.  Here:
//   * requests are rounded to size classes (512 B granules below 1 MiB, 2 MiB granules above), so
//     near-equal shapes reuse each other's blocks;
//   * a freed block carries the stream it was freed on and an event recorded there at free time, plus one
//     event per OTHER stream that consumed it while it was live (record_stream — a tensor uploaded on a copy
//     stream, read by kernels on a compute stream, dropped by the host right after enqueueing).  It is handed
//     out again immediately on the free stream when no other stream used it (stream order makes that safe),
//     and otherwise — or on another stream — once every one of its events has completed.  No host
//     synchronisation on the hot path;
//   * on device OOM the cache is emptied and the allocation retried once.
// The device is reached through a small virtual interface so the caching logic is unit-tested on a
// fake device without a GPU (hptb_alloc_selftest).
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace hptb {

struct DeviceApi {
  virtual ~DeviceApi() {}
  virtual int malloc(void** p, size_t bytes) = 0;      // 0 ok, 2 = out of memory, other = error
  virtual int free(void* p) = 0;
  virtual int event_create(void** ev) = 0;
  virtual int event_destroy(void* ev) = 0;
  virtual int event_record(void* ev, void* stream) = 0;
  virtual int event_done(void* ev, bool* done) = 0;
  virtual int device_sync() = 0;
};

struct AllocStats {
  uint64_t bytes_in_use = 0, bytes_cached = 0, bytes_reserved_peak = 0;
  uint64_t n_alloc = 0, n_cache_hit = 0, n_device_malloc = 0, n_device_free = 0;
};

class CachingAllocator {
 public:
  explicit CachingAllocator(DeviceApi* api) : api_(api) {}
  ~CachingAllocator();
  // returns 0 ok, 2 OOM, else device error code
  int allocate(size_t bytes, void* stream, void** out);
  int release(void* ptr, void* stream);  // returns 1 if ptr is unknown
  // `ptr` (a live block) is being used by work enqueued on `stream`, which is not the stream it will be freed on:
  // the block is not reused until that work has completed.  Returns 1 if ptr is unknown.
  int record_stream(void* ptr, void* stream);
  int empty_cache();
  AllocStats stats();
  static size_t round_size(size_t bytes);

 private:
  struct Block {
    void* ptr;
    size_t size;
    void* stream;
    void* event;                  // recorded on `stream` at release; null while in use
    std::vector<void*> users;     // live: other streams that consumed the block (record_stream)
    std::vector<void*> user_evs;  // cached: one event per such stream, recorded at release
  };
  void* take_event();
  bool events_done(const Block& b, bool include_own);
  int empty_cache_locked();
  DeviceApi* api_;
  std::mutex mu_;
  std::multimap<size_t, Block> free_;             // size class → cached blocks
  std::unordered_map<void*, Block> live_;         // blocks handed out
  std::vector<void*> event_pool_;
  AllocStats st_;
};

}  // namespace hptb
