#include "allocator.h"

namespace hptb {

size_t CachingAllocator::round_size(size_t bytes) {
  if (bytes == 0) bytes = 1;
  const size_t small = 512, large = size_t(2) << 20;
  if (bytes < (size_t(1) << 20)) return (bytes + small - 1) / small * small;
  return (bytes + large - 1) / large * large;
}

CachingAllocator::~CachingAllocator() {
  empty_cache();
  for (void* ev : event_pool_) api_->event_destroy(ev);
}

void* CachingAllocator::take_event() {
  void* ev = nullptr;
  if (!event_pool_.empty()) { ev = event_pool_.back(); event_pool_.pop_back(); }
  else if (api_->event_create(&ev) != 0) ev = nullptr;
  return ev;
}

bool CachingAllocator::events_done(const Block& b, bool include_own) {
  bool done = false;
  if (include_own && !(b.event && api_->event_done(b.event, &done) == 0 && done)) return false;
  for (void* ev : b.user_evs) {
    done = false;
    if (!(ev && api_->event_done(ev, &done) == 0 && done)) return false;
  }
  return true;
}

int CachingAllocator::allocate(size_t bytes, void* stream, void** out) {
  size_t sz = round_size(bytes);
  std::lock_guard<std::mutex> g(mu_);
  st_.n_alloc++;
  // 1. same size class: prefer a block last used on this stream, else one whose event completed
  auto range = free_.equal_range(sz);
  auto pick = free_.end();
  for (auto it = range.first; it != range.second; ++it) {
    // stream order covers the free stream only: work recorded from other streams must have completed
    if (it->second.stream == stream && events_done(it->second, false)) { pick = it; break; }
  }
  if (pick == free_.end()) {
    for (auto it = range.first; it != range.second; ++it) {
      if (events_done(it->second, true)) { pick = it; break; }
    }
  }
  if (pick != free_.end()) {
    Block b = pick->second;
    free_.erase(pick);
    if (b.event) { event_pool_.push_back(b.event); b.event = nullptr; }
    for (void* ev : b.user_evs) event_pool_.push_back(ev);
    b.user_evs.clear();
    b.users.clear();
    b.stream = stream;
    live_[b.ptr] = b;
    st_.n_cache_hit++;
    st_.bytes_cached -= b.size;
    st_.bytes_in_use += b.size;
    *out = b.ptr;
    return 0;
  }
  // 2. miss: device malloc, retry once after emptying the cache
  void* p = nullptr;
  int rc = api_->malloc(&p, sz);
  if (rc == 2) {
    empty_cache_locked();
    rc = api_->malloc(&p, sz);
  }
  if (rc != 0) return rc;
  st_.n_device_malloc++;
  Block b{p, sz, stream, nullptr, {}, {}};
  live_[p] = b;
  st_.bytes_in_use += sz;
  uint64_t reserved = st_.bytes_in_use + st_.bytes_cached;
  if (reserved > st_.bytes_reserved_peak) st_.bytes_reserved_peak = reserved;
  *out = p;
  return 0;
}

int CachingAllocator::release(void* ptr, void* stream) {
  if (!ptr) return 0;
  std::lock_guard<std::mutex> g(mu_);
  auto it = live_.find(ptr);
  if (it == live_.end()) return 1;
  Block b = it->second;
  live_.erase(it);
  void* ev = take_event();
  b.stream = stream;
  b.event = ev;
  if (ev) api_->event_record(ev, stream);
  for (void* us : b.users) {
    if (us == stream) continue;
    void* uev = take_event();
    if (uev) api_->event_record(uev, us);
    b.user_evs.push_back(uev);  // a null event never reads as done: the block then waits for empty_cache's device sync
  }
  b.users.clear();
  st_.bytes_in_use -= b.size;
  st_.bytes_cached += b.size;
  free_.emplace(b.size, b);
  return 0;
}

int CachingAllocator::record_stream(void* ptr, void* stream) {
  if (!ptr) return 0;
  std::lock_guard<std::mutex> g(mu_);
  auto it = live_.find(ptr);
  if (it == live_.end()) return 1;
  Block& b = it->second;
  for (void* us : b.users)
    if (us == stream) return 0;
  b.users.push_back(stream);
  return 0;
}

int CachingAllocator::empty_cache() {
  std::lock_guard<std::mutex> g(mu_);
  return empty_cache_locked();
}

int CachingAllocator::empty_cache_locked() {
  if (free_.empty()) return 0;
  api_->device_sync();  // cached blocks may still be referenced by queued work
  for (auto& kv : free_) {
    api_->free(kv.second.ptr);
    st_.n_device_free++;
    st_.bytes_cached -= kv.second.size;
    if (kv.second.event) event_pool_.push_back(kv.second.event);
    for (void* ev : kv.second.user_evs)
      if (ev) event_pool_.push_back(ev);
  }
  free_.clear();
  return 0;
}

AllocStats CachingAllocator::stats() {
  std::lock_guard<std::mutex> g(mu_);
  return st_;
}

}  // namespace hptb
