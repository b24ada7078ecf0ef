// context.h — per-device context: device properties, the caching allocator, reduction scratch.
#pragma once
#include "allocator.h"
#include "common.h"

struct hptb_ctx {
  int device = 0;
  int sm_count = 148;
  int max_smem_optin = 0;
  hptb::DeviceApi* api = nullptr;
  hptb::CachingAllocator* alloc = nullptr;
  // Zero-initialised ticket counters for single-launch multi-block reductions.  The last block of
  // each reduction resets its ticket to 0, so one stream-ordered buffer serves every launch on a
  // stream; launches on different streams get distinct buffers from the pool.
};

namespace hptb {
// number of device kernels launched by this library in this process (evidence for bench.py's gpu_launches)
void count_launches(int n);
// RAII device scratch from the context's pool, returned to the pool on the same stream.
struct Scratch {
  hptb_ctx* ctx = nullptr;
  void* ptr = nullptr;
  void* stream = nullptr;
  hptb_status get(hptb_ctx* c, size_t bytes, void* s);
  ~Scratch();
};
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev);
  ~DeviceGuard();
};
}  // namespace hptb
