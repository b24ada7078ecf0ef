// context.h — per-device context: device properties, the caching allocator, reduction scratch.
#pragma once
#include <atomic>

#include "allocator.h"
#include "common.h"

struct hptb_ctx {
  int device = 0;
  int sm_count = 148;
  int max_smem_optin = 0;
  hptb::DeviceApi* api = nullptr;
  hptb::CachingAllocator* alloc = nullptr;
  // Snake-order passes (pass_direction below): base pointer and direction of the last streaming pass launched
  // through this context.
  std::atomic<const void*> last_pass_in{nullptr};
  std::atomic<int> last_pass_rev{0};
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
// Snake order.  A pass over a tensor larger than the 126 MB L2 leaves its TAIL resident; when the next launch
// streams the same tensor again (config 2: max then argmax of one view; mean then sum_square; …) it starts where
// the previous one ended, so those lines come from L2 instead of HBM (measured: max + argmax of f32 [8192,8192]
// 80.0 → 67.9 µs, tools/experiments/snake_order.py).  Kernels that can walk their outputs backwards (the lean row
// reductions) ask for the direction; every other streaming launch just records its input.  Results do not depend
// on the direction.  HPTB_NO_SNAKE=1 disables it.
bool pass_direction(hptb_ctx* ctx, const void* in, size_t bytes, bool can_reverse);
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev);
  ~DeviceGuard();
};
}  // namespace hptb
