// reduce_plan.h — type-erased reduction request (host only).
#pragma once
#include "context.h"
#include "layout.h"

namespace hptb {
struct ReducePlan {
  Collapsed c;  // operand 0 = out (stride 0 on reduced dims), operand 1 = in
  const void* in = nullptr;
  void* out = nullptr;
  void* out2 = nullptr;  // second output of the fused mean/var extension (same layout as out)
  double count = 1.0;
  int fold_out = 0;
  hptb_ctx* ctx = nullptr;
};
typedef hptb_status (*ReduceLauncher)(const ReducePlan&, cudaStream_t);

// zero-initialised, self-resetting ticket counters for single-launch split reductions (one buffer per stream)
uint32_t* ctx_tickets(hptb_ctx* ctx, cudaStream_t stream, size_t n);
void ctx_tickets_destroy(hptb_ctx* ctx);
}  // namespace hptb
