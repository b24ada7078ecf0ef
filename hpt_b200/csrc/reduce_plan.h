// reduce_plan.h — type-erased reduction request (host only).
#pragma once
#include "context.h"
#include "layout.h"
#include "xchg.cuh"

namespace hptb {
constexpr int kPlanReduce = 0;   // reduce `in` into `out`
constexpr int kPlanCombine = 1;  // sharded: exchange / combine bare accumulators, apply post, write `out` (reduce.cuh)
struct ReducePlan {
  Collapsed c;  // operand 0 = out (stride 0 on reduced dims), operand 1 = in
  const void* in = nullptr;
  void* out = nullptr;
  void* out2 = nullptr;  // second output of the fused mean/var extension (same layout as out)
  double count = 1.0;
  int fold_out = 0;
  int reverse = 0;  // snake order: kernels that can, walk their outputs backwards (context.h pass_direction)
  hptb_ctx* ctx = nullptr;
  // ---- sharded reductions (comm.cpp): the exchange of per-rank accumulators, fused into the kernel where the launch
  // shape allows (xchg.cuh) -----------------------------------------------------------------------------------------
  int mode = kPlanReduce;
  const XchgParams* xchg = nullptr;  // peer mailboxes of this call; nullptr = no exchange
  void* raw_out = nullptr;           // accumulator scratch addressed with `out`'s element offsets (unfused shapes)
  bool* fused = nullptr;             // result: true = the kernel exchanged and wrote `out`; false = accumulators in raw_out
  size_t* acc_bytes = nullptr;       // result: sizeof(accumulator) of this (op, dtype)
  // kPlanCombine: `in` = this rank's comb_M accumulators (gathered == 0: exchanged through the mailboxes) or every
  // rank's, [gathered][comb_M] (NCCL all-gather); `out` dims innermost first
  int gathered = 0;
  int64_t comb_M = 0;
  int comb_nk = 0;
  int64_t comb_shape[HPTB_MAX_DIMS] = {0}, comb_stride[HPTB_MAX_DIMS] = {0};
};
typedef hptb_status (*ReduceLauncher)(const ReducePlan&, cudaStream_t);

// elementwise → reduce fusion (fused_inst.cu)
struct FusedPlan {
  Collapsed c;  // operand 0 = out (stride 0 on reduced dims), 1 = lhs, 2 = rhs (broadcast strides)
  const void* lhs = nullptr;
  const void* rhs = nullptr;
  void* out = nullptr;
  double count = 1.0;
  int fold_out = 0;
  int red_op = 0;
  hptb_ctx* ctx = nullptr;
};
typedef hptb_status (*FusedLauncher)(const FusedPlan&, cudaStream_t);

// zero-initialised, self-resetting ticket counters for single-launch split reductions (one buffer per stream)
uint32_t* ctx_tickets(hptb_ctx* ctx, cudaStream_t stream, size_t n);
void ctx_tickets_destroy(hptb_ctx* ctx);
}  // namespace hptb
