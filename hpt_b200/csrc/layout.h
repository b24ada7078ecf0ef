// layout.h — the shape/stride collapse + broadcast-folding pass (host side).
//
// Generalises Layout::coalesce_dims (hpt-common/src/layout/layout_utils.rs:384-410),
// to_broadcast_layout (:73-95) and split_groups_by_axes / rearrange_array
// (hpt/src/backends/common/reduce.rs:1-96) into one pass over up to 4 operands, so that every
// elementwise or reduce call ends in one of three launch classes with all parameters passed by
// value (no device-side shape/stride tables, no H2D copies — compare cuda_utils.rs:379-399).
#pragma once
#include "common.h"

namespace hptb {

constexpr int kMaxOperands = 4;

struct Collapsed {
  int ndim = 0;
  int nops = 0;
  int launch_class = HPTB_CLASS_CONTIGUOUS;
  int64_t shape[HPTB_MAX_DIMS] = {0};
  int64_t strides[kMaxOperands][HPTB_MAX_DIMS] = {{0}};
  uint8_t reduced[HPTB_MAX_DIMS] = {0};
  int64_t numel = 1;
};

// Broadcast `t` to `shape` (ndim entries): per-dim strides with 0 on broadcast dims.
// Errors if t is not broadcastable to shape (hpt-common/src/shape/shape_utils.rs:370-400).
hptb_status broadcast_strides(const hptb_tensor& t, const int64_t* shape, int ndim, int64_t* strides_out);

hptb_status broadcast_shape(const int64_t* a, int na, const int64_t* b, int nb, int64_t* out, int* nout);

// The pass.  strides[op][dim] are element strides in the common `shape`; operand 0 is the output
// (for reductions its stride on reduced dims must be 0).  `reduced` may be null.
//  1. drop size-1 dims
//  2. order dims: kept dims by |out stride| descending, then reduced dims by |input stride|
//     descending (reductions are order-independent; elementwise order is dictated by the output)
//  3. merge neighbours i,i+1 of the same kind when stride[i] == stride[i+1]*shape[i+1] in every operand
//  4. classify (elementwise): contiguous / inner-contiguous / strided
void collapse(int ndim, const int64_t* shape, int nops, const int64_t (*strides)[HPTB_MAX_DIMS],
              const uint8_t* reduced, Collapsed* out);

}  // namespace hptb
