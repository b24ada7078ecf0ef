// map_plan.h — type-erased launch request handed from the C-ABI layer to an elementwise launcher.
#pragma once
#include "layout.h"

namespace hptb {

struct MapPlan {
  Collapsed c;  // operand 0 = out, 1.. = inputs
  void* ptr[3] = {nullptr, nullptr, nullptr};
  double alpha = 0.0, beta = 0.0;
  int sm_count = 148;
  // runtime-typed ("dyn") launchers only: element types of the inputs and the operator
  int in_dtype[2] = {-1, -1};
  int op = 0;
};

typedef hptb_status (*MapLauncher)(const MapPlan&, cudaStream_t);

// Internal status of a specialised launcher: "this layout is not one of mine" (unaligned rows, odd inner
// extents).  The API layer then takes the runtime-typed launcher, which handles every layout.
constexpr hptb_status HPTB_FALLBACK = (hptb_status)-1;

}  // namespace hptb
