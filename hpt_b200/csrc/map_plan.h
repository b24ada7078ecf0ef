// map_plan.h — type-erased launch request handed from the C-ABI layer to an elementwise launcher.
#pragma once
#include "layout.h"

namespace hptb {

struct MapPlan {
  Collapsed c;  // operand 0 = out, 1.. = inputs
  void* ptr[3] = {nullptr, nullptr, nullptr};
  double alpha = 0.0, beta = 0.0;
  int sm_count = 148;
};

typedef hptb_status (*MapLauncher)(const MapPlan&, cudaStream_t);

}  // namespace hptb
