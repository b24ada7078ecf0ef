// cast_inst.cu — same-dtype strided gather (`contiguous()`, `to_cpu` of views): a bit move, so one set of
// specialised kernels per ELEMENT SIZE.  Exports `hptb_copy_same(element size) -> launcher`.  Replaces
// strided_copy_<T> (hpt-cudakernels/src/strided_copy.cu); dtype-converting copies (`astype`) take the
// runtime-typed kernel (dyn_inst.cu), whose conversions follow hpt-macros/src/scalar_convert.rs.
#include "dtypes_x.h"
#include "elementwise.cuh"
#include "ops.cuh"

namespace hptb {
namespace {
template <typename T>
struct Inst {
  static hptb_status launch(const MapPlan& plan, cudaStream_t s) {
    typedef CastFn<T, T> F;
    return launch_map<1, F, T, T, T>(plan, F{}, s);
  }
};
}  // namespace
}  // namespace hptb

extern "C" hptb::MapLauncher hptb_copy_same(int esz) {
  using namespace hptb;
  switch (esz) {
    case 1: return &Inst<uint8_t>::launch;
    case 2: return &Inst<uint16_t>::launch;
    case 4: return &Inst<uint32_t>::launch;
    case 8: return &Inst<uint64_t>::launch;
    default: return nullptr;
  }
}
