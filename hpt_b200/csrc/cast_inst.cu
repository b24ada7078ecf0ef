// cast_inst.cu — strided gather / dtype conversion (`contiguous()`, `to_cpu` of views, `astype`);
// one translation unit per source dtype (-DHPTB_LHS, -DHPTB_LHSNAME), exports
// `hptb_cast_<src>(dst dtype) -> launcher`.  Replaces strided_copy_<T>
// (hpt-cudakernels/src/strided_copy.cu); conversions follow hpt-macros/src/scalar_convert.rs.
#include "dtypes_x.h"
#include "elementwise.cuh"
#include "ops.cuh"

namespace hptb {
namespace {
template <typename O>
struct Inst {
  typedef HPTB_LHS A;
  static hptb_status launch(const MapPlan& plan, cudaStream_t s) {
    typedef CastFn<O, A> F;
    return launch_map<1, F, O, A, A>(plan, F{}, s);
  }
};
}  // namespace
}  // namespace hptb

extern "C" hptb::MapLauncher HPTB_CAT(hptb_cast_, HPTB_LHSNAME)(int dst) {
  using namespace hptb;
  switch (dst) {
#define X(T, N, E) \
  case E: return &Inst<T>::launch;
    HPTB_FOR_DTYPES(X)
#undef X
    default: return nullptr;
  }
}
