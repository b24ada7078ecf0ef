// unary_inst.cu — one translation unit per FloatUnaryOps op; compiled with
//   -DHPTB_OPENUM=<hptb_unary_op> -DHPTB_OPNAME=<name>
// and exports `hptb_unary_<op>(in dtype) -> launcher` for the float dtypes (T → T): the vector-only specialised
// kernels.  Integer / bool inputs (→ f16/f32/f64 by FloatOutUnaryPromote) and unaligned layouts take the
// runtime-typed kernel (dyn_inst.cu).
#include "dtypes_x.h"
#include "elementwise.cuh"
#include "ops.cuh"
#include "promote.h"

namespace hptb {
namespace {
template <typename T>
struct Inst {
  static hptb_status launch(const MapPlan& plan, cudaStream_t s) {
    typedef UnaryFn<HPTB_OPENUM, T, T> F;
    F f;
    f.alpha = (compute_t<T>)plan.alpha;
    f.beta = (compute_t<T>)plan.beta;
    return launch_map<1, F, T, T, T>(plan, f, s);
  }
};
}  // namespace
}  // namespace hptb

extern "C" hptb::MapLauncher HPTB_CAT(hptb_unary_, HPTB_OPNAME)(int in) {
  using namespace hptb;
  switch (in) {
    case HPTB_F16: return &Inst<f16>::launch;
    case HPTB_BF16: return &Inst<bf16>::launch;
    case HPTB_F32: return &Inst<float>::launch;
    case HPTB_F64: return &Inst<double>::launch;
    default: return nullptr;
  }
}
