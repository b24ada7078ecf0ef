// unary_inst.cu — one translation unit per FloatUnaryOps op; compiled with
//   -DHPTB_OPENUM=<hptb_unary_op> -DHPTB_OPNAME=<name>
// and exports `hptb_unary_<op>(in dtype) -> launcher`.  13 input dtypes × 3 kernels each.
#include "dtypes_x.h"
#include "elementwise.cuh"
#include "ops.cuh"
#include "promote.h"

namespace hptb {
namespace {
template <typename A>
struct Inst {
  static constexpr int odt = promote_ct(dtype_of<A>::value, 0, HPTB_PROMOTE_FLOAT_UNARY);
  typedef typename type_of_dtype<odt>::type O;
  static hptb_status launch(const MapPlan& plan, cudaStream_t s) {
    typedef UnaryFn<HPTB_OPENUM, O, A> F;
    F f;
    f.alpha = (compute_t<O>)plan.alpha;
    f.beta = (compute_t<O>)plan.beta;
    return launch_map<1, F, O, A, A>(plan, f, s);
  }
};
}  // namespace
}  // namespace hptb

extern "C" hptb::MapLauncher HPTB_CAT(hptb_unary_, HPTB_OPNAME)(int in) {
  using namespace hptb;
  switch (in) {
#define X(T, N, E) \
  case E: return &Inst<T>::launch;
    HPTB_FOR_DTYPES(X)
#undef X
    default: return nullptr;
  }
}
