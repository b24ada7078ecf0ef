// binary_inst.cu — one translation unit per (binary op, lhs dtype); compiled with
//   -DHPTB_OP=<functor> -DHPTB_OPNAME=<name> -DHPTB_KIND=<promote kind> -DHPTB_BOOL_OK=<0|1>
//   -DHPTB_LHS=<c++ type> -DHPTB_LHSNAME=<short name>
// and exports `hptb_binary_<op>_<lhs>(rhs dtype) -> launcher`.  13 rhs dtypes × 3 kernels each.
#include "dtypes_x.h"
#include "elementwise.cuh"
#include "ops.cuh"
#include "promote.h"

namespace hptb {
namespace {
template <typename R>
struct Inst {
  typedef HPTB_LHS L;
  static constexpr int odt = promote_ct(dtype_of<L>::value, dtype_of<R>::value, HPTB_KIND);
  typedef typename type_of_dtype<odt>::type O;
  static hptb_status launch(const MapPlan& plan, cudaStream_t s) {
    typedef BinaryFn<HPTB_OP, O, L, R> F;
    return launch_map<2, F, O, L, R>(plan, F{}, s);
  }
  static MapLauncher get() {
    if constexpr (odt == HPTB_BOOL && !HPTB_BOOL_OK) return nullptr;
    else return &launch;
  }
};
}  // namespace
}  // namespace hptb

extern "C" hptb::MapLauncher HPTB_CAT4(hptb_binary_, HPTB_OPNAME, _, HPTB_LHSNAME)(int rhs) {
  using namespace hptb;
  switch (rhs) {
#define X(T, N, E) \
  case E: return Inst<T>::get();
    HPTB_FOR_DTYPES(X)
#undef X
    default: return nullptr;
  }
}
