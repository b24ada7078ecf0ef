// binary_inst.cu — one translation unit per binary op; compiled with
//   -DHPTB_OP=<functor> -DHPTB_OPNAME=<name> -DHPTB_KIND=<promote kind> -DHPTB_BOOL_OK=<0|1>
// and exports `hptb_binary_<op>(dtype) -> launcher` for the SAME-dtype pairs (T, T) → T: the vector-only
// specialised kernels (map_rows_kernel with 128-bit accesses, map_tiled_kernel for permuted operands).
// Mixed-dtype pairs, pairs whose Output differs from T (int / int → float for div) and unaligned layouts are
// served by the runtime-typed kernel (dyn_inst.cu).
#include "dtypes_x.h"
#include "elementwise.cuh"
#include "ops.cuh"
#include "promote.h"

namespace hptb {
namespace {
template <typename T>
struct Inst {
  static constexpr int odt = promote_ct(dtype_of<T>::value, dtype_of<T>::value, HPTB_KIND);
  static hptb_status launch(const MapPlan& plan, cudaStream_t s) {
    typedef BinaryFn<HPTB_OP, T, T, T> F;
    return launch_map<2, F, T, T, T>(plan, F{}, s);
  }
  static MapLauncher get() {
    if constexpr (odt != dtype_of<T>::value) return nullptr;
    else if constexpr (odt == HPTB_BOOL && !HPTB_BOOL_OK) return nullptr;
    else return &launch;
  }
};
}  // namespace
}  // namespace hptb

extern "C" hptb::MapLauncher HPTB_CAT(hptb_binary_, HPTB_OPNAME)(int dt) {
  using namespace hptb;
  switch (dt) {
#define X(T, N, E) \
  case E: return Inst<T>::get();
    HPTB_FOR_DTYPES(X)
#undef X
    default: return nullptr;
  }
}
