// binary_inst.cu — one translation unit per binary op; compiled with
//   -DHPTB_OP=<functor> -DHPTB_OPNAME=<name> -DHPTB_KIND=<promote kind> -DHPTB_BOOL_OK=<0|1>
// and exports `hptb_binary_<op>(dtype) -> launcher` for the SAME-dtype pairs (T, T) → T: the vector-only
// specialised kernels (map_rows_kernel with 128-bit accesses, map_tiled_kernel for permuted operands).
// Other mixed-dtype pairs, pairs whose Output differs from T (int / int → float for div) and unaligned layouts
// are served by the runtime-typed kernel (dyn_inst.cu); `hptb_binary_<op>_mixed(lhs, rhs)` covers a short list.
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

// A short list of mixed-dtype pairs that get their own specialised kernels as well (the pairs a model actually
// produces: f32 against integer indices / f64 / half weights, BASELINE config 4's f32 ⊕ i64 → f64); every
// other pair runs on the runtime-typed kernel.
template <typename L, typename R>
struct Mixed {
  static constexpr int odt = promote_ct(dtype_of<L>::value, dtype_of<R>::value, HPTB_KIND);
  typedef typename type_of_dtype<odt>::type O;
  static hptb_status launch(const MapPlan& plan, cudaStream_t s) {
    typedef BinaryFn<HPTB_OP, O, L, R> F;
    return launch_map<2, F, O, L, R>(plan, F{}, s);
  }
};
#define HPTB_MIXED_PAIRS(X)                                                                       \
  X(float, int64_t) X(int64_t, float) X(float, int32_t) X(int32_t, float) X(float, double) X(double, float) \
  X(f16, float) X(float, f16) X(bf16, float) X(float, bf16) X(int32_t, int64_t) X(int64_t, int32_t)       \
  X(double, int64_t) X(int64_t, double)
}  // namespace
}  // namespace hptb

extern "C" hptb::MapLauncher HPTB_CAT(HPTB_CAT(hptb_binary_, HPTB_OPNAME), _mixed)(int lhs, int rhs) {
  using namespace hptb;
#define X(L, R) \
  if (lhs == dtype_of<L>::value && rhs == dtype_of<R>::value) return &Mixed<L, R>::launch;
  HPTB_MIXED_PAIRS(X)
#undef X
  return nullptr;
}

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
