// nunary_inst.cu — one translation unit per NormalUaryOps op (+ BITNOT); compiled with
//   -DHPTB_OPENUM=<hptb_unary_op> -DHPTB_OPNAME=<name>
// and exports `hptb_nunary_<op>(dtype) -> launcher` for all 13 dtypes (T → T): the vector-only specialised
// kernels (contiguous / inner-contiguous rows and the shared-memory transposing tile kernel for permuted views).
// Replaces the NVRTC-generated kernels of uary_fn_with_out_simd (hpt/src/backends/cuda/tensor_internal/
// normal_out_unary.rs); unaligned or odd layouts take the runtime-typed kernel (dyn_inst.cu).
#include "dtypes_x.h"
#include "elementwise.cuh"
#include "ops.cuh"

#include <limits>

namespace hptb {
namespace {
// scalar parameter (alpha / beta: leaky_relu slope, clamp bounds) in the compute type, on the HOST: Rust `as` from
// f64 for integers (saturating, NaN → 0), `!= 0` for bool
template <typename C>
C host_param(double v) {
  if constexpr (is_bool_t<C>::value) return b8{(uint8_t)(v != 0.0)};
  else if constexpr (std::is_integral<C>::value) {
    if (v != v) return (C)0;
    if (v <= (double)std::numeric_limits<C>::min()) return std::numeric_limits<C>::min();
    if (v >= (double)std::numeric_limits<C>::max()) return std::numeric_limits<C>::max();
    return (C)v;
  } else return (C)v;
}

template <typename T>
struct Inst {
  static hptb_status launch(const MapPlan& plan, cudaStream_t s) {
    typedef NormalUnaryFn<HPTB_OPENUM, T> F;
    F f;
    f.alpha = host_param<compute_t<T>>(plan.alpha);
    f.beta = host_param<compute_t<T>>(plan.beta);
    return launch_map<1, F, T, T, T>(plan, f, s);
  }
};
}  // namespace
}  // namespace hptb

extern "C" hptb::MapLauncher HPTB_CAT(hptb_nunary_, HPTB_OPNAME)(int dt) {
  using namespace hptb;
  if (HPTB_OPENUM == HPTB_BITNOT && dt > HPTB_U64) return nullptr;
  switch (dt) {
#define X(T, N, E) \
  case E: return &Inst<T>::launch;
    HPTB_FOR_DTYPES(X)
#undef X
    default: return nullptr;
  }
}
