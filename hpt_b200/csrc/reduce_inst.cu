// reduce_inst.cu — one translation unit per reduction op (-DHPTB_OPENUM, -DHPTB_OPNAME); exports
// `hptb_reduce_<op>(in dtype) -> launcher` for the 13 input dtypes.
#include "dtypes_x.h"
#include "reduce.cuh"

namespace hptb {
namespace {
template <typename T>
struct Inst {
  static hptb_status launch(const ReducePlan& plan, cudaStream_t s) {
    return launch_reduce<ReduceOp<HPTB_OPENUM, T>, T>(plan, s);
  }
};
}  // namespace
}  // namespace hptb

extern "C" hptb::ReduceLauncher HPTB_CAT(hptb_reduce_, HPTB_OPNAME)(int in) {
  using namespace hptb;
  switch (in) {
#define X(T, N, E) \
  case E: return &Inst<T>::launch;
    HPTB_FOR_DTYPES(X)
#undef X
    default: return nullptr;
  }
}

#ifdef HPTB_LOGSUMEXP_LONG
// second instantiation of the logsumexp unit: bare-ex2 accumulation for long reductions (reduce.cuh LogSumExpOp)
namespace hptb {
namespace {
template <typename T>
struct InstLong {
  static hptb_status launch(const ReducePlan& plan, cudaStream_t s) {
    return launch_reduce<LogSumExpOp<T, true>, T>(plan, s);
  }
};
}  // namespace
}  // namespace hptb
extern "C" hptb::ReduceLauncher hptb_reduce_logsumexp_long(int in) {
  using namespace hptb;
  switch (in) {  // f64-computed inputs have no cheaper form
#define X(T, N, E) \
  case E: return std::is_same<typename LogSumExpOp<T, true>::Acc, float>::value ? &InstLong<T>::launch : nullptr;
    HPTB_FOR_DTYPES(X)
#undef X
    default: return nullptr;
  }
}
#endif
