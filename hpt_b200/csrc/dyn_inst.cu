// dyn_inst.cu — one translation unit per OUTPUT dtype (-DHPTB_OUT=<c++ type> -DHPTB_OUTNAME=<short name>
// -DHPTB_OUT_FLOAT=<0|1>): the runtime-typed elementwise kernels (elementwise_dyn.cuh) for binary ops, dtype
// conversion, unary ops and compares (in the promoted type).  Exports `hptb_dyn_{binary,cast,unary,cmp}_<out>()`.
#include "dtypes_x.h"
#include "elementwise_dyn.cuh"

namespace hptb {
namespace {
typedef HPTB_OUT O;
constexpr int kMaxIn = sizeof(O) >= 4 ? (int)sizeof(O) : 4;  // widest input a binary / unary op with Output O can see
hptb_status launch_binary(const MapPlan& plan, cudaStream_t s) { return launch_map_dyn<2, kMaxIn, DynBinaryFn<O>, O>(plan, s); }
hptb_status launch_cast(const MapPlan& plan, cudaStream_t s) { return launch_map_dyn<1, 8, DynCastFn<O>, O>(plan, s); }
// unary: FloatUnaryOps on float outputs (integer inputs promote), NormalUaryOps / BITNOT on every output dtype
hptb_status launch_unary(const MapPlan& plan, cudaStream_t s) { return launch_map_dyn<1, kMaxIn, DynUnaryFn<O>, O>(plan, s); }
// compare: O is the promoted type of the two inputs, the stored type is bool
hptb_status launch_cmp(const MapPlan& plan, cudaStream_t s) { return launch_map_dyn<2, kMaxIn, DynCmpFn<O>, O, b8>(plan, s); }
}  // namespace
}  // namespace hptb

extern "C" hptb::MapLauncher HPTB_CAT(hptb_dyn_binary_, HPTB_OUTNAME)() { return &hptb::launch_binary; }
extern "C" hptb::MapLauncher HPTB_CAT(hptb_dyn_cast_, HPTB_OUTNAME)() { return &hptb::launch_cast; }
extern "C" hptb::MapLauncher HPTB_CAT(hptb_dyn_unary_, HPTB_OUTNAME)() { return &hptb::launch_unary; }
extern "C" hptb::MapLauncher HPTB_CAT(hptb_dyn_cmp_, HPTB_OUTNAME)() { return &hptb::launch_cmp; }
