// api_elementwise.cpp — C-ABI entry points for binary / unary / copy, plus the host helpers
// (promotion, broadcasting, axis processing, collapse) that the Rust shim shares with the library.
//
// Host-side flow of hptb_binary mirrors binary_fn_precompiled
// (hpt/src/backends/cuda/utils/binary/binary_normal.rs:371-544) without its four branches: scalar
// operands, same-shape contiguous operands and general broadcasts all become strides fed to the
// collapse pass; there is no D2H read of a scalar operand (:394,:439) and no table upload (:520-523).
#include <cstdlib>

#include "context.h"
#include "dtypes_x.h"
#include "layout.h"
#include "promote.h"

#include "map_plan.h"

// launcher getters exported by the instantiation units
// (weak: a development build may leave units out; a missing unit reports HPTB_ERR_DTYPE)
#define HPTB_WEAK __attribute__((weak))
extern "C" {
// specialised, vector-only kernels: same-dtype binary (T,T)→T, float unary T→T, same-dtype copy by element size
#define XB(F, NAME, E, K, B)                               \
  HPTB_WEAK hptb::MapLauncher hptb_binary_##NAME(int);     \
  HPTB_WEAK hptb::MapLauncher hptb_binary_##NAME##_mixed(int, int);
HPTB_FOR_BINARY_OPS(XB)
#undef XB
#define XU(NAME, E) HPTB_WEAK hptb::MapLauncher hptb_unary_##NAME(int);
HPTB_FOR_UNARY_OPS(XU)
#undef XU
// NormalUaryOps: T → T for every dtype
#define XU(NAME, E) HPTB_WEAK hptb::MapLauncher hptb_nunary_##NAME(int);
HPTB_FOR_NORMAL_UNARY_OPS(XU)
#undef XU
HPTB_WEAK hptb::MapLauncher hptb_copy_same(int);
// runtime-typed kernels, one set per OUTPUT dtype (mixed-dtype pairs, integer-input unary, astype, odd layouts)
#define XD(T, N, E)                                      \
  HPTB_WEAK hptb::MapLauncher hptb_dyn_binary_##N(void); \
  HPTB_WEAK hptb::MapLauncher hptb_dyn_cast_##N(void);   \
  HPTB_WEAK hptb::MapLauncher hptb_dyn_unary_##N(void);  \
  HPTB_WEAK hptb::MapLauncher hptb_dyn_cmp_##N(void);
HPTB_FOR_DTYPES(XD)
#undef XD
}

namespace hptb {

int promote(int lhs, int rhs, int kind) {
  if (!dtype_valid(lhs)) return -1;
  if (kind == HPTB_PROMOTE_FLOAT_UNARY) return kFloatOutUnary[lhs];
  if (!dtype_valid(rhs)) return -1;
  if (kind == HPTB_PROMOTE_NORMAL) return kNormalOut[lhs][rhs];
  if (kind == HPTB_PROMOTE_FLOAT_BINARY) return kFloatOutBinary[lhs][rhs];
  return -1;
}

typedef MapLauncher (*Getter)(int);
typedef MapLauncher (*DynGetter)(void);

static Getter binary_getter(int op) {
  switch (op) {
#define XB(F, NAME, E, K, B) \
  case E: return hptb_binary_##NAME;
    HPTB_FOR_BINARY_OPS(XB)
#undef XB
    default: return nullptr;
  }
}

typedef MapLauncher (*MixedGetter)(int, int);
static MixedGetter binary_mixed_getter(int op) {
  switch (op) {
#define XB(F, NAME, E, K, B) \
  case E: return hptb_binary_##NAME##_mixed;
    HPTB_FOR_BINARY_OPS(XB)
#undef XB
    default: return nullptr;
  }
}

static Getter unary_getter(int op) {
  switch (op) {
#define XU(NAME, E) \
  case E: return hptb_unary_##NAME;
    HPTB_FOR_UNARY_OPS(XU)
#undef XU
#define XU(NAME, E) \
  case E: return hptb_nunary_##NAME;
    HPTB_FOR_NORMAL_UNARY_OPS(XU)
#undef XU
    default: return nullptr;
  }
}

enum DynKind { kDynBinary, kDynCast, kDynUnary, kDynCmp };
static MapLauncher dyn_launcher(DynKind kind, int out_dtype) {
  DynGetter g = nullptr;
  switch (out_dtype) {
#define XD(T, N, E) \
  case E: g = kind == kDynBinary ? hptb_dyn_binary_##N : kind == kDynCast ? hptb_dyn_cast_##N : kind == kDynUnary ? hptb_dyn_unary_##N : hptb_dyn_cmp_##N; break;
    HPTB_FOR_DTYPES(XD)
#undef XD
    default: break;
  }
  return g ? g() : nullptr;
}

static int binary_kind(int op) {
  return (op == HPTB_DIV || op == HPTB_POW || op == HPTB_HYPOT) ? HPTB_PROMOTE_FLOAT_BINARY : HPTB_PROMOTE_NORMAL;
}
static bool is_bit_op(int op) { return op >= HPTB_BITAND && op <= HPTB_SHR; }
static bool is_int_or_bool(int dt) { return dt >= HPTB_BOOL && dt <= HPTB_U64; }

// shared tail of every elementwise entry: broadcast inputs to out's shape, collapse, launch.  `fast` (may be
// null) is the specialised same-dtype launcher; it may decline a layout (HPTB_FALLBACK), and `dyn` — the
// runtime-typed launcher for out's dtype — takes everything else.
static hptb_status run_map(hptb_ctx* ctx, MapLauncher fast, MapLauncher dyn, int op, hptb_tensor* out, const hptb_tensor* in0,
                           const hptb_tensor* in1, double alpha, double beta, void* stream) {
  int64_t strides[kMaxOperands][HPTB_MAX_DIMS] = {{0}};
  const int nd = out->ndim;
  for (int i = 0; i < nd; ++i) strides[0][i] = out->strides[i];
  HPTB_TRY(broadcast_strides(*in0, out->shape, nd, strides[1]));
  int nops = 2;
  if (in1) { HPTB_TRY(broadcast_strides(*in1, out->shape, nd, strides[2])); nops = 3; }
  MapPlan plan;
  collapse(nd, out->shape, nops, strides, nullptr, &plan.c);
  plan.ptr[0] = out->data;
  plan.ptr[1] = in0->data;
  plan.ptr[2] = in1 ? in1->data : nullptr;
  plan.alpha = alpha;
  plan.beta = beta;
  plan.sm_count = ctx->sm_count;
  plan.in_dtype[0] = in0->dtype;
  plan.in_dtype[1] = in1 ? in1->dtype : -1;
  plan.op = op;
  pass_direction(ctx, in0->data, 0, false);  // a forward streaming pass over in0 (snake order, context.h)
  DeviceGuard g(ctx->device);
  hptb_status st = fast ? fast(plan, (cudaStream_t)stream) : HPTB_FALLBACK;
  if (st == HPTB_FALLBACK) {
    if (!dyn) return fail(HPTB_ERR_DTYPE, "elementwise: no kernel for output dtype %s", dtype_name(out->dtype));
    st = dyn(plan, (cudaStream_t)stream);
  }
  if (st == HPTB_OK && plan.c.numel > 0) count_launches(1);
  return st;
}

}  // namespace hptb

using namespace hptb;

extern "C" {

int hptb_promote(int lhs, int rhs, int kind) { return promote(lhs, rhs, kind); }

int hptb_binary_out_dtype(int op, int lhs, int rhs) {
  if (op < 0 || op >= HPTB_BINARY_COUNT) return -1;
  if (is_bit_op(op) && !(is_int_or_bool(lhs) && is_int_or_bool(rhs))) return -1;  // BitWiseOut: integer / bool only
  int o = promote(lhs, rhs, binary_kind(op));
  if (o == HPTB_BOOL && (op == HPTB_SUB || op == HPTB_REM || op == HPTB_DIV)) return -1;  // _bool.rs:31-49 panics
  return o;
}
int hptb_unary_out_dtype(int op, int in) {
  if (op < 0 || op >= HPTB_UNARY_COUNT || !dtype_valid(in)) return -1;
  if (op == HPTB_BITNOT) return is_int_or_bool(in) ? in : -1;
  if (op >= HPTB_FLOAT_UNARY_COUNT) return in;  // NormalUaryOps: T → T
  return promote(in, in, HPTB_PROMOTE_FLOAT_UNARY);
}

hptb_status hptb_broadcast_shape(const int64_t* a, int na, const int64_t* b, int nb, int64_t* out, int* nout) {
  if ((!a && na) || (!b && nb) || !out || !nout) return fail(HPTB_ERR_INVALID, "broadcast_shape: null argument");
  return broadcast_shape(a, na, b, nb, out, nout);
}

hptb_status hptb_process_axes(const int64_t* axes, int naxes, int ndim, int32_t* out) {
  if ((!axes && naxes) || !out) return fail(HPTB_ERR_INVALID, "process_axes: null argument");
  for (int i = 0; i < naxes; ++i) {
    for (int j = 0; j < i; ++j)
      if (axes[j] == axes[i]) return fail(HPTB_ERR_AXIS, "Axis %lld is duplicated", (long long)axes[i]);
    int64_t a = axes[i] < 0 ? axes[i] + ndim : axes[i];
    if (ndim > 0 && (a < 0 || a >= ndim))
      // message format of ShapeError::DimOutOfRange (hpt-common/src/error/shape.rs:62), pinned by
      // hpt-tests/src/hpt_common/axis.rs:6-36 ("got" is the value after adding ndim to a negative axis)
      return fail(HPTB_ERR_AXIS, "Dimension out of range: expected in 0..%d, got %lld", ndim, (long long)a);
    out[i] = (int32_t)a;
  }
  return HPTB_OK;
}

hptb_status hptb_reduce_shape(const int64_t* shape, int ndim, const int32_t* axes, int naxes, int keep_dims,
                              int64_t* out_shape, int* out_ndim) {
  if ((!shape && ndim) || (!axes && naxes) || !out_shape || !out_ndim)
    return fail(HPTB_ERR_INVALID, "reduce_shape: null argument");
  int n = 0;
  for (int i = 0; i < ndim; ++i) {
    bool red = false;
    for (int j = 0; j < naxes; ++j) red |= axes[j] == i;
    if (red) { if (keep_dims) out_shape[n++] = 1; }
    else out_shape[n++] = shape[i];
  }
  if (n == 0) { out_shape[0] = 1; n = 1; }  // layout_utils.rs:342-347
  *out_ndim = n;
  return HPTB_OK;
}

hptb_status hptb_collapse(const hptb_tensor* const* operands, int n_operands, const uint8_t* reduce_mask,
                          hptb_collapse_plan* plan) {
  if (!operands || !plan || n_operands < 1 || n_operands > kMaxOperands)
    return fail(HPTB_ERR_INVALID, "collapse: bad arguments");
  for (int o = 0; o < n_operands; ++o) HPTB_TRY(validate_tensor(operands[o], "collapse operand"));
  // common shape = shape of the largest-rank operand broadcast with the others; for reductions
  // (mask given) operand 1 (the input) defines it and operand 0 (the output) has stride 0 on reduced dims
  int64_t shape[HPTB_MAX_DIMS];
  int nd = 0;
  int64_t strides[kMaxOperands][HPTB_MAX_DIMS] = {{0}};
  if (reduce_mask) {
    if (n_operands != 2) return fail(HPTB_ERR_INVALID, "collapse: a reduce mask needs exactly (out, in)");
    const hptb_tensor* in = operands[1];
    nd = in->ndim;
    int k = 0;
    for (int i = 0; i < nd; ++i) {
      shape[i] = in->shape[i];
      strides[1][i] = in->strides[i];
      if (reduce_mask[i]) strides[0][i] = 0;
      else {
        if (k >= operands[0]->ndim || operands[0]->shape[k] != in->shape[i])
          return fail(HPTB_ERR_SHAPE, "collapse: output shape does not match the kept dims");
        strides[0][i] = operands[0]->strides[k++];
      }
    }
  } else {
    nd = operands[0]->ndim;
    for (int i = 0; i < nd; ++i) { shape[i] = operands[0]->shape[i]; strides[0][i] = operands[0]->strides[i]; }
    for (int o = 1; o < n_operands; ++o) HPTB_TRY(broadcast_strides(*operands[o], shape, nd, strides[o]));
  }
  Collapsed c;
  collapse(nd, shape, n_operands, strides, reduce_mask, &c);
  memset(plan, 0, sizeof(*plan));
  plan->ndim = c.ndim;
  plan->launch_class = c.launch_class;
  plan->n_operands = n_operands;
  for (int i = 0; i < c.ndim; ++i) {
    plan->shape[i] = c.shape[i];
    plan->reduced[i] = c.reduced[i];
    for (int o = 0; o < n_operands; ++o) plan->strides[o][i] = c.strides[o][i];
  }
  return HPTB_OK;
}

hptb_status hptb_binary(hptb_ctx* ctx, int op, const hptb_tensor* lhs, const hptb_tensor* rhs, hptb_tensor* out,
                        void* stream) {
  if (!ctx) return fail(HPTB_ERR_INVALID, "binary: null ctx");
  if (op < 0 || op >= HPTB_BINARY_COUNT) return fail(HPTB_ERR_INVALID, "binary: bad op %d", op);
  HPTB_TRY(validate_tensor(lhs, "binary lhs"));
  HPTB_TRY(validate_tensor(rhs, "binary rhs"));
  HPTB_TRY(validate_tensor(out, "binary out"));
  int odt = hptb_binary_out_dtype(op, lhs->dtype, rhs->dtype);
  if (odt < 0)
    return fail(HPTB_ERR_DTYPE, "binary op %d is not supported for (%s, %s)", op, dtype_name(lhs->dtype), dtype_name(rhs->dtype));
  if (out->dtype != odt)
    return fail(HPTB_ERR_DTYPE, "binary: out dtype is %s but (%s, %s) promotes to %s", dtype_name(out->dtype),
                dtype_name(lhs->dtype), dtype_name(rhs->dtype), dtype_name(odt));
  int64_t bshape[HPTB_MAX_DIMS];
  int bn = 0;
  HPTB_TRY(broadcast_shape(lhs->shape, lhs->ndim, rhs->shape, rhs->ndim, bshape, &bn));
  bool same = bn == out->ndim;
  for (int i = 0; same && i < bn; ++i) same = bshape[i] == out->shape[i];
  if (!same) return fail(HPTB_ERR_SHAPE, "binary: out shape does not equal the broadcast shape of the operands");
  MapLauncher fast = nullptr;
  if (lhs->dtype == rhs->dtype && lhs->dtype == odt) {
    Getter g = binary_getter(op);
    fast = g ? g(odt) : nullptr;
  } else if (lhs->dtype != rhs->dtype) {
    MixedGetter g = binary_mixed_getter(op);
    fast = g ? g(lhs->dtype, rhs->dtype) : nullptr;
  }
  return run_map(ctx, fast, dyn_launcher(kDynBinary, odt), op, out, lhs, rhs, 0.0, 0.0, stream);
}

hptb_status hptb_compare(hptb_ctx* ctx, int op, const hptb_tensor* lhs, const hptb_tensor* rhs, hptb_tensor* out,
                         void* stream) {
  if (!ctx) return fail(HPTB_ERR_INVALID, "compare: null ctx");
  if (op < 0 || op >= HPTB_CMP_COUNT) return fail(HPTB_ERR_INVALID, "compare: bad op %d", op);
  HPTB_TRY(validate_tensor(lhs, "compare lhs"));
  HPTB_TRY(validate_tensor(rhs, "compare rhs"));
  HPTB_TRY(validate_tensor(out, "compare out"));
  if (out->dtype != HPTB_BOOL) return fail(HPTB_ERR_DTYPE, "compare: out dtype is %s, expected bool", dtype_name(out->dtype));
  const int pdt = promote(lhs->dtype, rhs->dtype, HPTB_PROMOTE_NORMAL);  // the type the comparison runs in
  if (pdt < 0) return fail(HPTB_ERR_DTYPE, "compare: bad operand dtype");
  int64_t bshape[HPTB_MAX_DIMS];
  int bn = 0;
  HPTB_TRY(broadcast_shape(lhs->shape, lhs->ndim, rhs->shape, rhs->ndim, bshape, &bn));
  bool same = bn == out->ndim;
  for (int i = 0; same && i < bn; ++i) same = bshape[i] == out->shape[i];
  if (!same) return fail(HPTB_ERR_SHAPE, "compare: out shape does not equal the broadcast shape of the operands");
  return run_map(ctx, nullptr, dyn_launcher(kDynCmp, pdt), op, out, lhs, rhs, 0.0, 0.0, stream);
}

hptb_status hptb_unary(hptb_ctx* ctx, int op, const hptb_tensor* in, hptb_tensor* out, double alpha, double beta,
                       void* stream) {
  if (!ctx) return fail(HPTB_ERR_INVALID, "unary: null ctx");
  if (op < 0 || op >= HPTB_UNARY_COUNT) return fail(HPTB_ERR_INVALID, "unary: bad op %d", op);
  HPTB_TRY(validate_tensor(in, "unary in"));
  HPTB_TRY(validate_tensor(out, "unary out"));
  int odt = hptb_unary_out_dtype(op, in->dtype);
  if (odt < 0) return fail(HPTB_ERR_DTYPE, "unary op %d is not supported for %s", op, dtype_name(in->dtype));
  if (op == HPTB_CLAMP && !(alpha <= beta)) return fail(HPTB_ERR_INVALID, "clamp: min must be <= max and neither may be NaN");
  if (out->dtype != odt)
    return fail(HPTB_ERR_DTYPE, "unary: out dtype is %s but %s promotes to %s", dtype_name(out->dtype), dtype_name(in->dtype),
                dtype_name(odt));
  bool same = in->ndim == out->ndim;
  for (int i = 0; same && i < in->ndim; ++i) same = in->shape[i] == out->shape[i];
  if (!same) return fail(HPTB_ERR_SHAPE, "unary: out shape differs from the input shape");
  MapLauncher fast = nullptr;
  if (in->dtype == odt) {
    Getter ug = unary_getter(op);
    fast = ug ? ug(odt) : nullptr;
  }
  return run_map(ctx, fast, dyn_launcher(kDynUnary, odt), op, out, in, nullptr, alpha, beta, stream);
}

hptb_status hptb_copy(hptb_ctx* ctx, const hptb_tensor* in, hptb_tensor* out, void* stream) {
  if (!ctx) return fail(HPTB_ERR_INVALID, "copy: null ctx");
  HPTB_TRY(validate_tensor(in, "copy in"));
  HPTB_TRY(validate_tensor(out, "copy out"));
  if (!dtype_valid(in->dtype) || !dtype_valid(out->dtype)) return fail(HPTB_ERR_DTYPE, "copy: bad dtype");
  MapLauncher fast = nullptr;
  if (in->dtype == out->dtype && hptb_copy_same) fast = hptb_copy_same((int)dtype_size(in->dtype));
  // `in` may broadcast into `out` (used by fill-from-tensor and expand().contiguous())
  return run_map(ctx, fast, dyn_launcher(kDynCast, out->dtype), 0, out, in, nullptr, 0.0, 0.0, stream);
}

}  // extern "C"
