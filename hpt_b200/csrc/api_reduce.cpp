// api_reduce.cpp — C-ABI entry point for axis reductions and the host planner in front of reduce.cuh.
//
// Host flow mirrors reduce / reduce2 / reduce3 → contiguous_reduce / uncontiguous_reduce
// (hpt/src/backends/cuda/utils/reduce/reduce.rs:62-838) and reduce_prepare (reduce_utils.rs:18-74):
// validate axes, form the (out, in) stride pair with stride 0 on reduced dims, collapse, launch.
// No per-call cuMemAlloc (reduce.rs:272-273, :446-451): scratch and tickets come from the context.
#include <mutex>
#include <unordered_map>
#include <vector>

#include <algorithm>
#include "dtypes_x.h"
#include "map_plan.h"
#include "promote.h"
#include "reduce_plan.h"

#define HPTB_WEAK __attribute__((weak))
extern "C" {
HPTB_WEAK hptb::ReduceLauncher hptb_reduce_sum(int);
HPTB_WEAK hptb::ReduceLauncher hptb_reduce_mean(int);
HPTB_WEAK hptb::ReduceLauncher hptb_reduce_max(int);
HPTB_WEAK hptb::ReduceLauncher hptb_reduce_min(int);
HPTB_WEAK hptb::ReduceLauncher hptb_reduce_argmax(int);
HPTB_WEAK hptb::ReduceLauncher hptb_reduce_argmin(int);
HPTB_WEAK hptb::ReduceLauncher hptb_reduce_logsumexp(int);
HPTB_WEAK hptb::ReduceLauncher hptb_reduce_logsumexp_long(int);
HPTB_WEAK hptb::ReduceLauncher hptb_reduce_sum_square(int);
HPTB_WEAK hptb::ReduceLauncher hptb_reduce_prod(int);
HPTB_WEAK hptb::ReduceLauncher hptb_reduce_reducel1(int);
HPTB_WEAK hptb::ReduceLauncher hptb_reduce_nansum(int);
HPTB_WEAK hptb::ReduceLauncher hptb_reduce_nanprod(int);
HPTB_WEAK hptb::ReduceLauncher hptb_reduce_all(int);
HPTB_WEAK hptb::ReduceLauncher hptb_reduce_any(int);
HPTB_WEAK hptb::ReduceLauncher hptb_reduce_reducel2(int);
HPTB_WEAK hptb::ReduceLauncher hptb_reduce_reducel3(int);
HPTB_WEAK hptb::FusedLauncher hptb_fused_add(int);
HPTB_WEAK hptb::FusedLauncher hptb_fused_sub(int);
HPTB_WEAK hptb::FusedLauncher hptb_fused_mul(int);
int hptb_binary_out_dtype(int op, int lhs, int rhs);
hptb_status hptb_binary(hptb_ctx*, int, const hptb_tensor*, const hptb_tensor*, hptb_tensor*, void*);
}

namespace hptb {

namespace {
struct TicketBuf {
  uint32_t* ptr = nullptr;
  size_t n = 0;
};
struct TicketState {
  std::mutex mu;
  std::unordered_map<void*, TicketBuf> by_stream;
  std::vector<void*> retired;
};
std::mutex g_mu;
std::unordered_map<hptb_ctx*, TicketState*> g_states;

TicketState* state_of(hptb_ctx* ctx) {
  std::lock_guard<std::mutex> g(g_mu);
  auto it = g_states.find(ctx);
  if (it != g_states.end()) return it->second;
  TicketState* s = new TicketState();
  g_states[ctx] = s;
  return s;
}
}  // namespace

uint32_t* ctx_tickets(hptb_ctx* ctx, cudaStream_t stream, size_t n) {
  TicketState* st = state_of(ctx);
  std::lock_guard<std::mutex> g(st->mu);
  TicketBuf& b = st->by_stream[(void*)stream];
  if (b.n >= n) return b.ptr;
  size_t cap = 1 << 16;
  while (cap < n) cap <<= 1;
  void* p = nullptr;
  if (cudaMalloc(&p, cap * sizeof(uint32_t)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (cudaMemsetAsync(p, 0, cap * sizeof(uint32_t), stream) != cudaSuccess) { cudaFree(p); return nullptr; }
  if (b.ptr) st->retired.push_back(b.ptr);  // earlier launches on the stream may still use it
  b.ptr = (uint32_t*)p;
  b.n = cap;
  return b.ptr;
}

void ctx_tickets_destroy(hptb_ctx* ctx) {
  TicketState* st = nullptr;
  {
    std::lock_guard<std::mutex> g(g_mu);
    auto it = g_states.find(ctx);
    if (it == g_states.end()) return;
    st = it->second;
    g_states.erase(it);
  }
  cudaDeviceSynchronize();
  for (auto& kv : st->by_stream) cudaFree(kv.second.ptr);
  for (void* p : st->retired) cudaFree(p);
  delete st;
}

static ReduceLauncher reduce_launcher(int op, int dt) {
  typedef ReduceLauncher (*G)(int);
  G g = nullptr;
  switch (op) {
    case HPTB_SUM: g = hptb_reduce_sum; break;
    case HPTB_MEAN: g = hptb_reduce_mean; break;
    case HPTB_MAX: g = hptb_reduce_max; break;
    case HPTB_MIN: g = hptb_reduce_min; break;
    case HPTB_ARGMAX: g = hptb_reduce_argmax; break;
    case HPTB_ARGMIN: g = hptb_reduce_argmin; break;
    case HPTB_LOGSUMEXP: g = hptb_reduce_logsumexp; break;
    case HPTB_SUM_SQUARE: g = hptb_reduce_sum_square; break;
    case HPTB_PROD: g = hptb_reduce_prod; break;
    case HPTB_REDUCEL1: g = hptb_reduce_reducel1; break;
    case HPTB_NANSUM: g = hptb_reduce_nansum; break;
    case HPTB_NANPROD: g = hptb_reduce_nanprod; break;
    case HPTB_ALL: g = hptb_reduce_all; break;
    case HPTB_ANY: g = hptb_reduce_any; break;
    case HPTB_REDUCEL2: g = hptb_reduce_reducel2; break;
    case HPTB_REDUCEL3: g = hptb_reduce_reducel3; break;
    default: break;
  }
  return g ? g(dt) : nullptr;
}

hptb_status build_reduce_plan(hptb_ctx* ctx, const hptb_tensor* in, const int32_t* axes, int naxes,
                              const hptb_tensor* out, ReducePlan* plan) {
  uint8_t mask[HPTB_MAX_DIMS] = {0};
  for (int i = 0; i < naxes; ++i) {
    int a = axes[i];
    if (in->ndim > 0 && (a < 0 || a >= in->ndim)) return fail(HPTB_ERR_AXIS, "reduce: axis %d out of range for ndim %d", a, in->ndim);
    if (in->ndim == 0) continue;
    if (mask[a]) return fail(HPTB_ERR_AXIS, "reduce: axis %d is duplicated", a);
    mask[a] = 1;
  }
  // expected output shape (keep_dims = false; all reduced → [1])
  int64_t oshape[HPTB_MAX_DIMS];
  int on = 0;
  double count = 1.0;
  for (int i = 0; i < in->ndim; ++i) {
    if (mask[i]) count *= (double)in->shape[i];
    else oshape[on++] = in->shape[i];
  }
  bool scalar_out = on == 0;
  if (scalar_out) { oshape[0] = 1; on = 1; }
  bool same = out->ndim == on;
  for (int i = 0; same && i < on; ++i) same = out->shape[i] == oshape[i];
  if (!same) return fail(HPTB_ERR_SHAPE, "reduce: out shape does not match the reduced shape of the input");
  int64_t strides[kMaxOperands][HPTB_MAX_DIMS] = {{0}};
  int k = 0;
  for (int i = 0; i < in->ndim; ++i) {
    strides[1][i] = in->strides[i];
    strides[0][i] = mask[i] ? 0 : out->strides[k++];
  }
  collapse(in->ndim, in->shape, 2, strides, mask, &plan->c);
  plan->in = in->data;
  plan->out = out->data;
  plan->count = count;
  plan->ctx = ctx;
  return HPTB_OK;
}

}  // namespace hptb

namespace hptb {
// Host-side routing of one reduction (pure: looks at shapes, strides, dtypes and the alignment of in->data only).
//
// PEEL — rows that start off the 16-byte boundary (a[5:8000, 3:8100].sum(1)): every row has the same misalignment when
// the other strides are multiples of a pack, so a row is head (< one pack) + aligned body + tail.  The body takes the
// vector kernels; head and tail are folded into `out` by two tiny launches (init_out = 0).  Only where `out` holds
// the accumulator exactly (f32 / f64 / integers) and the op folds associatively.  (Scalar loads through the general
// kernel: 26 issued instructions per element, 82.7 µs for that f32 window against 40.4 µs peeled; torch 52.7 µs.)
//
// TWO_STEP — transposing reduction: the input's fastest KEPT dim is not the output's fastest dim
// (x.permute(2,0,1).sum(2)): the lanes that read a full line would each write to a different line, and no kernel
// class is coalesced on both sides (0.13 of peak through the general kernel).  The output is the small side, so
// reduce into a scratch laid out in the INPUT's dim order (coalesced cols / rows kernels apply) and gather it
// into `out` afterwards (319 → 49 µs).
void plan_reduce_route(int op, const hptb_tensor* in, const int32_t* axes, int naxes, const hptb_tensor* out, int init_out,
                       double count_override, hptb_reduce_route_t* route) {
  memset(route, 0, sizeof(*route));
  route->kind = HPTB_ROUTE_DIRECT;
  if (init_out && naxes >= 1 && in->ndim >= 1 && in->data && count_override < 0 && count_override != -2.0) {
    const int last = in->ndim - 1;
    const size_t esz = dtype_size(in->dtype);
    const int64_t pack = esz <= 8 ? (int64_t)(16 / esz) : 1;
    const bool fold_ok = op == HPTB_SUM || op == HPTB_MAX || op == HPTB_MIN || op == HPTB_PROD || op == HPTB_SUM_SQUARE ||
                         op == HPTB_REDUCEL1 || op == HPTB_NANSUM || op == HPTB_NANPROD || op == HPTB_ALL || op == HPTB_ANY;
    // ops whose OUTPUT is not their accumulator (mean: Σ then ÷ n; logsumexp: Σexp then ln; l2 / l3: power sum then root)
    // and half-precision dtypes (f32 accumulator, half output) cannot fold through `out`: their pieces are combined in a
    // scratch of ACCUMULATORS and finished by the combine kernel (PEEL_RAW)
    const bool raw_ok = op == HPTB_MEAN || op == HPTB_LOGSUMEXP || op == HPTB_REDUCEL2 || op == HPTB_REDUCEL3;
    const bool half = in->dtype == HPTB_F16 || in->dtype == HPTB_BF16;
    const bool op_ok = fold_ok || raw_ok;
    const bool dt_ok = true;
    bool has_last = false;
    for (int i = 0; i < naxes; ++i) has_last |= axes[i] == last;
    const uintptr_t addr = reinterpret_cast<uintptr_t>(in->data);
    if (op_ok && dt_ok && has_last && pack >= 2 && in->strides[last] == 1 && in->shape[last] >= 64 && numel(*in) >= (1 << 16) &&
        addr % esz == 0 && addr % 16 != 0) {
      bool rows_ok = true;
      for (int d = 0; d < last; ++d)
        if (in->shape[d] > 1 && (uint64_t)(std::llabs(in->strides[d]) * (int64_t)esz) % 16) rows_ok = false;
      if (rows_ok) {
        const int64_t L = in->shape[last];
        route->kind = (raw_ok || half) ? HPTB_ROUTE_PEEL_RAW : HPTB_ROUTE_PEEL;
        route->head = pack - (int64_t)((addr % 16) / esz);
        route->body = ((L - route->head) / pack) * pack;
        route->tail = L - route->head - route->body;
        return;
      }
    }
  }
  if (init_out && count_override != -2.0 && in->ndim > 0 && out->ndim >= 2) {
    uint8_t mask[HPTB_MAX_DIMS] = {0};
    bool ok = true;
    for (int i = 0; i < naxes; ++i) {
      if (axes[i] < 0 || axes[i] >= in->ndim) { ok = false; break; }
      mask[axes[i]] = 1;
    }
    int src_dim[HPTB_MAX_DIMS], nk = 0;
    double red = 1.0;
    for (int i = 0; ok && i < in->ndim; ++i) {
      if (mask[i]) red *= (double)in->shape[i];
      else src_dim[nk++] = i;
    }
    if (ok && nk == out->ndim && red >= 8.0) {
      int fin = -1, fout = -1;  // fastest kept dim on the input side / on the output side (extent > 1)
      for (int j = 0; j < nk; ++j) {
        if (out->shape[j] <= 1) continue;
        if (fin < 0 || std::llabs(in->strides[src_dim[j]]) < std::llabs(in->strides[src_dim[fin]])) fin = j;
        if (fout < 0 || std::llabs(out->strides[j]) < std::llabs(out->strides[fout])) fout = j;
      }
      if (fin >= 0 && fin != fout && std::llabs(in->strides[src_dim[fin]]) == 1 && out->shape[fin] >= 32 && numel(*out) > 0) {
        int order[HPTB_MAX_DIMS];  // out dims from the slowest to the fastest input stride
        for (int j = 0; j < nk; ++j) order[j] = j;
        std::sort(order, order + nk, [&](int a, int b) {
          return std::llabs(in->strides[src_dim[a]]) > std::llabs(in->strides[src_dim[b]]);
        });
        int64_t st = 1;
        for (int j = nk - 1; j >= 0; --j) { route->scratch_strides[order[j]] = st; st *= out->shape[order[j]]; }
        route->kind = HPTB_ROUTE_TWO_STEP;
      }
    }
  }
}
}  // namespace hptb

namespace hptb {
hptb_status reduce_impl(hptb_ctx* ctx, int op, const hptb_tensor* in, const int32_t* axes, int naxes, hptb_tensor* out,
                        int init_out, double count_override, void* stream);
hptb_status reduce_pieces(hptb_ctx* ctx, int op, const hptb_tensor* in, const int32_t* axes, int naxes, const hptb_tensor* lay, double count,
                          void* raw, bool accumulate, void* stream);
hptb_status reduce_combine(hptb_ctx* ctx, int op, int in_dtype, double count, const void* partials, int gathered, const XchgParams* x,
                           const hptb_tensor* out, void* stream);
}
using namespace hptb;

extern "C" {

int hptb_reduce_out_dtype(int op, int in) {
  if (!dtype_valid(in)) return -1;
  switch (op) {
    case HPTB_SUM: case HPTB_MAX: case HPTB_MIN: case HPTB_SUM_SQUARE: case HPTB_PROD:
    case HPTB_REDUCEL1: case HPTB_NANSUM: case HPTB_NANPROD: return in;
    case HPTB_ALL: case HPTB_ANY: return HPTB_BOOL;
    case HPTB_MEAN: case HPTB_LOGSUMEXP: case HPTB_REDUCEL2: case HPTB_REDUCEL3: return kFloatOutBinary[in][in];
    case HPTB_ARGMAX: case HPTB_ARGMIN: return HPTB_I64;
    default: return -1;
  }
}

hptb_status hptb_reduce(hptb_ctx* ctx, int op, const hptb_tensor* in, const int32_t* axes, int naxes, hptb_tensor* out,
                        int init_out, void* stream) {
  return reduce_impl(ctx, op, in, axes, naxes, out, init_out, -1.0, stream);
}

hptb_status hptb_reduce_route(int op, const hptb_tensor* in, const int32_t* axes, int naxes, const hptb_tensor* out,
                              int init_out, hptb_reduce_route_t* route) {
  if (!in || !out || !route || (!axes && naxes)) return fail(HPTB_ERR_INVALID, "reduce_route: null argument");
  HPTB_TRY(validate_tensor(in, "reduce_route in"));
  HPTB_TRY(validate_tensor(out, "reduce_route out"));
  plan_reduce_route(op, in, axes, naxes, out, init_out, -1.0, route);
  return HPTB_OK;
}

}  // extern "C"

namespace hptb {
hptb_status reduce_impl(hptb_ctx* ctx, int op, const hptb_tensor* in, const int32_t* axes, int naxes, hptb_tensor* out,
                        int init_out, double count_override, void* stream) {
  if (!ctx) return fail(HPTB_ERR_INVALID, "reduce: null ctx");
  if (op < 0 || op >= HPTB_REDUCE_COUNT) return fail(HPTB_ERR_INVALID, "reduce: bad op %d", op);
  if (!axes && naxes) return fail(HPTB_ERR_INVALID, "reduce: null axes");
  HPTB_TRY(validate_tensor(in, "reduce in"));
  HPTB_TRY(validate_tensor(out, "reduce out"));
  if ((op == HPTB_ARGMAX || op == HPTB_ARGMIN) && naxes != 1)
    return fail(HPTB_ERR_AXIS, "argmax/argmin take exactly one axis (got %d)", naxes);
  int odt = hptb_reduce_out_dtype(op, in->dtype);
  if (out->dtype != odt)
    return fail(HPTB_ERR_DTYPE, "reduce: out dtype is %s, expected %s", dtype_name(out->dtype), dtype_name(odt));
  ReduceLauncher fn = reduce_launcher(op, in->dtype);
  if (!fn) return fail(HPTB_ERR_DTYPE, "reduce: no kernel for op %d on %s", op, dtype_name(in->dtype));
  // host-side routing (plan_reduce_route above): peel misaligned rows, or reduce into an input-ordered scratch
  hptb_reduce_route_t route;
  plan_reduce_route(op, in, axes, naxes, out, init_out, count_override, &route);
  if (route.kind == HPTB_ROUTE_PEEL) {
    const int last = in->ndim - 1;
    const int64_t esz = (int64_t)dtype_size(in->dtype);
    hptb_tensor part = *in;
    part.data = static_cast<char*>(in->data) + route.head * esz;
    part.shape[last] = route.body;
    HPTB_TRY(reduce_impl(ctx, op, &part, axes, naxes, out, 1, count_override, stream));
    part.data = in->data;
    part.shape[last] = route.head;
    HPTB_TRY(reduce_impl(ctx, op, &part, axes, naxes, out, 0, count_override, stream));
    if (route.tail > 0) {
      part.data = static_cast<char*>(in->data) + (route.head + route.body) * esz;
      part.shape[last] = route.tail;
      HPTB_TRY(reduce_impl(ctx, op, &part, axes, naxes, out, 0, count_override, stream));
    }
    return HPTB_OK;
  }
  if (route.kind == HPTB_ROUTE_PEEL_RAW) {
    // head / aligned body / tail reduced into a scratch of accumulators laid out like a row-major `out`, then one small
    // kernel applies the op's post step and writes `out` (any strides)
    const int last = in->ndim - 1;
    const int64_t esz = (int64_t)dtype_size(in->dtype);
    const int64_t M = numel(*out);
    if (M == 0) return HPTB_OK;
    double count = 1.0;
    for (int i = 0; i < naxes; ++i) count *= (double)in->shape[axes[i]];
    hptb_tensor lay = *out;
    int64_t st = 1;
    for (int i = lay.ndim - 1; i >= 0; --i) { lay.strides[i] = st; st *= lay.shape[i]; }
    lay.data = nullptr;
    Scratch raw;
    HPTB_TRY(raw.get(ctx, (size_t)M * 16, stream));
    hptb_tensor part = *in;
    part.data = static_cast<char*>(in->data) + route.head * esz;
    part.shape[last] = route.body;
    HPTB_TRY(reduce_pieces(ctx, op, &part, axes, naxes, &lay, count, raw.ptr, false, stream));
    part.data = in->data;
    part.shape[last] = route.head;
    HPTB_TRY(reduce_pieces(ctx, op, &part, axes, naxes, &lay, count, raw.ptr, true, stream));
    if (route.tail > 0) {
      part.data = static_cast<char*>(in->data) + (route.head + route.body) * esz;
      part.shape[last] = route.tail;
      HPTB_TRY(reduce_pieces(ctx, op, &part, axes, naxes, &lay, count, raw.ptr, true, stream));
    }
    return reduce_combine(ctx, op, in->dtype, count, raw.ptr, 1, nullptr, out, stream);
  }
  if (route.kind == HPTB_ROUTE_TWO_STEP) {
    hptb_tensor tmp = *out;
    for (int j = 0; j < out->ndim; ++j) tmp.strides[j] = route.scratch_strides[j];
    Scratch sc;
    HPTB_TRY(sc.get(ctx, (size_t)numel(*out) * dtype_size(out->dtype), stream));
    tmp.data = sc.ptr;
    HPTB_TRY(reduce_impl(ctx, op, in, axes, naxes, &tmp, 1, count_override, stream));
    return hptb_copy(ctx, &tmp, out, stream);
  }
  ReducePlan plan;
  HPTB_TRY(build_reduce_plan(ctx, in, axes, naxes, out, &plan));
  plan.fold_out = init_out ? 0 : 1;
  if (op == HPTB_LOGSUMEXP && plan.count >= 512.0 && hptb_reduce_logsumexp_long) {  // kLogSumExpLongMin (reduce.cuh)
    if (ReduceLauncher lf = hptb_reduce_logsumexp_long(in->dtype)) fn = lf;
  }
  if (count_override > 0) plan.count = count_override;  // sharded mean: divide the local Σ by the GLOBAL count
  if (count_override == -2.0 && (op == HPTB_REDUCEL2 || op == HPTB_REDUCEL3)) plan.count = -2.0;  // kPartialPowerSum
  plan.reverse = pass_direction(ctx, in->data, (size_t)numel(*in) * dtype_size(in->dtype), true) ? 1 : 0;
  DeviceGuard g(ctx->device);
  hptb_status st = fn(plan, (cudaStream_t)stream);
  if (st == HPTB_OK) count_launches(1);
  return st;
}
}  // namespace hptb

// ---- sharded reductions: the two halves comm.cpp composes (xchg.cuh) ------------------------------------------------
namespace hptb {
static ReduceLauncher sharded_launcher(int op, int in_dtype, double count) {
  ReduceLauncher fn = reduce_launcher(op, in_dtype);
  if (fn && op == HPTB_LOGSUMEXP && count >= 512.0 && hptb_reduce_logsumexp_long)
    if (ReduceLauncher lf = hptb_reduce_logsumexp_long(in_dtype)) fn = lf;
  return fn;
}

// Local pass.  `lay` is `out`'s shape with ROW-MAJOR strides; lay->data is the real output when that is how `out` is laid
// out (the kernel may then exchange in its epilogue and write the result: *fused = true), else NULL.  Otherwise the bare
// accumulators land in `raw` at the same row-major offsets.  `count` is the GLOBAL element count per output.
hptb_status reduce_for_exchange(hptb_ctx* ctx, int op, const hptb_tensor* in, const int32_t* axes, int naxes, const hptb_tensor* lay,
                                double count, const XchgParams* x, void* raw, bool* fused, size_t* acc_bytes, void* stream) {
  ReduceLauncher fn = sharded_launcher(op, in->dtype, count);
  if (!fn) return fail(HPTB_ERR_DTYPE, "reduce_sharded: no kernel for op %d on %s", op, dtype_name(in->dtype));
  ReducePlan plan;
  HPTB_TRY(build_reduce_plan(ctx, in, axes, naxes, lay, &plan));
  plan.count = count;
  plan.fold_out = 0;
  plan.xchg = x;
  plan.raw_out = raw;
  plan.fused = fused;
  plan.acc_bytes = acc_bytes;
  plan.reverse = pass_direction(ctx, in->data, (size_t)numel(*in) * dtype_size(in->dtype), true) ? 1 : 0;
  DeviceGuard g(ctx->device);
  hptb_status st = fn(plan, (cudaStream_t)stream);
  if (st == HPTB_OK) count_launches(1);
  return st;
}

// One piece of a peeled reduction (PEEL_RAW): bare accumulators into `raw` (row-major `lay`), stored or combined with
// what the previous pieces left there.
hptb_status reduce_pieces(hptb_ctx* ctx, int op, const hptb_tensor* in, const int32_t* axes, int naxes, const hptb_tensor* lay, double count,
                          void* raw, bool accumulate, void* stream) {
  ReduceLauncher fn = sharded_launcher(op, in->dtype, count);
  if (!fn) return fail(HPTB_ERR_DTYPE, "reduce: no kernel for op %d on %s", op, dtype_name(in->dtype));
  ReducePlan plan;
  HPTB_TRY(build_reduce_plan(ctx, in, axes, naxes, lay, &plan));
  plan.count = count;
  plan.fold_out = accumulate ? 3 : 0;  // kFoldRawAcc (reduce.cuh)
  plan.raw_out = raw;
  DeviceGuard g(ctx->device);
  hptb_status st = fn(plan, (cudaStream_t)stream);
  if (st == HPTB_OK) count_launches(1);
  return st;
}

// Exchange (gathered == 0: through the peer mailboxes of `x`) or take the NCCL-gathered accumulators ([gathered][M]),
// combine in rank order, apply the op's post step and write `out` (any strides).
hptb_status reduce_combine(hptb_ctx* ctx, int op, int in_dtype, double count, const void* partials, int gathered, const XchgParams* x,
                           const hptb_tensor* out, void* stream) {
  ReduceLauncher fn = sharded_launcher(op, in_dtype, count);
  if (!fn) return fail(HPTB_ERR_DTYPE, "reduce_sharded: no kernel for op %d on %s", op, dtype_name(in_dtype));
  ReducePlan plan;
  plan.mode = kPlanCombine;
  plan.ctx = ctx;
  plan.in = partials;
  plan.out = out->data;
  plan.count = count;
  plan.xchg = x;
  plan.gathered = gathered;
  plan.comb_M = numel(*out);
  plan.comb_nk = out->ndim;
  for (int i = 0; i < out->ndim; ++i) {
    plan.comb_shape[i] = out->shape[out->ndim - 1 - i];
    plan.comb_stride[i] = out->strides[out->ndim - 1 - i];
  }
  DeviceGuard g(ctx->device);
  hptb_status st = fn(plan, (cudaStream_t)stream);
  if (st == HPTB_OK) count_launches(1);
  return st;
}
}  // namespace hptb

// ---- elementwise → reduce fusion ------------------------------------------------------------------------------------
extern "C" hptb_status hptb_binary_reduce(hptb_ctx* ctx, int bin_op, int red_op, const hptb_tensor* lhs, const hptb_tensor* rhs,
                                          const int32_t* axes, int naxes, hptb_tensor* out, int init_out, void* stream) {
  if (!ctx) return fail(HPTB_ERR_INVALID, "binary_reduce: null ctx");
  if (!axes && naxes) return fail(HPTB_ERR_INVALID, "binary_reduce: null axes");
  if (red_op < 0 || red_op >= HPTB_REDUCE_COUNT) return fail(HPTB_ERR_INVALID, "binary_reduce: bad reduce op %d", red_op);
  HPTB_TRY(validate_tensor(lhs, "binary_reduce lhs"));
  HPTB_TRY(validate_tensor(rhs, "binary_reduce rhs"));
  HPTB_TRY(validate_tensor(out, "binary_reduce out"));
  const int mid = hptb_binary_out_dtype(bin_op, lhs->dtype, rhs->dtype);
  if (mid < 0)
    return fail(HPTB_ERR_DTYPE, "binary op %d is not supported for (%s, %s)", bin_op, dtype_name(lhs->dtype), dtype_name(rhs->dtype));
  const int odt = hptb_reduce_out_dtype(red_op, mid);
  if (out->dtype != odt) return fail(HPTB_ERR_DTYPE, "binary_reduce: out dtype is %s, expected %s", dtype_name(out->dtype), dtype_name(odt));
  // the elementwise result (never materialised on the fast path): broadcast shape of the operands
  hptb_tensor tmp;
  memset(&tmp, 0, sizeof(tmp));
  HPTB_TRY(broadcast_shape(lhs->shape, lhs->ndim, rhs->shape, rhs->ndim, tmp.shape, &tmp.ndim));
  tmp.dtype = mid;
  int64_t numel_mid = 1;
  for (int i = tmp.ndim - 1; i >= 0; --i) { tmp.strides[i] = numel_mid; numel_mid *= tmp.shape[i]; }
  FusedLauncher fn = nullptr;
  if (lhs->dtype == mid && rhs->dtype == mid) {
    if (bin_op == HPTB_ADD && hptb_fused_add) fn = hptb_fused_add(mid);
    else if (bin_op == HPTB_SUB && hptb_fused_sub) fn = hptb_fused_sub(mid);
    else if (bin_op == HPTB_MUL && hptb_fused_mul) fn = hptb_fused_mul(mid);
  }
  if (fn && numel_mid > 0) {
    // same plan as hptb_reduce on the (virtual) elementwise result, with both inputs' broadcast strides
    uint8_t mask[HPTB_MAX_DIMS] = {0};
    bool ok = true;
    for (int i = 0; i < naxes && ok; ++i) {
      if (axes[i] < 0 || axes[i] >= tmp.ndim || mask[axes[i]]) ok = false;
      else mask[axes[i]] = 1;
    }
    if (!ok) return fail(HPTB_ERR_AXIS, "binary_reduce: bad axes");
    int64_t oshape[HPTB_MAX_DIMS];
    int on = 0;
    double count = 1.0;
    for (int i = 0; i < tmp.ndim; ++i) {
      if (mask[i]) count *= (double)tmp.shape[i];
      else oshape[on++] = tmp.shape[i];
    }
    if (on == 0) { oshape[0] = 1; on = 1; }
    bool same = out->ndim == on;
    for (int i = 0; same && i < on; ++i) same = out->shape[i] == oshape[i];
    if (!same) return fail(HPTB_ERR_SHAPE, "binary_reduce: out shape does not match the reduced broadcast shape");
    int64_t strides[kMaxOperands][HPTB_MAX_DIMS] = {{0}};
    int k = 0;
    for (int i = 0; i < tmp.ndim; ++i) strides[0][i] = mask[i] ? 0 : out->strides[k++];
    HPTB_TRY(broadcast_strides(*lhs, tmp.shape, tmp.ndim, strides[1]));
    HPTB_TRY(broadcast_strides(*rhs, tmp.shape, tmp.ndim, strides[2]));
    FusedPlan plan;
    collapse(tmp.ndim, tmp.shape, 3, strides, mask, &plan.c);
    plan.lhs = lhs->data;
    plan.rhs = rhs->data;
    plan.out = out->data;
    plan.count = count;
    plan.fold_out = init_out ? 0 : 1;
    plan.red_op = red_op;
    plan.ctx = ctx;
    DeviceGuard g(ctx->device);
    hptb_status st = fn(plan, (cudaStream_t)stream);
    if (st == HPTB_OK) { count_launches(1); return HPTB_OK; }
    if (st != HPTB_FALLBACK) return st;
  }
  // composition through a temporary: exactly the two calls the fused kernel stands for
  Scratch scratch;
  HPTB_TRY(scratch.get(ctx, (size_t)(numel_mid > 0 ? numel_mid : 1) * dtype_size(mid), stream));
  tmp.data = scratch.ptr;
  HPTB_TRY(hptb_binary(ctx, bin_op, lhs, rhs, &tmp, stream));
  return hptb_reduce(ctx, red_op, &tmp, axes, naxes, out, init_out, stream);
}
