// elementwise_dyn.cuh — the runtime-typed elementwise kernel.
//
// Hpt's NormalOut / FloatOutBinary / Cast semantics are "cast every input to the promoted Output type, then
// run the op in Output" (hpt-macros/src/normal_out.rs:72-135, hpt-macros/src/scalar_convert.rs:39-255).  So a
// kernel only has to be specialised on the OUTPUT type: the inputs' element types are launch parameters, the
// loads are raw byte moves (switch on element size, all issued before anything is converted so the memory-level
// parallelism of the specialised kernels is kept) and the conversion is a warp-uniform switch on the dtype.
// One kernel per output dtype replaces the 13×13 (lhs, rhs) grid of per-pair kernels the reference instantiates
// (hpt-cudakernels/src/binary/binary_template.cuh:107-138: `{op}_{A}_{B}_contiguous` …) — the operator is a
// launch parameter as well (uniform switch per pack).
//
// It serves every mixed-dtype pair (e.g. f32 ⊕ i64 → f64, BASELINE config 4), FloatUnaryOps on integer inputs,
// dtype-converting copies, and any layout the vector-only specialised kernels hand back (unaligned views, odd
// row lengths, strided inner dims): VEC elements per thread with 128-bit-or-narrower accesses when every
// operand's inner stride is 1 (or 0) and aligned, one element per thread with arbitrary strides otherwise.
#pragma once
#include "elementwise.cuh"
#include "ops.cuh"

namespace hptb {

struct DynExtra {
  int32_t dtype[2];  // hptb_dtype of the inputs
  int32_t esz[2];    // their element sizes in bytes
  int32_t op;        // hptb_binary_op / hptb_unary_op (ignored by casts)
  int32_t pad;
  double alpha, beta;
};

// raw storage for N elements of up to MAXB bytes each (MAXB = 4 or 8)
template <int N, int MAXB = 8> struct alignas(16) RawPack {
  uint32_t w[(N * MAXB + 3) / 4 < 2 ? 2 : (N * MAXB + 3) / 4];
};

// N elements of `esz` bytes from p (aligned to min(16, N·esz)); cached = allocate in L1 (re-read operand)
template <int N, int MAXB>
__device__ __forceinline__ void load_raw(RawPack<N, MAXB>& r, const unsigned char* p, int esz, bool cached) {
#define HPTB_RAW_CASE(ESZ)                                                                     \
  {                                                                                            \
    constexpr int bytes = N * ESZ;                                                             \
    constexpr int chunk = bytes > 16 ? 16 : bytes;                                             \
    typedef typename VecBytes<chunk>::type V;                                                  \
    const V* s = reinterpret_cast<const V*>(p);                                                \
    V* d = reinterpret_cast<V*>(r.w);                                                          \
    _Pragma("unroll") for (int i = 0; i < bytes / chunk; ++i) d[i] = cached ? __ldg(s + i) : ldg_stream(s + i); \
  }
  switch (esz) {
    case 1: HPTB_RAW_CASE(1) break;
    case 2: HPTB_RAW_CASE(2) break;
    case 4: HPTB_RAW_CASE(4) break;
    default:
      if constexpr (MAXB >= 8) HPTB_RAW_CASE(8)
      break;
  }
#undef HPTB_RAW_CASE
}

// place one element (in `one`) at position k of a raw pack (k is a compile-time constant after unrolling)
template <int N, int MAXB>
__device__ __forceinline__ void raw_insert(RawPack<N, MAXB>& r, int k, const RawPack<1, MAXB>& one, int esz) {
  switch (esz) {
    case 1: { const uint32_t sh = (k & 3) * 8; r.w[k >> 2] = (r.w[k >> 2] & ~(0xffu << sh)) | ((one.w[0] & 0xffu) << sh); } break;
    case 2: { const uint32_t sh = (k & 1) * 16; r.w[k >> 1] = (r.w[k >> 1] & ~(0xffffu << sh)) | ((one.w[0] & 0xffffu) << sh); } break;
    case 4: r.w[k] = one.w[0]; break;
    default:
      if constexpr (MAXB >= 8) { r.w[2 * k] = one.w[0]; r.w[2 * k + 1] = one.w[1]; }
      break;
  }
}

// convert the N raw elements of runtime dtype `dt` to O (Hpt `Cast` semantics)
template <typename O, int N, int MAXB>
__device__ __forceinline__ void unpack_raw(O (&v)[N], const RawPack<N, MAXB>& r, int dt) {
  switch (dt) {
#define X(T, NAME, E)                                                              \
  case E: {                                                                        \
    if constexpr (sizeof(T) <= MAXB) {                                             \
      const T* t = reinterpret_cast<const T*>(r.w);                                \
      _Pragma("unroll") for (int k = 0; k < N; ++k) v[k] = cast<O>(t[k]);          \
    }                                                                              \
  } break;
    HPTB_FOR_DTYPES(X)
#undef X
    default: break;
  }
}

// the same for all UNROLL packs of one operand under a single switch
template <typename O, int N, int U, int MAXB>
__device__ __forceinline__ void unpack_all(O (&v)[U][N], const RawPack<N, MAXB> (&r)[U], int dt) {
  switch (dt) {
#define X(T, NAME, E)                                                                \
  case E: {                                                                          \
    if constexpr (sizeof(T) <= MAXB) {                                               \
      _Pragma("unroll") for (int u = 0; u < U; ++u) {                                \
        const T* t = reinterpret_cast<const T*>(r[u].w);                             \
        _Pragma("unroll") for (int k = 0; k < N; ++k) v[u][k] = cast<O>(t[k]);       \
      }                                                                              \
    }                                                                                \
  } break;
    HPTB_FOR_DTYPES(X)
#undef X
    default: break;
  }
}

// ---- operators with the op code as a launch parameter ---------------------------------------------------
template <typename O>
struct DynBinaryFn {
  typedef compute_t<O> C;
  static __device__ __forceinline__ O apply(O a, O b, const DynExtra& x) {
    const C ca = to_compute<O>(a), cb = to_compute<O>(b);
    C r;
    switch (x.op) {
      case HPTB_ADD: r = OpAdd::apply<C>(ca, cb); break;
      case HPTB_SUB: r = OpSub::apply<C>(ca, cb); break;
      case HPTB_MUL: r = OpMul::apply<C>(ca, cb); break;
      case HPTB_REM: r = OpRem::apply<C>(ca, cb); break;
      case HPTB_DIV: r = OpDiv::apply<C>(ca, cb); break;
      case HPTB_MAXIMUM: r = OpMax::apply<C>(ca, cb); break;
      case HPTB_MINIMUM: r = OpMin::apply<C>(ca, cb); break;
      case HPTB_POW: r = OpPow::apply<C>(ca, cb); break;      // float outputs only (host-checked)
      case HPTB_HYPOT: r = OpHypot::apply<C>(ca, cb); break;
      case HPTB_BITAND: r = OpBitAnd::apply<C>(ca, cb); break;  // bool / integer outputs only
      case HPTB_BITOR: r = OpBitOr::apply<C>(ca, cb); break;
      case HPTB_BITXOR: r = OpBitXor::apply<C>(ca, cb); break;
      case HPTB_SHL: r = OpShl::apply<C>(ca, cb); break;
      default: r = OpShr::apply<C>(ca, cb); break;
    }
    return from_compute<O>(r);
  }
};

// scalar parameter of a unary op in the compute type (integers: Rust `as` from f64; bool: != 0)
template <typename C>
__device__ __forceinline__ C dyn_param(double v) {
  if constexpr (is_bool_t<C>::value) return b8{(uint8_t)(v != 0.0)};
  else if constexpr (std::is_integral<C>::value) return float_to_int_sat<C, double>(v);
  else return (C)v;
}

template <typename O>
struct DynUnaryFn {
  typedef compute_t<O> C;
  static __device__ __forceinline__ O apply(O a, O, const DynExtra& x) {
    const C c = to_compute<O>(a), al = dyn_param<C>(x.alpha), be = dyn_param<C>(x.beta);
    if (x.op >= HPTB_FLOAT_UNARY_COUNT) return from_compute<O>(normal_unary<C>(x.op, c, al, be));
    if constexpr (std::is_floating_point<C>::value) {
      C r;
      switch (x.op) {
#define XU(NAME, E) \
  case E: r = UnaryOp<E>::apply(c, al, be); break;
        HPTB_FOR_UNARY_OPS(XU)
#undef XU
        default: r = c; break;
      }
      return from_compute<O>(r);
    } else {
      return a;  // FloatUnaryOps never produce an integer output (host-checked)
    }
  }
};

// TensorCmp: the kernel is specialised on the PROMOTED type P (both inputs are cast to it), stores bool
template <typename P>
struct DynCmpFn {
  static __device__ __forceinline__ b8 apply(P a, P b, const DynExtra& x) { return cmp_apply<P>(x.op, a, b); }
};

template <typename O>
struct DynCastFn {
  static __device__ __forceinline__ O apply(O a, O, const DynExtra&) { return a; }
};

// ---- kernel ------------------------------------------------------------------------------------------------
// O = the type both inputs are converted to and the operator runs in; S = the stored type (O, or bool for compares)
template <int NIN, int VEC, int UNROLL, int MAXB, typename Fn, typename O, typename S = O>
__global__ void __launch_bounds__(kMapThreads)
map_dyn_kernel(S* __restrict__ out, const unsigned char* __restrict__ a, const unsigned char* __restrict__ b, RowsParams p,
               DynExtra x) {
  pdl_prologue();
  const int64_t c0 = (int64_t)blockIdx.x * (kMapThreads * UNROLL) + threadIdx.x;
  RawPack<VEC, MAXB> ra[UNROLL], rb[UNROLL];
  int64_t oo[UNROLL];
  int32_t cnt[UNROLL];
  const int esa = x.esz[0], esb = NIN == 2 ? x.esz[1] : 1;
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    const int64_t c = c0 + (int64_t)u * kMapThreads;
    cnt[u] = 0;
    if (c < p.total_chunks) {
      int64_t row = 0, col = c;
      int64_t off[3] = {0, 0, 0};
      if (p.nouter > 0) {
        if (!p.use64) { row = p.cpr_div.div((uint32_t)c); col = c - row * p.cpr; }
        else { row = c / p.cpr; col = c - row * p.cpr; }
        walk_outer<3>(row, p.nouter, p.use64, p.outer_shape, p.outer_div, p.outer_stride, off);
      }
      const int64_t e = col * VEC;
      const int64_t left = p.inner - e;
      cnt[u] = left >= VEC ? VEC : (int32_t)left;
      oo[u] = off[0] + e * p.inner_stride[0];
      const unsigned char* ap = a + (off[1] + e * p.inner_stride[1]) * esa;
      if (VEC == 1 || p.inner_stride[1] == 0) {
        RawPack<1, MAXB> one = {{0u, 0u}};
        load_raw<1, MAXB>(one, ap, esa, true);
        ra[u].w[0] = one.w[0];
        ra[u].w[1] = one.w[1];
      } else if (cnt[u] == VEC) {
        load_raw<VEC, MAXB>(ra[u], ap, esa, p.reuse[1] != 0);
      } else {  // ragged tail of a 1-D tensor: element-wise
#pragma unroll
        for (int k = 0; k < (int)(sizeof(ra[u].w) / 4); ++k) ra[u].w[k] = 0u;
#pragma unroll
        for (int k = 0; k < VEC; ++k)
          if (k < cnt[u]) {
            RawPack<1, MAXB> one = {{0u, 0u}};
            load_raw<1, MAXB>(one, ap + k * esa, esa, true);
            raw_insert<VEC, MAXB>(ra[u], k, one, esa);
          }
      }
      if constexpr (NIN == 2) {
        const unsigned char* bp = b + (off[2] + e * p.inner_stride[2]) * esb;
        if (VEC == 1 || p.inner_stride[2] == 0) {
          RawPack<1, MAXB> one = {{0u, 0u}};
          load_raw<1, MAXB>(one, bp, esb, true);
          rb[u].w[0] = one.w[0];
          rb[u].w[1] = one.w[1];
        } else if (cnt[u] == VEC) {
          load_raw<VEC, MAXB>(rb[u], bp, esb, p.reuse[2] != 0);
        } else {
#pragma unroll
          for (int k = 0; k < (int)(sizeof(rb[u].w) / 4); ++k) rb[u].w[k] = 0u;
#pragma unroll
          for (int k = 0; k < VEC; ++k)
            if (k < cnt[u]) {
              RawPack<1, MAXB> one = {{0u, 0u}};
              load_raw<1, MAXB>(one, bp + k * esb, esb, true);
              raw_insert<VEC, MAXB>(rb[u], k, one, esb);
            }
        }
      }
    }
  }
  // convert: ONE warp-uniform switch per operand per thread (not per pack), then the operator
  O va[UNROLL][VEC], vb[UNROLL][VEC];
  unpack_all<O, VEC, UNROLL, MAXB>(va, ra, x.dtype[0]);
  if constexpr (NIN == 2) unpack_all<O, VEC, UNROLL, MAXB>(vb, rb, x.dtype[1]);
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    if (cnt[u] == 0) continue;
    if (VEC > 1 && p.inner_stride[1] == 0) {
#pragma unroll
      for (int k = 1; k < VEC; ++k) va[u][k] = va[u][0];
    }
    if constexpr (NIN == 2) {
      if (VEC > 1 && p.inner_stride[2] == 0) {
#pragma unroll
        for (int k = 1; k < VEC; ++k) vb[u][k] = vb[u][0];
      }
    }
    Pack<S, VEC> po;
#pragma unroll
    for (int k = 0; k < VEC; ++k) po.v[k] = Fn::apply(va[u][k], NIN == 2 ? vb[u][k] : va[u][k], x);
    if (cnt[u] == VEC) store_pack<S, VEC>(out + oo[u], po);
    else {
#pragma unroll
      for (int k = 0; k < VEC; ++k)
        if (k < cnt[u]) out[oo[u] + k] = po.v[k];
    }
  }
}

// elements per thread: the OUTPUT does 16-byte stores (8-byte for 1- and 2-byte outputs, which keeps the raw
// input pack at ≤ 32 bytes even for an 8-byte source)
template <typename O>
constexpr int dyn_vec_width() {
  return sizeof(O) == 8 ? 2 : 4;
}

// MAXB = the widest input element the launcher can be handed: 8 for casts; for binary / unary outputs the
// promotion tables never pair a narrower Output with an 8-byte input (but i32/u32 ⊕ f16 → f16 is a 4-byte input)
template <int NIN, int MAXB, typename Fn, typename O, typename S = O>
hptb_status launch_map_dyn(const MapPlan& plan, cudaStream_t stream) {
  const Collapsed& c = plan.c;
  if (c.numel == 0) return HPTB_OK;
  S* out = static_cast<S*>(plan.ptr[0]);
  const unsigned char* a = static_cast<const unsigned char*>(plan.ptr[1]);
  const unsigned char* b = NIN == 2 ? static_cast<const unsigned char*>(plan.ptr[2]) : a;
  constexpr int VEC = dyn_vec_width<O>();
  DynExtra x;
  memset(&x, 0, sizeof(x));
  size_t esz[3] = {sizeof(S), 1, 1};
  for (int i = 0; i < NIN; ++i) {
    if (!dtype_valid(plan.in_dtype[i])) return fail(HPTB_ERR_INVALID, "elementwise: bad input dtype");
    x.dtype[i] = plan.in_dtype[i];
    x.esz[i] = (int32_t)dtype_size(plan.in_dtype[i]);
    if (x.esz[i] > MAXB) return fail(HPTB_ERR_DTYPE, "elementwise: %s input is wider than this kernel accepts", dtype_name(plan.in_dtype[i]));
    esz[i + 1] = dtype_size(plan.in_dtype[i]);
  }
  x.op = plan.op;
  x.alpha = plan.alpha;
  x.beta = plan.beta;

  RowsParams p;
  memset(&p, 0, sizeof(p));
  const int nd = c.ndim;
  p.inner = nd ? c.shape[nd - 1] : 1;
  p.nouter = nd ? nd - 1 : 0;
  bool big = false;
  for (int o = 0; o <= NIN; ++o) {
    int64_t s = nd ? c.strides[o][nd - 1] : 1;
    if (s > 0x7fffffffLL || s < -0x7fffffffLL) return fail(HPTB_ERR_UNSUPPORTED, "elementwise: inner stride exceeds 31 bits");
    p.inner_stride[o] = (int32_t)s;
  }
  for (int i = 0; i < p.nouter; ++i) {
    int d = nd - 2 - i;
    if (!fits_u32(c.shape[d])) big = true;
    p.outer_shape[i] = (uint32_t)c.shape[d];
    p.outer_div[i] = FastDiv((uint32_t)c.shape[d]);
    for (int o = 0; o <= NIN; ++o) p.outer_stride[o][i] = c.strides[o][d];
  }
  bool vec_ok = p.inner_stride[0] == 1;
  for (int o = 0; o <= NIN && vec_ok; ++o) {
    if (p.inner_stride[o] == 0 && o > 0) continue;
    if (p.inner_stride[o] != 1) { vec_ok = false; break; }
    size_t align = esz[o] * VEC > 16 ? 16 : esz[o] * VEC;
    if (reinterpret_cast<uintptr_t>(plan.ptr[o]) % align) vec_ok = false;
    for (int i = 0; i < p.nouter; ++i) {
      if ((uint64_t)(std::llabs(p.outer_stride[o][i]) * (int64_t)esz[o]) % align) vec_ok = false;
      if (o > 0 && p.outer_stride[o][i] == 0) p.reuse[o] = 1;
    }
  }
  if (vec_ok && p.nouter > 0 && p.inner % VEC) vec_ok = false;
  const int vec = vec_ok ? VEC : 1;
  p.cpr = (p.inner + vec - 1) / vec;
  int64_t rows = 1;
  for (int i = 0; i < p.nouter; ++i) rows *= c.shape[nd - 2 - i];
  p.total_chunks = rows * p.cpr;
  if (!fits_u32(p.cpr) || p.total_chunks >= (int64_t(1) << 32)) big = true;
  p.use64 = big ? 1 : 0;
  p.cpr_div = FastDiv(big ? 1u : (uint32_t)p.cpr);
  constexpr int UNROLL = 4;
  int64_t blocks = (p.total_chunks + kMapThreads * UNROLL - 1) / (kMapThreads * UNROLL);
  if (blocks > 0x7fffffffLL) return fail(HPTB_ERR_UNSUPPORTED, "elementwise: tensor too large for one launch");
  if (vec_ok) HPTB_CUDA_CHECK(launch_kernel(map_dyn_kernel<NIN, VEC, UNROLL, MAXB, Fn, O, S>, dim3((unsigned)blocks), dim3(kMapThreads), 0, stream, out, a, b, p, x));
  else HPTB_CUDA_CHECK(launch_kernel(map_dyn_kernel<NIN, 1, UNROLL, MAXB, Fn, O, S>, dim3((unsigned)blocks), dim3(kMapThreads), 0, stream, out, a, b, p, x));
  return HPTB_OK;
}

}  // namespace hptb
