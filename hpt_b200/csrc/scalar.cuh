// scalar.cuh — device scalar types, Hpt cast semantics and vector load/store helpers.
//
// Cast semantics restate hpt-macros/src/scalar_convert.rs:39-255 (the `Cast` trait = Rust `as`):
//   int→int wraps, float→int saturates with NaN→0, x→bool is `x != 0`, bool→x is 1/0,
//   →f16/bf16 goes through f32 for bool/i8/u8/i16/u16/i32 and through f64 for u32/i64/u64,
//   f16↔bf16 through f32, half→int through f32.
// Arithmetic semantics restate hpt-types/src/scalars/impls.rs:29-70 (wrapping ints),
// hpt-types/src/scalars/_bool.rs:25-63 (bool add=OR, mul=AND, max=OR, min=AND) and
// hpt-types/src/scalars/_bf16.rs:28-66 (half types compute in f32 and round to nearest even).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <type_traits>

namespace hptb {

// bool is stored as one byte; a distinct type keeps it apart from u8 in templates.
struct alignas(1) b8 {
  uint8_t v;
};
typedef __half f16;
typedef __nv_bfloat16 bf16;

template <typename T> struct is_half : std::false_type {};
template <> struct is_half<f16> : std::true_type {};
template <> struct is_half<bf16> : std::true_type {};
template <typename T> struct is_float_t : std::integral_constant<bool, std::is_floating_point<T>::value || is_half<T>::value> {};
template <typename T> struct is_bool_t : std::is_same<T, b8> {};
template <typename T> struct is_int_t : std::integral_constant<bool, std::is_integral<T>::value> {};

// compute ("Intermediate") type: f32 for the half types, otherwise the type itself.
template <typename T> struct compute_of { typedef T type; };
template <> struct compute_of<f16> { typedef float type; };
template <> struct compute_of<bf16> { typedef float type; };
template <typename T> using compute_t = typename compute_of<T>::type;

// ---- float → int, Rust `as`: saturating, NaN → 0 ----------------------------------------------
template <typename I, typename F>
__device__ __forceinline__ I float_to_int_sat(F x) {
  constexpr I lo = std::numeric_limits<I>::min();
  constexpr I hi = std::numeric_limits<I>::max();
  if (x != x) return (I)0;
  if (x <= (F)lo) return lo;
  if (x >= (F)hi) return hi;
  return (I)x;  // in range: truncation toward zero
}

template <typename To, typename From, typename Enable = void> struct Caster;

template <typename To, typename From>
__device__ __forceinline__ To cast(From x) {
  return Caster<To, From>::run(x);
}

// identity
template <typename T> struct Caster<T, T, void> {
  static __device__ __forceinline__ T run(T x) { return x; }
};
// int → int (wrapping), int → f32/f64 (round to nearest)
template <typename To, typename From>
struct Caster<To, From,
              typename std::enable_if<!std::is_same<To, From>::value && is_int_t<From>::value &&
                                      (is_int_t<To>::value || std::is_floating_point<To>::value)>::type> {
  static __device__ __forceinline__ To run(From x) { return static_cast<To>(x); }
};
// f32/f64 → int
template <typename To, typename From>
struct Caster<To, From, typename std::enable_if<is_int_t<To>::value && std::is_floating_point<From>::value>::type> {
  static __device__ __forceinline__ To run(From x) { return float_to_int_sat<To, From>(x); }
};
// f32 ↔ f64
template <> struct Caster<double, float, void> {
  static __device__ __forceinline__ double run(float x) { return (double)x; }
};
template <> struct Caster<float, double, void> {
  static __device__ __forceinline__ float run(double x) { return (float)x; }
};
// half → f32 / f64 / int / other half
template <> struct Caster<float, f16, void> {
  static __device__ __forceinline__ float run(f16 x) { return __half2float(x); }
};
template <> struct Caster<float, bf16, void> {
  static __device__ __forceinline__ float run(bf16 x) { return __bfloat162float(x); }
};
template <> struct Caster<double, f16, void> {
  static __device__ __forceinline__ double run(f16 x) { return (double)__half2float(x); }
};
template <> struct Caster<double, bf16, void> {
  static __device__ __forceinline__ double run(bf16 x) { return (double)__bfloat162float(x); }
};
template <typename To, typename From>
struct Caster<To, From, typename std::enable_if<is_int_t<To>::value && is_half<From>::value>::type> {
  static __device__ __forceinline__ To run(From x) { return float_to_int_sat<To, float>(cast<float>(x)); }
};
template <> struct Caster<f16, bf16, void> {
  static __device__ __forceinline__ f16 run(bf16 x) { return __float2half_rn(__bfloat162float(x)); }
};
template <> struct Caster<bf16, f16, void> {
  static __device__ __forceinline__ bf16 run(f16 x) { return __float2bfloat16_rn(__half2float(x)); }
};
// f32 / f64 → half
template <> struct Caster<f16, float, void> {
  static __device__ __forceinline__ f16 run(float x) { return __float2half_rn(x); }
};
template <> struct Caster<bf16, float, void> {
  static __device__ __forceinline__ bf16 run(float x) { return __float2bfloat16_rn(x); }
};
template <> struct Caster<f16, double, void> {
  static __device__ __forceinline__ f16 run(double x) { return __double2half(x); }
};
template <> struct Caster<bf16, double, void> {
  static __device__ __forceinline__ bf16 run(double x) { return __double2bfloat16(x); }
};
// int → half: via f32 for ≤32-bit signed and ≤16-bit unsigned, via f64 for u32/i64/u64
template <typename To, typename From>
struct Caster<To, From, typename std::enable_if<is_half<To>::value && is_int_t<From>::value>::type> {
  static __device__ __forceinline__ To run(From x) {
    constexpr bool via_f64 = sizeof(From) == 8 || std::is_same<From, uint32_t>::value;
    if (via_f64) return cast<To>((double)x);
    return cast<To>((float)x);
  }
};
// bool → anything
template <typename To>
struct Caster<To, b8, typename std::enable_if<!is_bool_t<To>::value>::type> {
  static __device__ __forceinline__ To run(b8 x) {
    if constexpr (is_half<To>::value) return cast<To>(x.v ? 1.0f : 0.0f);
    else return x.v ? (To)1 : (To)0;
  }
};
// anything → bool
template <typename From>
struct Caster<b8, From, typename std::enable_if<!is_bool_t<From>::value>::type> {
  static __device__ __forceinline__ b8 run(From x) {
    b8 r;
    if constexpr (is_half<From>::value) r.v = cast<float>(x) != 0.0f;
    else r.v = x != (From)0;
    return r;
  }
};

// widen to / narrow from the compute type
template <typename T>
__device__ __forceinline__ compute_t<T> to_compute(T x) {
  return cast<compute_t<T>>(x);
}
template <typename T>
__device__ __forceinline__ T from_compute(compute_t<T> x) {
  return cast<T>(x);
}

// ---- identities used by reductions (hpt-types/src/dtype.rs:333-480: INF/NEG_INF are MAX/MIN for ints) ----
template <typename T> struct Limits {
  static __device__ __forceinline__ T lowest() { return std::numeric_limits<T>::lowest(); }
  static __device__ __forceinline__ T highest() { return std::numeric_limits<T>::max(); }
};
template <> struct Limits<float> {
  static __device__ __forceinline__ float lowest() { return -__int_as_float(0x7f800000); }
  static __device__ __forceinline__ float highest() { return __int_as_float(0x7f800000); }
};
template <> struct Limits<double> {
  static __device__ __forceinline__ double lowest() { return -__longlong_as_double(0x7ff0000000000000LL); }
  static __device__ __forceinline__ double highest() { return __longlong_as_double(0x7ff0000000000000LL); }
};
template <> struct Limits<b8> {
  static __device__ __forceinline__ b8 lowest() { return b8{0}; }
  static __device__ __forceinline__ b8 highest() { return b8{1}; }
};

// ---- vector load / store of N elements of T (N*sizeof(T) bytes, naturally aligned up to 16 B) -------------
template <int BYTES> struct VecBytes;
template <> struct VecBytes<1> { typedef uint8_t type; };
template <> struct VecBytes<2> { typedef uint16_t type; };
template <> struct VecBytes<4> { typedef uint32_t type; };
template <> struct VecBytes<8> { typedef uint2 type; };
template <> struct VecBytes<16> { typedef uint4 type; };

template <typename T, int N>
struct alignas((sizeof(T) * N) > 16 ? 16 : (sizeof(T) * N)) Pack {
  T v[N];
};

// streaming (read-once) loads: ld.global.nc with no L1 allocation
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ldg_stream(const uint2* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ldg_stream(const uint32_t* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ uint16_t ldg_stream(const uint16_t* p) {
  uint16_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ uint8_t ldg_stream(const uint8_t* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(r) : "l"(p));
  return (uint8_t)r;
}
__device__ __forceinline__ void stg_stream(uint4* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream(uint2* p, uint2 v) {
  asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void stg_stream(uint32_t* p, uint32_t v) {
  asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void stg_stream(uint16_t* p, uint16_t v) {
  asm volatile("st.global.L1::no_allocate.u16 [%0], %1;" ::"l"(p), "h"(v) : "memory");
}
__device__ __forceinline__ void stg_stream(uint8_t* p, uint8_t v) {
  asm volatile("st.global.L1::no_allocate.u8 [%0], %1;" ::"l"(p), "r"((uint32_t)v) : "memory");
}

template <typename T, int N>
__device__ __forceinline__ void load_pack(Pack<T, N>& dst, const T* src) {
  constexpr int bytes = sizeof(T) * N;
  constexpr int chunk = bytes > 16 ? 16 : bytes;
  typedef typename VecBytes<chunk>::type V;
  const V* s = reinterpret_cast<const V*>(src);
  V* d = reinterpret_cast<V*>(&dst);
#pragma unroll
  for (int i = 0; i < bytes / chunk; ++i) d[i] = ldg_stream(s + i);
}
// re-read operands (a row broadcast over the outer dims): plain read-only loads that DO allocate in L1, so the
// repeats are served by the SM's own L1 instead of crossing to L2 every time
template <typename T, int N>
__device__ __forceinline__ void load_pack_cached(Pack<T, N>& dst, const T* src) {
  constexpr int bytes = sizeof(T) * N;
  constexpr int chunk = bytes > 16 ? 16 : bytes;
  typedef typename VecBytes<chunk>::type V;
  const V* s = reinterpret_cast<const V*>(src);
  V* d = reinterpret_cast<V*>(&dst);
#pragma unroll
  for (int i = 0; i < bytes / chunk; ++i) d[i] = __ldg(s + i);
}
template <typename T, int N>
__device__ __forceinline__ void store_pack(T* dst, const Pack<T, N>& src) {
  constexpr int bytes = sizeof(T) * N;
  constexpr int chunk = bytes > 16 ? 16 : bytes;
  typedef typename VecBytes<chunk>::type V;
  V* d = reinterpret_cast<V*>(dst);
  const V* s = reinterpret_cast<const V*>(&src);
#pragma unroll
  for (int i = 0; i < bytes / chunk; ++i) stg_stream(d + i, s[i]);
}
template <typename T>
__device__ __forceinline__ T load_one(const T* src) {
  Pack<T, 1> p;
  load_pack<T, 1>(p, src);
  return p.v[0];
}

// exp for the streaming kernels (softmax, logsumexp): ex2(x·log2e) with the rounding error of the product folded back in
// (hi + lo = x·log2e to ~2^-48; exp = ex2(hi)·(1 + lo·ln2)), 7 instructions against libdevice expf's 12, same
// accuracy class (MUFU.EX2 ≤ 2 ulp).  Inputs below −110 (including −inf: masked logits, padding lanes) give 0;
// NaN propagates.
__device__ __forceinline__ float fast_expf(float x) {
  x = x < -110.0f ? -110.0f : x;  // a select, not fmaxf: NaN must survive
  const float hi = x * 1.4426950408889634f;
  const float lo = fmaf(x, 1.4426950408889634f, -hi) + x * 1.9259629911266175e-8f;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(hi));
  return fmaf(e, lo * 0.6931471805599453f, e);
}
// the same for arguments that may overflow (naive logsumexp): +inf stays +inf instead of inf·lo − inf = NaN
__device__ __forceinline__ float fast_expf_ovf(float x) {
  const float xc = x < -110.0f ? -110.0f : x;
  const float hi = xc * 1.4426950408889634f;
  const float lo = fmaf(xc, 1.4426950408889634f, -hi) + xc * 1.9259629911266175e-8f;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(hi));
  const float r = fmaf(e, lo * 0.6931471805599453f, e);
  return e > 3.0e38f ? e : r;
}

// C++ type ↔ hptb_dtype
template <typename T> struct dtype_of;
#define HPTB_DTYPE_OF(T, E) \
  template <> struct dtype_of<T> { static constexpr int value = E; };
HPTB_DTYPE_OF(b8, 0) HPTB_DTYPE_OF(int8_t, 1) HPTB_DTYPE_OF(int16_t, 2) HPTB_DTYPE_OF(int32_t, 3)
HPTB_DTYPE_OF(int64_t, 4) HPTB_DTYPE_OF(uint8_t, 5) HPTB_DTYPE_OF(uint16_t, 6) HPTB_DTYPE_OF(uint32_t, 7)
HPTB_DTYPE_OF(uint64_t, 8) HPTB_DTYPE_OF(f16, 9) HPTB_DTYPE_OF(bf16, 10) HPTB_DTYPE_OF(float, 11)
HPTB_DTYPE_OF(double, 12)
#undef HPTB_DTYPE_OF

template <int E> struct type_of_dtype;
#define HPTB_TYPE_OF(E, T) \
  template <> struct type_of_dtype<E> { typedef T type; };
HPTB_TYPE_OF(0, b8) HPTB_TYPE_OF(1, int8_t) HPTB_TYPE_OF(2, int16_t) HPTB_TYPE_OF(3, int32_t)
HPTB_TYPE_OF(4, int64_t) HPTB_TYPE_OF(5, uint8_t) HPTB_TYPE_OF(6, uint16_t) HPTB_TYPE_OF(7, uint32_t)
HPTB_TYPE_OF(8, uint64_t) HPTB_TYPE_OF(9, f16) HPTB_TYPE_OF(10, bf16) HPTB_TYPE_OF(11, float)
HPTB_TYPE_OF(12, double)
#undef HPTB_TYPE_OF

}  // namespace hptb
