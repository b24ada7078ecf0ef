// common.h — host-side shared definitions for libhpt_b200 (no CUDA device code in here).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/hpt_b200.h"

namespace hptb {

// thread-local error message behind hptb_last_error()
void set_error(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
hptb_status fail(hptb_status st, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
const char* last_error();

#define HPTB_CUDA_CHECK(expr)                                                                      \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return ::hptb::fail(HPTB_ERR_CUDA, "%s failed: %s (%d) at %s:%d", #expr, cudaGetErrorString(_e), \
                          (int)_e, __FILE__, __LINE__);                                            \
  } while (0)

#define HPTB_TRY(expr)                 \
  do {                                 \
    hptb_status _s = (expr);           \
    if (_s != HPTB_OK) return _s;      \
  } while (0)

inline size_t dtype_size(int dt) {
  switch (dt) {
    case HPTB_BOOL: case HPTB_I8: case HPTB_U8: return 1;
    case HPTB_I16: case HPTB_U16: case HPTB_F16: case HPTB_BF16: return 2;
    case HPTB_I32: case HPTB_U32: case HPTB_F32: return 4;
    case HPTB_I64: case HPTB_U64: case HPTB_F64: return 8;
    default: return 0;
  }
}
inline bool dtype_valid(int dt) { return dt >= 0 && dt < HPTB_DTYPE_COUNT; }
inline bool dtype_is_float(int dt) { return dt >= HPTB_F16 && dt <= HPTB_F64; }
const char* dtype_name(int dt);

inline int64_t numel(const hptb_tensor& t) {
  int64_t n = 1;
  for (int i = 0; i < t.ndim; ++i) n *= t.shape[i];
  return n;
}

hptb_status validate_tensor(const hptb_tensor* t, const char* what);

// promotion tables (promote.cpp)
int promote(int lhs, int rhs, int kind);

// 32-bit magic-number division: q = n / d for 0 <= n < 2^32, 1 <= d < 2^31.
// (Own derivation: round-up method, m = ceil(2^(32+s)/d) - 2^32 with s = ceil(log2 d).)
struct FastDiv {
  uint32_t d, m, s;
  FastDiv() : d(1), m(0), s(0) {}
  explicit FastDiv(uint32_t div) : d(div) {
    if (div <= 1) { d = 1; m = 0; s = 0; return; }
    uint32_t l = 0;
    while ((1ull << l) < div) ++l;  // ceil(log2 d)
    s = l;
    uint64_t p = 1ull << l;
    m = (uint32_t)(((p - div) << 32) / div + 1);
  }
#ifdef __CUDACC__
  __host__ __device__
#endif
  inline uint32_t div(uint32_t n) const {
#ifdef __CUDA_ARCH__
    uint32_t t = __umulhi(n, m);
#else
    uint32_t t = (uint32_t)(((uint64_t)n * m) >> 32);
#endif
    // (t + n) can overflow 32 bits: use the (n - t)/2 + t form
    return (((n - t) >> (s ? 1 : 0)) + t) >> (s ? s - 1 : 0);
  }
};

}  // namespace hptb
