// oracle_cpu.cpp — C++17/OpenMP restatement of Hpt's CPU path for the benchmarked configurations.
//
// TEST / BASELINE INFRASTRUCTURE ONLY.  Nothing under hpt_b200/ links or loads this; it is built into
// oracle/_build/liboracle_cpu.so and used (a) by tests/ as a second checker beside oracle/hpt_oracle.py and
// (b) by bench.py as the timed CPU baseline (`cpu_baseline`, kind "port") and as `--impl reference`.
// Hpt itself is Rust and cannot be built in this image (no cargo/rustc), so this is a
// restatement, not the reference binary: "parity pinned" only through the checks listed in
// oracle/hpt_oracle.py's header.
//
// Structure kept from the reference (paths relative to the Hpt repository):
//   contiguous binary   hpt/src/backends/cpu/utils/binary/binary_normal.rs:202-252  rayon par_chunks of one SIMD
//                       vector → OpenMP static chunks + `omp simd`
//   broadcast binary    binary_normal.rs:253-280 + hpt-iterator/src/par_strided.rs:59-100: rows split over
//                       threads, inner loop over the last dim with per-operand last stride
//   unary               hpt/src/backends/cpu/utils/unary/unary.rs:25-98: vector body + scalar tail; the
//                       transcendental is a ~1-ulp vector routine (Hpt: SLEEF-u10 port,
//                       hpt-types/src/vectors/arch_simd/_256bit/avx2/f32x8.rs:181-183; here glibc libmvec)
//   reductions          hpt/src/backends/cpu/utils/reduce/reduce.rs:367-…, cpu/kernels/reduce.rs: one thread per
//                       block of output rows, inner-axis SIMD accumulate, accumulation in the OUTPUT dtype
//   argmax/argmin       hpt/src/backends/cpu/kernels/argreduce_kernels.rs:2-71: scalar strict-compare scan
//   softmax             hpt/src/backends/cpu/kernels/softmax.rs:204-310: per-row max → exp → Σ → divide
//   logsumexp           hpt/src/backends/cpu/tensor_internal/common_reduce.rs:451-480: ln Σ exp(x), no shift
//
// For the permuted unary of config 2 the LOGICAL result is produced (the reference's CPU path walks
// `a.t().sin()` in memory order, a known defect not copied: SURVEY.md §8c).
#include <immintrin.h>
#include <omp.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

// glibc libmvec AVX2 entry points (vector ABI names); ≤ 4 ulp by glibc's statement, ~1 ulp in practice for
// sinf/expf — the stand-in for Hpt's SLEEF-u10 port.
extern "C" __m256 _ZGVdN8v_sinf(__m256);
extern "C" __m256 _ZGVdN8v_expf(__m256);

static inline void vec_unary(int op, const float* src, float* dst, int64_t n) {
  int64_t i = 0;
  if (op == 0) {
    for (; i + 8 <= n; i += 8) _mm256_storeu_ps(dst + i, _ZGVdN8v_sinf(_mm256_loadu_ps(src + i)));
    for (; i < n; ++i) dst[i] = sinf(src[i]);
  } else {
    for (; i + 8 <= n; i += 8) _mm256_storeu_ps(dst + i, _ZGVdN8v_expf(_mm256_loadu_ps(src + i)));
    for (; i < n; ++i) dst[i] = expf(src[i]);
  }
}

extern "C" {

int orc_num_threads(void) { return omp_get_max_threads(); }
void orc_set_num_threads(int n) { omp_set_num_threads(n); }

// ---- unary, contiguous: op 0 = sin, 1 = exp ---------------------------------------------------------------
void orc_unary_f32(int op, const float* in, float* out, int64_t n) {
#pragma omp parallel
  {
    int nt = omp_get_num_threads(), t = omp_get_thread_num();
    int64_t per = ((n + nt - 1) / nt + 7) / 8 * 8;
    int64_t lo = per * t, hi = lo + per < n ? lo + per : n;
    if (lo < hi) vec_unary(op, in + lo, out + lo, hi - lo);
  }
}

// unary over a 2-D strided view (rows x cols, element strides sr, sc) into a contiguous [rows, cols] output.
// Rows of the OUTPUT are split over threads; each thread gathers a row of the view (per-operand last stride,
// as hpt-iterator's strided walk does), then runs the vector body on it.
void orc_unary_f32_strided2d(int op, const float* in, int64_t rows, int64_t cols, int64_t sr, int64_t sc, float* out) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < rows; ++r) {
    const float* src = in + r * sr;
    float* dst = out + r * cols;
    if (sc == 1) { vec_unary(op, src, dst, cols); continue; }
    for (int64_t c = 0; c < cols; ++c) dst[c] = src[c * sc];
    vec_unary(op, dst, dst, cols);
  }
}

// ---- binary ------------------------------------------------------------------------------------------------
// out[r, c] = a[r, c] + b[c]   (config 1: [R, C] + [1, C])
void orc_add_f32_bcast_row(const float* a, const float* b, float* out, int64_t rows, int64_t cols) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < rows; ++r) {
    const float* ar = a + r * cols;
    float* o = out + r * cols;
#pragma omp simd
    for (int64_t c = 0; c < cols; ++c) o[c] = ar[c] + b[c];
  }
}
// out[r, c] = (f64)a[r, c] + (f64)k[c]   (config 4: f32 ⊕ i64 → f64, both cast to Output first)
void orc_add_f32_i64_bcast_row(const float* a, const int64_t* k, double* out, int64_t rows, int64_t cols) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < rows; ++r) {
    const float* ar = a + r * cols;
    double* o = out + r * cols;
#pragma omp simd
    for (int64_t c = 0; c < cols; ++c) o[c] = (double)ar[c] + (double)k[c];
  }
}

// ---- reductions ----------------------------------------------------------------------------------------------
// sum over the last axis of [rows, cols] (f32 accumulate, as Hpt: accumulation in the output dtype)
void orc_sum_f32_rows(const float* in, int64_t rows, int64_t cols, float* out) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < rows; ++r) {
    const float* src = in + r * cols;
    float acc = 0.f;
#pragma omp simd reduction(+ : acc)
    for (int64_t c = 0; c < cols; ++c) acc += src[c];
    out[r] = acc;
  }
}
// full sum: per-thread partial over a contiguous chunk, then combine
float orc_sum_f32_all(const float* in, int64_t n) {
  double total = 0.0;
#pragma omp parallel
  {
    float acc = 0.f;
#pragma omp for simd schedule(static) nowait
    for (int64_t i = 0; i < n; ++i) acc += in[i];
#pragma omp atomic
    total += (double)acc;
  }
  return (float)total;
}
// sum over axis 0 of [rows, cols] → [cols]: threads split columns, rows walked outermost (keeps SIMD on columns)
void orc_sum_f32_cols(const float* in, int64_t rows, int64_t cols, float* out) {
#pragma omp parallel
  {
    int nt = omp_get_num_threads(), t = omp_get_thread_num();
    int64_t c0 = cols * t / nt, c1 = cols * (t + 1) / nt;
    for (int64_t c = c0; c < c1; ++c) out[c] = 0.f;
    for (int64_t r = 0; r < rows; ++r) {
      const float* src = in + r * cols;
#pragma omp simd
      for (int64_t c = c0; c < c1; ++c) out[c] += src[c];
    }
  }
}
// max / argmax of a 2-D strided view over its axis 0: out[j] over i of in[i*s_red + j*s_out]  (config 2 passes
// s_red = 1, s_out = ld: each output scans one contiguous memory row)
void orc_max_f32_axis0(const float* in, int64_t n_red, int64_t n_out, int64_t s_red, int64_t s_out, float* out) {
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < n_out; ++j) {
    const float* src = in + j * s_out;
    float m = -std::numeric_limits<float>::infinity();
    if (s_red == 1) {
#pragma omp simd reduction(max : m)
      for (int64_t i = 0; i < n_red; ++i) m = src[i] > m ? src[i] : m;
    } else {
      for (int64_t i = 0; i < n_red; ++i) m = src[i * s_red] > m ? src[i * s_red] : m;
    }
    out[j] = m;
  }
}
void orc_argmax_f32_axis0(const float* in, int64_t n_red, int64_t n_out, int64_t s_red, int64_t s_out, int64_t* out) {
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < n_out; ++j) {
    const float* src = in + j * s_out;
    float m = -std::numeric_limits<float>::infinity();
    int64_t idx = 0;
    for (int64_t i = 0; i < n_red; ++i) {  // strict `>`: first index wins, NaN never wins
      float v = src[i * s_red];
      if (v > m) { m = v; idx = i; }
    }
    out[j] = idx;
  }
}
// mean over (N, H, W) of an NCHW bf16 tensor viewed NHWC → [C]; accumulation in bf16 (as Hpt's CPU path:
// output dtype bf16, every add rounds to bf16).  bf16 carried as uint16 bit patterns.
static inline float bf16_to_f32(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
static inline uint16_t f32_to_bf16(float f) {
  uint32_t u; memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
void orc_mean_bf16_nchw_channels(const uint16_t* in, int64_t N, int64_t C, int64_t HW, uint16_t* out, int hpt_rounding) {
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < C; ++c) {
    if (hpt_rounding) {
      uint16_t acc = 0;
      for (int64_t n = 0; n < N; ++n) {
        const uint16_t* src = in + (n * C + c) * HW;
        for (int64_t i = 0; i < HW; ++i) acc = f32_to_bf16(bf16_to_f32(acc) + bf16_to_f32(src[i]));
      }
      out[c] = f32_to_bf16(bf16_to_f32(acc) / bf16_to_f32(f32_to_bf16((float)(N * HW))));
    } else {
      float acc = 0.f;
      for (int64_t n = 0; n < N; ++n) {
        const uint16_t* src = in + (n * C + c) * HW;
#pragma omp simd reduction(+ : acc)
        for (int64_t i = 0; i < HW; ++i) acc += bf16_to_f32(src[i]);
      }
      out[c] = f32_to_bf16(acc / (float)(N * HW));
    }
  }
}

// ---- softmax / logsumexp over the last axis of [rows, cols] ---------------------------------------------------------
void orc_softmax_f32_rows(const float* in, int64_t rows, int64_t cols, float* out) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < rows; ++r) {
    const float* src = in + r * cols;
    float* dst = out + r * cols;
    float m = -std::numeric_limits<float>::infinity();
#pragma omp simd reduction(max : m)
    for (int64_t c = 0; c < cols; ++c) m = src[c] > m ? src[c] : m;
    for (int64_t c = 0; c < cols; ++c) dst[c] = src[c] - m;
    vec_unary(1, dst, dst, cols);
    float s = 0.f;
#pragma omp simd reduction(+ : s)
    for (int64_t c = 0; c < cols; ++c) s += dst[c];
#pragma omp simd
    for (int64_t c = 0; c < cols; ++c) dst[c] = dst[c] / s;
  }
}
void orc_logsumexp_f32_rows(const float* in, int64_t rows, int64_t cols, float* out) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < rows; ++r) {
    const float* src = in + r * cols;
    float s = 0.f;
    float buf[1024];
    for (int64_t c0 = 0; c0 < cols; c0 += 1024) {
      int64_t nb = cols - c0 < 1024 ? cols - c0 : 1024;
      vec_unary(1, src + c0, buf, nb);
#pragma omp simd reduction(+ : s)
      for (int64_t c = 0; c < nb; ++c) s += buf[c];
    }
    out[r] = logf(s);
  }
}

}  // extern "C"
