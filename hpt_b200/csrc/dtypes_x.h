// dtypes_x.h — X-macro over the 13 scalar dtypes of the hot path: X(c++ type, short name, enum value)
#pragma once
#define HPTB_FOR_DTYPES(X) \
  X(b8, bool, 0) X(int8_t, i8, 1) X(int16_t, i16, 2) X(int32_t, i32, 3) X(int64_t, i64, 4) \
  X(uint8_t, u8, 5) X(uint16_t, u16, 6) X(uint32_t, u32, 7) X(uint64_t, u64, 8)            \
  X(f16, f16, 9) X(bf16, bf16, 10) X(float, f32, 11) X(double, f64, 12)

#define HPTB_CAT2(a, b) a##b
#define HPTB_CAT(a, b) HPTB_CAT2(a, b)
#define HPTB_CAT4(a, b, c, d) HPTB_CAT(HPTB_CAT(a, b), HPTB_CAT(c, d))

// binary ops: X(functor, name, enum, promote kind, bool output allowed)
#define HPTB_FOR_BINARY_OPS(X)                                                       \
  X(OpAdd, add, HPTB_ADD, 0, 1) X(OpSub, sub, HPTB_SUB, 0, 0) X(OpMul, mul, HPTB_MUL, 0, 1) \
  X(OpRem, rem, HPTB_REM, 0, 0) X(OpDiv, div, HPTB_DIV, 1, 0) X(OpMax, maximum, HPTB_MAXIMUM, 0, 1) \
  X(OpMin, minimum, HPTB_MINIMUM, 0, 1) X(OpPow, pow, HPTB_POW, 1, 0) X(OpHypot, hypot, HPTB_HYPOT, 1, 0)

// unary ops: X(name, enum)
#define HPTB_FOR_UNARY_OPS(X)                                                                  \
  X(sin, HPTB_SIN) X(cos, HPTB_COS) X(tan, HPTB_TAN) X(asin, HPTB_ASIN) X(acos, HPTB_ACOS)     \
  X(atan, HPTB_ATAN) X(sinh, HPTB_SINH) X(cosh, HPTB_COSH) X(tanh, HPTB_TANH)                  \
  X(asinh, HPTB_ASINH) X(acosh, HPTB_ACOSH) X(atanh, HPTB_ATANH) X(exp, HPTB_EXP)              \
  X(exp2, HPTB_EXP2) X(exp10, HPTB_EXP10) X(ln, HPTB_LN) X(log2, HPTB_LOG2) X(log10, HPTB_LOG10) \
  X(sqrt, HPTB_SQRT) X(cbrt, HPTB_CBRT) X(recip, HPTB_RECIP) X(erf, HPTB_ERF)                  \
  X(sigmoid, HPTB_SIGMOID) X(gelu, HPTB_GELU) X(selu, HPTB_SELU) X(elu, HPTB_ELU)              \
  X(celu, HPTB_CELU) X(mish, HPTB_MISH) X(softplus, HPTB_SOFTPLUS) X(softsign, HPTB_SOFTSIGN)  \
  X(hard_sigmoid, HPTB_HARD_SIGMOID) X(hard_swish, HPTB_HARD_SWISH)

// NormalUaryOps (+ BITNOT): X(name, enum) — out dtype = in dtype
#define HPTB_FOR_NORMAL_UNARY_OPS(X)                                                                   \
  X(floor, HPTB_FLOOR) X(ceil, HPTB_CEIL) X(round, HPTB_ROUND) X(trunc, HPTB_TRUNC) X(abs, HPTB_ABS)    \
  X(neg, HPTB_NEG) X(sign, HPTB_SIGN) X(square, HPTB_SQUARE) X(relu, HPTB_RELU) X(relu6, HPTB_RELU6)    \
  X(leaky_relu, HPTB_LEAKY_RELU) X(clamp, HPTB_CLAMP) X(bitnot, HPTB_BITNOT)
