/* lean_math.h — log / exp / pow of the CRNN hot path, ONE definition for the CUDA kernels and for host code.
 *
 * The RHS spends its transcendental calls on NS (log) and NR (exp) lanes of a warp, so what they cost is ISSUE
 * SLOTS: CUDA's log/exp are ~85/~60 instructions each, a third of them moving 64-bit literals into registers.
 * These keep the classic argument reductions (log: m in [sqrt(1/2), sqrt 2), s = f/(2+f), odd series in s; exp:
 * k = rint(x/ln 2), Taylor degree 13 on |r| <= ln2/2), read their constants from one table (constant bank on the
 * device) and send anything outside the plain range (zero, negative, subnormal, inf, NaN; |x| >= ~700 for exp) to
 * the library call.
 *
 * BIT-REPRODUCIBLE across host and device: every operation is an explicitly rounded add / mul / fma (the macros
 * below stop nvcc from contracting; host translation units that include this file are built with
 * -ffp-contract=off), and the reciprocal inside log is a linear seed plus two cubically convergent corrections —
 * FMAs only, no MUFU — so the CPU oracle can run the SAME functions (oracle/lean_math_host.c) and
 * step-count comparisons do not hinge on glibc-vs-CUDA last-ulp differences.  Accuracy (tests/test_lean_math_cpu.py,
 * against mpmath): <= 1 ulp for both, the same class as the library versions.
 */
#ifndef CRNN_LEAN_MATH_H
#define CRNN_LEAN_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define CRNN_LM_FN __host__ __device__ __forceinline__
#else
#define CRNN_LM_FN static inline
#endif

/* table: 0..6 log series, 7/8 ln2 hi/lo, 9 1/ln2, 10 rint magic, 11..24 1/13! .. 1/0!, 25/26 reciprocal seed,
 * 27 ln 10, 28 1/ln 10 */
#define CRNN_LM_TABLE                                                                                              \
  6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01,          \
  1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01,                                    \
  6.93147180369123816490e-01, 1.90821492927058770002e-10,                                                          \
  1.4426950408889634074, 6755399441055744.0,                                                                       \
  1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0,         \
  1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0, 1.0,                                    \
  0.9850615000483448, -0.23901599922648414,                                                                        \
  2.302585092994045684, 0.4342944819032518277

#if defined(__CUDACC__)
__constant__ double c_lm[30] = {CRNN_LM_TABLE, 0.0};
#endif
static const double crnn_lm_host[30] = {CRNN_LM_TABLE, 0.0};

#if defined(__CUDA_ARCH__)
#define LM_C(i) c_lm[i]
#define LM_MUL(a, b) __dmul_rn((a), (b))
#define LM_ADD(a, b) __dadd_rn((a), (b))
#define LM_SUB(a, b) __dsub_rn((a), (b))
#define LM_FMA(a, b, c) __fma_rn((a), (b), (c))
#define LM_HI(x) __double2hiint(x)
#define LM_LO(x) __double2loint(x)
#define LM_MK(hi, lo) __hiloint2double((hi), (lo))
#else
#define LM_C(i) crnn_lm_host[i]
#define LM_MUL(a, b) ((a) * (b))
#define LM_ADD(a, b) ((a) + (b))
#define LM_SUB(a, b) ((a) - (b))
#define LM_FMA(a, b, c) fma((a), (b), (c))
static inline int crnn_lm_hi(double x) { uint64_t u; memcpy(&u, &x, 8); return (int)(uint32_t)(u >> 32); }
static inline int crnn_lm_lo(double x) { uint64_t u; memcpy(&u, &x, 8); return (int)(uint32_t)u; }
static inline double crnn_lm_mk(int hi, int lo) {
  uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint64_t)(uint32_t)lo; double x; memcpy(&x, &u, 8); return x;
}
#define LM_HI(x) crnn_lm_hi(x)
#define LM_LO(x) crnn_lm_lo(x)
#define LM_MK(hi, lo) crnn_lm_mk((hi), (lo))
#endif

/* out-of-range arguments: the library call, out of line on the device (rare, keeps the hot code small) */
#if defined(__CUDACC__)
static __device__ __noinline__ double lib_log(double x) { return log(x); }
static __device__ __noinline__ double lib_exp(double x) { return exp(x); }
#endif
#if defined(__CUDA_ARCH__)
#define LM_LIBLOG(x) lib_log(x)
#define LM_LIBEXP(x) lib_exp(x)
#else
#define LM_LIBLOG(x) log(x)
#define LM_LIBEXP(x) exp(x)
#endif

CRNN_LM_FN double lean_log(double x) {
  const int hi = LM_HI(x), lo = LM_LO(x);
  if ((unsigned)(hi - 0x00100000) >= 0x7fe00000u) return LM_LIBLOG(x);
  const int hx = hi & 0xfffff, i = (hx + 0x95f64) & 0x100000;
  const double m = LM_MK(hx | (i ^ 0x3ff00000), lo);
  const double dk = (double)((hi >> 20) - 1023 + (i >> 20));
  const double f = LM_SUB(m, 1.0), d = LM_ADD(2.0, f);   /* d in [1.707, 2.414] */
  /* r = 1/d: minimax linear seed (1.6e-2), two corrections r <- r(1 + e + e^2), e = 1 - d r  (e -> e^3) */
  double r = LM_FMA(d, LM_C(26), LM_C(25));
  double e = LM_FMA(-d, r, 1.0);
  r = LM_FMA(r, LM_FMA(e, e, e), r);
  e = LM_FMA(-d, r, 1.0);
  r = LM_FMA(r, LM_FMA(e, e, e), r);
  const double s = LM_MUL(f, r), z = LM_MUL(s, s), w = LM_MUL(z, z);
  const double t1 = LM_MUL(w, LM_FMA(w, LM_FMA(w, LM_C(5), LM_C(3)), LM_C(1)));
  const double R = LM_FMA(z, LM_FMA(w, LM_FMA(w, LM_FMA(w, LM_C(6), LM_C(4)), LM_C(2)), LM_C(0)), t1);
  const double hfsq = LM_MUL(LM_MUL(0.5, f), f);
  const double inner = LM_FMA(s, LM_ADD(hfsq, R), LM_MUL(dk, LM_C(8)));
  return LM_FMA(dk, LM_C(7), -LM_SUB(LM_SUB(hfsq, inner), f));
}

CRNN_LM_FN double lean_exp(double x) {
  if ((unsigned)(LM_HI(x) & 0x7fffffff) >= 0x4085e000u) return LM_LIBEXP(x);
  const double t = LM_FMA(x, LM_C(9), LM_C(10));
  const double kf = LM_SUB(t, LM_C(10));
  double r = LM_FMA(kf, -LM_C(7), x);
  r = LM_FMA(kf, -LM_C(8), r);
  double p = LM_C(11);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int n = 12; n <= 24; ++n) p = LM_FMA(p, r, LM_C(n));
  return LM_MK(LM_HI(p) + (int)((unsigned)LM_LO(t) << 20), LM_LO(p));
}

/* "compute, then fix" forms: the fast path runs unconditionally (straight-line code a scheduler can interleave with
 * independent work) and the rare out-of-range argument is patched afterwards.  Same values as lean_log / lean_exp. */
CRNN_LM_FN double lean_log_cf(double x) {
  const int hi = LM_HI(x), lo = LM_LO(x);
  const int hx = hi & 0xfffff, i = (hx + 0x95f64) & 0x100000;
  const double m = LM_MK(hx | (i ^ 0x3ff00000), lo);
  const double dk = (double)((hi >> 20) - 1023 + (i >> 20));
  const double f = LM_SUB(m, 1.0), d = LM_ADD(2.0, f);
  double r = LM_FMA(d, LM_C(26), LM_C(25));
  double e = LM_FMA(-d, r, 1.0);
  r = LM_FMA(r, LM_FMA(e, e, e), r);
  e = LM_FMA(-d, r, 1.0);
  r = LM_FMA(r, LM_FMA(e, e, e), r);
  const double s = LM_MUL(f, r), z = LM_MUL(s, s), w = LM_MUL(z, z);
  const double t1 = LM_MUL(w, LM_FMA(w, LM_FMA(w, LM_C(5), LM_C(3)), LM_C(1)));
  const double R = LM_FMA(z, LM_FMA(w, LM_FMA(w, LM_FMA(w, LM_C(6), LM_C(4)), LM_C(2)), LM_C(0)), t1);
  const double hfsq = LM_MUL(LM_MUL(0.5, f), f);
  const double inner = LM_FMA(s, LM_ADD(hfsq, R), LM_MUL(dk, LM_C(8)));
  double y = LM_FMA(dk, LM_C(7), -LM_SUB(LM_SUB(hfsq, inner), f));
  if ((unsigned)(hi - 0x00100000) >= 0x7fe00000u) y = LM_LIBLOG(x);
  return y;
}

CRNN_LM_FN double lean_exp_cf(double x) {
  const double t = LM_FMA(x, LM_C(9), LM_C(10));
  const double kf = LM_SUB(t, LM_C(10));
  double r = LM_FMA(kf, -LM_C(7), x);
  r = LM_FMA(kf, -LM_C(8), r);
  double p = LM_C(11);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int n = 12; n <= 24; ++n) p = LM_FMA(p, r, LM_C(n));
  double y = LM_MK(LM_HI(p) + (int)((unsigned)LM_LO(t) << 20), LM_LO(p));
  if ((unsigned)(LM_HI(x) & 0x7fffffff) >= 0x4085e000u) y = LM_LIBEXP(x);
  return y;
}

/* x^y for x > 0 (step-size controller exponents; OrdinaryDiffEq uses its own `fastpow` there): a few ulp */
CRNN_LM_FN double lean_pow(double x, double y) { return lean_exp(LM_MUL(y, lean_log(x))); }
/* the initial-step heuristic's 10^(-(2 + log10 d)/order) */
CRNN_LM_FN double lean_log10(double x) { return LM_MUL(lean_log(x), LM_C(28)); }
CRNN_LM_FN double lean_exp10(double x) { return lean_exp(LM_MUL(x, LM_C(27))); }

#endif /* CRNN_LEAN_MATH_H */
