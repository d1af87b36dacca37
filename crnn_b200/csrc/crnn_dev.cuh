// crnn_dev.cuh — device-side building blocks of the CRNN engine (sm_100a).
//
// Everything is fp64 on the CUDA cores: the largest contraction on this path is
// 9x8, so tensor cores do not apply.  Dimensions are template
// parameters so that the state, the 7 Tsit5 stage vectors and the weights live
// in registers / the constant bank with fully unrolled loops.
#pragma once
#include <cstdint>
#include <cmath>
#include "../../include/crnn_b200.h"
#include "lean_math.h"

namespace crnn {

template <int NS_, int NR_, int KIND_>
struct Cfg {
  static constexpr int NS = NS_;                         // species (rows of w_out)
  static constexpr int NR = NR_;                         // reactions
  static constexpr int KIND = KIND_;                     // crnn_rhs_kind
  static constexpr int N = NS_ + (KIND_ == 1 ? 1 : 0);   // state length
  static constexpr int NIN = N;                          // rows of w_in
  static constexpr int NW = NR_ * (NIN + 1 + NS_);       // flat weight count
};

// Weights travel as a by-value kernel parameter: they sit in the constant bank
// and feed DFMA directly as c[][] operands (no load instructions).
// w_out (and the w_out rows of the seed) are pre-multiplied by out_scale on the host.
template <class C>
struct ModelP {
  double w_in[C::NIN * C::NR];
  double w_b[C::NR];
  double w_out[C::NS * C::NR];
  double lb, ub, gas_R;
};

template <class C>
struct SolveP {
  double abstol[C::N], reltol[C::N];
  double inv_yscale[C::N];  // per state row (1 where unused)
  double t0, t1, pred_lo, pred_hi;
  double inv_qmin, inv_qmax, gamma, beta1, beta2, inv_order;
  double beta1_ros, beta2_ros;  // AutoTsit5: PI exponents while the Rosenbrock23 half runs
  double qs_min, qs_max;        // step_accept_controller!'s dead-band: qs_min <= q <= qs_max keeps dt
  double norm_cnt;              // divisor of the dual-aware norms: totallength(u) = N*(1+np), or N (crnn_opts)
  double eig_cnt;               // AutoSwitch eigenvalue estimate: N * (number of columns in the norm)
  long long maxiters;
  const double* saveat;  // device [n_save]
  const int* row2obs;    // device [N]: observation slot of state row i, or -1
  int n_save, n_obs, incl_sens, loss_kind;
};

// ---- Tsit5 (Tsitouras 2011) — SURVEY App. C.1/C.2 ----
namespace ts {
constexpr double a21 = 0.161;
constexpr double a31 = -0.008480655492356989, a32 = 0.335480655492357;
constexpr double a41 = 2.8971530571054935, a42 = -6.359448489975075, a43 = 4.3622954328695815;
constexpr double a51 = 5.325864828439257, a52 = -11.748883564062828, a53 = 7.4955393428898365,
                 a54 = -0.09249506636175525;
constexpr double a61 = 5.86145544294642, a62 = -12.92096931784711, a63 = 8.159367898576159,
                 a64 = -0.071584973281401, a65 = -0.028269050394068383;
constexpr double a71 = 0.09646076681806523, a72 = 0.01, a73 = 0.4798896504144996, a74 = 1.379008574103742,
                 a75 = -3.290069515436081, a76 = 2.324710524099774;
constexpr double bt1 = -0.00178001105222577714, bt2 = -0.0008164344596567469, bt3 = 0.007880878010261995,
                 bt4 = -0.1447110071732629, bt5 = 0.5823571654525552, bt6 = -0.45808210592918697,
                 bt7 = 0.015151515151515152;
// dense output b_i(theta) = theta*(r_i1 + theta*(r_i2 + theta*(r_i3 + theta*r_i4)))
constexpr double r11 = 1.0, r12 = -2.763706197274826, r13 = 2.9132554618219126, r14 = -1.0530884977290216;
constexpr double r22 = 0.13169999999999998, r23 = -0.2234, r24 = 0.1017;
constexpr double r32 = 3.9302962368947516, r33 = -5.941033872131505, r34 = 2.490627285651253;
constexpr double r42 = -12.411077166933676, r43 = 30.33818863028232, r44 = -16.548102889244902;
constexpr double r52 = 37.50931341651104, r53 = -88.1789048947664, r54 = 47.37952196281928;
constexpr double r62 = -27.896526289197286, r63 = 65.09189467479366, r64 = -34.87065786149661;
constexpr double r72 = 1.5, r73 = -4.0, r74 = 2.5;
// coefficients read from constant memory: as literals each one costs two register moves per use
__constant__ double c_dense[22] = {r11, r12, r13, r14, r22, r23, r24, r32, r33, r34, r42, r43, r44,
                                   r52, r53, r54, r62, r63, r64, r72, r73, r74};
__device__ __forceinline__ void dense_b(double th, double (&b)[7]) {
  b[0] = th * (c_dense[0] + th * (c_dense[1] + th * (c_dense[2] + th * c_dense[3])));
  const double t2 = th * th;
#pragma unroll
  for (int j = 1; j < 7; ++j) b[j] = t2 * (c_dense[1 + 3 * j] + th * (c_dense[2 + 3 * j] + th * c_dense[3 + 3 * j]));
}
}  // namespace ts

// Julia's min/max propagate NaN (fmin/fmax drop it): a NaN RHS must surface as a NaN dt.
__device__ __forceinline__ double jmin(double a, double b) { return a < b ? a : (b <= a ? b : a + b); }
__device__ __forceinline__ double jmax(double a, double b) { return a > b ? a : (b >= a ? b : a + b); }

// Julia Base.clamp semantics (NaN propagates).
__device__ __forceinline__ double clampd(double v, double lo, double hi) {
  return v > hi ? hi : (v < lo ? lo : v);
}

// lean_log / lean_exp / lean_pow / lean_log10 / lean_exp10: lean_math.h — one definition shared with host code
// (the CPU oracle's shared-math build runs the very same functions, bit for bit).
// out-of-line copy for cold call sites of the instruction-fetch-bound kernels
static __device__ __noinline__ double lean_log_nl(double x) { return lean_log(x); }

// OrdinaryDiffEq PI controller.
template <class C>
__device__ __forceinline__ double pi_controller(const SolveP<C>& sp, double EEst, double qold, double& q11) {
  if (EEst == 0.0) { q11 = 0.0; return sp.inv_qmax; }
  q11 = lean_pow(EEst, sp.beta1);
  double q = q11 / lean_pow(qold, sp.beta2);
  return jmax(sp.inv_qmax, jmin(sp.inv_qmin, q / sp.gamma));
}

__device__ __forceinline__ double snap_t(double tnew, double tend) {
  if (fabs(tnew - tend) < 100.0 * 2.220446049250313e-16 * fmax(fabs(tnew), fabs(tend))) return tend;
  return tnew;
}

__device__ __forceinline__ double ulp_of(double a) {
  a = fabs(a);
  return __longlong_as_double(__double_as_longlong(a) + 1) - a;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

// ------------------------------------------------------------------------------------------
// RHS value, thread-local (thread-per-trajectory kernels).
//   F0: case1.jl:80-83, case3.jl:162-166, rober_crnn.jl:113-116 ; F1: case2.jl:113-118
// bT[j] = w_b[j] (+ w_in[NS,j] * (-1/(R T)) for F1) is constant along a trajectory (dT/dt = 0).
// ------------------------------------------------------------------------------------------
template <class C>
__device__ __forceinline__ void rhs_value(const ModelP<C>& mp, const double (&bT)[C::NR],
                                          const double (&u)[C::NS], double (&du)[C::NS]) {
  double x[C::NS], r[C::NR];
#pragma unroll
  for (int i = 0; i < C::NS; ++i) x[i] = lean_log(clampd(u[i], mp.lb, mp.ub));
#pragma unroll
  for (int j = 0; j < C::NR; ++j) {
    double z = bT[j];
#pragma unroll
    for (int i = 0; i < C::NS; ++i) z = fma(mp.w_in[i + C::NIN * j], x[i], z);
    r[j] = lean_exp(z);
  }
#pragma unroll
  for (int i = 0; i < C::NS; ++i) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < C::NR; ++j) s = fma(mp.w_out[i + C::NS * j], r[j], s);
    du[i] = s;
  }
}

template <class C>
__device__ __forceinline__ void make_bT(const ModelP<C>& mp, double Tval, double (&bT)[C::NR]) {
#pragma unroll
  for (int j = 0; j < C::NR; ++j) {
    bT[j] = mp.w_b[j];
    if (C::KIND == 1) bT[j] = fma(mp.w_in[C::NS + C::NIN * j], -1.0 / (mp.gas_R * Tval), bT[j]);
  }
}

}  // namespace crnn
