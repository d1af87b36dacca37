// kernel_train.cuh — the device side of the on-device training loop.
//
// The reference's epoch loop (case2/case2.jl:192-207) is  for i_exp in randperm(n_exp_train):  grad = ForwardDiff.gradient(...);
// update!(opt, p, grad)  — batch size 1, ~74 000 optimiser steps for the committed case2 checkpoint.  At that grain a step is
// launch- and round-trip-bound, not compute-bound, so the whole step lives on the device: a p2vec kernel turns the parameter
// vector into the physical weights and the structured seed columns, the forward-sensitivity kernel reads them from device
// memory (k_tsit5_sens<..., DEVW>), the gradient is reduced, and the optimiser kernel below applies
// Flux's ExpDecay -> ADAM / NADAM -> WeightDecay chain (case2.jl:31-32, case3.jl:20, rober_crnn.jl:19; formulas as in
// crnn_b200/optim.py) with the 2-norm clip of rober_crnn.jl:220-223.  Nothing returns to the host between steps.
#pragma once
#include "crnn_dev.cuh"
#include "kernel_tsit5_sens.cuh"
#include "kernel_rosenbrock23_sens.cuh"

namespace crnn {

struct TrainP {
  int optimiser;            // 0 ADAM, 1 NADAM
  int np;
  double eta, beta1, beta2, eps, weight_decay;
  double expdecay_decay, expdecay_clip, grad_max;
  long long expdecay_step;  // <= 0: no ExpDecay link
};

// optimiser state in device memory: m[np] | v[np] | beta1^t | beta2^t | ExpDecay eta | ExpDecay count
__device__ __forceinline__ double* opt_m(double* st) { return st; }
__device__ __forceinline__ double* opt_v(double* st, int np) { return st + np; }
__device__ __forceinline__ double* opt_tail(double* st, int np) { return st + 2 * np; }

// p2vec of case2/case2.jl:91-99 for the forward-sensitivity kernel: physical weights (ModelP) and the structured seed columns
// (one column per parameter: rows a_j = dW_in[i_in, j], b_j = db_j of SensSmem::seed, and the R1Desc entries).
//   slope = p[25]*100; w_b = p[1:3]*slope; w_out = reshape(p[4:21], 6, 3); Ea = |p[22:24]*slope|; w_in = [clamp(-w_out, 0, 4); Ea']
template <class C>
__global__ void k_p2vec_case2(const double* __restrict__ p, double lb, double ub, double gas_R, ModelP<C>* __restrict__ mp,
                              double* __restrict__ rows /* [2*NR][32] */, R1Desc* __restrict__ desc /* [32] */) {
  constexpr int NS = C::NS, NR = C::NR, NIN = C::NIN;
  static_assert(NS == 6 && NR == 3 && C::KIND == 1, "case2 dimensions");
  const int t = threadIdx.x;
  for (int q = t; q < 2 * NR * 32; q += blockDim.x) rows[q] = 0.0;
  if (t < 32) { R1Desc d{}; d.o = 0.0; d.i_in = 0; d.i_out = 0; d.j_out = 0; d.pad = 0; desc[t] = d; }
  __syncthreads();
  const double p_slope = p[NR * (NS + 2)], slope = p_slope * 100.0;
  if (t == 0) { mp->lb = lb; mp->ub = ub; mp->gas_R = gas_R; }
  if (t < NR) {
    const int j = t;
    mp->w_b[j] = p[j] * slope;
    rows[(NR + j) * 32 + (1 + j)] = slope;                                   // d w_b[j] / d p[j]
    rows[(NR + j) * 32 + (1 + NR * (NS + 2))] = p[j] * 100.0;                // d w_b[j] / d p[slope]
    const double ea = p[NR * (NS + 1) + j] * slope;
    const double sg = signbit(ea) ? -1.0 : 1.0;                              // abs(dual): sign from signbit
    mp->w_in[NS + NIN * j] = ea * sg;
    rows[j * 32 + (1 + NR * (NS + 1) + j)] = sg * slope;                     // d Ea[j] / d p[Ea_j]
    rows[j * 32 + (1 + NR * (NS + 2))] = sg * (p[NR * (NS + 1) + j] * 100.0);  // d Ea[j] / d p[slope]
    desc[1 + NR * (NS + 1) + j].i_in = NS;
    if (j == 0) desc[1 + NR * (NS + 2)].i_in = NS;
  }
  if (t < NS * NR) {
    const int i = t % NS, j = t / NS;
    const double wo = p[NR + i + NS * j];
    mp->w_out[i + NS * j] = wo;                                               // no out_scale in case2
    const double wi = -wo;
    mp->w_in[i + NIN * j] = wi > 4.0 ? 4.0 : (wi < 0.0 ? 0.0 : wi);
    const int c = 1 + NR + i + NS * j;
    R1Desc d{}; d.o = 1.0; d.i_in = i; d.i_out = i; d.j_out = j; d.pad = 0;
    desc[c] = d;
    rows[j * 32 + c] = (wi >= 0.0 && wi <= 4.0) ? -1.0 : 0.0;               // clamp(dual): derivative 1 on the closed interval
  }
}

// p2vec of case1/case1.jl:70-78: w_b = p[1:nr] .+ b0 (b0 = -10, :70), w_out = reshape(p[nr+1:end], ns, nr), w_in = clamp(-w_out, 0, 2.5)
template <class C>
__global__ void k_p2vec_case1(const double* __restrict__ p, double lb, double ub, double b0, ModelP<C>* __restrict__ mp,
                              double* __restrict__ rows /* [2*NR][32] */, R1Desc* __restrict__ desc /* [32] */) {
  constexpr int NS = C::NS, NR = C::NR, NIN = C::NIN;
  static_assert(C::KIND == 0 && NIN == NS, "an F0 model");
  const int t = threadIdx.x;
  for (int q = t; q < 2 * NR * 32; q += blockDim.x) rows[q] = 0.0;
  if (t < 32) { R1Desc d{}; d.o = 0.0; d.i_in = 0; d.i_out = 0; d.j_out = 0; d.pad = 0; desc[t] = d; }
  __syncthreads();
  if (t == 0) { mp->lb = lb; mp->ub = ub; mp->gas_R = 0.0; }
  if (t < NR) {
    mp->w_b[t] = p[t] + b0;
    rows[(NR + t) * 32 + (1 + t)] = 1.0;
  }
  if (t < NS * NR) {
    const int i = t % NS, j = t / NS;
    const double wo = p[NR + i + NS * j];
    mp->w_out[i + NS * j] = wo;
    const double wi = -wo;
    mp->w_in[i + NIN * j] = wi > 2.5 ? 2.5 : (wi < 0.0 ? 0.0 : wi);
    const int c = 1 + NR + i + NS * j;
    R1Desc d{}; d.o = 1.0; d.i_in = i; d.i_out = i; d.j_out = j; d.pad = 0;
    desc[c] = d;
    rows[j * 32 + c] = (wi >= 0.0 && wi <= 2.5) ? -1.0 : 0.0;
  }
}

// p2vec of case3/case3.jl:42-53 (9 species, 8 reactions, np = 153; five warps share a trajectory: COLS = 160 seed columns):
//   w_b = p[1:nr];  w_in_raw = reshape(p[nr(ns+1)+1 : nr(2ns+1)], ns, nr);  w_out = -w_in_raw .* abs(reshape(p[nr+1 : nr(ns+1)], ns, nr));
//   w_in = clamp(w_in_raw, 0, 4);  p[end] is not used by the weights.  out_scale (dy_std of the RHS, case3.jl:165) is folded into
//   w_out and into the seed entries as pack<C>() / plan_r1 do on the host path.
template <class C, int COLS>
__global__ void k_p2vec_case3(const double* __restrict__ p, double lb, double ub, const double* __restrict__ oscale,
                              ModelP<C>* __restrict__ mp, double* __restrict__ rows /* [2*NR][COLS] */, R1Desc* __restrict__ desc /* [COLS] */) {
  constexpr int NS = C::NS, NR = C::NR, NIN = C::NIN;
  static_assert(C::KIND == 0 && NIN == NS && 2 + NR * (2 * NS + 1) <= COLS, "an F0 model whose columns fit");
  const int t = threadIdx.x;
  for (int q = t; q < 2 * NR * COLS; q += blockDim.x) rows[q] = 0.0;
  for (int q = t; q < COLS; q += blockDim.x) { R1Desc d{}; d.o = 0.0; d.i_in = 0; d.i_out = 0; d.j_out = 0; d.pad = 0; desc[q] = d; }
  __syncthreads();
  if (t == 0) { mp->lb = lb; mp->ub = ub; mp->gas_R = 0.0; }
  if (t < NR) {
    mp->w_b[t] = p[t];
    rows[(NR + t) * COLS + (1 + t)] = 1.0;
  }
  if (t < NS * NR) {
    const int i = t % NS, j = t / NS;
    const double os = oscale ? oscale[i] : 1.0;
    const double a = p[NR * (NS + 1) + i + NS * j], b = p[NR + i + NS * j];   // w_in_raw, w_out_raw
    const double sg = signbit(b) ? -1.0 : 1.0, ab = b * sg;                    // abs(dual): sign from signbit
    mp->w_out[i + NS * j] = ((-a) * ab) * os;
    mp->w_in[i + NIN * j] = a > 4.0 ? 4.0 : (a < 0.0 ? 0.0 : a);
    const int c_out = 1 + NR + i + NS * j, c_in = 1 + NR * (NS + 1) + i + NS * j;
    R1Desc d{}; d.pad = 0; d.i_in = i; d.i_out = i; d.j_out = j;
    d.o = ((-a) * sg) * os;  desc[c_out] = d;                                  // d w_out[i,j] / d w_out_raw[i,j]; no w_in row
    d.o = (-ab) * os;        desc[c_in] = d;                                   // d w_out[i,j] / d w_in_raw[i,j]
    rows[j * COLS + c_in] = (a >= 0.0 && a <= 4.0) ? 1.0 : 0.0;                // clamp(dual): derivative 1 on the closed interval
  }
}

// p2vec of robertson/rober_crnn.jl:85-96 (3 species, 6 reactions, np = 43: two column tiles, COLS = 64):
//   slope = abs(p[end]);  w_b = p[1:nr] .* (10 slope);  w_in_raw = reshape(p[nr(ns+1)+1 : nr(2ns+1)], ns, nr);
//   w_out = -w_in_raw .* 10 .^ reshape(p[nr+1 : nr(ns+1)], ns, nr);  w_in = clamp(w_in_raw, 0, 2.5);  out_scale = dydt_scale (:81-82,115)
template <class C, int COLS>
__global__ void k_p2vec_robertson(const double* __restrict__ p, double lb, double ub, const double* __restrict__ oscale,
                                  ModelP<C>* __restrict__ mp, double* __restrict__ rows /* [2*NR][COLS] */, R1Desc* __restrict__ desc /* [COLS] */) {
  constexpr int NS = C::NS, NR = C::NR, NIN = C::NIN;
  static_assert(C::KIND == 0 && NIN == NS && 2 + NR * (2 * NS + 1) <= COLS, "an F0 model whose columns fit");
  constexpr int K_SLOPE = NR * (2 * NS + 1);
  const int t = threadIdx.x;
  for (int q = t; q < 2 * NR * COLS; q += blockDim.x) rows[q] = 0.0;
  for (int q = t; q < COLS; q += blockDim.x) { R1Desc d{}; d.o = 0.0; d.i_in = 0; d.i_out = 0; d.j_out = 0; d.pad = 0; desc[q] = d; }
  __syncthreads();
  if (t == 0) { mp->lb = lb; mp->ub = ub; mp->gas_R = 0.0; }
  const double sgs = signbit(p[K_SLOPE]) ? -1.0 : 1.0, slope = p[K_SLOPE] * sgs;   // abs(dual): sign from signbit
  if (t < NR) {
    mp->w_b[t] = p[t] * (slope * 10.0);
    rows[(NR + t) * COLS + (1 + t)] = slope * 10.0;
    rows[(NR + t) * COLS + (1 + K_SLOPE)] = p[t] * (sgs * 10.0);
  }
  if (t < NS * NR) {
    const int i = t % NS, j = t / NS;
    const double os = oscale ? oscale[i] : 1.0;
    const double a = p[NR * (NS + 1) + i + NS * j], b = p[NR + i + NS * j];   // w_in_raw, w_out_raw
    const double v = pow(10.0, b);
    mp->w_out[i + NS * j] = ((-a) * v) * os;
    mp->w_in[i + NIN * j] = a > 2.5 ? 2.5 : (a < 0.0 ? 0.0 : a);
    const int c_out = 1 + NR + i + NS * j, c_in = 1 + NR * (NS + 1) + i + NS * j;
    R1Desc d{}; d.pad = 0; d.i_in = i; d.i_out = i; d.j_out = j;
    d.o = ((-a) * (v * 2.302585092994046)) * os;  desc[c_out] = d;           // d w_out / d w_out_raw = -a 10^b ln 10; no w_in row
    d.o = (-v) * os;                              desc[c_in] = d;            // d w_out / d w_in_raw
    rows[j * COLS + c_in] = (a >= 0.0 && a <= 2.5) ? 1.0 : 0.0;               // clamp(dual): derivative 1 on the closed interval
  }
}

// [sum of the finite losses, number of them] of one step's experiments (fixed order: deterministic)
static __global__ void __launch_bounds__(256) k_train_loss_sum(const double* __restrict__ loss, int n, double* __restrict__ out) {
  __shared__ double s_sum[256], s_cnt[256];
  double s = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) { const double v = loss[i]; if (v == v && fabs(v) != INFINITY) { s += v; c += 1.0; } }
  s_sum[threadIdx.x] = s; s_cnt[threadIdx.x] = c;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) { s_sum[threadIdx.x] += s_sum[threadIdx.x + w]; s_cnt[threadIdx.x] += s_cnt[threadIdx.x + w]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = s_sum[0]; out[1] = s_cnt[0]; }
}

// one optimiser step on the device: p -= chain(grad), grad = grad_sum / n_ok (mean over the step's experiments)
static __global__ void __launch_bounds__(256)
k_optim_step(const TrainP T, const double* __restrict__ grad_sum, const double* __restrict__ loss_sum /* [sum, n_ok] */,
             double* __restrict__ p, double* __restrict__ st, double* __restrict__ step_loss, double* __restrict__ step_gnorm,
             long long step) {
  __shared__ double red[256];
  const int t = threadIdx.x, np = T.np;
  const double n_ok = loss_sum[1] > 0.0 ? loss_sum[1] : 1.0;
  double g = t < np ? grad_sum[t] / n_ok : 0.0;
  red[t] = g * g;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) { if (t < w) red[t] += red[t + w]; __syncthreads(); }
  const double gn = sqrt(red[0]);
  if (T.grad_max > 0.0 && gn > T.grad_max) g = g / gn * T.grad_max;          // rober_crnn.jl:220-223
  double* tail = opt_tail(st, np);
  const double b1p = tail[0], b2p = tail[1];
  double eta_d = tail[2], cnt = tail[3];
  if (T.expdecay_step > 0) {                                                  // Flux.ExpDecay, placed before ADAM (case2.jl:31-32)
    cnt += 1.0;
    if (((long long)cnt) % T.expdecay_step == 0 && eta_d > T.expdecay_clip) eta_d = fmax(eta_d * T.expdecay_decay, T.expdecay_clip);
    g = g * eta_d;
  }
  if (t < np) {
    // every product and sum rounded once (no FMA contraction): the arithmetic of Flux's broadcasts, so a step equals the
    // host mirror's (crnn_b200/optim.py) bit for bit given the same gradient
    auto mul = [](double a, double b) { return __dmul_rn(a, b); };
    auto add = [](double a, double b) { return __dadd_rn(a, b); };
    double* m = opt_m(st); double* v = opt_v(st, np);
    const double mt = add(mul(T.beta1, m[t]), mul(1.0 - T.beta1, g));
    const double vt = add(mul(T.beta2, v[t]), mul(mul(1.0 - T.beta2, g), g));
    m[t] = mt; v[t] = vt;
    double d;
    if (T.optimiser == 0) d = mul(mt / (1.0 - b1p) / add(sqrt(vt / (1.0 - b2p)), T.eps), T.eta);
    else d = mul(add(mul(T.beta1, mt) / (1.0 - mul(T.beta1, b1p)), mul(1.0 - T.beta1, g) / (1.0 - b1p)) /
                 add(sqrt(mul(vt, T.beta2) / (1.0 - b2p)), T.eps), T.eta);
    d = add(d, mul(T.weight_decay, p[t]));                                    // Flux.WeightDecay after ADAM (ADAMW)
    p[t] -= d;
  }
  __syncthreads();
  if (t == 0) {
    tail[0] = b1p * T.beta1; tail[1] = b2p * T.beta2; tail[2] = eta_d; tail[3] = cnt;
    if (step_loss) step_loss[step] = loss_sum[0] / n_ok;
    if (step_gnorm) step_gnorm[step] = gn;
  }
}

}  // namespace crnn
