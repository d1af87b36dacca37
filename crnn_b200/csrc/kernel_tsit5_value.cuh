// kernel_tsit5_value.cuh — predict_neuralode for a batch: one THREAD owns one trajectory.
//
// Replaces `solve(prob, Tsit5(), u0=u0, p=p, saveat=tsteps)` of predict_neuralode
// (case1/case1.jl:92-97, case2/case2.jl:124-128, case3/case3.jl:172-176).  The state, the
// seven stage vectors and the CRNN RHS live in registers; the whole adaptive integration
// (init-dt, PI controller, dense-output saveat) runs in-kernel with no launch per step.
#pragma once
#include "crnn_dev.cuh"

namespace crnn {

template <class C>
__global__ void __launch_bounds__(128)
k_tsit5_value(const __grid_constant__ ModelP<C> mp, const __grid_constant__ SolveP<C> sp,
              const double* __restrict__ u0, const int* __restrict__ n_save_used, long long ntraj,
              double* __restrict__ pred, int* __restrict__ n_saved, int* __restrict__ retcode,
              crnn_stats* __restrict__ stats) {
  constexpr int NS = C::NS, NR = C::NR, N = C::N;
  const long long traj = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (traj >= ntraj) return;

  double u[NS], un[NS], k1[NS], k2[NS], k3[NS], k4[NS], k5[NS], k6[NS], k7[NS], tmp[NS];
  double Tval = 0.0;
#pragma unroll
  for (int i = 0; i < NS; ++i) u[i] = u0[traj * N + i];
  if (C::KIND == 1) Tval = u0[traj * N + NS];
  double bT[NR];
  make_bT<C>(mp, Tval, bT);

  int nsave = sp.n_save;
  double tend = sp.t1;
  if (n_save_used) {
    int q = n_save_used[traj];
    if (q > 0 && q <= sp.n_save) { nsave = q; tend = __ldg(sp.saveat + q - 1); }
  }
  const double t0 = sp.t0;
  const double dtmax = tend - t0;
  const double dtmin = fmax(ulp_of(t0), ulp_of(tend));
  double* mypred = pred ? pred + (size_t)traj * sp.n_obs * sp.n_save : nullptr;

  auto save = [&](int ks, const double (&y)[NS]) {
    if (!mypred) return;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      int q = __ldg(sp.row2obs + i);
      if (q >= 0) {
        double v = (i < NS) ? y[i < NS ? i : 0] : Tval;
        mypred[q + sp.n_obs * ks] = clampd(v, sp.pred_lo, sp.pred_hi);
      }
    }
  };

  // ---- initial step size (Hairer-Wanner / ode_determine_initdt) ----
  int n_rhs = 0, n_acc = 0, n_rej = 0;
  rhs_value<C>(mp, bT, u, k1); ++n_rhs;
  double dt;
  {
    double s0 = 0.0, s1 = 0.0, sk[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      sk[i] = sp.abstol[i] + fabs(u[i]) * sp.reltol[i];
      double a = u[i] / sk[i], b = k1[i] / sk[i];
      s0 = fma(a, a, s0); s1 = fma(b, b, s1);
    }
    if (C::KIND == 1) { double a = Tval / (sp.abstol[NS] + fabs(Tval) * sp.reltol[NS]); s0 = fma(a, a, s0); }
    double d0 = sqrt(s0 / N), d1 = sqrt(s1 / N);
    double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
    dt0 = jmin(dt0, dtmax);
#pragma unroll
    for (int i = 0; i < NS; ++i) tmp[i] = fma(dt0, k1[i], u[i]);
    rhs_value<C>(mp, bT, tmp, k2); ++n_rhs;
    double s2 = 0.0;
#pragma unroll
    for (int i = 0; i < NS; ++i) { double b = (k2[i] - k1[i]) / sk[i]; s2 = fma(b, b, s2); }
    double d2 = sqrt(s2 / N) / dt0;
    double dm = jmax(d1, d2);
    double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : lean_exp10(-(2.0 + lean_log10(dm)) * sp.inv_order);
    dt = jmin(jmin(100.0 * dt0, dt1), dtmax);
  }

  double t = t0, qold = 1e-4, dt_last = 0.0;
  int isave = 0, ret = CRNN_RET_DEFAULT;
  long long iter = 0;
  while (isave < nsave && __ldg(sp.saveat + isave) <= t0) { save(isave, u); ++isave; }

  while (t < tend) {
    ++iter;
    if (dt != dt) { ret = CRNN_RET_DTNAN; break; }
    if (iter > sp.maxiters) { ret = CRNN_RET_MAXITERS; break; }
    dt = jmin(dt, dtmax);
    dt = jmin(dt, tend - t);
    if (dt <= dtmin && tend - t > dtmin) { ret = CRNN_RET_DTLESSTHANMIN; break; }
    bool bad = false;
#pragma unroll
    for (int i = 0; i < NS; ++i) bad |= (u[i] != u[i]);
    if (bad) { ret = CRNN_RET_UNSTABLE; break; }

#pragma unroll
    for (int i = 0; i < NS; ++i) tmp[i] = fma(dt, ts::a21 * k1[i], u[i]);
    rhs_value<C>(mp, bT, tmp, k2);
#pragma unroll
    for (int i = 0; i < NS; ++i) tmp[i] = fma(dt, fma(ts::a32, k2[i], ts::a31 * k1[i]), u[i]);
    rhs_value<C>(mp, bT, tmp, k3);
#pragma unroll
    for (int i = 0; i < NS; ++i)
      tmp[i] = fma(dt, fma(ts::a43, k3[i], fma(ts::a42, k2[i], ts::a41 * k1[i])), u[i]);
    rhs_value<C>(mp, bT, tmp, k4);
#pragma unroll
    for (int i = 0; i < NS; ++i)
      tmp[i] = fma(dt, fma(ts::a54, k4[i], fma(ts::a53, k3[i], fma(ts::a52, k2[i], ts::a51 * k1[i]))), u[i]);
    rhs_value<C>(mp, bT, tmp, k5);
#pragma unroll
    for (int i = 0; i < NS; ++i)
      tmp[i] = fma(dt, fma(ts::a65, k5[i], fma(ts::a64, k4[i], fma(ts::a63, k3[i], fma(ts::a62, k2[i], ts::a61 * k1[i])))), u[i]);
    rhs_value<C>(mp, bT, tmp, k6);
#pragma unroll
    for (int i = 0; i < NS; ++i)
      un[i] = fma(dt, fma(ts::a76, k6[i], fma(ts::a75, k5[i], fma(ts::a74, k4[i], fma(ts::a73, k3[i], fma(ts::a72, k2[i], ts::a71 * k1[i]))))), u[i]);
    rhs_value<C>(mp, bT, un, k7);
    n_rhs += 6;

    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      double e = dt * fma(ts::bt7, k7[i], fma(ts::bt6, k6[i], fma(ts::bt5, k5[i], fma(ts::bt4, k4[i],
                      fma(ts::bt3, k3[i], fma(ts::bt2, k2[i], ts::bt1 * k1[i]))))));
      double sc = fma(fmax(fabs(u[i]), fabs(un[i])), sp.reltol[i], sp.abstol[i]);
      double q = e / sc;
      acc = fma(q, q, acc);
    }
    const double EEst = sqrt(acc / N);
    double q11;
    const double q = pi_controller<C>(sp, EEst, qold, q11);
    dt_last = dt;
    if (EEst <= 1.0) {
      ++n_acc;
      qold = jmax(EEst, 1e-4);
      const double dtnew = dt / q;
      const double tprev = t;
      t = snap_t(t + dt, tend);
      while (isave < nsave) {
        const double tsv = __ldg(sp.saveat + isave);
        if (!(tsv <= t)) break;
        if (tsv == t) {
          save(isave, un);
        } else {
          double b[7];
          ts::dense_b((tsv - tprev) / dt, b);
#pragma unroll
          for (int i = 0; i < NS; ++i)
            tmp[i] = fma(dt, fma(b[6], k7[i], fma(b[5], k6[i], fma(b[4], k5[i], fma(b[3], k4[i],
                          fma(b[2], k3[i], fma(b[1], k2[i], b[0] * k1[i])))))), u[i]);
          save(isave, tmp);
        }
        ++isave;
      }
#pragma unroll
      for (int i = 0; i < NS; ++i) { u[i] = un[i]; k1[i] = k7[i]; }
      dt = jmin(dtnew, dtmax);
    } else {
      ++n_rej;
      dt = dt / jmin(sp.inv_qmin, q11 / sp.gamma);
    }
  }
  if (ret == CRNN_RET_DEFAULT) ret = CRNN_RET_SUCCESS;
  if (mypred)
    for (int ks = isave; ks < sp.n_save; ++ks)
      for (int q = 0; q < sp.n_obs; ++q) mypred[q + sp.n_obs * ks] = 0.0;
  if (n_saved) n_saved[traj] = isave;
  if (retcode) retcode[traj] = ret;
  if (stats) {
    crnn_stats s;
    s.n_accept = n_acc; s.n_reject = n_rej; s.n_rhs = n_rhs; s.n_jac = 0;
    s.t_reached = t; s.dt_last = dt_last;
    stats[traj] = s;
  }
}

}  // namespace crnn
