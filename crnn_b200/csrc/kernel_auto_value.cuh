// kernel_auto_value.cuh — predict_neuralode with AutoTsit5(Rosenbrock23()): one THREAD owns one trajectory.
//
// The composite algorithm of case2/case2.jl:26 (and HyChem/crnn_pyrolysis_mass.jl:29) for the dimension-specialised
// configurations: the Tsit5 step of kernel_tsit5_value.cuh and the Rosenbrock23 step of kernel_rosenbrock23.cuh in
// one kernel, selected per thread by OrdinaryDiffEq's AutoSwitch counter (semantics in oracle/crnn_oracle.c::solve_one
// and DESIGN.md §3.2d).  Stage storage is shared between the two halves: k[0..6] are Tsit5's k1..k7, or
// f0, k1, k2, k3, f1, f2 under Rosenbrock23.  For states this small the lane-per-component kernel (k_wide_solve) keeps
// 3-9 of 32 lanes busy; here every lane integrates its own trajectory.
#pragma once
#include "crnn_dev.cuh"
#include "kernel_rosenbrock23.cuh"

namespace crnn {

template <class C>
__global__ void __launch_bounds__(128)
k_auto_value(const __grid_constant__ ModelP<C> mp, const __grid_constant__ SolveP<C> sp,
             const double* __restrict__ u0, const int* __restrict__ n_save_used, long long ntraj,
             double* __restrict__ pred, int* __restrict__ n_saved, int* __restrict__ retcode,
             crnn_stats* __restrict__ stats) {
  constexpr int NS = C::NS, NR = C::NR, N = C::N;
  const long long traj = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (traj >= ntraj) return;

  double u[NS], un[NS], k[7][NS], tmp[NS];
  double r0[NR], dx0[NS], r1[NR], dx1[NS];
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

  int n_rhs = 0, n_acc = 0, n_rej = 0, n_jac = 0;
  rhs_value_full<C>(mp, bT, u, k[0], r0, dx0); ++n_rhs;
  double dt;
  {
    double s0 = 0.0, s1 = 0.0, sk[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      sk[i] = sp.abstol[i] + fabs(u[i]) * sp.reltol[i];
      double a = u[i] / sk[i], b = k[0][i] / sk[i];
      s0 = fma(a, a, s0); s1 = fma(b, b, s1);
    }
    if (C::KIND == 1) { double a = Tval / (sp.abstol[NS] + fabs(Tval) * sp.reltol[NS]); s0 = fma(a, a, s0); }
    double d0 = sqrt(s0 / N), d1 = sqrt(s1 / N);
    double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
    dt0 = jmin(dt0, dtmax);
#pragma unroll
    for (int i = 0; i < NS; ++i) tmp[i] = fma(dt0, k[0][i], u[i]);
    rhs_value<C>(mp, bT, tmp, k[1]); ++n_rhs;
    double s2 = 0.0;
#pragma unroll
    for (int i = 0; i < NS; ++i) { double b = (k[1][i] - k[0][i]) / sk[i]; s2 = fma(b, b, s2); }
    double d2 = sqrt(s2 / N) / dt0;
    double dm = jmax(d1, d2);
    double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : lean_exp10(-(2.0 + lean_log10(dm)) * sp.inv_order);
    dt = jmin(jmin(100.0 * dt0, dt1), dtmax);
  }

  double t = t0, qold = 1e-4, dt_last = 0.0, eigen_est = 0.0;
  int isave = 0, ret = CRNN_RET_DEFAULT, sw_count = 0;
  bool rosen = false;
  long long iter = 0;
  while (isave < nsave && __ldg(sp.saveat + isave) <= t0) { save(isave, u); ++isave; }

  while (t < tend) {
    ++iter;
    if (iter > 1) {  // AutoSwitch choice function, on the dt the controller just proposed
      const bool stiff = fabs(eigen_est * dt / 3.5068) > 0.9;
      sw_count = stiff ? (sw_count < 0 ? 1 : sw_count + 1) : (sw_count > 0 ? -1 : sw_count - 1);
      bool want = rosen;
      if (!rosen && sw_count > 10) { dt = dt * 2.0; want = true; }
      else if (rosen && sw_count < -3) { dt = dt / 2.0; want = false; }
      if (want != rosen) {
        rosen = want;
        rhs_value_full<C>(mp, bT, u, k[0], r0, dx0); ++n_rhs;  // initialize!(new cache): fsalfirst = f(uprev)
      }
    }
    if (dt != dt) { ret = CRNN_RET_DTNAN; break; }
    if (iter > sp.maxiters) { ret = CRNN_RET_MAXITERS; break; }
    dt = jmin(dt, dtmax);
    dt = jmin(dt, tend - t);
    if (dt <= dtmin && tend - t > dtmin) { ret = CRNN_RET_DTLESSTHANMIN; break; }
    bool bad = false;
#pragma unroll
    for (int i = 0; i < NS; ++i) bad |= (u[i] != u[i]);
    if (bad) { ret = CRNN_RET_UNSTABLE; break; }

    double acc = 0.0;
    if (!rosen) {
      // ---- Tsit5 (kernel_tsit5_value.cuh) ----
#pragma unroll
      for (int i = 0; i < NS; ++i) tmp[i] = fma(dt, ts::a21 * k[0][i], u[i]);
      rhs_value<C>(mp, bT, tmp, k[1]);
#pragma unroll
      for (int i = 0; i < NS; ++i) tmp[i] = fma(dt, fma(ts::a32, k[1][i], ts::a31 * k[0][i]), u[i]);
      rhs_value<C>(mp, bT, tmp, k[2]);
#pragma unroll
      for (int i = 0; i < NS; ++i)
        tmp[i] = fma(dt, fma(ts::a43, k[2][i], fma(ts::a42, k[1][i], ts::a41 * k[0][i])), u[i]);
      rhs_value<C>(mp, bT, tmp, k[3]);
#pragma unroll
      for (int i = 0; i < NS; ++i)
        tmp[i] = fma(dt, fma(ts::a54, k[3][i], fma(ts::a53, k[2][i], fma(ts::a52, k[1][i], ts::a51 * k[0][i]))), u[i]);
      rhs_value<C>(mp, bT, tmp, k[4]);
#pragma unroll
      for (int i = 0; i < NS; ++i)
        tmp[i] = fma(dt, fma(ts::a65, k[4][i], fma(ts::a64, k[3][i], fma(ts::a63, k[2][i], fma(ts::a62, k[1][i], ts::a61 * k[0][i])))), u[i]);
      rhs_value<C>(mp, bT, tmp, k[5]);
#pragma unroll
      for (int i = 0; i < NS; ++i)
        un[i] = fma(dt, fma(ts::a76, k[5][i], fma(ts::a75, k[4][i], fma(ts::a74, k[3][i], fma(ts::a73, k[2][i], fma(ts::a72, k[1][i], ts::a71 * k[0][i]))))), u[i]);
      rhs_value_full<C>(mp, bT, un, k[6], r1, dx1);
      n_rhs += 6;
      double num = 0.0, den = 0.0;  // eigen_est = ||k7 - k6|| / ||u_{n+1} - g6||, g6 = the stage-6 state (still in tmp)
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        const double e = dt * fma(ts::bt7, k[6][i], fma(ts::bt6, k[5][i], fma(ts::bt5, k[4][i], fma(ts::bt4, k[3][i],
                              fma(ts::bt3, k[2][i], fma(ts::bt2, k[1][i], ts::bt1 * k[0][i]))))));
        const double sc = fma(fmax(fabs(u[i]), fabs(un[i])), sp.reltol[i], sp.abstol[i]);
        const double q = e / sc;
        acc = fma(q, q, acc);
        const double a = k[6][i] - k[5][i], b = un[i] - tmp[i];
        num = fma(a, a, num); den = fma(b, b, den);
      }
      eigen_est = sqrt(num / N) / sqrt(den / N);
    } else {
      // ---- Rosenbrock23 (kernel_rosenbrock23.cuh): k[0]=f0, k[1..3]=k1..k3, k[4]=f1, k[5]=f2 ----
      const double g = rb::d * dt;
      double W[NS][NS];
      int piv[NS];
      double eig = 0.0;
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        double rowsum = 0.0;
#pragma unroll
        for (int l = 0; l < NS; ++l) {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < NR; ++j) s = fma(mp.w_out[i + NS * j] * r0[j], mp.w_in[l + C::NIN * j], s);
          const double Jil = s * dx0[l];
          rowsum += fabs(Jil);
          W[i][l] = (i == l ? 1.0 : 0.0) - g * Jil;
        }
        if (C::KIND == 1) {  // the temperature column takes part in opnorm(J, Inf) (its row is zero)
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < NR; ++j) s = fma(mp.w_out[i + NS * j] * r0[j], mp.w_in[NS + C::NIN * j], s);
          rowsum += fabs(s * (1.0 / (mp.gas_R * Tval * Tval)));
        }
        eig = fmax(eig, rowsum);
      }
      eigen_est = eig;
      ++n_jac;
      lu_factor<NS>(W, piv);
#pragma unroll
      for (int i = 0; i < NS; ++i) k[1][i] = k[0][i];
      lu_solve<NS>(W, piv, k[1]);
#pragma unroll
      for (int i = 0; i < NS; ++i) tmp[i] = fma(0.5 * dt, k[1][i], u[i]);
      rhs_value<C>(mp, bT, tmp, k[4]);
#pragma unroll
      for (int i = 0; i < NS; ++i) k[2][i] = k[4][i] - k[1][i];
      lu_solve<NS>(W, piv, k[2]);
#pragma unroll
      for (int i = 0; i < NS; ++i) { k[2][i] += k[1][i]; un[i] = fma(dt, k[2][i], u[i]); }
      rhs_value_full<C>(mp, bT, un, k[5], r1, dx1);
      n_rhs += 2;
#pragma unroll
      for (int i = 0; i < NS; ++i) k[3][i] = k[5][i] - rb::e32 * (k[2][i] - k[4][i]) - 2.0 * (k[1][i] - k[0][i]);
      lu_solve<NS>(W, piv, k[3]);
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        const double e = dt / 6.0 * (k[1][i] - 2.0 * k[2][i] + k[3][i]);
        const double sc = fma(fmax(fabs(u[i]), fabs(un[i])), sp.reltol[i], sp.abstol[i]);
        const double q = e / sc;
        acc = fma(q, q, acc);
      }
    }
    const double EEst = sqrt(acc / N);
    const double b1 = rosen ? sp.beta1_ros : sp.beta1, b2 = rosen ? sp.beta2_ros : sp.beta2;
    double q11, q;
    if (EEst == 0.0) { q11 = 0.0; q = sp.inv_qmax; }
    else {
      q11 = lean_pow(EEst, b1);
      q = jmax(sp.inv_qmax, jmin(sp.inv_qmin, q11 / lean_pow(qold, b2) / sp.gamma));
    }
    dt_last = dt;
    if (EEst <= 1.0) {
      ++n_acc;
      qold = jmax(EEst, 1e-4);
      const double dtnew = dt / (q >= sp.qs_min && q <= sp.qs_max ? 1.0 : q);  // steady-state dead-band
      const double tprev = t;
      t = snap_t(t + dt, tend);
      while (isave < nsave) {
        const double tsv = __ldg(sp.saveat + isave);
        if (!(tsv <= t)) break;
        if (tsv == t) {
          save(isave, un);
        } else if (!rosen) {
          double b[7];
          ts::dense_b((tsv - tprev) / dt, b);
#pragma unroll
          for (int i = 0; i < NS; ++i)
            tmp[i] = fma(dt, fma(b[6], k[6][i], fma(b[5], k[5][i], fma(b[4], k[4][i], fma(b[3], k[3][i],
                          fma(b[2], k[2][i], fma(b[1], k[1][i], b[0] * k[0][i])))))), u[i]);
          save(isave, tmp);
        } else {
          const double th = (tsv - tprev) / dt;
          const double c1 = th * (1.0 - th) * rb::inv_1m2d, c2 = th * (th - 2.0 * rb::d) * rb::inv_1m2d;
#pragma unroll
          for (int i = 0; i < NS; ++i) tmp[i] = fma(dt, fma(c2, k[2][i], c1 * k[1][i]), u[i]);
          save(isave, tmp);
        }
        ++isave;
      }
#pragma unroll
      for (int i = 0; i < NS; ++i) { u[i] = un[i]; k[0][i] = rosen ? k[5][i] : k[6][i]; dx0[i] = dx1[i]; }
#pragma unroll
      for (int j = 0; j < NR; ++j) r0[j] = r1[j];
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
    s.n_accept = n_acc; s.n_reject = n_rej; s.n_rhs = n_rhs; s.n_jac = n_jac;
    s.t_reached = t; s.dt_last = dt_last;
    stats[traj] = s;
  }
}

}  // namespace crnn
