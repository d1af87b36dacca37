// kernel_rosenbrock23.cuh — stiff predict path: one THREAD owns one trajectory.
//
// Replaces `solve(_prob, Rosenbrock23(autodiff=true), saveat=tsteps[1:sample], maxiters=...)`
// (robertson/rober_crnn.jl:123-136).  Shampine-Reichelt ode23s with the
// ANALYTIC CRNN Jacobian J = W_out diag(r) W_in' diag(dx/du) (App. B.2) built from the r/dx
// the RHS already produced, W = I - d*dt*J factored by a fully unrolled in-register LU with
// partial pivoting, three triangular solves, two new RHS evaluations per step (FSAL).
// For F1 the temperature row of J is zero and k_T = 0, so the linear system is NS x NS.
#pragma once
#include "crnn_dev.cuh"

namespace crnn {

// RHS that also hands back r = exp(z) and dx_i = d log(clamp u_i)/du_i for the Jacobian.
template <class C>
__device__ __forceinline__ void rhs_value_full(const ModelP<C>& mp, const double (&bT)[C::NR],
                                               const double (&u)[C::NS], double (&du)[C::NS],
                                               double (&r)[C::NR], double (&dx)[C::NS]) {
  double x[C::NS];
#pragma unroll
  for (int i = 0; i < C::NS; ++i) {
    const double uc = clampd(u[i], mp.lb, mp.ub);
    x[i] = lean_log(uc);
    dx[i] = (u[i] >= mp.lb && u[i] <= mp.ub) ? 1.0 / uc : 0.0;
  }
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

// In-register LU with partial pivoting (same pivot rule as the oracle: first strict maximum).
template <int NS>
__device__ __forceinline__ void lu_factor(double (&A)[NS][NS], int (&piv)[NS]) {
#pragma unroll
  for (int k = 0; k < NS; ++k) {
    int p = k;
    double best = fabs(A[k][k]);
#pragma unroll
    for (int i = k + 1; i < NS; ++i) {
      const double v = fabs(A[i][k]);
      if (v > best) { best = v; p = i; }
    }
    piv[k] = p;
#pragma unroll
    for (int i = k + 1; i < NS; ++i) {
      const bool sw = (p == i);
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const double a = A[k][j], b = A[i][j];
        A[k][j] = sw ? b : a;
        A[i][j] = sw ? a : b;
      }
    }
    const double d = 1.0 / A[k][k];
    A[k][k] = d;  // the diagonal keeps 1/u_kk: back-substitution multiplies (an fp64 division is ~30 instructions)
#pragma unroll
    for (int i = k + 1; i < NS; ++i) {
      const double l = A[i][k] * d;
      A[i][k] = l;
#pragma unroll
      for (int j = k + 1; j < NS; ++j) A[i][j] = fma(-l, A[k][j], A[i][j]);
    }
  }
}

template <int NS>
__device__ __forceinline__ void lu_solve(const double (&A)[NS][NS], const int (&piv)[NS], double (&b)[NS]) {
#pragma unroll
  for (int k = 0; k < NS; ++k) {
#pragma unroll
    for (int i = k + 1; i < NS; ++i) {
      const bool sw = (piv[k] == i);
      const double x = b[k], y = b[i];
      b[k] = sw ? y : x;
      b[i] = sw ? x : y;
    }
  }
#pragma unroll
  for (int i = 1; i < NS; ++i) {
    double s = b[i];
#pragma unroll
    for (int j = 0; j < i; ++j) s = fma(-A[i][j], b[j], s);
    b[i] = s;
  }
#pragma unroll
  for (int i = NS - 1; i >= 0; --i) {
    double s = b[i];
#pragma unroll
    for (int j = i + 1; j < NS; ++j) s = fma(-A[i][j], b[j], s);
    b[i] = s * A[i][i];  // A[i][i] holds 1/u_ii (lu_factor)
  }
}

namespace rb {
constexpr double d = 0.29289321881345254;   // 1/(2+sqrt(2))
constexpr double e32 = 7.414213562373095;   // 6+sqrt(2)
constexpr double inv_1m2d = 2.414213562373095;  // 1/(1-2d)
}  // namespace rb

template <class C>
__global__ void __launch_bounds__(128)
k_rosenbrock23_value(const __grid_constant__ ModelP<C> mp, const __grid_constant__ SolveP<C> sp,
                     const double* __restrict__ u0, const int* __restrict__ n_save_used, long long ntraj,
                     double* __restrict__ pred, int* __restrict__ n_saved, int* __restrict__ retcode,
                     crnn_stats* __restrict__ stats) {
  constexpr int NS = C::NS, NR = C::NR, N = C::N;
  const long long traj = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (traj >= ntraj) return;

  double u[NS], un[NS], f0[NS], f1[NS], f2[NS], k1[NS], k2[NS], k3[NS], tmp[NS];
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
  rhs_value_full<C>(mp, bT, u, f0, r0, dx0); ++n_rhs;
  double dt;
  {
    double s0 = 0.0, s1 = 0.0, sk[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      sk[i] = sp.abstol[i] + fabs(u[i]) * sp.reltol[i];
      double a = u[i] / sk[i], b = f0[i] / sk[i];
      s0 = fma(a, a, s0); s1 = fma(b, b, s1);
    }
    if (C::KIND == 1) { double a = Tval / (sp.abstol[NS] + fabs(Tval) * sp.reltol[NS]); s0 = fma(a, a, s0); }
    double d0 = sqrt(s0 / N), d1 = sqrt(s1 / N);
    double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
    dt0 = jmin(dt0, dtmax);
#pragma unroll
    for (int i = 0; i < NS; ++i) tmp[i] = fma(dt0, f0[i], u[i]);
    rhs_value_full<C>(mp, bT, tmp, f1, r1, dx1); ++n_rhs;
    double s2 = 0.0;
#pragma unroll
    for (int i = 0; i < NS; ++i) { double b = (f1[i] - f0[i]) / sk[i]; s2 = fma(b, b, s2); }
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

    // W = I - gamma*J,  J[i][l] = (sum_j w_out[i,j] r_j w_in[l,j]) * dx_l
    const double g = rb::d * dt;
    double W[NS][NS];
    int piv[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i)
#pragma unroll
      for (int l = 0; l < NS; ++l) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < NR; ++j) s = fma(mp.w_out[i + NS * j] * r0[j], mp.w_in[l + C::NIN * j], s);
        W[i][l] = (i == l ? 1.0 : 0.0) - g * (s * dx0[l]);
      }
    ++n_jac;
    lu_factor<NS>(W, piv);
#pragma unroll
    for (int i = 0; i < NS; ++i) k1[i] = f0[i];
    lu_solve<NS>(W, piv, k1);
#pragma unroll
    for (int i = 0; i < NS; ++i) tmp[i] = fma(0.5 * dt, k1[i], u[i]);
    rhs_value_full<C>(mp, bT, tmp, f1, r1, dx1);
#pragma unroll
    for (int i = 0; i < NS; ++i) k2[i] = f1[i] - k1[i];
    lu_solve<NS>(W, piv, k2);
#pragma unroll
    for (int i = 0; i < NS; ++i) { k2[i] += k1[i]; un[i] = fma(dt, k2[i], u[i]); }
    rhs_value_full<C>(mp, bT, un, f2, r1, dx1);
    n_rhs += 2;
#pragma unroll
    for (int i = 0; i < NS; ++i) k3[i] = f2[i] - rb::e32 * (k2[i] - f1[i]) - 2.0 * (k1[i] - f0[i]);
    lu_solve<NS>(W, piv, k3);

    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      double e = dt / 6.0 * (k1[i] - 2.0 * k2[i] + k3[i]);
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
      const double dtnew = dt / (q >= sp.qs_min && q <= sp.qs_max ? 1.0 : q);  // steady-state dead-band
      const double tprev = t;
      t = snap_t(t + dt, tend);
      while (isave < nsave) {
        const double tsv = __ldg(sp.saveat + isave);
        if (!(tsv <= t)) break;
        if (tsv == t) {
          save(isave, un);
        } else {
          const double th = (tsv - tprev) / dt;
          const double c1 = th * (1.0 - th) * rb::inv_1m2d, c2 = th * (th - 2.0 * rb::d) * rb::inv_1m2d;
#pragma unroll
          for (int i = 0; i < NS; ++i) tmp[i] = fma(dt, fma(c2, k2[i], c1 * k1[i]), u[i]);
          save(isave, tmp);
        }
        ++isave;
      }
#pragma unroll
      for (int i = 0; i < NS; ++i) { u[i] = un[i]; f0[i] = f2[i]; dx0[i] = dx1[i]; }
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
