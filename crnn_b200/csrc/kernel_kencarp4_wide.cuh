// kernel_kencarp4_wide.cuh — stiff predict path for LARGE states (BASELINE config 5, "HyChem-sized"):
// one WARP owns one trajectory and lane i owns state component i (n_state <= 32, n_reac <= 32,
// dimensions are RUNTIME values, no template instantiation per model).
//
// KenCarp4 (ESDIRK4(3)6L[2]SA, Kennedy & Carpenter 2003) is not used anywhere in the
// reference — BASELINE asks for it — so its policies are ours and documented in oracle/crnn_oracle.c
// (solve_one_kencarp4), which this kernel mirrors operation by operation:
//   * W = I - gamma*dt*J(u_n) with the ANALYTIC CRNN Jacobian, one LU per step attempt, factored
//     cooperatively in shared memory (lane = row, partial pivoting by a warp arg-max, first strict maximum);
//   * simplified Newton per implicit stage (predictor z_i = z_{i-1}, eta*|dz| < 1/100, <= 10 iterations,
//     divergence => reject and halve dt), triangular solves with one shuffle broadcast per pivot;
//   * smoothed embedded error estimate, PI controller (order 4), Hermite dense output at the save points.
// The RHS is naturally lane-parallel here: lane i takes log(u_i), lane j takes exp(z_j) — one log and
// one exp issued per evaluation for the whole state.
#pragma once
#include "crnn_dev.cuh"
#include "wide_common.cuh"

namespace crnn {

namespace kc {
constexpr double g = 0.25;
__constant__ double A[6][5] = {
    {0, 0, 0, 0, 0},
    {0.25, 0, 0, 0, 0},
    {8611.0 / 62500.0, -1743.0 / 31250.0, 0, 0, 0},
    {5012029.0 / 34652500.0, -654441.0 / 2922500.0, 174375.0 / 388108.0, 0, 0},
    {15267082809.0 / 155376265600.0, -71443401.0 / 120774400.0, 730878875.0 / 902184768.0, 2285395.0 / 8070912.0, 0},
    {82889.0 / 524892.0, 0.0, 15625.0 / 83664.0, 69875.0 / 102672.0, -2260.0 / 8211.0}};
__constant__ double C[6] = {0.0, 0.5, 83.0 / 250.0, 31.0 / 50.0, 17.0 / 20.0, 1.0};
__constant__ double BHAT[6] = {4586570599.0 / 29645900160.0, 0.0, 178811875.0 / 945068544.0,
                               814220225.0 / 1159782912.0, -3700637.0 / 11593932.0, 61727.0 / 225920.0};
}  // namespace kc

template <int WARPS, bool F2, bool SPARSE = false>
__global__ void __launch_bounds__(WARPS * 32, 512 / (WARPS * 32))
k_kencarp4_wide(const __grid_constant__ WideP P, const double* __restrict__ u0,
                const int* __restrict__ n_save_used, long long ntraj, double* __restrict__ pred,
                int* __restrict__ n_saved, int* __restrict__ retcode, crnn_stats* __restrict__ stats,
                unsigned long long* __restrict__ queue) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WideBlock& sb = *reinterpret_cast<WideBlock*>(smem_raw);
  using KcWarp = WideWarpT<0>;  // the six stage values live in registers: 1.8 KB less shared memory per warp -> 16 warps/SM
  KcWarp* wws = reinterpret_cast<KcWarp*>(smem_raw + sizeof(WideBlock));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  KcWarp& ww = wws[warp];
  const int n = P.n, ns = P.ns, nin = P.nin, nr = P.nr;

  wide_block_init(P, sb);

  const bool isp = lane < ns;                      // lane owns a species
  const double my_at = lane < n ? P.abstol[lane] : 1.0, my_rt = lane < n ? P.reltol[lane] : 0.0;
  const int my_obs = lane < n ? P.row2obs[lane] : -1;

  const double my_mw = (F2 && lane < ns) ? __ldg(P.mw + lane) : 1.0;
  // f(y, t), W = I - gdt*J and the triangular solves come from wide_common.cuh (all RHS flavours)
  int tab_seg = 0;  // F2: segment hint of the T(t), P(t) lookup
  auto rhs = [&](double tt, double y, WideAux& ax) -> double { return wide_rhs<F2, SPARSE>(P, sb, ww, lane, my_mw, tt, y, ax, tab_seg); };
  // W^{-1} explicitly (Gauss-Jordan, in place) and mat-vec "solves": ~20 Newton solves share one factorisation
  auto lusolve = [&](double b) -> double { return wide_invmul(ww, lane, ns, b); };
  auto build_lu = [&](const double* rsrc, const WideAux& ax, double gdt) { wide_build_inv<F2, SPARSE>(P, sb, ww, lane, rsrc, ax, gdt); };
  // rms over the n state components of v_i / (atol_i + max(|a_i|,|b_i|) rtol_i)
  auto wrms = [&](double v, double a, double b) -> double {
    double q = 0.0;
    if (lane < n) { const double sc = my_at + fmax(fabs(a), fabs(b)) * my_rt; q = v / sc; q *= q; }
    return sqrt(wsum(q) / n);
  };
  while (true) {
    unsigned long long tq = 0;
    if (lane == 0) tq = atomicAdd(queue, 1ull);
    const long long traj = (long long)__shfl_sync(0xffffffffu, tq, 0);
    if (traj >= ntraj) break;

    double u = lane < n ? __ldg(u0 + traj * n + lane) : 0.0;
    int nsave = P.n_save;
    double tend = P.t1;
    if (n_save_used) {
      int q = __ldg(n_save_used + traj);
      if (q > 0 && q <= P.n_save) { nsave = q; tend = __ldg(P.saveat + q - 1); }
    }
    const double t0 = P.t0, dtmax = tend - t0;
    const double dtmin = fmax(ulp_of(t0), ulp_of(tend));
    double* mypred = pred ? pred + (size_t)traj * P.n_obs * P.n_save : nullptr;
    auto save = [&](int ks, double y) {
      if (mypred && my_obs >= 0) mypred[my_obs + P.n_obs * ks] = clampd(y, P.pred_lo, P.pred_hi);
    };

    int n_rhs = 0, n_acc = 0, n_rej = 0, n_jac = 0;
    WideAux a0, as;
    double f0 = rhs(t0, u, a0); ++n_rhs;
    ww.r0[lane] = ww.r[lane];
    // ---- initial step (Hairer-Wanner, order 4) ----
    double dt;
    {
      const double sk = my_at + fabs(u) * my_rt;
      double a = 0.0, b = 0.0;
      if (lane < n) { a = u / sk; a *= a; b = f0 / sk; b *= b; }
      const double d0 = sqrt(wsum(a) / n), d1 = sqrt(wsum(b) / n);
      double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
      dt0 = jmin(dt0, dtmax);
      const double f1p = rhs(t0 + dt0, fma(dt0, f0, u), as); ++n_rhs;
      double c = 0.0;
      if (lane < n) { c = (f1p - f0) / sk; c *= c; }
      const double d2 = sqrt(wsum(c) / n) / dt0;
      const double dm = jmax(d1, d2);
      const double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : lean_exp10(-(2.0 + lean_log10(dm)) * P.inv_order);
      dt = jmin(jmin(100.0 * dt0, dt1), dtmax);
    }
    double t = t0, qold = 1e-4, eta_old = 1.0, dt_last = 0.0;
    int isave = 0, ret = CRNN_RET_DEFAULT;
    long long iter = 0;
    while (isave < nsave && __ldg(P.saveat + isave) <= t0) { save(isave, u); ++isave; }

    while (t < tend) {
      ++iter;
      if (dt != dt) { ret = CRNN_RET_DTNAN; break; }
      if (iter > P.maxiters) { ret = CRNN_RET_MAXITERS; break; }
      dt = jmin(dt, dtmax);
      dt = jmin(dt, tend - t);
      if (dt <= dtmin && tend - t > dtmin) { ret = CRNN_RET_DTLESSTHANMIN; break; }
      if (__any_sync(0xffffffffu, lane < n && u != u)) { ret = CRNN_RET_UNSTABLE; break; }

      // ---- W = I - g dt J(u_n), J[i][l] = sum_j w_out[i,j] r0_j w_in[l,j] dx0_l ; LU ----
      const double gdt = kc::g * dt;
      build_lu(ww.r0, a0, gdt);
      ++n_jac;

      // ---- stages ----
      double z[6], tmp = u, yk = u;
      z[0] = dt * f0;
      bool ok = true, refreshed = false;
#pragma unroll 1
      for (int s = 1; s < 6 && ok; ++s) {
        tmp = u;
#pragma unroll
        for (int j = 0; j < 5; ++j)
          if (j < s) tmp = fma(kc::A[s][j], z[j], tmp);
        double zs = z[s - 1];
        bool conv = false;
#pragma unroll 1
        for (int attempt = 0; attempt < 2 && !conv; ++attempt) {
          double ndz_prev = 0.0, eta = lean_pow(fmax(eta_old, 2.220446049250313e-16), 0.8);
#pragma unroll 1
          for (int it = 1; it <= 10; ++it) {
            yk = fma(kc::g, zs, tmp);
            double dz = fma(dt, rhs(t + kc::C[s] * dt, yk, as), -zs); ++n_rhs;
            dz = lusolve(dz);
            const double ndz = wrms(dz, u, yk);
            zs += dz;
            if (it > 1) {
              const double theta = ndz / ndz_prev;
              if (!(theta <= 2.0)) break;
              eta = theta / (1.0 - theta);
            }
            if ((eta >= 0.0 && eta * ndz < 0.01) || ndz == 0.0) { conv = true; eta_old = eta; break; }
            ndz_prev = ndz;
          }
          if (!conv) {
            // a stage value that crossed the clamp sees a very different Jacobian: once per step
            // attempt rebuild W at the last iterate and redo this stage (oracle: solve_one_kencarp4)
            if (refreshed || __any_sync(0xffffffffu, lane < n && zs != zs)) break;
            refreshed = true;
            yk = fma(kc::g, zs, tmp);
            (void)rhs(t + kc::C[s] * dt, yk, as); ++n_rhs;
            build_lu(ww.r, as, gdt);
            ++n_jac;
          }
        }
        // z[s] = zs with a compile-time index
#pragma unroll
        for (int j = 1; j < 6; ++j)
          if (j == s) z[j] = zs;
        if (!conv) ok = false;
      }
      dt_last = dt;
      if (!ok) { ++n_rej; dt = dt / 2.0; continue; }
      const double un = fma(kc::g, z[5], tmp);
      double e = (kc::A[5][0] - kc::BHAT[0]) * z[0];   // the oracle's order: stages 1..5 ascending, the diagonal term last
#pragma unroll
      for (int j = 1; j < 5; ++j) e = fma(kc::A[5][j] - kc::BHAT[j], z[j], e);
      e = fma(kc::g - kc::BHAT[5], z[5], e);
      e = lusolve(isp ? e : 0.0);
      const double EEst = wrms(e, u, un);
      double q11, q;
      if (EEst == 0.0) { q11 = 0.0; q = P.inv_qmax; }
      else {
        q11 = lean_pow(EEst, P.beta1);
        q = jmax(P.inv_qmax, jmin(P.inv_qmin, q11 / lean_pow(qold, P.beta2) / P.gamma));
      }
      if (EEst <= 1.0) {
        ++n_acc;
        qold = jmax(EEst, 1e-4);
        const double dtnew = dt / (q >= P.qs_min && q <= P.qs_max ? 1.0 : q), tprev = t;  // steady-state dead-band
        t = snap_t(t + dt, tend);
        WideAux a1;
        const double f1 = rhs(t, un, a1); ++n_rhs;
        while (isave < nsave) {
          const double tsv = __ldg(P.saveat + isave);
          if (!(tsv <= t)) break;
          if (tsv == t) save(isave, un);
          else {
            const double th = (tsv - tprev) / dt;
            save(isave, (1.0 - th) * u + th * un +
                            th * (th - 1.0) * ((1.0 - 2.0 * th) * (un - u) + (th - 1.0) * dt * f0 + th * dt * f1));
          }
          ++isave;
        }
        u = un; f0 = f1; a0 = a1;
        __syncwarp();
        ww.r0[lane] = ww.r[lane];
        dt = jmin(dtnew, dtmax);
      } else {
        ++n_rej;
        dt = dt / jmin(P.inv_qmin, q11 / P.gamma);
      }
    }
    if (ret == CRNN_RET_DEFAULT) ret = CRNN_RET_SUCCESS;
    if (mypred && my_obs >= 0)
      for (int ks = isave; ks < P.n_save; ++ks) mypred[my_obs + P.n_obs * ks] = 0.0;
    if (lane == 0) {
      if (n_saved) n_saved[traj] = isave;
      if (retcode) retcode[traj] = ret;
      if (stats) {
        crnn_stats s;
        s.n_accept = n_acc; s.n_reject = n_rej; s.n_rhs = n_rhs; s.n_jac = n_jac;
        s.t_reached = t; s.dt_last = dt_last;
        stats[traj] = s;
      }
    }
    __syncwarp();
  }
}

}  // namespace crnn
