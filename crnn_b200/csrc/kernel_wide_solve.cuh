// kernel_wide_solve.cuh — the GENERIC predict path: any CRNN the reference scripts define, with runtime
// dimensions (n_state <= 32, n_reac <= 32), every RHS flavour (F0, F1 Arrhenius/T-state, F2 HyChem mass
// fractions under tabulated T(t), P(t)) and the three explicit/linearly-implicit steppers of the reference:
//   Tsit5                              case1/case1.jl:28, case3/case3.jl:29
//   Rosenbrock23 (analytic J, df/dt)   robertson/rober_crnn.jl:33
//   AutoTsit5(Rosenbrock23())          case2/case2.jl:26, HyChem/crnn_pyrolysis_mass.jl:29
// One WARP owns one trajectory and lane i owns state component i (the layout of k_kencarp4_wide): the RHS
// issues ONE log and ONE exp per evaluation for the whole state, the Jacobian is assembled analytically in
// shared memory and LU-factored cooperatively.  It mirrors oracle/crnn_oracle.c::solve_one operation by
// operation (stage order, norm, PI controller, AutoSwitch counter, dense output).
//
// The dimension-specialised thread-per-trajectory kernels (kernel_tsit5_value.cuh, kernel_rosenbrock23.cuh)
// stay the fast path for the instantiated configurations; this kernel serves everything else.
#pragma once
#include "crnn_dev.cuh"
#include "wide_common.cuh"
#include "kernel_tsit5_adjoint.cuh"  // tsc:: tableau in constant memory

namespace crnn {

// STIFF selects the stiff stepper: 0 = Rosenbrock23 (alg ROSENBROCK23 / AUTO_TSIT5_ROS23), 1 = TRBDF2 (alg TRBDF2 /
// AUTO_TSIT5_TRBDF2: Cathode/src/network.jl:102, yeast_glycolysis.jl:33) - separate instantiations, so neither pays for the
// other's code (these kernels are instruction-fetch bound).
template <int WARPS, bool F2, int STIFF = 0, bool OBS = false, bool MLP = false>
__global__ void __launch_bounds__(WARPS * 32, WARPS <= 4 ? 3 : 2)
k_wide_solve(const __grid_constant__ WideP P, const double* __restrict__ u0, const int* __restrict__ n_save_used,
             long long ntraj, double* __restrict__ pred, int* __restrict__ n_saved, int* __restrict__ retcode,
             crnn_stats* __restrict__ stats, unsigned long long* __restrict__ queue) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WideBlock& sb = *reinterpret_cast<WideBlock*>(smem_raw);
  WideWarp* wws = reinterpret_cast<WideWarp*>(smem_raw + sizeof(WideBlock));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WideWarp& ww = wws[warp];
  const int n = P.n, ns = P.ns, nin = P.nin, nr = P.nr;

  wide_block_init(P, sb);

  const bool autosw = (P.alg == CRNN_ALG_AUTO_TSIT5_ROS23 || P.alg == CRNN_ALG_AUTO_TSIT5_TRBDF2);
  const double my_at = lane < n ? P.abstol[lane] : 1.0, my_rt = lane < n ? P.reltol[lane] : 0.0;
  const int my_obs = lane < n ? P.row2obs[lane] : -1;
  const double my_mw = (F2 && lane < ns) ? __ldg(P.mw + lane) : 1.0;
  double* const kk = &ww.k[0][lane];  // this lane's stage column, stride KW_MAXN
#define KS(s) kk[(s) * KW_MAXN]

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
      const int q = __ldg(n_save_used + traj);
      if (q > 0 && q <= P.n_save) { nsave = q; tend = __ldg(P.saveat + q - 1); }
    }
    const double t0 = P.t0, dtmax = tend - t0;
    const double dtmin = fmax(ulp_of(t0), ulp_of(tend));
    double* mypred = pred ? pred + (size_t)traj * P.n_obs * P.n_save : nullptr;
    int n_rhs = 0, n_acc = 0, n_rej = 0, n_jac = 0, tab_seg = 0;
    WideAux a0, as;  // by-products at u_n / at the last evaluation
    // a saved state: the clamped rows of u, or - with the observable post-map (heat release = HRR_getter(ts, sol) * w_delH,
    // Cathode/src/network.jl:82-91,121) - y = sum_j w_obs[j] r_j(u(ts), ts) from one more evaluation at the saved state
    // (not counted in n_rhs, like the oracle's emit_save; the step's own by-products in ww.r are put back)
    auto save = [&](int ks, double tsv, double y) {
      if (OBS) {   // its own instantiations: the extra RHS copy is code the other models should not carry
        __syncwarp();
        const double rkeep = ww.r[lane];
        WideAux ao; int seg2 = tab_seg;
        (void)wide_rhs<F2, false, MLP>(P, sb, ww, lane, my_mw, tsv, y, ao, seg2);
        const double obs = wsum(lane < nr ? __ldg(P.w_obs + lane) * ww.r[lane] : 0.0);
        __syncwarp();
        ww.r[lane] = rkeep;
        __syncwarp();
        if (mypred && lane == 0) mypred[P.n_obs * ks] = clampd(obs, P.pred_lo, P.pred_hi);
        return;
      }
      if (mypred && my_obs >= 0) mypred[my_obs + P.n_obs * ks] = clampd(y, P.pred_lo, P.pred_hi);
    };
    KS(0) = wide_rhs<F2, false, MLP>(P, sb, ww, lane, my_mw, t0, u, a0, tab_seg); ++n_rhs;
    ww.r0[lane] = ww.r[lane];
    // ---- initial step (Hairer-Wanner; the order of the FIRST algorithm) ----
    double dt;
    {
      const double f0 = KS(0);
      const double sk = my_at + fabs(u) * my_rt;
      double a = 0.0, b = 0.0;
      if (lane < n) { a = u / sk; a *= a; b = f0 / sk; b *= b; }
      const double d0 = sqrt(wsum(a) / n), d1 = sqrt(wsum(b) / n);
      double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
      dt0 = jmin(dt0, dtmax);
      const double f1p = wide_rhs<F2, false, MLP>(P, sb, ww, lane, my_mw, t0 + dt0, fma(dt0, f0, u), as, tab_seg); ++n_rhs;
      double c = 0.0;
      if (lane < n) { c = (f1p - f0) / sk; c *= c; }
      const double d2 = sqrt(wsum(c) / n) / dt0;
      const double dm = jmax(d1, d2);
      const double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : lean_exp10(-(2.0 + lean_log10(dm)) * P.inv_order);
      dt = jmin(jmin(100.0 * dt0, dt1), dtmax);
    }
    double t = t0, qold = 1e-4, dt_last = 0.0, eigen_est = 0.0;
    int isave = 0, ret = CRNN_RET_DEFAULT, sw_count = 0;
    bool rosen = (P.alg == CRNN_ALG_ROSENBROCK23 || P.alg == CRNN_ALG_TRBDF2);   // "the stiff stepper is active"
    double eta_old = 1.0;   // TRBDF2: the Newton solver's eta, kept across steps
    long long iter = 0;
    while (isave < nsave && __ldg(P.saveat + isave) <= t0) { save(isave, __ldg(P.saveat + isave), u); ++isave; }

    while (t < tend) {
      ++iter;
      if (autosw && iter > 1) {  // OrdinaryDiffEq AutoSwitch (oracle solve_one header comment)
        const bool stiff = fabs(eigen_est * dt / 3.5068) > 0.9;
        sw_count = stiff ? (sw_count < 0 ? 1 : sw_count + 1) : (sw_count > 0 ? -1 : sw_count - 1);
        bool want = rosen;
        if (!rosen && sw_count > 10) { dt = dt * 2.0; want = true; }
        else if (rosen && sw_count < -3) { dt = dt / 2.0; want = false; }
        if (want != rosen) {
          rosen = want;
          KS(0) = wide_rhs<F2, false, MLP>(P, sb, ww, lane, my_mw, t, u, a0, tab_seg); ++n_rhs;  // initialize!: fsalfirst = f(uprev)
          __syncwarp();
          ww.r0[lane] = ww.r[lane];
        }
      }
      if (dt != dt) { ret = CRNN_RET_DTNAN; break; }
      if (iter > P.maxiters) { ret = CRNN_RET_MAXITERS; break; }
      dt = jmin(dt, dtmax);
      dt = jmin(dt, tend - t);
      if (dt <= dtmin && tend - t > dtmin) { ret = CRNN_RET_DTLESSTHANMIN; break; }
      if (__any_sync(0xffffffffu, lane < n && u != u)) { ret = CRNN_RET_UNSTABLE; break; }

      double un, e;
      if (!rosen) {
        // ---- Tsit5 ----
        double g6 = u;
#pragma unroll 1
        for (int s = 1; s < 7; ++s) {
          double acc = tsc::A[s][0] * KS(0);
          for (int j = 1; j < s; ++j) acc = fma(tsc::A[s][j], KS(j), acc);
          const double y = fma(dt, acc, u);
          if (s == 5) g6 = y;
          un = y;
          KS(s) = wide_rhs<F2, false, MLP>(P, sb, ww, lane, my_mw, t + tsc::C[s] * dt, y, as, tab_seg); ++n_rhs;
        }
        double acc = tsc::BT[0] * KS(0);
#pragma unroll
        for (int j = 1; j < 7; ++j) acc = fma(tsc::BT[j], KS(j), acc);
        e = dt * acc;
        if (autosw) {
          double a = 0.0, b = 0.0;
          if (lane < n) { a = KS(6) - KS(5); a *= a; b = un - g6; b *= b; }
          eigen_est = sqrt(wsum(a) / n) / sqrt(wsum(b) / n);
        }
      } else if (STIFF == 1) {
        // ---- TRBDF2 as an ESDIRK (oracle solve_one, "TRBDF2" branch): z1 = dt f0 (FSAL), z_g at t + gamma dt, z3 at t + dt,
        //      W = I - d dt J(u_n) inverted explicitly (one Gauss-Jordan per attempt, every Newton solve a mat-vec) ----
        constexpr double s2 = 1.4142135623730951, gam = 2.0 - s2, d = 1.0 - s2 / 2.0, w = s2 / 4.0;
        constexpr double bt1 = (1.0 - s2) / 3.0, bt2 = 1.0 / 3.0, bt3 = (s2 - 2.0) / 3.0, al1 = -s2 / 2.0, al2 = 1.0 + s2 / 2.0;
        const double gdt = d * dt;
        auto rhs_fd = [&](double tt, double yy) -> double { WideAux ax; return wide_rhs<F2, false, MLP>(P, sb, ww, lane, my_mw, tt, yy, ax, tab_seg); };
        double eig;
        if (MLP) { eig = wide_assemble_W_fd(P, ww, lane, t, u, gdt, rhs_fd); wide_invert(ww, lane, ns); }   // the scripts' autodiff=false
        else eig = wide_build_inv<F2>(P, sb, ww, lane, ww.r0, a0, gdt);
        ++n_jac;
        if (autosw) eigen_est = eig;
        const double z1 = dt * KS(0);
        double zg = z1, z3 = 0.0, tmp = u;
        bool ok = true, refreshed = false;
#pragma unroll 1
        for (int stg = 0; stg < 2 && ok; ++stg) {
          double zs;
          if (stg == 0) { tmp = fma(d, z1, u); zs = z1; }
          else { tmp = fma(w, zg, fma(w, z1, u)); zs = fma(al2, zg, al1 * z1); }
          const double tst = t + (stg ? 1.0 : gam) * dt;
          bool conv = false;
#pragma unroll 1
          for (int attempt = 0; attempt < 2 && !conv; ++attempt) {
            double ndz_prev = 0.0, eta = lean_pow(fmax(eta_old, 2.220446049250313e-16), 0.8);
#pragma unroll 1
            for (int it = 1; it <= 10; ++it) {
              const double yk = fma(d, zs, tmp);
              double dz = fma(dt, wide_rhs<F2, false, MLP>(P, sb, ww, lane, my_mw, tst, yk, as, tab_seg), -zs); ++n_rhs;
              dz = wide_invmul(ww, lane, ns, dz);
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
              if (refreshed || __any_sync(0xffffffffu, lane < n && zs != zs)) break;
              refreshed = true;
              (void)wide_rhs<F2, false, MLP>(P, sb, ww, lane, my_mw, tst, fma(d, zs, tmp), as, tab_seg); ++n_rhs;
              if (MLP) { (void)wide_assemble_W_fd(P, ww, lane, tst, fma(d, zs, tmp), gdt, rhs_fd); wide_invert(ww, lane, ns); }
              else (void)wide_build_inv<F2>(P, sb, ww, lane, ww.r, as, gdt);
              ++n_jac;
            }
          }
          if (stg == 0) zg = zs; else z3 = zs;
          if (!conv) ok = false;
        }
        if (!ok) { dt_last = dt; ++n_rej; dt = dt / 2.0; continue; }
        un = fma(d, z3, tmp);
        e = wide_invmul(ww, lane, ns, lane < ns ? fma(bt3, z3, fma(bt2, zg, bt1 * z1)) : 0.0);
        KS(5) = z3 / dt;   // fsallast
      } else {
        // ---- Rosenbrock23 = ode23s: KS(0)=f0, KS(1..3)=k1..k3, KS(4)=f1, KS(5)=f2 ----
        const double d = 1.0 / (2.0 + 1.4142135623730951), e32 = 6.0 + 1.4142135623730951;
        const double g = d * dt;
        const double dTv = wide_time_deriv<F2>(P, sb, ww, lane, t, ww.r0, a0, tab_seg);
        double eig;
        if (MLP) {
          auto rhs_fd = [&](double tt, double yy) -> double { WideAux ax; return wide_rhs<F2, false, MLP>(P, sb, ww, lane, my_mw, tt, yy, ax, tab_seg); };
          eig = wide_assemble_W_fd(P, ww, lane, t, u, g, rhs_fd); wide_factor_lu(ww, lane, ns);
        } else eig = wide_build_lu<F2>(P, sb, ww, lane, ww.r0, a0, g);
        ++n_jac;
        if (autosw) eigen_est = eig;
        const double f0 = KS(0);
        const double k1 = wide_lusolve(ww, lane, ns, fma(g, dTv, f0));
        const double f1 = wide_rhs<F2, false, MLP>(P, sb, ww, lane, my_mw, t + 0.5 * dt, fma(0.5 * dt, k1, u), as, tab_seg); ++n_rhs;
        const double k2 = wide_lusolve(ww, lane, ns, f1 - k1) + k1;
        un = fma(dt, k2, u);
        const double f2 = wide_rhs<F2, false, MLP>(P, sb, ww, lane, my_mw, t + dt, un, as, tab_seg); ++n_rhs;
        const double k3 = wide_lusolve(ww, lane, ns, f2 - e32 * (k2 - f1) - 2.0 * (k1 - f0) + dt * dTv);
        e = dt / 6.0 * (k1 - 2.0 * k2 + k3);
        KS(1) = k1; KS(2) = k2; KS(5) = f2;
      }
      const double EEst = wrms(e, u, un);
      const double b1 = (autosw && rosen) ? P.beta1_ros : P.beta1, b2 = (autosw && rosen) ? P.beta2_ros : P.beta2;
      double q11, q;
      if (EEst == 0.0) { q11 = 0.0; q = P.inv_qmax; }
      else {
        q11 = lean_pow(EEst, b1);
        q = jmax(P.inv_qmax, jmin(P.inv_qmin, q11 / lean_pow(qold, b2) / P.gamma));
      }
      dt_last = dt;
      if (EEst <= 1.0) {
        ++n_acc;
        qold = jmax(EEst, 1e-4);
        const double dtnew = dt / (q >= P.qs_min && q <= P.qs_max ? 1.0 : q), tprev = t;  // steady-state dead-band
        t = snap_t(t + dt, tend);
        if (STIFF == 1 && rosen) {   // the next attempt's analytic Jacobian needs the RHS by-products at u_{n+1}
          (void)wide_rhs<F2, false, MLP>(P, sb, ww, lane, my_mw, t, un, as, tab_seg); ++n_rhs;
        }
        while (isave < nsave) {
          const double tsv = __ldg(P.saveat + isave);
          if (!(tsv <= t)) break;
          if (tsv == t) save(isave, tsv, un);
          else {
            const double th = (tsv - tprev) / dt;
            if (!rosen) {
              double acc = 0.0;
#pragma unroll
              for (int s = 0; s < 7; ++s) {
                const double bs = th * (tsc::R[s][0] + th * (tsc::R[s][1] + th * (tsc::R[s][2] + th * tsc::R[s][3])));
                acc = s == 0 ? bs * KS(0) : fma(bs, KS(s), acc);
              }
              save(isave, tsv, fma(dt, acc, u));
            } else if (STIFF == 1) {   // Hermite on (u_n, fsalfirst) .. (u_{n+1}, fsallast)
              save(isave, tsv, (1.0 - th) * u + th * un +
                              th * (th - 1.0) * ((1.0 - 2.0 * th) * (un - u) + (th - 1.0) * dt * KS(0) + th * dt * KS(5)));
            } else {
              const double d = 1.0 / (2.0 + 1.4142135623730951);
              const double c1 = th * (1.0 - th) / (1.0 - 2.0 * d), c2 = th * (th - 2.0 * d) / (1.0 - 2.0 * d);
              save(isave, tsv, u + dt * (c1 * KS(1) + c2 * KS(2)));
            }
          }
          ++isave;
        }
        u = un;
        KS(0) = rosen ? KS(5) : KS(6);  // FSAL
        a0 = as;
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
#undef KS
}

}  // namespace crnn
