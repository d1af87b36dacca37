// kernel_rosenbrock23_sens.cuh — stiff loss + forward-mode gradient: one WARP owns one trajectory.
//
// Replaces `ForwardDiff.gradient(x -> loss_neuralode(x, i_exp; sample), p)` over
// `solve(_prob, Rosenbrock23(autodiff=true), saveat=tsteps[1:sample], maxiters=...)`
// (robertson/rober_crnn.jl:123-144,219): duals pushed through Shampine-Reichelt ode23s
// (SURVEY App. C.4).  With autodiff=true under an outer ForwardDiff the Jacobian itself is
// dual-valued, so the sensitivity of each linear solve W k = rhs carries the directional
// second derivative dJ[S_c, dW_c] k (SURVEY §7.3 "Rosenbrock sensitivities"):
//     W kdot_c = rhsdot_c + gamma * dJ_c * k .
// Same lane layout as k_tsit5_sens (lane = dual column, column 0 = value, structured R1 seed
// columns), same phase-machine discipline: ONE instance of the warp-cooperative RHS, ONE of the
// linear-solve block, ONE of the norm reduction, ONE of the loss/gradient block.
//   * J = W_out diag(r) W_in' diag(dx) is rebuilt analytically from the r / dx cached at u_n
//     (FSAL), W = I - d*dt*J is LU-factored redundantly in every lane's registers (NS <= 6),
//     the value lane's k is broadcast through shared memory for the dJ*k products.
//   * dense output of order 2 from k1, k2; loss and gradient fused at the save points.
#pragma once
#include "crnn_dev.cuh"
#include "kernel_rosenbrock23.cuh"
#include "kernel_tsit5_sens.cuh"

namespace crnn {

template <class C, int CT>
struct alignas(16) RosWarpBuf {
  double K[5][CT][C::NS][32];  // slots: f0 / f2 (roles swap on accept), k1, k2, f1
  double red[C::NS][32];
  double y[C::N], x[C::N], dx[C::N], r[C::NR], g[C::N];
  double x0[C::N], dx0[C::N], r0[C::NR];  // RHS intermediates cached at u_n (Jacobian, dJ*k)
  double v[C::N];                         // value lane's k for the dJ*k products
  double term[2][C::N];
  double rp[C::NS][8];                    // partial row sums of the norm reduction (RP lanes share one row of `red`)
};

// DEVW = true: weights come from device memory (mp_dev, written by a p2vec kernel of the on-device training loop, kernel_train.cuh)
template <class C, int CT, int WARPS, int MINB, bool DEVW = false>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_rosenbrock23_sens(const __grid_constant__ ModelP<C> mp, const __grid_constant__ SolveP<C> sp,
                    const double* __restrict__ seed_dev, const R1Desc* __restrict__ desc_dev, int ncol,
                    const double* __restrict__ u0, const int* __restrict__ n_save_used, long long ntraj,
                    const double* __restrict__ data, double* __restrict__ loss, double* __restrict__ grad_each,
                    double* __restrict__ pred, int* __restrict__ n_saved, int* __restrict__ retcode,
                    crnn_stats* __restrict__ stats, unsigned long long* __restrict__ queue,
                    const long long* __restrict__ in_idx, const ModelP<C>* __restrict__ mp_dev = nullptr) {
  constexpr int NS = C::NS, NR = C::NR, N = C::N, NIN = C::NIN;
  static_assert(NS <= 6, "per-lane register LU");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SensSmem<C, CT, true>& sm = *reinterpret_cast<SensSmem<C, CT, true>*>(smem_raw);
  RosWarpBuf<C, CT>* wbs = reinterpret_cast<RosWarpBuf<C, CT>*>(smem_raw + sizeof(SensSmem<C, CT, true>));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  RosWarpBuf<C, CT>& wb = wbs[warp];

  for (int q = threadIdx.x; q < 2 * NR * 32 * CT; q += blockDim.x) (&sm.seed[0][0])[q] = seed_dev[q];
  if (DEVW) {
    for (int q = threadIdx.x; q < (int)(sizeof(ModelP<C>) / sizeof(double)); q += blockDim.x)
      reinterpret_cast<double*>(&sm.mpw)[q] = reinterpret_cast<const double*>(mp_dev)[q];
    __syncthreads();
  }
  const ModelP<C>& M = DEVW ? sm.mpw : mp;   // compile-time choice: shared-memory loads or constant-bank operands
  for (int q = threadIdx.x; q < NIN * NR; q += blockDim.x) sm.w_in[q] = M.w_in[q];
  for (int q = threadIdx.x; q < NR; q += blockDim.x) sm.w_b[q] = M.w_b[q];
  for (int q = threadIdx.x; q < N; q += blockDim.x) {
    sm.inv_ys[q] = sp.inv_yscale[q];
    sm.row2obs[q] = sp.row2obs[q];
    sm.abstol[q] = sp.abstol[q];
    sm.reltol[q] = sp.reltol[q];
  }
  __syncthreads();
  const int np = ncol - 1;

  bool isval[CT], live[CT];
  double d_o[CT];
  int d_iin[CT], d_iout[CT], d_jout[CT];
#pragma unroll
  for (int t = 0; t < CT; ++t) {
    isval[t] = (t == 0 && lane == 0);
    const R1Desc d = desc_dev[lane + 32 * t];
    d_o[t] = d.o; d_iin[t] = d.i_in; d_iout[t] = d.i_out; d_jout[t] = d.j_out;
    live[t] = (lane + 32 * t) < ncol && (sp.incl_sens || isval[t]);
  }
  double my_at = 0.0, my_rt = 0.0;
  if (lane < NS) { my_at = sm.abstol[lane]; my_rt = sm.reltol[lane]; }

  // phases: F0/F1 = initial-step probes; K1 = factor W and solve k1; E1 = f1 and k2; E2 = f2, k3, error
  constexpr int PH_F0 = 0, PH_F1 = 1, PH_K1 = 2, PH_E1 = 3, PH_E2 = 4, PH_SAVE = 5;
  constexpr int S_K1 = 1, S_K2 = 2, S_F1 = 3;

  while (true) {
    unsigned long long tq = 0;
    if (lane == 0) tq = atomicAdd(queue, 1ull);
    const long long traj = (long long)__shfl_sync(0xffffffffu, tq, 0);
    if (traj >= ntraj) break;
    const long long src = in_idx ? __ldg(in_idx + traj) : traj;  // dataset row of the inputs (outputs stay at traj)
    const double* __restrict__ u0t = u0 + src * N;

    double U[CT][NS], Y[CT][NS], KO[CT][NS];
    double Tval = 0.0, xT = 0.0, mybT = 0.0, my_sk = 1.0, my_u0 = 0.0;
    if (C::KIND == 1) { Tval = __ldg(u0t + NS); xT = -1.0 / (M.gas_R * Tval); }
    if (lane < NR) {
      mybT = sm.w_b[lane];
      if (C::KIND == 1) mybT = fma(sm.w_in[NS + NIN * lane], xT, mybT);
    }
    if (C::KIND == 1 && lane == 0) { wb.x[NS] = xT; wb.x0[NS] = xT; }
#pragma unroll
    for (int t = 0; t < CT; ++t)
#pragma unroll
      for (int i = 0; i < NS; ++i) U[t][i] = isval[t] ? __ldg(u0t + i) : 0.0;
    if (lane < NS) {
      my_u0 = __ldg(u0t + lane);
      my_sk = my_at + fabs(my_u0) * my_rt;
    }

    int nsave = sp.n_save;
    double tend = sp.t1;
    if (n_save_used) {
      int q = __ldg(n_save_used + traj);
      if (q > 0 && q <= sp.n_save) { nsave = q; tend = __ldg(sp.saveat + q - 1); }
    }
    const double t0 = sp.t0, dtmax = tend - t0;
    const double dtmin = fmax(ulp_of(t0), ulp_of(tend));
    const size_t pbase = (size_t)traj * sp.n_obs * sp.n_save;
    const double* __restrict__ datat = data + (size_t)src * sp.n_obs * sp.n_save - pbase;  // datat + off reads row src

    int n_rhs = 0, n_acc = 0, n_rej = 0, n_jac = 0;
    double G[CT], loss_acc = 0.0;
#pragma unroll
    for (int t = 0; t < CT; ++t) G[t] = 0.0;
    double asum = my_u0 * my_u0, bsum = 0.0;
    double t = t0, tprev = t0, dt = 0.0, dt0 = 0.0, d1 = 0.0, dtnew = 0.0, qold = 1e-4, dt_last = 0.0, gam = 0.0;
    int isave = 0, ret = CRNN_RET_DEFAULT, phase = PH_F0, f0s = 0;  // f0s: slot of f0 (0 or 4), f2 in 4-f0s
    long long iter = 0;
    double W[NS][NS];
    int piv[NS];

    while (true) {
      const bool do_eval = (phase == PH_F0 || phase == PH_F1 || phase == PH_E1 || phase == PH_E2);
      if (do_eval) {
        // ---- KO = f(Y) on all columns: the single RHS instance ----
        if (phase == PH_F0) {
#pragma unroll
          for (int tt = 0; tt < CT; ++tt)
#pragma unroll
            for (int i = 0; i < NS; ++i) Y[tt][i] = U[tt][i];
        }
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < NS; ++i) wb.y[i] = Y[0][i];
        }
        __syncwarp();
        if (lane < NS) {
          const double yi = wb.y[lane];
          const double uc = clampd(yi, M.lb, M.ub);
          const bool inside = (yi >= M.lb) && (yi <= M.ub);
          wb.x[lane] = lean_log(uc);
          wb.dx[lane] = inside ? __drcp_rn(uc) : 0.0;
        }
        __syncwarp();
        if (lane < NR) {
          double z = mybT;
#pragma unroll
          for (int i = 0; i < NS; ++i) z = fma(sm.w_in[i + NIN * lane], wb.x[i], z);
          wb.r[lane] = lean_exp(z);
        }
        __syncwarp();
        {
          double dx[NS], r[NR];
#pragma unroll
          for (int i = 0; i < NS; ++i) dx[i] = wb.dx[i];
#pragma unroll
          for (int j = 0; j < NR; ++j) r[j] = wb.r[j];
#pragma unroll
          for (int tt = 0; tt < CT; ++tt) {
            const int lc = lane + 32 * tt;
            double q[NR];
            const double xin = wb.x[d_iin[tt]];
#pragma unroll
            for (int j = 0; j < NR; ++j) {
              double zd = fma(sm.seed[j][lc], xin, sm.seed[NR + j][lc]);
#pragma unroll
              for (int i = 0; i < NS; ++i) zd = fma(M.w_in[i + NIN * j], Y[tt][i] * dx[i], zd);
              if (isval[tt]) zd = 1.0;
              q[j] = r[j] * zd;
            }
            const double ro = d_o[tt] * wb.r[d_jout[tt]];
#pragma unroll
            for (int i = 0; i < NS; ++i) {
              double s = 0.0;
#pragma unroll
              for (int j = 0; j < NR; ++j) s = fma(M.w_out[i + NS * j], q[j], s);
              if (d_iout[tt] == i) s += ro;
              KO[tt][i] = s;
            }
          }
        }
        ++n_rhs;
        // store: F0 -> f0 ; F1 -> scratch slot F1 ; E1 -> f1 ; E2 -> f2
        {
          const int dst = (phase == PH_F0) ? f0s : (phase == PH_E2 ? 4 - f0s : S_F1);
#pragma unroll
          for (int tt = 0; tt < CT; ++tt)
#pragma unroll
            for (int i = 0; i < NS; ++i) wb.K[dst][tt][i][lane] = KO[tt][i];
        }
        if (phase == PH_F0) {  // cache the RHS intermediates at u_n for the Jacobian and dJ*k
          __syncwarp();
          if (lane < NS) { wb.x0[lane] = wb.x[lane]; wb.dx0[lane] = wb.dx[lane]; }
          if (lane < NR) wb.r0[lane] = wb.r[lane];
          __syncwarp();
        }
      }

      if (phase == PH_K1 || phase == PH_E1 || phase == PH_E2) {
        // ---- the single linear-solve instance: k = W \ (rhs + gamma * dJ * k_value) ----
        if (phase == PH_K1) {
          gam = rb::d * dt;
          double dx0[NS], r0[NR];
#pragma unroll
          for (int i = 0; i < NS; ++i) dx0[i] = wb.dx0[i];
#pragma unroll
          for (int j = 0; j < NR; ++j) r0[j] = wb.r0[j];
#pragma unroll
          for (int i = 0; i < NS; ++i)
#pragma unroll
            for (int l = 0; l < NS; ++l) {
              double s = 0.0;
#pragma unroll
              for (int j = 0; j < NR; ++j) s = fma(M.w_out[i + NS * j] * r0[j], M.w_in[l + NIN * j], s);
              W[i][l] = (i == l ? 1.0 : 0.0) - gam * (s * dx0[l]);
            }
          ++n_jac;
          lu_factor<NS>(W, piv);
        }
        // right-hand side without the dJ term (per column)
#pragma unroll
        for (int tt = 0; tt < CT; ++tt)
#pragma unroll
          for (int i = 0; i < NS; ++i) {
            const double f0 = wb.K[f0s][tt][i][lane];
            if (phase == PH_K1) {
              KO[tt][i] = f0;
            } else if (phase == PH_E1) {
              KO[tt][i] = KO[tt][i] - wb.K[S_K1][tt][i][lane];  // f1 - k1
            } else {
              const double k1 = wb.K[S_K1][tt][i][lane], k2 = wb.K[S_K2][tt][i][lane], f1 = wb.K[S_F1][tt][i][lane];
              KO[tt][i] = KO[tt][i] - rb::e32 * (k2 - f1) - 2.0 * (k1 - f0);  // f2 - e32 (k2 - f1) - 2 (k1 - f0)
            }
          }
        // value lane solves first and publishes its k
        double vk[NS];
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < NS; ++i) vk[i] = KO[0][i];
          lu_solve<NS>(W, piv, vk);
#pragma unroll
          for (int i = 0; i < NS; ++i) wb.v[i] = vk[i];
        }
        __syncwarp();
        {
          double v[NS], dx0[NS], r0[NR], aj[NR];
#pragma unroll
          for (int i = 0; i < NS; ++i) { v[i] = wb.v[i]; dx0[i] = wb.dx0[i]; }
#pragma unroll
          for (int j = 0; j < NR; ++j) {
            r0[j] = wb.r0[j];
            double a = 0.0;
#pragma unroll
            for (int i = 0; i < NS; ++i) a = fma(M.w_in[i + NIN * j], v[i] * dx0[i], a);
            aj[j] = a;
          }
#pragma unroll
          for (int tt = 0; tt < CT; ++tt) {
            const int lc = lane + 32 * tt;
            if (!isval[tt]) {
              // dJ[S, dW] v  (oracle: djac_vec); d2x = -dx^2 inside the clamp
              const int iin = d_iin[tt];
              const double xin = wb.x0[iin];
              const double vdx_in = (iin < NS) ? wb.v[iin] * wb.dx0[iin] : 0.0;
              double q2[NR];
#pragma unroll
              for (int j = 0; j < NR; ++j) {
                const double sa = sm.seed[j][lc];
                double zd = fma(sa, xin, sm.seed[NR + j][lc]);
                double ad = sa * vdx_in;
#pragma unroll
                for (int i = 0; i < NS; ++i) {
                  const double w = M.w_in[i + NIN * j];
                  const double sdx = U[tt][i] * dx0[i];
                  zd = fma(w, sdx, zd);
                  ad = fma(w, -(v[i] * dx0[i]) * sdx, ad);  // v_i * d2x_i * S_i
                }
                q2[j] = r0[j] * fma(zd, aj[j], ad);
              }
              double djo = 0.0;
              {
                // seed w_out part: dW_out[i_out, j_out] * (r0 .* a)[j_out]
                const int jo = d_jout[tt];
                double ajo = 0.0;
#pragma unroll
                for (int j = 0; j < NR; ++j) ajo = (jo == j) ? aj[j] : ajo;
                djo = d_o[tt] * (wb.r0[jo] * ajo);
              }
#pragma unroll
              for (int i = 0; i < NS; ++i) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j < NR; ++j) s = fma(M.w_out[i + NS * j], q2[j], s);
                if (d_iout[tt] == i) s += djo;
                KO[tt][i] = fma(gam, s, KO[tt][i]);
              }
              double kk[NS];
#pragma unroll
              for (int i = 0; i < NS; ++i) kk[i] = KO[tt][i];
              lu_solve<NS>(W, piv, kk);
#pragma unroll
              for (int i = 0; i < NS; ++i) KO[tt][i] = kk[i];
            } else {
#pragma unroll
              for (int i = 0; i < NS; ++i) KO[tt][i] = vk[i];
            }
          }
        }
        // post-processing per phase
        if (phase == PH_K1) {
#pragma unroll
          for (int tt = 0; tt < CT; ++tt)
#pragma unroll
            for (int i = 0; i < NS; ++i) {
              wb.K[S_K1][tt][i][lane] = KO[tt][i];
              Y[tt][i] = fma(0.5 * dt, KO[tt][i], U[tt][i]);
            }
          phase = PH_E1;
          continue;
        }
        if (phase == PH_E1) {
#pragma unroll
          for (int tt = 0; tt < CT; ++tt)
#pragma unroll
            for (int i = 0; i < NS; ++i) {
              const double k2 = KO[tt][i] + wb.K[S_K1][tt][i][lane];
              wb.K[S_K2][tt][i][lane] = k2;
              Y[tt][i] = fma(dt, k2, U[tt][i]);
            }
          phase = PH_E2;
          continue;
        }
        // PH_E2: KO = k3 -> error vector
#pragma unroll
        for (int tt = 0; tt < CT; ++tt)
#pragma unroll
          for (int i = 0; i < NS; ++i)
            KO[tt][i] = dt / 6.0 * (wb.K[S_K1][tt][i][lane] - 2.0 * wb.K[S_K2][tt][i][lane] + KO[tt][i]);
      } else if (phase == PH_F1) {
#pragma unroll
        for (int tt = 0; tt < CT; ++tt)
#pragma unroll
          for (int i = 0; i < NS; ++i) KO[tt][i] -= wb.K[f0s][tt][i][lane];
      }

      if (phase == PH_F0 || phase == PH_F1 || phase == PH_E2) {
        // ---- the single norm instance (as in k_tsit5_sens) ----
        double rsum = 0.0;
        const int npass = (phase == PH_E2) ? 2 : 1;
#pragma unroll 1
        for (int pass = 0; pass < npass; ++pass) {
          __syncwarp();
#pragma unroll
          for (int i = 0; i < NS; ++i) {
            double sa = 0.0;
#pragma unroll
            for (int tt = 0; tt < CT; ++tt) {
              const double vv = pass == 0 ? KO[tt][i] : Y[tt][i];
              sa = live[tt] ? fma(vv, vv, sa) : sa;
            }
            wb.red[i][lane] = sa;
          }
          __syncwarp();
          // RP lanes share a row: partial sums of 32/RP skewed (conflict-free) columns, then lane i adds them
          // (k_tsit5_sens does the same; one lane summing all 32 entries was 10 % of this kernel's instructions)
          constexpr int RP = NS <= 4 ? 8 : 4, SEG = 32 / RP;
          if (lane < NS * RP) {
            const int row = lane / RP, part = lane % RP;
            double ps = 0.0;
#pragma unroll
            for (int k = 0; k < SEG; ++k) ps += wb.red[row][(part * SEG + k + row) & 31];
            wb.rp[row][part] = ps;
          }
          __syncwarp();
          if (lane < NS) {
            double tot = 0.0;
#pragma unroll
            for (int pq = 0; pq < RP; ++pq) tot += wb.rp[lane][pq];
            if (pass == 0) rsum = tot; else bsum = tot;
          }
        }
        double term0 = 0.0, term1 = 0.0;
        if (lane < NS) {
          if (phase == PH_E2) {
            const double sc = fma(sqrt(fmax(asum, bsum)), my_rt, my_at);
            term0 = rsum / (sc * sc);
          } else {
            const double a = my_u0 / my_sk;
            term0 = rsum / (my_sk * my_sk);
            term1 = a * a;
          }
          wb.term[0][lane] = term0;
          wb.term[1][lane] = term1;
        }
        __syncwarp();
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int i = 0; i < NS; ++i) { s0 += wb.term[0][i]; s1 += wb.term[1][i]; }

        if (phase == PH_F0) {
          if (C::KIND == 1) { const double a = Tval / (sm.abstol[NS] + fabs(Tval) * sm.reltol[NS]); s1 = fma(a, a, s1); }
          const double d0 = sqrt(s1 / sp.norm_cnt);
          d1 = sqrt(s0 / sp.norm_cnt);
          dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
          dt0 = jmin(dt0, dtmax);
#pragma unroll
          for (int tt = 0; tt < CT; ++tt)
#pragma unroll
            for (int i = 0; i < NS; ++i) Y[tt][i] = fma(dt0, wb.K[f0s][tt][i][lane], U[tt][i]);
          phase = PH_F1;
        } else if (phase == PH_F1) {
          const double d2 = sqrt(s0 / sp.norm_cnt) / dt0;
          const double dm = jmax(d1, d2);
          const double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : lean_exp10(-(2.0 + lean_log10(dm)) * sp.inv_order);
          dt = jmin(jmin(100.0 * dt0, dt1), dtmax);
          dtnew = dt;
          // pseudo-step that saves t0: proposed state = U, f2 = f0 (commit is then a no-op)
#pragma unroll
          for (int tt = 0; tt < CT; ++tt)
#pragma unroll
            for (int i = 0; i < NS; ++i) {
              Y[tt][i] = U[tt][i];
              wb.K[4 - f0s][tt][i][lane] = wb.K[f0s][tt][i][lane];
            }
          // the eval at u0 + dt0 f0 overwrote x/dx/r: restore them from the u_n cache
          __syncwarp();
          if (lane < NS) { wb.x[lane] = wb.x0[lane]; wb.dx[lane] = wb.dx0[lane]; }
          if (lane < NR) wb.r[lane] = wb.r0[lane];
          __syncwarp();
          bsum = asum;
          phase = PH_SAVE;
        } else {
          const double EEst = sqrt(s0 / sp.norm_cnt);
          double q11;
          const double q = pi_controller<C>(sp, EEst, qold, q11);
          dt_last = dt;
          if (EEst <= 1.0) {
            ++n_acc;
            qold = jmax(EEst, 1e-4);
            dtnew = dt / (q >= sp.qs_min && q <= sp.qs_max ? 1.0 : q);  // steady-state dead-band
            tprev = t;
            t = snap_t(t + dt, tend);
            phase = PH_SAVE;
          } else {
            ++n_rej;
            dt = dt / jmin(sp.inv_qmin, q11 / sp.gamma);
            phase = PH_K1;
          }
        }
      } else if (phase == PH_SAVE) {
        // ---- saves in (tprev, t] by the order-2 dense output, loss + gradient, then commit ----
        while (isave < nsave) {
          const double tsv = __ldg(sp.saveat + isave);
          if (!(tsv <= t)) break;
          if (tsv == t) {
#pragma unroll
            for (int tt = 0; tt < CT; ++tt)
#pragma unroll
              for (int i = 0; i < NS; ++i) KO[tt][i] = Y[tt][i];
          } else {
            const double th = (tsv - tprev) / dt;
            const double c1 = th * (1.0 - th) * rb::inv_1m2d, c2 = th * (th - 2.0 * rb::d) * rb::inv_1m2d;
#pragma unroll
            for (int tt = 0; tt < CT; ++tt)
#pragma unroll
              for (int i = 0; i < NS; ++i)
                KO[tt][i] = fma(dt, fma(c2, wb.K[S_K2][tt][i][lane], c1 * wb.K[S_K1][tt][i][lane]), U[tt][i]);
          }
          if (lane == 0) {
#pragma unroll
            for (int i = 0; i < NS; ++i) wb.y[i] = KO[0][i];
          }
          __syncwarp();
          if (lane < N) {
            const int q = sm.row2obs[lane];
            double g = 0.0;
            if (q >= 0) {
              const double y = (lane < NS) ? wb.y[lane] : Tval;
              const double yc = clampd(y, sp.pred_lo, sp.pred_hi);
              const bool inside = (y >= sp.pred_lo) && (y <= sp.pred_hi);
              const size_t off = pbase + q + (size_t)sp.n_obs * isave;
              if (pred) pred[off] = yc;
              const double d = __ldg(datat + off);
              double diff;
              if (sp.loss_kind == CRNN_LOSS_MAE_SCALED) {
                const double iy = sm.inv_ys[lane];
                diff = d * iy - yc * iy;
                g = signbit(diff) ? iy : -iy;
              } else {
                diff = lean_log(clampd(d, sp.pred_lo, sp.pred_hi)) - lean_log(yc);
                g = (signbit(diff) ? 1.0 : -1.0) / yc;
              }
              loss_acc += fabs(diff);
              if (!inside) g = 0.0;
            }
            wb.g[lane] = g;
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < NS; ++i) {
            const double g = wb.g[i];
#pragma unroll
            for (int tt = 0; tt < CT; ++tt) G[tt] = fma(g, KO[tt][i], G[tt]);
          }
          ++isave;
        }
        // commit: u_n <- u_{n+1}; f0 <- f2 (swap slot roles); cache x/dx/r of the f2 evaluation
#pragma unroll
        for (int tt = 0; tt < CT; ++tt)
#pragma unroll
          for (int i = 0; i < NS; ++i) U[tt][i] = Y[tt][i];
        f0s = 4 - f0s;
        __syncwarp();
        if (lane < NS) { wb.x0[lane] = wb.x[lane]; wb.dx0[lane] = wb.dx[lane]; }
        if (lane < NR) wb.r0[lane] = wb.r[lane];
        __syncwarp();
        asum = bsum;
        dt = jmin(dtnew, dtmax);
        phase = PH_K1;
      }

      if (phase == PH_K1) {  // loopheader! + check_error! before every step attempt
        if (!(t < tend)) break;
        ++iter;
        if (dt != dt) { ret = CRNN_RET_DTNAN; break; }
        if (iter > sp.maxiters) { ret = CRNN_RET_MAXITERS; break; }
        dt = jmin(dt, dtmax);
        dt = jmin(dt, tend - t);
        if (dt <= dtmin && tend - t > dtmin) { ret = CRNN_RET_DTLESSTHANMIN; break; }
        bool bad = false;
#pragma unroll
        for (int tt = 0; tt < CT; ++tt)
#pragma unroll
          for (int i = 0; i < NS; ++i) bad |= (U[tt][i] != U[tt][i]);
        if (__any_sync(0xffffffffu, bad)) { ret = CRNN_RET_UNSTABLE; break; }
      }
    }
    if (ret == CRNN_RET_DEFAULT) ret = CRNN_RET_SUCCESS;

    const double cnt = (double)sp.n_obs * (double)isave;
    const double ltot = warp_sum(loss_acc);
    if (lane == 0) {
      loss[traj] = isave > 0 ? ltot / cnt : __longlong_as_double(0x7ff8000000000000LL);
      if (n_saved) n_saved[traj] = isave;
      if (retcode) retcode[traj] = ret;
      if (stats) {
        crnn_stats s;
        s.n_accept = n_acc; s.n_reject = n_rej; s.n_rhs = n_rhs; s.n_jac = n_jac;
        s.t_reached = t; s.dt_last = dt_last;
        stats[traj] = s;
      }
    }
#pragma unroll
    for (int tt = 0; tt < CT; ++tt) {
      const int c = lane + 32 * tt;
      if (c >= 1 && c < ncol) grad_each[(size_t)traj * np + (c - 1)] = isave > 0 ? G[tt] / cnt : 0.0;
    }
    if (pred && isave < sp.n_save) {
      for (int q = isave * sp.n_obs + lane; q < sp.n_save * sp.n_obs; q += 32) pred[pbase + q] = 0.0;
    }
    __syncwarp();
  }
}

}  // namespace crnn
