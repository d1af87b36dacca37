// kernel_tsit5_sens_pl.cuh — k_tsit5_sens with a PIPELINED value path (one warp per trajectory, WPT = 1).
// EXPERIMENT, NOT THE DEFAULT (CRNN_B200_SENS_PIPELINED=1 selects it): bit-identical results, but measured 7.50-7.67 ms against
// 6.68 ms for kernel_tsit5_sens.cuh on BASELINE configs[1] (profiles/r2_sens_v16_pipelined_rejected.txt) - the redundant value
// path costs 7.5 % more instructions and its shared-memory hand-offs more short-scoreboard stalls than the overlap wins.
//
// Same arithmetic, same phase machine and same outputs as kernel_tsit5_sens.cuh (see its header for the lane layout, the
// structured seeds and the code-size discipline).  What changes is WHEN the value path runs.  There, every stage was
//   stage sums (32 lanes) | barrier | log on NS lanes | barrier | mat-vec + exp on NR lanes | barrier | column RHS (32 lanes)
// and ncu showed the kernel bound by dependent-issue latency, not by issue slots (profiles/r2_sens_v15_rejected.txt): the
// Horner chains of log / exp ran alone on 3-6 lanes behind divergent branches while the other 26 lanes waited.  Here
//   * every lane evaluates the value path redundantly for component lane % NS / reaction lane % NR - no divergence, no
//     idle lanes, the same instruction count;
//   * the value column's stage derivatives are mirrored in a tiny per-warp array (vk), so the value inputs of stage s+1
//     (state, log, reciprocal) are computed in the SAME basic block as the column RHS of stage s, and the exponentials of
//     stage s in the same block as the columns' stage sums: two independent instruction streams the scheduler interleaves;
//   * two warp barriers per stage instead of four.
// Bitwise the same values as before: the mirrored derivative is computed with the operation order of lane 0's column.
#pragma once
#include "crnn_dev.cuh"
#include "kernel_tsit5_sens.cuh"

namespace crnn {

template <class C, int CT, bool R1>
struct alignas(16) SensSmemP {
  double seed[R1 ? 2 * C::NR : C::NW][32 * CT];  // dW/dp (dense, or the a_j / b_j rows); column 0 = value = 0
  double w_in[C::NIN * C::NR];
  double w_out[C::NS * C::NR];                   // rows addressed by lane % NS in the value path
  double w_b[C::NR];
  double inv_ys[C::N];
  double abstol[C::N], reltol[C::N];
  int row2obs[C::N];
};

template <class C, int CT>
struct alignas(16) WarpBufP {
  double K[7][CT][C::NS][32];
  double red[C::NS][32];      // column-sum scratch of the dual-aware norms
  double vk[7][C::NS];        // the value column's stage derivatives, slot-indexed like K
  double vu[C::N];            // value state u_n
  double vx[2][C::N];         // double-buffered value inputs x (vx[.][NS] = -1/(R T) for F1) ...
  double vdx[2][C::N];        // ... and d x / d u
  double vr[C::NR];           // r of the current phase
  double y[C::N];             // save phase: interpolated value state
  double g[C::N];
  double term[2][C::N];
  double rp[C::NS][8];        // partial row sums (RP lanes share one row of `red`)
  double cold[4];             // rarely-read scalars kept out of registers: tend, dtmax, dtmin, dt of the last attempt
};

template <class C, int CT, int WARPS, int MINB, bool R1>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_tsit5_sens_pl(const __grid_constant__ ModelP<C> mp, const __grid_constant__ SolveP<C> sp,
                const double* __restrict__ seed_dev, const R1Desc* __restrict__ desc_dev, int ncol,
                const double* __restrict__ u0, const int* __restrict__ n_save_used, long long ntraj,
                const double* __restrict__ data, double* __restrict__ loss, double* __restrict__ grad_each,
                double* __restrict__ pred, int* __restrict__ n_saved, int* __restrict__ retcode,
                crnn_stats* __restrict__ stats, unsigned long long* __restrict__ queue,
                const long long* __restrict__ in_idx) {
  constexpr int NS = C::NS, NR = C::NR, N = C::N, NIN = C::NIN, NW = C::NW;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SensSmemP<C, CT, R1>& sm = *reinterpret_cast<SensSmemP<C, CT, R1>*>(smem_raw);
  WarpBufP<C, CT>* wbs = reinterpret_cast<WarpBufP<C, CT>*>(smem_raw + sizeof(SensSmemP<C, CT, R1>));
  unsigned lane_u, tid_u;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane_u));
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid_u));
  const int lane = (int)lane_u, warp = (int)(tid_u >> 5);
  const int i0 = lane % NS, j0 = lane % NR;   // the value-path component / reaction this lane evaluates (redundantly)
  WarpBufP<C, CT>& wb = wbs[warp];
  auto kload = [&](int slot, int tt, double (&v)[NS]) {
    const double* base = &wb.K[slot][tt][0][0];
#pragma unroll
    for (int p = 0; p < NS / 2; ++p) {
      const double2 w = reinterpret_cast<const double2*>(base)[p * 32 + lane];
      v[2 * p] = w.x; v[2 * p + 1] = w.y;
    }
    if (NS & 1) v[NS - 1] = base[(NS - 1) * 32 + lane];
  };
  auto kstore = [&](int slot, int tt, const double (&v)[NS]) {
    double* base = &wb.K[slot][tt][0][0];
#pragma unroll
    for (int p = 0; p < NS / 2; ++p) reinterpret_cast<double2*>(base)[p * 32 + lane] = make_double2(v[2 * p], v[2 * p + 1]);
    if (NS & 1) base[(NS - 1) * 32 + lane] = v[NS - 1];
  };

  for (int q = threadIdx.x; q < (R1 ? 2 * NR : NW) * 32 * CT; q += blockDim.x) (&sm.seed[0][0])[q] = seed_dev[q];
  for (int q = threadIdx.x; q < NIN * NR; q += blockDim.x) sm.w_in[q] = mp.w_in[q];
  for (int q = threadIdx.x; q < NS * NR; q += blockDim.x) sm.w_out[q] = mp.w_out[q];
  for (int q = threadIdx.x; q < NR; q += blockDim.x) sm.w_b[q] = mp.w_b[q];
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
    d_o[t] = 0.0; d_iin[t] = d_iout[t] = d_jout[t] = 0;
    if (R1) {
      const R1Desc d = desc_dev[lane + 32 * t];
      d_o[t] = d.o; d_iin[t] = d.i_in; d_iout[t] = d.i_out; d_jout[t] = d.j_out;
    }
    live[t] = (lane + 32 * t) < ncol && (sp.incl_sens || isval[t]);
  }

  constexpr int PH_F0 = 0, PH_F1 = 7, PH_SAVE = 8;  // phases 1..6 are the Tsit5 stages

  // value inputs of one component: x = log(clamp y), dx = 1/clamp y inside the clamp
  auto value_inputs = [&](double yi, double& xi, double& dxi) {
    const double uc = clampd(yi, mp.lb, mp.ub);
    const bool inside = (yi >= mp.lb) && (yi <= mp.ub);
    xi = lean_log_cf(uc);
    dxi = inside ? __drcp_rn(uc) : 0.0;
  };

  while (true) {
    unsigned long long tq = 0;
    if (lane == 0) tq = atomicAdd(queue, 1ull);
    const long long traj = (long long)__shfl_sync(0xffffffffu, tq, 0);
    if (traj >= ntraj) break;
    const double* __restrict__ u0t = u0 + (in_idx ? __ldg(in_idx + traj) : traj) * N;

    double U[CT][NS], Y[CT][NS], KO[CT][NS];
    double xT = 0.0;
    if (C::KIND == 1) xT = -1.0 / (mp.gas_R * __ldg(u0t + NS));
    double mybT = sm.w_b[j0];
    if (C::KIND == 1) mybT = fma(sm.w_in[NS + NIN * j0], xT, mybT);
    if (lane < N) wb.vu[lane] = __ldg(u0t + lane);
    if (lane < NS) {
#pragma unroll
      for (int j = 0; j < 7; ++j) wb.vk[j][lane] = 0.0;   // finite: the straight-line look-ahead multiplies stale slots by 0
    }
    if (C::KIND == 1 && lane == 0) { wb.vx[0][NS] = xT; wb.vx[1][NS] = xT; }
#pragma unroll
    for (int t = 0; t < CT; ++t)
#pragma unroll
      for (int i = 0; i < NS; ++i) U[t][i] = isval[t] ? __ldg(u0t + i) : 0.0;

    int nsave = sp.n_save;
    const double t0 = sp.t0;
    {
      double tend = sp.t1;
      if (n_save_used) {
        int q = __ldg(n_save_used + traj);
        if (q > 0 && q <= sp.n_save) { nsave = q; tend = __ldg(sp.saveat + q - 1); }
      }
      if (lane == 0) { wb.cold[0] = tend; wb.cold[1] = tend - t0; wb.cold[2] = fmax(ulp_of(t0), ulp_of(tend)); wb.cold[3] = 0.0; }
      __syncwarp();
    }
    const size_t pbase = (size_t)traj * sp.n_obs * sp.n_save;
    const double* __restrict__ datat = data + (size_t)(in_idx ? __ldg(in_idx + traj) : traj) * sp.n_obs * sp.n_save;

    int n_acc = 0, n_rej = 0;
    double G[CT], loss_acc = 0.0;
#pragma unroll
    for (int t = 0; t < CT; ++t) G[t] = 0.0;
    double asum = 0.0, bsum = 0.0;  // lane i < NS: dual magnitude^2 of u_i at t_n / t_{n+1}
    if (lane < NS) { const double v = __ldg(u0t + lane); asum = v * v; }
    double t = t0, tprev = t0, dt = 0.0, dtnew = 0.0, lqold = lean_log(1e-4);  // log(qoldinit)
    int isave = 0, ret = CRNN_RET_DEFAULT, phase = PH_F0, k1s = 0;  // k1s: slot of K1 (0 or 6), K7 in 6-k1s
    int cur = 0;          // which half of vx / vdx holds this phase's value inputs
    bool ahead = false;   // true: they were computed during the previous phase
    const int my_q = (lane < N) ? sm.row2obs[lane] : -1;
    double ts_next = __ldg(sp.saveat), d_next = 0.0;
    if (my_q >= 0) d_next = __ldg(datat + my_q);

    while (true) {
      if (phase != PH_SAVE) {
        const int nj = (phase == PH_F1) ? 1 : phase;
        const double h = dt;
        // ---- value inputs of this phase (state, log, reciprocal) unless the previous phase already computed them ----
        if (!ahead) {
          double acc = 0.0;
          if (nj > 0) {
            acc = c_tsA[phase][0] * wb.vk[k1s][i0];
#pragma unroll 2
            for (int j = 1; j < nj; ++j) acc = fma(c_tsA[phase][j], wb.vk[j][i0], acc);
          }
          double xi, dxi;
          value_inputs(fma(h, acc, wb.vu[i0]), xi, dxi);
          wb.vx[cur][i0] = xi; wb.vdx[cur][i0] = dxi;
          __syncwarp();
        }
        // ---- block 1: r_j0 = exp(W_in' x + b)  ||  stage state of the columns, Y = U + h * sum_{j<nj} A[phase][j] K_j ----
        double rj;
        {
          double z = mybT;
#pragma unroll
          for (int i = 0; i < NS; ++i) z = fma(sm.w_in[i + NIN * j0], wb.vx[cur][i], z);
          rj = lean_exp_cf(z);
        }
        {
          double kv[NS];
          if (nj > 0) {  // j = 0 reads K1 from its FSAL slot; the rest are slots 1..nj-1
            const double a0 = c_tsA[phase][0];
#pragma unroll
            for (int tt = 0; tt < CT; ++tt) {
              kload(k1s, tt, kv);
#pragma unroll
              for (int i = 0; i < NS; ++i) KO[tt][i] = a0 * kv[i];
            }
#pragma unroll 2
            for (int j = 1; j < nj; ++j) {
              const double a = c_tsA[phase][j];
#pragma unroll
              for (int tt = 0; tt < CT; ++tt) {
                kload(j, tt, kv);
#pragma unroll
                for (int i = 0; i < NS; ++i) KO[tt][i] = fma(a, kv[i], KO[tt][i]);
              }
            }
          } else {
#pragma unroll
            for (int tt = 0; tt < CT; ++tt)
#pragma unroll
              for (int i = 0; i < NS; ++i) KO[tt][i] = 0.0;
          }
#pragma unroll
          for (int tt = 0; tt < CT; ++tt)
#pragma unroll
            for (int i = 0; i < NS; ++i) Y[tt][i] = fma(h, KO[tt][i], U[tt][i]);
        }
        wb.vr[j0] = rj;
        __syncwarp();

        // ---- block 2: KO = f(Y) on this lane's columns  ||  the value column's K and the value inputs of the next stage ----
        const bool look = (phase >= 1 && phase <= 5);
        double kval, xn = 0.0, dxn = 0.0;
        {
          double dx[NS], r[NR];
#pragma unroll
          for (int i = 0; i < NS; ++i) dx[i] = wb.vdx[cur][i];
#pragma unroll
          for (int j = 0; j < NR; ++j) r[j] = wb.vr[j];
          // value column's derivative for component i0, in the operation order of lane 0's column below (bitwise equal)
          kval = 0.0;
#pragma unroll
          for (int j = 0; j < NR; ++j) kval = fma(sm.w_out[i0 + NS * j], r[j], kval);
          if (look) {
            // straight-line: rows of c_tsA are zero beyond their stage and vk slots are always finite, so the terms j >= phase
            // add exact zeros; the newest derivative (slot `phase`, not stored yet) enters from the register
            double acc = c_tsA[phase + 1][0] * wb.vk[k1s][i0];
#pragma unroll
            for (int j = 1; j < 6; ++j) acc = fma(c_tsA[phase + 1][j], j == phase ? kval : wb.vk[j][i0], acc);
            value_inputs(fma(h, acc, wb.vu[i0]), xn, dxn);
          }
#pragma unroll
          for (int tt = 0; tt < CT; ++tt) {
            const int lc = lane + 32 * tt;
            double sd[NS], q[NR];
#pragma unroll
            for (int i = 0; i < NS; ++i) sd[i] = Y[tt][i] * dx[i];
            const double xin = R1 ? wb.vx[cur][d_iin[tt]] : 0.0;
#pragma unroll
            for (int j = 0; j < NR; ++j) {
              double zd = R1 ? fma(sm.seed[j][lc], xin, sm.seed[NR + j][lc]) : sm.seed[NIN * NR + j][lc];
#pragma unroll
              for (int i = 0; i < NS; ++i) zd = fma(mp.w_in[i + NIN * j], sd[i], zd);
              if (!R1) {
#pragma unroll
                for (int i = 0; i < NS; ++i) zd = fma(sm.seed[i + NIN * j][lc], wb.vx[cur][i], zd);
                if (C::KIND == 1) zd = fma(sm.seed[NS + NIN * j][lc], xT, zd);
              }
              if (isval[tt]) zd = 1.0;
              q[j] = r[j] * zd;
            }
            const double rov = R1 ? d_o[tt] * wb.vr[d_jout[tt]] : 0.0;
#pragma unroll
            for (int i = 0; i < NS; ++i) {
              double s = 0.0;
#pragma unroll
              for (int j = 0; j < NR; ++j) s = fma(mp.w_out[i + NS * j], q[j], s);
              if (!R1) {
#pragma unroll
                for (int j = 0; j < NR; ++j) s = fma(sm.seed[NIN * NR + NR + i + NS * j][lc], r[j], s);
              } else if (d_iout[tt] == i) {
                s += rov;
              }
              KO[tt][i] = s;
            }
          }
        }
        // ---- store KO into its stage slot (F0 -> K1 ; F1 -> slot 1, scratch ; stage s -> K_{s+1}), mirror the value column ----
        {
          const int dst = (phase == PH_F0) ? k1s : (phase == PH_F1) ? 1 : (phase == 6 ? 6 - k1s : phase);
#pragma unroll
          for (int tt = 0; tt < CT; ++tt) kstore(dst, tt, KO[tt]);
          wb.vk[dst][i0] = kval;
        }
        if (look) { wb.vx[cur ^ 1][i0] = xn; wb.vdx[cur ^ 1][i0] = dxn; }
        __syncwarp();

        if (look) {
          ++phase; cur ^= 1; ahead = true;
        } else {
          ahead = false;
          // ---- phases that need a norm: F0 (|f0|), F1 (|f1 - f0|), stage 6 (error estimate) ----
          if (phase == PH_F1) {
#pragma unroll
            for (int tt = 0; tt < CT; ++tt) {
              double kv[NS];
              kload(k1s, tt, kv);
#pragma unroll
              for (int i = 0; i < NS; ++i) KO[tt][i] -= kv[i];
            }
          } else if (phase == 6) {
#pragma unroll
            for (int tt = 0; tt < CT; ++tt)
#pragma unroll
              for (int i = 0; i < NS; ++i) KO[tt][i] *= c_tsBT[6];
#pragma unroll 1
            for (int j = 0; j < 6; ++j) {
              const double a = c_tsBT[j];
              const int slot = (j == 0) ? k1s : j;
#pragma unroll
              for (int tt = 0; tt < CT; ++tt) {
                double kv[NS];
                kload(slot, tt, kv);
#pragma unroll
                for (int i = 0; i < NS; ++i) KO[tt][i] = fma(a, kv[i], KO[tt][i]);
              }
            }
#pragma unroll
            for (int tt = 0; tt < CT; ++tt)
#pragma unroll
              for (int i = 0; i < NS; ++i) KO[tt][i] *= dt;
          }
          double rsum = 0.0;
          const int npass = (phase == 6) ? 2 : 1;
#pragma unroll 1
          for (int pass = 0; pass < npass; ++pass) {
            __syncwarp();
#pragma unroll
            for (int i = 0; i < NS; ++i) {
              double sa = 0.0;
#pragma unroll
              for (int tt = 0; tt < CT; ++tt) {
                const double v = pass == 0 ? KO[tt][i] : Y[tt][i];
                sa = live[tt] ? fma(v, v, sa) : sa;
              }
              wb.red[i][lane] = sa;
            }
            __syncwarp();
            constexpr int RP = NS <= 4 ? 8 : (NS <= 8 ? 4 : (NS <= 16 ? 2 : 1)), SEG = 32 / RP;
            if (lane < NS * RP) {
              const int row = lane / RP, part = lane % RP;
              double ps = 0.0;
#pragma unroll
              for (int k = 0; k < SEG; ++k) ps += wb.red[row][(part * SEG + k + row) & 31];
              wb.rp[row][part] = ps;
            }
            __syncwarp();
            double tot = 0.0;
            if (lane < NS) {
#pragma unroll
              for (int p = 0; p < RP; ++p) tot += wb.rp[lane][p];
            }
            if (lane < NS) { if (pass == 0) rsum = tot; else bsum = tot; }
          }
          double term0 = 0.0, term1 = 0.0;
          if (lane < NS) {
            if (phase == 6) {
              const double sc = fma(sqrt(fmax(asum, bsum)), sm.reltol[lane], sm.abstol[lane]);
              term0 = rsum / (sc * sc);
            } else {
              const double my_u0 = __ldg(u0t + lane), my_sk = sm.abstol[lane] + fabs(my_u0) * sm.reltol[lane];
              const double a = my_u0 / my_sk;
              term0 = rsum / (my_sk * my_sk);
              term1 = a * a;
            }
            wb.term[0][lane] = term0; wb.term[1][lane] = term1;
          }
          __syncwarp();
          double s0 = 0.0, s1 = 0.0;
#pragma unroll
          for (int i = 0; i < NS; ++i) { s0 += wb.term[0][i]; s1 += wb.term[1][i]; }

          if (phase == PH_F0) {
            if (C::KIND == 1) {
              const double Tval = __ldg(u0t + NS), a = Tval / (sm.abstol[NS] + fabs(Tval) * sm.reltol[NS]);
              s1 = fma(a, a, s1);
            }
            const double d0 = sqrt(s1 / sp.norm_cnt);
            const double d1 = sqrt(s0 / sp.norm_cnt);
            const double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
            dt = jmin(dt0, wb.cold[1]);
            dtnew = d1;
            phase = PH_F1;
          } else if (phase == PH_F1) {
            const double dt0 = dt, d1 = dtnew;
            const double d2 = sqrt(s0 / sp.norm_cnt) / dt0;
            const double dm = jmax(d1, d2);
            const double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : lean_exp10(-(2.0 + lean_log10(dm)) * sp.inv_order);
            dt = jmin(jmin(100.0 * dt0, dt1), wb.cold[1]);
            dtnew = dt;
            // pseudo-step: proposed state = U, K7 = K1, so the commit below is a no-op
#pragma unroll
            for (int tt = 0; tt < CT; ++tt) {
              double kv[NS];
              kload(k1s, tt, kv);
              kstore(6 - k1s, tt, kv);
#pragma unroll
              for (int i = 0; i < NS; ++i) Y[tt][i] = U[tt][i];
            }
            wb.vk[6 - k1s][i0] = wb.vk[k1s][i0];
            bsum = asum;
            phase = PH_SAVE;
          } else {
            const double EEst = sqrt(s0 / sp.norm_cnt);
            double q11 = 0.0, q = sp.inv_qmax, lE = 0.0;
            if (EEst != 0.0) {
              lE = lean_log(EEst);
              q11 = lean_exp(LM_MUL(sp.beta1, lE));
              q = jmax(sp.inv_qmax, jmin(sp.inv_qmin, q11 / lean_exp(LM_MUL(sp.beta2, lqold)) / sp.gamma));
            }
            if (isval[0]) wb.cold[3] = dt;
            if (EEst <= 1.0) {
              ++n_acc;
              lqold = (EEst > 1e-4) ? lE : lean_log(1e-4);
              dtnew = dt / q;
              tprev = t;
              t = snap_t(t + dt, wb.cold[0]);
              phase = PH_SAVE;
            } else {
              ++n_rej;
              dt = dt / jmin(sp.inv_qmin, q11 / sp.gamma);
              phase = 1;
            }
          }
        }
      } else {
        // ---- SAVE phase: every save time in (tprev, t] via the dense interpolant, loss and gradient fused, then commit ----
        while (isave < nsave) {
          const double tsv = ts_next;
          if (!(tsv <= t)) break;
          if (tsv == t) {
#pragma unroll
            for (int tt = 0; tt < CT; ++tt)
#pragma unroll
              for (int i = 0; i < NS; ++i) KO[tt][i] = Y[tt][i];
          } else {
            double b[7];
            ts::dense_b((tsv - tprev) / dt, b);
#pragma unroll
            for (int tt = 0; tt < CT; ++tt) {
              double kv[NS];
              kload(k1s, tt, kv);
#pragma unroll
              for (int i = 0; i < NS; ++i) KO[tt][i] = b[0] * kv[i];
#pragma unroll
              for (int j = 1; j < 7; ++j) {
                kload(j == 6 ? 6 - k1s : j, tt, kv);
#pragma unroll
                for (int i = 0; i < NS; ++i) KO[tt][i] = fma(b[j], kv[i], KO[tt][i]);
              }
#pragma unroll
              for (int i = 0; i < NS; ++i) KO[tt][i] = fma(dt, KO[tt][i], U[tt][i]);
            }
          }
          if (isval[0]) {
#pragma unroll
            for (int i = 0; i < NS; ++i) wb.y[i] = KO[0][i];
          }
          __syncwarp();
          if (lane < N) {
            const int q = my_q;
            double g = 0.0;
            if (q >= 0) {
              const double y = (lane < NS) ? wb.y[lane] : __ldg(u0t + NS);  // the T row never changes
              const double yc = clampd(y, sp.pred_lo, sp.pred_hi);
              const bool inside = (y >= sp.pred_lo) && (y <= sp.pred_hi);
              const size_t off = pbase + q + (size_t)sp.n_obs * isave;
              if (pred) pred[off] = yc;
              const double d = d_next;
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
          if (isave < nsave) {
            ts_next = __ldg(sp.saveat + isave);
            if (my_q >= 0) d_next = __ldg(datat + my_q + (size_t)sp.n_obs * isave);
          }
        }
        // commit: u_n <- u_{n+1}, K1 <- K7 (FSAL: swap the slot roles, no copy), value state mirrored for the look-ahead
#pragma unroll
        for (int tt = 0; tt < CT; ++tt)
#pragma unroll
          for (int i = 0; i < NS; ++i) U[tt][i] = Y[tt][i];
        if (isval[0]) {
#pragma unroll
          for (int i = 0; i < NS; ++i) wb.vu[i] = Y[0][i];
        }
        __syncwarp();
        k1s = 6 - k1s;
        asum = bsum;
        dt = jmin(dtnew, wb.cold[1]);
        phase = 1;
        ahead = false;
      }

      if (phase == 1 && !ahead) {  // loopheader! + check_error! before every step attempt
        const double tend = wb.cold[0], dtmin = wb.cold[2];
        if (!(t < tend)) break;
        if (dt != dt) { ret = CRNN_RET_DTNAN; break; }
        if ((long long)n_acc + n_rej + 1 > sp.maxiters) { ret = CRNN_RET_MAXITERS; break; }
        dt = jmin(dt, wb.cold[1]);
        dt = jmin(dt, tend - t);
        if (dt <= dtmin && tend - t > dtmin) { ret = CRNN_RET_DTLESSTHANMIN; break; }
        bool bad = false;
#pragma unroll
        for (int tt = 0; tt < CT; ++tt)
#pragma unroll
          for (int i = 0; i < NS; ++i) bad |= (U[tt][i] != U[tt][i]);
        bad = __any_sync(0xffffffffu, bad);
        if (bad) { ret = CRNN_RET_UNSTABLE; break; }
      }
    }
    if (ret == CRNN_RET_DEFAULT) ret = CRNN_RET_SUCCESS;

    // ---- per-trajectory outputs ----
    const double cnt = (double)sp.n_obs * (double)isave;
    const double ltot = warp_sum(loss_acc);
    if (isval[0]) {
      loss[traj] = isave > 0 ? ltot / cnt : __longlong_as_double(0x7ff8000000000000LL);
      if (n_saved) n_saved[traj] = isave;
      if (retcode) retcode[traj] = ret;
      if (stats) {
        crnn_stats s;
        s.n_accept = n_acc; s.n_reject = n_rej; s.n_rhs = 2 + 6 * (n_acc + n_rej); s.n_jac = 0;
        s.t_reached = t; s.dt_last = wb.cold[3];
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
