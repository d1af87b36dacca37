// kernel_tsit5_sens.cuh — loss + forward-mode gradient for a batch: one WARP owns one trajectory.
//
// Replaces `ForwardDiff.gradient(x -> loss_neuralode(x, i_exp), p)` (case2/case2.jl:132-137,195;
// Zygote.forwarddiff in case1/case1.jl:195-199): duals pushed through the adaptive solver.
// Lane l of the warp owns dual column l (+32 per column tile): column 0 is the value, column
// c >= 1 the partial d/dp_c.  All columns share one step sequence; the partials take part in
// the error norm exactly like DiffEqBase's norm over Dual arrays.
//
// Per stage the value path (NS logs, NR exps) is spread over lanes — lane i takes log(u_i),
// lane j takes exp(z_j) — and broadcast through a few bytes of per-warp shared memory, so the
// warp issues ONE log and ONE exp per stage instead of NS+NR; every lane then applies
// J(u)*S + (df/dW)*dW/dp_c matrix-free to its own column.
// The loss (MAE-scaled or MAE-log) and its gradient are fused at each save point, so the
// n_state x n_save x np sensitivity tensor never leaves registers.
// A persistent grid pulls trajectory indices from a global atomic queue (step counts vary ~3x).
#pragma once
#include "crnn_dev.cuh"

namespace crnn {

// Tsit5 coefficients addressed by a RUNTIME stage index (uniform constant-bank loads).
// row s (1..6): a_{s+1,1..s}; row 0: {} ; row 7: {1} (Euler probe of the initial-step heuristic)
__constant__ double c_tsA[8][6] = {
    {0, 0, 0, 0, 0, 0},
    {ts::a21, 0, 0, 0, 0, 0},
    {ts::a31, ts::a32, 0, 0, 0, 0},
    {ts::a41, ts::a42, ts::a43, 0, 0, 0},
    {ts::a51, ts::a52, ts::a53, ts::a54, 0, 0},
    {ts::a61, ts::a62, ts::a63, ts::a64, ts::a65, 0},
    {ts::a71, ts::a72, ts::a73, ts::a74, ts::a75, ts::a76},
    {1.0, 0, 0, 0, 0, 0}};
__constant__ double c_tsBT[7] = {ts::bt1, ts::bt2, ts::bt3, ts::bt4, ts::bt5, ts::bt6, ts::bt7};

template <class C, int CT>
struct alignas(16) SensSmem {
  double seed[C::NW][32 * CT];  // dW/dp, zero padded; column 0 (value lane) is zero
  double w_in[C::NIN * C::NR];
  double w_b[C::NR];
  double yscale[C::N];
  double abstol[C::N], reltol[C::N];
  double dense_r[7][4];         // Tsit5 dense-output polynomial coefficients (lane j takes b_j)
  int row2obs[C::N];
};

// Per-warp working set.  K holds the seven stage derivatives of every column: lane l owns
// K[slot][tile][i][l], so all accesses are conflict-free 256 B rows.
template <class C, int CT>
struct alignas(16) WarpBuf {
  double K[7][CT][C::NS][32];
  double red[C::NS][32];  // column-sum scratch for the dual-aware norms
  double y[C::N];
  double x[C::N];
  double dx[C::N];
  double r[C::NR];
  double g[C::N];
  double b[8];            // dense-output weights b_j(theta)
  double term[C::N];
};

template <class C, int CT, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_tsit5_sens(const __grid_constant__ ModelP<C> mp, const __grid_constant__ SolveP<C> sp,
             const double* __restrict__ seed_dev, int ncol,
             const double* __restrict__ u0, const int* __restrict__ n_save_used, long long ntraj,
             const double* __restrict__ data, double* __restrict__ loss, double* __restrict__ grad_each,
             double* __restrict__ pred, int* __restrict__ n_saved, int* __restrict__ retcode,
             crnn_stats* __restrict__ stats, unsigned long long* __restrict__ queue) {
  constexpr int NS = C::NS, NR = C::NR, N = C::N, NIN = C::NIN, NW = C::NW;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SensSmem<C, CT>& sm = *reinterpret_cast<SensSmem<C, CT>*>(smem_raw);
  WarpBuf<C, CT>* wbs = reinterpret_cast<WarpBuf<C, CT>*>(smem_raw + sizeof(SensSmem<C, CT>));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WarpBuf<C, CT>& wb = wbs[warp];

  for (int q = threadIdx.x; q < NW * 32 * CT; q += blockDim.x) (&sm.seed[0][0])[q] = seed_dev[q];
  for (int q = threadIdx.x; q < NIN * NR; q += blockDim.x) sm.w_in[q] = mp.w_in[q];
  for (int q = threadIdx.x; q < NR; q += blockDim.x) sm.w_b[q] = mp.w_b[q];
  for (int q = threadIdx.x; q < N; q += blockDim.x) {
    sm.yscale[q] = 1.0 / sp.inv_yscale[q];
    sm.row2obs[q] = sp.row2obs[q];
    sm.abstol[q] = sp.abstol[q];
    sm.reltol[q] = sp.reltol[q];
  }
  if (threadIdx.x == 0) {
    const double rr[7][4] = {{ts::r11, ts::r12, ts::r13, ts::r14}, {0.0, ts::r22, ts::r23, ts::r24},
                             {0.0, ts::r32, ts::r33, ts::r34}, {0.0, ts::r42, ts::r43, ts::r44},
                             {0.0, ts::r52, ts::r53, ts::r54}, {0.0, ts::r62, ts::r63, ts::r64},
                             {0.0, ts::r72, ts::r73, ts::r74}};
    for (int j = 0; j < 7; ++j)
      for (int k = 0; k < 4; ++k) sm.dense_r[j][k] = rr[j][k];
  }
  __syncthreads();
  const int np = ncol - 1;

  // columns owned by this lane; live = takes part in norms
  bool isval[CT], live[CT];
#pragma unroll
  for (int t = 0; t < CT; ++t) {
    isval[t] = (t == 0 && lane == 0);
    live[t] = (lane + 32 * t) < ncol && (sp.incl_sens || isval[t]);
  }
  double my_at = 0.0, my_rt = 0.0;
  if (lane < NS) { my_at = sm.abstol[lane]; my_rt = sm.reltol[lane]; }

  // Phase machine (one warp-uniform `phase` per trajectory).  Every RHS evaluation — the two
  // of the initial-step heuristic and the six Tsit5 stages — goes through ONE instance of the
  // warp-cooperative RHS inside a runtime stage loop, the stage vectors live in shared memory
  // and the tableau in the constant bank, so the hot loop is a few KB of code (the fully
  // unrolled first version was 160 KB of SASS and stalled on instruction fetch: profiles/).
  constexpr int PH_F0 = 0, PH_F1 = 7, PH_SAVE = 8;   // phases 1..6 are the Tsit5 stages

  while (true) {
    unsigned long long tq = 0;
    if (lane == 0) tq = atomicAdd(queue, 1ull);
    const long long traj = (long long)__shfl_sync(0xffffffffu, tq, 0);
    if (traj >= ntraj) break;

    // U: state; Y: stage state (holds the proposed u_{n+1} from stage 6 through the save phase);
    // KO: RHS output / scratch
    double U[CT][NS], Y[CT][NS], KO[CT][NS];
    double Tval = 0.0, xT = 0.0, mybT = 0.0, my_sk = 1.0, my_u0 = 0.0;
    if (C::KIND == 1) { Tval = __ldg(u0 + traj * N + NS); xT = -1.0 / (mp.gas_R * Tval); }
    if (lane < NR) {
      mybT = sm.w_b[lane];
      if (C::KIND == 1) mybT = fma(sm.w_in[NS + NIN * lane], xT, mybT);
    }
#pragma unroll
    for (int t = 0; t < CT; ++t)
#pragma unroll
      for (int i = 0; i < NS; ++i) U[t][i] = isval[t] ? __ldg(u0 + traj * N + i) : 0.0;
    if (lane < NS) {
      my_u0 = __ldg(u0 + traj * N + lane);
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

    // lane i < NS returns sum over participating columns of V[.][i]^2 (others: garbage-free 0)
    auto colsq = [&](const double (&V)[CT][NS]) -> double {
      __syncwarp();
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        double s = 0.0;
#pragma unroll
        for (int t = 0; t < CT; ++t) s = live[t] ? fma(V[t][i], V[t][i], s) : s;
        wb.red[i][lane] = s;
      }
      __syncwarp();
      double tot = 0.0;
      if (lane < NS) {
#pragma unroll 8
        for (int k = 0; k < 32; ++k) tot += wb.red[lane][(k + lane) & 31];  // skewed: conflict-free
      }
      return tot;
    };
    // all lanes: sum_i term_i, term held by lane i < NS
    auto sum_terms = [&](double term) -> double {
      __syncwarp();
      if (lane < NS) wb.term[lane] = term;
      __syncwarp();
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < NS; ++i) s += wb.term[i];
      return s;
    };

    int n_rhs = 0, n_acc = 0, n_rej = 0;
    double G[CT], loss_acc = 0.0;
#pragma unroll
    for (int t = 0; t < CT; ++t) G[t] = 0.0;
    double asum = my_u0 * my_u0, bsum = 0.0;  // lane i: dual magnitude^2 of u_i at t_n / t_{n+1}
    double t = t0, tprev = t0, dt = 0.0, dt0 = 0.0, d1 = 0.0, dtnew = 0.0, qold = 1e-4, dt_last = 0.0;
    int isave = 0, ret = CRNN_RET_DEFAULT, phase = PH_F0, k1s = 0;  // k1s: slot of K1 (0 or 6), K7 in 6-k1s
    long long iter = 0;

    while (true) {
      if (phase != PH_SAVE) {
        // ---- stage state Y = U + h * sum_{j<nj} A[phase][j] K_j ----
        {
          const int nj = (phase == PH_F1) ? 1 : phase;
          const double h = (phase == PH_F1) ? dt0 : dt;
#pragma unroll
          for (int tt = 0; tt < CT; ++tt)
#pragma unroll
            for (int i = 0; i < NS; ++i) KO[tt][i] = 0.0;
          for (int j = 0; j < nj; ++j) {
            const double a = c_tsA[phase][j];
            const int slot = (j == 0) ? k1s : j;
#pragma unroll
            for (int tt = 0; tt < CT; ++tt)
#pragma unroll
              for (int i = 0; i < NS; ++i) KO[tt][i] = fma(a, wb.K[slot][tt][i][lane], KO[tt][i]);
          }
#pragma unroll
          for (int tt = 0; tt < CT; ++tt)
#pragma unroll
            for (int i = 0; i < NS; ++i) Y[tt][i] = fma(h, KO[tt][i], U[tt][i]);
        }

        // ---- KO = f(Y) on all columns: the single RHS instance (3 warp barriers) ----
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < NS; ++i) wb.y[i] = Y[0][i];
        }
        __syncwarp();
        if (lane < NS) {
          const double yi = wb.y[lane];
          const double uc = clampd(yi, mp.lb, mp.ub);
          const bool inside = (yi >= mp.lb) && (yi <= mp.ub);
          wb.x[lane] = log(uc);
          wb.dx[lane] = inside ? 1.0 / uc : 0.0;
        }
        __syncwarp();
        if (lane < NR) {
          double z = mybT;
#pragma unroll
          for (int i = 0; i < NS; ++i) z = fma(sm.w_in[i + NIN * lane], wb.x[i], z);
          wb.r[lane] = exp(z);
        }
        __syncwarp();
        {
          double x[NIN], dx[NS], r[NR];
#pragma unroll
          for (int i = 0; i < NS; ++i) { x[i] = wb.x[i]; dx[i] = wb.dx[i]; }
          if (C::KIND == 1) x[NS] = xT;
#pragma unroll
          for (int j = 0; j < NR; ++j) r[j] = wb.r[j];
#pragma unroll
          for (int tt = 0; tt < CT; ++tt) {
            const int lc = lane + 32 * tt;
            double sd[NS], q[NR];
#pragma unroll
            for (int i = 0; i < NS; ++i) sd[i] = Y[tt][i] * dx[i];
#pragma unroll
            for (int j = 0; j < NR; ++j) {
              double zd = sm.seed[NIN * NR + j][lc];
#pragma unroll
              for (int i = 0; i < NS; ++i) zd = fma(mp.w_in[i + NIN * j], sd[i], zd);
#pragma unroll
              for (int i = 0; i < NIN; ++i) zd = fma(sm.seed[i + NIN * j][lc], x[i], zd);
              if (isval[tt]) zd = 1.0;
              q[j] = r[j] * zd;
            }
#pragma unroll
            for (int i = 0; i < NS; ++i) {
              double s = 0.0;
#pragma unroll
              for (int j = 0; j < NR; ++j) s = fma(mp.w_out[i + NS * j], q[j], s);
#pragma unroll
              for (int j = 0; j < NR; ++j) s = fma(sm.seed[NIN * NR + NR + i + NS * j][lc], r[j], s);
              KO[tt][i] = s;
            }
          }
        }
        ++n_rhs;

        // ---- store KO into its stage slot ----
        {
          // F0 -> K1 ; F1 -> slot 1 (scratch, free until stage 1 writes K2) ; stage s -> K_{s+1}
          const int dst = (phase == PH_F0) ? k1s : (phase == PH_F1) ? 1 : (phase == 6 ? 6 - k1s : phase);
#pragma unroll
          for (int tt = 0; tt < CT; ++tt)
#pragma unroll
            for (int i = 0; i < NS; ++i) wb.K[dst][tt][i][lane] = KO[tt][i];
        }

        if (phase == PH_F0) {
          // initial step size, part 1 (ode_determine_initdt)
          const double f2 = colsq(KO);
          double t0s = 0.0, t1s = 0.0;
          if (lane < NS) {
            const double a = my_u0 / my_sk;
            t0s = a * a;
            t1s = f2 / (my_sk * my_sk);
          }
          if (C::KIND == 1 && lane == NS) {
            const double a = Tval / (sm.abstol[NS] + fabs(Tval) * sm.reltol[NS]);
            t0s = a * a;
          }
          // two sums through the same scratch
          __syncwarp();
          if (lane < N) { wb.term[lane] = t0s; wb.g[lane] = t1s; }
          __syncwarp();
          double s0 = 0.0, s1 = 0.0;
#pragma unroll
          for (int i = 0; i < N; ++i) { s0 += wb.term[i]; s1 += wb.g[i]; }
          const double d0 = sqrt(s0 / N);
          d1 = sqrt(s1 / N);
          dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
          dt0 = jmin(dt0, dtmax);
          phase = PH_F1;
        } else if (phase == PH_F1) {
          // initial step size, part 2; then the pseudo-step that saves t0
#pragma unroll
          for (int tt = 0; tt < CT; ++tt)
#pragma unroll
            for (int i = 0; i < NS; ++i) KO[tt][i] -= wb.K[k1s][tt][i][lane];
          const double f2 = colsq(KO);
          const double s2 = sum_terms(lane < NS ? f2 / (my_sk * my_sk) : 0.0);
          const double d2 = sqrt(s2 / N) / dt0;
          const double dm = jmax(d1, d2);
          const double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(dm)) * sp.inv_order);
          dt = jmin(jmin(100.0 * dt0, dt1), dtmax);
          dtnew = dt;
          // pseudo-step: proposed state = U, K7 = K1, so the commit below is a no-op
#pragma unroll
          for (int tt = 0; tt < CT; ++tt)
#pragma unroll
            for (int i = 0; i < NS; ++i) {
              Y[tt][i] = U[tt][i];
              wb.K[6 - k1s][tt][i][lane] = wb.K[k1s][tt][i][lane];
            }
          bsum = asum;
          phase = PH_SAVE;
        } else if (phase < 6) {
          ++phase;
        } else {
          // all seven stages done: error estimate, PI controller, accept / reject
#pragma unroll
          for (int tt = 0; tt < CT; ++tt)
#pragma unroll
            for (int i = 0; i < NS; ++i) KO[tt][i] *= c_tsBT[6];
          for (int j = 0; j < 6; ++j) {
            const double a = c_tsBT[j];
            const int slot = (j == 0) ? k1s : j;
#pragma unroll
            for (int tt = 0; tt < CT; ++tt)
#pragma unroll
              for (int i = 0; i < NS; ++i) KO[tt][i] = fma(a, wb.K[slot][tt][i][lane], KO[tt][i]);
          }
#pragma unroll
          for (int tt = 0; tt < CT; ++tt)
#pragma unroll
            for (int i = 0; i < NS; ++i) KO[tt][i] *= dt;
          const double e2 = colsq(KO);
          bsum = colsq(Y);
          double term = 0.0;
          if (lane < NS) {
            // max(|u0|,|u1|) with dual magnitudes; sqrt is monotone, so one sqrt serves both
            const double sc = fma(sqrt(fmax(asum, bsum)), my_rt, my_at);
            term = e2 / (sc * sc);
          }
          const double EEst = sqrt(sum_terms(term) / N);
          double q11;
          const double q = pi_controller<C>(sp, EEst, qold, q11);
          dt_last = dt;
          if (EEst <= 1.0) {
            ++n_acc;
            qold = jmax(EEst, 1e-4);
            dtnew = dt / q;
            tprev = t;
            t = snap_t(t + dt, tend);
            phase = PH_SAVE;
          } else {
            ++n_rej;
            dt = dt / jmin(sp.inv_qmin, q11 / sp.gamma);
            phase = 1;
          }
        }
      } else {
        // ---- SAVE phase: every save time in (tprev, t] via the dense interpolant, loss and
        //      gradient fused (single instance), then commit the accepted step ----
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
            __syncwarp();
            if (lane < 7)
              wb.b[lane] = th * (sm.dense_r[lane][0] + th * (sm.dense_r[lane][1] +
                                 th * (sm.dense_r[lane][2] + th * sm.dense_r[lane][3])));
            __syncwarp();
#pragma unroll
            for (int tt = 0; tt < CT; ++tt)
#pragma unroll
              for (int i = 0; i < NS; ++i) KO[tt][i] = 0.0;
            for (int j = 0; j < 7; ++j) {
              const double bj = wb.b[j];
              const int slot = (j == 0) ? k1s : (j == 6 ? 6 - k1s : j);
#pragma unroll
              for (int tt = 0; tt < CT; ++tt)
#pragma unroll
                for (int i = 0; i < NS; ++i) KO[tt][i] = fma(bj, wb.K[slot][tt][i][lane], KO[tt][i]);
            }
#pragma unroll
            for (int tt = 0; tt < CT; ++tt)
#pragma unroll
              for (int i = 0; i < NS; ++i) KO[tt][i] = fma(dt, KO[tt][i], U[tt][i]);
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
              const double d = __ldg(data + off);
              double diff;
              if (sp.loss_kind == CRNN_LOSS_MAE_SCALED) {
                const double ys = sm.yscale[lane];
                diff = d / ys - yc / ys;
                g = (signbit(diff) ? 1.0 : -1.0) / ys;
              } else {
                diff = log(clampd(d, sp.pred_lo, sp.pred_hi)) - log(yc);
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
        // commit: u_n <- u_{n+1}, K1 <- K7 (FSAL: swap the slot roles, no copy)
#pragma unroll
        for (int tt = 0; tt < CT; ++tt)
#pragma unroll
          for (int i = 0; i < NS; ++i) U[tt][i] = Y[tt][i];
        k1s = 6 - k1s;
        asum = bsum;
        dt = jmin(dtnew, dtmax);
        phase = 1;
      }

      if (phase == 1) {  // loopheader! + check_error! before every step attempt
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

    // ---- per-trajectory outputs ----
    const double cnt = (double)sp.n_obs * (double)isave;
    const double ltot = warp_sum(loss_acc);
    if (lane == 0) {
      loss[traj] = isave > 0 ? ltot / cnt : __longlong_as_double(0x7ff8000000000000LL);
      if (n_saved) n_saved[traj] = isave;
      if (retcode) retcode[traj] = ret;
      if (stats) {
        crnn_stats s;
        s.n_accept = n_acc; s.n_reject = n_rej; s.n_rhs = n_rhs; s.n_jac = 0;
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

// Deterministic reduction of per-trajectory gradients: grad_each[N][np] -> grad_sum[np].
// Pass 1: block b sums its contiguous trajectory slab in index order -> partial[b][np].
// Pass 2 (same kernel, last block to finish): sums the partials in block order.
static __global__ void __launch_bounds__(256)
k_grad_reduce(const double* __restrict__ grad_each, long long ntraj, int np, double* __restrict__ partial,
              double* __restrict__ grad_sum, unsigned int* __restrict__ done) {
  const int nb = gridDim.x;
  const long long per = (ntraj + nb - 1) / nb;
  const long long lo = per * blockIdx.x, hi = (lo + per < ntraj) ? lo + per : ntraj;
  for (int c = threadIdx.x; c < np; c += blockDim.x) {
    double s = 0.0;
    for (long long i = lo; i < hi; ++i) s += grad_each[i * np + c];
    partial[(size_t)blockIdx.x * np + c] = s;
  }
  __threadfence();
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(done, 1u) == (unsigned)nb - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  for (int c = threadIdx.x; c < np; c += blockDim.x) {
    double s = 0.0;
    for (int b = 0; b < nb; ++b) s += partial[(size_t)b * np + c];
    grad_sum[c] = s;
  }
  if (threadIdx.x == 0) *done = 0;
}

}  // namespace crnn
