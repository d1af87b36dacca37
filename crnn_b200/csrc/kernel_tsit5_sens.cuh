// kernel_tsit5_sens.cuh — loss + forward-mode gradient for a batch: one WARP owns one trajectory.
//
// Replaces `ForwardDiff.gradient(x -> loss_neuralode(x, i_exp), p)` (case2/case2.jl:132-137,195;
// Zygote.forwarddiff in case1/case1.jl:195-199): duals pushed through the adaptive solver.
// Lane l of the warp owns dual column l (+32 per column tile): column 0 is the value, column
// c >= 1 the partial d/dp_c.  All columns share one step sequence; the partials take part in the
// error norm exactly like DiffEqBase's norm over Dual arrays.
//
// Per stage the value path (NS logs, NR exps) is spread over lanes — lane i takes log(u_i),
// lane j takes exp(z_j) — and broadcast through a few bytes of per-warp shared memory, so the
// warp issues ONE log and ONE exp per stage instead of NS+NR; every lane then applies
// J(u)*S + (df/dW)*dW/dp_c matrix-free to its own column.
//
// Two seed layouts (template flag R1):
//   R1 = false  dense dW/dp column per lane, NW shared-memory loads per stage (any p2vec);
//   R1 = true   structured columns: dW_in non-zero in ONE row i (any reactions), db arbitrary,
//               dW_out non-zero in at most ONE entry (i', j') — the shape of every p2vec in the
//               reference (a weight and its mirrored w_in entry, a bias, an activation energy, or
//               a slope that rescales all biases / activation energies).  The forcing term is
//               W_out*(r .* (a_j*x_i + b_j)) + o*r_j'*e_i': 2*NR seed loads instead of NW.
//
// Code-size discipline (profiles/r1_*): the whole integration is a phase machine around ONE
// inlined instance of the RHS, ONE of the norm reduction and ONE of the loss/gradient block,
// with the stage vectors in shared memory and the tableau in the constant bank.  The first,
// fully unrolled version was 160 KB of SASS and spent 11 of 15 stall cycles per instruction
// waiting for instruction fetch.
// The loss (MAE-scaled or MAE-log) and its gradient are fused at each save point, so the
// n_state x n_save x np sensitivity tensor never leaves the SM.
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

// One structured seed column (R1 layout), one entry per lane and tile; its per-reaction parts
// a_j = dW_in[i_in, j] and b_j = db_j sit in SensSmem::seed rows [0,NR) and [NR,2NR).
struct R1Desc {
  double o;  // dW_out[i_out, j_out] (out_scale folded in)
  int i_in, i_out, j_out, pad;
};

template <class C, int CT, bool R1, int WPT = 1>
struct alignas(16) SensSmem {
  double seed[R1 ? 2 * C::NR : C::NW][32 * CT * WPT];  // dW/dp (dense, or the a_j / b_j rows); column 0 = value = 0
  double w_in[C::NIN * C::NR];
  double w_b[C::NR];
  double inv_ys[C::N];
  double abstol[C::N], reltol[C::N];
  double dense_r[7][4];  // Tsit5 dense-output polynomial coefficients (lane j takes b_j)
  int row2obs[C::N];
  ModelP<C> mpw;         // DEVW: the model read from DEVICE memory (on-device training loop, kernel_train.cuh)
};

// Per-warp working set.  K holds the seven stage derivatives of every column.  Inside a
// (slot, tile) block the components of lane l are stored as lane-interleaved PAIRS -
// (2p, 2p+1) at double2 index p*32 + l, an odd last component as a plain row - so one
// LDS.128 / STS.128 moves two components (kload / kstore below), conflict-free.
template <class C, int CT>
struct alignas(16) WarpBuf {
  double K[7][CT][C::NS][32];
  double red[C::NS][32];      // column-sum scratch of the dual-aware norms
  double y[C::N];
  double x[C::N];             // x[NS] = -1/(R T) for F1
  double dx[C::N];
  double r[C::NR];
  double g[C::N];
  double b[8];                // dense-output weights b_j(theta)
  double term[2][C::N];
  double rp[C::NS][8];        // partial row sums (RP lanes share one row of `red`)
  double cold[4];             // rarely-read scalars kept out of registers: tend, dtmax, dtmin, dt of the last attempt
  double part[8][C::N];       // WPT > 1: per-warp partial row sums of the trajectory's warp group
  long long trajslot;         // WPT > 1: trajectory index broadcast
  int flag, pad;              // WPT > 1: NaN flag of the group
};

// WPT warps share one trajectory (32*CT*WPT dual columns): the group synchronises on a named
// barrier instead of __syncwarp, the small broadcast arrays live in the group leader's WarpBuf.
// AUTO = true (WPT = 1 only): the kernel also carries OrdinaryDiffEq's AutoSwitch stiffness monitor of
// AutoTsit5(Rosenbrock23()) (case2/case2.jl:26; oracle solve_one header).  As long as the composite stays on Tsit5 - always,
// on the trained case2 CRNN - this IS the composite algorithm, at this kernel's speed; a trajectory whose counter asks for
// Rosenbrock23 is abandoned and its index appended to sel_list, and the host re-runs exactly those through the generic
// composite kernel (kernel_gen_sens.cuh), which overwrites their outputs.
// DEVW = true: weights come from device memory (mp_dev, written by a p2vec kernel of the on-device training loop) instead
// of the by-value kernel parameter; they are staged in shared memory, everything else is unchanged.
template <class C, int CT, int WARPS, int MINB, bool R1, int WPT = 1, bool AUTO = false, bool DEVW = false>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_tsit5_sens(const __grid_constant__ ModelP<C> mp, const __grid_constant__ SolveP<C> sp,
             const double* __restrict__ seed_dev, const R1Desc* __restrict__ desc_dev, int ncol,
             const double* __restrict__ u0, const int* __restrict__ n_save_used, long long ntraj,
             const double* __restrict__ data, double* __restrict__ loss, double* __restrict__ grad_each,
             double* __restrict__ pred, int* __restrict__ n_saved, int* __restrict__ retcode,
             crnn_stats* __restrict__ stats, unsigned long long* __restrict__ queue,
             const long long* __restrict__ in_idx, long long* __restrict__ sel_list = nullptr,
             unsigned int* __restrict__ sel_count = nullptr, const ModelP<C>* __restrict__ mp_dev = nullptr) {
  static_assert(!AUTO || WPT == 1, "the AutoSwitch monitor is built for one warp per trajectory");
  // in_idx (or NULL): trajectory `traj` of this call reads u0 / data of dataset row in_idx[traj] (crnn_loss_grad_indexed);
  // every output stays at position traj
  constexpr int NS = C::NS, NR = C::NR, N = C::N, NIN = C::NIN, NW = C::NW;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  static_assert(WPT >= 1 && WPT <= 8 && WARPS % WPT == 0 && (WPT == 1 || CT == 1), "bad warp grouping");
  SensSmem<C, CT, R1, WPT>& sm = *reinterpret_cast<SensSmem<C, CT, R1, WPT>*>(smem_raw);
  WarpBuf<C, CT>* wbs = reinterpret_cast<WarpBuf<C, CT>*>(smem_raw + sizeof(SensSmem<C, CT, R1, WPT>));
  // volatile: keeps lane/warp in registers (the compiler otherwise re-derives them from %tid all over the loop)
  unsigned lane_u, tid_u;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane_u));
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid_u));
  const int lane = (int)lane_u, warp = (int)(tid_u >> 5);
  const int wig = warp % WPT, grp = warp / WPT;  // warp in group, group in block
  WarpBuf<C, CT>& wb = wbs[warp];        // own: K, red
  WarpBuf<C, CT>& gb = wbs[warp - wig];  // group leader's: broadcast arrays
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
  auto gsync = [&]() {
    if (WPT == 1) __syncwarp();
    else if (grp == 0) asm volatile("bar.sync 1, %0;" ::"n"(WPT * 32) : "memory");  // immediate ids: a register id
    else asm volatile("bar.sync 2, %0;" ::"n"(WPT * 32) : "memory");                 // would reserve all 16 barriers
  };

  for (int q = threadIdx.x; q < (R1 ? 2 * NR : NW) * 32 * CT * WPT; q += blockDim.x) (&sm.seed[0][0])[q] = seed_dev[q];
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
  double d_o[CT];
  int d_iin[CT], d_iout[CT], d_jout[CT];
#pragma unroll
  for (int t = 0; t < CT; ++t) {
    isval[t] = (wig == 0 && t == 0 && lane == 0);
    d_o[t] = 0.0; d_iin[t] = d_iout[t] = d_jout[t] = 0;
    if (R1) {
      const R1Desc d = desc_dev[lane + 32 * (t + CT * wig)];
      d_o[t] = d.o; d_iin[t] = d.i_in; d_iout[t] = d.i_out; d_jout[t] = d.j_out;
    }
    live[t] = (lane + 32 * (t + CT * wig)) < ncol && (sp.incl_sens || isval[t]);
  }

  constexpr int PH_F0 = 0, PH_F1 = 7, PH_SAVE = 8;  // phases 1..6 are the Tsit5 stages

  while (true) {
    long long traj;
    if (WPT == 1) {
      unsigned long long tq = 0;
      if (lane == 0) tq = atomicAdd(queue, 1ull);
      traj = (long long)__shfl_sync(0xffffffffu, tq, 0);
    } else {
      if (isval[0]) { gb.trajslot = (long long)atomicAdd(queue, 1ull); gb.flag = 0; }
      gsync();
      traj = gb.trajslot;
      gsync();
    }
    if (traj >= ntraj) break;
    const double* __restrict__ u0t = u0 + (in_idx ? __ldg(in_idx + traj) : traj) * N;

    // U: state; Y: stage state (holds the proposed u_{n+1} from stage 6 through the save phase);
    // KO: RHS output / scratch
    double U[CT][NS], Y[CT][NS], KO[CT][NS];
    double xT = 0.0, mybT = 0.0;
    if (C::KIND == 1) xT = -1.0 / (M.gas_R * __ldg(u0t + NS));
    if (lane < NR) {
      mybT = sm.w_b[lane];
      if (C::KIND == 1) mybT = fma(sm.w_in[NS + NIN * lane], xT, mybT);
    }
    if (C::KIND == 1 && isval[0]) gb.x[NS] = xT;
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

    int n_acc = 0, n_rej = 0;  // RHS evaluations = 2 + 6 * attempts, loop iterations = attempts (derived, not counted)
    double G[CT], loss_acc = 0.0;
#pragma unroll
    for (int t = 0; t < CT; ++t) G[t] = 0.0;
    double asum = 0.0, bsum = 0.0;  // lane i < NS: dual magnitude^2 of u_i at t_n / t_{n+1}
    if (lane < NS) { const double v = __ldg(u0t + lane); asum = v * v; }
    // during the two initial-step phases dt holds dt0 and dtnew holds d1 (both are free until the first step)
    double t = t0, tprev = t0, dt = 0.0, dtnew = 0.0, lqold = lean_log(1e-4);  // log(qoldinit)
    int isave = 0, ret = CRNN_RET_DEFAULT, phase = PH_F0, k1s = 0;  // k1s: slot of K1 (0 or 6), K7 in 6-k1s
    double eigen_est = 0.0, den_l = 0.0;  // AUTO: |k7 - k6| / |u_{n+1} - g6|, this lane's share of the denominator
    int sw_count = 0;
    bool needs_composite = false;
    // the next save time and this lane's next target are fetched one save ahead: their global-load
    // latency then overlaps the step in between instead of stalling the save phase
    const int my_q = (wig == 0 && lane < N) ? sm.row2obs[lane] : -1;
    double ts_next = __ldg(sp.saveat), d_next = 0.0;
    if (my_q >= 0) d_next = __ldg(datat + my_q);

    while (true) {
      if (phase != PH_SAVE) {
        // ---- stage state Y = U + h * sum_{j<nj} A[phase][j] K_j ----
        {
          const int nj = (phase == PH_F1) ? 1 : phase;
          const double h = dt;
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
          if (AUTO && phase == 5) {   // g6 (the stage-6 state) parked in the K7 slot, free until stage 6 stores K7
#pragma unroll
            for (int tt = 0; tt < CT; ++tt) kstore(6 - k1s, tt, Y[tt]);
          }
          if (AUTO && phase == 6) {
            den_l = 0.0;
#pragma unroll
            for (int tt = 0; tt < CT; ++tt) {
              kload(6 - k1s, tt, kv);
#pragma unroll
              for (int i = 0; i < NS; ++i) { const double dd = Y[tt][i] - kv[i]; den_l = live[tt] ? fma(dd, dd, den_l) : den_l; }
            }
          }
        }

        // ---- KO = f(Y) on all columns: the single RHS instance (3 warp barriers) ----
        if (isval[0]) {
#pragma unroll
          for (int i = 0; i < NS; ++i) gb.y[i] = Y[0][i];
        }
        gsync();
        if (wig == 0 && lane < NS) {
          const double yi = gb.y[lane];
          const double uc = clampd(yi, M.lb, M.ub);
          const bool inside = (yi >= M.lb) && (yi <= M.ub);
          gb.x[lane] = lean_log(uc);
          gb.dx[lane] = inside ? __drcp_rn(uc) : 0.0;
        }
        gsync();
        if (wig == 0 && lane < NR) {
          double z = mybT;
#pragma unroll
          for (int i = 0; i < NS; ++i) z = fma(sm.w_in[i + NIN * lane], gb.x[i], z);
          gb.r[lane] = lean_exp(z);
        }
        gsync();
        {
          double dx[NS], r[NR];
#pragma unroll
          for (int i = 0; i < NS; ++i) dx[i] = gb.dx[i];
#pragma unroll
          for (int j = 0; j < NR; ++j) r[j] = gb.r[j];
#pragma unroll
          for (int tt = 0; tt < CT; ++tt) {
            const int lc = lane + 32 * (tt + CT * wig);
            double sd[NS], q[NR];
#pragma unroll
            for (int i = 0; i < NS; ++i) sd[i] = Y[tt][i] * dx[i];
            const double xin = R1 ? gb.x[d_iin[tt]] : 0.0;
#pragma unroll
            for (int j = 0; j < NR; ++j) {
              double zd = R1 ? fma(sm.seed[j][lc], xin, sm.seed[NR + j][lc]) : sm.seed[NIN * NR + j][lc];
#pragma unroll
              for (int i = 0; i < NS; ++i) zd = fma(M.w_in[i + NIN * j], sd[i], zd);
              if (!R1) {
#pragma unroll
                for (int i = 0; i < NS; ++i) zd = fma(sm.seed[i + NIN * j][lc], gb.x[i], zd);
                if (C::KIND == 1) zd = fma(sm.seed[NS + NIN * j][lc], xT, zd);
              }
              if (isval[tt]) zd = 1.0;
              q[j] = r[j] * zd;
            }
            const double ro = R1 ? d_o[tt] * gb.r[d_jout[tt]] : 0.0;
#pragma unroll
            for (int i = 0; i < NS; ++i) {
              double s = 0.0;
#pragma unroll
              for (int j = 0; j < NR; ++j) s = fma(M.w_out[i + NS * j], q[j], s);
              if (!R1) {
#pragma unroll
                for (int j = 0; j < NR; ++j) s = fma(sm.seed[NIN * NR + NR + i + NS * j][lc], r[j], s);
              } else if (d_iout[tt] == i) {
                s += ro;
              }
              KO[tt][i] = s;
            }
          }
        }

        // ---- store KO into its stage slot ----
        {
          // F0 -> K1 ; F1 -> slot 1 (scratch, free until stage 1 writes K2) ; stage s -> K_{s+1}
          const int dst = (phase == PH_F0) ? k1s : (phase == PH_F1) ? 1 : (phase == 6 ? 6 - k1s : phase);
          if (AUTO && phase == 6) {   // eigen_est = |k7 - k6| / |u_{n+1} - g6| over the columns that take part in the norm
            double num_l = 0.0;
#pragma unroll
            for (int tt = 0; tt < CT; ++tt) {
              double kv[NS];
              kload(5, tt, kv);
#pragma unroll
              for (int i = 0; i < NS; ++i) { const double dd = KO[tt][i] - kv[i]; num_l = live[tt] ? fma(dd, dd, num_l) : num_l; }
            }
            eigen_est = sqrt(warp_sum(num_l) / sp.eig_cnt) / sqrt(warp_sum(den_l) / sp.eig_cnt);
          }
#pragma unroll
          for (int tt = 0; tt < CT; ++tt) kstore(dst, tt, KO[tt]);
        }

        if (phase >= 1 && phase < 6) {
          ++phase;
        } else {
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
          // (a) lane i < NS <- sum over participating columns of KO[.][i]^2 (pass 0) and, for the
          //     error estimate, of Y[.][i]^2 (pass 1: dual magnitude of u_{n+1}).  One code instance.
          double rsum = 0.0;
          const int npass = (phase == 6) ? 2 : 1;
#pragma unroll 1
          for (int pass = 0; pass < npass; ++pass) {
            gsync();
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
            // RP lanes share a row: partial sums of 32/RP skewed (conflict-free) columns, then lane i adds them
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
              if (WPT > 1) gb.part[wig][lane] = tot;
            }
            if (WPT > 1) {  // combine the group's per-warp partial sums (same order in every warp)
              gsync();
              tot = 0.0;
              if (lane < NS) {
#pragma unroll
                for (int w = 0; w < WPT; ++w) tot += gb.part[w][lane];
              }
            }
            if (lane < NS) { if (pass == 0) rsum = tot; else bsum = tot; }
          }
          // (b) per-row terms (lane i < NS), then their sums in every lane
          double term0 = 0.0, term1 = 0.0;
          if (lane < NS) {
            if (phase == 6) {
              // max(|u0|,|u1|) with dual magnitudes; sqrt is monotone, so one sqrt serves both
              const double sc = fma(sqrt(fmax(asum, bsum)), sm.reltol[lane], sm.abstol[lane]);
              term0 = rsum / (sc * sc);
            } else {
              const double my_u0 = __ldg(u0t + lane), my_sk = sm.abstol[lane] + fabs(my_u0) * sm.reltol[lane];
              const double a = my_u0 / my_sk;
              term0 = rsum / (my_sk * my_sk);
              term1 = a * a;
            }
            if (wig == 0) { gb.term[0][lane] = term0; gb.term[1][lane] = term1; }
          }
          gsync();
          double s0 = 0.0, s1 = 0.0;
#pragma unroll
          for (int i = 0; i < NS; ++i) { s0 += gb.term[0][i]; s1 += gb.term[1][i]; }

          if (phase == PH_F0) {
            // initial step size, part 1 (ode_determine_initdt)
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
            // initial step size, part 2; then the pseudo-step that saves t0
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
            bsum = asum;
            phase = PH_SAVE;
          } else {
            // all seven stages done: PI controller, accept / reject
            const double EEst = sqrt(s0 / sp.norm_cnt);
            // PI controller (pi_controller of crnn_dev.cuh) with log(qold) carried between steps: qold is the previous
            // step's max(EEst, 1e-4), whose log this code already took — same values, one lean_log less per step
            double q11 = 0.0, q = sp.inv_qmax, lE = 0.0;
            if (EEst != 0.0) {
              lE = lean_log(EEst);
              q11 = lean_exp(LM_MUL(sp.beta1, lE));
              q = jmax(sp.inv_qmax, jmin(sp.inv_qmin, q11 / lean_exp(LM_MUL(sp.beta2, lqold)) / sp.gamma));
            }
            if (isval[0]) wb.cold[3] = dt;
            if (EEst <= 1.0) {
              ++n_acc;
              lqold = (EEst > 1e-4) ? lE : lean_log(1e-4);   // log(qold), qold = max(EEst, qoldinit)
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
        // ---- SAVE phase: every save time in (tprev, t] via the dense interpolant, loss and
        //      gradient fused (single instance), then commit the accepted step ----
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
            ts::dense_b((tsv - tprev) / dt, b);  // every lane evaluates the seven weights itself: no barrier
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
            for (int i = 0; i < NS; ++i) gb.y[i] = KO[0][i];
          }
          gsync();
          if (wig == 0 && lane < N) {
            const int q = my_q;
            double g = 0.0;
            if (q >= 0) {
              const double y = (lane < NS) ? gb.y[lane] : __ldg(u0t + NS);  // the T row never changes
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
            gb.g[lane] = g;
          }
          gsync();
#pragma unroll
          for (int i = 0; i < NS; ++i) {
            const double g = gb.g[i];
#pragma unroll
            for (int tt = 0; tt < CT; ++tt) G[tt] = fma(g, KO[tt][i], G[tt]);
          }
          ++isave;
          if (isave < nsave) {
            ts_next = __ldg(sp.saveat + isave);
            if (my_q >= 0) d_next = __ldg(datat + my_q + (size_t)sp.n_obs * isave);
          }
        }
        // commit: u_n <- u_{n+1}, K1 <- K7 (FSAL: swap the slot roles, no copy)
#pragma unroll
        for (int tt = 0; tt < CT; ++tt)
#pragma unroll
          for (int i = 0; i < NS; ++i) U[tt][i] = Y[tt][i];
        k1s = 6 - k1s;
        asum = bsum;
        dt = jmin(dtnew, wb.cold[1]);
        phase = 1;
      }

      if (phase == 1) {  // loopheader! + check_error! before every step attempt
        const double tend = wb.cold[0], dtmin = wb.cold[2];
        if (!(t < tend)) break;
        if (AUTO && n_acc + n_rej > 0) {  // AutoSwitch choice function, before every attempt but the first, on the proposed dt
          const bool stiff = fabs(eigen_est * dt / 3.5068) > 0.9;
          sw_count = stiff ? (sw_count < 0 ? 1 : sw_count + 1) : (sw_count > 0 ? -1 : sw_count - 1);
          if (sw_count > 10) { needs_composite = true; break; }
        }
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
        if (WPT > 1) {  // any warp of the group
          if (bad && lane == 0) gb.flag = 1;
          gsync();
          bad = gb.flag != 0;
        }
        if (bad) { ret = CRNN_RET_UNSTABLE; break; }
      }
    }
    if (ret == CRNN_RET_DEFAULT) ret = CRNN_RET_SUCCESS;
    if (AUTO && needs_composite) {
      ret = CRNN_RET_DEFAULT;   // provisional: the composite kernel rewrites every output of this trajectory
      if (lane == 0) sel_list[atomicAdd(sel_count, 1u)] = traj;
    }

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
      const int c = lane + 32 * (tt + CT * wig);
      if (c >= 1 && c < ncol) grad_each[(size_t)traj * np + (c - 1)] = isave > 0 ? G[tt] / cnt : 0.0;
    }
    if (wig == 0 && pred && isave < sp.n_save) {
      for (int q = isave * sp.n_obs + lane; q < sp.n_save * sp.n_obs; q += 32) pred[pbase + q] = 0.0;
    }
    gsync();
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
