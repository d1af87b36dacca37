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

template <class C, int CT>
struct alignas(16) SensSmem {
  double seed[C::NW][32 * CT];  // dW/dp, zero padded; column 0 (value lane) is zero
  double w_in[C::NIN * C::NR];
  double w_b[C::NR];
  double yscale[C::N];
  int row2obs[C::N];
};

template <class C>
struct WarpBuf {
  double y[C::N];
  double x[C::N];
  double dx[C::N];
  double r[C::NR];
  double g[C::N];
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
  WarpBuf<C>* wbs = reinterpret_cast<WarpBuf<C>*>(smem_raw + sizeof(SensSmem<C, CT>));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WarpBuf<C>& wb = wbs[warp];

  for (int q = threadIdx.x; q < NW * 32 * CT; q += blockDim.x) (&sm.seed[0][0])[q] = seed_dev[q];
  for (int q = threadIdx.x; q < NIN * NR; q += blockDim.x) sm.w_in[q] = mp.w_in[q];
  for (int q = threadIdx.x; q < NR; q += blockDim.x) sm.w_b[q] = mp.w_b[q];
  for (int q = threadIdx.x; q < N; q += blockDim.x) {
    sm.yscale[q] = 1.0 / sp.inv_yscale[q];
    sm.row2obs[q] = sp.row2obs[q];
  }
  __syncthreads();
  const int np = ncol - 1;

  // columns owned by this lane; mask = takes part in norms
  bool isval[CT], live[CT];
#pragma unroll
  for (int t = 0; t < CT; ++t) {
    isval[t] = (t == 0 && lane == 0);
    live[t] = (lane + 32 * t) < ncol && (sp.incl_sens || isval[t]);
  }

  while (true) {
    unsigned long long tq = 0;
    if (lane == 0) tq = atomicAdd(queue, 1ull);
    const long long traj = (long long)__shfl_sync(0xffffffffu, tq, 0);
    if (traj >= ntraj) break;

    double U[CT][NS], Un[CT][NS], K1[CT][NS], K2[CT][NS], K3[CT][NS], K4[CT][NS], K5[CT][NS], K6[CT][NS],
        K7[CT][NS], Y[CT][NS];
    double u0v[NS], Tval = 0.0, xT = 0.0, mybT = 0.0;
#pragma unroll
    for (int i = 0; i < NS; ++i) u0v[i] = __ldg(u0 + traj * N + i);
    if (C::KIND == 1) { Tval = __ldg(u0 + traj * N + NS); xT = -1.0 / (mp.gas_R * Tval); }
    if (lane < NR) {
      mybT = sm.w_b[lane];
      if (C::KIND == 1) mybT = fma(sm.w_in[NS + NIN * lane], xT, mybT);
    }
#pragma unroll
    for (int t = 0; t < CT; ++t)
#pragma unroll
      for (int i = 0; i < NS; ++i) U[t][i] = isval[t] ? u0v[i] : 0.0;

    int nsave = sp.n_save;
    double tend = sp.t1;
    if (n_save_used) {
      int q = __ldg(n_save_used + traj);
      if (q > 0 && q <= sp.n_save) { nsave = q; tend = __ldg(sp.saveat + q - 1); }
    }
    const double t0 = sp.t0, dtmax = tend - t0;
    const double dtmin = fmax(ulp_of(t0), ulp_of(tend));
    const size_t pbase = (size_t)traj * sp.n_obs * sp.n_save;

    // f on all columns at stage state Yin -> Kout (warp-cooperative, 3 warp barriers)
    auto eval = [&](const double (&Yin)[CT][NS], double (&Kout)[CT][NS]) {
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NS; ++i) wb.y[i] = Yin[0][i];
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
      double x[NIN], dx[NS], r[NR];
#pragma unroll
      for (int i = 0; i < NS; ++i) { x[i] = wb.x[i]; dx[i] = wb.dx[i]; }
      if (C::KIND == 1) x[NS] = xT;
#pragma unroll
      for (int j = 0; j < NR; ++j) r[j] = wb.r[j];
#pragma unroll
      for (int t = 0; t < CT; ++t) {
        const int lc = lane + 32 * t;
        double sd[NS], q[NR];
#pragma unroll
        for (int i = 0; i < NS; ++i) sd[i] = Yin[t][i] * dx[i];
#pragma unroll
        for (int j = 0; j < NR; ++j) {
          double zd = sm.seed[NIN * NR + j][lc];
#pragma unroll
          for (int i = 0; i < NS; ++i) zd = fma(mp.w_in[i + NIN * j], sd[i], zd);
#pragma unroll
          for (int i = 0; i < NIN; ++i) zd = fma(sm.seed[i + NIN * j][lc], x[i], zd);
          if (isval[t]) zd = 1.0;
          q[j] = r[j] * zd;
        }
#pragma unroll
        for (int i = 0; i < NS; ++i) {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < NR; ++j) s = fma(mp.w_out[i + NS * j], q[j], s);
#pragma unroll
          for (int j = 0; j < NR; ++j) s = fma(sm.seed[NIN * NR + NR + i + NS * j][lc], r[j], s);
          Kout[t][i] = s;
        }
      }
    };

    // sum over (participating) columns of v[t][i]^2, for every i
    auto colsq = [&](const double (&V)[CT][NS], double (&out)[NS]) {
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        double s = 0.0;
#pragma unroll
        for (int t = 0; t < CT; ++t) s = live[t] ? fma(V[t][i], V[t][i], s) : s;
        out[i] = warp_sum(s);
      }
    };

    int n_rhs = 0, n_acc = 0, n_rej = 0;
    eval(U, K1); ++n_rhs;
    double dt;
    {
      double sk[NS], f2[NS], s0 = 0.0, s1 = 0.0;
      colsq(K1, f2);
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        sk[i] = sp.abstol[i] + fabs(u0v[i]) * sp.reltol[i];
        const double a = u0v[i] / sk[i];
        s0 = fma(a, a, s0);
        s1 += f2[i] / (sk[i] * sk[i]);
      }
      if (C::KIND == 1) { const double a = Tval / (sp.abstol[NS] + fabs(Tval) * sp.reltol[NS]); s0 = fma(a, a, s0); }
      const double d0 = sqrt(s0 / N), d1 = sqrt(s1 / N);
      double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
      dt0 = fmin(dt0, dtmax);
#pragma unroll
      for (int t = 0; t < CT; ++t)
#pragma unroll
        for (int i = 0; i < NS; ++i) Y[t][i] = fma(dt0, K1[t][i], U[t][i]);
      eval(Y, K2); ++n_rhs;
#pragma unroll
      for (int t = 0; t < CT; ++t)
#pragma unroll
        for (int i = 0; i < NS; ++i) Y[t][i] = K2[t][i] - K1[t][i];
      colsq(Y, f2);
      double s2 = 0.0;
#pragma unroll
      for (int i = 0; i < NS; ++i) s2 += f2[i] / (sk[i] * sk[i]);
      const double d2 = sqrt(s2 / N) / dt0;
      const double dm = fmax(d1, d2);
      const double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(dm)) * sp.inv_order);
      dt = fmin(fmin(100.0 * dt0, dt1), dtmax);
    }

    double G[CT], loss_acc = 0.0;
#pragma unroll
    for (int t = 0; t < CT; ++t) G[t] = 0.0;

    // publish one save column: value lane -> loss/pred, all lanes -> gradient
    auto emit = [&](int ks, const double (&Ys)[CT][NS]) {
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NS; ++i) wb.y[i] = Ys[0][i];
      }
      __syncwarp();
      if (lane < N) {
        const int q = sm.row2obs[lane];
        double g = 0.0;
        if (q >= 0) {
          const double y = (lane < NS) ? wb.y[lane] : Tval;
          const double yc = clampd(y, sp.pred_lo, sp.pred_hi);
          const bool inside = (y >= sp.pred_lo) && (y <= sp.pred_hi);
          const size_t off = pbase + q + (size_t)sp.n_obs * ks;
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
        for (int t = 0; t < CT; ++t) G[t] = fma(g, Ys[t][i], G[t]);
      }
    };

    double asum[NS];
    colsq(U, asum);
    double t = t0, qold = 1e-4, dt_last = 0.0;
    int isave = 0, ret = CRNN_RET_DEFAULT;
    long long iter = 0;
    while (isave < nsave && __ldg(sp.saveat + isave) <= t0) { emit(isave, U); ++isave; }

    while (t < tend) {
      ++iter;
      if (dt != dt) { ret = CRNN_RET_DTNAN; break; }
      if (iter > sp.maxiters) { ret = CRNN_RET_MAXITERS; break; }
      dt = fmin(dt, dtmax);
      dt = fmin(dt, tend - t);
      if (dt <= dtmin && tend - t > dtmin) { ret = CRNN_RET_DTLESSTHANMIN; break; }
      bool bad = false;
#pragma unroll
      for (int tt = 0; tt < CT; ++tt)
#pragma unroll
        for (int i = 0; i < NS; ++i) bad |= (U[tt][i] != U[tt][i]);
      if (__any_sync(0xffffffffu, bad)) { ret = CRNN_RET_UNSTABLE; break; }

#define CRNN_STAGE(OUT, EXPR)                                   \
  _Pragma("unroll") for (int tt = 0; tt < CT; ++tt)             \
  _Pragma("unroll") for (int i = 0; i < NS; ++i) OUT[tt][i] = fma(dt, (EXPR), U[tt][i]);
      CRNN_STAGE(Y, ts::a21 * K1[tt][i]);
      eval(Y, K2);
      CRNN_STAGE(Y, fma(ts::a32, K2[tt][i], ts::a31 * K1[tt][i]));
      eval(Y, K3);
      CRNN_STAGE(Y, fma(ts::a43, K3[tt][i], fma(ts::a42, K2[tt][i], ts::a41 * K1[tt][i])));
      eval(Y, K4);
      CRNN_STAGE(Y, fma(ts::a54, K4[tt][i], fma(ts::a53, K3[tt][i], fma(ts::a52, K2[tt][i], ts::a51 * K1[tt][i]))));
      eval(Y, K5);
      CRNN_STAGE(Y, fma(ts::a65, K5[tt][i], fma(ts::a64, K4[tt][i], fma(ts::a63, K3[tt][i],
                    fma(ts::a62, K2[tt][i], ts::a61 * K1[tt][i])))));
      eval(Y, K6);
      CRNN_STAGE(Un, fma(ts::a76, K6[tt][i], fma(ts::a75, K5[tt][i], fma(ts::a74, K4[tt][i],
                     fma(ts::a73, K3[tt][i], fma(ts::a72, K2[tt][i], ts::a71 * K1[tt][i]))))));
      eval(Un, K7);
#undef CRNN_STAGE
      n_rhs += 6;

#pragma unroll
      for (int tt = 0; tt < CT; ++tt)
#pragma unroll
        for (int i = 0; i < NS; ++i)
          Y[tt][i] = dt * fma(ts::bt7, K7[tt][i], fma(ts::bt6, K6[tt][i], fma(ts::bt5, K5[tt][i],
                          fma(ts::bt4, K4[tt][i], fma(ts::bt3, K3[tt][i], fma(ts::bt2, K2[tt][i], ts::bt1 * K1[tt][i]))))));
      double e2[NS], bsum[NS];
      colsq(Y, e2);
      colsq(Un, bsum);
      double acc = 0.0;
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        const double sc = fma(fmax(sqrt(asum[i]), sqrt(bsum[i])), sp.reltol[i], sp.abstol[i]);
        acc += e2[i] / (sc * sc);
      }
      const double EEst = sqrt(acc / N);
      double q11;
      const double q = pi_controller<C>(sp, EEst, qold, q11);
      dt_last = dt;
      if (EEst <= 1.0) {
        ++n_acc;
        qold = fmax(EEst, 1e-4);
        const double dtnew = dt / q;
        const double tprev = t;
        t = snap_t(t + dt, tend);
        while (isave < nsave) {
          const double tsv = __ldg(sp.saveat + isave);
          if (!(tsv <= t)) break;
          if (tsv == t) {
            emit(isave, Un);
          } else {
            double b[7];
            ts::dense_b((tsv - tprev) / dt, b);
#pragma unroll
            for (int tt = 0; tt < CT; ++tt)
#pragma unroll
              for (int i = 0; i < NS; ++i)
                Y[tt][i] = fma(dt, fma(b[6], K7[tt][i], fma(b[5], K6[tt][i], fma(b[4], K5[tt][i], fma(b[3], K4[tt][i],
                               fma(b[2], K3[tt][i], fma(b[1], K2[tt][i], b[0] * K1[tt][i])))))), U[tt][i]);
            emit(isave, Y);
          }
          ++isave;
        }
#pragma unroll
        for (int tt = 0; tt < CT; ++tt)
#pragma unroll
          for (int i = 0; i < NS; ++i) { U[tt][i] = Un[tt][i]; K1[tt][i] = K7[tt][i]; }
#pragma unroll
        for (int i = 0; i < NS; ++i) asum[i] = bsum[i];
        dt = fmin(dtnew, dtmax);
      } else {
        ++n_rej;
        dt = dt / fmin(sp.inv_qmin, q11 / sp.gamma);
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
__global__ void __launch_bounds__(256)
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
