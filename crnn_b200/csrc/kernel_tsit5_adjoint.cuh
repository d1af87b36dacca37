// kernel_tsit5_adjoint.cuh — loss + gradient by the INTERPOLATING ADJOINT (BASELINE config 4): one WARP
// owns one trajectory, lane i owns state component i (n_state <= 32, n_reac <= 32, runtime dimensions).
//
// The reference never runs an adjoint (its scripts differentiate forward-mode only), so the
// policies are ours and are spelled out in oracle/crnn_oracle.c (solve_one_adjoint), which this kernel
// mirrors step for step:
//   forward   Tsit5 value solve; every accepted step's (t_n, dt_n, u_n, k1..k7) is recorded — in SHARED
//             memory for the first `cap_s` steps (the usual case: nothing touches HBM), spilling to a
//             per-warp global scratch beyond that;
//   backward  lambda' = -J(u(t))^T lambda by adaptive Tsit5 (error control on lambda only) from t_reached
//             to t0, stopping at every save time where lambda jumps by dL/du(t_k) (loss fused here);
//             u(t) from the recorded dense output; the parameter quadrature rides the step's own
//             b-weights as three outer products in physical-weight space
//                 G_in[i,j] = int x_i g_j r_j, G_b[j] = int g_j r_j, G_out[i,j] = int s_i lambda_i r_j.
//   output    vec(G) per trajectory; k_grad_reduce sums over trajectories (deterministic order) and
//             k_seed_contract applies dW/dp^T.
// Cost is independent of np: the right tool for case3 (np = 153) and larger parameter vectors.
#pragma once
#include "crnn_dev.cuh"
#include "wide_common.cuh"  // WideP, WideBlock, KW_MAXN, F2 tables

namespace crnn {

namespace tsc {
__constant__ double A[7][6] = {
    {0, 0, 0, 0, 0, 0},
    {ts::a21, 0, 0, 0, 0, 0},
    {ts::a31, ts::a32, 0, 0, 0, 0},
    {ts::a41, ts::a42, ts::a43, 0, 0, 0},
    {ts::a51, ts::a52, ts::a53, ts::a54, 0, 0},
    {ts::a61, ts::a62, ts::a63, ts::a64, ts::a65, 0},
    {ts::a71, ts::a72, ts::a73, ts::a74, ts::a75, ts::a76}};
__constant__ double BT[7] = {ts::bt1, ts::bt2, ts::bt3, ts::bt4, ts::bt5, ts::bt6, ts::bt7};
__constant__ double C[7] = {0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0};
__constant__ double R[7][4] = {{ts::r11, ts::r12, ts::r13, ts::r14}, {0.0, ts::r22, ts::r23, ts::r24},
                               {0.0, ts::r32, ts::r33, ts::r34}, {0.0, ts::r42, ts::r43, ts::r44},
                               {0.0, ts::r52, ts::r53, ts::r54}, {0.0, ts::r62, ts::r63, ts::r64},
                               {0.0, ts::r72, ts::r73, ts::r74}};
}  // namespace tsc

struct AdjP {
  WideP w;                 // dimensions, tolerances, weights (w_out WITH out_scale folded), saveat, row2obs
  const double* scale;     // device [n_species] out_scale (1 if none): G_out needs it separately
  const double* inv_ys;    // device [n_state] 1/yscale per state row (1 where unused)
  double* scratch;         // global overflow record, per warp slot
  int cap_s, cap_g;        // record capacity (steps) in shared / global memory
  int nw, loss_kind;
  int discrete;            // 1: discrete adjoint (reverse-mode through the recorded steps), 0: interpolating adjoint
  int mlp_extra;           // F4: doubles per warp for the MLP's activations / activation derivatives / deltas (32 + (3L + 1) * mlp_stride), else 0
  int mlp_stride;          // F4: slots per layer in those arrays (the widest layer, rounded up to even)
  int mlp_np, mlp_np_raw;  // F4: number of MLP parameters, rounded up to even (a per-block shared-memory copy) and as is; else 0
  int gs_len;              // length of the stage accumulator GS: (nw rounded up to even) for the interpolating adjoint, 0 for the discrete one
};

constexpr int ADJ_MAX_ENT = 16;  // quadrature entries per lane: n_w <= 512

// MLP: the F4 flavour (MLP-augmented inputs, yeast_glycolysis.jl:128-142): the forward pass evaluates the Flux chain lane-per-neuron,
// the adjoint RHS goes back through it (oracle adj_rhs_f4) and the quadrature runs in the extended weight space
// [vec(w_in); w_b; vec(w_out); w_J; mlp_params] - its own instantiation.
template <int WARPS, bool F2, int MINB = 2, bool MLP = false>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_tsit5_adjoint(const __grid_constant__ AdjP P, const double* __restrict__ u0, const int* __restrict__ n_save_used,
                long long ntraj, const double* __restrict__ data, double* __restrict__ loss,
                double* __restrict__ gw_each, double* __restrict__ pred, int* __restrict__ n_saved,
                int* __restrict__ retcode, crnn_stats* __restrict__ stats, unsigned long long* __restrict__ queue,
                const long long* __restrict__ in_idx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const WideP& W = P.w;
  const int n = W.n, ns = W.ns, nin = W.nin, nr = W.nr, nw = P.nw;
  const int stride = 8 * n + 2;  // doubles per recorded step: t, dt, u, k1..k7
  WideBlockLite& sb = *reinterpret_cast<WideBlockLite*>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // block: WideBlockLite | (F4: MLP parameters)    per warp: x[32] r[32] lam[32] gr[32] chi[32] sl[32] one[2] | (F4: MLP arrays) |
  // GW[nw] | (interpolating: GS[nw]) | record[cap_s][stride]
  constexpr int ADJ_FIXED = 6 * 32 + 2;   // crnn_api.cu::loss_grad_adjoint sizes the launch with the same number
  const size_t per_warp = ADJ_FIXED + (MLP ? P.mlp_extra : 0) + (size_t)((nw + 1) & ~1) + (size_t)P.gs_len + (size_t)P.cap_s * stride;
  double* s_mw = reinterpret_cast<double*>(smem_raw + sizeof(WideBlockLite));   // F4: the block's copy of the MLP parameters
  double* wbase = s_mw + (MLP ? P.mlp_np : 0) + per_warp * warp;
  double* s_x = wbase; double* s_r = wbase + 32; double* s_lam = wbase + 64; double* s_gr = wbase + 96;
  double* s_chi = wbase + 128;
  double* s_sl = wbase + 160;   // scale_i * lambda_i / rho: the left factor of the G_out outer product
  // F4: v[32] (cotangents of the augmented input rows; also the state broadcast) | a_l[ms], l = 0..L | act'_l[ms], l < L | delta_l[ms], l < L
  // (ms = P.mlp_stride slots per layer)
  const int ML = MLP ? W.mlp_layers : 0;
  const int ms = MLP ? P.mlp_stride : 0;
  double* s_v = wbase + ADJ_FIXED;
  auto A_ = [&](int l) -> double* { return wbase + ADJ_FIXED + 32 + ms * l; };
  auto SP_ = [&](int l) -> double* { return wbase + ADJ_FIXED + 32 + ms * (ML + 1 + l); };
  auto DL_ = [&](int l) -> double* { return wbase + ADJ_FIXED + 32 + ms * (2 * ML + 1 + l); };
  double* GW = wbase + ADJ_FIXED + (MLP ? P.mlp_extra : 0); double* GS = GW + ((nw + 1) & ~1);
  if (lane == 0) wbase[192] = 1.0;   // the unit left factor of the G_b entries
  double* rec_s = GS + P.gs_len;
  if (MLP)
    for (int q = threadIdx.x; q < P.mlp_np; q += blockDim.x) s_mw[q] = q < P.mlp_np_raw ? W.mlp_params[q] : 0.0;
  double* rec_g = P.scratch + ((size_t)blockIdx.x * WARPS + warp) * (size_t)P.cap_g * stride;

  for (int q = threadIdx.x; q < KW_MAXN * KW_MAXN; q += blockDim.x) {
    const int i = q / KW_MAXN, j = q % KW_MAXN;
    sb.w_inT[i][j] = (i < nin && j < nr) ? W.w_inT[i * KW_MAXN + j] : 0.0;
    sb.w_out[i][j] = (i < nr && j < ns) ? W.w_out[j + ns * i] : 0.0;  // [reaction][species], scaled
  }
  for (int q = threadIdx.x; q < KW_MAXN; q += blockDim.x) sb.w_b[q] = q < nr ? W.w_b[q] : 0.0;
  __syncthreads();

  const bool isp = lane < ns;
  const double my_at = lane < n ? W.abstol[lane] : 1.0, my_rt = lane < n ? W.reltol[lane] : 0.0;
  const int my_obs = lane < n ? W.row2obs[lane] : -1;
  const double my_iys = lane < n ? P.inv_ys[lane] : 1.0;
  const double my_scale = isp ? P.scale[lane] : 0.0;
  constexpr bool f2 = F2;
  const double my_mw = (f2 && isp) ? __ldg(W.mw + lane) : 1.0;
  // quadrature entries: lane owns e = lane + 32*q.  Every entry is a product of two per-warp shared-memory values; the table
  // (one per block, in shared memory - sixteen registers per thread when it was a local array) holds their offsets into wbase:
  //   G_in[i,j] = x_i * (g_j r_j)     G_b[j] = 1 * (g_j r_j)     G_out[i,j] = (scale_i lambda_i) * r_j
  for (int e = threadIdx.x; e < 512; e += blockDim.x) {
    int code = 0;
    if (e < nin * nr) code = ((0 + e % nin) << 16) | (96 + e / nin);
    else if (e < nin * nr + nr) code = (192 << 16) | (96 + e - nin * nr);
    else if (e < nin * nr + nr + ns * nr) { const int f = e - nin * nr - nr; code = ((160 + f % ns) << 16) | (32 + f / ns); }
    else if (MLP && e < nin * nr + nr + ns * nr + ns) code = ((160 + (e - nin * nr - nr - ns * nr)) << 16) | 192;   // w_J[i]: (scale_i lambda_i) * 1
    else if (MLP && e < nw) {   // W_l[k, i]: delta_l[k] * a_l[i];  b_l[k]: delta_l[k] * 1
      int f = e - (nin * nr + nr + ns * nr + ns);
      for (int l = 0; l < ML; ++l) {
        const int din = W.mlp_dims[l], dout = W.mlp_dims[l + 1];
        const int dl_off = ADJ_FIXED + 32 + ms * (2 * ML + 1 + l), a_off = ADJ_FIXED + 32 + ms * l;
        if (f < din * dout) { code = ((dl_off + f % dout) << 16) | (a_off + f / dout); break; }
        f -= din * dout;
        if (f < dout) { code = ((dl_off + f) << 16) | 192; break; }
        f -= dout;
      }
    }
    sb.ent[e] = code;
  }
  __syncthreads();

  auto rec_ptr = [&](int step) -> double* {
    return step < P.cap_s ? rec_s + (size_t)step * stride : rec_g + (size_t)(step - P.cap_s) * stride;
  };
  // F4: the Flux chain of the state y (lane i holds y_i), lane k = neuron k, activations a_l and activation derivatives act'_l kept in
  // the per-warp arrays for the way back; returns the value of augmented input row `lane` (state row or MLP output).  Same
  // arithmetic as wide_mlp_aug / the oracle's mlp_eval.
  auto mlp_aug = [&](double y) -> double {
    __syncwarp();
    s_v[lane] = y;
    __syncwarp();
    double a = lane < W.mlp_dims[0] ? s_v[__ldg(W.mlp_in_idx + lane)] : 0.0;
    const double* w = s_mw;
#pragma unroll 1
    for (int l = 0; l < ML; ++l) {
      const int din = W.mlp_dims[l], dout = W.mlp_dims[l + 1];
      if (lane < din) A_(l)[lane] = a;
      __syncwarp();
      double sacc = 0.0;
      if (lane < dout) {
        const double* al = A_(l);
        const double* wp = w + lane;
#pragma unroll 4
        for (int i = 0; i < din; ++i, wp += dout) sacc = fma(*wp, al[i], sacc);
        sacc += *wp;
        // the activation and, for the way back, its derivative at the pre-activation (gelu' in NNlib's tanh form, softplus' = sigma, exp' = exp)
        if (l + 1 < ML) {
          const double x = sacc;
          const double th = 1.0 - 2.0 / (lean_exp(2.0 * (0.7978845608028654 * (x + 0.044715 * (x * x * x)))) + 1.0);
          SP_(l)[lane] = 0.5 * (1.0 + th) + 0.5 * x * (1.0 - th * th) * (0.7978845608028654 * (1.0 + 3.0 * 0.044715 * (x * x)));
          sacc = 0.5 * x * (1.0 + th);
        } else if (W.mlp_act_out == 0) {
          SP_(l)[lane] = 1.0 / (1.0 + lean_exp(-sacc));
          sacc = lean_log(1.0 + lean_exp(-fabs(sacc))) + (sacc > 0.0 ? sacc : 0.0);
        } else {
          sacc = lean_exp(sacc);
          SP_(l)[lane] = sacc;
        }
      }
      a = sacc;
      w += din * dout + dout;
    }
    if (lane < W.mlp_dims[ML]) A_(ML)[lane] = a;
    __syncwarp();
    if (lane >= nin) return 0.0;
    const int src = __ldg(W.aug_src + lane);
    return src >= 0 ? s_v[src] : A_(ML)[-1 - src];
  };
  // forward RHS: lane i holds y_i -> f_i (x, r left in shared memory)
  int tab_seg = 0;  // F2: segment hint of the T(t), P(t) lookup
  auto rhs = [&](double tt, double y) -> double {
    if (MLP) {
      const double v = mlp_aug(y);
      s_x[lane] = lane < nin ? lean_log(clampd(v, W.lb, W.ub)) : 0.0;
      __syncwarp();
      if (lane < nr) {
        double z = sb.w_b[lane];
#pragma unroll 2
        for (int i = 0; i < nin; ++i) z = fma(sb.w_inT[i][lane], s_x[i], z);
        s_r[lane] = lean_exp(z);
      }
      __syncwarp();
      double f = 0.0;
      if (isp) {
#pragma unroll 2
        for (int j = 0; j < nr; ++j) f = fma(sb.w_out[j][lane], s_r[j], f);
        if (W.w_J) f += __ldg(W.w_J + lane);
      }
      return f;
    }
    __syncwarp();
    double xi = 0.0, rho = 1.0;
    if (f2) {  // HyChem mass fractions (kernel_wide_solve.cuh::wide_rhs)
      const TabVal tv = wide_tab(W, tt, tab_seg);
      const double ymw = isp ? clampd(y, W.lb, W.ub) / my_mw : 0.0;
      const double S = wsum(ymw);
      rho = tv.P / (kGasRu * tv.T * S);
      if (isp) xi = lean_log(clampd(rho * ymw * 1e3, W.lb, W.ub));
      else if (lane == ns) xi = -1.0 / W.gas_R / tv.T;
      else if (lane == ns + 1) xi = lean_log(tv.T);
    } else if (isp) xi = lean_log(clampd(y, W.lb, W.ub));
    else if (W.kind == 1 && lane == ns) xi = -1.0 / (W.gas_R * y);
    s_x[lane] = xi;
    __syncwarp();
    if (lane < nr) {
      double z = sb.w_b[lane];
#pragma unroll 2
      for (int i = 0; i < nin; ++i) z = fma(sb.w_inT[i][lane], s_x[i], z);
      s_r[lane] = lean_exp(z);
    }
    __syncwarp();
    double f = 0.0;
    if (isp)
#pragma unroll 2
      for (int j = 0; j < nr; ++j) f = fma(sb.w_out[j][lane], s_r[j], f);
    if (f2) f = f / rho;
    return f;
  };
  // u_i(ts) from the recorded dense output of step `ir`
  auto dense_u = [&](int ir, double tsx) -> double {
    const double* r0 = rec_ptr(ir);
    const double th = (tsx - r0[0]) / r0[1];
    double acc = 0.0;
    if (lane < n) {
#pragma unroll
      for (int q7 = 0; q7 < 7; ++q7) {
        const double b = th * (tsc::R[q7][0] + th * (tsc::R[q7][1] + th * (tsc::R[q7][2] + th * tsc::R[q7][3])));
        acc = fma(b, r0[2 + (1 + q7) * n + lane], acc);
      }
      acc = fma(r0[1], acc, r0[2 + lane]);
    }
    return acc;
  };
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

    const long long src = in_idx ? __ldg(in_idx + traj) : traj;  // dataset row of the inputs (outputs stay at traj)
    double u = lane < n ? __ldg(u0 + src * n + lane) : 0.0;
    const double u_init = u;
    int nsave = W.n_save;
    double tend = W.t1;
    if (n_save_used) {
      int q = __ldg(n_save_used + traj);
      if (q > 0 && q <= W.n_save) { nsave = q; tend = __ldg(W.saveat + q - 1); }
    }
    const double t0 = W.t0, dtmax = tend - t0;
    const double dtmin = fmax(ulp_of(t0), ulp_of(tend));
    const size_t pbase = (size_t)traj * W.n_obs * W.n_save;
    const double* __restrict__ datat = data + (size_t)src * W.n_obs * W.n_save - pbase;  // datat + off reads row src

    // ================= forward: Tsit5 value solve, recording every accepted step =================
    int n_rhs = 0, n_acc = 0, n_rej = 0, n_back = 0, nrec = 0;
    double k[7];
    k[0] = rhs(t0, u); ++n_rhs;
    double dt;
    {
      const double sk = my_at + fabs(u) * my_rt;
      double a = 0.0, b = 0.0;
      if (lane < n) { a = u / sk; a *= a; b = k[0] / sk; b *= b; }
      const double d0 = sqrt(wsum(a) / n), d1 = sqrt(wsum(b) / n);
      double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
      dt0 = jmin(dt0, dtmax);
      const double f1 = rhs(t0 + dt0, fma(dt0, k[0], u)); ++n_rhs;
      double c = 0.0;
      if (lane < n) { c = (f1 - k[0]) / sk; c *= c; }
      const double d2 = sqrt(wsum(c) / n) / dt0;
      const double dm = jmax(d1, d2);
      const double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : lean_exp10(-(2.0 + lean_log10(dm)) * W.inv_order);
      dt = jmin(jmin(100.0 * dt0, dt1), dtmax);
    }
    double t = t0, qold = 1e-4, dt_last = 0.0;
    int isave = 0, ret = CRNN_RET_DEFAULT;
    long long iter = 0;
    while (isave < nsave && __ldg(W.saveat + isave) <= t0) ++isave;
    while (t < tend) {
      ++iter;
      if (dt != dt) { ret = CRNN_RET_DTNAN; break; }
      if (iter > W.maxiters) { ret = CRNN_RET_MAXITERS; break; }
      dt = jmin(dt, dtmax);
      dt = jmin(dt, tend - t);
      if (dt <= dtmin && tend - t > dtmin) { ret = CRNN_RET_DTLESSTHANMIN; break; }
      if (__any_sync(0xffffffffu, lane < n && u != u)) { ret = CRNN_RET_UNSTABLE; break; }
      double un = u;
#pragma unroll 1
      for (int s = 1; s < 7; ++s) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j)
          if (j < s) acc = (j == 0) ? tsc::A[s][0] * k[0] : fma(tsc::A[s][j], k[j], acc);
        un = fma(dt, acc, u);
        const double f = rhs(t + tsc::C[s] * dt, un); ++n_rhs;
#pragma unroll
        for (int j = 1; j < 7; ++j)
          if (j == s) k[j] = f;
      }
      double e = tsc::BT[0] * k[0];
#pragma unroll
      for (int j = 1; j < 7; ++j) e = fma(tsc::BT[j], k[j], e);
      const double EEst = wrms(dt * e, u, un);
      double q11, q;
      if (EEst == 0.0) { q11 = 0.0; q = W.inv_qmax; }
      else { q11 = lean_pow(EEst, W.beta1); q = jmax(W.inv_qmax, jmin(W.inv_qmin, q11 / lean_pow(qold, W.beta2) / W.gamma)); }
      dt_last = dt;
      if (EEst <= 1.0) {
        ++n_acc;
        qold = jmax(EEst, 1e-4);
        const double dtnew = dt / q;
        if (nrec >= P.cap_s + P.cap_g) { ret = CRNN_RET_MAXITERS; break; }  // record capacity (documented limit)
        double* r0 = rec_ptr(nrec);
        if (lane == 0) { r0[0] = t; r0[1] = dt; }
        if (lane < n) {
          r0[2 + lane] = u;
#pragma unroll
          for (int j = 0; j < 7; ++j) r0[2 + (1 + j) * n + lane] = k[j];
        }
        ++nrec;
        t = snap_t(t + dt, tend);
        while (isave < nsave && __ldg(W.saveat + isave) <= t) ++isave;
        u = un; k[0] = k[6];
        dt = jmin(dtnew, dtmax);
      } else {
        ++n_rej;
        dt = dt / jmin(W.inv_qmin, q11 / W.gamma);
      }
    }
    if (ret == CRNN_RET_DEFAULT) ret = CRNN_RET_SUCCESS;
    const double t_reached = t;
    const double cnt = (double)W.n_obs * (double)isave;
    __syncwarp();

    // ================= backward: adjoint sweep with jumps at the save times =================
    for (int e = lane; e < nw; e += 32) GW[e] = 0.0;
    double lam = 0.0, loss_acc = 0.0;
    double cur = t_reached, bdt = 0.0, bq = 1e-4;
    int ir = nrec - 1;
    bool have_dt = false;
    long long biter = 0;
    double fwd_t = __longlong_as_double(0x7ff8000000000000LL);   // time of the evaluation point whose forward by-products are cached (NaN: none)
    double K[7];
    // adjoint RHS at state component u_i (lane), multiplier value in s_lam: returns (J^T lambda)_i and leaves
    // x in s_x, r in s_r, g.*r in s_gr
    // F2 (oracle adj_rhs): s_lam holds lambda_i / rho (MW_i s_i are folded into w_out and P.scale), and the
    // density couples every species: + rr_l * sum_j (WS_j - 1) g_j r_j with WS_j = sum_i w_in[i,j] chiC_i.
    // `reuse`: this evaluation is at the SAME (t, u) as the previous one - stages 6 and 7 of a backward step share c = 1, and the
    // next step starts where this one ended - so the forward half (dense-output state, log, w_in mat-vec, exp) is not repeated:
    // x, r, chiC are still in shared memory, dx / rr / 1/rho / WS in the registers below.  Same bits: they are functions of (t, u).
    double c_dxi = 0.0, c_rrl = 0.0, c_inv_rho = 1.0, c_ws = 0.0;
    auto adj_rhs = [&](double tt, double ui, double li, bool reuse) -> double {
      if (MLP) {   // F4 (oracle adj_rhs_f4)
        if (!reuse) {
          const double v = mlp_aug(ui);
          double xi = 0.0, dxi = 0.0;
          if (lane < nin) {
            const double vc = clampd(v, W.lb, W.ub);
            xi = lean_log(vc);
            dxi = (v >= W.lb && v <= W.ub) ? 1.0 / vc : 0.0;
          }
          s_x[lane] = xi;
          c_dxi = dxi;
        }
        __syncwarp();
        s_lam[lane] = isp ? li : 0.0;
        s_sl[lane] = my_scale * (isp ? li : 0.0);
        __syncwarp();
        if (lane < nr) {
          double gs = 0.0, r;
#pragma unroll 2
          for (int i = 0; i < ns; ++i) gs = fma(sb.w_out[lane][i], s_lam[i], gs);
          if (!reuse) {
            double z = sb.w_b[lane];
#pragma unroll 2
            for (int i = 0; i < nin; ++i) z = fma(sb.w_inT[i][lane], s_x[i], z);
            r = lean_exp(z);
            s_r[lane] = r;
          } else r = s_r[lane];
          s_gr[lane] = gs * r;
        }
        __syncwarp();
        {   // cotangent of every augmented input row
          double sacc = 0.0;
          if (lane < nin)
#pragma unroll 2
            for (int j = 0; j < nr; ++j) sacc = fma(sb.w_inT[lane][j], s_gr[j], sacc);
          s_v[lane] = c_dxi * sacc;
        }
        __syncwarp();
        double dl = 0.0, dh = 0.0;
        for (int q = 0; q < nin; ++q) {   // state rows collect directly, hidden rows through the chain (ascending q, like the oracle)
          const int src = __ldg(W.aug_src + q);
          if (src == lane) dl += s_v[q];
          if (-1 - src == lane) dh += s_v[q];
        }
        if (lane < W.mlp_dims[ML]) DL_(ML - 1)[lane] = dh * SP_(ML - 1)[lane];
        int woff = P.mlp_np_raw;   // walks down the layers of the block's parameter copy
#pragma unroll 1
        for (int l = ML - 1; l >= 0; --l) {
          const int din = W.mlp_dims[l], dout = W.mlp_dims[l + 1];
          woff -= din * dout + dout;
          __syncwarp();
          double sacc = 0.0;
          if (lane < din) {
            const double* wl = s_mw + woff + dout * lane;
            const double* dl_l = DL_(l);
#pragma unroll 4
            for (int k = 0; k < dout; ++k) sacc = fma(wl[k], dl_l[k], sacc);
          }
          if (l > 0) {
            if (lane < din) DL_(l - 1)[lane] = sacc * SP_(l - 1)[lane];
          } else {
            __syncwarp();
            s_v[lane] = lane < din ? sacc : 0.0;   // d / d (MLP input i); s_v's cotangents have been consumed
            __syncwarp();
            for (int i = 0; i < din; ++i)
              if (__ldg(W.mlp_in_idx + i) == lane) dl += s_v[i];
          }
        }
        __syncwarp();
        return isp ? dl : 0.0;
      }
      __syncwarp();
      if (!reuse) {
        double xi = 0.0, dxi = 0.0, rrl = 0.0, chiC = 0.0, inv_rho = 1.0;
        if (f2) {
          const TabVal tv = wide_tab(W, tt, tab_seg);
          double Y = 1.0, chi = 0.0, ymw = 0.0;
          if (isp) { Y = clampd(ui, W.lb, W.ub); chi = (ui >= W.lb && ui <= W.ub) ? 1.0 : 0.0; ymw = Y / my_mw; }
          const double S = wsum(ymw);
          const double rho = tv.P / (kGasRu * tv.T * S);
          inv_rho = 1.0 / rho;
          if (isp) {
            const double C = rho * ymw * 1e3;
            chiC = (C >= W.lb && C <= W.ub) ? 1.0 : 0.0;
            xi = lean_log(clampd(C, W.lb, W.ub));
            dxi = chiC * chi / Y;
            rrl = -chi / (my_mw * S);
          } else if (lane == ns) xi = -1.0 / W.gas_R / tv.T;
          else if (lane == ns + 1) xi = lean_log(tv.T);
          s_chi[lane] = chiC;
        } else if (isp) {
          const double uc = clampd(ui, W.lb, W.ub);
          xi = lean_log(uc);
          dxi = (ui >= W.lb && ui <= W.ub) ? __drcp_rn(uc) : 0.0;
        } else if (W.kind == 1 && lane == ns) {
          xi = -1.0 / (W.gas_R * ui);
        }
        s_x[lane] = xi;
        c_dxi = dxi; c_rrl = rrl; c_inv_rho = inv_rho;
      }
      s_lam[lane] = isp ? li * c_inv_rho : 0.0;
      s_sl[lane] = my_scale * (isp ? li * c_inv_rho : 0.0);
      __syncwarp();
      double brk = 0.0;
      if (lane < nr) {
        double gs = 0.0, r;
#pragma unroll 2
        for (int i = 0; i < ns; ++i) gs = fma(sb.w_out[lane][i], s_lam[i], gs);
        if (!reuse) {
          double z = sb.w_b[lane];
#pragma unroll 2
          for (int i = 0; i < nin; ++i) z = fma(sb.w_inT[i][lane], s_x[i], z);
          r = lean_exp(z);
          s_r[lane] = r;
          if (f2) {
            double ws = 0.0;
            for (int i = 0; i < ns; ++i) ws = fma(sb.w_inT[i][lane], s_chi[i], ws);
            c_ws = ws;
          }
        } else r = s_r[lane];
        s_gr[lane] = gs * r;
        if (f2) brk = (c_ws - 1.0) * (gs * r);
      }
      if (f2) brk = wsum(brk);
      __syncwarp();
      double s = 0.0;
      if (isp)
#pragma unroll 2
        for (int j = 0; j < nr; ++j) s = fma(sb.w_inT[lane][j], s_gr[j], s);
      return fma(c_rrl, brk, c_dxi * s);
    };

    // loss term and pred of save column kk (value y_i in this lane); returns this lane's dL/du_i(t_k)
    auto save_term = [&](int kk, double y) -> double {
      double gret = 0.0;
      if (my_obs >= 0) {
        const double yc = clampd(y, W.pred_lo, W.pred_hi);
        const bool inside = (y >= W.pred_lo) && (y <= W.pred_hi);
        const size_t off = pbase + my_obs + (size_t)W.n_obs * kk;
        if (pred) pred[off] = yc;
        const double d = __ldg(datat + off);
        double diff, g;
        if (P.loss_kind == CRNN_LOSS_MAE_SCALED) { diff = d * my_iys - yc * my_iys; g = signbit(diff) ? my_iys : -my_iys; }
        else { diff = lean_log_nl(clampd(d, W.pred_lo, W.pred_hi)) - lean_log_nl(yc); g = (signbit(diff) ? 1.0 : -1.0) / yc; }  // one out-of-line copy of log
        loss_acc += fabs(diff);
        if (inside && isp) gret = g / cnt;
      }
      return gret;
    };
    // GW/GS-style accumulation of the three outer products left in shared memory by adj_rhs
    auto accum = [&](double* dst, double bw) {
#pragma unroll
      for (int q = 0; q < ADJ_MAX_ENT; ++q) {
        if (32 * q >= nw) break;  // uniform: the remaining entries are empty for every lane
        const int e = lane + 32 * q;
        if (e < nw) {
          const int code = sb.ent[e];
          const double val = wbase[code >> 16] * wbase[code & 0xffff];
          dst[e] = fma(bw, val, dst[e]);
        }
      }
    };

    if (P.discrete && isave > 0) {
      // ---- discrete adjoint: reverse-mode through the recorded steps and the dense-output saves ----
      double ubar = 0.0;
      int ks = isave - 1;
#pragma unroll 1
      for (int st = nrec - 1; st >= 0; --st) {
        const double* r0 = rec_ptr(st);
        const double tn = r0[0], h = r0[1];
        const double tnext = (st + 1 < nrec) ? rec_ptr(st + 1)[0] : t_reached;
        const double unext = (st + 1 < nrec) ? (lane < n ? rec_ptr(st + 1)[2 + lane] : 0.0) : u;
        double kbar[7], ubn = 0.0;
#pragma unroll
        for (int j = 0; j < 7; ++j) kbar[j] = 0.0;
        while (ks >= 0) {
          const double tsv = __ldg(W.saveat + ks);
          if (!(tsv > tn)) break;
          if (tsv == tnext) {
            ubar += save_term(ks, unext);
          } else {
            const double g = save_term(ks, dense_u(st, tsv));
            const double th = (tsv - tn) / h;
            ubn += g;
#pragma unroll
            for (int q7 = 0; q7 < 7; ++q7) {
              const double b = th * (tsc::R[q7][0] + th * (tsc::R[q7][1] + th * (tsc::R[q7][2] + th * tsc::R[q7][3])));
              kbar[q7] = fma(h * b, g, kbar[q7]);
            }
          }
          --ks;
        }
        double kk[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) kk[j] = lane < n ? r0[2 + (1 + j) * n + lane] : 0.0;
        const double un0 = lane < n ? r0[2 + lane] : 0.0;
#pragma unroll 1
        for (int j = 6; j >= 0; --j) {
          // stage point Y_j = u_n + h sum_{l<j} a_jl k_l  (j = 6: u_{n+1}, whose f is k7 of the dense output)
          double acc = 0.0;
#pragma unroll
          for (int l = 0; l < 6; ++l)
            if (l < j) acc = fma(tsc::A[j][l], kk[l], acc);
          double kb = 0.0;
#pragma unroll
          for (int l = 0; l < 7; ++l)
            if (l == j) kb = kbar[l];
          const double vj = adj_rhs(tn + tsc::C[j] * h, fma(h, acc, un0), kb, false); ++n_rhs;
          accum(GW, 1.0);
          if (j == 6) {
            ubar += vj;
#pragma unroll
            for (int l = 0; l < 6; ++l) kbar[l] = fma(h * tsc::A[6][l], ubar, kbar[l]);
            ubn += ubar;
          } else {
            ubn += vj;
#pragma unroll
            for (int l = 0; l < 6; ++l)
              if (l < j) kbar[l] = fma(h * tsc::A[j][l], vj, kbar[l]);
          }
        }
        ubar = ubn;
        ++n_back;
      }
      while (ks >= 0) { (void)save_term(ks, u_init); --ks; }  // saves at t0: loss only, u0 does not depend on p
    } else if (isave > 0) {
      for (int ks = isave; ks >= 0; --ks) {
        const double tlo = (ks > 0) ? __ldg(W.saveat + ks - 1) : t0;
        if (ks < isave) {
          while (cur > tlo) {
            if (++biter > W.maxiters) { ret = CRNN_RET_MAXITERS; break; }
            if (!have_dt) {  // Hairer initial step of the lambda system at `cur`
              while (ir > 0 && rec_ptr(ir)[0] >= cur) --ir;
              K[0] = adj_rhs(cur, dense_u(ir, cur), lam, false); ++n_rhs;
              fwd_t = __longlong_as_double(0x7ff8000000000000LL);   // (this probe picks its record differently from the stage loop)
              const double sk = my_at + fabs(lam) * my_rt;
              double a = 0.0, b = 0.0;
              if (lane < n) { a = lam / sk; a *= a; b = K[0] / sk; b *= b; }
              const double d0 = sqrt(wsum(a) / n), d1 = sqrt(wsum(b) / n);
              bdt = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
              have_dt = true;
            }
            const double h = jmin(bdt, cur - tlo);
            if (!(h > 0.0)) break;
            for (int e = lane; e < nw; e += 32) GS[e] = 0.0;
            double ln = lam;
#pragma unroll 1
            for (int s = 0; s < 7; ++s) {
              double y = lam;
              if (s > 0) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < 6; ++j)
                  if (j < s) acc = (j == 0) ? tsc::A[s][0] * K[0] : fma(tsc::A[s][j], K[j], acc);
                y = fma(h, acc, lam);
              }
              if (s == 6) ln = y;
              const double tsx = cur - tsc::C[s] * h;
              const bool reuse = (tsx == fwd_t);   // same evaluation point as the previous call (never true for NaN)
              double uu = 0.0;
              if (!reuse) {
                while (ir > 0 && rec_ptr(ir)[0] > tsx) --ir;
                while (ir < nrec - 1 && rec_ptr(ir)[0] + rec_ptr(ir)[1] < tsx) ++ir;
                uu = dense_u(ir, tsx);
                fwd_t = tsx;
              }
              const double f = adj_rhs(tsx, uu, y, reuse); ++n_rhs;
#pragma unroll
              for (int j = 0; j < 7; ++j)
                if (j == s) K[j] = f;
              if (s < 6) accum(GS, tsc::A[6][s]);  // quadrature: b-weights of the 5th-order solution (b7 = 0)
            }
            double e = tsc::BT[0] * K[0];
#pragma unroll
            for (int j = 1; j < 7; ++j) e = fma(tsc::BT[j], K[j], e);
            const double EEst = wrms(h * e, lam, ln);
            double q11, q;
            if (EEst == 0.0) { q11 = 0.0; q = W.inv_qmax; }
            else { q11 = lean_pow(EEst, W.beta1); q = jmax(W.inv_qmax, jmin(W.inv_qmin, q11 / lean_pow(bq, W.beta2) / W.gamma)); }
            if (EEst <= 1.0) {
              bq = jmax(EEst, 1e-4);
              for (int ee = lane; ee < nw; ee += 32) GW[ee] = fma(h, GS[ee], GW[ee]);
              lam = ln;
              const double nxt = cur - h;
              cur = (fabs(nxt - tlo) < 100.0 * 2.220446049250313e-16 * fmax(fabs(cur), fabs(tlo))) ? tlo : nxt;
              bdt = jmin(h / q, dtmax);
              ++n_back;
            } else {
              bdt = h / jmin(W.inv_qmin, q11 / W.gamma);
            }
          }
        }
        if (ks > 0) {
          // jump at save ks-1: lambda += dL/du(t_k); loss term and pred fused here
          const int kk = ks - 1;
          const double tsv = __ldg(W.saveat + kk);
          int jr = ir;
          while (jr > 0 && rec_ptr(jr)[0] >= tsv) --jr;
          while (jr < nrec - 1 && rec_ptr(jr)[0] + rec_ptr(jr)[1] < tsv) ++jr;
          double y;
          if (tsv <= t0) y = u_init;                                        // t0 itself is saved exactly
          else if (kk == isave - 1 && tsv == t_reached) y = u;              // last save = end state, exactly
          else y = dense_u(jr, tsv);
          lam += save_term(kk, y);
          if (tsv < cur) cur = tsv;
        }
      }
    }
    const double ltot = wsum(loss_acc);
    __syncwarp();
    for (int e = lane; e < nw; e += 32) gw_each[(size_t)traj * nw + e] = isave > 0 ? GW[e] : 0.0;
    if (lane == 0) {
      loss[traj] = isave > 0 ? ltot / cnt : __longlong_as_double(0x7ff8000000000000LL);
      if (n_saved) n_saved[traj] = isave;
      if (retcode) retcode[traj] = ret;
      if (stats) {
        crnn_stats s;
        s.n_accept = n_acc; s.n_reject = n_rej; s.n_rhs = n_rhs; s.n_jac = n_back;
        s.t_reached = t_reached; s.dt_last = dt_last;
        stats[traj] = s;
      }
    }
    if (pred && my_obs >= 0)
      for (int ks = isave; ks < W.n_save; ++ks) pred[pbase + my_obs + (size_t)W.n_obs * ks] = 0.0;
    __syncwarp();
  }
}

// grad_sum[c] = sum_w dW_dp[w, c] * gw_sum[w]   (dW_dp col-major [nw, np], with out_scale NOT folded: G_out carries it)
static __global__ void k_seed_contract(const double* __restrict__ seed, const double* __restrict__ gw_sum, int nw, int np,
                                       double* __restrict__ grad_sum) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < np; c += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s = fma(seed[w + (size_t)nw * c], gw_sum[w], s);
    grad_sum[c] = s;
  }
}

}  // namespace crnn
