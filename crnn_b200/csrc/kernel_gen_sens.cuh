// kernel_gen_sens.cuh — the GENERIC forward-sensitivity path: ForwardDiff.gradient(loss_neuralode) for ANY CRNN of the
// reference (runtime dimensions n_state, n_in, n_reac <= 32; RHS flavours F0, F1, F2), any number of parameters up to
// 255, and the three steppers the scripts differentiate through:
//   Tsit5                        case1/case1.jl:28,195-199 ; case3/case3.jl:29,265-270
//   Rosenbrock23(autodiff=true)  robertson/rober_crnn.jl:33,219   (nested-dual Jacobian: the D^2 f terms below)
//   AutoTsit5(Rosenbrock23())    case2/case2.jl:26,195 ; HyChem/crnn_pyrolysis_mass.jl:29,201  (what those scripts run)
// The dimension-specialised warp-per-trajectory kernels (kernel_tsit5_sens.cuh, kernel_rosenbrock23_sens.cuh) stay the
// fast path for the instantiated configurations; this kernel serves everything they refuse: gradients through the
// composite algorithm, stiff gradients of F2 and of models with more than 6 species, pruned / resized CRNNs.
//
// One BLOCK owns one trajectory (the north star's "one block for HyChem-sized state").  Warp 0 carries the VALUE
// column cooperatively, lane i = state component i (log / exp issued once per evaluation for the whole state, analytic
// Jacobian + cooperative LU of wide_common.cuh).  Every THREAD of the block owns one dual column c = tid + 1 and applies
// the linearised RHS to it matrix-free, reading the value path's by-products (x, dx, r, ...) as shared-memory
// broadcasts; its state and stage vectors live in shared memory as [slot][component][column] (conflict-free).
// Rosenbrock23's dual linear solves reuse the value path's LU: W kdot_c = rhsdot_c + gamma dt (D^2 f[(S_c, dW_c), (k, tau)]),
// the mixed second derivative restated in closed form for every flavour (oracle/crnn_oracle.c::djac_vec), its time
// component tau carrying the dual part of df/dt for the non-autonomous F2.
// The partials enter the error norm like DiffEqBase's norm over Dual arrays (block-wide row sums), the loss and its
// gradient are fused at the save points.  It mirrors oracle/crnn_oracle.c::solve_one stage by stage.
#pragma once
#include "crnn_dev.cuh"
#include "wide_common.cuh"
#include "kernel_tsit5_sens.cuh"     // R1Desc
#include "kernel_tsit5_adjoint.cuh"  // tsc:: tableau in constant memory

namespace crnn {

struct GenP {
  WideP w;                 // dimensions, tolerances, controller, weights (out_scale and, for F2, MW folded into w_out), tables
  const double* inv_ys;    // device [n] 1/yscale per state row (1 where unused)
  const double* seed_rows; // device [2*nr][cols]: a_j = dW_in[i_in, j] then b_j = db_j of the column a thread owns
  const R1Desc* desc;      // device [cols]
  double norm_cnt;         // divisor of the norms: n*(1+np), or n
  int np, cols, loss_kind, incl_sens;
  // observable post-map y = sum_j w_obs[j] r_j (heat release, Cathode/src/network.jl:82-91,121): seed_rows then has
  // 3*nr rows, the last nr being d w_obs_j / d p_c
  const double* w_obs;     // device [nr] or NULL
  // parameter-batched mode (crnn_loss_grad_particles): trajectory = experiment e + n_exp * particle p
  const double* pw;        // device, per particle: w_inT [nin][32] | w_b [nr] | w_out [ns x nr] | w_obs [nr]
  long long pw_stride, seed_stride;   // doubles per particle in pw / seed_rows (desc: cols per particle)
  int n_part, n_exp, tab_per_exp, pad;
};

struct alignas(16) GenPoint {  // by-products of one value-path evaluation, broadcast to the column threads
  double x[KW_MAXN + 2];    // inputs of W_in
  double dx[KW_MAXN];       // F0/F1: d x_i / d u_i ; F2: chi_i / Y_i
  double d2[KW_MAXN];       // F0/F1: -d2 x_i / d u_i^2 ; F2: chi_i / Y_i^2
  double chiC[KW_MAXN];     // F2: 1 inside the clamp of C_i ; F0/F1: 1
  double chimw[KW_MAXN];    // F2: chi_i / MW_i
  double r[KW_MAXN];
  double wdot[KW_MAXN];     // sum_j w_out[i,j] r_j (scaled)
  double inv_rho, rho, Ssum, T, Pr, Td, Pd, pad;
};

struct alignas(16) GenDir2 {   // the second direction (v, tau) of D^2 f, prepared by warp 0
  double x2[KW_MAXN + 2], vd2[KW_MAXN], z2[KW_MAXN], r2[KW_MAXN], w2[KW_MAXN];
  double lr2, s2;
};

struct alignas(16) GenShared {
  WideBlock sb;
  WideWarp ww;
  GenPoint cur, base;        // at the last evaluation / at u_n
  GenDir2 d2;
  double red[3 * KW_MAXN][8];
  double rows[3 * KW_MAXN];
  GenPoint tmp;              // at a save point (observable post-map)
  double wobs[KW_MAXN];
  double g[KW_MAXN];         // d loss / d yhat_i at the current save point
  double bcast[8];           // scalars broadcast from thread 0: EEst, dt, ...
  long long traj;
  int ibcast[6];
};

// f(y, t) of the value column (lane i holds y_i), filling the broadcast by-products `pc` and this lane's WideAux.
template <bool F2>
__device__ __forceinline__ double gen_rhs(const WideP& P, const WideBlock& sb, GenPoint& pc, int lane, double mw,
                                          double t, double y, WideAux& a, int& seg, size_t toff = 0) {
  const int ns = P.ns, nin = P.nin, nr = P.nr;
  const bool isp = lane < ns;
  __syncwarp();
  double xi = 0.0, dxi = 0.0, d2i = 0.0, chiC = 1.0, chimw = 0.0, rho = 1.0;
  a.dx = 0.0; a.rr = 0.0; a.chiC = 0.0; a.inv_rho = 1.0;
  if (F2) {
    // F2: HyChem mass fractions with the density map; F5 (Cathode/src/network.jl:68-80): same inputs without it
    const bool dens = (P.kind == CRNN_RHS_F2_MASSFRAC_TP);
    TabVal tv = wide_tab(P, t, seg, toff);
    if (!dens) { tv.P = 1.0; tv.Pd = 0.0; }
    double Y = 1.0, chi = 0.0, ymw = 0.0;
    if (isp) { Y = clampd(y, P.lb, P.ub); chi = (y >= P.lb && y <= P.ub) ? 1.0 : 0.0; ymw = dens ? Y / mw : 0.0; }
    const double S = dens ? wsum(ymw) : 1.0;
    rho = dens ? tv.P / (kGasRu * tv.T * S) : 1.0;
    chiC = 0.0;
    if (isp) {
      const double C = dens ? rho * ymw * 1e3 : Y;
      chiC = (C >= P.lb && C <= P.ub) ? 1.0 : 0.0;
      xi = lean_log(clampd(C, P.lb, P.ub));
      // chi, chiC are 0 or 1: x * (1 / Y) has the bits of x / Y, and a zero numerator no longer takes the division's slow path
      // (a consumed species sits on the clamp: chi = 0 on every evaluation of the Cathode model)
      const double rY = 1.0 / Y;
      dxi = chi * rY; d2i = chi * (1.0 / (Y * Y)); chimw = dens ? chi * (1.0 / mw) : 0.0;
      a.chiC = chiC; a.dx = (chiC * chi) * rY; a.rr = dens ? -chi * (1.0 / (mw * S)) : 0.0;
    } else if (lane == ns) {
      xi = -1.0 / P.gas_R / tv.T;
    } else if (lane == ns + 1) {
      xi = lean_log(tv.T);
    }
    a.inv_rho = 1.0 / rho;
    if (lane == 0) { pc.inv_rho = a.inv_rho; pc.rho = rho; pc.Ssum = S; pc.T = tv.T; pc.Pr = tv.P; pc.Td = tv.Td; pc.Pd = tv.Pd; }
  } else {
    if (isp) {
      const double uc = clampd(y, P.lb, P.ub);
      const bool inside = (y >= P.lb) && (y <= P.ub);
      xi = lean_log(uc);
      dxi = inside ? __drcp_rn(uc) : 0.0;
      d2i = inside ? 1.0 / (uc * uc) : 0.0;
    } else if (P.kind == 1 && lane == ns) {
      xi = -1.0 / (P.gas_R * y);
      dxi = 1.0 / (P.gas_R * y * y);
      d2i = 2.0 / (P.gas_R * y * y * y);
    }
    a.dx = dxi;
    if (lane == 0) { pc.inv_rho = 1.0; pc.rho = 1.0; pc.Ssum = 1.0; pc.Td = 0.0; pc.Pd = 0.0; pc.T = 1.0; pc.Pr = 1.0; }
  }
  if (lane < nin) pc.x[lane] = xi;
  pc.dx[lane] = dxi; pc.d2[lane] = d2i; pc.chiC[lane] = chiC; pc.chimw[lane] = chimw;
  __syncwarp();
  if (lane < nr) {
    double z = sb.w_b[lane];
#pragma unroll 2
    for (int i = 0; i < nin; ++i) z = fma(sb.w_inT[i][lane], pc.x[i], z);
    pc.r[lane] = lean_exp(z);
  }
  __syncwarp();
  double f = 0.0;
  if (isp) {
#pragma unroll 2
    for (int j = 0; j < nr; ++j) f = fma(sb.w_out[j][lane], pc.r[j], f);
  }
  a.wdot = f;
  pc.wdot[lane] = f;
  if (F2) f = f / rho;
  return f;
}

// warp 0: the arrays of the second direction (v lane-distributed, time component tau) at the point `pc`
template <bool F2>
__device__ __forceinline__ void gen_dir2(const WideP& P, const WideBlock& sb, const GenPoint& pc, GenDir2& d, int lane,
                                         double v, double tau) {
  const int ns = P.ns, nin = P.nin, nr = P.nr, nsd = F2 ? ns : P.n;
  __syncwarp();
  double lr2 = 0.0, s2 = 0.0;
  if (F2 && P.kind == CRNN_RHS_F2_MASSFRAC_TP) {
    s2 = wsum(lane < ns ? pc.chimw[lane] * v : 0.0);
    lr2 = tau * (pc.Pd / pc.Pr - pc.Td / pc.T) - s2 / pc.Ssum;
  }
  double x2 = 0.0;
  if (lane < nsd) x2 = pc.chiC[lane] * (lr2 + v * pc.dx[lane]);
  else if (F2 && lane == ns) x2 = tau * pc.Td / (P.gas_R * pc.T * pc.T);
  else if (F2 && lane == ns + 1) x2 = tau * pc.Td / pc.T;
  if (lane < nin) d.x2[lane] = x2;
  d.vd2[lane] = lane < nsd ? v * pc.d2[lane] : 0.0;
  if (lane == 0) { d.lr2 = lr2; d.s2 = s2; }
  __syncwarp();
  if (lane < nr) {
    double z = 0.0;
#pragma unroll 2
    for (int i = 0; i < nin; ++i) z = fma(sb.w_inT[i][lane], d.x2[i], z);
    d.z2[lane] = z; d.r2[lane] = pc.r[lane] * z;
  }
  __syncwarp();
  double w2 = 0.0;
  if (lane < ns) {
#pragma unroll 2
    for (int j = 0; j < nr; ++j) w2 = fma(sb.w_out[j][lane], d.r2[j], w2);
  }
  d.w2[lane] = w2;
  __syncwarp();
}

// One column: dst <- f'[(a, dW_c)] (D2 = false) or D^2 f[(a, dW_c), (v, tau)] (D2 = true) at the point pc.
// a / dst / k0 point at [component 0] of this thread's column in a slot; components are `cs` doubles apart.
template <bool F2, bool D2>
__device__ __forceinline__ void col_apply(const WideP& P, const WideBlock& sb, const GenPoint& pc, const GenDir2& d,
                                          const double* __restrict__ srow, int cols, int tid, const R1Desc& ds,
                                          const double* a, double* dst, const double* k0, int cs) {
  const int n = P.n, ns = P.ns, nr = P.nr, nsd = F2 ? ns : n;
  double lr1 = 0.0, lr2 = 0.0, lr12 = 0.0;
  if (F2) {
    double s1 = 0.0;
    if (P.kind == CRNN_RHS_F2_MASSFRAC_TP) {   // F5 has no density map: chimw = 0, every term below is zero
      for (int l = 0; l < ns; ++l) s1 = fma(pc.chimw[l], a[l * cs], s1);
      lr1 = -s1 / pc.Ssum;
      if (D2) { lr2 = d.lr2; lr12 = s1 * d.s2 / (pc.Ssum * pc.Ssum); }
    }
  }
  for (int i = 0; i < ns; ++i) dst[i * cs] = 0.0;
  const double xin = pc.x[ds.i_in], x2in = D2 ? d.x2[ds.i_in] : 0.0;
  for (int j0 = 0; j0 < nr; j0 += 8) {
    double z1[8], z12[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int j = j0 + q;
      double sa = 0.0, sbb = 0.0;
      if (j < nr) { sa = __ldg(srow + (size_t)j * cols + tid); sbb = __ldg(srow + (size_t)(nr + j) * cols + tid); }
      z1[q] = fma(sa, xin, sbb);
      z12[q] = D2 ? sa * x2in : 0.0;
    }
    for (int i = 0; i < nsd; ++i) {
      const double ai = a[i * cs];
      const double x1 = pc.chiC[i] * (lr1 + ai * pc.dx[i]);
      const double x12 = D2 ? pc.chiC[i] * (lr12 - ai * d.vd2[i]) : 0.0;
      const double2* w = reinterpret_cast<const double2*>(&sb.w_inT[i][j0]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double2 v = w[q];
        z1[2 * q] = fma(v.x, x1, z1[2 * q]); z1[2 * q + 1] = fma(v.y, x1, z1[2 * q + 1]);
        if (D2) { z12[2 * q] = fma(v.x, x12, z12[2 * q]); z12[2 * q + 1] = fma(v.y, x12, z12[2 * q + 1]); }
      }
    }
    double val[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int j = j0 + q;
      val[q] = 0.0;
      if (j < nr) val[q] = D2 ? pc.r[j] * fma(z1[q], d.z2[j], z12[q]) : pc.r[j] * z1[q];
    }
    for (int i = 0; i < ns; ++i) {
      double acc = dst[i * cs];
#pragma unroll
      for (int q = 0; q < 8; ++q) acc = fma(sb.w_out[(j0 + q) & (KW_MAXN - 1)][i], val[q], acc);  // val = 0 beyond nr
      dst[i * cs] = acc;
    }
  }
  if (ds.o != 0.0) dst[ds.i_out * cs] = fma(ds.o, D2 ? d.r2[ds.j_out] : pc.r[ds.j_out], dst[ds.i_out * cs]);
  if (F2) {
    for (int i = 0; i < ns; ++i) {
      const double w = dst[i * cs];
      if (!D2) dst[i * cs] = (w - pc.wdot[i] * lr1) * pc.inv_rho;
      else {
        const double w1 = fma(k0[i * cs], pc.rho, pc.wdot[i] * lr1);  // wdot'_i from f'_i = (wdot' - wdot lr') / rho
        dst[i * cs] = ((lr1 * lr2 - lr12) * pc.wdot[i] - lr1 * d.w2[i] - lr2 * w1 + w) * pc.inv_rho;
      }
    }
  }
  for (int i = ns; i < n; ++i) dst[i * cs] = 0.0;
}

// b <- W^{-1} b for one column with the LU of the value path (ww.A / piv / dinv); loops ordered like oracle lu_solve
__device__ __forceinline__ void col_lusolve(const WideWarp& ww, int ns, double* b, int cs) {
  for (int k = 0; k < ns; ++k) {
    const int p = ww.piv[k];
    if (p != k) { const double t = b[k * cs]; b[k * cs] = b[p * cs]; b[p * cs] = t; }
  }
  for (int i = 1; i < ns; ++i) {
    double s = b[i * cs];
    for (int j = 0; j < i; ++j) s = fma(-ww.A[i][j], b[j * cs], s);
    b[i * cs] = s;
  }
  for (int i = ns - 1; i >= 0; --i) {
    double s = b[i * cs];
    for (int j = ns - 1; j > i; --j) s = fma(-ww.A[i][j], b[j * cs], s);
    b[i * cs] = s * ww.dinv[i];
  }
}

// One column at a save point: d/d eps of the observable y = sum_j w_obs[j] r_j along (a, dW_c), a_i = colval(i):
// sum_j ( w_obs[j] r_j z'_j + d w_obs_j r_j ),  z'_j = a_j x[i_in] + b_j + sum_i w_in[i,j] chiC_i (lr' + a_i dx_i).
template <bool F2, class CV>
__device__ __forceinline__ double col_obs(const WideP& P, const WideBlock& sb, const GenPoint& pc, const double* wobs,
                                          const double* __restrict__ srow, int cols, int tid, const R1Desc& ds, CV colval) {
  const int n = P.n, ns = P.ns, nr = P.nr, nsd = F2 ? ns : n;
  double lr1 = 0.0;
  if (F2 && P.kind == CRNN_RHS_F2_MASSFRAC_TP) {
    double s1 = 0.0;
    for (int l = 0; l < ns; ++l) s1 = fma(pc.chimw[l], colval(l), s1);
    lr1 = -s1 / pc.Ssum;
  }
  const double xin = pc.x[ds.i_in];
  double dy = 0.0;
  for (int j0 = 0; j0 < nr; j0 += 8) {
    double z1[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int j = j0 + q;
      double sa = 0.0, sbb = 0.0;
      if (j < nr) { sa = __ldg(srow + (size_t)j * cols + tid); sbb = __ldg(srow + (size_t)(nr + j) * cols + tid); }
      z1[q] = fma(sa, xin, sbb);
    }
    for (int i = 0; i < nsd; ++i) {
      const double x1 = pc.chiC[i] * (lr1 + colval(i) * pc.dx[i]);
      const double2* w = reinterpret_cast<const double2*>(&sb.w_inT[i][j0]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double2 v = w[q];
        z1[2 * q] = fma(v.x, x1, z1[2 * q]); z1[2 * q + 1] = fma(v.y, x1, z1[2 * q + 1]);
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int j = j0 + q;
      if (j < nr) dy = fma(pc.r[j], fma(wobs[j], z1[q], __ldg(srow + (size_t)(2 * nr + j) * cols + tid)), dy);
    }
  }
  return dy;
}

// STIFF: the stiff stepper of this instantiation - 0 Rosenbrock23 (alg ROSENBROCK23 / AUTO_TSIT5_ROS23), 1 TRBDF2 (alg TRBDF2 /
// AUTO_TSIT5_TRBDF2: the Cathode scripts' training path, Cathode/src/network.jl:102 + src_333/network.jl:232).
template <bool F2, int STIFF = 0>
__global__ void __launch_bounds__(256, 1)
k_gen_sens(const __grid_constant__ GenP G, const double* __restrict__ u0, const int* __restrict__ n_save_used,
           long long ntraj, const double* __restrict__ data, double* __restrict__ loss, double* __restrict__ grad_each,
           double* __restrict__ pred, int* __restrict__ n_saved, int* __restrict__ retcode,
           crnn_stats* __restrict__ stats, unsigned long long* __restrict__ queue, const long long* __restrict__ in_idx,
           const long long* __restrict__ sel = nullptr, const unsigned int* __restrict__ sel_count = nullptr) {
  // sel / sel_count (or NULL): only the trajectories sel[0 .. *sel_count) of this call are solved - the ones a specialised
  // Tsit5 kernel's AutoSwitch monitor handed over (kernel_tsit5_sens.cuh, AUTO); the count lives in device memory, so
  // no host synchronisation sits between the two launches
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GenShared& S = *reinterpret_cast<GenShared*>(smem_raw);
  double* const colbase = reinterpret_cast<double*>(smem_raw + sizeof(GenShared));
  const WideP& P = G.w;
  WideBlock& sb = S.sb;
  WideWarp& ww = S.ww;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int n = P.n, ns = P.ns, nin = P.nin, nr = P.nr, np = G.np, cols = G.cols;
  const int cs = cols;                       // distance between components of a column
  const bool active = tid < np;              // this thread owns dual column tid + 1
  const bool w0 = warp == 0;
  // column slots: [slot][component][column]
  enum { SL_U = 0, SL_Y = 1, SL_K = 2 };     // K0..K6 = slots 2..8
  auto slot = [&](int s) -> double* { return colbase + ((size_t)s * n) * cols + tid; };

  const bool has_obs = G.w_obs != nullptr || (G.n_part > 0 && G.pw_stride > (long long)nin * KW_MAXN + nr + (long long)ns * nr);
  auto load_weights = [&](const double* winT, const double* wb, const double* wout, const double* wobs) {
    for (int q = tid; q < KW_MAXN * KW_MAXN; q += blockDim.x) {
      const int i = q / KW_MAXN, j = q % KW_MAXN;
      sb.w_inT[i][j] = (i < nin && j < nr) ? winT[i * KW_MAXN + j] : 0.0;
      sb.w_inJ[j][i] = sb.w_inT[i][j];
      sb.w_out[i][j] = (i < nr && j < ns) ? wout[j + ns * i] : 0.0;  // [reaction][species]
    }
    for (int q = tid; q < KW_MAXN; q += blockDim.x) {
      sb.w_b[q] = q < nr ? wb[q] : 0.0;
      S.wobs[q] = (wobs && q < nr) ? wobs[q] : 0.0;
    }
    __syncthreads();
  };
  if (G.n_part == 0) load_weights(P.w_inT, P.w_b, P.w_out, G.w_obs);

  const bool autosw = (P.alg == CRNN_ALG_AUTO_TSIT5_ROS23 || P.alg == CRNN_ALG_AUTO_TSIT5_TRBDF2);
  const bool incl = G.incl_sens != 0;
  const double my_at = lane < n ? P.abstol[lane] : 1.0, my_rt = lane < n ? P.reltol[lane] : 0.0;
  const int my_obs = (w0 && lane < n) ? P.row2obs[lane] : -1;
  const double my_mw = (F2 && lane < ns) ? __ldg(P.mw + lane) : 1.0;
  const double my_iys = (w0 && lane < n) ? __ldg(G.inv_ys + lane) : 1.0;
  R1Desc ds{}; ds.o = 0.0;
  if (active && G.n_part == 0) ds = G.desc[tid];
  const double* srow = G.seed_rows;
  double* const kk = &ww.k[0][lane];
#define KS(s) kk[(s) * KW_MAXN]
#define CKS(s) slot(SL_K + (s))

  // block-wide sums over the dual columns of f(row) for `nrow` rows -> S.rows[row]
  auto row_sums = [&](int nrow, auto f) {
    for (int rw = 0; rw < nrow; ++rw) {
      const double v = wsum(active ? f(rw) : 0.0);
      if (lane == 0) S.red[rw][warp] = v;
    }
    __syncthreads();
    if (tid < nrow) {
      double s = 0.0;
      for (int w = 0; w < nwarps; ++w) s += S.red[tid][w];
      S.rows[tid] = s;
    }
    __syncthreads();
  };
  // dual-aware norm: sqrt( sum_i (vv_i^2 + sum_c colv_c,i^2) / sk_i^2 / norm_cnt ), sk from |u0| only (initial step)
  // or from max(|u|,|un|) (error estimate); result broadcast in S.bcast[0]

  while (true) {
    if (tid == 0) S.traj = (long long)atomicAdd(queue, 1ull);
    __syncthreads();
    long long traj = S.traj;
    __syncthreads();
    if (sel) {
      if (traj >= (long long)*sel_count) break;
      traj = sel[traj];
    } else if (traj >= ntraj) break;
    long long src = in_idx ? __ldg(in_idx + traj) : traj;
    size_t toff = 0;
    if (G.n_part > 0) {   // trajectory = experiment e + n_exp * particle p: this particle's weights and seed columns
      const long long pp = traj / G.n_exp;
      src = traj - pp * G.n_exp;
      const double* pwp = G.pw + pp * G.pw_stride;
      load_weights(pwp, pwp + (size_t)nin * KW_MAXN, pwp + (size_t)nin * KW_MAXN + nr,
                   has_obs ? pwp + (size_t)nin * KW_MAXN + nr + (size_t)ns * nr : nullptr);
      srow = G.seed_rows + pp * G.seed_stride;
      if (active) ds = G.desc[pp * cols + tid];
    }
    if (G.tab_per_exp) toff = (size_t)src * P.n_tab;

    double u = (w0 && lane < n) ? __ldg(u0 + src * n + lane) : 0.0;
    const double u_init = u;
    int nsave = P.n_save;
    double tend = P.t1;
    if (n_save_used) {
      const int q = __ldg(n_save_used + traj);
      if (q > 0 && q <= P.n_save) { nsave = q; tend = __ldg(P.saveat + q - 1); }
    }
    const double t0 = P.t0, dtmax = tend - t0;
    const double dtmin = fmax(ulp_of(t0), ulp_of(tend));
    const size_t pbase = (size_t)traj * P.n_obs * P.n_save;
    const double* datat = data + (size_t)src * P.n_obs * P.n_save;

    int n_rhs = 0, n_acc = 0, n_rej = 0, n_jac = 0, tab_seg = 0;
    double Gc = 0.0, loss_acc = 0.0;
    WideAux a0, as;
    if (active) for (int i = 0; i < n; ++i) slot(SL_U)[i * cs] = 0.0;   // sensitivities of u0 are zero
    // ---- f0 on all columns ----
    if (w0) { KS(0) = gen_rhs<F2>(P, sb, S.cur, lane, my_mw, t0, u, as, tab_seg, toff); a0 = as; }
    ++n_rhs;
    __syncthreads();
    if (active) col_apply<F2, false>(P, sb, S.cur, S.d2, srow, cols, tid, ds, slot(SL_U), CKS(0), nullptr, cs);
    for (int q = tid; q < (int)(sizeof(GenPoint) / sizeof(double)); q += blockDim.x)
      reinterpret_cast<double*>(&S.base)[q] = reinterpret_cast<const double*>(&S.cur)[q];
    __syncthreads();
    // ---- initial step (Hairer-Wanner with the dual-aware norms) ----
    double dt;
    {
      // d0 = |u0/sk|, d1 = |f0/sk|: the partials of u0 are zero, those of f0 are K0_c
      if (incl) row_sums(n, [&](int i) { const double v = CKS(0)[i * cs]; return v * v; });
      double d0 = 0.0, d1 = 0.0, dt0 = 0.0;
      const double sk = my_at + fabs(u_init) * my_rt;
      if (w0) {
        double a = 0.0, b = 0.0;
        if (lane < n) { a = u_init / sk; a *= a; const double f0 = KS(0); b = (f0 * f0 + (incl ? S.rows[lane] : 0.0)) / (sk * sk); }
        d0 = sqrt(wsum(a) / G.norm_cnt); d1 = sqrt(wsum(b) / G.norm_cnt);
        dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
        dt0 = jmin(dt0, dtmax);
        if (lane == 0) S.bcast[0] = dt0;
      }
      __syncthreads();
      dt0 = S.bcast[0];
      // f(u0 + dt0 f0, t0 + dt0) on all columns: Y = U + dt0 K0
      if (active) for (int i = 0; i < n; ++i) slot(SL_Y)[i * cs] = fma(dt0, CKS(0)[i * cs], slot(SL_U)[i * cs]);
      double f1p = 0.0;
      if (w0) f1p = gen_rhs<F2>(P, sb, S.cur, lane, my_mw, t0 + dt0, fma(dt0, KS(0), u), as, tab_seg, toff);
      ++n_rhs;
      __syncthreads();
      if (active) col_apply<F2, false>(P, sb, S.cur, S.d2, srow, cols, tid, ds, slot(SL_Y), CKS(1), nullptr, cs);
      if (incl) row_sums(n, [&](int i) { const double v = CKS(1)[i * cs] - CKS(0)[i * cs]; return v * v; });
      else __syncthreads();
      if (w0) {
        double c = 0.0;
        if (lane < n) { const double dv = f1p - KS(0); c = (dv * dv + (incl ? S.rows[lane] : 0.0)) / (sk * sk); }
        const double d2n = sqrt(wsum(c) / G.norm_cnt) / dt0;
        const double dm = jmax(d1, d2n);
        const double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : lean_exp10(-(2.0 + lean_log10(dm)) * P.inv_order);
        if (lane == 0) S.bcast[0] = jmin(jmin(100.0 * dt0, dt1), dtmax);
      }
      __syncthreads();
      dt = S.bcast[0];
      __syncthreads();
    }
    double t = t0, qold = 1e-4, dt_last = 0.0, eigen_est = 0.0;
    int isave = 0, ret = CRNN_RET_DEFAULT, sw_count = 0;
    bool rosen = (P.alg == CRNN_ALG_ROSENBROCK23 || P.alg == CRNN_ALG_TRBDF2);   // "the stiff stepper is active"
    double eta_old = 1.0;   // TRBDF2: the Newton solver's eta, kept across steps (every thread holds the same value)
    long long iter = 0;

    // one save point at time tsv: value column yv (lane-distributed in warp 0), this thread's column value via colval(i)
    auto loss_term = [&](double d, double yc, double iys, double& term, double& g) {
      if (G.loss_kind == CRNN_LOSS_MAE_SCALED) { const double diff = d * iys - yc * iys; term = fabs(diff); g = signbit(diff) ? iys : -iys; }
      else if (G.loss_kind == CRNN_LOSS_MSE) { const double diff = d * iys - yc * iys; term = diff * diff; g = -2.0 * diff * iys; }
      else { const double diff = lean_log(clampd(d, P.pred_lo, P.pred_hi)) - lean_log(yc); term = fabs(diff); g = (signbit(diff) ? 1.0 : -1.0) / yc; }
    };
    auto emit_save = [&](double tsv, double yv, auto colval) {
      if (has_obs) {
        // observable post-map: y = sum_j w_obs[j] r_j(u(tsv), tsv) - one more value-path evaluation at the saved state
        if (w0) {
          WideAux ax; int sg = tab_seg;
          (void)gen_rhs<F2>(P, sb, S.tmp, lane, my_mw, tsv, yv, ax, sg, toff);
          const double y = wsum(lane < nr ? S.wobs[lane] * S.tmp.r[lane] : 0.0);
          if (lane == 0) {
            const double yc = clampd(y, P.pred_lo, P.pred_hi);
            const bool inside = (y >= P.pred_lo) && (y <= P.pred_hi);
            const size_t off = (size_t)P.n_obs * isave;
            if (pred) pred[pbase + off] = yc;
            double term, g;
            loss_term(__ldg(datat + off), yc, __ldg(G.inv_ys), term, g);
            loss_acc += term;
            S.g[0] = inside ? g : 0.0;
          }
        }
        __syncthreads();
        if (active) Gc = fma(S.g[0], col_obs<F2>(P, sb, S.tmp, S.wobs, srow, cols, tid, ds, colval), Gc);
        __syncthreads();
        return;
      }
      if (w0) {
        double g = 0.0;
        if (my_obs >= 0) {
          const double yc = clampd(yv, P.pred_lo, P.pred_hi);
          const bool inside = (yv >= P.pred_lo) && (yv <= P.pred_hi);
          const size_t off = (size_t)my_obs + (size_t)P.n_obs * isave;
          if (pred) pred[pbase + off] = yc;
          double term;
          loss_term(__ldg(datat + off), yc, my_iys, term, g);
          loss_acc += term;
          if (!inside) g = 0.0;
        }
        S.g[lane] = g;
      }
      __syncthreads();
      if (active) for (int i = 0; i < n; ++i) Gc = fma(S.g[i], colval(i), Gc);
      __syncthreads();
    };

    while (isave < nsave && __ldg(P.saveat + isave) <= t0) {
      emit_save(__ldg(P.saveat + isave), u, [&](int i) { return slot(SL_U)[i * cs]; });
      ++isave;
    }

    while (t < tend) {
      ++iter;
      if (autosw && iter > 1) {  // OrdinaryDiffEq AutoSwitch (oracle solve_one header comment)
        const bool stiff = fabs(eigen_est * dt / 3.5068) > 0.9;
        sw_count = stiff ? (sw_count < 0 ? 1 : sw_count + 1) : (sw_count > 0 ? -1 : sw_count - 1);
        bool want = rosen;
        if (!rosen && sw_count > 10) { dt = dt * 2.0; want = true; }
        else if (rosen && sw_count < -3) { dt = dt / 2.0; want = false; }
        if (want != rosen) {
          rosen = want;   // initialize!(integrator, new cache): fsalfirst = f(uprev) on all columns
          if (w0) { KS(0) = gen_rhs<F2>(P, sb, S.base, lane, my_mw, t, u, a0, tab_seg, toff); }
          ++n_rhs;
          __syncthreads();
          if (active) col_apply<F2, false>(P, sb, S.base, S.d2, srow, cols, tid, ds, slot(SL_U), CKS(0), nullptr, cs);
          __syncthreads();
        }
      }
      if (dt != dt) { ret = CRNN_RET_DTNAN; break; }
      if (iter > P.maxiters) { ret = CRNN_RET_MAXITERS; break; }
      dt = jmin(dt, dtmax);
      dt = jmin(dt, tend - t);
      if (dt <= dtmin && tend - t > dtmin) { ret = CRNN_RET_DTLESSTHANMIN; break; }
      {  // has_nan over all columns
        bool bad = w0 && lane < n && u != u;
        if (active) for (int i = 0; i < n; ++i) { const double v = slot(SL_U)[i * cs]; bad |= (v != v); }
        if (__syncthreads_or(bad ? 1 : 0)) { ret = CRNN_RET_UNSTABLE; break; }
      }

      double un = 0.0, e = 0.0;
      if (!rosen) {
        // ---- Tsit5: stage s state into the Y slot, K_s from it ----
        double g6 = u, den_c = 0.0;
#pragma unroll 1
        for (int s = 1; s < 7; ++s) {
          if (active) {
            for (int i = 0; i < n; ++i) {
              double acc = tsc::A[s][0] * CKS(0)[i * cs];
              for (int j = 1; j < s; ++j) acc = fma(tsc::A[s][j], CKS(j)[i * cs], acc);
              const double y = fma(dt, acc, slot(SL_U)[i * cs]);
              if (s == 6 && autosw) { const double b = y - slot(SL_Y)[i * cs]; den_c = fma(b, b, den_c); }  // Y still holds g6
              slot(SL_Y)[i * cs] = y;
            }
          }
          if (w0) {
            double acc = tsc::A[s][0] * KS(0);
            for (int j = 1; j < s; ++j) acc = fma(tsc::A[s][j], KS(j), acc);
            const double y = fma(dt, acc, u);
            if (s == 5) g6 = y;
            un = y;
            KS(s) = gen_rhs<F2>(P, sb, S.cur, lane, my_mw, t + tsc::C[s] * dt, y, as, tab_seg, toff);
          }
          ++n_rhs;
          __syncthreads();
          if (active) col_apply<F2, false>(P, sb, S.cur, S.d2, srow, cols, tid, ds, slot(SL_Y), CKS(s), nullptr, cs);
          __syncthreads();
        }
        if (w0) {
          double acc = tsc::BT[0] * KS(0);
#pragma unroll
          for (int j = 1; j < 7; ++j) acc = fma(tsc::BT[j], KS(j), acc);
          e = dt * acc;
        }
        if (autosw) {  // eigen_est = |k7 - k6| / |u_{n+1} - g6| over the columns that take part in the norm
          if (incl) row_sums(2, [&](int rw) {
            if (rw == 1) return den_c;
            double num = 0.0;
            for (int i = 0; i < n; ++i) { const double a = CKS(6)[i * cs] - CKS(5)[i * cs]; num = fma(a, a, num); }
            return num;
          });
          if (w0) {
            double a = 0.0, b = 0.0;
            if (lane < n) { a = KS(6) - KS(5); a *= a; b = un - g6; b *= b; }
            const double num = wsum(a) + (incl ? S.rows[0] : 0.0), den = wsum(b) + (incl ? S.rows[1] : 0.0);
            if (lane == 0) S.bcast[1] = sqrt(num / G.norm_cnt) / sqrt(den / G.norm_cnt);
          }
          __syncthreads();
          eigen_est = S.bcast[1];
        }
      } else if (STIFF == 1) {
        // ---- TRBDF2 with dual columns (oracle solve_one, TRBDF2 branch): duals through the SAME simplified-Newton iterations.
        //      W dz = r(z), W = I - gdt J(u_n); column c:  W dz'_c = dt f'(y)[(y'_c, dW_c)] - z'_c + gdt D^2 f(u_n)[(S_c, dW_c), (dz, 0)];
        //      the convergence test and the smoothed error estimate use the dual-aware norm.  No Jacobian refresh with columns
        //      (its dual would need D^2 f at the refresh point): Newton failure => dt/2.
        //      slots: K0 fsalfirst, K1 z1, K2 z_gamma, K3 z3, K4 tmp (then the error columns), K5 f' / dz (then fsallast), K6 D^2 ----
        constexpr double s2 = 1.4142135623730951, gam = 2.0 - s2, d = 1.0 - s2 / 2.0, w = s2 / 4.0;
        constexpr double bt1 = (1.0 - s2) / 3.0, bt2 = 1.0 / 3.0, bt3 = (s2 - 2.0) / 3.0, al1 = -s2 / 2.0, al2 = 1.0 + s2 / 2.0;
        const double gdt = d * dt;
        double z1 = 0.0, zg = 0.0, z3 = 0.0, tmpv = 0.0, zs = 0.0;
        if (w0) {
          const double eig = wide_build_lu<F2>(P, sb, ww, lane, S.base.r, a0, gdt);
          if (autosw && lane == 0) S.bcast[1] = eig;
          z1 = dt * KS(0);
        }
        ++n_jac;
        if (active) for (int i = 0; i < n; ++i) CKS(1)[i * cs] = dt * CKS(0)[i * cs];
        __syncthreads();
        if (autosw) eigen_est = S.bcast[1];
        bool ok = true;
#pragma unroll 1
        for (int stg = 0; stg < 2 && ok; ++stg) {
          double* const ZS = CKS(stg ? 3 : 2);
          if (w0) {
            if (stg == 0) { tmpv = fma(d, z1, u); zs = z1; }
            else { tmpv = fma(w, zg, fma(w, z1, u)); zs = fma(al2, zg, al1 * z1); }
          }
          if (active) for (int i = 0; i < n; ++i) {
            const double uc = slot(SL_U)[i * cs], c1 = CKS(1)[i * cs];
            if (stg == 0) { CKS(4)[i * cs] = fma(d, c1, uc); ZS[i * cs] = c1; }
            else { const double cg = CKS(2)[i * cs]; CKS(4)[i * cs] = fma(w, cg, fma(w, c1, uc)); ZS[i * cs] = fma(al2, cg, al1 * c1); }
          }
          const double tst = t + (stg ? 1.0 : gam) * dt;
          bool conv = false;
          double ndz_prev = 0.0, eta = lean_pow(fmax(eta_old, 2.220446049250313e-16), 0.8);
#pragma unroll 1
          for (int it = 1; it <= 10; ++it) {
            double dzv = 0.0, yk = 0.0;
            if (w0) {
              yk = fma(d, zs, tmpv);
              const double fk = gen_rhs<F2>(P, sb, S.cur, lane, my_mw, tst, yk, as, tab_seg, toff);
              dzv = wide_lusolve(ww, lane, ns, fma(dt, fk, -zs));
              gen_dir2<F2>(P, sb, S.base, S.d2, lane, dzv, 0.0);                       // second direction (dz, tau = 0)
            }
            ++n_rhs;
            if (active) for (int i = 0; i < n; ++i) slot(SL_Y)[i * cs] = fma(d, ZS[i * cs], CKS(4)[i * cs]);
            __syncthreads();
            if (active) {
              col_apply<F2, false>(P, sb, S.cur, S.d2, srow, cols, tid, ds, slot(SL_Y), CKS(5), nullptr, cs);
              col_apply<F2, true>(P, sb, S.base, S.d2, srow, cols, tid, ds, slot(SL_U), CKS(6), CKS(0), cs);
              for (int i = 0; i < n; ++i) CKS(5)[i * cs] = fma(gdt, CKS(6)[i * cs], fma(dt, CKS(5)[i * cs], -ZS[i * cs]));
              col_lusolve(ww, ns, CKS(5), cs);
            }
            if (incl) {
              row_sums(3 * n, [&](int rw) {
                const int i = rw % n, what = rw / n;
                const double v = what == 0 ? CKS(5)[i * cs] : (what == 1 ? slot(SL_U)[i * cs] : slot(SL_Y)[i * cs]);
                return v * v;
              });
            } else __syncthreads();
            if (w0) {
              double term = 0.0;
              if (lane < n) {
                const double e2 = dzv * dzv + (incl ? S.rows[lane] : 0.0);
                const double a2 = u * u + (incl ? S.rows[n + lane] : 0.0), b2 = yk * yk + (incl ? S.rows[2 * n + lane] : 0.0);
                const double sc = my_at + fmax(sqrt(a2), sqrt(b2)) * my_rt;
                term = e2 / (sc * sc);
              }
              const double nd = sqrt(wsum(term) / G.norm_cnt);
              if (lane == 0) S.bcast[2] = nd;
              zs += dzv;
            }
            if (active) for (int i = 0; i < n; ++i) ZS[i * cs] += CKS(5)[i * cs];
            __syncthreads();
            const double ndz = S.bcast[2];
            if (it > 1) {
              const double theta = ndz / ndz_prev;
              if (!(theta <= 2.0)) break;
              eta = theta / (1.0 - theta);
            }
            if ((eta >= 0.0 && eta * ndz < 0.01) || ndz == 0.0) { conv = true; eta_old = eta; break; }
            ndz_prev = ndz;
          }
          if (w0) { if (stg == 0) zg = zs; else z3 = zs; }
          if (!conv) ok = false;
        }
        if (!ok) { dt_last = dt; ++n_rej; dt = dt / 2.0; __syncthreads(); continue; }
        if (w0) {
          un = fma(d, z3, tmpv);
          e = wide_lusolve(ww, lane, ns, lane < ns ? fma(bt3, z3, fma(bt2, zg, bt1 * z1)) : 0.0);   // smooth_est
          gen_dir2<F2>(P, sb, S.base, S.d2, lane, e, 0.0);
          KS(1) = z1; KS(2) = zg; KS(3) = z3; KS(5) = z3 / dt;                                        // fsallast
        }
        if (active) for (int i = 0; i < n; ++i) slot(SL_Y)[i * cs] = fma(d, CKS(3)[i * cs], CKS(4)[i * cs]);   // u_{n+1} of the column
        __syncthreads();
        if (active) {
          col_apply<F2, true>(P, sb, S.base, S.d2, srow, cols, tid, ds, slot(SL_U), CKS(6), CKS(0), cs);
          for (int i = 0; i < n; ++i) {
            const double c1 = CKS(1)[i * cs], cg = CKS(2)[i * cs], c3 = CKS(3)[i * cs];
            CKS(4)[i * cs] = fma(gdt, CKS(6)[i * cs], fma(bt3, c3, fma(bt2, cg, bt1 * c1)));
            CKS(5)[i * cs] = c3 / dt;
          }
          col_lusolve(ww, ns, CKS(4), cs);                                                             // the error columns
        }
        __syncthreads();
      } else {
        // ---- Rosenbrock23 = ode23s: K0=f0, K1..K3=k1..k3, K4=f1, K5=f2; caches at u_n in S.base ----
        const double d = 1.0 / (2.0 + 1.4142135623730951), e32 = 6.0 + 1.4142135623730951;
        const double g = d * dt;
        double k1 = 0.0, k2 = 0.0, k3 = 0.0, f1 = 0.0, f2 = 0.0, dTv = 0.0;
        if (w0) {
          dTv = wide_time_deriv<F2>(P, sb, ww, lane, t, S.base.r, a0, tab_seg, toff);
          const double eig = wide_build_lu<F2>(P, sb, ww, lane, S.base.r, a0, g);
          if (autosw && lane == 0) S.bcast[1] = eig;
          k1 = wide_lusolve(ww, lane, ns, fma(g, dTv, KS(0)));
          gen_dir2<F2>(P, sb, S.base, S.d2, lane, k1, 1.0);          // (k1, tau = 1): J'k1 and the dual part of g*dT
        }
        ++n_jac;
        __syncthreads();
        if (autosw) eigen_est = S.bcast[1];
        if (active) {
          col_apply<F2, true>(P, sb, S.base, S.d2, srow, cols, tid, ds, slot(SL_U), CKS(1), CKS(0), cs);
          for (int i = 0; i < n; ++i) CKS(1)[i * cs] = fma(g, CKS(1)[i * cs], CKS(0)[i * cs]);
          col_lusolve(ww, ns, CKS(1), cs);
          for (int i = 0; i < n; ++i) slot(SL_Y)[i * cs] = fma(0.5 * dt, CKS(1)[i * cs], slot(SL_U)[i * cs]);
        }
        __syncthreads();   // columns are done with d2 (k1)
        if (w0) f1 = gen_rhs<F2>(P, sb, S.cur, lane, my_mw, t + 0.5 * dt, fma(0.5 * dt, k1, u), as, tab_seg, toff);
        ++n_rhs;
        __syncthreads();
        if (w0) {
          const double k2v = wide_lusolve(ww, lane, ns, f1 - k1);
          k2 = k2v + k1;
          gen_dir2<F2>(P, sb, S.base, S.d2, lane, k2v, 0.0);         // (k2 - k1, 0)
        }
        if (active) col_apply<F2, false>(P, sb, S.cur, S.d2, srow, cols, tid, ds, slot(SL_Y), CKS(4), nullptr, cs);
        __syncthreads();
        if (active) {
          col_apply<F2, true>(P, sb, S.base, S.d2, srow, cols, tid, ds, slot(SL_U), CKS(2), CKS(0), cs);
          for (int i = 0; i < n; ++i) CKS(2)[i * cs] = fma(g, CKS(2)[i * cs], CKS(4)[i * cs] - CKS(1)[i * cs]);
          col_lusolve(ww, ns, CKS(2), cs);
          for (int i = 0; i < n; ++i) {
            const double v = CKS(2)[i * cs] + CKS(1)[i * cs];
            CKS(2)[i * cs] = v;
            slot(SL_Y)[i * cs] = fma(dt, v, slot(SL_U)[i * cs]);     // u_{n+1} of the column
          }
        }
        __syncthreads();
        if (w0) { un = fma(dt, k2, u); f2 = gen_rhs<F2>(P, sb, S.cur, lane, my_mw, t + dt, un, as, tab_seg, toff); }
        ++n_rhs;
        __syncthreads();
        if (w0) {
          k3 = wide_lusolve(ww, lane, ns, f2 - e32 * (k2 - f1) - 2.0 * (k1 - KS(0)) + dt * dTv);
          gen_dir2<F2>(P, sb, S.base, S.d2, lane, k3, 1.0 / d);      // (k3, tau = dt / g): J'k3 and the dual part of dt*dT
          e = dt / 6.0 * (k1 - 2.0 * k2 + k3);
          KS(1) = k1; KS(2) = k2; KS(3) = k3; KS(4) = f1; KS(5) = f2;
        }
        if (active) col_apply<F2, false>(P, sb, S.cur, S.d2, srow, cols, tid, ds, slot(SL_Y), CKS(5), nullptr, cs);
        __syncthreads();
        if (active) {
          col_apply<F2, true>(P, sb, S.base, S.d2, srow, cols, tid, ds, slot(SL_U), CKS(3), CKS(0), cs);
          for (int i = 0; i < n; ++i) {
            const double rhs = CKS(5)[i * cs] - e32 * (CKS(2)[i * cs] - CKS(4)[i * cs]) - 2.0 * (CKS(1)[i * cs] - CKS(0)[i * cs]);
            CKS(3)[i * cs] = fma(g, CKS(3)[i * cs], rhs);
          }
          col_lusolve(ww, ns, CKS(3), cs);
        }
        __syncthreads();
      }

      // ---- error estimate with the dual-aware norm ----
      if (incl) {
        row_sums(3 * n, [&](int rw) {
          const int i = rw % n, what = rw / n;
          double v;
          if (what == 0) {
            if (!rosen) {
              double acc = tsc::BT[0] * CKS(0)[i * cs];
              for (int j = 1; j < 7; ++j) acc = fma(tsc::BT[j], CKS(j)[i * cs], acc);
              v = dt * acc;
            } else if (STIFF == 1) v = CKS(4)[i * cs];
            else v = dt / 6.0 * (CKS(1)[i * cs] - 2.0 * CKS(2)[i * cs] + CKS(3)[i * cs]);
          } else if (what == 1) v = slot(SL_U)[i * cs];
          else v = slot(SL_Y)[i * cs];
          return v * v;
        });
      }
      if (w0) {
        double term = 0.0;
        if (lane < n) {
          const double e2 = e * e + (incl ? S.rows[lane] : 0.0);
          const double a2 = u * u + (incl ? S.rows[n + lane] : 0.0), b2 = un * un + (incl ? S.rows[2 * n + lane] : 0.0);
          const double sc = my_at + fmax(sqrt(a2), sqrt(b2)) * my_rt;
          term = e2 / (sc * sc);
        }
        const double EE = sqrt(wsum(term) / G.norm_cnt);
        if (lane == 0) S.bcast[0] = EE;
      }
      __syncthreads();
      const double EEst = S.bcast[0];
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
        if (STIFF == 1 && rosen) {   // the next attempt's analytic Jacobian needs the by-products at u_{n+1} (fsallast stays z3/dt)
          if (w0) (void)gen_rhs<F2>(P, sb, S.cur, lane, my_mw, t, un, as, tab_seg, toff);
          ++n_rhs;
          __syncthreads();
        }
        while (isave < nsave) {
          const double tsv = __ldg(P.saveat + isave);
          if (!(tsv <= t)) break;
          if (tsv == t) {
            emit_save(tsv, un, [&](int i) { return slot(SL_Y)[i * cs]; });
          } else {
            const double th = (tsv - tprev) / dt;
            if (!rosen) {
              double bs[7];
#pragma unroll
              for (int s = 0; s < 7; ++s) bs[s] = th * (tsc::R[s][0] + th * (tsc::R[s][1] + th * (tsc::R[s][2] + th * tsc::R[s][3])));
              double yv = 0.0;
              if (w0) {
                double acc = bs[0] * KS(0);
#pragma unroll
                for (int s = 1; s < 7; ++s) acc = fma(bs[s], KS(s), acc);
                yv = fma(dt, acc, u);
              }
              emit_save(tsv, yv, [&](int i) {
                double acc = bs[0] * CKS(0)[i * cs];
#pragma unroll
                for (int s = 1; s < 7; ++s) acc = fma(bs[s], CKS(s)[i * cs], acc);
                return fma(dt, acc, slot(SL_U)[i * cs]);
              });
            } else if (STIFF == 1) {   // Hermite on (u_n, fsalfirst) .. (u_{n+1}, fsallast), every column
              auto herm = [&](double a, double b, double fa, double fb) {
                return (1.0 - th) * a + th * b + th * (th - 1.0) * ((1.0 - 2.0 * th) * (b - a) + (th - 1.0) * dt * fa + th * dt * fb);
              };
              const double yv = w0 ? herm(u, un, KS(0), KS(5)) : 0.0;
              emit_save(tsv, yv, [&](int i) { return herm(slot(SL_U)[i * cs], slot(SL_Y)[i * cs], CKS(0)[i * cs], CKS(5)[i * cs]); });
            } else {
              const double d = 1.0 / (2.0 + 1.4142135623730951);
              const double c1 = th * (1.0 - th) / (1.0 - 2.0 * d), c2 = th * (th - 2.0 * d) / (1.0 - 2.0 * d);
              const double yv = w0 ? u + dt * (c1 * KS(1) + c2 * KS(2)) : 0.0;
              emit_save(tsv, yv, [&](int i) { return slot(SL_U)[i * cs] + dt * (c1 * CKS(1)[i * cs] + c2 * CKS(2)[i * cs]); });
            }
          }
          ++isave;
        }
        // commit: u_n <- u_{n+1}, FSAL, the cache of the last evaluation becomes the cache at u_n
        if (w0) { u = un; KS(0) = rosen ? KS(5) : KS(6); a0 = as; }
        if (active) {
          const int fs = rosen ? 5 : 6;
          for (int i = 0; i < n; ++i) { slot(SL_U)[i * cs] = slot(SL_Y)[i * cs]; CKS(0)[i * cs] = CKS(fs)[i * cs]; }
        }
        for (int q = tid; q < (int)(sizeof(GenPoint) / sizeof(double)); q += blockDim.x)
          reinterpret_cast<double*>(&S.base)[q] = reinterpret_cast<const double*>(&S.cur)[q];
        __syncthreads();
        dt = jmin(dtnew, dtmax);
      } else {
        ++n_rej;
        dt = dt / jmin(P.inv_qmin, q11 / P.gamma);
      }
    }
    if (ret == CRNN_RET_DEFAULT) ret = CRNN_RET_SUCCESS;

    // ---- per-trajectory outputs ----
    const double cnt = (double)P.n_obs * (double)isave;
    if (w0) {
      const double ltot = wsum(loss_acc);
      if (pred && (has_obs ? lane == 0 : my_obs >= 0))
        for (int ks = isave; ks < P.n_save; ++ks) pred[pbase + (has_obs ? 0 : my_obs) + (size_t)P.n_obs * ks] = 0.0;
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
    }
    if (active) grad_each[(size_t)traj * np + tid] = isave > 0 ? Gc / cnt : 0.0;
    __syncthreads();
  }
#undef KS
#undef CKS
}

}  // namespace crnn
