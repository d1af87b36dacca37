// wide_common.cuh — building blocks of the lane-per-component kernels (k_wide_solve, k_kencarp4_wide,
// k_tsit5_adjoint): one WARP owns one trajectory, lane i owns state component i, dimensions are RUNTIME values
// (n_state, n_in, n_reac <= 32).  The RHS of every flavour (F0, F1, F2), its analytic Jacobian assembled in
// shared memory, df/dt for the non-autonomous F2, the cooperative LU and the triangular solves live here.
#pragma once
#include "crnn_dev.cuh"

namespace crnn {

constexpr int KW_MAXN = 32;

struct WideP {
  double abstol[KW_MAXN], reltol[KW_MAXN];
  double lb, ub, gas_R;
  double t0, t1, pred_lo, pred_hi;
  double inv_qmin, inv_qmax, gamma, beta1, beta2, inv_order;
  double qs_min, qs_max;  // step_accept_controller!'s dead-band
  long long maxiters;
  const double* w_inT;   // device [n_in][nrp]: w_in transposed (reaction fastest), nrp = 32
  const double* w_b;     // device [n_reac]
  const double* w_out;   // device [n_species x n_reac] col-major, out_scale folded in
  const double* saveat;  // device [n_save]
  const int* row2obs;    // device [n_state]
  int n, ns, nin, nr, kind;
  int n_save, n_obs;
  // generic solve path (kernel_wide_solve.cuh)
  int alg, n_tab;
  double beta1_ros, beta2_ros;  // PI exponents while AutoTsit5 runs Rosenbrock23 (beta1/beta2 above: Tsit5)
  const double* mw;      // device [n_species]           (F2)
  const double* tab_t;   // device [n_tab] knots         (F2)
  const double* tab_T;   // device [n_tab]
  const double* tab_P;   // device [n_tab]
  const double* w_obs;   // device [n_reac] or NULL: observable post-map y = sum_j w_obs[j] r_j(u(ts), ts) (k_wide_solve's saves)
  // F4 (MLP-augmented inputs, yeast_glycolysis.jl:128-142 / rober_crnn_qssa.jl:111-126): device arrays, Flux.destructure order
  int mlp_layers, mlp_act_out;
  int mlp_dims[10];
  const int* mlp_in_idx;     // [d0]
  const int* aug_src;        // [n_in]
  const double* mlp_params;
  const double* w_J;         // [n_species] or NULL
};

// ---- shared by the lane-per-component kernels (k_wide_solve, k_tsit5_adjoint): F2 tables ----
constexpr double kGasRu = 8.31446261815324e3;  // HyChem/crnn_pyrolysis_mass.jl:108

struct TabVal { double T, P, Td, Pd; };

// Interpolations.LinearInterpolation(tab_t, v)(t) and its slope; segment = last one whose left knot is <= t.
// `seg` is the caller's hint (the segment of its previous lookup): stage times move monotonically inside a step, so the
// hint nearly always holds and the six dependent loads of the binary search are skipped.
// `toff`: offset of this trajectory's T / P tables (per-experiment temperature programmes; 0 = the model's own).
__device__ __forceinline__ TabVal wide_tab(const WideP& P, double t, int& seg, size_t toff = 0) {
  int lo = seg;
  double ta = __ldg(P.tab_t + lo), tb = __ldg(P.tab_t + lo + 1);
  if (!(ta <= t && (t < tb || lo == P.n_tab - 2))) {
    lo = 0;
    int hi = P.n_tab - 1;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(P.tab_t + mid) <= t) lo = mid; else hi = mid;
    }
    ta = __ldg(P.tab_t + lo); tb = __ldg(P.tab_t + lo + 1);
    seg = lo;
  }
  const double h = tb - ta, w = (t - ta) / h;
  const double T0 = __ldg(P.tab_T + toff + lo), T1 = __ldg(P.tab_T + toff + lo + 1);
  const double P0 = __ldg(P.tab_P + toff + lo), P1 = __ldg(P.tab_P + toff + lo + 1);
  TabVal v;
  v.T = T0 + w * (T1 - T0); v.P = P0 + w * (P1 - P0);
  v.Td = (T1 - T0) / h; v.Pd = (P1 - P0) / h;
  return v;
}

// out-of-line warp sum for the lane-per-component kernels: they reduce at dozens of sites (norms, Newton tests), and every
// inlined copy is ten shuffles plus their convergence-barrier code — instruction-cache pressure is what binds these kernels
static __device__ __noinline__ double wsum(double v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, m));
  return v;
}

// Partial pivoting: among the lanes with `cand`, the one whose `a` (>= 0) is largest, ties to the smallest `idx` (< 2^23);
// returns (idx << 8) | lane of the winner in every lane.  Three REDUX instead of a 5-round butterfly of 64-bit shuffles.
// (A NaN wins - the matrix is lost either way and the trajectory ends as Unstable.)
__device__ __forceinline__ int warp_argmax_abs(double a, bool cand, int idx, int lane) {
  const int hi = cand ? __double2hiint(a) : -1;
  const int m1 = __reduce_max_sync(0xffffffffu, hi);
  const bool c1 = cand && hi == m1;
  const unsigned lo = c1 ? (unsigned)__double2loint(a) : 0u;
  const unsigned m2 = __reduce_max_sync(0xffffffffu, lo);
  const int key = (c1 && lo == m2) ? ((idx << 8) | lane) : 0x7fffffff;
  return __reduce_min_sync(0xffffffffu, key);
}

template <int NK>   // NK stage-vector slots (k_wide_solve: 7; k_kencarp4_wide keeps its stages in registers: 0)
struct alignas(16) WideWarpT {
  double A[KW_MAXN][KW_MAXN + 1];  // W and its LU (row i is lane i's; +1 pad: conflict-free columns)
  double x[KW_MAXN], r[KW_MAXN], r0[KW_MAXN];
  int perm[KW_MAXN];
  int piv[KW_MAXN];   // pivot row chosen at elimination step k (LAPACK ipiv): lets a thread permute its own right-hand side in place
  // generic solve path: broadcast slots of the per-lane Jacobian factors, stage vectors
  double bdx[KW_MAXN], brr[KW_MAXN], bchi[KW_MAXN], ws[KW_MAXN];
  double dinv[KW_MAXN];  // 1/u_kk of the LU: back-substitution multiplies instead of dividing
  double k[NK > 0 ? NK : 1][NK > 0 ? KW_MAXN : 2];
};
using WideWarp = WideWarpT<7>;

// Index lists of the NON-ZERO weights, built once per block (wide_block_init).  A CRNN is sparse by construction - w_in =
// clamp(-w_out, 0, ..) is zero wherever a species is not a reactant, pruned models (case*_pruning.jl) more so, the 30-reaction
// HyChem-sized model has <= 3 inputs per reaction - and the two mat-vecs of the RHS are a sixth of the stiff kernels'
// instructions when done densely.  The lists keep the dense loop's order (increasing index), so the sums are the same bits.
constexpr int KW_SPMAX = 12;   // longest list kept; a denser row / column falls back to the dense loop
struct WideSparse {
  unsigned char idx_in[KW_SPMAX][KW_MAXN];   // [k][j]: k-th input row i with w_in[i,j] != 0
  unsigned char idx_out[KW_SPMAX][KW_MAXN];  // [k][i]: k-th reaction j with w_out[i,j] != 0
  unsigned char cnt_in[KW_MAXN], cnt_out[KW_MAXN];
  int max_in, max_out, use_in, use_out;
};

struct alignas(16) WideBlock {
  double w_inT[KW_MAXN][KW_MAXN];  // [i][j]
  double w_inJ[KW_MAXN][KW_MAXN];  // [j][i]: the Jacobian assembly reads 8 consecutive inputs of a reaction with 4 LDS.128
  double w_out[KW_MAXN][KW_MAXN];  // [j][i]: lane i reads consecutive addresses for fixed j
  double w_b[KW_MAXN];
  WideSparse sp;
};

// weights into shared memory + the sparse index lists; every thread of the block calls it
__device__ __forceinline__ void wide_block_init(const WideP& P, WideBlock& sb) {
  const int ns = P.ns, nin = P.nin, nr = P.nr;
  for (int q = threadIdx.x; q < KW_MAXN * KW_MAXN; q += blockDim.x) {
    const int i = q / KW_MAXN, j = q % KW_MAXN;
    sb.w_inT[i][j] = (i < nin && j < nr) ? P.w_inT[i * KW_MAXN + j] : 0.0;
    sb.w_inJ[j][i] = sb.w_inT[i][j];
    sb.w_out[i][j] = (i < nr && j < ns) ? P.w_out[j + ns * i] : 0.0;  // [reaction][species]
  }
  for (int q = threadIdx.x; q < KW_MAXN; q += blockDim.x) sb.w_b[q] = q < nr ? P.w_b[q] : 0.0;
  __syncthreads();
  if (threadIdx.x < KW_MAXN) {
    const int j = threadIdx.x;
    int c = 0;
    for (int i = 0; i < nin; ++i)
      if (sb.w_inT[i][j] != 0.0) { if (c < KW_SPMAX) sb.sp.idx_in[c][j] = (unsigned char)i; ++c; }
    sb.sp.cnt_in[j] = (unsigned char)c;
  } else if (threadIdx.x < 2 * KW_MAXN) {
    const int i = threadIdx.x - KW_MAXN;
    int c = 0;
    for (int j = 0; j < nr; ++j)
      if (sb.w_out[j][i] != 0.0) { if (c < KW_SPMAX) sb.sp.idx_out[c][i] = (unsigned char)j; ++c; }
    sb.sp.cnt_out[i] = (unsigned char)c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int mi = 0, mo = 0;
    for (int q = 0; q < KW_MAXN; ++q) { mi = max(mi, (int)sb.sp.cnt_in[q]); mo = max(mo, (int)sb.sp.cnt_out[q]); }
    sb.sp.max_in = mi; sb.sp.max_out = mo;
    // a list entry costs an index load and a gathered operand more than a dense term
    sb.sp.use_in = (mi <= KW_SPMAX && 5 * mi < 3 * nin) ? 1 : 0;
    sb.sp.use_out = (mo <= KW_SPMAX && 5 * mo < 3 * nr) ? 1 : 0;
  }
  __syncthreads();
}

// the adjoint kernel assembles no Jacobian: no reaction-major copy of w_in (8 KB more room for its step record)
// rows padded to 33: the adjoint's transposed mat-vecs read these with the LANE as the row index (g_j = sum_i w_out[j][i] mu_i on
// lane j, (J^T lambda)_l = sum_j w_in[l][j] g_j r_j on lane l) - with 32-double rows every lane hit the same bank (37 % of the
// kernel's shared-memory wavefronts were conflict replays); both access directions are conflict-free at 33
struct alignas(16) WideBlockLite {
  double w_inT[KW_MAXN][KW_MAXN + 1];  // [i][j]
  double w_out[KW_MAXN][KW_MAXN + 1];  // [j][i]
  double w_b[KW_MAXN];
  int ent[512];                        // quadrature entry e < n_w: offsets (left << 16) | right of its two factors (kernel_tsit5_adjoint.cuh)
};

struct WideAux {  // per-lane by-products of one RHS evaluation (what the Jacobian needs)
  double dx;       // d x_l / d u_l (F2: at fixed density)
  double rr;       // F2: d log(rho) / d u_l = -chi_l / (MW_l S)
  double wdot;     // sum_j w_out[l,j] r_j (scaled; F2: before the 1/rho)
  double inv_rho;  // F2
  double chiC;     // F2: 1 if lb <= C_l <= ub
};

// F4: the value of input row `lane` of the CRNN - a state row or an output of the Flux MLP of the state (oracle mlp_eval):
// lane k of a layer computes neuron k from the previous activations broadcast through shared memory (ww.ws / ww.bchi as scratch);
// gelu in NNlib's tanh form, tanh(y) = 1 - 2 / (exp(2y) + 1), softplus = log(1 + exp(-|x|)) + max(x, 0) with the lean functions.
template <class WW>
__device__ __forceinline__ double wide_mlp_aug(const WideP& P, WW& ww, int lane, double y) {
  __syncwarp();
  ww.ws[lane] = y;
  __syncwarp();
  double a = lane < P.mlp_dims[0] ? ww.ws[__ldg(P.mlp_in_idx + lane)] : 0.0;
  const double* w = P.mlp_params;
#pragma unroll 1
  for (int l = 0; l < P.mlp_layers; ++l) {
    const int din = P.mlp_dims[l], dout = P.mlp_dims[l + 1];
    __syncwarp();
    ww.bchi[lane] = a;
    __syncwarp();
    double s = 0.0;
    if (lane < dout) {
      const double* wp = w + lane;
#pragma unroll 4
      for (int i = 0; i < din; ++i, wp += dout) s = fma(__ldg(wp), ww.bchi[i], s);
      s += __ldg(wp);
      if (l + 1 < P.mlp_layers) {
        const double th = 1.0 - 2.0 / (lean_exp(2.0 * (0.7978845608028654 * (s + 0.044715 * (s * s * s)))) + 1.0);
        s = 0.5 * s * (1.0 + th);
      } else {
        s = P.mlp_act_out == 0 ? lean_log(1.0 + lean_exp(-fabs(s))) + (s > 0.0 ? s : 0.0) : lean_exp(s);
      }
    }
    a = s;
    w += din * dout + dout;
  }
  __syncwarp();
  ww.bchi[lane] = a;
  __syncwarp();
  if (lane >= P.nin) return 0.0;
  const int src = __ldg(P.aug_src + lane);
  return src >= 0 ? ww.ws[src] : ww.bchi[-1 - src];
}

// f(y, t): lane i holds y_i in, f_i out; leaves x in ww.x and r in ww.r.  SPARSE: the instantiation that consults the index
// lists (large models; small ones keep the dense-only code - these kernels are instruction-fetch bound).
template <bool F2, bool SPARSE = false, bool MLP = false, class WW>
__device__ __forceinline__ double wide_rhs(const WideP& P, const WideBlock& sb, WW& ww, int lane, double mw,
                                           double t, double y, WideAux& a, int& seg) {
  const int ns = P.ns, nin = P.nin, nr = P.nr;
  const bool isp = lane < ns;
  __syncwarp();
  double xi = 0.0, rho = 1.0;
  a.dx = 0.0; a.rr = 0.0; a.chiC = 0.0; a.inv_rho = 1.0;
  if (F2) {
    const bool dens = (P.kind == CRNN_RHS_F2_MASSFRAC_TP);   // F5 (Cathode/src/network.jl:68-80): the same inputs, no density map
    const TabVal tv = wide_tab(P, t, seg);
    double Y = 1.0, chi = 0.0, ymw = 0.0;
    if (isp) { Y = clampd(y, P.lb, P.ub); chi = (y >= P.lb && y <= P.ub) ? 1.0 : 0.0; ymw = dens ? Y / mw : 0.0; }
    const double S = dens ? wsum(ymw) : 1.0;
    rho = dens ? tv.P / (kGasRu * tv.T * S) : 1.0;
    if (isp) {
      const double C = dens ? rho * ymw * 1e3 : Y;
      a.chiC = (C >= P.lb && C <= P.ub) ? 1.0 : 0.0;
      xi = lean_log(clampd(C, P.lb, P.ub));
      a.dx = a.chiC * chi / Y;
      a.rr = dens ? -chi / (mw * S) : 0.0;
    } else if (lane == ns) {
      xi = -1.0 / P.gas_R / tv.T;
    } else if (lane == ns + 1) {
      xi = lean_log(tv.T);
    }
    a.inv_rho = 1.0 / rho;
  } else if (MLP) {
    const double v = wide_mlp_aug(P, ww, lane, y);
    if (lane < nin) xi = lean_log(clampd(v, P.lb, P.ub));   // no dx: this flavour's Jacobian is taken by finite differences
  } else if (isp) {
    const double uc = clampd(y, P.lb, P.ub);
    xi = lean_log(uc);
    a.dx = (y >= P.lb && y <= P.ub) ? __drcp_rn(uc) : 0.0;
  } else if (!F2 && P.kind == 1 && lane == ns) {
    xi = -1.0 / (P.gas_R * y);
    a.dx = 1.0 / (P.gas_R * y * y);
  }
  ww.x[lane] = xi;
  __syncwarp();
  if (lane < nr) {
    double z = sb.w_b[lane];
    if (SPARSE && sb.sp.use_in) {
      const int c = sb.sp.cnt_in[lane], cm = sb.sp.max_in;
#pragma unroll 2
      for (int k = 0; k < cm; ++k)
        if (k < c) { const int i = sb.sp.idx_in[k][lane]; z = fma(sb.w_inT[i][lane], ww.x[i], z); }
    } else {
#pragma unroll 2
      for (int i = 0; i < nin; ++i) z = fma(sb.w_inT[i][lane], ww.x[i], z);
    }
    ww.r[lane] = lean_exp(z);
  }
  __syncwarp();
  double f = 0.0;
  if (isp) {
    if (SPARSE && sb.sp.use_out) {
      const int c = sb.sp.cnt_out[lane], cm = sb.sp.max_out;
#pragma unroll 2
      for (int k = 0; k < cm; ++k)
        if (k < c) { const int j = sb.sp.idx_out[k][lane]; f = fma(sb.w_out[j][lane], ww.r[j], f); }
    } else {
#pragma unroll 2
      for (int j = 0; j < nr; ++j) f = fma(sb.w_out[j][lane], ww.r[j], f);
    }
  }
  a.wdot = f;
  if (F2) f = f / rho;
  if (MLP && isp && P.w_J) f += __ldg(P.w_J + lane);   // .+ w_J (yeast_glycolysis.jl:131); out_scale is folded into w_out only
  return f;
}

// df/dt at fixed u from the by-products of the evaluation at (u, t) (r in rsrc): F2 only, 0 otherwise
template <bool F2, class WW>
__device__ __forceinline__ double wide_time_deriv(const WideP& P, const WideBlock& sb, WW& ww, int lane, double t,
                                                  const double* rsrc, const WideAux& a, int& seg, size_t toff = 0) {
  if (!F2) return 0.0;
  const int ns = P.ns, nr = P.nr;
  const TabVal tv = wide_tab(P, t, seg, toff);
  const double rr = (P.kind == CRNN_RHS_F2_MASSFRAC_TP) ? tv.Pd / tv.P - tv.Td / tv.T : 0.0;  // F5: no density map
  __syncwarp();
  ww.bchi[lane] = a.chiC;
  __syncwarp();
  if (lane < nr) {
    double zd = 0.0;
    for (int i = 0; i < ns; ++i) zd = fma(sb.w_inT[i][lane], ww.bchi[i] * rr, zd);
    zd = fma(sb.w_inT[ns][lane], tv.Td / (P.gas_R * tv.T * tv.T), zd);
    zd = fma(sb.w_inT[ns + 1][lane], tv.Td / tv.T, zd);
    ww.ws[lane] = rsrc[lane] * zd;
  }
  __syncwarp();
  double s = 0.0;
  if (lane < ns) {
    for (int j = 0; j < nr; ++j) s = fma(sb.w_out[j][lane], ww.ws[j], s);
    s = (s - a.wdot * rr) * a.inv_rho;
  }
  __syncwarp();
  return s;
}

// W = I - gdt*J(u) from the RHS by-products (r in rsrc, this lane's aux), cooperative LU with partial
// pivoting (first strict maximum, like oracle lu_factor); returns opnorm(J, Inf).
template <bool F2, bool SPARSE = false, class WW>
__device__ __forceinline__ double wide_assemble_W(const WideP& P, const WideBlock& sb, WW& ww, int lane,
                                                  const double* rsrc, const WideAux& a, double gdt) {
  const int n = P.n, ns = P.ns, nr = P.nr;
  const bool isp = lane < ns;
  __syncwarp();
  ww.bdx[lane] = a.dx; ww.brr[lane] = a.rr; ww.bchi[lane] = a.chiC;
  __syncwarp();
  if (SPARSE && !F2 && sb.sp.use_in && sb.sp.use_out) {
    // J[i][l] = dx_l sum_j w_out[i,j] r_j w_in[l,j] over the NON-ZERO weights only: row i walks its reactions (idx_out), every
    // reaction its inputs (idx_in), accumulating in the lane's own row of A.  For a fixed (i, l) the reactions still arrive in
    // ascending order and the skipped terms are exact zeros: the same bits as the dense loop, ~350 instead of ~1800 instructions
    // per assembly on the 30-reaction model (<= 3 inputs per reaction, <= 10 reactions per species).
    double rowsum = 0.0;
    if (isp) {
      double* row = &ww.A[lane][0];
      for (int l = 0; l < n; ++l) row[l] = 0.0;
      const int c = sb.sp.cnt_out[lane], cm = sb.sp.max_out, qm = sb.sp.max_in;
      for (int k = 0; k < cm; ++k)
        if (k < c) {
          const int j = sb.sp.idx_out[k][lane];
          const double wr = sb.w_out[j][lane] * rsrc[j];
          const int ci = sb.sp.cnt_in[j];
          for (int q = 0; q < qm; ++q)
            if (q < ci) { const int l = sb.sp.idx_in[q][j]; row[l] = fma(wr, sb.w_inT[l][j], row[l]); }
        }
      for (int l = 0; l < n; ++l) {
        const double Jil = row[l] * ww.bdx[l];
        rowsum += fabs(Jil);
        if (l < ns) row[l] = (lane == l ? 1.0 : 0.0) - gdt * Jil;
      }
    }
    const double eig = warp_max(rowsum);
    __syncwarp();
    return eig;
  }
  if (F2 && lane < nr) {
    double ws = 0.0;
    for (int i = 0; i < ns; ++i) ws = fma(sb.w_inT[i][lane], ww.bchi[i], ws);
    ww.ws[lane] = ws;
  }
  __syncwarp();
  double rowsum = 0.0;
  if (isp) {
    double coef = 0.0;
    if (F2) {
      for (int j = 0; j < nr; ++j) coef = fma(sb.w_out[j][lane] * rsrc[j], ww.ws[j], coef);
      coef -= a.wdot;
    }
    // J[i][l] = (sum_j w_out[i,j] r_j w_in[l,j]) dx_l (+ F2 density terms), eight columns l per pass over the reactions
    for (int l0 = 0; l0 < n; l0 += 8) {
      double s[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) s[q] = 0.0;
      for (int j = 0; j < nr; ++j) {
        const double wr = sb.w_out[j][lane] * rsrc[j];
        const double2* wj = reinterpret_cast<const double2*>(&sb.w_inJ[j][l0]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double2 v = wj[q];
          s[2 * q] = fma(wr, v.x, s[2 * q]);
          s[2 * q + 1] = fma(wr, v.y, s[2 * q + 1]);
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int l = l0 + q;
        if (l < n) {
          double Jil = s[q] * ww.bdx[l];
          if (F2) Jil = (Jil + coef * ww.brr[l]) * a.inv_rho;
          rowsum += fabs(Jil);
          if (l < ns) ww.A[lane][l] = (lane == l ? 1.0 : 0.0) - gdt * Jil;
        }
      }
    }
  }
  const double eig = warp_max(rowsum);
  __syncwarp();
  return eig;
}

// cooperative LU of the ns x ns matrix in ww.A (lane = row, arg-max pivoting by shuffles, reciprocal diagonal in ww.dinv)
template <class WW>
__device__ __forceinline__ void wide_factor_lu(WW& ww, int lane, int ns) {
  const bool isp = lane < ns;
  ww.perm[lane] = lane;
  __syncwarp();
  for (int k = 0; k < ns; ++k) {
    const int bi = warp_argmax_abs((lane >= k && isp) ? fabs(ww.A[lane][k]) : 0.0, lane >= k && isp, lane, lane) >> 8;
    if (bi != k) {
      if (isp) { const double tmpv = ww.A[k][lane]; ww.A[k][lane] = ww.A[bi][lane]; ww.A[bi][lane] = tmpv; }
      if (lane == 0) { const int tp = ww.perm[k]; ww.perm[k] = ww.perm[bi]; ww.perm[bi] = tp; }
    }
    if (lane == 0) ww.piv[k] = bi;
    __syncwarp();
    const double rk = 1.0 / ww.A[k][k];
    if (lane == k) ww.dinv[k] = rk;
    if (lane > k && isp) {
      const double l = ww.A[lane][k] * rk;
      ww.A[lane][k] = l;
#pragma unroll 2
      for (int j = k + 1; j < ns; ++j) ww.A[lane][j] = fma(-l, ww.A[k][j], ww.A[lane][j]);
    }
    __syncwarp();
  }
}

// W = I - gdt*J(u) from the RHS by-products (r in rsrc, this lane's aux), cooperative LU with partial
// pivoting (first strict maximum, like oracle lu_factor); returns opnorm(J, Inf).
template <bool F2, class WW>
__device__ __forceinline__ double wide_build_lu(const WideP& P, const WideBlock& sb, WW& ww, int lane,
                                                const double* rsrc, const WideAux& a, double gdt) {
  const double eig = wide_assemble_W<F2>(P, sb, ww, lane, rsrc, a, gdt);
  wide_factor_lu(ww, lane, P.ns);
  return eig;
}

// W = I - gdt*J(u) with J by FORWARD FINITE DIFFERENCES (F4: TRBDF2(autodiff=false) / Rosenbrock23(autodiff=false),
// yeast_glycolysis.jl:33, rober_crnn_qssa.jl:30; oracle jac_value): step max(sqrt(eps)|u_l|, sqrt(eps)), f(u) evaluated afresh,
// the n + 1 evaluations not counted in n_rhs.  `rhs(t, y)` is the caller's evaluation; returns opnorm(J, Inf).
template <class WW, class RHS>
__device__ __forceinline__ double wide_assemble_W_fd(const WideP& P, WW& ww, int lane, double t, double u, double gdt, RHS rhs) {
  const int n = P.n, ns = P.ns;
  const double f0 = rhs(t, u);
  double rowsum = 0.0;
#pragma unroll 1
  for (int l = 0; l < n; ++l) {
    const double ul = __shfl_sync(0xffffffffu, u, l);
    const double eps = fmax(1.4901161193847656e-8 * fabs(ul), 1.4901161193847656e-8);
    const double fl = rhs(t, lane == l ? u + eps : u);
    const double Jil = (fl - f0) / eps;
    if (lane < n) {
      rowsum += fabs(Jil);
      if (l < ns && lane < ns) ww.A[lane][l] = (lane == l ? 1.0 : 0.0) - gdt * Jil;
    }
  }
  const double eig = warp_max(rowsum);
  __syncwarp();
  return eig;
}

// b <- W^{-1} b with the factored W in ww.A (lane i holds b_i)
template <class WW>
__device__ __forceinline__ double wide_lusolve(const WW& ww, int lane, int ns, double b) {
  const bool isp = lane < ns;
  b = __shfl_sync(0xffffffffu, b, ww.perm[lane]);
#pragma unroll 2
  for (int k = 0; k + 1 < ns; ++k) {
    const double bk = __shfl_sync(0xffffffffu, b, k);
    if (lane > k && isp) b = fma(-ww.A[lane][k], bk, b);
  }
#pragma unroll 2
  for (int k = ns - 1; k >= 0; --k) {
    if (lane == k) b = b * ww.dinv[k];
    const double bk = __shfl_sync(0xffffffffu, b, k);
    if (lane < k) b = fma(-ww.A[lane][k], bk, b);
  }
  return isp ? b : 0.0;
}

// Gauss-Jordan inversion of the ns x ns matrix in ww.A with every lane's ROW IN REGISTERS (ns > 16: BASELINE config 5).
// The shared-memory form below reads and writes each row once per pivot (ns^2 * 16 B of shared-memory traffic per pivot,
// 41 % of k_kencarp4_wide's wavefronts); here only the scaled pivot row passes through shared memory (16 STS.128 by the
// pivot lane, 16 broadcast LDS.128 by all).  Two devices keep every register index a compile-time constant:
//   * the columns live in a RING that is shifted left by one per pivot (the shift is free: the update writes a[m] from
//     a[m+1]), so the pivot column is always a[0] and the finished column of the inverse enters at a[31];
//   * rows never move between lanes: partial pivoting picks the pivot LANE, and each lane tracks the logical index its row
//     would have after LAPACK-style exchanges (needed for the tie-break and for where the row is finally stored).
// Same operations on the same numbers as the shared-memory form and as the oracle's kc_factor (inverse form).
template <class WW>
__device__ __noinline__ void wide_gj_regs(WW& ww, int lane, int ns) {
  const bool isp = lane < ns;
  double a[KW_MAXN];
#pragma unroll
  for (int j = 0; j < KW_MAXN; ++j) a[j] = (isp && j < ns) ? ww.A[lane][j] : 0.0;
  double2* buf = reinterpret_cast<double2*>(ww.ws);
  int lidx = lane;
  bool used = false;
#pragma unroll 1
  for (int k = 0; k < ns; ++k) {
    const int key = warp_argmax_abs(fabs(a[0]), isp && !used, lidx, lane);
    const int pl = key & 0xff, bi = key >> 8;   // the pivot row's lane / logical index
    if (lane != pl && lidx == k) lidx = bi;     // the exchange rows k <-> bi, in logical indices only
    if (lane == 0) ww.piv[k] = bi;
    const double pinv = 1.0 / __shfl_sync(0xffffffffu, a[0], pl);
    __syncwarp();
    if (lane == pl) {
      lidx = k; used = true;
#pragma unroll
      for (int m = 0; m < KW_MAXN - 1; ++m) a[m] = a[m + 1] * pinv;   // scaled and shifted
      a[KW_MAXN - 1] = pinv;
#pragma unroll
      for (int q = 0; q < KW_MAXN / 2; ++q) buf[q] = make_double2(a[2 * q], a[2 * q + 1]);
    }
    __syncwarp();
    if (lane != pl) {
      const double f = a[0];
#pragma unroll
      for (int q = 0; q < KW_MAXN / 2; ++q) {
        const double2 r = buf[q];
        a[2 * q] = fma(-f, r.x, a[2 * q + 1]);
        if (2 * q + 2 < KW_MAXN) a[2 * q + 1] = fma(-f, r.y, a[2 * q + 2]);
        else a[2 * q + 1] = -f * r.y;
      }
    }
  }
  __syncwarp();
  // after ns shifts ring slot m holds column m - (32 - ns)
  if (isp) {
#pragma unroll
    for (int m = 0; m < KW_MAXN; ++m) {
      const int c = m - (KW_MAXN - ns);
      if (c >= 0) ww.A[lidx][c] = a[m];
    }
  }
  __syncwarp();
}

// the ns x ns matrix in ww.A replaced by its inverse
template <class WW>
__device__ __forceinline__ void wide_invert(WW& ww, int lane, int ns) {
  const bool isp = lane < ns;
  if (ns > 16) wide_gj_regs(ww, lane, ns);
  else
  for (int k = 0; k < ns; ++k) {
    const int bi = warp_argmax_abs((lane >= k && isp) ? fabs(ww.A[lane][k]) : 0.0, lane >= k && isp, lane, lane) >> 8;
    if (bi != k && isp) { const double tmpv = ww.A[k][lane]; ww.A[k][lane] = ww.A[bi][lane]; ww.A[bi][lane] = tmpv; }
    if (lane == 0) ww.piv[k] = bi;
    __syncwarp();
    const double pinv = 1.0 / ww.A[k][k];
    __syncwarp();
    if (isp) ww.A[k][lane] = (lane == k) ? pinv : ww.A[k][lane] * pinv;   // lane = column: scale the pivot row
    __syncwarp();
    if (isp && lane != k) {                                                  // lane = row: eliminate column k
      const double f = ww.A[lane][k];
#pragma unroll 4
      for (int j = 0; j < ns; ++j) {
        const double akj = ww.A[k][j];
        ww.A[lane][j] = (j == k) ? -f * akj : fma(-f, akj, ww.A[lane][j]);
      }
    }
    __syncwarp();
  }
  for (int k = ns - 1; k >= 0; --k) {   // undo the row exchanges as column exchanges, in reverse order
    const int p = ww.piv[k];
    if (p != k && isp) { const double tmpv = ww.A[lane][k]; ww.A[lane][k] = ww.A[lane][p]; ww.A[lane][p] = tmpv; }
  }
  __syncwarp();
}

// W^{-1} in place of W (ww.A) by Gauss-Jordan elimination with partial pivoting, lane = row.  A solve then is one
// mat-vec with independent loads (wide_invmul) instead of 2*ns dependent shuffle + FMA steps (wide_lusolve): the
// simplified-Newton iterations of k_kencarp4_wide call it ~20 times per factorisation, and those dependent chains were
// what bound that kernel.  Mirrored by the oracle's named switch crnn_oracle_set_kc4_inverse.
template <bool F2, bool SPARSE = false, class WW>
__device__ __forceinline__ double wide_build_inv(const WideP& P, const WideBlock& sb, WW& ww, int lane,
                                                 const double* rsrc, const WideAux& a, double gdt) {
  const double eig = wide_assemble_W<F2, SPARSE>(P, sb, ww, lane, rsrc, a, gdt);   // opnorm(J, Inf)
  wide_invert(ww, lane, P.ns);
  return eig;
}

// b <- W^{-1} b with the explicit inverse in ww.A (lane i holds b_i); scratch: ww.ws
template <class WW>
__device__ __forceinline__ double wide_invmul(WW& ww, int lane, int ns, double b) {
  __syncwarp();
  ww.ws[lane] = b;
  __syncwarp();
  double s = 0.0;
  if (lane < ns) {
#pragma unroll 6
    for (int j = 0; j < ns; ++j) s = fma(ww.A[lane][j], ww.ws[j], s);
  }
  return s;
}

}  // namespace crnn
