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

struct alignas(16) WideBlock {
  double w_inT[KW_MAXN][KW_MAXN];  // [i][j]
  double w_inJ[KW_MAXN][KW_MAXN];  // [j][i]: the Jacobian assembly reads 8 consecutive inputs of a reaction with 4 LDS.128
  double w_out[KW_MAXN][KW_MAXN];  // [j][i]: lane i reads consecutive addresses for fixed j
  double w_b[KW_MAXN];
};

// the adjoint kernel assembles no Jacobian: no reaction-major copy of w_in (8 KB more room for its step record)
struct alignas(16) WideBlockLite {
  double w_inT[KW_MAXN][KW_MAXN];  // [i][j]
  double w_out[KW_MAXN][KW_MAXN];  // [j][i]
  double w_b[KW_MAXN];
};

struct WideAux {  // per-lane by-products of one RHS evaluation (what the Jacobian needs)
  double dx;       // d x_l / d u_l (F2: at fixed density)
  double rr;       // F2: d log(rho) / d u_l = -chi_l / (MW_l S)
  double wdot;     // sum_j w_out[l,j] r_j (scaled; F2: before the 1/rho)
  double inv_rho;  // F2
  double chiC;     // F2: 1 if lb <= C_l <= ub
};

// f(y, t): lane i holds y_i in, f_i out; leaves x in ww.x and r in ww.r.
template <bool F2, class WW>
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
#pragma unroll 2
    for (int i = 0; i < nin; ++i) z = fma(sb.w_inT[i][lane], ww.x[i], z);
    ww.r[lane] = lean_exp(z);
  }
  __syncwarp();
  double f = 0.0;
  if (isp) {
#pragma unroll 2
    for (int j = 0; j < nr; ++j) f = fma(sb.w_out[j][lane], ww.r[j], f);
  }
  a.wdot = f;
  if (F2) f = f / rho;
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
template <bool F2, class WW>
__device__ __forceinline__ double wide_assemble_W(const WideP& P, const WideBlock& sb, WW& ww, int lane,
                                                  const double* rsrc, const WideAux& a, double gdt) {
  const int n = P.n, ns = P.ns, nr = P.nr;
  const bool isp = lane < ns;
  __syncwarp();
  ww.bdx[lane] = a.dx; ww.brr[lane] = a.rr; ww.bchi[lane] = a.chiC;
  __syncwarp();
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

// W = I - gdt*J(u) from the RHS by-products (r in rsrc, this lane's aux), cooperative LU with partial
// pivoting (first strict maximum, like oracle lu_factor); returns opnorm(J, Inf).
template <bool F2, class WW>
__device__ __forceinline__ double wide_build_lu(const WideP& P, const WideBlock& sb, WW& ww, int lane,
                                                const double* rsrc, const WideAux& a, double gdt) {
  const int ns = P.ns;
  const bool isp = lane < ns;
  const double eig = wide_assemble_W<F2>(P, sb, ww, lane, rsrc, a, gdt);
  ww.perm[lane] = lane;
  __syncwarp();
  for (int k = 0; k < ns; ++k) {
    double best = (lane >= k && isp) ? fabs(ww.A[lane][k]) : -1.0;
    int bi = lane;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, m);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, m);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
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

// W^{-1} in place of W (ww.A) by Gauss-Jordan elimination with partial pivoting, lane = row.  A solve then is one
// mat-vec with independent loads (wide_invmul) instead of 2*ns dependent shuffle + FMA steps (wide_lusolve): the
// simplified-Newton iterations of k_kencarp4_wide call it ~20 times per factorisation, and those dependent chains were
// what bound that kernel.  Mirrored by the oracle's named switch crnn_oracle_set_kc4_inverse.
template <bool F2, class WW>
__device__ __forceinline__ double wide_build_inv(const WideP& P, const WideBlock& sb, WW& ww, int lane,
                                                 const double* rsrc, const WideAux& a, double gdt) {
  const int ns = P.ns;
  const bool isp = lane < ns;
  const double eig = wide_assemble_W<F2>(P, sb, ww, lane, rsrc, a, gdt);   // opnorm(J, Inf)
  for (int k = 0; k < ns; ++k) {
    double best = (lane >= k && isp) ? fabs(ww.A[lane][k]) : -1.0;
    int bi = lane;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, m);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, m);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
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
