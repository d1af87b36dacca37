// crnn_api.cu — the C-ABI of libcrnn_b200.so (declared in include/crnn_b200.h).
//
// Entry points only: validation and dispatch on (n_species, n_reac, rhs_kind) to the
// dimension-specialised engines, which are compiled one translation unit per configuration
// (inst.cu with -DCRNN_NS/-DCRNN_NR/-DCRNN_KIND) so the build parallelises.
#include "crnn_host.cuh"
#include "kernel_kencarp4_wide.cuh"
#include "kernel_tsit5_adjoint.cuh"
#include "kernel_wide_solve.cuh"
#include "kernel_gen_sens.cuh"

namespace crnn_host {
#define X(NS_, NR_, K_)                                                                                     \
  extern template int solve_impl<Cfg<NS_, NR_, K_>>(crnn_handle*, const crnn_model*, const crnn_opts*,      \
                                                    const HostIO&, int64_t);                                \
  extern template int loss_grad_impl<Cfg<NS_, NR_, K_>>(crnn_handle*, const crnn_model*, const crnn_opts*,  \
                                                        const double*, int, const double*, int,             \
                                                        const HostIO&, int64_t, double*, const AutoHook*);     \
  extern template int train_impl<Cfg<NS_, NR_, K_>>(crnn_handle*, const crnn_model*, const crnn_opts*,      \
                                                    const crnn_train_opts*, const crnn_dataset*, const int64_t*, int64_t, \
                                                    const double*, int32_t, double*, double*, double*, double*);
CRNN_FOR_EACH_CFG(X)
#undef X
}  // namespace crnn_host
using namespace crnn_host;

namespace {

// Fills the generic-dimension parameter block shared by the lane-per-component kernels and uploads
// its device arrays: w_inT [nin][32] | w_b | w_out (scaled) | saveat | row2obs | extra doubles.
int build_wide(crnn_handle* h, const crnn_model* m, const crnn_opts* o, int order, const std::vector<double>& extra,
               cudaStream_t st, WideP& P, const double** extra_dev, DevBuf* blob_buf = nullptr) {
  DevBuf& cfgbuf = blob_buf ? *blob_buf : h->cfg;
  if (m->n_state > KW_MAXN || m->n_reac > KW_MAXN || m->n_in > KW_MAXN)
    return fail(h, CRNN_ERR_UNSUPPORTED, "this solver's kernel supports n_state, n_in, n_reac <= 32");
  const int n = m->n_state, ns = m->n_species, nin = m->n_in, nr = m->n_reac;
  const bool dens = (m->rhs_kind == CRNN_RHS_F2_MASSFRAC_TP);
  const bool f2 = dens || m->rhs_kind == CRNN_RHS_F5_TRAMP;   // inputs from T(t) tables
  const int ntab = f2 ? m->n_tab : 0;
  std::vector<int> row2obs(n, -1);
  for (int q = 0; q < o->n_obs; ++q) {
    const int r = o->obs_idx[q];
    if (r < 0 || r >= n) return fail(h, CRNN_ERR_BAD_ARG, "obs_idx out of range");
    if (row2obs[r] >= 0) return fail(h, CRNN_ERR_BAD_ARG, "obs_idx has duplicates");
    row2obs[r] = q;
  }
  for (int i = 0; i < KW_MAXN; ++i) {
    P.abstol[i] = i < n ? o->abstol[o->n_abstol > 1 ? i : 0] : 1.0;
    P.reltol[i] = i < n ? o->reltol[o->n_reltol > 1 ? i : 0] : 0.0;
  }
  P.lb = m->lb; P.ub = m->ub; P.gas_R = m->gas_R;
  P.t0 = o->t0; P.t1 = o->t1; P.pred_lo = o->pred_clamp_lo; P.pred_hi = o->pred_clamp_hi;
  const double qmin = o->qmin > 0 ? o->qmin : 0.2, qmax = o->qmax > 0 ? o->qmax : 10.0;
  P.inv_qmin = 1.0 / qmin; P.inv_qmax = 1.0 / qmax;
  P.gamma = o->gamma > 0 ? o->gamma : 0.9;
  P.beta2 = o->beta2 > 0 ? o->beta2 : 2.0 / (5.0 * order);
  P.beta1 = o->beta1 > 0 ? o->beta1 : 7.0 / (10.0 * order);
  P.inv_order = 1.0 / order;
  {
    const bool implicit_alg = (o->alg == CRNN_ALG_ROSENBROCK23 || o->alg == CRNN_ALG_KENCARP4 || o->alg == CRNN_ALG_TRBDF2);
    P.qs_min = o->qsteady_min > 0 ? o->qsteady_min : 1.0;
    P.qs_max = o->qsteady_max > 0 ? o->qsteady_max : (implicit_alg ? 1.2 : 1.0);
  }
  P.maxiters = o->maxiters;
  P.n = n; P.ns = ns; P.nin = nin; P.nr = nr; P.kind = m->rhs_kind;
  P.n_save = o->n_save; P.n_obs = o->n_obs;
  P.alg = o->alg; P.n_tab = ntab;
  P.beta2_ros = o->beta2 > 0 ? o->beta2 : 2.0 / (5.0 * 2.0);
  P.beta1_ros = o->beta1 > 0 ? o->beta1 : 7.0 / (10.0 * 2.0);
  const size_t n_r2o = (n + 1) / 2 + 1;
  // F4: MLP parameters | w_J | mlp_in_idx, aug_src (ints packed behind)
  const bool f4 = (m->rhs_kind == CRNN_RHS_F4_MLP_AUG);
  size_t n_mlp_par = 0;
  if (f4) for (int l = 0; l < m->mlp_n_layers; ++l) n_mlp_par += (size_t)m->mlp_dims[l] * m->mlp_dims[l + 1] + m->mlp_dims[l + 1];
  const size_t n_mlp = f4 ? n_mlp_par + ns + (size_t)(m->mlp_dims[0] + nin + 1) / 2 + 2 : 0;
  std::vector<double> blob((size_t)nin * KW_MAXN + nr + (size_t)ns * nr + o->n_save + n_r2o + extra.size() + 2 +
                           (f2 ? ns + 3 * (size_t)ntab : 0) + (m->w_obs ? nr : 0) + n_mlp, 0.0);
  double* p_winT = blob.data();
  double* p_wb = p_winT + (size_t)nin * KW_MAXN;
  double* p_wout = p_wb + nr;
  double* p_save = p_wout + (size_t)ns * nr;
  int* p_r2o = reinterpret_cast<int*>(p_save + o->n_save);
  double* p_extra = p_save + o->n_save + n_r2o;
  for (int j = 0; j < nr; ++j) {
    for (int i = 0; i < nin; ++i) p_winT[(size_t)i * KW_MAXN + j] = m->w_in[i + nin * j];
    p_wb[j] = m->w_b[j];
    // out_scale (and, for F2, the molar mass of `wdot * l_MW / density * dydt_scale`) folded into the rows of w_out
    for (int i = 0; i < ns; ++i)
      p_wout[i + ns * j] = m->w_out[i + ns * j] * (dens ? m->mw[i] : 1.0) * (m->out_scale ? m->out_scale[i] : 1.0);
  }
  for (int k = 0; k < o->n_save; ++k) p_save[k] = o->saveat[k];
  for (int i = 0; i < n; ++i) p_r2o[i] = row2obs[i];
  for (size_t q = 0; q < extra.size(); ++q) p_extra[q] = extra[q];
  double* p_f2 = p_extra + extra.size() + 1;
  if (f2) {
    for (int i = 0; i < ns; ++i) p_f2[i] = dens ? m->mw[i] : 1.0;
    for (int k = 0; k < ntab; ++k) {
      p_f2[ns + k] = m->tab_t[k]; p_f2[ns + ntab + k] = m->tab_T[k]; p_f2[ns + 2 * ntab + k] = dens ? m->tab_P[k] : 1.0;
    }
  }
  double* p_obs = p_f2 + (f2 ? ns + 3 * (size_t)ntab : 0);
  if (m->w_obs) for (int j = 0; j < nr; ++j) p_obs[j] = m->w_obs[j];
  double* p_mlp = p_obs + (m->w_obs ? nr : 0);
  if (f4) {
    for (size_t q = 0; q < n_mlp_par; ++q) p_mlp[q] = m->mlp_params[q];
    for (int i = 0; i < ns; ++i) p_mlp[n_mlp_par + i] = m->w_J ? m->w_J[i] * (m->out_scale ? m->out_scale[i] : 1.0) : 0.0;
    int* pi = reinterpret_cast<int*>(p_mlp + n_mlp_par + ns);
    for (int i = 0; i < m->mlp_dims[0]; ++i) pi[i] = m->mlp_in_idx[i];
    for (int i = 0; i < nin; ++i) pi[m->mlp_dims[0] + i] = m->aug_src[i];
  }
  CK(cfgbuf.reserve(std::max<size_t>(blob.size() * sizeof(double), 4096)));
  CK(cudaMemcpyAsync(cfgbuf.p, blob.data(), blob.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  double* d = cfgbuf.as<double>();
  P.w_obs = m->w_obs ? d + (p_obs - blob.data()) : nullptr;
  if (f4) {
    const double* d_mlp = d + (p_mlp - blob.data());
    P.mlp_layers = m->mlp_n_layers; P.mlp_act_out = m->mlp_act_out;
    for (int l = 0; l <= m->mlp_n_layers; ++l) P.mlp_dims[l] = m->mlp_dims[l];
    P.mlp_params = d_mlp; P.w_J = m->w_J ? d_mlp + n_mlp_par : nullptr;
    P.mlp_in_idx = reinterpret_cast<const int*>(d_mlp + n_mlp_par + ns);
    P.aug_src = P.mlp_in_idx + m->mlp_dims[0];
  }
  P.w_inT = d; P.w_b = d + (p_wb - blob.data()); P.w_out = d + (p_wout - blob.data());
  P.saveat = d + (p_save - blob.data());
  P.row2obs = reinterpret_cast<const int*>(d + (p_save - blob.data()) + o->n_save);
  if (extra_dev) *extra_dev = d + (p_extra - blob.data());
  if (f2) {
    const double* df2 = d + (p_f2 - blob.data());
    P.mw = df2; P.tab_t = df2 + ns; P.tab_T = df2 + ns + ntab; P.tab_P = df2 + ns + 2 * ntab;
  }
  return CRNN_OK;
}

// Generic predict path (kernel_wide_solve.cuh): Tsit5 / Rosenbrock23 / AutoTsit5(Rosenbrock23) for any
// dimensions <= 32 and every RHS flavour.
int solve_wide(crnn_handle* h, const crnn_model* m, const crnn_opts* o, const HostIO& io, int64_t N) {
  if (m->w_obs && o->n_obs != 1) return fail(h, CRNN_ERR_BAD_ARG, "the observable post-map has one output (n_obs = 1)");
  WideP P{};
  cudaStream_t st = o->buffers_on_device ? (cudaStream_t)o->stream : h->s_compute;
  const int order = (o->alg == CRNN_ALG_ROSENBROCK23 || o->alg == CRNN_ALG_TRBDF2) ? 2 : 5;
  int rcw = build_wide(h, m, o, order, {}, st, P, nullptr);
  if (rcw) return rcw;
  constexpr int WARPS = 7;   // two blocks of seven warps per SM (shared memory: 24.8 KB per block + 12.4 KB per warp)
  const bool tab = (m->rhs_kind == CRNN_RHS_F2_MASSFRAC_TP || m->rhs_kind == CRNN_RHS_F5_TRAMP);
  const bool trb = (o->alg == CRNN_ALG_TRBDF2 || o->alg == CRNN_ALG_AUTO_TSIT5_TRBDF2);   // TRBDF2 as the stiff stepper
  if (m->w_obs && !tab) return fail(h, CRNN_ERR_UNSUPPORTED, "the observable post-map of the predict path is built for the tabulated-input flavours (F5: Cathode)");
  auto kern = (m->rhs_kind == CRNN_RHS_F4_MLP_AUG)   // MLP-augmented inputs (yeast / QSSA): own instantiations, finite-difference Jacobian
                  ? (trb ? k_wide_solve<WARPS, false, 1, false, true> : k_wide_solve<WARPS, false, 0, false, true>)
              : trb ? (tab ? (m->w_obs ? k_wide_solve<WARPS, true, 1, true> : k_wide_solve<WARPS, true, 1, false>) : k_wide_solve<WARPS, false, 1, false>)
                  : (tab ? (m->w_obs ? k_wide_solve<WARPS, true, 0, true> : k_wide_solve<WARPS, true, 0, false>) : k_wide_solve<WARPS, false, 0, false>);
  const size_t smem = sizeof(WideBlock) + WARPS * sizeof(WideWarp);
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int bps = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, WARPS * 32, smem));
  if (bps < 1) bps = 1;
  return run_batch(h, m, o, io, N, false, 0, nullptr, [&](const BatchPtrs& b, cudaStream_t s) -> int {
    if (b.n == 0) return (int)CRNN_OK;
    const long long want = (b.n + WARPS - 1) / WARPS;
    const unsigned blocks = (unsigned)std::min<long long>((long long)h->num_sms * bps, want);
    unsigned long long* queue = h->ctr.as<unsigned long long>() + b.qslot;
    CK(cudaMemsetAsync(queue, 0, sizeof(unsigned long long), s));
    ProfScope prof(h, s);
    kern<<<blocks, WARPS * 32, smem, s>>>(P, b.u0, b.nsu, b.n, b.pred, b.n_saved, b.retcode, b.stats, queue);
    CK(cudaGetLastError());
    h->launches++;
    return (int)CRNN_OK;
  });
}

// KenCarp4 (BASELINE config 5): generic-dimension warp-per-trajectory kernel, value path only.
int solve_kencarp4(crnn_handle* h, const crnn_model* m, const crnn_opts* o, const HostIO& io, int64_t N) {
  if (m->rhs_kind == CRNN_RHS_F5_TRAMP || m->rhs_kind == CRNN_RHS_F4_MLP_AUG || m->w_obs)
    return fail(h, CRNN_ERR_UNSUPPORTED, "KenCarp4 serves F0 / F1 / F2 without an observable post-map");
  WideP P{};
  cudaStream_t st = o->buffers_on_device ? (cudaStream_t)o->stream : h->s_compute;
  int rcw = build_wide(h, m, o, 4, {}, st, P, nullptr);
  if (rcw) return rcw;
  constexpr int WARPS = 8;   // two blocks of eight warps per SM
  // large models: the instantiation whose RHS mat-vecs run over the non-zero weights (wide_common.cuh, WideSparse)
  const bool big = m->n_in > 16 || m->n_reac > 16;
  auto kern = m->rhs_kind == CRNN_RHS_F2_MASSFRAC_TP ? (big ? k_kencarp4_wide<WARPS, true, true> : k_kencarp4_wide<WARPS, true, false>)
                                                     : (big ? k_kencarp4_wide<WARPS, false, true> : k_kencarp4_wide<WARPS, false, false>);
  const size_t smem = sizeof(WideBlock) + WARPS * sizeof(WideWarpT<0>);
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int bps = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, WARPS * 32, smem));
  if (bps < 1) bps = 1;
  return run_batch(h, m, o, io, N, false, 0, nullptr, [&](const BatchPtrs& b, cudaStream_t s) -> int {
    if (b.n == 0) return (int)CRNN_OK;
    const long long want = (b.n + WARPS - 1) / WARPS;
    const unsigned blocks = (unsigned)std::min<long long>((long long)h->num_sms * bps, want);
    unsigned long long* queue = h->ctr.as<unsigned long long>() + b.qslot;
    CK(cudaMemsetAsync(queue, 0, sizeof(unsigned long long), s));
    ProfScope prof(h, s);
    kern<<<blocks, WARPS * 32, smem, s>>>(P, b.u0, b.nsu, b.n, b.pred, b.n_saved, b.retcode, b.stats, queue);
    CK(cudaGetLastError());
    h->launches++;
    return (int)CRNN_OK;
  });
}


// Interpolating adjoint (BASELINE config 4): Tsit5, generic dimensions, any np (cost independent of np).
int loss_grad_adjoint(crnn_handle* h, const crnn_model* m, const crnn_opts* o, const double* dW_dp, int np,
                      const double* yscale, int loss_kind, const HostIO& io, int64_t N, double* grad_sum) {
  if (o->alg != CRNN_ALG_TSIT5) return fail(h, CRNN_ERR_UNSUPPORTED, "the adjoint is implemented for Tsit5");
  if (m->rhs_kind == CRNN_RHS_F5_TRAMP || m->w_obs || loss_kind == CRNN_LOSS_MSE)
    return fail(h, CRNN_ERR_UNSUPPORTED, "the adjoint kernels serve F0 / F1 / F2 with the MAE losses on the states (use sens_mode FORWARD)");
  const int n = m->n_state, ns = m->n_species, nr = m->n_reac;
  // F4: the quadrature runs in the extended weight space [vec(w_in); w_b; vec(w_out); w_J; mlp_params]
  const bool f4 = (m->rhs_kind == CRNN_RHS_F4_MLP_AUG);
  int nw = nr * (m->n_in + 1 + ns);
  if (f4) {
    nw += ns;
    for (int l = 0; l < m->mlp_n_layers; ++l) nw += m->mlp_dims[l] * m->mlp_dims[l + 1] + m->mlp_dims[l + 1];
  }
  if (nw > 32 * ADJ_MAX_ENT) return fail(h, CRNN_ERR_UNSUPPORTED, "adjoint kernel supports n_w <= 512");
  // extra device doubles: scale[ns] | inv_ys[n] | seed [nw*np]
  std::vector<double> extra((size_t)ns + n + (size_t)nw * np, 1.0);
  // scale[i] multiplies lambda_i in the w_out quadrature: out_scale, times the molar mass for F2 (f_i = wdot_i MW_i / rho s_i)
  for (int i = 0; i < ns; ++i)
    extra[i] = (m->out_scale ? m->out_scale[i] : 1.0) * (m->rhs_kind == CRNN_RHS_F2_MASSFRAC_TP ? m->mw[i] : 1.0);
  for (int q = 0; q < o->n_obs; ++q) {
    const int r = o->obs_idx[q];
    if (r >= 0 && r < n && loss_kind == CRNN_LOSS_MAE_SCALED) extra[ns + r] = 1.0 / yscale[q];
  }
  for (size_t q = 0; q < (size_t)nw * np; ++q) extra[(size_t)ns + n + q] = dW_dp[q];
  AdjP P{};
  cudaStream_t st = o->buffers_on_device ? (cudaStream_t)o->stream : h->s_compute;
  const double* extra_dev = nullptr;
  int rcw = build_wide(h, m, o, 5, extra, st, P.w, &extra_dev);
  if (rcw) return rcw;
  P.scale = extra_dev; P.inv_ys = extra_dev + ns;
  const double* seed_dev = extra_dev + ns + n;
  P.nw = nw; P.loss_kind = loss_kind;
  P.discrete = (o->sens_mode == CRNN_SENS_DISCRETE_ADJOINT) ? 1 : 0;
  if (f4) {   // per-warp MLP arrays packed to the widest layer; a per-block copy of the parameters
    int widest = 2, npar = 0;
    for (int l = 0; l <= m->mlp_n_layers; ++l) widest = std::max(widest, (int)m->mlp_dims[l]);
    for (int l = 0; l < m->mlp_n_layers; ++l) npar += m->mlp_dims[l] * m->mlp_dims[l + 1] + m->mlp_dims[l + 1];
    P.mlp_stride = (widest + 1) & ~1;
    P.mlp_extra = 32 + (3 * m->mlp_n_layers + 1) * P.mlp_stride;
    P.mlp_np_raw = npar;
    P.mlp_np = (npar + 1) & ~1;
  }
  P.gs_len = P.discrete ? 0 : ((nw + 1) & ~1);   // the stage accumulator of the interpolating adjoint's quadrature
  constexpr int WARPS = 4;
  // blocks per SM: four (16 warps, 128 registers with 116 B spilled) unless CRNN_B200_ADJ_BLOCKS=2 asks for the two-block build
  static const bool want_four = [] { const char* e = std::getenv("CRNN_B200_ADJ_BLOCKS"); return !e || std::atoi(e) != 2; }();
  const bool f2 = m->rhs_kind == CRNN_RHS_F2_MASSFRAC_TP;
  const int stride = 8 * n + 2;
  const size_t fixed_pw = (6 * 32 + 2 + (size_t)P.mlp_extra + (size_t)((nw + 1) & ~1) + (size_t)P.gs_len) * sizeof(double);   // kernel_tsit5_adjoint.cuh: ADJ_FIXED
  const size_t block_fixed = sizeof(WideBlockLite) + (size_t)P.mlp_np * sizeof(double);
  // forward-record capacity in shared memory for `nb` blocks of 4 warps per SM (steps per warp; < 0: does not fit)
  auto cap_for = [&](int nb) -> long long {
    const size_t budget = (size_t)(227 * 1024 / nb) - 2048 - block_fixed;
    if (fixed_pw * WARPS > budget) return -1;
    return (long long)std::min<size_t>(256, (budget / WARPS - fixed_pw) / (stride * sizeof(double)));
  };
  // large models (n_w towards 512) leave a four-block build no room for the record: they keep two blocks per SM
  const bool four_blocks = want_four && cap_for(4) >= 4;
  // F4: the MLP state adds (3L + 2)*32 doubles per warp; three blocks (12 warps per SM) with a short record beat two with a long one
  static const int f4_blocks_env = [] { const char* e = std::getenv("CRNN_B200_ADJ_F4_BLOCKS"); return e ? std::atoi(e) : 0; }();
  const bool three_blocks = f4 && !four_blocks && f4_blocks_env != 2 && cap_for(3) >= 2;
  const long long cap = cap_for(four_blocks ? 4 : three_blocks ? 3 : 2);
  if (cap < 1) return fail(h, CRNN_ERR_UNSUPPORTED, "model too large for the adjoint kernel's shared memory");
  auto kern = f4 ? (four_blocks ? k_tsit5_adjoint<WARPS, false, 4, true> : three_blocks ? k_tsit5_adjoint<WARPS, false, 3, true>
                                                                                        : k_tsit5_adjoint<WARPS, false, 2, true>)
              : four_blocks ? (f2 ? k_tsit5_adjoint<WARPS, true, 4> : k_tsit5_adjoint<WARPS, false, 4>)
                            : (f2 ? k_tsit5_adjoint<WARPS, true, 2> : k_tsit5_adjoint<WARPS, false, 2>);
  P.cap_s = (int)cap;
  P.cap_g = 512;
  const size_t smem = block_fixed + WARPS * (fixed_pw + (size_t)P.cap_s * stride * sizeof(double));
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int bps = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, WARPS * 32, smem));
  if (bps < 1) bps = 1;
  const long long max_blocks = (long long)h->num_sms * bps;
  // overflow record (steps beyond cap_s): one region per block and warp, and one copy PER PIPELINE SLOT — the chunks of the
  // host-buffer path run their kernels concurrently on the slots' streams and must not share it
  const size_t scratch_doubles = (size_t)max_blocks * WARPS * P.cap_g * stride;
  CK(h->adj_scratch.reserve(scratch_doubles * (o->buffers_on_device ? 1 : kPipe) * sizeof(double)));
  P.scratch = h->adj_scratch.as<double>();
  PostFn post = [=](const double* red, double* out, cudaStream_t s) -> int {
    k_seed_contract<<<1, 256, 0, s>>>(seed_dev, red, nw, np, out);
    CK(cudaGetLastError());
    h->launches++;
    return CRNN_OK;
  };
  return run_batch(h, m, o, io, N, true, nw, grad_sum, [&](const BatchPtrs& b, cudaStream_t s) -> int {
    if (b.n == 0) return (int)CRNN_OK;
    const long long want = (b.n + WARPS - 1) / WARPS;
    const unsigned blocks = (unsigned)std::min<long long>(max_blocks, want);
    unsigned long long* queue = h->ctr.as<unsigned long long>() + b.qslot;
    CK(cudaMemsetAsync(queue, 0, sizeof(unsigned long long), s));
    AdjP Pl = P;
    if (b.qslot >= 2) Pl.scratch = P.scratch + scratch_doubles * (size_t)(b.qslot - 2);
    ProfScope prof(h, s);
    kern<<<blocks, WARPS * 32, smem, s>>>(Pl, b.u0, b.nsu, b.n, b.data, b.loss, b.grad_each, b.pred, b.n_saved,
                                       b.retcode, b.stats, queue, b.in_idx);
    CK(cudaGetLastError());
    h->launches++;
    return (int)CRNN_OK;
  }, np, post);
}

// Structured ("R1") form of one seed matrix for the generic forward kernel: every column touches w_in in at most ONE row
// and w_out in at most ONE entry - the shape of every p2vec of the reference.  rows: [(2 or 3)*nr][cols]
// (a_j = dW_in[i_in, j]; b_j = db_j; with an observable post-map also d w_obs_j), desc: [cols].
int gen_plan_seed(crnn_handle* h, const crnn_model* m, const double* dW_dp, int np, int cols, bool has_obs,
                  R1Desc* desc, double* rows) {
  const int ns = m->n_species, nin = m->n_in, nr = m->n_reac;
  const bool dens = (m->rhs_kind == CRNN_RHS_F2_MASSFRAC_TP);
  const int off_b = nin * nr, off_out = off_b + nr, off_obs = off_out + ns * nr, nw = off_obs + (has_obs ? nr : 0);
  for (int c = 0; c < np; ++c) {
    const double* s = dW_dp + (size_t)nw * c;
    R1Desc d{};
    int i_in = -1, n_out = 0;
    for (int j = 0; j < nr; ++j) {
      for (int i = 0; i < nin; ++i)
        if (s[i + nin * j] != 0.0) {
          if (i_in >= 0 && i_in != i) return fail(h, CRNN_ERR_UNSUPPORTED, "the generic forward-sensitivity kernel needs structured seed columns (one w_in row per parameter)");
          i_in = i;
        }
      for (int i = 0; i < ns; ++i)
        if (s[off_out + i + ns * j] != 0.0) {
          if (++n_out > 1) return fail(h, CRNN_ERR_UNSUPPORTED, "the generic forward-sensitivity kernel needs structured seed columns (one w_out entry per parameter)");
          d.i_out = i; d.j_out = j;
          d.o = s[off_out + i + ns * j] * (m->out_scale ? m->out_scale[i] : 1.0) * (dens ? m->mw[i] : 1.0);
        }
    }
    d.i_in = i_in < 0 ? 0 : i_in;
    desc[c] = d;
    for (int j = 0; j < nr; ++j) {
      rows[(size_t)j * cols + c] = i_in < 0 ? 0.0 : s[i_in + nin * j];
      rows[(size_t)(nr + j) * cols + c] = s[off_b + j];
      if (has_obs) rows[(size_t)(2 * nr + j) * cols + c] = s[off_obs + j];
    }
  }
  return CRNN_OK;
}

int gen_launch_cfg(crnn_handle* h, const crnn_model* m, int np, int& cols, size_t& smem, long long& max_blocks, bool f2k, bool trb,
                   void (**kern_out)(GenP, const double*, const int*, long long, const double*, double*, double*, double*, int*,
                                     int*, crnn_stats*, unsigned long long*, const long long*, const long long*,
                                     const unsigned int*)) {
  if (np < 1 || np > 255) return fail(h, CRNN_ERR_UNSUPPORTED, "the generic forward-sensitivity kernel supports 1 <= np <= 255");
  if (m->n_state > KW_MAXN || m->n_reac > KW_MAXN || m->n_in > KW_MAXN)
    return fail(h, CRNN_ERR_UNSUPPORTED, "this solver's kernel supports n_state, n_in, n_reac <= 32");
  cols = 32 * ((np + 31) / 32);
  smem = sizeof(GenShared) + (size_t)9 * m->n_state * cols * sizeof(double);
  if (smem > 227 * 1024)
    return fail(h, CRNN_ERR_UNSUPPORTED, "n_state * np too large for the generic forward-sensitivity kernel's shared memory");
  auto kern = trb ? (f2k ? k_gen_sens<true, 1> : k_gen_sens<false, 1>) : (f2k ? k_gen_sens<true, 0> : k_gen_sens<false, 0>);
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int bps = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, cols, smem));
  if (bps < 1) bps = 1;
  max_blocks = (long long)h->num_sms * bps;
  *kern_out = kern;
  return CRNN_OK;
}

// Generic forward sensitivities (kernel_gen_sens.cuh): any dimensions <= 32, F0 / F1 / F2 / F5, Tsit5 / Rosenbrock23 /
// AutoTsit5(Rosenbrock23), np <= 255, structured seed columns, optional observable post-map, all three losses.
// Returns CRNN_ERR_UNSUPPORTED when the model cannot be served so that the caller can try the adjoint route.
struct GenLaunch {
  GenP G{};
  void (*kern)(GenP, const double*, const int*, long long, const double*, double*, double*, double*, int*, int*, crnn_stats*,
               unsigned long long*, const long long*, const long long*, const unsigned int*) = nullptr;
  int cols = 0; size_t smem = 0; long long max_blocks = 0;
};

int gen_enqueue(crnn_handle* h, const GenLaunch& L, const BatchPtrs& b, cudaStream_t s, const long long* sel,
                const unsigned int* sel_count) {
  if (b.n == 0) return CRNN_OK;
  const unsigned blocks = (unsigned)std::min<long long>(L.max_blocks, b.n);
  // the hand-over launch gets its own work-queue counter (slot 5 + qslot would collide: use the high half of the ctr block)
  unsigned long long* queue = h->ctr.as<unsigned long long>() + (sel ? 8 + b.qslot : b.qslot);
  CK(cudaMemsetAsync(queue, 0, sizeof(unsigned long long), s));
  ProfScope prof(h, s);
  L.kern<<<blocks, L.cols, L.smem, s>>>(L.G, b.u0, b.nsu, b.n, b.data, b.loss, b.grad_each, b.pred, b.n_saved, b.retcode,
                                        b.stats, queue, b.in_idx, sel, sel_count);
  CK(cudaGetLastError());
  h->launches++;
  return CRNN_OK;
}

int gen_prepare(crnn_handle* h, const crnn_model* m, const crnn_opts* o, const double* dW_dp, int np,
                const double* yscale, int loss_kind, cudaStream_t st, DevBuf* blob_buf, GenLaunch& L);

int loss_grad_generic(crnn_handle* h, const crnn_model* m, const crnn_opts* o, const double* dW_dp, int np,
                      const double* yscale, int loss_kind, const HostIO& io, int64_t N, double* grad_sum) {
  GenLaunch L;
  cudaStream_t st = o->buffers_on_device ? (cudaStream_t)o->stream : h->s_compute;
  int rc = gen_prepare(h, m, o, dW_dp, np, yscale, loss_kind, st, nullptr, L);
  if (rc) return rc;
  return run_batch(h, m, o, io, N, true, np, grad_sum,
                   [&](const BatchPtrs& b, cudaStream_t s) -> int { return gen_enqueue(h, L, b, s, nullptr, nullptr); });
}

int gen_prepare(crnn_handle* h, const crnn_model* m, const crnn_opts* o, const double* dW_dp, int np,
                const double* yscale, int loss_kind, cudaStream_t st, DevBuf* blob_buf, GenLaunch& L) {
  const bool trb = (o->alg == CRNN_ALG_TRBDF2 || o->alg == CRNN_ALG_AUTO_TSIT5_TRBDF2);
  if (o->alg != CRNN_ALG_TSIT5 && o->alg != CRNN_ALG_ROSENBROCK23 && o->alg != CRNN_ALG_AUTO_TSIT5_ROS23 && !trb)
    return fail(h, CRNN_ERR_UNSUPPORTED, "forward sensitivities are implemented for Tsit5, Rosenbrock23, TRBDF2 and the AutoTsit5 composites");
  if (trb && m->rhs_kind == CRNN_RHS_F2_MASSFRAC_TP)
    return fail(h, CRNN_ERR_UNSUPPORTED, "sensitivities through TRBDF2 are built for F0 / F1 / F5 (the Cathode scripts); the HyChem flavour uses Rosenbrock23 (crnn_pyrolysis_mass.jl:29)");
  const int n = m->n_state, nr = m->n_reac;
  const bool f2k = (m->rhs_kind == CRNN_RHS_F2_MASSFRAC_TP || m->rhs_kind == CRNN_RHS_F5_TRAMP);
  const bool has_obs = m->w_obs != nullptr;
  int cols = 0; size_t smem = 0; long long max_blocks = 0;
  void (*kern)(GenP, const double*, const int*, long long, const double*, double*, double*, double*, int*, int*, crnn_stats*,
               unsigned long long*, const long long*, const long long*, const unsigned int*) = nullptr;
  int rc = gen_launch_cfg(h, m, np, cols, smem, max_blocks, f2k, trb, &kern);
  if (rc) return rc;
  const int nrow = (has_obs ? 3 : 2) * nr;
  std::vector<R1Desc> desc(cols, R1Desc{});
  std::vector<double> rows((size_t)nrow * cols, 0.0);
  rc = gen_plan_seed(h, m, dW_dp, np, cols, has_obs, desc.data(), rows.data());
  if (rc) return rc;
  // extra device doubles: inv_ys[n] | rows | desc[cols] (3 doubles each) | w_obs[nr]
  static_assert(sizeof(R1Desc) == 24, "R1Desc is packed as three doubles");
  std::vector<double> extra((size_t)n + rows.size() + 3 * (size_t)cols + nr, 1.0);
  if (has_obs) { if (loss_kind != CRNN_LOSS_MAE_LOG && yscale) extra[0] = 1.0 / yscale[0]; }
  else
    for (int q = 0; q < o->n_obs; ++q) {
      const int r = o->obs_idx[q];
      if (r >= 0 && r < n && loss_kind != CRNN_LOSS_MAE_LOG) extra[r] = 1.0 / yscale[q];
    }
  std::memcpy(extra.data() + n, rows.data(), rows.size() * sizeof(double));
  std::memcpy(extra.data() + n + rows.size(), desc.data(), desc.size() * sizeof(R1Desc));
  for (int j = 0; j < nr; ++j) extra[(size_t)n + rows.size() + 3 * (size_t)cols + j] = has_obs ? m->w_obs[j] : 0.0;
  GenP& G = L.G;
  const double* extra_dev = nullptr;
  const int order = (o->alg == CRNN_ALG_ROSENBROCK23 || o->alg == CRNN_ALG_TRBDF2) ? 2 : 5;
  int rcw = build_wide(h, m, o, order, extra, st, G.w, &extra_dev, blob_buf);
  if (rcw) return rcw;
  G.inv_ys = extra_dev; G.seed_rows = extra_dev + n;
  G.desc = reinterpret_cast<const R1Desc*>(extra_dev + n + rows.size());
  G.w_obs = has_obs ? extra_dev + n + rows.size() + 3 * (size_t)cols : nullptr;
  G.np = np; G.cols = cols; G.loss_kind = loss_kind; G.incl_sens = o->err_norm_includes_sens ? 1 : 0;
  G.norm_cnt = (double)n * ((o->err_norm_includes_sens && !o->err_norm_mean_over_state_only) ? (double)(np + 1) : 1.0);
  L.kern = kern; L.cols = cols; L.smem = smem; L.max_blocks = max_blocks;
  return CRNN_OK;
}

// grad[c + np*p] = sum over the experiments e of grad_each[(e + E*p)*np + c], in experiment order (deterministic)
__global__ void k_particle_reduce(const double* __restrict__ grad_each, int np, int E, int P, double* __restrict__ grad) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= np * P) return;
  const int c = q % np, p = q / np;
  double s = 0.0;
  for (int e = 0; e < E; ++e) s += grad_each[((size_t)e + (size_t)E * p) * np + c];
  grad[q] = s;
}

// lean_math.h on the device, elementwise (crnn_debug_lean_math)
__global__ void k_lean_math(int op, const double* __restrict__ x, const double* __restrict__ y2, double* __restrict__ y, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = x[i];
  y[i] = op == 0 ? lean_log(v) : op == 1 ? lean_exp(v) : op == 2 ? lean_pow(v, y2[i]) : op == 3 ? lean_log10(v) : lean_exp10(v);
}

}  // namespace

namespace crnn_host {
// Validation + dispatch shared by crnn_loss_grad_batch and crnn_loss_grad_indexed (crnn_dataset.cu).
int loss_grad_core(crnn_handle* h, const crnn_model* m, const crnn_opts* o, const double* dW_dp, int32_t np,
                   const double* yscale, int32_t loss_kind, const HostIO& io, int64_t N, double* grad_sum) {
  int rc = validate(h, m, o, N);
  if (rc) return rc;
  if (m->rhs_kind == CRNN_RHS_F4_MLP_AUG && o->sens_mode != CRNN_SENS_INTERP_ADJOINT && o->sens_mode != CRNN_SENS_DISCRETE_ADJOINT)
    return fail(h, CRNN_ERR_UNSUPPORTED, "gradients of the MLP-augmented RHS (F4) come from the adjoint sens_modes (Tsit5): forward mode is not built");
  if (N > 0 && (!io.u0 || !io.data || !io.loss)) return fail(h, CRNN_ERR_BAD_ARG, "null u0/data/loss");
  if (np < 0 || (np > 0 && !dW_dp)) return fail(h, CRNN_ERR_BAD_ARG, "bad seed matrix");
  if (loss_kind != CRNN_LOSS_MAE_SCALED && loss_kind != CRNN_LOSS_MAE_LOG && loss_kind != CRNN_LOSS_MSE)
    return fail(h, CRNN_ERR_BAD_ARG, "bad loss_kind");
  if (loss_kind != CRNN_LOSS_MAE_LOG && !yscale) return fail(h, CRNN_ERR_BAD_ARG, "null yscale");
  if (o->sens_mode != CRNN_SENS_FORWARD && o->sens_mode != CRNN_SENS_INTERP_ADJOINT &&
      o->sens_mode != CRNN_SENS_DISCRETE_ADJOINT)
    return fail(h, CRNN_ERR_UNSUPPORTED, "sens_mode must be FORWARD, INTERP_ADJOINT or DISCRETE_ADJOINT");
  if (o->n_obs == 0 || o->n_save == 0) return fail(h, CRNN_ERR_BAD_ARG, "loss needs n_obs > 0 and n_save > 0");
  CK(cudaSetDevice(h->device));
  h->last_grad_np = -1; h->last_grad_n = -1;
  if (o->sens_mode == CRNN_SENS_INTERP_ADJOINT || o->sens_mode == CRNN_SENS_DISCRETE_ADJOINT)
    return loss_grad_adjoint(h, m, o, dW_dp, np, yscale, loss_kind, io, N, grad_sum);
  h->last_grad_np = np; h->last_grad_n = N;
  // the dimension-specialised warp-per-trajectory kernels where they exist (Tsit5: np <= 255; Rosenbrock23: n_species <= 6,
  // np <= 63) ...
  const char* force = std::getenv("CRNN_B200_FORCE_GENERIC");
  const bool spec_alg = o->alg == CRNN_ALG_TSIT5 || (o->alg == CRNN_ALG_ROSENBROCK23 && m->n_species <= 6 && np <= 63);
  if (spec_alg && (m->rhs_kind == CRNN_RHS_F0 || m->rhs_kind == CRNN_RHS_F1_ARRH_TSTATE) && !m->w_obs &&
      loss_kind != CRNN_LOSS_MSE && !(force && force[0] == '1')) {
#define X(NS_, NR_, K_)                                                              \
  if (m->n_species == NS_ && m->n_reac == NR_ && m->rhs_kind == K_)                  \
    return loss_grad_impl<Cfg<NS_, NR_, K_>>(h, m, o, dW_dp, np, yscale, loss_kind, io, N, grad_sum, nullptr);
    CRNN_FOR_EACH_CFG(X)
#undef X
  }
  // AutoTsit5(Rosenbrock23) on a model with a specialised kernel (case2.jl:26 as written): Tsit5 with the AutoSwitch monitor,
  // the trajectories that would switch handed over to the generic composite kernel on the same stream
  if (o->alg == CRNN_ALG_AUTO_TSIT5_ROS23 && (m->rhs_kind == CRNN_RHS_F0 || m->rhs_kind == CRNN_RHS_F1_ARRH_TSTATE) && !m->w_obs &&
      loss_kind != CRNN_LOSS_MSE && np >= 1 && np <= 63 && !(force && force[0] == '1')) {
    bool have_cfg = false;
#define X(NS_, NR_, K_) have_cfg = have_cfg || (m->n_species == NS_ && m->n_reac == NR_ && m->rhs_kind == K_);
    CRNN_FOR_EACH_CFG(X)
#undef X
    GenLaunch L;
    cudaStream_t stg = o->buffers_on_device ? (cudaStream_t)o->stream : h->s_compute;
    if (have_cfg && gen_prepare(h, m, o, dW_dp, np, yscale, loss_kind, stg, &h->cfg2, L) == CRNN_OK) {
      AutoHook hook;
      hook.stride = (size_t)std::max<int64_t>(N, 1);
      const size_t nslot = 2 + kPipe;
      CK(h->auto_sel.reserve(nslot * hook.stride * sizeof(long long) + 64));
      hook.count = h->auto_sel.as<unsigned int>();
      hook.sel = reinterpret_cast<long long*>(h->auto_sel.as<char>() + 64);
      hook.fallback = [h, &L](const BatchPtrs& b, cudaStream_t s, const long long* sel, const unsigned int* cnt) -> int {
        return gen_enqueue(h, L, b, s, sel, cnt);
      };
      int rca = CRNN_ERR_UNSUPPORTED;
#define X(NS_, NR_, K_)                                                              \
  if (m->n_species == NS_ && m->n_reac == NR_ && m->rhs_kind == K_)                  \
    rca = loss_grad_impl<Cfg<NS_, NR_, K_>>(h, m, o, dW_dp, np, yscale, loss_kind, io, N, grad_sum, &hook);
      CRNN_FOR_EACH_CFG(X)
#undef X
      if (rca != CRNN_ERR_UNSUPPORTED) return rca;
    }
  }
  // ... and the generic block-per-trajectory kernel for everything else: gradients through AutoTsit5(Rosenbrock23), stiff
  // gradients of F2 and of models with more than 6 species, any (n_species, n_reac) <= 32
  {
    const int rcg = loss_grad_generic(h, m, o, dW_dp, np, yscale, loss_kind, io, N, grad_sum);
    if (rcg != CRNN_ERR_UNSUPPORTED) return rcg;
  }
  // Not servable in forward mode (dense seed, np > 255, shared memory).  With the value-only error norm the forward-mode
  // gradient IS the derivative of the recorded step sequence, i.e. what the discrete adjoint computes.
  h->last_grad_np = -1; h->last_grad_n = -1;
  if (o->alg == CRNN_ALG_TSIT5 && !o->err_norm_includes_sens) {
    crnn_opts oa = *o;
    oa.sens_mode = CRNN_SENS_DISCRETE_ADJOINT;
    return loss_grad_adjoint(h, m, &oa, dW_dp, np, yscale, loss_kind, io, N, grad_sum);
  }
  return fail(h, CRNN_ERR_UNSUPPORTED, "forward mode cannot serve this call (" + h->err + "): use an adjoint sens_mode, or "
              "err_norm_includes_sens = 0 with Tsit5 (served by the discrete adjoint)");
}
}  // namespace crnn_host

extern "C" {

int crnn_debug_lean_math(crnn_handle* h, int32_t op, const double* x, const double* x2, double* y, int64_t n) {
  if (!h || !x || !y || n < 0 || op < 0 || op > 4 || (op == 2 && !x2)) return CRNN_ERR_BAD_ARG;
  if (n == 0) return CRNN_OK;
  CK(cudaSetDevice(h->device));
  DevBuf bx, by, b2;
  CK(bx.reserve(n * sizeof(double))); CK(by.reserve(n * sizeof(double)));
  // same stream as the kernel: a synchronous cudaMemcpy from pageable memory may return before its DMA has landed, and
  // s_compute (non-blocking) does not wait for the legacy default stream
  CK(cudaMemcpyAsync(bx.p, x, n * sizeof(double), cudaMemcpyHostToDevice, h->s_compute));
  if (op == 2) { CK(b2.reserve(n * sizeof(double))); CK(cudaMemcpyAsync(b2.p, x2, n * sizeof(double), cudaMemcpyHostToDevice, h->s_compute)); }
  k_lean_math<<<(unsigned)((n + 255) / 256), 256, 0, h->s_compute>>>(op, bx.as<double>(), b2.as<double>(), by.as<double>(), n);
  CK(cudaGetLastError());
  h->launches++;
  CK(cudaMemcpyAsync(y, by.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->s_compute));
  CK(cudaStreamSynchronize(h->s_compute));
  bx.release(); by.release(); b2.release();
  return CRNN_OK;
}

int crnn_version(void) { return CRNN_B200_VERSION; }

int crnn_create(crnn_handle** out, int device_id) {
  if (!out) return CRNN_ERR_BAD_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return CRNN_ERR_NO_DEVICE;
  int dev = device_id;
  if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) return CRNN_ERR_NO_DEVICE;
  if (dev >= ndev) return CRNN_ERR_BAD_ARG;
  if (cudaSetDevice(dev) != cudaSuccess) return CRNN_ERR_CUDA;
  crnn_handle* h = new crnn_handle();
  h->device = dev;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { delete h; return CRNN_ERR_CUDA; }
  h->num_sms = prop.multiProcessorCount;
  bool ok = cudaStreamCreateWithFlags(&h->s_compute, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&h->ev_cfg, cudaEventDisableTiming) == cudaSuccess;
  for (int s = 0; s < kPipe && ok; ++s) ok = cudaStreamCreateWithFlags(&h->s_slot[s], cudaStreamNonBlocking) == cudaSuccess;
  for (int s = 0; s < kPipe && ok; ++s)
    ok = cudaEventCreateWithFlags(&h->ev_in[s], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&h->ev_done[s], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&h->ev_out[s], cudaEventDisableTiming) == cudaSuccess;
  ok = ok && h->ctr.reserve(256) == cudaSuccess && cudaMemset(h->ctr.p, 0, 256) == cudaSuccess;
  if (!ok) { crnn_destroy(h); return CRNN_ERR_CUDA; }
  *out = h;
  return CRNN_OK;
}

void crnn_destroy(crnn_handle* h) {
  if (!h) return;
  if (!h->kids.empty()) {  // multi-device parent: communicators first, then the per-device handles
    multi_release_comms(h);
    for (crnn_handle* k : h->kids) crnn_destroy(k);
    delete h;
    return;
  }
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  DevBuf* bufs[] = {&h->cfg, &h->seed, &h->desc, &h->ctr, &h->partial, &h->d_grad_each, &h->d_grad_sum, &h->d_grad_out, &h->adj_scratch,
                    &h->d_loss, &h->d_nsaved, &h->d_ret, &h->d_stats, &h->d_idx, &h->d_nsu_ix, &h->d_result, &h->cfg2, &h->auto_sel, &h->train};
  for (DevBuf* b : bufs) b->release();
  for (int s = 0; s < kPipe; ++s) {
    DevBuf* sb[] = {&h->d_u0[s], &h->d_nsu[s], &h->d_data[s], &h->d_pred[s]};
    for (DevBuf* b : sb) b->release();
    if (h->ev_in[s]) cudaEventDestroy(h->ev_in[s]);
    if (h->ev_done[s]) cudaEventDestroy(h->ev_done[s]);
    if (h->ev_out[s]) cudaEventDestroy(h->ev_out[s]);
    if (h->s_slot[s]) cudaStreamDestroy(h->s_slot[s]);
  }
  for (auto& pr : h->prof_events) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  if (h->ev_cfg) cudaEventDestroy(h->ev_cfg);
  if (h->s_compute) cudaStreamDestroy(h->s_compute);
  if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
  if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
  delete h;
}

const char* crnn_last_error(const crnn_handle* h) { return h ? h->err.c_str() : "null handle"; }

int64_t crnn_launch_count(const crnn_handle* h) {
  if (!h) return 0;
  int64_t n = h->launches;
  for (const crnn_handle* k : h->kids) n += k->launches;
  return n;
}

int crnn_copy_grad_each(crnn_handle* h, double* dst, int64_t N, int32_t np, int32_t on_device, void* stream) {
  if (!h || !dst || N < 0 || np <= 0) return CRNN_ERR_BAD_ARG;
  if (!h->kids.empty()) return fail(h, CRNN_ERR_UNSUPPORTED, "crnn_copy_grad_each needs a single-device handle");
  const size_t bytes = (size_t)N * np * sizeof(double);
  if (h->last_grad_np != np || h->last_grad_n != N || h->d_grad_each.cap < bytes)
    return fail(h, CRNN_ERR_BAD_ARG, "no forward-mode gradients of that shape from the last call");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = on_device ? (cudaStream_t)stream : h->s_compute;
  CK(cudaMemcpyAsync(dst, h->d_grad_each.p, bytes, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
  if (!on_device) CK(cudaStreamSynchronize(st));
  return CRNN_OK;
}

int crnn_profile_begin(crnn_handle* h) {
  if (!h) return CRNN_ERR_BAD_ARG;
  for (crnn_handle* k : h->kids) crnn_profile_begin(k);
  h->profiling = true;
  h->prof_used = 0;
  return CRNN_OK;
}

int crnn_profile_end(crnn_handle* h, double* total_ms, int64_t* n_launches) {
  if (!h) return CRNN_ERR_BAD_ARG;
  h->profiling = false;
  double tot = 0.0;
  int64_t nk = 0;
  for (crnn_handle* k : h->kids) {  // multi-device parent: sums over its devices
    double ms = 0.0; int64_t n = 0;
    int rc = crnn_profile_end(k, &ms, &n);
    if (rc) { h->err = k->err; return rc; }
    tot += ms; nk += n;
  }
  for (size_t q = 0; q < h->prof_used; ++q) {
    float ms = 0.f;
    CK(cudaEventSynchronize(h->prof_events[q].second));
    CK(cudaEventElapsedTime(&ms, h->prof_events[q].first, h->prof_events[q].second));
    tot += ms;
  }
  if (total_ms) *total_ms = tot;
  if (n_launches) *n_launches = (int64_t)h->prof_used + nk;
  h->prof_used = 0;
  return CRNN_OK;
}

int crnn_train_steps(crnn_handle* h, const crnn_model* m, const crnn_opts* o, const crnn_train_opts* t,
                     const crnn_dataset* ds, const int64_t* order, int64_t n_steps, const double* yscale,
                     int32_t loss_kind, double* p, double* opt_state, double* step_loss, double* step_gnorm) {
  if (!h) return CRNN_ERR_BAD_ARG;
  if (!h->kids.empty()) return fail(h, CRNN_ERR_UNSUPPORTED, "crnn_train_steps needs a single-device handle");
  if (!m || !o || !t || !ds || !p || !opt_state || (n_steps > 0 && !order)) return fail(h, CRNN_ERR_BAD_ARG, "crnn_train_steps: null argument");
  if (ds->owner != h) return fail(h, CRNN_ERR_BAD_ARG, "dataset does not belong to this handle");
  if (m->n_state != ds->n_state || o->n_obs != ds->n_obs || o->n_save != ds->n_save)
    return fail(h, CRNN_ERR_BAD_ARG, "model / opts do not match the dataset's (n_state, n_obs, n_save)");
  if (loss_kind != CRNN_LOSS_MAE_SCALED && loss_kind != CRNN_LOSS_MAE_LOG) return fail(h, CRNN_ERR_UNSUPPORTED, "loss_kind");
  if (loss_kind == CRNN_LOSS_MAE_SCALED && !yscale) return fail(h, CRNN_ERR_BAD_ARG, "null yscale");
  {  // validate() wants weight pointers: borrow a zero block
    std::vector<double> zw((size_t)m->n_reac * (m->n_in + 1 + m->n_species), 0.0);
    crnn_model mm = *m; mm.w_in = zw.data(); mm.w_b = zw.data(); mm.w_out = zw.data();
    int rc = validate(h, &mm, o, n_steps);
    if (rc) return rc;
  }
#define X(NS_, NR_, K_)                                                              \
  if (m->n_species == NS_ && m->n_reac == NR_ && m->rhs_kind == K_)                  \
    return train_impl<Cfg<NS_, NR_, K_>>(h, m, o, t, ds, order, n_steps, yscale, loss_kind, p, opt_state, step_loss, step_gnorm);
  CRNN_FOR_EACH_CFG(X)
#undef X
  return fail(h, CRNN_ERR_UNSUPPORTED, "the on-device training loop needs a model with a specialised kernel");
}

int crnn_loss_grad_particles(crnn_handle* h, const crnn_model* m, const crnn_opts* o, const double* weights,
                             const double* dW_dp, int32_t np, int32_t P, const double* u0, int32_t E,
                             const int32_t* n_save_used, const double* data, const double* tab_T, const double* tab_P,
                             const double* yscale, int32_t loss_kind, double* loss, double* grad, int32_t* n_saved,
                             int32_t* retcode, crnn_stats* stats) {
  if (!h) return CRNN_ERR_BAD_ARG;
  if (!h->kids.empty()) return fail(h, CRNN_ERR_UNSUPPORTED, "crnn_loss_grad_particles needs a single-device handle");
  if (!m || !o || !weights || !dW_dp || !u0 || !data || !loss || !grad || P < 1 || E < 1 || np < 1)
    return fail(h, CRNN_ERR_BAD_ARG, "crnn_loss_grad_particles: null argument or empty batch");
  if (o->buffers_on_device) return fail(h, CRNN_ERR_BAD_ARG, "crnn_loss_grad_particles takes host buffers");
  if (loss_kind != CRNN_LOSS_MAE_SCALED && loss_kind != CRNN_LOSS_MAE_LOG && loss_kind != CRNN_LOSS_MSE)
    return fail(h, CRNN_ERR_BAD_ARG, "bad loss_kind");
  if (loss_kind != CRNN_LOSS_MAE_LOG && !yscale) return fail(h, CRNN_ERR_BAD_ARG, "null yscale");
  if (o->sens_mode != CRNN_SENS_FORWARD) return fail(h, CRNN_ERR_UNSUPPORTED, "particles: forward mode only");
  const int n = m->n_state, ns = m->n_species, nin = m->n_in, nr = m->n_reac;
  const bool has_obs = m->w_obs != nullptr;
  const int off_b = nin * nr, off_out = off_b + nr, off_obs = off_out + ns * nr, nw = off_obs + (has_obs ? nr : 0);
  // the model seen by validation / build_wide: particle 0's weights
  crnn_model m0 = *m;
  m0.w_in = weights; m0.w_b = weights + off_b; m0.w_out = weights + off_out;
  if (has_obs) m0.w_obs = weights + off_obs;
  int rc = validate(h, &m0, o, (int64_t)P * E);
  if (rc) return rc;
  if (o->n_obs == 0 || o->n_save == 0) return fail(h, CRNN_ERR_BAD_ARG, "loss needs n_obs > 0 and n_save > 0");
  const bool trb = (o->alg == CRNN_ALG_TRBDF2 || o->alg == CRNN_ALG_AUTO_TSIT5_TRBDF2);   // src_333/network.jl: AutoTsit5(TRBDF2)
  if (o->alg != CRNN_ALG_TSIT5 && o->alg != CRNN_ALG_ROSENBROCK23 && o->alg != CRNN_ALG_AUTO_TSIT5_ROS23 && !trb)
    return fail(h, CRNN_ERR_UNSUPPORTED, "forward sensitivities are implemented for Tsit5, Rosenbrock23, TRBDF2 and the AutoTsit5 composites");
  const bool dens = (m->rhs_kind == CRNN_RHS_F2_MASSFRAC_TP);
  if (trb && dens) return fail(h, CRNN_ERR_UNSUPPORTED, "sensitivities through TRBDF2 are built for F0 / F1 / F5");
  const bool f2k = dens || m->rhs_kind == CRNN_RHS_F5_TRAMP;
  if ((tab_T || tab_P) && !f2k) return fail(h, CRNN_ERR_BAD_ARG, "per-experiment tables need rhs_kind F2 or F5");
  CK(cudaSetDevice(h->device));
  int cols = 0; size_t smem = 0; long long max_blocks = 0;
  void (*kern)(GenP, const double*, const int*, long long, const double*, double*, double*, double*, int*, int*, crnn_stats*,
               unsigned long long*, const long long*, const long long*, const unsigned int*) = nullptr;
  rc = gen_launch_cfg(h, &m0, np, cols, smem, max_blocks, f2k, trb, &kern);
  if (rc) return rc;
  cudaStream_t st = h->s_compute;
  // ---- per-particle blocks: weights in the kernel's layout, structured seed rows, descriptors ----
  const size_t pw_stride = (size_t)nin * KW_MAXN + nr + (size_t)ns * nr + (has_obs ? nr : 0);
  const int nrow = (has_obs ? 3 : 2) * nr;
  const size_t seed_stride = (size_t)nrow * cols;
  std::vector<double> pw(pw_stride * P, 0.0), rows(seed_stride * P, 0.0);
  std::vector<R1Desc> desc((size_t)cols * P, R1Desc{});
  for (int p = 0; p < P; ++p) {
    const double* w = weights + (size_t)nw * p;
    double* d = pw.data() + pw_stride * p;
    for (int j = 0; j < nr; ++j) {
      for (int i = 0; i < nin; ++i) d[(size_t)i * KW_MAXN + j] = w[i + nin * j];
      d[(size_t)nin * KW_MAXN + j] = w[off_b + j];
      for (int i = 0; i < ns; ++i)
        d[(size_t)nin * KW_MAXN + nr + i + ns * j] = w[off_out + i + ns * j] * (dens ? m->mw[i] : 1.0) * (m->out_scale ? m->out_scale[i] : 1.0);
      if (has_obs) d[(size_t)nin * KW_MAXN + nr + (size_t)ns * nr + j] = w[off_obs + j];
    }
    rc = gen_plan_seed(h, &m0, dW_dp + (size_t)nw * np * p, np, cols, has_obs, desc.data() + (size_t)cols * p, rows.data() + seed_stride * p);
    if (rc) return rc;
  }
  // per-experiment tables
  const int ntab = f2k ? m->n_tab : 0;
  const bool per_exp = f2k && tab_T != nullptr;
  std::vector<double> tabs;
  if (per_exp) {
    tabs.assign((size_t)2 * ntab * E, 1.0);
    for (size_t q = 0; q < (size_t)ntab * E; ++q) { tabs[q] = tab_T[q]; tabs[(size_t)ntab * E + q] = (dens && tab_P) ? tab_P[q] : 1.0; }
  }
  // device staging (reuses the handle's buffers): [pw | rows | desc | tabs | u0 | data | nsu] in adj_scratch
  const int64_t N = (int64_t)P * E;
  const size_t ps = (size_t)o->n_obs * o->n_save;
  const size_t nd_pw = pw.size(), nd_rows = rows.size(), nd_desc = 3 * desc.size(), nd_tabs = tabs.size(),
               nd_u0 = (size_t)n * E, nd_data = ps * E, nd_nsu = n_save_used ? ((size_t)N + 1) / 2 : 0;
  CK(h->adj_scratch.reserve((nd_pw + nd_rows + nd_desc + nd_tabs + nd_u0 + nd_data + nd_nsu + 8) * sizeof(double)));
  double* base = h->adj_scratch.as<double>();
  double* d_pw = base; double* d_rows = d_pw + nd_pw; double* d_desc = d_rows + nd_rows; double* d_tabs = d_desc + nd_desc;
  double* d_u0 = d_tabs + nd_tabs; double* d_data = d_u0 + nd_u0; int* d_nsu = reinterpret_cast<int*>(d_data + nd_data);
  CK(cudaMemcpyAsync(d_pw, pw.data(), nd_pw * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_rows, rows.data(), nd_rows * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_desc, desc.data(), desc.size() * sizeof(R1Desc), cudaMemcpyHostToDevice, st));
  if (per_exp) CK(cudaMemcpyAsync(d_tabs, tabs.data(), nd_tabs * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_u0, u0, nd_u0 * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_data, data, nd_data * sizeof(double), cudaMemcpyHostToDevice, st));
  std::vector<int> nsu_full;
  if (n_save_used) {   // per experiment -> per trajectory
    nsu_full.resize(N);
    for (int64_t q = 0; q < N; ++q) nsu_full[q] = n_save_used[q % E];
    CK(cudaMemcpyAsync(d_nsu, nsu_full.data(), N * sizeof(int), cudaMemcpyHostToDevice, st));
  }
  std::vector<double> extra(n, 1.0);   // inv_ys
  if (has_obs) { if (loss_kind != CRNN_LOSS_MAE_LOG) extra[0] = 1.0 / yscale[0]; }
  else
    for (int q = 0; q < o->n_obs; ++q) {
      const int r = o->obs_idx[q];
      if (r >= 0 && r < n && loss_kind != CRNN_LOSS_MAE_LOG) extra[r] = 1.0 / yscale[q];
    }
  GenP G{};
  const double* extra_dev = nullptr;
  const int order = (o->alg == CRNN_ALG_ROSENBROCK23 || o->alg == CRNN_ALG_TRBDF2) ? 2 : 5;
  rc = build_wide(h, &m0, o, order, extra, st, G.w, &extra_dev);
  if (rc) return rc;
  if (per_exp) { G.w.tab_T = d_tabs; G.w.tab_P = d_tabs + (size_t)ntab * E; }
  G.inv_ys = extra_dev; G.seed_rows = d_rows; G.desc = reinterpret_cast<const R1Desc*>(d_desc);
  G.w_obs = nullptr; G.pw = d_pw; G.pw_stride = (long long)pw_stride; G.seed_stride = (long long)seed_stride;
  G.n_part = P; G.n_exp = E; G.tab_per_exp = per_exp ? 1 : 0;
  G.np = np; G.cols = cols; G.loss_kind = loss_kind; G.incl_sens = o->err_norm_includes_sens ? 1 : 0;
  G.norm_cnt = (double)n * ((o->err_norm_includes_sens && !o->err_norm_mean_over_state_only) ? (double)(np + 1) : 1.0);
  CK(h->d_grad_each.reserve((size_t)N * np * sizeof(double)));
  CK(h->d_loss.reserve(N * sizeof(double)));
  CK(h->d_nsaved.reserve(N * sizeof(int)));
  CK(h->d_ret.reserve(N * sizeof(int)));
  if (stats) CK(h->d_stats.reserve(N * sizeof(crnn_stats)));
  CK(h->d_grad_sum.reserve((size_t)np * P * sizeof(double)));
  h->last_grad_np = -1; h->last_grad_n = -1;
  unsigned long long* queue = h->ctr.as<unsigned long long>();
  CK(cudaMemsetAsync(queue, 0, sizeof(unsigned long long), st));
  {
    ProfScope prof(h, st);
    kern<<<(unsigned)std::min<long long>(max_blocks, N), cols, smem, st>>>(
        G, d_u0, n_save_used ? d_nsu : nullptr, N, d_data, h->d_loss.as<double>(), h->d_grad_each.as<double>(), nullptr,
        h->d_nsaved.as<int>(), h->d_ret.as<int>(), stats ? h->d_stats.as<crnn_stats>() : nullptr, queue, nullptr, nullptr, nullptr);
    CK(cudaGetLastError());
    h->launches++;
  }
  k_particle_reduce<<<(np * P + 255) / 256, 256, 0, st>>>(h->d_grad_each.as<double>(), np, E, P, h->d_grad_sum.as<double>());
  CK(cudaGetLastError());
  h->launches++;
  CK(cudaMemcpyAsync(grad, h->d_grad_sum.p, (size_t)np * P * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(loss, h->d_loss.p, N * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (n_saved) CK(cudaMemcpyAsync(n_saved, h->d_nsaved.p, N * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (retcode) CK(cudaMemcpyAsync(retcode, h->d_ret.p, N * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (stats) CK(cudaMemcpyAsync(stats, h->d_stats.p, N * sizeof(crnn_stats), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return CRNN_OK;
}

int crnn_solve_batch(crnn_handle* h, const crnn_model* m, const crnn_opts* o, const double* u0, int64_t N,
                     const int32_t* n_save_used, double* pred, int32_t* n_saved, int32_t* retcode,
                     crnn_stats* stats) {
  if (!h) return CRNN_ERR_BAD_ARG;
  if (!h->kids.empty()) return multi_solve_batch(h, m, o, u0, N, n_save_used, pred, n_saved, retcode, stats);
  int rc = validate(h, m, o, N);
  if (rc) return rc;
  if (N > 0 && !u0) return fail(h, CRNN_ERR_BAD_ARG, "null u0");
  if (o->alg != CRNN_ALG_TSIT5 && o->alg != CRNN_ALG_ROSENBROCK23 && o->alg != CRNN_ALG_KENCARP4 &&
      o->alg != CRNN_ALG_AUTO_TSIT5_ROS23 && o->alg != CRNN_ALG_TRBDF2 && o->alg != CRNN_ALG_AUTO_TSIT5_TRBDF2)
    return fail(h, CRNN_ERR_UNSUPPORTED, "alg not supported by solve_batch");
  CK(cudaSetDevice(h->device));
  HostIO io{u0, n_save_used, nullptr, pred, nullptr, n_saved, retcode, stats};
  if (o->alg == CRNN_ALG_KENCARP4) return solve_kencarp4(h, m, o, io, N);
  if (o->alg == CRNN_ALG_TRBDF2 || o->alg == CRNN_ALG_AUTO_TSIT5_TRBDF2 || m->rhs_kind == CRNN_RHS_F4_MLP_AUG) return solve_wide(h, m, o, io, N);
  // dimension-specialised thread-per-trajectory kernels for the instantiated configurations ...
  const char* force = std::getenv("CRNN_B200_FORCE_WIDE");
  if (m->rhs_kind != CRNN_RHS_F2_MASSFRAC_TP && !(force && force[0] == '1')) {
#define X(NS_, NR_, K_)                                                              \
  if (m->n_species == NS_ && m->n_reac == NR_ && m->rhs_kind == K_)                  \
    return solve_impl<Cfg<NS_, NR_, K_>>(h, m, o, io, N);
    CRNN_FOR_EACH_CFG(X)
#undef X
  }
  // ... and the generic warp-per-trajectory kernel for everything else (any dimensions <= 32, F2)
  return solve_wide(h, m, o, io, N);
}

int crnn_loss_grad_batch(crnn_handle* h, const crnn_model* m, const crnn_opts* o, const double* dW_dp, int32_t np,
                         const double* u0, int64_t N, const int32_t* n_save_used, const double* data,
                         const double* yscale, int32_t loss_kind, double* loss, double* grad_sum, double* pred,
                         int32_t* n_saved, int32_t* retcode, crnn_stats* stats) {
  if (!h) return CRNN_ERR_BAD_ARG;
  if (!h->kids.empty()) return multi_loss_grad_batch(h, m, o, dW_dp, np, u0, N, n_save_used, data, yscale, loss_kind, loss,
                                                     grad_sum, pred, n_saved, retcode, stats);
  HostIO io{u0, n_save_used, data, pred, loss, n_saved, retcode, stats};
  return loss_grad_core(h, m, o, dW_dp, np, yscale, loss_kind, io, N, grad_sum);
}

}  // extern "C"
