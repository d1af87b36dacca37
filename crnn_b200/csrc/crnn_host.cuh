#pragma once
// crnn_host.cuh — host side of the engine shared by crnn_api.cu and the per-configuration
// instantiation units (inst.cu): argument validation, packing of the model/solver
// options into by-value kernel parameters, dispatch to the dimension-specialised
// kernels, and (for host buffers) a chunked H2D -> kernel -> D2H pipeline on two
// streams so that PCIe copies overlap the solve.  No CPU compute path exists here.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "../../include/crnn_b200.h"
#include "crnn_dev.cuh"
#include "kernel_tsit5_value.cuh"
#include "kernel_tsit5_sens.cuh"
#include "kernel_tsit5_sens_pl.cuh"
#include "kernel_rosenbrock23.cuh"
#include "kernel_rosenbrock23_sens.cuh"
#include "kernel_auto_value.cuh"
#include "kernel_train.cuh"

using namespace crnn;

// ------------------------------------------------------------------------------------------
// Dimension dispatch.  (n_species, n_reac, rhs_kind) of every model the reference scripts and
// their generating mechanisms use:
//   case1 5/4/F0, case2 6/3/F1, case3 9/8/F0, robertson CRNN 3/6/F0, robertson truth 3/3/F0,
//   gene-regulatory 9/15/F0.
// ------------------------------------------------------------------------------------------
#define CRNN_FOR_EACH_CFG(X) \
  X(5, 4, 0) X(6, 3, 1) X(9, 8, 0) X(3, 6, 0) X(3, 3, 0) X(9, 15, 0)


namespace crnn_host {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

constexpr int kPipe = 2;  // pipeline depth of the host-buffer path

}  // namespace crnn_host
using crnn_host::DevBuf;
using crnn_host::kPipe;

struct crnn_handle {
  int device = 0;
  int num_sms = 0;
  std::string err;
  int64_t launches = 0;
  int64_t last_grad_n = -1;  // shape of the per-trajectory gradients left in d_grad_each (crnn_copy_grad_each)
  int last_grad_np = -1;
  cudaStream_t s_compute = nullptr, s_h2d = nullptr, s_d2h = nullptr;
  cudaStream_t s_slot[kPipe] = {};  // one compute stream per pipeline slot: chunk c+1 fills the SMs chunk c's tail leaves idle
  cudaEvent_t ev_in[kPipe] = {}, ev_done[kPipe] = {}, ev_out[kPipe] = {}, ev_cfg = nullptr;
  // small per-call device state.  ctr: [0] work queue (device mode), [1] reduce ticket, [2+s] queue of slot s
  DevBuf cfg, seed, desc, ctr, partial;
  // staging for the host-buffer path (per pipeline slot) and full-batch gradients
  DevBuf d_u0[kPipe], d_nsu[kPipe], d_data[kPipe], d_pred[kPipe];
  // small per-trajectory outputs: full batch on the device, copied back once at the end (a D2H into
  // pageable host memory blocks the host thread, which would serialise the chunk pipeline)
  DevBuf d_loss, d_nsaved, d_ret, d_stats;
  DevBuf d_grad_each, d_grad_sum, d_grad_out, adj_scratch;
  // AutoTsit5 fast path: config blob of the generic composite kernel, hand-over lists (one per work-queue slot)
  DevBuf cfg2, auto_sel;
  // on-device training loop: p | optimiser state | ModelP | seed rows | descriptors | order | per-step loss / gnorm
  DevBuf train;
  // device-resident dataset path (crnn_loss_grad_indexed): row indices, n_save_used, [sum loss, n finite, grad(np)]
  DevBuf d_idx, d_nsu_ix, d_result;
  // multi-device parent (crnn_create_multi): the children own all per-device state, the parent launches nothing itself
  std::vector<crnn_handle*> kids;
  std::vector<void*> comms;  // ncclComm_t per child (NCCL is dlopen'ed: crnn_dataset.cu)
  // optional kernel timing (crnn_profile_begin/_end)
  bool profiling = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
  size_t prof_used = 0;
};

struct crnn_dataset {
  crnn_handle* owner = nullptr;
  int n_state = 0, n_obs = 0, n_save = 0;
  int64_t N = 0;
  std::vector<int64_t> lo;       // shard bounds, size n_dev + 1
  std::vector<DevBuf> u0, data;  // per device
};

#define CK(call)                                                                               \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e__);                            \
      return CRNN_ERR_CUDA;                                                                    \
    }                                                                                          \
  } while (0)

namespace crnn_host {

struct HostIO;
// crnn_api.cu: validation + dispatch of one loss/gradient call on ONE device
int loss_grad_core(crnn_handle* h, const crnn_model* m, const crnn_opts* o, const double* dW_dp, int32_t np,
                   const double* yscale, int32_t loss_kind, const HostIO& io, int64_t N, double* grad_sum);
// crnn_dataset.cu: the host-buffer entry points on a multi-device handle (contiguous shards, one host thread per device)
int multi_loss_grad_batch(crnn_handle* h, const crnn_model* m, const crnn_opts* o, const double* dW_dp, int32_t np,
                          const double* u0, int64_t N, const int32_t* n_save_used, const double* data, const double* yscale,
                          int32_t loss_kind, double* loss, double* grad_sum, double* pred, int32_t* n_saved,
                          int32_t* retcode, crnn_stats* stats);
void multi_release_comms(crnn_handle* h);
int multi_solve_batch(crnn_handle* h, const crnn_model* m, const crnn_opts* o, const double* u0, int64_t N,
                      const int32_t* n_save_used, double* pred, int32_t* n_saved, int32_t* retcode, crnn_stats* stats);

inline int fail(crnn_handle* h, int code, const std::string& msg) {
  h->err = msg;
  return code;
}

// brackets one solver-kernel launch with events when profiling is on
struct ProfScope {
  crnn_handle* h; cudaStream_t st; cudaEvent_t e1 = nullptr;
  ProfScope(crnn_handle* h_, cudaStream_t st_) : h(h_), st(st_) {
    if (!h->profiling) return;
    if (h->prof_used == h->prof_events.size()) {
      cudaEvent_t a, b;
      if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
      h->prof_events.emplace_back(a, b);
    }
    auto& pr = h->prof_events[h->prof_used++];
    cudaEventRecord(pr.first, st);
    e1 = pr.second;
  }
  ~ProfScope() { if (e1) cudaEventRecord(e1, st); }
};

struct Packed {
  std::vector<double> w_out_scaled, seed_pad;
  std::vector<int> row2obs;
  std::vector<double> inv_ys;
};

inline int validate(crnn_handle* h, const crnn_model* m, const crnn_opts* o, int64_t N) {
  if (!m || !o) return fail(h, CRNN_ERR_BAD_ARG, "null model/opts");
  if (N < 0) return fail(h, CRNN_ERR_BAD_ARG, "negative N");
  if (m->rhs_kind != CRNN_RHS_F0 && m->rhs_kind != CRNN_RHS_F1_ARRH_TSTATE && m->rhs_kind != CRNN_RHS_F2_MASSFRAC_TP &&
      m->rhs_kind != CRNN_RHS_F5_TRAMP && m->rhs_kind != CRNN_RHS_F4_MLP_AUG)
    return fail(h, CRNN_ERR_UNSUPPORTED, "rhs_kind not supported");
  const bool f2 = (m->rhs_kind == CRNN_RHS_F2_MASSFRAC_TP || m->rhs_kind == CRNN_RHS_F5_TRAMP);  // inputs from T(t) tables
  const bool f4 = (m->rhs_kind == CRNN_RHS_F4_MLP_AUG);
  if ((!f4 && m->n_in != m->n_state + (f2 ? 2 : 0)) || m->n_state != m->n_species + (m->rhs_kind == CRNN_RHS_F1_ARRH_TSTATE ? 1 : 0))
    return fail(h, CRNN_ERR_BAD_ARG, "inconsistent n_state / n_species / n_in for rhs_kind");
  if (f4) {
    if (m->mlp_n_layers < 1 || m->mlp_n_layers > 8 || !m->mlp_dims || !m->mlp_in_idx || !m->mlp_params || !m->aug_src || m->w_obs)
      return fail(h, CRNN_ERR_BAD_ARG, "F4 needs mlp_dims / mlp_in_idx / mlp_params / aug_src (1..8 layers) and no observable post-map");
    if (m->n_in < 1 || m->n_in > 32) return fail(h, CRNN_ERR_BAD_ARG, "F4: n_in must be 1..32");
    for (int l = 0; l <= m->mlp_n_layers; ++l)
      if (m->mlp_dims[l] < 1 || m->mlp_dims[l] > 32) return fail(h, CRNN_ERR_BAD_ARG, "F4: MLP layer widths must be 1..32");
    for (int i = 0; i < m->mlp_dims[0]; ++i)
      if (m->mlp_in_idx[i] < 0 || m->mlp_in_idx[i] >= m->n_state) return fail(h, CRNN_ERR_BAD_ARG, "F4: mlp_in_idx out of range");
    for (int i = 0; i < m->n_in; ++i)
      if (m->aug_src[i] >= m->n_state || m->aug_src[i] < -m->mlp_dims[m->mlp_n_layers]) return fail(h, CRNN_ERR_BAD_ARG, "F4: aug_src out of range");
  }
  if (f2) {
    if (!m->tab_t || !m->tab_T || m->n_tab < 2 || (m->rhs_kind == CRNN_RHS_F2_MASSFRAC_TP && (!m->mw || !m->tab_P)))
      return fail(h, CRNN_ERR_BAD_ARG, "F2 needs mw and the tab_t / tab_T / tab_P tables, F5 tab_t / tab_T (n_tab >= 2)");
    for (int k = 1; k < m->n_tab; ++k)
      if (!(m->tab_t[k] > m->tab_t[k - 1])) return fail(h, CRNN_ERR_BAD_ARG, "tab_t must be strictly ascending");
    if (m->tab_t[0] > o->t0 || m->tab_t[m->n_tab - 1] < o->t1)
      return fail(h, CRNN_ERR_BAD_ARG, "tab_t must cover [t0, t1]");
  }
  if (!m->w_in || !m->w_b || !m->w_out) return fail(h, CRNN_ERR_BAD_ARG, "null weights");
  if (!(m->lb > 0.0) || !(m->ub > m->lb)) return fail(h, CRNN_ERR_BAD_ARG, "need 0 < lb < ub");
  if (o->n_save < 0 || (o->n_save > 0 && !o->saveat)) return fail(h, CRNN_ERR_BAD_ARG, "bad saveat");
  for (int k = 0; k < o->n_save; ++k) {
    if (k > 0 && o->saveat[k] < o->saveat[k - 1]) return fail(h, CRNN_ERR_BAD_ARG, "saveat must be ascending");
    if (o->saveat[k] < o->t0 || o->saveat[k] > o->t1) return fail(h, CRNN_ERR_BAD_ARG, "saveat outside [t0,t1]");
  }
  if (!(o->t1 > o->t0)) return fail(h, CRNN_ERR_BAD_ARG, "need t1 > t0");
  if ((o->n_abstol != 1 && o->n_abstol != m->n_state) || (o->n_reltol != 1 && o->n_reltol != m->n_state) ||
      !o->abstol || !o->reltol)
    return fail(h, CRNN_ERR_BAD_ARG, "abstol/reltol must have 1 or n_state entries");
  if (o->n_obs < 0 || o->n_obs > m->n_state || (o->n_obs > 0 && !o->obs_idx))
    return fail(h, CRNN_ERR_BAD_ARG, "bad obs_idx");
  if (m->w_obs && o->n_obs != 1) return fail(h, CRNN_ERR_BAD_ARG, "the observable post-map w_obs defines ONE observed quantity: n_obs must be 1");
  if (o->maxiters <= 0) return fail(h, CRNN_ERR_BAD_ARG, "maxiters must be positive");
  return CRNN_OK;
}

template <class C>
int pack(crnn_handle* h, const crnn_model* m, const crnn_opts* o, const double* yscale, int loss_kind,
         ModelP<C>& mp, SolveP<C>& sp, Packed& pk) {
  for (int q = 0; q < C::NIN * C::NR; ++q) mp.w_in[q] = m->w_in[q];
  for (int q = 0; q < C::NR; ++q) mp.w_b[q] = m->w_b[q];
  for (int j = 0; j < C::NR; ++j)
    for (int i = 0; i < C::NS; ++i)
      mp.w_out[i + C::NS * j] = m->w_out[i + C::NS * j] * (m->out_scale ? m->out_scale[i] : 1.0);
  mp.lb = m->lb; mp.ub = m->ub; mp.gas_R = m->gas_R;
  const int order = (o->alg == CRNN_ALG_TSIT5 || o->alg == CRNN_ALG_AUTO_TSIT5_ROS23) ? 5 : (o->alg == CRNN_ALG_ROSENBROCK23 ? 2 : 4);
  for (int i = 0; i < C::N; ++i) {
    sp.abstol[i] = o->abstol[o->n_abstol > 1 ? i : 0];
    sp.reltol[i] = o->reltol[o->n_reltol > 1 ? i : 0];
    sp.inv_yscale[i] = 1.0;
  }
  pk.row2obs.assign(C::N, -1);
  for (int q = 0; q < o->n_obs; ++q) {
    int r = o->obs_idx[q];
    if (r < 0 || r >= C::N) return fail(h, CRNN_ERR_BAD_ARG, "obs_idx out of range");
    if (pk.row2obs[r] >= 0) return fail(h, CRNN_ERR_BAD_ARG, "obs_idx has duplicates");
    pk.row2obs[r] = q;
    if (yscale && loss_kind == CRNN_LOSS_MAE_SCALED) sp.inv_yscale[r] = 1.0 / yscale[q];
  }
  sp.t0 = o->t0; sp.t1 = o->t1;
  sp.pred_lo = o->pred_clamp_lo; sp.pred_hi = o->pred_clamp_hi;
  const double qmin = o->qmin > 0 ? o->qmin : 0.2, qmax = o->qmax > 0 ? o->qmax : 10.0;
  sp.inv_qmin = 1.0 / qmin; sp.inv_qmax = 1.0 / qmax;
  sp.gamma = o->gamma > 0 ? o->gamma : 0.9;
  sp.beta2 = o->beta2 > 0 ? o->beta2 : 2.0 / (5.0 * order);
  sp.beta1 = o->beta1 > 0 ? o->beta1 : 7.0 / (10.0 * order);
  sp.inv_order = 1.0 / order;
  sp.beta2_ros = o->beta2 > 0 ? o->beta2 : 2.0 / (5.0 * 2.0);
  sp.beta1_ros = o->beta1 > 0 ? o->beta1 : 7.0 / (10.0 * 2.0);
  sp.maxiters = o->maxiters;
  sp.n_save = o->n_save; sp.n_obs = o->n_obs;
  sp.incl_sens = o->err_norm_includes_sens;
  // qsteady defaults: 1 / 1 for explicit and composite algorithms, 1 / (6//5) for the adaptive implicit ones
  const bool implicit_alg = (o->alg == CRNN_ALG_ROSENBROCK23 || o->alg == CRNN_ALG_KENCARP4);
  sp.qs_min = o->qsteady_min > 0 ? o->qsteady_min : 1.0;
  sp.qs_max = o->qsteady_max > 0 ? o->qsteady_max : (implicit_alg ? 1.2 : 1.0);
  sp.norm_cnt = (double)C::N;  // loss_grad_impl multiplies by (1 + np) when the partials share the mean
  sp.eig_cnt = (double)C::N;
  sp.loss_kind = loss_kind;
  return CRNN_OK;
}

// uploads saveat + row2obs into the handle's config blob; fills the device pointers of sp
template <class C>
int upload_cfg(crnn_handle* h, const crnn_opts* o, const Packed& pk, SolveP<C>& sp, cudaStream_t st) {
  const size_t off_r2o = ((size_t)o->n_save * sizeof(double) + 15) & ~size_t(15);
  const size_t bytes = off_r2o + C::N * sizeof(int);
  CK(h->cfg.reserve(std::max<size_t>(bytes, 4096)));
  if (o->n_save) CK(cudaMemcpyAsync(h->cfg.p, o->saveat, o->n_save * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync((char*)h->cfg.p + off_r2o, pk.row2obs.data(), C::N * sizeof(int), cudaMemcpyHostToDevice, st));
  sp.saveat = h->cfg.as<double>();
  sp.row2obs = reinterpret_cast<const int*>((char*)h->cfg.p + off_r2o);
  return CRNN_OK;
}

struct BatchPtrs;
// AutoTsit5(Rosenbrock23) on a specialised Tsit5 kernel: the kernel monitors stiffness and hands the trajectories that
// would switch over to `fallback` (the generic composite kernel), enqueued right behind it on the same stream.
struct AutoHook {
  long long* sel = nullptr;        // device [(2 + kPipe) * stride] hand-over lists, one region per work-queue slot
  unsigned int* count = nullptr;   // device [2 + kPipe]
  size_t stride = 0;
  std::function<int(const BatchPtrs&, cudaStream_t, const long long* sel, const unsigned int* count)> fallback;
};

struct BatchPtrs {  // device pointers of one (sub)batch
  const double* u0; const int* nsu; const double* data;
  double* pred; double* loss; int* n_saved; int* retcode; crnn_stats* stats;
  double* grad_each;
  long long n;
  int qslot;  // which work-queue counter of the handle this launch uses
  const long long* in_idx = nullptr;  // device [n] dataset rows of the inputs (crnn_loss_grad_indexed), or NULL
};

// ---------------- value path launchers ----------------
template <class C>
int launch_value(crnn_handle* h, int alg, const ModelP<C>& mp, const SolveP<C>& sp, const BatchPtrs& b,
                 cudaStream_t st) {
  if (b.n == 0) return CRNN_OK;
  const int threads = 128;
  const unsigned blocks = (unsigned)((b.n + threads - 1) / threads);
  ProfScope prof(h, st);
  if (alg == CRNN_ALG_TSIT5)
    k_tsit5_value<C><<<blocks, threads, 0, st>>>(mp, sp, b.u0, b.nsu, b.n, b.pred, b.n_saved, b.retcode, b.stats);
  else if (alg == CRNN_ALG_AUTO_TSIT5_ROS23)
    k_auto_value<C><<<blocks, threads, 0, st>>>(mp, sp, b.u0, b.nsu, b.n, b.pred, b.n_saved, b.retcode, b.stats);
  else
    k_rosenbrock23_value<C><<<blocks, threads, 0, st>>>(mp, sp, b.u0, b.nsu, b.n, b.pred, b.n_saved, b.retcode,
                                                       b.stats);
  CK(cudaGetLastError());
  h->launches++;
  return CRNN_OK;
}

// ---------------- sensitivity path launchers ----------------
template <class C, int CT, bool R1, int WPT = 1, int NGRP = 0>
int launch_sens(crnn_handle* h, const ModelP<C>& mp, const SolveP<C>& sp, int ncol, const BatchPtrs& b,
                cudaStream_t st, const AutoHook* hook = nullptr) {
  if (b.n == 0) return CRNN_OK;
  if constexpr (WPT == 1) {
    if (hook) {   // AutoTsit5(Rosenbrock23): Tsit5 + AutoSwitch monitor here, the composite kernel for what it hands over
      constexpr int WARPS_A = CT == 1 ? 8 : 4;
      auto kern = k_tsit5_sens<C, CT, WARPS_A, 2, R1, 1, true>;
      const size_t smem = sizeof(SensSmem<C, CT, R1, 1>) + WARPS_A * sizeof(WarpBuf<C, CT>);
      if (smem > 227 * 1024) return fail(h, CRNN_ERR_UNSUPPORTED, "model too large for the forward-sensitivity kernel's shared memory");
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int bps = 0;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, WARPS_A * 32, smem));
      if (bps < 1) bps = 1;
      const unsigned blocks = (unsigned)std::min<long long>((long long)h->num_sms * bps, (b.n + WARPS_A - 1) / WARPS_A);
      unsigned long long* queue = h->ctr.as<unsigned long long>() + b.qslot;
      long long* sel = hook->sel + (size_t)b.qslot * hook->stride;
      unsigned int* cnt = hook->count + b.qslot;
      CK(cudaMemsetAsync(queue, 0, sizeof(unsigned long long), st));
      CK(cudaMemsetAsync(cnt, 0, sizeof(unsigned int), st));
      {
        ProfScope prof(h, st);
        kern<<<blocks, WARPS_A * 32, smem, st>>>(mp, sp, h->seed.as<double>(), h->desc.as<R1Desc>(), ncol, b.u0, b.nsu, b.n,
                                              b.data, b.loss, b.grad_each, b.pred, b.n_saved, b.retcode, b.stats, queue,
                                              b.in_idx, sel, cnt, nullptr);
        CK(cudaGetLastError());
        h->launches++;
      }
      return hook->fallback(b, st, sel, cnt);
    }
  }
  // warps per block: 16 warps/SM in two blocks for the single-warp layouts; one or two warp
  // groups per block when WPT warps share a trajectory (NGRP overrides the number of groups)
  constexpr int WARPS = WPT == 1 ? (CT == 1 ? 8 : 4) : (NGRP > 0 ? NGRP * WPT : (WPT <= 4 ? 2 * WPT : WPT));
  constexpr int MINB = WPT == 1 ? 2 : 1;
  if constexpr (WPT == 1) {
    // CRNN_B200_SENS_PIPELINED=1 selects the experimental kernel with the pipelined value path (kernel_tsit5_sens_pl.cuh:
    // same results, measured 10-13 % SLOWER - kept for A/B measurements)
    static const bool pipelined = [] { const char* e = std::getenv("CRNN_B200_SENS_PIPELINED"); return e && e[0] == '1'; }();
    if (pipelined) {
      auto kern = k_tsit5_sens_pl<C, CT, WARPS, MINB, R1>;
      const size_t smem = sizeof(SensSmemP<C, CT, R1>) + WARPS * sizeof(WarpBufP<C, CT>);
      if (smem > 227 * 1024) return fail(h, CRNN_ERR_UNSUPPORTED, "model too large for the forward-sensitivity kernel's shared memory");
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int bps = 0;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, WARPS * 32, smem));
      if (bps < 1) bps = 1;
      unsigned blocks = (unsigned)std::min<long long>((long long)h->num_sms * bps, (b.n + WARPS - 1) / WARPS);
      unsigned long long* queue = h->ctr.as<unsigned long long>() + b.qslot;
      CK(cudaMemsetAsync(queue, 0, sizeof(unsigned long long), st));
      ProfScope prof(h, st);
      kern<<<blocks, WARPS * 32, smem, st>>>(mp, sp, h->seed.as<double>(), h->desc.as<R1Desc>(), ncol, b.u0, b.nsu, b.n,
                                          b.data, b.loss, b.grad_each, b.pred, b.n_saved, b.retcode, b.stats, queue, b.in_idx);
      CK(cudaGetLastError());
      h->launches++;
      return CRNN_OK;
    }
  }
  auto kern = k_tsit5_sens<C, CT, WARPS, MINB, R1, WPT>;
  const size_t smem = sizeof(SensSmem<C, CT, R1, WPT>) + WARPS * sizeof(WarpBuf<C, CT>);
  if (smem > 227 * 1024) return fail(h, CRNN_ERR_UNSUPPORTED, "model too large for the forward-sensitivity kernel's shared memory");
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int bps = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, WARPS * 32, smem));
  if (bps < 1) bps = 1;
  constexpr int GROUPS = WARPS / WPT;
  long long want = (b.n + GROUPS - 1) / GROUPS;
  unsigned blocks = (unsigned)std::min<long long>((long long)h->num_sms * bps, want);
  unsigned long long* queue = h->ctr.as<unsigned long long>() + b.qslot;
  CK(cudaMemsetAsync(queue, 0, sizeof(unsigned long long), st));
  ProfScope prof(h, st);
  kern<<<blocks, WARPS * 32, smem, st>>>(mp, sp, h->seed.as<double>(), h->desc.as<R1Desc>(), ncol, b.u0, b.nsu, b.n,
                                      b.data, b.loss, b.grad_each, b.pred, b.n_saved, b.retcode, b.stats, queue, b.in_idx,
                                      nullptr, nullptr, nullptr);
  CK(cudaGetLastError());
  h->launches++;
  return CRNN_OK;
}

// Rosenbrock23 + forward sensitivities (structured seeds, NS <= 6: per-lane register LU)
template <class C, int CT>
int launch_ros_sens(crnn_handle* h, const ModelP<C>& mp, const SolveP<C>& sp, int ncol, const BatchPtrs& b,
                    cudaStream_t st) {
  if constexpr (C::NS > 6) {
    return fail(h, CRNN_ERR_UNSUPPORTED, "Rosenbrock23 forward sensitivities need n_species <= 6");
  } else {
    if (b.n == 0) return CRNN_OK;
    constexpr int WARPS = 4, MINB = 3;
    auto kern = k_rosenbrock23_sens<C, CT, WARPS, MINB>;
    const size_t smem = sizeof(SensSmem<C, CT, true>) + WARPS * sizeof(RosWarpBuf<C, CT>);
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int bps = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, WARPS * 32, smem));
    if (bps < 1) bps = 1;
    long long want = (b.n + WARPS - 1) / WARPS;
    unsigned blocks = (unsigned)std::min<long long>((long long)h->num_sms * bps, want);
    unsigned long long* queue = h->ctr.as<unsigned long long>() + b.qslot;
    CK(cudaMemsetAsync(queue, 0, sizeof(unsigned long long), st));
    ProfScope prof(h, st);
    kern<<<blocks, WARPS * 32, smem, st>>>(mp, sp, h->seed.as<double>(), h->desc.as<R1Desc>(), ncol, b.u0, b.nsu, b.n,
                                        b.data, b.loss, b.grad_each, b.pred, b.n_saved, b.retcode, b.stats, queue, b.in_idx,
                                        nullptr);
    CK(cudaGetLastError());
    h->launches++;
    return CRNN_OK;
  }
}

inline int launch_grad_reduce(crnn_handle* h, const double* grad_each, long long n, int np, double* grad_sum_dev,
                              cudaStream_t st) {
  const int nb = (int)std::max<long long>(1, std::min<long long>(4LL * h->num_sms, (n + 63) / 64));
  CK(h->partial.reserve((size_t)nb * np * sizeof(double)));
  k_grad_reduce<<<nb, 256, 0, st>>>(grad_each, n, np, h->partial.as<double>(), grad_sum_dev,
                                    reinterpret_cast<unsigned int*>(h->ctr.as<unsigned long long>() + 1));
  CK(cudaGetLastError());
  h->launches++;
  return CRNN_OK;
}

// Structured ("R1") form of the seed matrix (kernel_tsit5_sens.cuh): usable when every column of
// dW/dp touches w_in in at most ONE row and w_out in at most ONE entry — true for every p2vec of
// the reference scripts.  Otherwise ok = false and the caller uses the dense layout.
struct R1Plan {
  std::vector<R1Desc> desc;   // [32*ct], lane 0 of tile 0 = value column
  std::vector<double> rows;   // [2*NR][32*ct]: a_j = dW_in[i_in, j], then b_j = db_j
  bool ok = false;
};

template <class C>
R1Plan plan_r1(const crnn_model* m, const double* dW_dp, int np) {
  R1Plan pl;
  const int tiles = (np + 1 + 31) / 32;
  const int width = 32 * (tiles <= 2 ? tiles : (tiles <= 3 ? 3 : (tiles <= 5 ? 5 : 8)));  // matches the WPT dispatch
  const int off_b = C::NIN * C::NR, off_out = off_b + C::NR;
  pl.desc.assign(width, R1Desc{});
  pl.rows.assign((size_t)2 * C::NR * width, 0.0);
  for (int c = 0; c < np; ++c) {
    const double* s = dW_dp + (size_t)C::NW * c;
    R1Desc d{};
    int i_in = -1, n_out = 0;
    for (int j = 0; j < C::NR; ++j) {
      for (int i = 0; i < C::NIN; ++i)
        if (s[i + C::NIN * j] != 0.0) {
          if (i_in >= 0 && i_in != i) return pl;
          i_in = i;
        }
      for (int i = 0; i < C::NS; ++i)
        if (s[off_out + i + C::NS * j] != 0.0) {
          if (++n_out > 1) return pl;
          d.i_out = i; d.j_out = j;
          d.o = s[off_out + i + C::NS * j] * (m->out_scale ? m->out_scale[i] : 1.0);
        }
    }
    d.i_in = i_in < 0 ? 0 : i_in;
    pl.desc[c + 1] = d;
    for (int j = 0; j < C::NR; ++j) {
      pl.rows[(size_t)j * width + c + 1] = i_in < 0 ? 0.0 : s[i_in + C::NIN * j];
      pl.rows[(size_t)(C::NR + j) * width + c + 1] = s[off_b + j];
    }
  }
  pl.ok = true;
  return pl;
}

// Pads dW/dp to [NW][32*CT] (column 0 = value lane = 0), folding out_scale into the w_out rows.
template <class C>
int upload_seed(crnn_handle* h, const crnn_model* m, const double* dW_dp, int np, int ct, Packed& pk,
                cudaStream_t st) {
  const int width = 32 * ct;
  pk.seed_pad.assign((size_t)C::NW * width, 0.0);
  const int off_out = C::NIN * C::NR + C::NR;
  for (int c = 0; c < np; ++c)
    for (int w = 0; w < C::NW; ++w) {
      double v = dW_dp[w + (size_t)C::NW * c];
      if (w >= off_out && m->out_scale) v *= m->out_scale[(w - off_out) % C::NS];
      pk.seed_pad[(size_t)w * width + (c + 1)] = v;
    }
  CK(h->seed.reserve(pk.seed_pad.size() * sizeof(double)));
  CK(cudaMemcpyAsync(h->seed.p, pk.seed_pad.data(), pk.seed_pad.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  return CRNN_OK;
}

// The generic driver: runs `launch(b, stream)` over the batch, either directly on device
// buffers or through the chunked host pipeline.
struct HostIO {
  const double* u0; const int32_t* nsu; const double* data;
  double* pred; double* loss; int32_t* n_saved; int32_t* retcode; crnn_stats* stats;
  const long long* in_idx = nullptr;  // device buffers only: dataset rows of the inputs
};

// `post` (optional): maps the reduced per-trajectory vector (length np, device) to the final gradient
// (length nout, device) — the adjoint path contracts vec(G) with dW/dp there.
using PostFn = std::function<int(const double* red_dev, double* out_dev, cudaStream_t)>;

template <class F>
int run_batch(crnn_handle* h, const crnn_model* m, const crnn_opts* o, const HostIO& io, int64_t N, bool want_loss,
              int np, double* grad_sum, F&& launch, int nout = -1, const PostFn& post = nullptr) {
  if (nout < 0) nout = np;
  const size_t ns = m->n_state, ps = (size_t)o->n_obs * o->n_save;
  if (o->buffers_on_device) {
    cudaStream_t st = (cudaStream_t)o->stream;
    BatchPtrs b{io.u0, io.nsu, io.data, io.pred, io.loss, io.n_saved, io.retcode, io.stats, nullptr, N, 0, io.in_idx};
    if (want_loss && np > 0) {
      CK(h->d_grad_each.reserve(std::max<size_t>(8, (size_t)N * np * sizeof(double))));
      b.grad_each = h->d_grad_each.as<double>();
    }
    int rc = launch(b, st);
    if (rc) return rc;
    if (want_loss && np > 0 && grad_sum) {
      if (N == 0) { CK(cudaMemsetAsync(grad_sum, 0, nout * sizeof(double), st)); }
      else if (!post) { rc = launch_grad_reduce(h, b.grad_each, N, np, grad_sum, st); if (rc) return rc; }
      else {
        CK(h->d_grad_sum.reserve(np * sizeof(double)));
        rc = launch_grad_reduce(h, b.grad_each, N, np, h->d_grad_sum.as<double>(), st); if (rc) return rc;
        rc = post(h->d_grad_sum.as<double>(), grad_sum, st); if (rc) return rc;
      }
    }
    return CRNN_OK;
  }

  // ---- host buffers: chunked, double-buffered pipeline (H2D stream || compute stream || D2H stream) ----
  if (want_loss && np > 0) CK(h->d_grad_each.reserve(std::max<size_t>(8, (size_t)N * np * sizeof(double))));
  CK(cudaEventRecord(h->ev_cfg, h->s_compute));  // model/option uploads were enqueued on s_compute
  for (int s = 0; s < kPipe; ++s) CK(cudaStreamWaitEvent(h->s_slot[s], h->ev_cfg, 0));
  int64_t nsplit = 4;  // full-size chunks per call; CRNN_B200_CHUNKS overrides (tuning knob of the host pipeline)
  if (const char* e = std::getenv("CRNN_B200_CHUNKS")) nsplit = std::max(1, std::atoi(e));
  const int64_t chunk = std::max<int64_t>(2048, (N + nsplit - 1) / nsplit);
  // chunk boundaries: the H2D of the FIRST chunk cannot overlap any compute, so the pipeline ramps up — 1/8, 1/4, 1/2 of
  // a chunk, then full chunks (CRNN_B200_RAMP = first-chunk divisor, 0 or 1 switches the ramp off)
  std::vector<int64_t> bounds{0};
  {
    int ramp = 8;
    if (const char* e = std::getenv("CRNN_B200_RAMP")) ramp = std::atoi(e);
    int64_t sz = ramp > 1 ? std::max<int64_t>(2048, chunk / ramp) : chunk;
    while (bounds.back() < N) {
      bounds.push_back(std::min<int64_t>(N, bounds.back() + std::min(sz, chunk)));
      sz *= 2;
    }
  }
  const int64_t nchunk = (int64_t)bounds.size() - 1;
  for (int s = 0; s < kPipe; ++s) {
    CK(h->d_u0[s].reserve(chunk * ns * sizeof(double)));
    if (io.nsu) CK(h->d_nsu[s].reserve(chunk * sizeof(int)));
    if (want_loss) CK(h->d_data[s].reserve(std::max<size_t>(8, chunk * ps * sizeof(double))));
    if (io.pred) CK(h->d_pred[s].reserve(std::max<size_t>(8, chunk * ps * sizeof(double))));
  }
  if (want_loss) CK(h->d_loss.reserve(std::max<size_t>(8, N * sizeof(double))));
  CK(h->d_nsaved.reserve(std::max<size_t>(8, N * sizeof(int))));
  CK(h->d_ret.reserve(std::max<size_t>(8, N * sizeof(int))));
  if (io.stats) CK(h->d_stats.reserve(std::max<size_t>(8, N * sizeof(crnn_stats))));
  for (int64_t c = 0; c < nchunk; ++c) {
    const int s = (int)(c % kPipe);
    const int64_t lo = bounds[c], n = bounds[c + 1] - lo;
    // slot reuse: inputs may be overwritten once the kernel of chunk c-kPipe is done,
    // outputs once their D2H copies are done
    if (c >= kPipe) CK(cudaStreamWaitEvent(h->s_h2d, h->ev_done[s], 0));
    CK(cudaMemcpyAsync(h->d_u0[s].p, io.u0 + lo * ns, n * ns * sizeof(double), cudaMemcpyHostToDevice, h->s_h2d));
    if (io.nsu) CK(cudaMemcpyAsync(h->d_nsu[s].p, io.nsu + lo, n * sizeof(int), cudaMemcpyHostToDevice, h->s_h2d));
    if (want_loss && ps) CK(cudaMemcpyAsync(h->d_data[s].p, io.data + lo * ps, n * ps * sizeof(double), cudaMemcpyHostToDevice, h->s_h2d));
    CK(cudaEventRecord(h->ev_in[s], h->s_h2d));
    cudaStream_t sc = h->s_slot[s];
    CK(cudaStreamWaitEvent(sc, h->ev_in[s], 0));
    if (c >= kPipe) CK(cudaStreamWaitEvent(sc, h->ev_out[s], 0));
    BatchPtrs b{h->d_u0[s].as<double>(), io.nsu ? h->d_nsu[s].as<int>() : nullptr,
                want_loss ? h->d_data[s].as<double>() : nullptr, io.pred ? h->d_pred[s].as<double>() : nullptr,
                want_loss ? h->d_loss.as<double>() + lo : nullptr, h->d_nsaved.as<int>() + lo, h->d_ret.as<int>() + lo,
                io.stats ? h->d_stats.as<crnn_stats>() + lo : nullptr,
                (want_loss && np > 0) ? h->d_grad_each.as<double>() + (size_t)lo * np : nullptr, n, 2 + s};
    int rc = launch(b, sc);
    if (rc) return rc;
    CK(cudaEventRecord(h->ev_done[s], sc));
    // D2H of this chunk's saved states on its own stream (a second copy engine; asynchronous when
    // the caller's buffer is pinned)
    CK(cudaStreamWaitEvent(h->s_d2h, h->ev_done[s], 0));
    if (io.pred && ps) CK(cudaMemcpyAsync(io.pred + lo * ps, h->d_pred[s].p, n * ps * sizeof(double), cudaMemcpyDeviceToHost, h->s_d2h));
    CK(cudaEventRecord(h->ev_out[s], h->s_d2h));
  }
  for (int s = 0; s < kPipe && s < nchunk; ++s) CK(cudaStreamWaitEvent(h->s_compute, h->ev_done[s], 0));
  if (want_loss && np > 0 && grad_sum) {
    if (N == 0) {
      std::memset(grad_sum, 0, nout * sizeof(double));
    } else {
      CK(h->d_grad_sum.reserve(np * sizeof(double)));
      int rc = launch_grad_reduce(h, h->d_grad_each.as<double>(), N, np, h->d_grad_sum.as<double>(), h->s_compute);
      if (rc) return rc;
      const double* src = h->d_grad_sum.as<double>();
      if (post) {
        CK(h->d_grad_out.reserve(std::max<size_t>(8, nout * sizeof(double))));
        rc = post(src, h->d_grad_out.as<double>(), h->s_compute);
        if (rc) return rc;
        src = h->d_grad_out.as<double>();
      }
      CK(cudaMemcpyAsync(grad_sum, src, nout * sizeof(double), cudaMemcpyDeviceToHost, h->s_compute));
    }
  }
  // small outputs: one copy each, after every chunk's kernel (s_compute has waited on them above)
  if (N > 0) {
    if (want_loss && io.loss) CK(cudaMemcpyAsync(io.loss, h->d_loss.p, N * sizeof(double), cudaMemcpyDeviceToHost, h->s_compute));
    if (io.n_saved) CK(cudaMemcpyAsync(io.n_saved, h->d_nsaved.p, N * sizeof(int), cudaMemcpyDeviceToHost, h->s_compute));
    if (io.retcode) CK(cudaMemcpyAsync(io.retcode, h->d_ret.p, N * sizeof(int), cudaMemcpyDeviceToHost, h->s_compute));
    if (io.stats) CK(cudaMemcpyAsync(io.stats, h->d_stats.p, N * sizeof(crnn_stats), cudaMemcpyDeviceToHost, h->s_compute));
  }
  CK(cudaStreamSynchronize(h->s_h2d));
  CK(cudaStreamSynchronize(h->s_d2h));
  for (int s = 0; s < kPipe; ++s) CK(cudaStreamSynchronize(h->s_slot[s]));
  CK(cudaStreamSynchronize(h->s_compute));
  return CRNN_OK;
}

}  // namespace


namespace crnn_host {

template <class C>
int solve_impl(crnn_handle* h, const crnn_model* m, const crnn_opts* o, const HostIO& io, int64_t N) {
  ModelP<C> mp; SolveP<C> sp; Packed pk;
  int rc = pack<C>(h, m, o, nullptr, 0, mp, sp, pk);
  if (rc) return rc;
  cudaStream_t st = o->buffers_on_device ? (cudaStream_t)o->stream : h->s_compute;
  rc = upload_cfg<C>(h, o, pk, sp, st);
  if (rc) return rc;
  const int alg = o->alg;
  return run_batch(h, m, o, io, N, false, 0, nullptr,
                   [&](const BatchPtrs& b, cudaStream_t s) { return launch_value<C>(h, alg, mp, sp, b, s); });
}

template <class C>
int loss_grad_impl(crnn_handle* h, const crnn_model* m, const crnn_opts* o, const double* dW_dp, int np,
                   const double* yscale, int loss_kind, const HostIO& io, int64_t N, double* grad_sum,
                   const AutoHook* hook) {
  if (o->alg != CRNN_ALG_TSIT5 && o->alg != CRNN_ALG_ROSENBROCK23 && !(hook && o->alg == CRNN_ALG_AUTO_TSIT5_ROS23))
    return fail(h, CRNN_ERR_UNSUPPORTED, "forward sensitivities are implemented for Tsit5 and Rosenbrock23");
  const bool ros = (o->alg == CRNN_ALG_ROSENBROCK23);
  const int ncol = np + 1;
  const int ct = (ncol + 31) / 32;  // 32-lane tiles of dual columns
  if (ct > 8) return fail(h, CRNN_ERR_UNSUPPORTED, "forward mode supports np <= 255");
  ModelP<C> mp; SolveP<C> sp; Packed pk;
  int rc = pack<C>(h, m, o, yscale, loss_kind, mp, sp, pk);
  if (rc) return rc;
  if (o->err_norm_includes_sens && !o->err_norm_mean_over_state_only) sp.norm_cnt = (double)C::N * ncol;  // totallength(u)
  if (o->err_norm_includes_sens) sp.eig_cnt = (double)C::N * ncol;
  if (hook && ct > 2) return fail(h, CRNN_ERR_UNSUPPORTED, "the AutoSwitch fast path serves np <= 63");
  cudaStream_t st = o->buffers_on_device ? (cudaStream_t)o->stream : h->s_compute;
  rc = upload_cfg<C>(h, o, pk, sp, st);
  if (rc) return rc;
  R1Plan pl = plan_r1<C>(m, dW_dp, np);
  const bool r1 = pl.ok;
  if (!r1 && (ct > 2 || ros))
    return fail(h, CRNN_ERR_UNSUPPORTED, "np > 63 and Rosenbrock23 sensitivities need a structured seed (one w_in row and one w_out entry per column)");
  if (ros && ct > 2) return fail(h, CRNN_ERR_UNSUPPORTED, "Rosenbrock23 forward sensitivities support np <= 63");
  if (r1) {
    CK(h->desc.reserve(pl.desc.size() * sizeof(R1Desc)));
    CK(cudaMemcpyAsync(h->desc.p, pl.desc.data(), pl.desc.size() * sizeof(R1Desc), cudaMemcpyHostToDevice, st));
    CK(h->seed.reserve(pl.rows.size() * sizeof(double)));
    CK(cudaMemcpyAsync(h->seed.p, pl.rows.data(), pl.rows.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  } else {
    rc = upload_seed<C>(h, m, dW_dp, np, ct, pk, st);
    if (rc) return rc;
    CK(h->desc.reserve(sizeof(R1Desc)));
  }
  return run_batch(h, m, o, io, N, true, np, grad_sum, [&](const BatchPtrs& b, cudaStream_t s) {
    if (ros) return ct == 1 ? launch_ros_sens<C, 1>(h, mp, sp, ncol, b, s) : launch_ros_sens<C, 2>(h, mp, sp, ncol, b, s);
    if (r1 && ct > 2) {  // several warps per trajectory
      if (ct <= 3) return launch_sens<C, 1, true, 3>(h, mp, sp, ncol, b, s);
      if (ct <= 5) {
        // two trajectories per block (and per SM) when their stage vectors fit the 227 KB: case3's np = 153 does
        // (224 KB), and a block of one group leaves the SM with a single trajectory in flight between barriers
        constexpr size_t smem2 = sizeof(SensSmem<C, 1, true, 5>) + 10 * sizeof(WarpBuf<C, 1>);
        if constexpr (smem2 <= 227 * 1024) return launch_sens<C, 1, true, 5, 2>(h, mp, sp, ncol, b, s);
        else return launch_sens<C, 1, true, 5>(h, mp, sp, ncol, b, s);
      }
      return launch_sens<C, 1, true, 8>(h, mp, sp, ncol, b, s);
    }
    if (r1) return ct == 1 ? launch_sens<C, 1, true>(h, mp, sp, ncol, b, s, hook) : launch_sens<C, 2, true>(h, mp, sp, ncol, b, s, hook);
    if (hook) return fail(h, CRNN_ERR_UNSUPPORTED, "the AutoSwitch fast path needs structured seed columns");
    return ct == 1 ? launch_sens<C, 1, false>(h, mp, sp, ncol, b, s) : launch_sens<C, 2, false>(h, mp, sp, ncol, b, s);
  });
}

// The on-device training loop (kernel_train.cuh): n_steps optimiser steps of `batch` experiments each, picked
// by `order` from a device-resident dataset; p and the optimiser state are read from and written back to host memory once.
template <class C>
int train_impl(crnn_handle* h, const crnn_model* m, const crnn_opts* o, const crnn_train_opts* t, const crnn_dataset* ds,
               const int64_t* order, int64_t n_steps, const double* yscale, int32_t loss_kind, double* p, double* opt_state,
               double* step_loss, double* step_gnorm) {
  // device p2vec kernels: 2 = case2/case2.jl:91-99 (6 species, 3 reactions, F1), 1 = case1/case1.jl:70-78 (5 species, 4 reactions, F0),
  // 3 = case3/case3.jl:42-53 (9 species, 8 reactions, F0 with out_scale; 153 parameters: five warps per trajectory),
  // 4 = robertson/rober_crnn.jl:85-96 (3 species, 6 reactions, F0 with out_scale; 43 parameters: Rosenbrock23 sensitivities, two column tiles)
  constexpr int PK = (C::NS == 6 && C::NR == 3 && C::KIND == 1) ? 2 : ((C::NS == 5 && C::NR == 4 && C::KIND == 0) ? 1 :
                     ((C::NS == 9 && C::NR == 8 && C::KIND == 0) ? 3 : ((C::NS == 3 && C::NR == 6 && C::KIND == 0) ? 4 : 0)));
  if constexpr (PK == 0) {
    return fail(h, CRNN_ERR_UNSUPPORTED, "the on-device training loop has device p2vec kernels for case1 (5 x 4, F0), case2 (6 x 3, F1), case3 (9 x 8, F0) and robertson (3 x 6, F0)");
  } else {
    constexpr int NP = PK == 2 ? C::NR * (C::NS + 2) + 1 : (PK >= 3 ? C::NR * (2 * C::NS + 1) + 1 : C::NR * (C::NS + 1));   // 25 / 153, 43 / 24
    constexpr bool ROS = (PK == 4);                      // the robertson script integrates with Rosenbrock23 (rober_crnn.jl:33)
    constexpr int WPT = PK == 3 ? 5 : 1;                 // warps sharing a trajectory (np + 1 columns over 32-lane tiles)
    constexpr int CT = ROS ? 2 : 1;                      // column tiles per lane
    constexpr int COLS = 32 * WPT * CT;
    if (t->p2vec_kind != PK) return fail(h, CRNN_ERR_UNSUPPORTED, "p2vec_kind does not match the model: 1 = case1.jl:70-78, 2 = case2.jl:91-99, 3 = case3.jl:42-53, 4 = rober_crnn.jl:85-96");
    if (o->alg != (ROS ? CRNN_ALG_ROSENBROCK23 : CRNN_ALG_TSIT5) || o->sens_mode != CRNN_SENS_FORWARD)
      return fail(h, CRNN_ERR_UNSUPPORTED, "the on-device training loop runs forward sensitivities through Tsit5 (case1 / case2 / case3) or Rosenbrock23 (robertson)");
    if (t->batch < 1 || n_steps < 0) return fail(h, CRNN_ERR_BAD_ARG, "bad batch / n_steps");
    if (PK < 3 && m->out_scale) return fail(h, CRNN_ERR_UNSUPPORTED, "case1 / case2 have no out_scale");
    if (t->n_save_used)
      for (int64_t q = 0; q < n_steps * t->batch; ++q)
        if (t->n_save_used[q] < 1 || t->n_save_used[q] > o->n_save) return fail(h, CRNN_ERR_BAD_ARG, "n_save_used: 1 .. n_save");
    const int batch = t->batch;
    for (int64_t q = 0; q < n_steps * batch; ++q)
      if (order[q] < 0 || order[q] >= ds->N) return fail(h, CRNN_ERR_BAD_ARG, "order: dataset row index out of range");
    ModelP<C> mp{}; SolveP<C> sp; Packed pk;
    crnn_model mm = *m;
    std::vector<double> zw(C::NW, 0.0);   // pack() wants weight pointers; the real weights come from the device p2vec
    mm.w_in = zw.data(); mm.w_b = zw.data(); mm.w_out = zw.data();
    int rc = pack<C>(h, &mm, o, yscale, loss_kind, mp, sp, pk);
    if (rc) return rc;
    const int ncol = NP + 1;
    if (o->err_norm_includes_sens && !o->err_norm_mean_over_state_only) sp.norm_cnt = (double)C::N * ncol;
    if (o->err_norm_includes_sens) sp.eig_cnt = (double)C::N * ncol;
    cudaStream_t st = h->s_compute;
    CK(cudaSetDevice(h->device));
    rc = upload_cfg<C>(h, o, pk, sp, st);
    if (rc) return rc;
    // device block: p[NP] | state[2NP+4] | ModelP | rows[2*NR*COLS] | desc[COLS] | loss_sum[2] | grad_sum[NP] | order | step_loss | step_gnorm | out_scale[NS]
    const size_t n_mp = (sizeof(ModelP<C>) + 7) / 8, n_rows = 2 * C::NR * COLS, n_desc = 3 * COLS;
    const size_t n_order = (size_t)n_steps * batch;
    const size_t total = NP + (2 * NP + 4) + n_mp + n_rows + n_desc + 2 + NP + n_order + 2 * (size_t)n_steps + 8 + C::NS + (n_order + 1) / 2;
    CK(h->train.reserve(total * sizeof(double)));
    double* d_p = h->train.as<double>(); double* d_st = d_p + NP;
    ModelP<C>* d_mp = reinterpret_cast<ModelP<C>*>(d_st + 2 * NP + 4);
    double* d_rows = reinterpret_cast<double*>(d_mp) + n_mp; R1Desc* d_desc = reinterpret_cast<R1Desc*>(d_rows + n_rows);
    double* d_lsum = d_rows + n_rows + n_desc; double* d_gsum = d_lsum + 2;
    long long* d_order = reinterpret_cast<long long*>(d_gsum + NP);
    double* d_sloss = reinterpret_cast<double*>(d_order + n_order); double* d_sgn = d_sloss + n_steps;
    double* d_oscale = d_sgn + n_steps;
    if (m->out_scale) CK(cudaMemcpyAsync(d_oscale, m->out_scale, C::NS * sizeof(double), cudaMemcpyHostToDevice, st));
    int* d_nsu = reinterpret_cast<int*>(d_oscale + C::NS);    // per visited experiment: save points used (or none)
    if (t->n_save_used && n_order) CK(cudaMemcpyAsync(d_nsu, t->n_save_used, n_order * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_p, p, NP * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_st, opt_state, (2 * NP + 4) * sizeof(double), cudaMemcpyHostToDevice, st));
    if (n_order) CK(cudaMemcpyAsync(d_order, order, n_order * sizeof(long long), cudaMemcpyHostToDevice, st));
    CK(h->d_loss.reserve(batch * sizeof(double)));
    CK(h->d_nsaved.reserve(batch * sizeof(int)));
    CK(h->d_ret.reserve(batch * sizeof(int)));
    CK(h->d_grad_each.reserve((size_t)batch * NP * sizeof(double)));
    h->last_grad_np = -1; h->last_grad_n = -1;
    TrainP T{};
    T.optimiser = t->optimiser; T.np = NP; T.eta = t->eta; T.beta1 = t->beta1; T.beta2 = t->beta2; T.eps = t->eps;
    T.weight_decay = t->weight_decay; T.expdecay_decay = t->expdecay_decay; T.expdecay_clip = t->expdecay_clip;
    T.expdecay_step = t->expdecay_eta > 0 ? t->expdecay_step : 0; T.grad_max = t->grad_max;
    // one warp per trajectory, eight per block - or, for case3, two groups of five warps per block; robertson: the
    // Rosenbrock23 kernel's four warps (launch_sens' / launch_ros_sens' shapes)
    constexpr int WARPS = ROS ? 4 : (WPT == 1 ? 8 : 2 * WPT);
    constexpr int GROUPS = WARPS / WPT;
    size_t smem = 0;
    if constexpr (ROS) {
      smem = sizeof(SensSmem<C, CT, true>) + WARPS * sizeof(RosWarpBuf<C, CT>);
      CK(cudaFuncSetAttribute(k_rosenbrock23_sens<C, CT, WARPS, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    } else {
      smem = sizeof(SensSmem<C, 1, true, WPT>) + WARPS * sizeof(WarpBuf<C, 1>);
      if (smem > 227 * 1024) return fail(h, CRNN_ERR_UNSUPPORTED, "model too large for the forward-sensitivity kernel's shared memory");
      CK(cudaFuncSetAttribute(k_tsit5_sens<C, 1, WARPS, (WPT == 1 ? 2 : 1), true, WPT, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const unsigned blocks = (unsigned)std::min<long long>((WPT == 1 ? 2LL : 1LL) * h->num_sms, (batch + GROUPS - 1) / GROUPS);
    unsigned long long* queue = h->ctr.as<unsigned long long>();
    const int nb_red = (int)std::max<long long>(1, std::min<long long>(4LL * h->num_sms, (batch + 63) / 64));
    CK(h->partial.reserve((size_t)nb_red * NP * sizeof(double)));
    for (int64_t s = 0; s < n_steps; ++s) {
      if constexpr (PK == 2) k_p2vec_case2<C><<<1, 64, 0, st>>>(d_p, m->lb, m->ub, m->gas_R, d_mp, d_rows, d_desc);
      else if constexpr (PK == 3) k_p2vec_case3<C, COLS><<<1, 128, 0, st>>>(d_p, m->lb, m->ub, m->out_scale ? d_oscale : nullptr, d_mp, d_rows, d_desc);
      else if constexpr (PK == 4) k_p2vec_robertson<C, COLS><<<1, 64, 0, st>>>(d_p, m->lb, m->ub, m->out_scale ? d_oscale : nullptr, d_mp, d_rows, d_desc);
      else k_p2vec_case1<C><<<1, 64, 0, st>>>(d_p, m->lb, m->ub, t->p2vec_b0, d_mp, d_rows, d_desc);
      CK(cudaMemsetAsync(queue, 0, sizeof(unsigned long long), st));
      const int* nsu = t->n_save_used ? d_nsu + s * batch : nullptr;
      if constexpr (ROS)
        k_rosenbrock23_sens<C, CT, WARPS, 3, true><<<blocks, WARPS * 32, smem, st>>>(
            mp, sp, d_rows, d_desc, ncol, ds->u0[0].as<double>(), nsu, batch, ds->data[0].as<double>(), h->d_loss.as<double>(),
            h->d_grad_each.as<double>(), nullptr, h->d_nsaved.as<int>(), h->d_ret.as<int>(), nullptr, queue, d_order + s * batch, d_mp);
      else
        k_tsit5_sens<C, 1, WARPS, (WPT == 1 ? 2 : 1), true, WPT, false, true><<<blocks, WARPS * 32, smem, st>>>(
            mp, sp, d_rows, d_desc, ncol, ds->u0[0].as<double>(), nsu, batch, ds->data[0].as<double>(), h->d_loss.as<double>(),
            h->d_grad_each.as<double>(), nullptr, h->d_nsaved.as<int>(), h->d_ret.as<int>(), nullptr, queue, d_order + s * batch,
            nullptr, nullptr, d_mp);
      k_grad_reduce<<<nb_red, 256, 0, st>>>(h->d_grad_each.as<double>(), batch, NP, h->partial.as<double>(), d_gsum,
                                            reinterpret_cast<unsigned int*>(h->ctr.as<unsigned long long>() + 1));
      k_train_loss_sum<<<1, 256, 0, st>>>(h->d_loss.as<double>(), batch, d_lsum);
      k_optim_step<<<1, 256, 0, st>>>(T, d_gsum, d_lsum, d_p, d_st, step_loss ? d_sloss : nullptr, step_gnorm ? d_sgn : nullptr, s);
      h->launches += 5;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(p, d_p, NP * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(opt_state, d_st, (2 * NP + 4) * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (step_loss && n_steps) CK(cudaMemcpyAsync(step_loss, d_sloss, n_steps * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (step_gnorm && n_steps) CK(cudaMemcpyAsync(step_gnorm, d_sgn, n_steps * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return CRNN_OK;
  }
}

}  // namespace crnn_host
