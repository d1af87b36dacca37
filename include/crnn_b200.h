/* crnn_b200.h — C-ABI of the B200-native batched CRNN neural-ODE engine.
 *
 * This is the drop-in boundary for the ONE hot path of DENG-MIT/CRNN: the
 * `solve(...)` call inside `predict_neuralode` and the
 * `ForwardDiff.gradient(x -> loss_neuralode(x, i_exp), p)` call in the training
 * loop.  The reference has no FFI of its own (it is 22 Julia scripts over
 * un-vendored SciML packages), so every entry point below cites the reference
 * call site it replaces (paths relative to the reference tree):
 *
 *   crnn_solve_batch       <- predict_neuralode: case1/case1.jl:92-97,
 *                             case2/case2.jl:124-128, case3/case3.jl:172-176,
 *                             robertson/rober_crnn.jl:123-136
 *   crnn_loss_grad_batch   <- ForwardDiff.gradient(loss_neuralode):
 *                             case2/case2.jl:132-137,195;
 *                             robertson/rober_crnn.jl:139-144,219;
 *                             Zygote.forwarddiff: case1/case1.jl:195-199,
 *                             case3/case3.jl:265-270
 *
 * Plain C: pointers and sizes only.  All matrices are COLUMN-MAJOR (Julia
 * layout), exactly what the scripts' `p2vec` returns.  Batched buffers are
 * laid out trajectory-slowest:  u0[n_state, N], data[n_obs, n_save, N],
 * pred[n_obs, n_save, N].
 *
 * Ownership: the caller owns every buffer.  With opts.buffers_on_device == 0
 * the batched buffers are host memory and the library does the H2D/D2H copies
 * itself (synchronous call).  With buffers_on_device == 1 they are device
 * pointers on the handle's device and the call only enqueues work on
 * opts.stream (a cudaStream_t; NULL = legacy default stream); the small
 * model/option arrays (weights, tolerances, saveat, obs_idx, seed) are always
 * host memory.
 *
 * Errors: 0 on success, a negative crnn_status otherwise (message through
 * crnn_last_error).  Solver outcomes are PER TRAJECTORY in retcode[] (values
 * mirror SciMLBase.ReturnCode, which robertson/rober_crnn.jl:130 tests), and
 * n_saved[i] says how many save columns of trajectory i are valid (the
 * reference tolerates truncated solutions: robertson/rober_crnn.jl:141).
 * A batch never aborts because one trajectory failed.
 */
#ifndef CRNN_B200_H
#define CRNN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRNN_B200_VERSION 200 /* 0.2.0: DiffEqBase dual norm (mean over n_state*(1+np)), qsteady dead-band, device-resident
                                 datasets, single-process multi-GPU handles */

typedef struct crnn_handle crnn_handle;

enum crnn_status {
  CRNN_OK = 0,
  CRNN_ERR_BAD_ARG = -1,
  CRNN_ERR_CUDA = -2,
  CRNN_ERR_UNSUPPORTED = -3,
  CRNN_ERR_NO_DEVICE = -4
};

/* RHS flavours. */
enum crnn_rhs_kind {
  CRNN_RHS_F0 = 0, /* du = s .* W_out*exp(W_in'*log(clamp(u,lb,ub)) + b): case1.jl:80-83, case3.jl:162-166, rober_crnn.jl:113-116 */
  CRNN_RHS_F1_ARRH_TSTATE = 1, /* state [X;T], x=[log clamp X; -1/(R T)], dT/dt=0: case2.jl:113-118 */
  CRNN_RHS_F2_MASSFRAC_TP = 2, /* mass fractions under tabulated T(t), P(t) (non-autonomous):
                                  Y=clamp(u,lb,ub), rho=P/(8314.46261815324 T sum(Y/MW)), C=rho Y/MW 1e3,
                                  x=[log clamp(C,lb,ub); -1/(R T); log T], du = W_out*exp(W_in'x+b) .* MW / rho .* out_scale:
                                  HyChem/crnn_pyrolysis_mass.jl:107-114,121-131 */
  CRNN_RHS_F4_MLP_AUG = 4,     /* CRNN whose input vector is AUGMENTED by a small MLP of the state ("neural closure" for hidden species):
                                  u_ = [u ; mlp(u)] (yeast-glycolysis/yeast_glycolysis.jl:128-142: 7 observed + 5 hidden species, du =
                                  (W_out*exp(W_in'*log(clamp(u_,lb,ub)) + b))[1:ns] .+ w_J) or u_ = [u1 ; mlp(u[1,3]) ; u3]
                                  (robertson/rober_crnn_qssa.jl:111-126); see the mlp_* / aug_src / w_J fields.  Predict path
                                  (crnn_solve_batch): the stiff steppers use the scripts' own FINITE-DIFFERENCE Jacobian
                                  (TRBDF2 / Rosenbrock23(autodiff=false)).  Loss + gradient (crnn_loss_grad_batch): the adjoint
                                  sens_modes with Tsit5, in the extended weight space [.. ; w_J ; mlp_params] */
  CRNN_RHS_F5_TRAMP = 3        /* species under a tabulated temperature programme T(t), no density map:
                                  x=[log clamp(u,lb,ub); -1/(gas_R T(t)); log T(t)], du = W_out*exp(W_in'x+b) .* out_scale:
                                  Cathode/src/network.jl:68-80 and Cathode_NCM333_UQ/src_333/network.jl:153-168 (there
                                  w_in = [diag(reaction orders); (Ea 1e5)'; b'], w_b = ln A, gas_R = 8.314, T = T0 + beta/60 t
                                  as a two-knot table); n_in = n_species + 2, tab_P unused */
};

enum crnn_alg {
  CRNN_ALG_TSIT5 = 0,        /* case1.jl:28, case3.jl:29, non-stiff half of case2.jl:26 */
  CRNN_ALG_ROSENBROCK23 = 1, /* rober_crnn.jl:33 */
  CRNN_ALG_KENCARP4 = 2,     /* BASELINE config 5 (not in the reference) */
  CRNN_ALG_AUTO_TSIT5_ROS23 = 3,/* AutoTsit5(Rosenbrock23()): case2.jl:26, HyChem/crnn_pyrolysis_mass.jl:29 — Tsit5 with
                                   OrdinaryDiffEq's AutoSwitch stiffness detection (maxstiffstep 10, maxnonstiffstep 3,
                                   tolerances 9/10, dtfac 2) switching to Rosenbrock23 (analytic J) and back;
                                   crnn_stats.n_jac counts the Rosenbrock23 step attempts */
  CRNN_ALG_TRBDF2 = 4,          /* TRBDF2 as an ESDIRK with simplified Newton (the stiff half of the next one) */
  CRNN_ALG_AUTO_TSIT5_TRBDF2 = 5 /* AutoTsit5(TRBDF2()): Cathode/src/network.jl:102, Cathode_NCM333_UQ/src_333/network.jl,
                                   yeast-glycolysis/yeast_glycolysis.jl:33 — the same AutoSwitch, TRBDF2 as the stiff stepper
                                   (predict, and forward sensitivities for F0 / F1 / F5: duals through the Newton iterations);
                                   crnn_stats.n_jac counts the Jacobian factorisations */
};

enum crnn_sens_mode {
  CRNN_SENS_NONE = 0,
  CRNN_SENS_FORWARD = 1,       /* duals-through-the-solver semantics: case2.jl:195 */
  CRNN_SENS_INTERP_ADJOINT = 2, /* BASELINE config 4 (not in the reference): continuous adjoint, jumps at save times */
  CRNN_SENS_DISCRETE_ADJOINT = 3 /* reverse-mode through the recorded Tsit5 steps + dense output: the forward-mode
                                    gradient (value-only error norm) at a cost independent of np and n_save */
};

enum crnn_loss_kind {
  CRNN_LOSS_MAE_SCALED = 0, /* mean|d/ys - clamp(pred)/ys|: case2.jl:132-137, rober_crnn.jl:139-144 */
  CRNN_LOSS_MAE_LOG = 1,    /* mean|log clamp d - log clamp pred|: case3.jl:183-190 */
  CRNN_LOSS_MSE = 2         /* mean (d - pred)^2 / ys^2: Cathode_NCM333_UQ/src_333/network.jl:262-275 */
};

/* Per-trajectory return codes (SciMLBase.ReturnCode values). */
enum crnn_retcode {
  CRNN_RET_DEFAULT = 0,
  CRNN_RET_SUCCESS = 1,
  CRNN_RET_TERMINATED = 2,
  CRNN_RET_DTNAN = 3,
  CRNN_RET_MAXITERS = 4,
  CRNN_RET_DTLESSTHANMIN = 5,
  CRNN_RET_UNSTABLE = 6
};

/* Physical CRNN weights, exactly what the scripts' p2vec returns. */
typedef struct crnn_model {
  int32_t n_state;   /* length of u (F1: n_species + 1) */
  int32_t n_species; /* rows of w_out */
  int32_t n_in;      /* rows of w_in (== n_state for F0 and F1, n_species + 2 for F2 and F5) */
  int32_t n_reac;    /* columns of w_in / w_out */
  int32_t rhs_kind;  /* crnn_rhs_kind */
  int32_t n_tab;     /* F2 only: length of tab_t / tab_T / tab_P (>= 2) */
  double lb, ub;     /* clamp bounds inside the RHS; ub may be +Inf (rober_crnn.jl:114) */
  double gas_R;      /* F1, F2: 1.98720425864083e-3 (case2.jl:113; a Float32 literal at crnn_pyrolysis_mass.jl:106) */
  const double* out_scale; /* [n_species] or NULL: dy_std_ (case3.jl:165), dydt_scale (rober_crnn.jl:115) */
  const double* w_in;      /* [n_in x n_reac] col-major */
  const double* w_b;       /* [n_reac] */
  const double* w_out;     /* [n_species x n_reac] col-major */
  /* F2 only (NULL otherwise): molar masses l_MW (crnn_pyrolysis_mass.jl:58) and the T(t), P(t) tables the script
   * interpolates linearly (itpT, itpP, :103-104); tab_t ascending and covering [t0, t1]. */
  const double* mw;        /* [n_species] */
  const double* tab_t;     /* [n_tab] */
  const double* tab_T;     /* [n_tab] */
  const double* tab_P;     /* [n_tab] */
  /* Observable post-map (NULL: the observed quantities are rows of u).  With w_obs the ONE observed quantity is
   * y = sum_j w_obs[j] r_j(u(t), t), r = exp(W_in'x + b): the heat release HRR_getter(ts, sol) * w_delH of
   * Cathode/src/network.jl:82-91,121 (n_obs must be 1, obs_idx is ignored).  The seed matrix then carries n_reac extra
   * rows, ordered [vec(w_in); w_b; vec(w_out); w_obs]. */
  const double* w_obs;     /* [n_reac] or NULL */
  /* F4 only (zero / NULL otherwise).  The MLP is Flux's Chain(Dense(d0, d1, gelu), ..., Dense(d_{L-1}, d_L, act_out)) with its
   * parameters in Flux.destructure order: per layer W (d_out x d_in, column-major), then b (yeast_glycolysis.jl:137-142). */
  int32_t mlp_n_layers;    /* L >= 1 */
  int32_t mlp_act_out;     /* last layer: 0 softplus (yeast :141), 1 exp (rober_crnn_qssa.jl:120); hidden layers: gelu (tanh form) */
  const int32_t* mlp_dims;   /* [L + 1]: d0 (number of state rows fed to the MLP), d1, ..., d_L (number of hidden species); all <= 32 */
  const int32_t* mlp_in_idx; /* [d0] 0-based state rows fed to the MLP (yeast: all; QSSA: rows 0 and 2) */
  const double* mlp_params;  /* [sum_l d_l*d_{l+1} + d_{l+1}] */
  const int32_t* aug_src;    /* [n_in] source of input row k of the CRNN: >= 0 state row, < 0 MLP output -1 - value */
  const double* w_J;         /* [n_species] additive source term (yeast :131,143) or NULL */
} crnn_model;

typedef struct crnn_opts {
  int32_t alg;                    /* crnn_alg */
  int32_t sens_mode;              /* crnn_sens_mode (loss_grad only) */
  int32_t err_norm_includes_sens; /* 1: dual partials take part in step-size control (DiffEqBase norm over Duals) */
  int32_t n_save;                 /* length of saveat */
  int32_t n_obs;                  /* length of obs_idx */
  int32_t n_abstol;               /* 1 or n_state (rober_crnn.jl:34 writes a vector) */
  int32_t n_reltol;               /* 1 or n_state */
  int32_t buffers_on_device;      /* 0: batched buffers are host memory; 1: device pointers */
  int64_t maxiters;               /* accepted+rejected step cap (case1.jl:32, rober_crnn.jl:30) */
  double t0, t1;                  /* tspan */
  double pred_clamp_lo, pred_clamp_hi; /* clamp applied to saved states: (-ub,ub) case2.jl:126, (lb,ub) case3.jl:174, (-Inf,Inf) robertson */
  const double* abstol;
  const double* reltol;
  const double* saveat;           /* [n_save], ascending, inside [t0,t1] */
  const int32_t* obs_idx;         /* [n_obs] 0-based rows of u that are observed (case2.jl:131) */
  double qmin, qmax, gamma, beta1, beta2; /* step controller; <= 0 selects the OrdinaryDiffEq defaults */
  void* stream;                   /* cudaStream_t, used when buffers_on_device == 1 */
  /* Dead-band of step_accept_controller!: qsteady_min <= q <= qsteady_max keeps dt.  <= 0 selects OrdinaryDiffEq's
   * defaults: 1 / 1 (no dead-band) for Tsit5 and the AutoTsit5 composite, 1 / 1.2 for the adaptive implicit
   * algorithms Rosenbrock23 and KenCarp4 (qsteady_max_default(::OrdinaryDiffEqAdaptiveImplicitAlgorithm) = 6//5). */
  double qsteady_min, qsteady_max;
  /* Divisor of the dual-aware norms when err_norm_includes_sens == 1.
   * 0 (default): DiffEqBase's ODE_DEFAULT_NORM over Dual arrays, sqrt(sum(sse) / totallength(u)) — the mean runs over
   *    all n_state*(1+np) numbers;
   * 1: the mean runs over the n_state rows only (what releases before 0.2.0 did). */
  int32_t err_norm_mean_over_state_only;
  int32_t reserved0;
} crnn_opts;

typedef struct crnn_stats {
  int32_t n_accept, n_reject, n_rhs, n_jac;
  double t_reached, dt_last;
} crnn_stats;

/* device_id < 0 selects the current CUDA device.  Fails with
 * CRNN_ERR_NO_DEVICE when no CUDA device is usable: there is NO CPU fallback. */
int crnn_create(crnn_handle** h, int device_id);
void crnn_destroy(crnn_handle* h);
const char* crnn_last_error(const crnn_handle* h);
int crnn_version(void);
/* Number of CUDA kernels this handle has launched since creation. */
int64_t crnn_launch_count(const crnn_handle* h);

/* Optional device-side timing of the solver kernels (the dominant launches): between
 * crnn_profile_begin and crnn_profile_end every solver-kernel launch is bracketed by CUDA
 * events on its own stream; _end synchronises and returns their summed duration and count.
 * Used by bench.py for the roofline line; off by default (no events recorded). */
int crnn_profile_begin(crnn_handle* h);
int crnn_profile_end(crnn_handle* h, double* total_ms, int64_t* n_launches);

/* predict_neuralode for a batch of N initial conditions.
 * Every alg (Tsit5, Rosenbrock23, KenCarp4, AutoTsit5(Rosenbrock23)) and every rhs_kind is served for n_state, n_in,
 * n_reac <= 32; the (n_species, n_reac, rhs_kind) of the reference scripts run dimension-specialised kernels.
 *   u0      [n_state, N]
 *   n_save_used [N] or NULL: trajectory i integrates only to
 *           saveat[n_save_used[i]-1] (random time truncation, rober_crnn.jl:218)
 *   pred    [n_obs, n_save, N]  clamped saved states (columns >= n_saved[i] are left zero)
 *   n_saved [N], retcode [N], stats [N] or NULL                              */
int crnn_solve_batch(crnn_handle* h, const crnn_model* m, const crnn_opts* o,
                     const double* u0, int64_t N, const int32_t* n_save_used,
                     double* pred, int32_t* n_saved, int32_t* retcode,
                     crnn_stats* stats);

/* loss_neuralode + its gradient for a batch.
 * sens_mode FORWARD (ForwardDiff semantics: every dual column rides the adaptive solve, partials in the error norm):
 *   Tsit5 (np <= 255) and Rosenbrock23 (np <= 63, n_species <= 6) for the reference scripts' dimensions, F0 / F1;
 *   with err_norm_includes_sens = 0 and Tsit5 every other model (<= 32, F2) is served by the discrete adjoint, which
 *   computes exactly that derivative.
 * sens_mode INTERP_ADJOINT / DISCRETE_ADJOINT: Tsit5, any dimensions <= 32, any np, n_w <= 512, F0 / F1 / F2 / F4, MAE losses.
 *   dW_dp   [n_w, np] col-major HOST seed matrix = Jacobian of p2vec, with the
 *           n_w = n_reac*(n_in + 1 + n_species) rows ordered
 *           [vec(w_in); w_b; vec(w_out)]
 *           (F4, adjoint modes: n_w += n_species + #mlp_params, rows [...; w_J; mlp_params] — the gradient covers the
 *           CRNN weights AND the Flux chain, yeast_glycolysis.jl:136-145,246)
 *   data    [n_obs, n_save, N], yscale [n_obs] (host; ignored for MAE_LOG)
 *   loss    [N] per-trajectory loss (NaN when a trajectory saved nothing)
 *   grad_sum[np] sum over trajectories of d loss_i / d p, accumulated in a fixed
 *           order (deterministic); host or device memory like the other
 *           batched buffers (opts.buffers_on_device)
 *   pred    may be NULL                                                     */
int crnn_loss_grad_batch(crnn_handle* h, const crnn_model* m, const crnn_opts* o,
                         const double* dW_dp, int32_t np,
                         const double* u0, int64_t N, const int32_t* n_save_used,
                         const double* data, const double* yscale, int32_t loss_kind,
                         double* loss, double* grad_sum, double* pred,
                         int32_t* n_saved, int32_t* retcode, crnn_stats* stats);

/* Per-trajectory gradients of the LAST crnn_loss_grad_batch call made with CRNN_SENS_FORWARD:
 * dst[np, N] (column-major: np fastest) <- d loss_i / d p.  What robertson/rober_crnn_lm.jl's
 * Levenberg-Marquardt variant needs (ForwardDiff.jacobian of the per-experiment losses, :216-218).
 * dst is host memory (on_device = 0) or device memory (on_device = 1, copied on `stream`). */
int crnn_copy_grad_each(crnn_handle* h, double* dst, int64_t N, int32_t np, int32_t on_device, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Device-resident training sets and multi-GPU handles.
 *
 * The scripts' epoch loop (case2/case2.jl:192-207) evaluates loss and gradient over the SAME u0_list /
 * ode_data_list (case2.jl:62-83) at every optimiser step; only p changes.  crnn_dataset_create uploads the
 * arrays once; crnn_loss_grad_indexed then moves only weights + seed in and [sum loss, n, grad_sum] out.
 * --------------------------------------------------------------------------------------------- */
typedef struct crnn_dataset crnn_dataset;

/* One handle over several CUDA devices of this process (device_ids == NULL: devices 0..n_devices-1).  Datasets are
 * split into contiguous shards, one per device; every device solves its rows on its own stream and the only exchange,
 * [sum loss, n, grad_sum] (np + 2 doubles), is an ncclAllReduce over NVLink enqueued on the compute streams
 * (ncclCommInitAll; libnccl.so.2 is loaded with dlopen, CRNN_ERR_UNSUPPORTED if n_devices > 1 and it is missing).
 * crnn_solve_batch / crnn_loss_grad_batch on such a handle take HOST buffers and shard them the same way. */
int crnn_create_multi(crnn_handle** h, const int32_t* device_ids, int32_t n_devices);
int32_t crnn_device_count(const crnn_handle* h);

/* u0 [n_state, N], data [n_obs, n_save, N] (host) -> device memory of the handle's device(s). */
int crnn_dataset_create(crnn_handle* h, const double* u0, const double* data, int32_t n_state, int32_t n_obs,
                        int32_t n_save, int64_t N, crnn_dataset** ds);
void crnn_dataset_destroy(crnn_dataset* ds);
int64_t crnn_dataset_size(const crnn_dataset* ds);

/* crnn_loss_grad_batch over rows idx[0..n_idx) of a dataset (idx == NULL: every row, in order) — the mini-batch
 * `for i_exp in randperm(n_exp_train)` picks (case2.jl:194).  All pointers are HOST memory.
 *   n_save_used [n_idx] or NULL (per picked row, rober_crnn.jl:218)
 *   loss_sum    [2]: sum of the finite per-trajectory losses and how many there were (mean loss = [0]/[1])
 *   grad_sum    [np]: sum over the picked rows of d loss_i / d p (all devices)
 *   loss, n_saved, retcode, stats: [n_idx] each or NULL (skipping them skips their device->host copies) */
int crnn_loss_grad_indexed(crnn_handle* h, const crnn_model* m, const crnn_opts* o, const double* dW_dp, int32_t np,
                           const crnn_dataset* ds, const int64_t* idx, int64_t n_idx, const int32_t* n_save_used,
                           const double* yscale, int32_t loss_kind, double* loss_sum, double* grad_sum, double* loss,
                           int32_t* n_saved, int32_t* retcode, crnn_stats* stats);

/* ---------------------------------------------------------------------------------------------
 * On-device training loop: the scripts' epoch loop
 *     for i_exp in randperm(n_exp_train); grad = ForwardDiff.gradient(x -> loss_neuralode(x, i_exp), p); update!(opt, p, grad); end
 * (case2/case2.jl:192-198) with NOTHING returning to the host between optimiser steps: a p2vec kernel (the script's own
 * map), the forward-sensitivity kernel reading its weights from device memory, the gradient reduction and Flux's
 * ExpDecay -> ADAM / NADAM -> WeightDecay chain (case2.jl:31-32, case3.jl:20, rober_crnn.jl:19) incl. the 2-norm clip
 * (rober_crnn.jl:220-223) are enqueued back to back.  The host supplies the visiting order (its own randperm).
 * --------------------------------------------------------------------------------------------- */
typedef struct crnn_train_opts {
  int32_t p2vec_kind;   /* which script's p2vec runs on the device: 1 = case1/case1.jl:70-78, 2 = case2/case2.jl:91-99,
                           3 = case3/case3.jl:42-53 (model.out_scale = dy_std is folded in on the device),
                           4 = robertson/rober_crnn.jl:85-96 (opts.alg = Rosenbrock23, out_scale = dydt_scale) */
  int32_t optimiser;    /* 0 ADAM (+ weight_decay = ADAMW), 1 NADAM */
  int32_t batch;        /* experiments per optimiser step (the scripts: 1) */
  int32_t reserved;
  double eta, beta1, beta2, eps, weight_decay;
  double expdecay_eta;  /* > 0: a Flux.ExpDecay link before ADAM (case2.jl:31); its running eta lives in opt_state */
  double expdecay_decay, expdecay_clip;
  int64_t expdecay_step;
  double grad_max;      /* > 0: clip the gradient's 2-norm (rober_crnn.jl:29,221) */
  double p2vec_b0;      /* p2vec_kind 1: the bias offset b0 of case1/case1.jl:70 (-10) */
  const int32_t* n_save_used;  /* HOST [n_steps * batch] or NULL: per visited experiment, the number of save points its loss uses
                                  (`sample = rand(batchsize:datasize)`, rober_crnn.jl:218; tspan ends at saveat[n - 1]) */
} crnn_train_opts;

/* n_steps optimiser steps; step s uses dataset rows order[s*batch .. (s+1)*batch).  Single-device handle, forward
 * sensitivities through Tsit5 (p2vec_kind 1-3) or Rosenbrock23 (p2vec_kind 4).
 *   m          dimensions, clamps, gas_R and out_scale of the model (weight pointers ignored: the device p2vec produces them)
 *   p          [np] in/out
 *   opt_state  [2*np + 4] in/out: ADAM m, v, beta1^t, beta2^t, ExpDecay's current eta and count
 *              (a fresh run: zeros, beta1, beta2, expdecay_eta, 0)
 *   step_loss, step_gnorm  [n_steps] or NULL: mean loss and gradient norm (before clipping) of every step */
int crnn_train_steps(crnn_handle* h, const crnn_model* m, const crnn_opts* o, const crnn_train_opts* t,
                     const crnn_dataset* ds, const int64_t* order, int64_t n_steps, const double* yscale,
                     int32_t loss_kind, double* p, double* opt_state, double* step_loss, double* step_gnorm);

/* Parameter-batched loss + gradient ("particles"): P parameter sets x E experiments in ONE launch, trajectory (p, e)
 * integrating experiment e with the weights of particle p - the loop `for j = 1:size(p)[1] ... ForwardDiff.gradient`
 * of Cathode_NCM333_UQ/src_333/network.jl:222-260 (100 SVGD particles x 5 data sets, sequential there).
 *   m         dimensions, rhs_kind, clamps, gas_R, out_scale, mw, tab_t of the model; its weight pointers are ignored
 *   weights   [n_w, P] col-major: [vec(w_in); w_b; vec(w_out); (w_obs)] of every particle (n_w incl. w_obs when m->w_obs != NULL)
 *   dW_dp     [n_w, np, P]: each particle's seed matrix (its own Jacobian of p2vec)
 *   u0        [n_state, E], data [n_obs, n_save, E], n_save_used [E] or NULL
 *   tab_T, tab_P  [n_tab, E] per-experiment tables (F2 / F5; NULL: the model's own tables for every experiment)
 *   loss      [E, P] (experiment fastest), grad [np, P] = sum over e of d loss(e, p) / d p_p,
 *   n_saved, retcode, stats: [E, P] or NULL.  All pointers are HOST memory.
 * Forward mode only (sens_mode FORWARD; Tsit5, Rosenbrock23, AutoTsit5(Rosenbrock23); structured seed columns). */
int crnn_loss_grad_particles(crnn_handle* h, const crnn_model* m, const crnn_opts* o, const double* weights,
                             const double* dW_dp, int32_t np, int32_t P, const double* u0, int32_t E,
                             const int32_t* n_save_used, const double* data, const double* tab_T, const double* tab_P,
                             const double* yscale, int32_t loss_kind, double* loss, double* grad, int32_t* n_saved,
                             int32_t* retcode, crnn_stats* stats);

/* Diagnostic: evaluates the engine's device log / exp / pow (crnn_b200/csrc/lean_math.h) elementwise on host arrays,
 * op 0: y = log(x), 1: exp(x), 2: x^x2, 3: log10(x), 4: 10^x.  The same header compiles for the host; the tests
 * use this entry point to show the two copies agree bit for bit. */
int crnn_debug_lean_math(crnn_handle* h, int32_t op, const double* x, const double* x2, double* y, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* CRNN_B200_H */
