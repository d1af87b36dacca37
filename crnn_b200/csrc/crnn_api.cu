// crnn_api.cu — the C-ABI of libcrnn_b200.so (declared in include/crnn_b200.h).
//
// Entry points only: validation and dispatch on (n_species, n_reac, rhs_kind) to the
// dimension-specialised engines, which are compiled one translation unit per configuration
// (inst.cu with -DCRNN_NS/-DCRNN_NR/-DCRNN_KIND) so the build parallelises.
#include "crnn_host.cuh"

namespace crnn_host {
#define X(NS_, NR_, K_)                                                                                     \
  extern template int solve_impl<Cfg<NS_, NR_, K_>>(crnn_handle*, const crnn_model*, const crnn_opts*,      \
                                                    const HostIO&, int64_t);                                \
  extern template int loss_grad_impl<Cfg<NS_, NR_, K_>>(crnn_handle*, const crnn_model*, const crnn_opts*,  \
                                                        const double*, int, const double*, int,             \
                                                        const HostIO&, int64_t, double*);
CRNN_FOR_EACH_CFG(X)
#undef X
}  // namespace crnn_host
using namespace crnn_host;

extern "C" {

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
  ok = ok && h->ctr.reserve(64) == cudaSuccess && cudaMemset(h->ctr.p, 0, 64) == cudaSuccess;
  if (!ok) { crnn_destroy(h); return CRNN_ERR_CUDA; }
  *out = h;
  return CRNN_OK;
}

void crnn_destroy(crnn_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  DevBuf* bufs[] = {&h->cfg, &h->seed, &h->desc, &h->ctr, &h->partial, &h->d_grad_each, &h->d_grad_sum};
  for (DevBuf* b : bufs) b->release();
  for (int s = 0; s < kPipe; ++s) {
    DevBuf* sb[] = {&h->d_u0[s], &h->d_nsu[s], &h->d_data[s], &h->d_pred[s], &h->d_loss[s], &h->d_nsaved[s],
                    &h->d_ret[s], &h->d_stats[s]};
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

int64_t crnn_launch_count(const crnn_handle* h) { return h ? h->launches : 0; }

int crnn_profile_begin(crnn_handle* h) {
  if (!h) return CRNN_ERR_BAD_ARG;
  h->profiling = true;
  h->prof_used = 0;
  return CRNN_OK;
}

int crnn_profile_end(crnn_handle* h, double* total_ms, int64_t* n_launches) {
  if (!h) return CRNN_ERR_BAD_ARG;
  h->profiling = false;
  double tot = 0.0;
  for (size_t q = 0; q < h->prof_used; ++q) {
    float ms = 0.f;
    CK(cudaEventSynchronize(h->prof_events[q].second));
    CK(cudaEventElapsedTime(&ms, h->prof_events[q].first, h->prof_events[q].second));
    tot += ms;
  }
  if (total_ms) *total_ms = tot;
  if (n_launches) *n_launches = (int64_t)h->prof_used;
  h->prof_used = 0;
  return CRNN_OK;
}

int crnn_solve_batch(crnn_handle* h, const crnn_model* m, const crnn_opts* o, const double* u0, int64_t N,
                     const int32_t* n_save_used, double* pred, int32_t* n_saved, int32_t* retcode,
                     crnn_stats* stats) {
  if (!h) return CRNN_ERR_BAD_ARG;
  int rc = validate(h, m, o, N);
  if (rc) return rc;
  if (N > 0 && !u0) return fail(h, CRNN_ERR_BAD_ARG, "null u0");
  if (o->alg != CRNN_ALG_TSIT5 && o->alg != CRNN_ALG_ROSENBROCK23)
    return fail(h, CRNN_ERR_UNSUPPORTED, "alg not supported by solve_batch");
  CK(cudaSetDevice(h->device));
  HostIO io{u0, n_save_used, nullptr, pred, nullptr, n_saved, retcode, stats};
#define X(NS_, NR_, K_)                                                              \
  if (m->n_species == NS_ && m->n_reac == NR_ && m->rhs_kind == K_)                  \
    return solve_impl<Cfg<NS_, NR_, K_>>(h, m, o, io, N);
  CRNN_FOR_EACH_CFG(X)
#undef X
  return fail(h, CRNN_ERR_UNSUPPORTED, "no kernel instantiated for this (n_species, n_reac, rhs_kind)");
}

int crnn_loss_grad_batch(crnn_handle* h, const crnn_model* m, const crnn_opts* o, const double* dW_dp, int32_t np,
                         const double* u0, int64_t N, const int32_t* n_save_used, const double* data,
                         const double* yscale, int32_t loss_kind, double* loss, double* grad_sum, double* pred,
                         int32_t* n_saved, int32_t* retcode, crnn_stats* stats) {
  if (!h) return CRNN_ERR_BAD_ARG;
  int rc = validate(h, m, o, N);
  if (rc) return rc;
  if (N > 0 && (!u0 || !data || !loss)) return fail(h, CRNN_ERR_BAD_ARG, "null u0/data/loss");
  if (np < 0 || (np > 0 && !dW_dp)) return fail(h, CRNN_ERR_BAD_ARG, "bad seed matrix");
  if (loss_kind != CRNN_LOSS_MAE_SCALED && loss_kind != CRNN_LOSS_MAE_LOG)
    return fail(h, CRNN_ERR_BAD_ARG, "bad loss_kind");
  if (loss_kind == CRNN_LOSS_MAE_SCALED && !yscale) return fail(h, CRNN_ERR_BAD_ARG, "null yscale");
  if (o->sens_mode != CRNN_SENS_FORWARD)
    return fail(h, CRNN_ERR_UNSUPPORTED, "only CRNN_SENS_FORWARD is implemented");
  if (o->n_obs == 0 || o->n_save == 0) return fail(h, CRNN_ERR_BAD_ARG, "loss needs n_obs > 0 and n_save > 0");
  CK(cudaSetDevice(h->device));
  HostIO io{u0, n_save_used, data, pred, loss, n_saved, retcode, stats};
#define X(NS_, NR_, K_)                                                              \
  if (m->n_species == NS_ && m->n_reac == NR_ && m->rhs_kind == K_)                  \
    return loss_grad_impl<Cfg<NS_, NR_, K_>>(h, m, o, dW_dp, np, yscale, loss_kind, io, N, grad_sum);
  CRNN_FOR_EACH_CFG(X)
#undef X
  return fail(h, CRNN_ERR_UNSUPPORTED, "no kernel instantiated for this (n_species, n_reac, rhs_kind)");
}

}  // extern "C"
