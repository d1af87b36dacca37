// crnn_dataset.cu — device-resident training sets and the single-process multi-GPU handle of the C-ABI.
//
// The reference's training loop (case2/case2.jl:192-207) evaluates loss + gradient over the SAME u0_list / ode_data_list
// (case2.jl:62-83) at every optimiser step; only p changes.  crnn_dataset_create uploads those arrays ONCE,
// crnn_loss_grad_indexed then moves only the weights + seed in and [sum loss, n, grad] out per step.
//
// crnn_create_multi builds one parent handle over several devices of this process ("multi-GPU fan-out is
// internal"): a dataset is split into contiguous shards, every device solves its rows on its own compute stream, and the
// one exchange of the path — [sum loss, n, grad_sum] (np + 2 doubles) — is an ncclAllReduce enqueued on those compute
// streams right after the per-device reduction kernels (NCCL from ncclCommInitAll, loaded with dlopen so that a
// single-GPU deployment has no NCCL dependency).  The message is <= 2.3 KB: pure latency, nothing to overlap.
#include <dlfcn.h>
#include <nccl.h>

#include <thread>

#include "crnn_host.cuh"


namespace {

// ---- NCCL through dlopen ----
struct Nccl {
  void* lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string why;
  bool ok = false;
};

Nccl& nccl() {
  static Nccl n = [] {
    Nccl q;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    if (const char* e = std::getenv("CRNN_B200_NCCL_LIB")) q.lib = dlopen(e, RTLD_NOW | RTLD_GLOBAL);
    for (const char* nm : names)
      if (!q.lib) q.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (!q.lib) { q.why = std::string("cannot load libnccl.so.2: ") + dlerror(); return q; }
    auto sym = [&](const char* s) { return dlsym(q.lib, s); };
    q.CommInitAll = (decltype(q.CommInitAll))sym("ncclCommInitAll");
    q.CommDestroy = (decltype(q.CommDestroy))sym("ncclCommDestroy");
    q.GroupStart = (decltype(q.GroupStart))sym("ncclGroupStart");
    q.GroupEnd = (decltype(q.GroupEnd))sym("ncclGroupEnd");
    q.AllReduce = (decltype(q.AllReduce))sym("ncclAllReduce");
    q.GetErrorString = (decltype(q.GetErrorString))sym("ncclGetErrorString");
    q.ok = q.CommInitAll && q.CommDestroy && q.GroupStart && q.GroupEnd && q.AllReduce && q.GetErrorString;
    if (!q.ok) q.why = "libnccl.so.2 lacks a required symbol";
    return q;
  }();
  return n;
}

#define NK(call)                                                                                   \
  do {                                                                                             \
    ncclResult_t r__ = (call);                                                                     \
    if (r__ != ncclSuccess) {                                                                      \
      h->err = std::string(#call) + ": " + nccl().GetErrorString(r__);                             \
      return CRNN_ERR_CUDA;                                                                        \
    }                                                                                              \
  } while (0)

// [sum of the finite losses, number of them] in a FIXED order: thread t adds elements t, t + 1024, ... then a fixed tree.
__global__ void __launch_bounds__(1024) k_loss_sum(const double* __restrict__ loss, long long n, double* __restrict__ out) {
  __shared__ double s_sum[1024], s_cnt[1024];
  double s = 0.0, c = 0.0;
  for (long long i = threadIdx.x; i < n; i += 1024) {
    const double v = loss[i];
    if (v == v && fabs(v) != INFINITY) { s += v; c += 1.0; }
  }
  s_sum[threadIdx.x] = s; s_cnt[threadIdx.x] = c;
  __syncthreads();
  for (int w = 512; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) { s_sum[threadIdx.x] += s_sum[threadIdx.x + w]; s_cnt[threadIdx.x] += s_cnt[threadIdx.x + w]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = s_sum[0]; out[1] = s_cnt[0]; }
}

std::vector<crnn_handle*> devices_of(crnn_handle* h) {
  if (h->kids.empty()) return {h};
  return h->kids;
}

// One device's share of an indexed call, enqueued on its compute stream (no synchronisation):
// result = [sum loss, n finite, grad_sum(np)] in k->d_result, per-trajectory outputs in k->d_loss / d_nsaved / d_ret / d_stats.
int enqueue_share(crnn_handle* k, const crnn_model* m, const crnn_opts* o, const double* dW_dp, int np, const crnn_dataset* ds,
                  int dev, const long long* idx_local /* host, or NULL: all rows of the shard */, int64_t n,
                  const int32_t* nsu /* host [n] or NULL */, const double* yscale, int loss_kind, bool want_stats) {
  crnn_handle* h = k;
  CK(cudaSetDevice(k->device));
  CK(k->d_result.reserve((size_t)(np + 2) * sizeof(double)));
  CK(cudaMemsetAsync(k->d_result.p, 0, (size_t)(np + 2) * sizeof(double), k->s_compute));
  if (n == 0) return CRNN_OK;
  const long long* d_idx = nullptr;
  if (idx_local) {
    CK(k->d_idx.reserve(n * sizeof(long long)));
    CK(cudaMemcpyAsync(k->d_idx.p, idx_local, n * sizeof(long long), cudaMemcpyHostToDevice, k->s_compute));
    d_idx = k->d_idx.as<long long>();
  }
  const int* d_nsu = nullptr;
  if (nsu) {
    CK(k->d_nsu_ix.reserve(n * sizeof(int)));
    CK(cudaMemcpyAsync(k->d_nsu_ix.p, nsu, n * sizeof(int), cudaMemcpyHostToDevice, k->s_compute));
    d_nsu = k->d_nsu_ix.as<int>();
  }
  CK(k->d_loss.reserve(n * sizeof(double)));
  CK(k->d_nsaved.reserve(n * sizeof(int)));
  CK(k->d_ret.reserve(n * sizeof(int)));
  if (want_stats) CK(k->d_stats.reserve(n * sizeof(crnn_stats)));
  crnn_opts od = *o;
  od.buffers_on_device = 1;
  od.stream = (void*)k->s_compute;
  crnn_host::HostIO io{ds->u0[dev].as<double>(), d_nsu, ds->data[dev].as<double>(), nullptr, k->d_loss.as<double>(),
                       k->d_nsaved.as<int>(), k->d_ret.as<int>(), want_stats ? k->d_stats.as<crnn_stats>() : nullptr, d_idx};
  int rc = crnn_host::loss_grad_core(k, m, &od, dW_dp, np, yscale, loss_kind, io, n, k->d_result.as<double>() + 2);
  if (rc) return rc;
  k_loss_sum<<<1, 1024, 0, k->s_compute>>>(k->d_loss.as<double>(), n, k->d_result.as<double>());
  CK(cudaGetLastError());
  k->launches++;
  return CRNN_OK;
}

}  // namespace

namespace crnn_host {

void multi_release_comms(crnn_handle* h) {
  for (size_t d = 0; d < h->comms.size(); ++d)
    if (h->comms[d]) { cudaSetDevice(h->kids[d]->device); nccl().CommDestroy((ncclComm_t)h->comms[d]); }
  h->comms.clear();
}

// Host-buffer entry points on a multi-device handle: contiguous shards, one host thread per device (each child runs its own
// chunked H2D -> kernel -> D2H pipeline); the per-device gradient sums are added on the host in device order.
int multi_solve_batch(crnn_handle* h, const crnn_model* m, const crnn_opts* o, const double* u0, int64_t N,
                      const int32_t* n_save_used, double* pred, int32_t* n_saved, int32_t* retcode, crnn_stats* stats) {
  if (!m || !o) return fail(h, CRNN_ERR_BAD_ARG, "null model/opts");
  if (o->buffers_on_device) return fail(h, CRNN_ERR_UNSUPPORTED, "device buffers need a single-device handle");
  const int nd = (int)h->kids.size();
  const size_t ns = m->n_state, ps = (size_t)o->n_obs * o->n_save;
  std::vector<int> rcs(nd, 0);
  std::vector<std::thread> th;
  for (int d = 0; d < nd; ++d) {
    const int64_t lo = N * d / nd, n = N * (d + 1) / nd - lo;
    th.emplace_back([=, &rcs] {
      rcs[d] = crnn_solve_batch(h->kids[d], m, o, u0 ? u0 + lo * ns : nullptr, n, n_save_used ? n_save_used + lo : nullptr,
                                pred ? pred + lo * ps : nullptr, n_saved ? n_saved + lo : nullptr,
                                retcode ? retcode + lo : nullptr, stats ? stats + lo : nullptr);
    });
  }
  for (auto& t : th) t.join();
  for (int d = 0; d < nd; ++d)
    if (rcs[d]) { h->err = "device " + std::to_string(h->kids[d]->device) + ": " + h->kids[d]->err; return rcs[d]; }
  return CRNN_OK;
}

int multi_loss_grad_batch(crnn_handle* h, const crnn_model* m, const crnn_opts* o, const double* dW_dp, int32_t np,
                          const double* u0, int64_t N, const int32_t* n_save_used, const double* data, const double* yscale,
                          int32_t loss_kind, double* loss, double* grad_sum, double* pred, int32_t* n_saved,
                          int32_t* retcode, crnn_stats* stats) {
  if (!m || !o) return fail(h, CRNN_ERR_BAD_ARG, "null model/opts");
  if (o->buffers_on_device) return fail(h, CRNN_ERR_UNSUPPORTED, "device buffers need a single-device handle");
  const int nd = (int)h->kids.size();
  const size_t ns = m->n_state, ps = (size_t)o->n_obs * o->n_save;
  std::vector<int> rcs(nd, 0);
  std::vector<std::vector<double>> g(nd, std::vector<double>(std::max(np, 1), 0.0));
  std::vector<std::thread> th;
  for (int d = 0; d < nd; ++d) {
    const int64_t lo = N * d / nd, n = N * (d + 1) / nd - lo;
    th.emplace_back([=, &rcs, &g] {
      rcs[d] = crnn_loss_grad_batch(h->kids[d], m, o, dW_dp, np, u0 ? u0 + lo * ns : nullptr, n,
                                    n_save_used ? n_save_used + lo : nullptr, data ? data + lo * ps : nullptr, yscale,
                                    loss_kind, loss ? loss + lo : nullptr, g[d].data(), pred ? pred + lo * ps : nullptr,
                                    n_saved ? n_saved + lo : nullptr, retcode ? retcode + lo : nullptr,
                                    stats ? stats + lo : nullptr);
    });
  }
  for (auto& t : th) t.join();
  for (int d = 0; d < nd; ++d)
    if (rcs[d]) { h->err = "device " + std::to_string(h->kids[d]->device) + ": " + h->kids[d]->err; return rcs[d]; }
  if (grad_sum)
    for (int c = 0; c < np; ++c) {
      double s = 0.0;
      for (int d = 0; d < nd; ++d) s += g[d][c];
      grad_sum[c] = s;
    }
  return CRNN_OK;
}

}  // namespace crnn_host

using crnn_host::fail;

extern "C" {

int crnn_create_multi(crnn_handle** out, const int32_t* device_ids, int32_t n_devices) {
  if (!out || n_devices < 1 || n_devices > 64) return CRNN_ERR_BAD_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return CRNN_ERR_NO_DEVICE;
  std::vector<int> ids(n_devices);
  for (int d = 0; d < n_devices; ++d) {
    ids[d] = device_ids ? device_ids[d] : d;
    if (ids[d] < 0 || ids[d] >= ndev) return CRNN_ERR_BAD_ARG;
    for (int e = 0; e < d; ++e)
      if (ids[e] == ids[d]) return CRNN_ERR_BAD_ARG;
  }
  if (n_devices > 1 && !nccl().ok) return CRNN_ERR_UNSUPPORTED;  // no communicator, no multi-device handle
  crnn_handle* h = new crnn_handle();
  h->device = ids[0];
  for (int d = 0; d < n_devices; ++d) {
    crnn_handle* k = nullptr;
    int rc = crnn_create(&k, ids[d]);
    if (rc) { crnn_destroy(h); return rc; }
    h->kids.push_back(k);
  }
  h->num_sms = h->kids[0]->num_sms;
  if (n_devices > 1) {
    std::vector<ncclComm_t> comms(n_devices);
    if (nccl().CommInitAll(comms.data(), n_devices, ids.data()) != ncclSuccess) { crnn_destroy(h); return CRNN_ERR_CUDA; }
    for (ncclComm_t c : comms) h->comms.push_back((void*)c);
  }
  *out = h;
  return CRNN_OK;
}

int32_t crnn_device_count(const crnn_handle* h) { return h ? (h->kids.empty() ? 1 : (int32_t)h->kids.size()) : 0; }

int crnn_dataset_create(crnn_handle* h, const double* u0, const double* data, int32_t n_state, int32_t n_obs, int32_t n_save,
                        int64_t N, crnn_dataset** out) {
  if (!h || !out) return CRNN_ERR_BAD_ARG;
  *out = nullptr;
  if (N < 0 || n_state < 1 || n_obs < 1 || n_save < 1 || (N > 0 && (!u0 || !data)))
    return fail(h, CRNN_ERR_BAD_ARG, "crnn_dataset_create: bad shape or null arrays");
  auto devs = devices_of(h);
  const int nd = (int)devs.size();
  crnn_dataset* ds = new crnn_dataset();
  ds->owner = h; ds->n_state = n_state; ds->n_obs = n_obs; ds->n_save = n_save; ds->N = N;
  ds->u0.resize(nd); ds->data.resize(nd);
  const size_t ps = (size_t)n_obs * n_save;
  for (int d = 0; d <= nd; ++d) ds->lo.push_back(N * d / nd);
  for (int d = 0; d < nd; ++d) {
    crnn_handle* k = devs[d];
    const int64_t lo = ds->lo[d], n = ds->lo[d + 1] - lo;
    cudaError_t e = cudaSetDevice(k->device);
    if (e == cudaSuccess) e = ds->u0[d].reserve(std::max<size_t>(8, n * n_state * sizeof(double)));
    if (e == cudaSuccess) e = ds->data[d].reserve(std::max<size_t>(8, n * ps * sizeof(double)));
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(ds->u0[d].p, u0 + lo * n_state, n * n_state * sizeof(double), cudaMemcpyHostToDevice, k->s_h2d);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(ds->data[d].p, data + lo * ps, n * ps * sizeof(double), cudaMemcpyHostToDevice, k->s_h2d);
    if (e != cudaSuccess) {
      h->err = std::string("crnn_dataset_create: ") + cudaGetErrorString(e);
      crnn_dataset_destroy(ds);
      return CRNN_ERR_CUDA;
    }
  }
  for (int d = 0; d < nd; ++d) {
    cudaSetDevice(devs[d]->device);
    if (cudaStreamSynchronize(devs[d]->s_h2d) != cudaSuccess) { crnn_dataset_destroy(ds); return fail(h, CRNN_ERR_CUDA, "dataset upload failed"); }
  }
  *out = ds;
  return CRNN_OK;
}

void crnn_dataset_destroy(crnn_dataset* ds) {
  if (!ds) return;
  auto devs = devices_of(ds->owner);
  for (size_t d = 0; d < ds->u0.size(); ++d) {
    cudaSetDevice(devs[d]->device);
    cudaDeviceSynchronize();
    ds->u0[d].release(); ds->data[d].release();
  }
  delete ds;
}

int64_t crnn_dataset_size(const crnn_dataset* ds) { return ds ? ds->N : 0; }

int crnn_loss_grad_indexed(crnn_handle* h, const crnn_model* m, const crnn_opts* o, const double* dW_dp, int32_t np,
                           const crnn_dataset* ds, const int64_t* idx, int64_t n_idx, const int32_t* n_save_used,
                           const double* yscale, int32_t loss_kind, double* loss_sum, double* grad_sum, double* loss,
                           int32_t* n_saved, int32_t* retcode, crnn_stats* stats) {
  if (!h) return CRNN_ERR_BAD_ARG;
  if (!ds || ds->owner != h) return fail(h, CRNN_ERR_BAD_ARG, "dataset does not belong to this handle");
  if (!m || !o) return fail(h, CRNN_ERR_BAD_ARG, "null model/opts");
  if (m->n_state != ds->n_state || o->n_obs != ds->n_obs || o->n_save != ds->n_save)
    return fail(h, CRNN_ERR_BAD_ARG, "model / opts do not match the dataset's (n_state, n_obs, n_save)");
  if (!idx) n_idx = ds->N;
  if (n_idx < 0) return fail(h, CRNN_ERR_BAD_ARG, "negative n_idx");
  if (np < 0 || (np > 0 && !grad_sum)) return fail(h, CRNN_ERR_BAD_ARG, "null grad_sum");
  auto devs = devices_of(h);
  const int nd = (int)devs.size();
  // rows per device: positions (into the caller's idx order) and shard-local row numbers
  std::vector<std::vector<long long>> pos(nd), rows(nd);
  std::vector<std::vector<int32_t>> nsu(nd);
  if (idx) {
    for (int64_t k = 0; k < n_idx; ++k) {
      const int64_t r = idx[k];
      if (r < 0 || r >= ds->N) return fail(h, CRNN_ERR_BAD_ARG, "dataset row index out of range");
      const int d = (int)(std::upper_bound(ds->lo.begin(), ds->lo.end(), r) - ds->lo.begin()) - 1;
      pos[d].push_back(k); rows[d].push_back(r - ds->lo[d]);
      if (n_save_used) nsu[d].push_back(n_save_used[k]);
    }
  }
  std::vector<int64_t> cnt(nd);
  for (int d = 0; d < nd; ++d) cnt[d] = idx ? (int64_t)rows[d].size() : ds->lo[d + 1] - ds->lo[d];
  // ---- enqueue every device's share on its compute stream ----
  for (int d = 0; d < nd; ++d) {
    const int32_t* nsu_d = !n_save_used ? nullptr : (idx ? nsu[d].data() : n_save_used + ds->lo[d]);
    int rc = enqueue_share(devs[d], m, o, dW_dp, np, ds, d, idx ? rows[d].data() : nullptr, cnt[d], nsu_d, yscale, loss_kind,
                           stats != nullptr);
    if (rc) { if (devs[d] != h) h->err = "device " + std::to_string(devs[d]->device) + ": " + devs[d]->err; return rc; }
  }
  // ---- the one exchange of the path: all-reduce of [sum loss, n, grad_sum] over NVLink, on the compute streams ----
  if (nd > 1) {
    NK(nccl().GroupStart());
    for (int d = 0; d < nd; ++d)
      NK(nccl().AllReduce(devs[d]->d_result.p, devs[d]->d_result.p, (size_t)np + 2, ncclDouble, ncclSum,
                          (ncclComm_t)h->comms[d], devs[d]->s_compute));
    NK(nccl().GroupEnd());
  }
  // ---- results: [sum loss, n, grad] from device 0, per-trajectory outputs (if asked for) from every device ----
  std::vector<double> res((size_t)np + 2);
  CK(cudaSetDevice(devs[0]->device));
  CK(cudaMemcpyAsync(res.data(), devs[0]->d_result.p, res.size() * sizeof(double), cudaMemcpyDeviceToHost, devs[0]->s_compute));
  std::vector<std::vector<double>> t_loss(nd);
  std::vector<std::vector<int32_t>> t_ns(nd), t_rc(nd);
  std::vector<std::vector<crnn_stats>> t_st(nd);
  for (int d = 0; d < nd; ++d) {
    crnn_handle* k = devs[d];
    const int64_t n = cnt[d];
    if (n == 0) continue;
    CK(cudaSetDevice(k->device));
    const bool direct = !idx;  // contiguous shard: copy straight into the caller's arrays
    const int64_t off = direct ? ds->lo[d] : 0;
    auto fetch = [&](auto* dst_user, auto& tmp, const void* src, size_t elem) -> int {
      if (!dst_user) return CRNN_OK;
      void* dst = direct ? (void*)(dst_user + off) : (tmp.resize(n), (void*)tmp.data());
      CK(cudaMemcpyAsync(dst, src, n * elem, cudaMemcpyDeviceToHost, k->s_compute));
      return CRNN_OK;
    };
    int rc = fetch(loss, t_loss[d], k->d_loss.p, sizeof(double)); if (rc) return rc;
    rc = fetch(n_saved, t_ns[d], k->d_nsaved.p, sizeof(int32_t)); if (rc) return rc;
    rc = fetch(retcode, t_rc[d], k->d_ret.p, sizeof(int32_t)); if (rc) return rc;
    rc = fetch(stats, t_st[d], k->d_stats.p, sizeof(crnn_stats)); if (rc) return rc;
  }
  for (int d = 0; d < nd; ++d) {
    CK(cudaSetDevice(devs[d]->device));
    CK(cudaStreamSynchronize(devs[d]->s_compute));
  }
  if (idx)
    for (int d = 0; d < nd; ++d)
      for (size_t q = 0; q < pos[d].size(); ++q) {
        const long long k = pos[d][q];
        if (loss) loss[k] = t_loss[d][q];
        if (n_saved) n_saved[k] = t_ns[d][q];
        if (retcode) retcode[k] = t_rc[d][q];
        if (stats) stats[k] = t_st[d][q];
      }
  if (loss_sum) { loss_sum[0] = res[0]; loss_sum[1] = res[1]; }
  for (int c = 0; c < np; ++c) grad_sum[c] = res[2 + c];
  return CRNN_OK;
}

}  // extern "C"
