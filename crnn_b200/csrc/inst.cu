// inst.cu — explicit instantiation of the engine for ONE (n_species, n_reac, rhs_kind);
// compiled once per configuration listed in the Makefile (keep in sync with CRNN_FOR_EACH_CFG).
#include "crnn_host.cuh"

namespace crnn_host {
using CfgT = crnn::Cfg<CRNN_NS, CRNN_NR, CRNN_KIND>;
template int solve_impl<CfgT>(crnn_handle*, const crnn_model*, const crnn_opts*, const HostIO&, int64_t);
template int loss_grad_impl<CfgT>(crnn_handle*, const crnn_model*, const crnn_opts*, const double*, int,
                                  const double*, int, const HostIO&, int64_t, double*, const AutoHook*);
template int train_impl<CfgT>(crnn_handle*, const crnn_model*, const crnn_opts*, const crnn_train_opts*, const crnn_dataset*,
                              const int64_t*, int64_t, const double*, int32_t, double*, double*, double*, double*);
}  // namespace crnn_host
