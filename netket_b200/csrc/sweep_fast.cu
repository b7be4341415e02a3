// Product-form fast path (placeholder until the kernel lands): reports "unsupported" so that nk_sweep and
// nk_eloc_ising_rbm use the generic theta-form kernels.
#include "kernels.cuh"

namespace nk {
bool sweep_fast_supported(const SweepKernelArgs &) { return false; }
int sweep_fast(cudaStream_t, const SweepKernelArgs &) {
  set_error("sweep_fast: not built");
  return NK_EUNSUPPORTED;
}
bool eloc_fast_supported(const nk_rbm_t &) { return false; }
int eloc_fast_ising(cudaStream_t, const nk_rbm_t &, const nk_ising_t &, const int8_t *, int64_t, void *, int32_t) {
  set_error("eloc_fast: not built");
  return NK_EUNSUPPORTED;
}
}  // namespace nk
