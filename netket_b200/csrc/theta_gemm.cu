// theta = sigma W + b.  Interim CUDA-core version (the tensor-core kernel replaces this file's body).
#include "kernels.cuh"

namespace nk {
int rbm_logpsi(cudaStream_t stream, const nk_rbm_t &rbm, const int8_t *sigma, int64_t B, void *out, void *theta_out);

int64_t theta_gemm_workspace_bytes(const nk_rbm_t &rbm, int64_t B) { return B * (rbm.dtype == NK_F32 ? 4 : 8); }

int theta_gemm(cudaStream_t stream, const nk_rbm_t &rbm, const int8_t *sigma, int64_t B, void *theta_out, void *workspace) {
  if (B == 0) return NK_OK;
  if (workspace == nullptr) {
    set_error("nk_theta_gemm: workspace is NULL");
    return NK_EINVAL;
  }
  return rbm_logpsi(stream, rbm, sigma, B, workspace, theta_out);
}
}  // namespace nk
