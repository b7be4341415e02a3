// Pieces of the quantum geometric tensor's matrix-vector product for the RBM (QGTOnTheFly,
// netket/optimizer/qgt/qgt_onthefly_logic.py:33-43:  S v = O^H ((O v - mean(O v)) / n) + shift v).
//
// O[s, :] = d log psi(sigma_s) / d p = [sigma_i tanh(theta_j) | tanh(theta_j) | sigma_i] in closed form, so
//   O v   = sum_j tanh(theta_sj) (sigma_s V + v_b)_j + sigma_s . v_a     -> the theta GEMM with (V, v_b) in place of (W, b)
//                                                                           (tcgen05 / DMMA), then the row dot below;
//   O^H w = the forces contraction with w in place of (E_loc - mean)       -> nk_forces_rbm.
// Both kernels here are HBM-bound streams over [Ns, M] arrays.
#include "kernels.cuh"

namespace nk {

template <typename T>
__global__ void __launch_bounds__(256) tanh_inplace_kernel(T *__restrict__ x, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = (T)tanh((double)x[i]);
}
template <>
__global__ void __launch_bounds__(256) tanh_inplace_kernel<float>(float *__restrict__ x, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = tanhf(x[i]);
}

// y[s] = sum_j t[s, j] g[s, j] + sum_i sigma[s, i] va[i]; one warp per sample, double accumulation
template <typename T>
__global__ void __launch_bounds__(256) jvp_dot_kernel(const T *__restrict__ t, const T *__restrict__ g, const int8_t *__restrict__ sigma,
                                                      const T *__restrict__ va, int N, int M, int64_t Ns, double *__restrict__ y,
                                                      double *__restrict__ y_sum) {
  const int lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  double total = 0.0;
  for (int64_t s = (int64_t)blockIdx.x * warps + (threadIdx.x >> 5); s < Ns; s += (int64_t)gridDim.x * warps) {
    const T *tr = t + s * M, *gr = g + s * M;
    double acc = 0.0;
    for (int j = lane; j < M; j += 32) acc = fma((double)tr[j], (double)gr[j], acc);
    if (va) {
      const int8_t *sr = sigma + s * N;
      for (int i = lane; i < N; i += 32) acc = fma((double)sr[i], (double)va[i], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) y[s] = acc;
    total += acc;
  }
  if (y_sum && lane == 0 && total != 0.0) atomicAdd(y_sum, total);  // sum over the device's samples (mean(O v) needs it)
}

static int grid_for(int64_t units, int per_block, int waves) {
  const int64_t need = (units + per_block - 1) / per_block;
  const int64_t cap = (int64_t)num_sms() * waves;
  return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

int rbm_tanh_inplace(cudaStream_t stream, void *x, int32_t dtype, int64_t n) {
  if (n == 0) return NK_OK;
  if (dtype == NK_F32)
    tanh_inplace_kernel<float><<<grid_for(n, 256, 16), 256, 0, stream>>>((float *)x, n);
  else
    tanh_inplace_kernel<double><<<grid_for(n, 256, 16), 256, 0, stream>>>((double *)x, n);
  NK_LAUNCH_OK();
  return NK_OK;
}

int rbm_jvp_dot(cudaStream_t stream, const nk_rbm_t &v, const int8_t *sigma, int64_t Ns, const void *t, const void *g, double *y,
                double *y_sum) {
  if (y_sum) NK_CUDA_OK(cudaMemsetAsync(y_sum, 0, sizeof(double), stream));
  if (Ns == 0) return NK_OK;
  const int grid = grid_for(Ns, 8, 16);
  if (v.dtype == NK_F32)
    jvp_dot_kernel<float><<<grid, 256, 0, stream>>>((const float *)t, (const float *)g, sigma, (const float *)v.a, v.N, v.M, Ns, y, y_sum);
  else
    jvp_dot_kernel<double><<<grid, 256, 0, stream>>>((const double *)t, (const double *)g, sigma, (const double *)v.a, v.N, v.M, Ns, y, y_sum);
  NK_LAUNCH_OK();
  return NK_OK;
}

}  // namespace nk
