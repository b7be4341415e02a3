// Stand-alone RBM.apply and local-energy kernels (generic theta-form path).
//
//  * rbm_logpsi_kernel   replaces netket/models/rbm.py:57-81 (+ log_cosh, netket/nn/activation.py:78-84)
//  * eloc_kernel         replaces local_value_kernel_jax (netket/vqs/mc/kernels.py:62-71) for Ising and
//                        2-site LocalOperator: connected configurations are never written to HBM.
#include "kernels.cuh"
#include "rbm_warp.cuh"

namespace nk {

template <typename T>
__global__ void __launch_bounds__(256) rbm_logpsi_kernel(const __grid_constant__ nk_rbm_t rbm, const int8_t *__restrict__ sigma,
                                                         int64_t B, T *__restrict__ out, T *__restrict__ theta_out, int n_pad) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const RbmView<T> r = make_view<T>(rbm);
  int8_t *sig = reinterpret_cast<int8_t *>(smem_raw) + (size_t)warp * n_pad;
  for (int64_t s = (int64_t)blockIdx.x * warps + warp; s < B; s += (int64_t)gridDim.x * warps) {
    for (int i = lane; i < r.N; i += 32) sig[i] = sigma[s * r.N + i];
    __syncwarp();
    T lp;
    if (theta_out != nullptr)
      lp = warp_theta_init<T, true>(r, sig, theta_out + s * r.M, lane);
    else
      lp = warp_theta_init<T, false>(r, sig, (T *)nullptr, lane);
    if (lane == 0) out[s] = lp;
    __syncwarp();
  }
}

struct ElocArgs {
  nk_rbm_t rbm;
  const int8_t *sigma;
  int64_t B;
  int32_t kind;  // 1 ising, 2 localop
  nk_ising_t ising;
  nk_localop_t localop;
  void *eloc_out;
  int32_t eloc_dtype;
  int32_t n_pad;
  const int *run_if_flag;  // run iff NULL or *run_if_flag != 0 (hand-over from the product-form kernel)
};

template <typename T>
__global__ void __launch_bounds__(256) eloc_generic_kernel(const __grid_constant__ ElocArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const RbmView<T> r = make_view<T>(p.rbm);
  T *theta = reinterpret_cast<T *>(smem_raw) + (size_t)warp * r.M;
  int8_t *sig = reinterpret_cast<int8_t *>(smem_raw + (size_t)warps * r.M * sizeof(T)) + (size_t)warp * p.n_pad;
  if (p.run_if_flag != nullptr && *p.run_if_flag == 0) return;
  for (int64_t s = (int64_t)blockIdx.x * warps + warp; s < p.B; s += (int64_t)gridDim.x * warps) {
    for (int i = lane; i < r.N; i += 32) sig[i] = p.sigma[s * r.N + i];
    __syncwarp();
    (void)warp_theta_init<T, true>(r, sig, theta, lane);
    __syncwarp();
    T e;
    if (p.kind == 1)
      e = warp_eloc_ising<T>(r, theta, sig, p.ising.edges, p.ising.n_edges, (T)p.ising.h, (T)p.ising.J, lane);
    else
      e = warp_eloc_localop<T>(r, theta, sig, p.localop, lane);
    if (lane == 0) store_as<T>(p.eloc_out, s, e, p.eloc_dtype);
    __syncwarp();
  }
}

template <typename K>
static int pick_grid(K kernel, int threads, size_t smem, int64_t units, int units_per_cta, int *grid) {
  int occ = 1;
  NK_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
  if (occ < 1) occ = 1;
  int64_t need = (units + units_per_cta - 1) / units_per_cta;
  int64_t cap = (int64_t)num_sms() * occ;
  *grid = (int)(need < cap ? need : cap);
  if (*grid < 1) *grid = 1;
  return NK_OK;
}

template <typename T>
static int launch_logpsi(cudaStream_t stream, const nk_rbm_t &rbm, const int8_t *sigma, int64_t B, void *out, void *theta_out) {
  const int n_pad = (rbm.N + 15) & ~15;
  const int warps = 8;
  const size_t smem = (size_t)warps * n_pad;
  int grid;
  int rc = pick_grid(rbm_logpsi_kernel<T>, warps * 32, smem, B, warps, &grid);
  if (rc) return rc;
  rbm_logpsi_kernel<T><<<grid, warps * 32, smem, stream>>>(rbm, sigma, B, (T *)out, (T *)theta_out, n_pad);
  NK_LAUNCH_OK();
  return NK_OK;
}

int rbm_logpsi(cudaStream_t stream, const nk_rbm_t &rbm, const int8_t *sigma, int64_t B, void *out, void *theta_out) {
  if (B == 0) return NK_OK;
  return rbm.dtype == NK_F32 ? launch_logpsi<float>(stream, rbm, sigma, B, out, theta_out)
                             : launch_logpsi<double>(stream, rbm, sigma, B, out, theta_out);
}

template <typename T>
static int launch_eloc(cudaStream_t stream, ElocArgs a) {
  const int N = a.rbm.N, M = a.rbm.M;
  a.n_pad = (N + 15) & ~15;
  const size_t per_warp = (size_t)M * sizeof(T) + a.n_pad;
  const size_t budget = 200 * 1024;
  int warps = 8;
  while (warps > 1 && warps * per_warp > budget) warps >>= 1;
  if (warps * per_warp > budget) {
    set_error("nk_eloc: M=%d too large for the generic path", M);
    return NK_EUNSUPPORTED;
  }
  const size_t smem = warps * per_warp;
  NK_CUDA_OK(cudaFuncSetAttribute(eloc_generic_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
  int grid;
  int rc = pick_grid(eloc_generic_kernel<T>, warps * 32, smem, a.B, warps, &grid);
  if (rc) return rc;
  eloc_generic_kernel<T><<<grid, warps * 32, smem, stream>>>(a);
  NK_LAUNCH_OK();
  return NK_OK;
}

int eloc_generic(cudaStream_t stream, const nk_rbm_t &rbm, const nk_ising_t *ising, const nk_localop_t *localop,
                 const int8_t *sigma, int64_t B, void *eloc_out, int32_t eloc_dtype, const int *run_if_flag) {
  if (B == 0) return NK_OK;
  ElocArgs a{};
  a.rbm = rbm;
  a.sigma = sigma;
  a.B = B;
  a.kind = ising != nullptr ? 1 : 2;
  if (ising) a.ising = *ising;
  if (localop) a.localop = *localop;
  a.eloc_out = eloc_out;
  a.eloc_dtype = eloc_dtype;
  a.run_if_flag = run_if_flag;
  return rbm.dtype == NK_F32 ? launch_eloc<float>(stream, a) : launch_eloc<double>(stream, a);
}

}  // namespace nk
