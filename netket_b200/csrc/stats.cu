// MC statistics partial sums (statistics(), netket/stats/mc_stats_old.py:52-196).
//
// Every quantity of the block statistics is a function of sums that add across chains, so a device only
// produces NK_STATS_NPARTIAL doubles; the caller all-reduces them (NCCL, ~64 bytes) and nk_stats_finalize
// does the scalar arithmetic.  Two phases keep jnp.var's accuracy: phase 0 gives the mean, phase 1 the
// second moments shifted by that mean (var(y) = sum((y-mu)^2)/n - (mean(y)-mu)^2 holds for any shift mu).
#include <math.h>

#include "kernels.cuh"

namespace nk {

__device__ __forceinline__ double block_sum(double v, double *sh) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (warp == 0) {
    r = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
    r = warp_sum(r);
  }
  return r;  // valid in warp 0
}

template <typename T>
__global__ void __launch_bounds__(256) stats_partial_kernel(const T *__restrict__ data, int64_t n_chains, int64_t L, int phase,
                                                            double mu, double *__restrict__ out, const int *__restrict__ run_if) {
  __shared__ double sh[8];
  if (run_if != nullptr && *run_if == 0) return;  // the sweep kernel in front of this one reduced its energies itself
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  double acc[NK_STATS_NPARTIAL];
#pragma unroll
  for (int q = 0; q < NK_STATS_NPARTIAL; ++q) acc[q] = 0.0;
  const int64_t l_block = (L / 32) > 1 ? (L / 32) : 1;  // l_block = max(1, L // batch_size)   (:100)
  const int64_t n_b = L / l_block;                       // blocks per chain                      (:28-35)
  const int64_t half = L / 2;                            // split-R_hat halves                    (:165-185)
  for (int64_t c = (int64_t)blockIdx.x * warps + warp; c < n_chains; c += (int64_t)gridDim.x * warps) {
    const T *row = data + c * L;
    if (phase == 0) {
      double s = 0.0;
      for (int64_t i = lane; i < L; i += 32) s += (double)row[i];
      acc[0] += s;  // lane partial; reduced below
    } else {
      // (x - mu) sums: total, squared, per-block, per-half
      double s1 = 0.0, s2 = 0.0;
      for (int64_t i = lane; i < L; i += 32) {
        const double d = (double)row[i] - mu;
        s1 += d;
        s2 += d * d;
      }
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      double bl1 = 0.0, bl2 = 0.0;
      for (int64_t bidx = 0; bidx < n_b; ++bidx) {
        double s = 0.0;
        for (int64_t i = lane; i < l_block; i += 32) s += (double)row[bidx * l_block + i] - mu;
        s = warp_sum(s) / (double)l_block;
        bl1 += s;
        bl2 += s * s;
      }
      double h1 = 0.0, h2 = 0.0;
      for (int hh = 0; hh < 2 && half > 0; ++hh) {
        double s = 0.0;
        for (int64_t i = lane; i < half; i += 32) s += (double)row[hh * half + i] - mu;
        s = warp_sum(s) / (double)half;
        h1 += s;
        h2 += s * s;
      }
      if (lane == 0) {
        const double m = s1 / (double)L;
        acc[0] += s2;
        acc[1] += m;
        acc[2] += m * m;
        acc[3] += bl1;
        acc[4] += bl2;
        acc[5] += h1;
        acc[6] += h2;
        acc[7] += s1;
      }
    }
  }
#pragma unroll
  for (int q = 0; q < NK_STATS_NPARTIAL; ++q) {
    if (phase == 0 && q > 0) break;
    double r = block_sum(acc[q], sh);
    if (threadIdx.x == 0) atomicAdd(out + q, r);
  }
}

// run_if != NULL: accumulate into `out` (not zeroed here) iff *run_if != 0 when the kernel starts (in-stream hand-over).
int stats_partial(cudaStream_t stream, const void *data, int32_t dtype, int64_t n_chains, int64_t L, int32_t phase, double shift,
                  double *out, const int *run_if) {
  if (run_if == nullptr) NK_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(double) * NK_STATS_NPARTIAL, stream));
  if (n_chains == 0 || L == 0) return NK_OK;
  const int warps = 8;
  const int64_t need = (n_chains + warps - 1) / warps;
  const int64_t cap = (int64_t)num_sms() * 4;
  const int grid = (int)(need < cap ? need : cap);
  if (dtype == NK_F32)
    stats_partial_kernel<float><<<grid, warps * 32, 0, stream>>>((const float *)data, n_chains, L, phase, shift, out, run_if);
  else
    stats_partial_kernel<double><<<grid, warps * 32, 0, stream>>>((const double *)data, n_chains, L, phase, shift, out, run_if);
  NK_LAUNCH_OK();
  return NK_OK;
}

// ---- integrated autocorrelation time per chain, Sokal's automatic window (the opt-in FFT variant of `statistics`:
// netket/stats/mc_stats.py:303-331 with netket/stats/_autocorr.py:40-86).  The reference evaluates the autocorrelation function
// acf[k] = sum_t d_t d_{t+k}, d = x - mean(x), with a zero-padded FFT; the same sums are taken directly here, lag by lag, and
// the scan stops at the window: tau(M) = 2 sum_{k<=M} acf[k]/acf[0] - 1 at the first M with M >= c tau(M) (`auto_window`:
// argmin(M < c tau) - 0 if every M qualifies, the last M if none does).  One warp per chain; out[0] += sum of the chains'
// tau, out[1] = max tau (as ordered bits, see tau_decode), out[2] += number of chains whose tau is NaN.
__device__ __forceinline__ unsigned long long tau_encode(double v) {  // order-preserving map double -> u64
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

template <typename T>
__global__ void __launch_bounds__(256) stats_tau_kernel(const T *__restrict__ data, int64_t n_chains, int64_t L, double c, double *__restrict__ out) {
  extern __shared__ double tsm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  double *d = tsm + (size_t)warp * L;
  for (int64_t ch = (int64_t)blockIdx.x * warps + warp; ch < n_chains; ch += (int64_t)gridDim.x * warps) {
    const T *row = data + ch * L;
    double s = 0.0;
    for (int64_t i = lane; i < L; i += 32) s += (double)row[i];
    const double mean = warp_sum(s) / (double)L;
    for (int64_t i = lane; i < L; i += 32) d[i] = (double)row[i] - mean;
    __syncwarp();
    double acf0 = 0.0, cum = 0.0, tau = 0.0, tau0 = 0.0;
    bool any_true = false, found = false;
    int64_t k = 0;
    for (; k < L; ++k) {
      double a = 0.0;
      for (int64_t t = lane; t + k < L; t += 32) a += d[t] * d[t + k];
      a = warp_sum(a);
      if (k == 0) acf0 = a;
      cum += a / acf0;
      tau = 2.0 * cum - 1.0;
      if (k == 0) tau0 = tau;
      if ((double)k < c * tau) {
        any_true = true;
      } else {  // first False entry of m: the window (if any entry is True at all)
        found = true;
        break;
      }
    }
    double res;
    if (found) {
      if (!any_true) {
        // no M with M < c tau: the reference takes the last M; finish the cumulative sum
        for (++k; k < L; ++k) {
          double a = 0.0;
          for (int64_t t = lane; t + k < L; t += 32) a += d[t] * d[t + k];
          cum += warp_sum(a) / acf0;
        }
        res = 2.0 * cum - 1.0;
      } else {
        res = tau;
      }
    } else {
      res = tau0;  // every M qualifies: argmin of an all-True mask is 0
    }
    if (lane == 0) {
      if (res != res) {
        atomicAdd(out + 2, 1.0);
      } else {
        atomicAdd(out + 0, res);
        atomicMax(reinterpret_cast<unsigned long long *>(out + 1), tau_encode(res));
      }
    }
    __syncwarp();
  }
}

int stats_tau(cudaStream_t stream, const void *data, int32_t dtype, int64_t n_chains, int64_t L, double c, double *out) {
  NK_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(double) * 3, stream));  // out[1] = 0 bits: below the encoding of every double
  if (n_chains == 0 || L == 0) return NK_OK;
  int warps = 8;
  while (warps > 1 && (size_t)warps * L * 8 > 200 * 1024) warps >>= 1;
  const size_t smem = (size_t)warps * L * 8;
  if (smem > 200 * 1024) {
    set_error("nk_stats_tau: chains of %lld samples do not fit shared memory (max 25600)", (long long)L);
    return NK_EUNSUPPORTED;
  }
  const int64_t need = (n_chains + warps - 1) / warps;
  const int64_t cap = (int64_t)num_sms() * 4;
  const int grid = (int)(need < cap ? need : cap);
  if (dtype == NK_F32) {
    NK_CUDA_OK(cudaFuncSetAttribute(stats_tau_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    stats_tau_kernel<float><<<grid, warps * 32, smem, stream>>>((const float *)data, n_chains, L, c, out);
  } else {
    NK_CUDA_OK(cudaFuncSetAttribute(stats_tau_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    stats_tau_kernel<double><<<grid, warps * 32, smem, stream>>>((const double *)data, n_chains, L, c, out);
  }
  NK_LAUNCH_OK();
  return NK_OK;
}

// Scalar arithmetic of _statistics (mc_stats_old.py:87-196) from globally reduced sums.
int stats_finalize(const double *p, double mean, int64_t n_chains, int64_t L, double *out) {
  const double nan = NAN;
  const double ts = (double)n_chains * (double)L;
  const double dm = p[7] / ts;  // residual mean shift (0 up to rounding when shift == mean)
  const double variance = p[0] / ts - dm * dm;
  const double nb = (double)n_chains;
  const double batch_var = p[2] / nb - (p[1] / nb) * (p[1] / nb);
  const int64_t l_block = (L / 32) > 1 ? (L / 32) : 1;
  const int64_t n_blocks = n_chains * (L / l_block);
  double block_var = nan;
  if (n_blocks > 0) block_var = p[4] / (double)n_blocks - (p[3] / (double)n_blocks) * (p[3] / (double)n_blocks);
  const double tau_batch = ((ts / nb) * batch_var / variance - 1.0) * 0.5;
  const double tau_block = n_blocks > 0 ? ((ts / (double)n_blocks) * block_var / variance - 1.0) * 0.5 : nan;
  const bool batch_good = (tau_batch < 6.0 * (double)L) && (n_chains >= 32);
  const bool block_good = (tau_block < 6.0 * (double)l_block) && (n_blocks >= 32);
  double err = nan, tau = nan;
  if (batch_good) {
    err = sqrt(batch_var / nb);
    tau = tau_batch > 0.0 ? tau_batch : 0.0;
  } else if (block_good) {
    err = sqrt(block_var / (double)n_blocks);
    tau = tau_block > 0.0 ? tau_block : 0.0;
  }
  double rhat = nan;
  if (n_chains > 1) {
    const double nh = 2.0 * nb;
    const double hv = (L / 2) > 0 ? p[6] / nh - (p[5] / nh) * (p[5] / nh) : nan;
    rhat = sqrt(((double)L - 1.0) / (double)L + hv / variance);
  }
  out[0] = mean + dm;  // dm: rounding noise when the shift was the mean itself, the whole correction for a one-pass shift
  out[1] = err;
  out[2] = variance;
  out[3] = tau;
  out[4] = rhat;
  return NK_OK;
}

}  // namespace nk
