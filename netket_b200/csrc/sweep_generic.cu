// Generic (theta-form) Metropolis sweep kernel: any N, M, fp32/fp64, LocalRule / ExchangeRule,
// optionally fused with the local energy of every recorded sample.
//
// Replaces MetropolisSampler._reset + _sample_chain (netket/sampler/metropolis.py:382-505):
// the reference runs `chain_length * sweep_size` iterations of {rule.transition, full forward pass,
// uniform, exp, select} as ~10 XLA kernels each; here one warp owns one chain for the whole call,
// keeps theta = W^T sigma + b in shared memory and applies rank-1 / rank-2 updates on accept.
#include "rbm_warp.cuh"
#include "kernels.cuh"

namespace nk {


// number of clusters with unequal spins (netket/sampler/rules/exchange.py:208-218)
__device__ __forceinline__ int warp_count_hoppable(const int8_t *sig, const int32_t *clusters, int C, int lane) {
  int cnt = 0;
  for (int c = lane; c < C; c += 32) cnt += (sig[clusters[2 * c]] != sig[clusters[2 * c + 1]]) ? 1 : 0;
  return warp_sum(cnt);
}

// k-th (0-based) hoppable cluster in cluster order
__device__ __forceinline__ int warp_select_hoppable(const int8_t *sig, const int32_t *clusters, int C, int k, int lane) {
  for (int base = 0; base < C; base += 32) {
    int c = base + lane;
    bool h = c < C && (sig[clusters[2 * c]] != sig[clusters[2 * c + 1]]);
    unsigned m = __ballot_sync(0xffffffffu, h);
    int pc = __popc(m);
    if (k < pc) return base + (int)__fns(m, 0, k + 1);
    k -= pc;
  }
  return -1;
}

// ExchangeRule(probabilities=p) (rules/exchange.py:155-182): total weight sum_c hoppable_c p_c ...
__device__ __forceinline__ double warp_weight_hoppable(const int8_t *sig, const int32_t *clusters, const double *prob, int C, int lane) {
  double w = 0.0;
  for (int c = lane; c < C; c += 32) w += (sig[clusters[2 * c]] != sig[clusters[2 * c + 1]]) ? prob[c] : 0.0;
  return warp_sum(w);
}

// ... and the first cluster (in cluster order) whose running weight reaches r: jax.random.choice's inverse-CDF pick,
// searchsorted(cumsum(hoppable * p), r)
__device__ __forceinline__ int warp_select_weighted(const int8_t *sig, const int32_t *clusters, const double *prob, int C, double r, int lane) {
  double run = 0.0;
  for (int base = 0; base < C; base += 32) {
    const int c = base + lane;
    double w = (c < C && sig[clusters[2 * c]] != sig[clusters[2 * c + 1]]) ? prob[c] : 0.0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += t;
    }
    const unsigned m = __ballot_sync(0xffffffffu, c < C && run + w >= r);
    if (m != 0u) return base + __ffs(m) - 1;
    run += __shfl_sync(0xffffffffu, w, 31);
  }
  return C - 1;
}

template <typename T>
__global__ void __launch_bounds__(256) sweep_generic_kernel(const __grid_constant__ SweepKernelArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps = blockDim.x >> 5;
  const RbmView<T> r = make_view<T>(p.rbm);
  const int N = r.N, M = r.M;
  T *theta = reinterpret_cast<T *>(smem_raw) + (size_t)warp * M;
  int8_t *sig = reinterpret_cast<int8_t *>(smem_raw + (size_t)warps * M * sizeof(T)) + (size_t)warp * p.n_pad;
  const T pw = (T)p.machine_pow;
  const int64_t T_total = (int64_t)(p.n_discard + p.chain_length) * p.sweep_size;
  if (p.run_if_flag != nullptr && *p.run_if_flag == 0) return;  // the fast kernel did the work

  for (int64_t chain = (int64_t)blockIdx.x * warps + warp; chain < p.B; chain += (int64_t)gridDim.x * warps) {
    for (int i = lane; i < N; i += 32) sig[i] = p.sigma[chain * N + i];
    __syncwarp();
    // _reset: log_prob = machine_pow * logpsi(sigma)  (metropolis.py:399-403); theta is rebuilt here
    T logpsi = warp_theta_init<T, true>(r, sig, theta, lane);
    __syncwarp();
    int64_t n_acc = 0;
    const uint64_t gchain = p.chain_offset + (uint64_t)chain;

    for (int64_t tt = 0; tt < T_total; tt += 32) {
      uint32_t w0_l = 0;
      T u_l = T(0);
      if (tt + lane < T_total) {
        if (p.stream_w0 != nullptr) {
          w0_l = p.stream_w0[(tt + lane) * p.B + chain];
          u_l = reinterpret_cast<const T *>(p.stream_u)[(tt + lane) * p.B + chain];
        } else {
          uint4 w = philox_words(p.seed, p.t0 + (uint64_t)(tt + lane), gchain, STREAM_STEP);
          w0_l = w.x;
          u_l = uniform_from_words<T>(w);
        }
      }
      const int nb = (int)min((int64_t)32, T_total - tt);
      for (int k = 0; k < nb; ++k) {
        const uint32_t w0 = __shfl_sync(0xffffffffu, w0_l, k);
        const T u = __shfl_sync(0xffffffffu, u_l, k);
        if (p.rule == NK_RULE_LOCAL) {
          // LocalRule.transition (rules/local.py:40-49): uniform site, deterministic flip
          const int i = (int)__umulhi(w0, (uint32_t)N);
          const T s = (T)sig[i];
          const T d = T(-2) * s;
          T delta = warp_sum(lane_delta_one(r, theta, i, d, lane));
          if (r.a != nullptr) delta = Math<T>::fma(d, r.a[i], delta);
          // accept = u < exp(logp' - logp), logp = machine_pow * Re logpsi  (metropolis.py:441-450)
          const bool accept = u < Math<T>::exp(pw * delta);
          if (accept) {
            warp_theta_update_one(r, theta, i, d, lane);
            if (lane == 0) sig[i] = (int8_t)(-(int)sig[i]);
            logpsi += delta;
            ++n_acc;
          }
          __syncwarp();
        } else {
          // ExchangeRule.transition (rules/exchange.py:143-184)
          const int C = p.n_clusters;
          const bool weighted = p.cluster_probs != nullptr;
          const int n_hop = weighted ? 0 : warp_count_hoppable(sig, p.clusters, C, lane);
          const double w_hop = weighted ? warp_weight_hoppable(sig, p.clusters, p.cluster_probs, C, lane) : 0.0;
          if (weighted ? w_hop > 0.0 : n_hop > 0) {
            int c;
            if (weighted) {
              c = warp_select_weighted(sig, p.clusters, p.cluster_probs, C, w_hop * (((double)w0 + 0.5) * 2.3283064365386963e-10), lane);
            } else {
              const int kth = (int)__umulhi(w0, (uint32_t)n_hop);
              c = warp_select_hoppable(sig, p.clusters, C, kth, lane);
            }
            const int si = p.clusters[2 * c], sj = p.clusters[2 * c + 1];
            const T vi = (T)sig[si], vj = (T)sig[sj];
            const T di = vj - vi, dj = vi - vj;  // sigma' - sigma
            T delta = warp_sum(lane_delta_two(r, theta, si, di, sj, dj, lane));
            if (r.a != nullptr) delta += di * r.a[si] + dj * r.a[sj];
            // correction log n_hop(sigma) - log n_hop(sigma')  (:177-182): swap, recount, swap back on reject
            __syncwarp();
            if (lane == 0) {
              int8_t tmp = sig[si];
              sig[si] = sig[sj];
              sig[sj] = tmp;
            }
            __syncwarp();
            T corr;
            if (weighted) {
              const double w_hop_p = warp_weight_hoppable(sig, p.clusters, p.cluster_probs, C, lane);
              corr = (T)(log(w_hop) - log(w_hop_p));
            } else {
              const int n_hop_p = warp_count_hoppable(sig, p.clusters, C, lane);
              corr = Math<T>::log((T)n_hop) - Math<T>::log((T)n_hop_p);
            }
            const bool accept = u < Math<T>::exp(pw * delta + corr);
            if (accept) {
              warp_theta_update_two(r, theta, si, di, sj, dj, lane);
              logpsi += delta;
              ++n_acc;
            } else {
              __syncwarp();
              if (lane == 0) {
                int8_t tmp = sig[si];
                sig[si] = sig[sj];
                sig[sj] = tmp;
              }
            }
            __syncwarp();
          }
        }
        // end of a sweep: record (metropolis.py:492-499: one sample per sweep, (n_chains, chain_length, N))
        const int64_t step = tt + k + 1;
        if (step % p.sweep_size == 0) {
          const int64_t sw = step / p.sweep_size - 1 - p.n_discard;
          if (sw >= 0) {
            const int64_t o = chain * p.chain_length + sw;
            if (p.samples_out != nullptr)
              for (int i = lane; i < N; i += 32) p.samples_out[o * N + i] = sig[i];
            if (p.logp_out != nullptr && lane == 0) reinterpret_cast<T *>(p.logp_out)[o] = pw * logpsi;
            if (p.tanh_out != nullptr)
              for (int j = lane; j < M; j += 32) reinterpret_cast<T *>(p.tanh_out)[o * M + j] = tanh(theta[j]);
            if (p.eloc_kind == 1) {
              T e = warp_eloc_ising<T>(r, theta, sig, p.ising.edges, p.ising.n_edges, (T)p.ising.h, (T)p.ising.J, lane);
              if (lane == 0) store_as<T>(p.eloc_out, o, e, p.eloc_dtype);
            } else if (p.eloc_kind == 2) {
              T e = warp_eloc_localop<T>(r, theta, sig, p.localop, lane);
              if (lane == 0) store_as<T>(p.eloc_out, o, e, p.eloc_dtype);
            }
          }
        }
      }
    }
    __syncwarp();
    for (int i = lane; i < N; i += 32) p.sigma[chain * N + i] = sig[i];
    if (lane == 0) {
      reinterpret_cast<T *>(p.log_prob)[chain] = pw * logpsi;
      p.n_accepted[chain] += n_acc;
    }
    __syncwarp();
  }
}

template <typename T>
static int launch_generic(cudaStream_t stream, const SweepKernelArgs &a) {
  const int N = a.rbm.N, M = a.rbm.M;
  const int n_pad = (N + 15) & ~15;
  const size_t per_warp = (size_t)M * sizeof(T) + n_pad;
  const size_t budget = 200 * 1024;
  int warps = 8;
  while (warps > 1 && warps * per_warp > budget) warps >>= 1;
  if (warps * per_warp > budget) {
    set_error("nk_sweep: M=%d too large for the generic path (theta does not fit shared memory)", M);
    return NK_EUNSUPPORTED;
  }
  SweepKernelArgs args = a;
  args.n_pad = n_pad;
  const size_t smem = warps * per_warp;
  NK_CUDA_OK(cudaFuncSetAttribute(sweep_generic_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
  int occ = 1;
  NK_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sweep_generic_kernel<T>, warps * 32, smem));
  if (occ < 1) occ = 1;
  const int64_t need = (a.B + warps - 1) / warps;
  const int64_t cap = (int64_t)num_sms() * occ;
  const int grid = (int)(need < cap ? need : cap);
  sweep_generic_kernel<T><<<grid, warps * 32, smem, stream>>>(args);
  NK_LAUNCH_OK();
  return NK_OK;
}

int sweep_generic(cudaStream_t stream, const SweepKernelArgs &a) {
  return a.rbm.dtype == NK_F32 ? launch_generic<float>(stream, a) : launch_generic<double>(stream, a);
}

}  // namespace nk
