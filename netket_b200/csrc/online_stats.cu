// Streaming MC statistics: per-chain Welford state + online autocovariance at lags 0..max_lag
// (OnlineStats, netket/_src/stats/online_stats/kernels.py:25-190 and accumulator.py:226-447).
//
// update:   one warp per chain.  The chain's stored samples and the new batch form one series z = [buffer | batch] in
//           shared memory; lane l owns the lags k = l, l+32, ...: the pair (t, t-k) counts when t lies in the batch and t-k
//           is a stored sample (the reference's "within-batch" and "cross-batch" windows are the two halves of that set).
//           Long batches go through in chunks of OS_CHUNK samples, the buffer rolling between chunks, which visits the
//           same pairs.  Reads of z[t-k] are consecutive across lanes, z[t] is a broadcast.
// summary:  everything the derived quantities need is a sum over chains, so a device produces 3 doubles (phase 0) and
//           4 + max_lag+1 doubles (phase 1); only those cross GPUs, nk_online_stats_finalize does the host arithmetic
//           (Geyer initial positive / monotone sequence on 65 numbers).
#include <math.h>

#include "kernels.cuh"

namespace nk {

constexpr int OS_CHUNK = 64;
constexpr int OS_WARPS = 8;

struct OsPtrs {
  const double *count, *mean, *M2, *cross, *m1, *m2, *pairs, *buf;
};
struct OsOut {
  double *count, *mean, *M2, *cross, *m1, *m2, *pairs, *buf;
};

template <typename T>
__global__ void __launch_bounds__(OS_WARPS * 32) online_update_kernel(OsPtrs in, OsOut out, const T *__restrict__ data, int64_t n_chains,
                                                                      int64_t n, int L, int buf_len, double decay, int warps) {
  extern __shared__ double os_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= warps) return;
  double *z = os_smem + (size_t)warp * (L + OS_CHUNK);
  const int nk = L > 0 ? L + 1 : 0;
  const bool single = n <= OS_CHUNK;  // the usual case (n = chain_length of one sampling call): one pass, sums stay in registers
  const int rounds = (nk + 31) / 32;
  for (int64_t c = (int64_t)blockIdx.x * warps + warp; c < n_chains; c += (int64_t)gridDim.x * warps) {
    const T *x = data + c * n;
    __syncwarp();
    if (single)
      for (int i = lane; i < (int)n; i += 32) z[L + i] = (double)x[i];
    for (int i = lane; i < L; i += 32) z[i] = in.buf[c * L + i];
    __syncwarp();
    // ---- parallel Welford merge of the batch (kernels.py:144-167)
    double s = 0.0;
    for (int64_t i = lane; i < n; i += 32) s += single ? z[L + i] : (double)x[i];
    const double bm = warp_sum(s) / (double)n;
    double q = 0.0;
    for (int64_t i = lane; i < n; i += 32) {
      const double d = (single ? z[L + i] : (double)x[i]) - bm;
      q += d * d;
    }
    q = warp_sum(q);
    if (lane == 0) {
      const double cnt = in.count[c] * decay, M2 = in.M2[c] * decay, mu = in.mean[c];
      const double tot = cnt + (double)n, safe = tot > 0.0 ? tot : 1.0, delta = bm - mu;
      out.mean[c] = mu + delta * ((double)n / safe);
      out.M2[c] = M2 + q + delta * delta * (cnt * (double)n / safe);
      out.count[c] = tot;
    }
    if (L == 0) continue;
    // ---- autocovariance sums (kernels.py:25-113)
    int first = L - buf_len;  // index in z of the oldest stored sample
    if (single) {
      const int cn = (int)n;
      for (int r = 0; r < rounds; ++r) {
        const int k = lane + 32 * r;
        if (k < nk) {
          const double o_cross = in.cross[c * nk + k], o_m1 = in.m1[c * nk + k], o_m2 = in.m2[c * nk + k], o_np = in.pairs[c * nk + k];
          double sc = 0.0, sl = 0.0, su = 0.0;
          int t0 = first + k - L;  // first batch position whose partner t-k is a stored sample
          t0 = t0 > 0 ? t0 : 0;
          for (int t = t0; t < cn; ++t) {
            const double cur = z[L + t], lag = z[L + t - k];
            sc = fma(cur, lag, sc);
            sl += lag;
            su += cur;
          }
          const int np = cn - t0 > 0 ? cn - t0 : 0;
          out.cross[c * nk + k] = fma(o_cross, decay, sc);
          out.m1[c * nk + k] = fma(o_m1, decay, sl);
          out.m2[c * nk + k] = fma(o_m2, decay, su);
          out.pairs[c * nk + k] = fma(o_np, decay, (double)np);
        }
      }
      for (int i = lane; i < L; i += 32) out.buf[c * L + i] = z[i + cn];  // the last L samples of z, right-aligned
      continue;
    }
    // long batches: chunks of OS_CHUNK samples, the buffer rolling between them; the sums of lag k = lane + 32 r live in out.*
    // between chunks, seeded with the decayed old sums
    for (int k = lane; k < nk; k += 32) {
      out.cross[c * nk + k] = in.cross[c * nk + k] * decay;
      out.m1[c * nk + k] = in.m1[c * nk + k] * decay;
      out.m2[c * nk + k] = in.m2[c * nk + k] * decay;
      out.pairs[c * nk + k] = in.pairs[c * nk + k] * decay;
    }
    for (int64_t pos = 0; pos < n; pos += OS_CHUNK) {
      const int cn = (int)((n - pos) < OS_CHUNK ? (n - pos) : OS_CHUNK);
      for (int i = lane; i < cn; i += 32) z[L + i] = (double)x[pos + i];
      __syncwarp();
      for (int r = 0; r < rounds; ++r) {
        const int k = lane + 32 * r;
        if (k < nk) {
          double sc = 0.0, sl = 0.0, su = 0.0;
          int t0 = first + k - L;
          t0 = t0 > 0 ? t0 : 0;
          for (int t = t0; t < cn; ++t) {
            const double cur = z[L + t], lag = z[L + t - k];
            sc = fma(cur, lag, sc);
            sl += lag;
            su += cur;
          }
          const int np = cn - t0 > 0 ? cn - t0 : 0;
          out.cross[c * nk + k] += sc;
          out.m1[c * nk + k] += sl;
          out.m2[c * nk + k] += su;
          out.pairs[c * nk + k] += (double)np;
        }
      }
      __syncwarp();
      // roll: keep the last L samples of z[0 .. L+cn), right-aligned
      for (int g = 0; g < L; g += 32) {
        const int i = g + lane;
        const double v = i < L ? z[i + cn] : 0.0;
        __syncwarp();
        if (i < L) z[i] = v;
      }
      first = first - cn > 0 ? first - cn : 0;
      __syncwarp();
    }
    for (int i = lane; i < L; i += 32) out.buf[c * L + i] = z[i];
  }
}

// phase 0: out[0] = sum count, out[1] = sum count * mean, out[2] = sum mean
// phase 1: out[0] = sum M2, out[1] = sum count (mean - gmean)^2, out[2] = sum (mean - mbar)^2, out[3] = sum M2 / max(count, 1),
//          out[4 + k] = sum over chains of  cross/n - (m1/n)(m2/n),  n = max(pairs, 1)        (accumulator.py:366-377)
__global__ void __launch_bounds__(OS_WARPS * 32) online_summary_kernel(OsPtrs in, int64_t n_chains, int L, int phase, double gmean,
                                                                       double mbar, double *__restrict__ out) {
  __shared__ double sh[OS_WARPS][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
  double a[4] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t c = tid; c < n_chains; c += nthr) {
    const double cnt = in.count[c], mu = in.mean[c];
    if (phase == 0) {
      a[0] += cnt;
      a[1] += cnt * mu;
      a[2] += mu;
    } else {
      const double M2 = in.M2[c];
      a[0] += M2;
      a[1] += cnt * (mu - gmean) * (mu - gmean);
      a[2] += (mu - mbar) * (mu - mbar);
      a[3] += M2 / (cnt > 1.0 ? cnt : 1.0);
    }
  }
  for (int qn = 0; qn < 4; ++qn) {
    const double r = warp_sum(a[qn]);
    __syncthreads();
    if (lane == 0) sh[warp][0] = r;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int w = 0; w < OS_WARPS; ++w) tot += sh[w][0];
      if (tot != 0.0) atomicAdd(out + qn, tot);
    }
  }
  if (phase == 0 || L == 0) return;
  const int nk = L + 1;
  const int64_t gw = (int64_t)blockIdx.x * OS_WARPS + warp, nw = (int64_t)gridDim.x * OS_WARPS;
  for (int k0 = 0; k0 < nk; k0 += 32) {
    const int k = k0 + lane;
    double acc = 0.0;
    if (k < nk)
      for (int64_t c = gw; c < n_chains; c += nw) {
        const double np = in.pairs[c * nk + k], nn = np > 1.0 ? np : 1.0;
        acc += in.cross[c * nk + k] / nn - (in.m1[c * nk + k] / nn) * (in.m2[c * nk + k] / nn);
      }
    __syncthreads();
    sh[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && k < nk) {
      double tot = 0.0;
      for (int w = 0; w < OS_WARPS; ++w) tot += sh[w][lane];
      atomicAdd(out + 4 + k, tot);
    }
  }
}

static OsPtrs in_ptrs(const nk_online_stats_t *s) {
  return OsPtrs{s->chain_count, s->chain_mean, s->chain_M2, s->cross_sum, s->m1_sum, s->m2_sum, s->pair_count, s->chain_buf};
}

int online_stats_update(cudaStream_t stream, const nk_online_stats_t *in, const nk_online_stats_t *out, const void *data, int32_t dtype,
                        int64_t n, double decay) {
  const int L = in->max_lag;
  const size_t per_warp = (size_t)(L + OS_CHUNK) * sizeof(double);
  int warps = (int)((48 * 1024) / per_warp);
  warps = warps > OS_WARPS ? OS_WARPS : warps;
  const int64_t need = (in->n_chains + warps - 1) / warps;
  const int64_t cap = (int64_t)num_sms() * 8;
  const int grid = (int)(need < cap ? need : cap);
  const OsPtrs ip = in_ptrs(in);
  const OsOut op{out->chain_count, out->chain_mean, out->chain_M2, out->cross_sum, out->m1_sum, out->m2_sum, out->pair_count, out->chain_buf};
  const size_t smem = per_warp * warps;
  if (dtype == NK_F32)
    online_update_kernel<float><<<grid, OS_WARPS * 32, smem, stream>>>(ip, op, (const float *)data, in->n_chains, n, L, in->buf_len, decay, warps);
  else
    online_update_kernel<double><<<grid, OS_WARPS * 32, smem, stream>>>(ip, op, (const double *)data, in->n_chains, n, L, in->buf_len, decay,
                                                                       warps);
  NK_LAUNCH_OK();
  return NK_OK;
}

int online_stats_summary(cudaStream_t stream, const nk_online_stats_t *s, int32_t phase, double gmean, double mbar, double *out) {
  const int n_out = phase == 0 ? 3 : 4 + (s->max_lag > 0 ? s->max_lag + 1 : 0);
  NK_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(double) * n_out, stream));
  if (s->n_chains == 0) return NK_OK;
  const int64_t need = (s->n_chains + OS_WARPS - 1) / OS_WARPS;
  const int64_t cap = (int64_t)num_sms() * 4;
  const int grid = (int)(need < cap ? need : cap);
  online_summary_kernel<<<grid, OS_WARPS * 32, 0, stream>>>(in_ptrs(s), s->n_chains, s->max_lag, phase, gmean, mbar, out);
  NK_LAUNCH_OK();
  return NK_OK;
}

// Host arithmetic of accumulator.py:240-447 and check_mc_convergence.py:243-272 on globally reduced sums.
int online_stats_finalize(const double *p0, const double *p1, int64_t n_chains, int64_t n_samples, int32_t L, double *o, double *acf) {
  const double nan = NAN;
  for (int i = 0; i < NK_ONLINE_NOUT; ++i) o[i] = nan;
  o[7] = 0.0;
  o[8] = 0.0;
  if (acf)
    for (int k = 0; k <= L && L > 0; ++k) acf[k] = nan;  // NaN = no autocorrelation function (accumulator.py:353-377 returns None)
  const double total = p0[0];
  if (total == 0.0) return NK_OK;
  const double nc = (double)n_chains;
  const double mean = p0[1] / total;
  const double variance = (p1[0] + p1[1]) / total;
  const double var_means = p1[2] / nc;  // jnp.var(chain_mean)
  // acf (accumulator.py:353-377)
  bool have_acf = false;
  if (L > 0) {
    const double c0 = p1[4] / nc;
    if (c0 > 0.0) {
      have_acf = true;
      for (int k = 0; k <= L; ++k) {
        const double r = (p1[4 + k] / nc) / c0;
        if (acf) acf[k] = r;
      }
    }
  }
  if (!have_acf && acf)
    for (int k = 0; k <= L && L > 0; ++k) acf[k] = nan;
  // Geyer initial positive sequence + initial monotone sequence (accumulator.py:312-351)
  double tau_acf = nan;
  bool saturated = false;
  const int m = have_acf ? (L + 1) / 2 : 0;
  if (m > 0) {
    const double c0 = p1[4];
    double sum = 0.0, running = INFINITY;
    int t = 0;
    for (; t < m; ++t) {
      const double P = (p1[4 + 2 * t] + p1[4 + 2 * t + 1]) / c0;
      if (!(P > 0.0)) break;
      running = P < running ? P : running;
      sum += running;
    }
    saturated = t == m;
    tau_acf = t == 0 ? 1.0 : fmax(2.0 * sum - 1.0, 1.0);
  }
  // batch estimate (accumulator.py:275-309)
  double tau_batch = nan;
  if (n_chains >= 2 && variance > 0.0) tau_batch = fmax((((double)n_samples / nc) * var_means / variance - 1.0) * 0.5, 0.0);
  // R_hat (accumulator.py:379-395)
  double rhat = nan;
  if (n_chains >= 2) {
    const double W = p1[3] / nc, N = total / nc;
    if (W > 0.0) rhat = sqrt((N - 1.0) / N + var_means / W);
  }
  // error of the mean (accumulator.py:430-447)
  double err = nan;
  if (n_chains > 1)
    err = sqrt(var_means / nc);
  else if (!isnan(tau_acf))
    err = sqrt(variance * tau_acf / (double)n_samples);
  o[0] = mean;
  o[1] = err;
  o[2] = variance;
  o[3] = isnan(tau_acf) ? tau_batch : tau_acf;
  o[4] = rhat;
  o[5] = tau_batch;
  o[6] = tau_acf;
  o[7] = saturated ? 1.0 : 0.0;
  // tau_corr_reliable (check_mc_convergence.py:258-272)
  o[8] = (!saturated && !isnan(tau_acf) && tau_acf > 0.0 && ((double)n_samples / nc) / tau_acf >= 50.0) ? 1.0 : 0.0;
  return NK_OK;
}

}  // namespace nk
