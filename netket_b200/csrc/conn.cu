// Connected elements (get_conn_padded) as stand-alone, HBM-write-bound kernels.
//
//  * ising_conn_kernel    IsingJax.get_conn_padded   netket/operator/_ising/jax.py:82-88,125-165
//  * ising_nconn_kernel   IsingJax.n_conn            netket/operator/_ising/jax.py:71-80,168-175
//  * localop_conn_kernel  LocalOperatorJax._get_conn_padded  netket/operator/_local_operator/jax.py:74-201,256-284
//
// Output volume is K*N bytes per sample (101x the input for 10x10 TFIM), so the kernels are organised around
// coalesced 32-bit stores: one warp owns one sample, keeps sigma in shared memory as words and streams the
// K rows out, patching the flipped byte(s) with an XOR (int8 +1 = 0x01, -1 = 0xFF, negation = XOR 0xFE).
#include "kernels.cuh"

namespace nk {

__device__ __forceinline__ uint32_t fastdiv(uint32_t x, uint32_t magic) { return __umulhi(x, magic); }
static inline uint32_t make_magic(uint32_t d) { return (uint32_t)(((1ull << 32) + d - 1) / d); }

// ------------------------------------------------------------------------------------------ Ising
struct IsingConnArgs {
  const int8_t *x;
  int64_t B;
  int32_t N, K;
  const int32_t *edges;
  int32_t n_edges;
  double h, J;
  int8_t *xp;
  void *mels;
  int32_t mel_dtype;
  int32_t n_pad;       // bytes of sigma per warp in smem (multiple of 16)
  uint32_t magic_row;  // magic for division by words-per-row (word path) or N (byte path)
};

template <bool WORDS>
__global__ void __launch_bounds__(256) ising_conn_kernel(const __grid_constant__ IsingConnArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  int8_t *sig = reinterpret_cast<int8_t *>(smem_raw) + (size_t)warp * p.n_pad;
  const int N = p.N, K = p.K;
  for (int64_t s = (int64_t)blockIdx.x * warps + warp; s < p.B; s += (int64_t)gridDim.x * warps) {
    for (int i = lane; i < N; i += 32) sig[i] = p.x[s * N + i];
    __syncwarp();
    // slot 0: J * sum_e (2*[s_i == s_j] - 1) = J * sum_e s_i s_j   (_ising_mels_jax, jax.py:125-137)
    int zz = 0;
    for (int e = lane; e < p.n_edges; e += 32) zz += (int)sig[p.edges[2 * e]] * (int)sig[p.edges[2 * e + 1]];
    zz = __reduce_add_sync(0xffffffffu, zz);
    const double mel0 = p.J * (double)zz;
    for (int k = lane; k < K; k += 32) {
      const double m = (k == 0) ? mel0 : -p.h;
      if (p.mel_dtype == NK_F64)
        reinterpret_cast<double *>(p.mels)[s * K + k] = m;
      else
        reinterpret_cast<float *>(p.mels)[s * K + k] = (float)m;
    }
    // slots: row 0 = sigma, row k = sigma with site k-1 flipped (eye(N+1, N, k=-1), jax.py:158-160)
    if (WORDS && N <= 128) {
      // rows of at most 32 words: lane l owns word column l of every row (no division; the rows of a sample are contiguous, so a
      // warp store covers N contiguous bytes).  Row k flips site k - 1: the lane whose word holds it patches that byte, i.e.
      // lane l patches rows 4l + 1 .. 4l + 4.  (The flat loop below spends ~13 instructions per stored word on index
      // arithmetic and was issue-bound at 0.78 of the HBM write bandwidth: ncu, profiles/r02_k4_ising_conn_ncu_raw.csv.)
      const uint32_t wpr = (uint32_t)N >> 2;
      if ((uint32_t)lane < wpr) {
        const uint32_t myw = reinterpret_cast<const uint32_t *>(sig)[lane];
        uint32_t *out = reinterpret_cast<uint32_t *>(p.xp + (size_t)s * K * N) + lane;
        const uint32_t first = 4u * (uint32_t)lane + 1u;  // first row this lane patches
        int k = 0;
        for (; k + 4 <= K; k += 4) {
#pragma unroll
          for (int d = 0; d < 4; ++d) {
            const uint32_t off = (uint32_t)(k + d) - first;
            out[(size_t)(k + d) * wpr] = off < 4u ? myw ^ (0xFEu << (off * 8u)) : myw;
          }
        }
        for (; k < K; ++k) {
          const uint32_t off = (uint32_t)k - first;
          out[(size_t)k * wpr] = off < 4u ? myw ^ (0xFEu << (off * 8u)) : myw;
        }
      }
    } else if (WORDS) {
      const uint32_t wpr = (uint32_t)N >> 2;
      const uint32_t total = (uint32_t)K * wpr;
      const uint32_t *sw = reinterpret_cast<const uint32_t *>(sig);
      uint32_t *out = reinterpret_cast<uint32_t *>(p.xp + (size_t)s * K * N);
      for (uint32_t w = lane; w < total; w += 32) {
        const uint32_t k = fastdiv(w, p.magic_row);
        const uint32_t wn = w - k * wpr;
        uint32_t v = sw[wn];
        const uint32_t site = k - 1u;  // wraps for k == 0 -> never matches
        if ((site >> 2) == wn && k != 0u) v ^= 0xFEu << ((site & 3u) * 8u);
        out[w] = v;
      }
    } else {
      const uint32_t total = (uint32_t)K * (uint32_t)N;
      int8_t *out = p.xp + (size_t)s * K * N;
      for (uint32_t idx = lane; idx < total; idx += 32) {
        const uint32_t k = fastdiv(idx, p.magic_row);
        const uint32_t n = idx - k * (uint32_t)N;
        int8_t v = sig[n];
        if (k != 0u && n == k - 1u) v = (int8_t)(-(int)v);
        out[idx] = v;
      }
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256) ising_nconn_kernel(const int8_t *__restrict__ x, int64_t B, int N, const int32_t *edges,
                                                          int n_edges, double h, double J, int32_t *out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  for (int64_t s = (int64_t)blockIdx.x * warps + warp; s < B; s += (int64_t)gridDim.x * warps) {
    int zz = 0;
    for (int e = lane; e < n_edges; e += 32) zz += (int)x[s * N + edges[2 * e]] * (int)x[s * N + edges[2 * e + 1]];
    zz = __reduce_add_sync(0xffffffffu, zz);
    if (lane == 0) out[s] = (h == 0.0 ? 0 : N) + ((J * (double)zz) != 0.0 ? 1 : 0);
  }
}

int ising_conn(cudaStream_t stream, const nk_ising_t &op, const int8_t *x, int64_t B, int32_t N, int8_t *xp, void *mels,
               int32_t mel_dtype) {
  if (B == 0) return NK_OK;
  IsingConnArgs a{};
  a.x = x;
  a.B = B;
  a.N = N;
  a.K = (op.h == 0.0) ? 1 : N + 1;
  a.edges = op.edges;
  a.n_edges = op.n_edges;
  a.h = op.h;
  a.J = op.J;
  a.xp = xp;
  a.mels = mels;
  a.mel_dtype = mel_dtype;
  a.n_pad = (N + 15) & ~15;
  const bool words = (N % 4 == 0);
  {  // fastdiv(x, magic) is exact while x * divisor < 2^32
    const uint64_t div = words ? (uint64_t)N / 4 : (uint64_t)N;
    if ((uint64_t)a.K * div * div >= (1ull << 32)) {
      set_error("nk_ising_conn: N=%d too large", N);
      return NK_EUNSUPPORTED;
    }
    a.magic_row = make_magic((uint32_t)div);
  }
  const int warps = 8;
  const size_t smem = (size_t)warps * a.n_pad;
  const int64_t need = (B + warps - 1) / warps;
  const int64_t cap = (int64_t)num_sms() * 8;
  const int grid = (int)(need < cap ? need : cap);
  if (words)
    ising_conn_kernel<true><<<grid, warps * 32, smem, stream>>>(a);
  else
    ising_conn_kernel<false><<<grid, warps * 32, smem, stream>>>(a);
  NK_LAUNCH_OK();
  return NK_OK;
}

int ising_n_conn(cudaStream_t stream, const nk_ising_t &op, const int8_t *x, int64_t B, int32_t N, int32_t *out) {
  if (B == 0) return NK_OK;
  const int warps = 8;
  const int64_t need = (B + warps - 1) / warps;
  const int64_t cap = (int64_t)num_sms() * 8;
  ising_nconn_kernel<<<(int)(need < cap ? need : cap), warps * 32, 0, stream>>>(x, B, N, op.edges, op.n_edges, op.h, op.J, out);
  NK_LAUNCH_OK();
  return NK_OK;
}

// ------------------------------------------------------------------------------------------ LocalOperator
struct LocalopConnArgs {
  nk_localop_t op;
  const int8_t *x;
  int64_t B;
  int32_t N, K;
  int8_t *xp;
  void *mels;
  int32_t mel_dtype;
  int32_t *nconn;
  int32_t n_pad;
  int32_t n_cand;  // number of candidate slots before compaction (excluding the trailing pad)
  uint32_t magic_row;
};

// Row descriptor kept in shared memory for every output slot: which sites are overwritten with which spin.
//   bits 0..13 site0, 14..27 site1, 28 new idx0, 29 new idx1, 30 two-site flag, 31 valid (otherwise row = sigma)
__device__ __forceinline__ uint32_t pack_row(int s0, int s1, int b0, int b1, bool two) {
  return (uint32_t)s0 | ((uint32_t)s1 << 14) | ((uint32_t)b0 << 28) | ((uint32_t)b1 << 29) | ((two ? 1u : 0u) << 30) | (1u << 31);
}

template <bool WORDS>
__global__ void __launch_bounds__(256) localop_conn_kernel(const __grid_constant__ LocalopConnArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const int N = p.N, K = p.K;
  uint32_t *rowdesc = reinterpret_cast<uint32_t *>(smem_raw) + (size_t)warp * K;
  int8_t *sig = reinterpret_cast<int8_t *>(smem_raw + (size_t)warps * K * sizeof(uint32_t)) + (size_t)warp * p.n_pad;
  const nk_localop_t &op = p.op;

  for (int64_t s = (int64_t)blockIdx.x * warps + warp; s < p.B; s += (int64_t)gridDim.x * warps) {
    for (int i = lane; i < N; i += 32) sig[i] = p.x[s * N + i];
    for (int k = lane; k < K; k += 32) rowdesc[k] = 0u;
    __syncwarp();

    int pos = 0;  // next free output slot (warp-uniform)
    // --- diagonal slot: constant + sum over terms of diag_mels[row]  (jax.py:117-131)
    if (op.nonzero_diagonal) {
      double d = 0.0;
      for (int g = 0; g < op.n_groups; ++g) {
        const nk_localop_group_t &G = op.groups[g];
        const int rows = 1 << G.n_sites;
        for (int o = lane; o < G.n_ops; o += 32) {
          const int s0 = G.acting_on[o * G.n_sites];
          const int x0 = sig[s0] > 0 ? 0 : 1;
          int row = x0;
          if (G.n_sites == 2) row = 2 * x0 + (sig[G.acting_on[o * 2 + 1]] > 0 ? 0 : 1);
          d += G.diag_mels[o * rows + row];
        }
      }
      // fixed-order reduction so that the diagonal is reproducible
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) d += __shfl_xor_sync(0xffffffffu, d, m);
      d += op.constant;
      if (fabs(d) > op.mel_cutoff) {
        if (lane == 0) {
          if (p.mel_dtype == NK_F64)
            reinterpret_cast<double *>(p.mels)[s * K] = d;
          else
            reinterpret_cast<float *>(p.mels)[s * K] = (float)d;
        }
        pos = 1;  // rowdesc[0] stays 0 => row = sigma
      }
    }
    // --- off-diagonal candidates in (group, term, entry) order; keep |mel| > cutoff, compact in order (jax.py:177-199)
    for (int g = 0; g < op.n_groups; ++g) {
      const nk_localop_group_t &G = op.groups[g];
      const int rows = 1 << G.n_sites;
      const int total = G.n_ops * G.ncmax;
      for (int base = 0; base < total; base += 32) {
        const int idx = base + lane;
        bool valid = false;
        double mel = 0.0;
        uint32_t desc = 0u;
        if (idx < total) {
          const int o = idx / G.ncmax, c = idx - o * G.ncmax;
          const int s0 = G.acting_on[o * G.n_sites];
          const int s1 = G.n_sites == 2 ? G.acting_on[o * 2 + 1] : s0;
          const int x0 = sig[s0] > 0 ? 0 : 1;
          const int x1 = sig[s1] > 0 ? 0 : 1;
          const int row = G.n_sites == 2 ? 2 * x0 + x1 : x0;
          if (c < G.n_conns[o * rows + row]) {
            mel = G.mels[((size_t)o * rows + row) * G.ncmax + c];
            valid = fabs(mel) > op.mel_cutoff;
            const int8_t *xpr = G.x_prime + (((size_t)o * rows + row) * G.ncmax + c) * G.n_sites;
            desc = pack_row(s0, s1, xpr[0], G.n_sites == 2 ? xpr[1] : xpr[0], G.n_sites == 2);
          }
        }
        const unsigned m = __ballot_sync(0xffffffffu, valid);
        if (valid) {
          const int k = pos + __popc(m & ((1u << lane) - 1u));
          if (k < K) {
            rowdesc[k] = desc;
            if (p.mel_dtype == NK_F64)
              reinterpret_cast<double *>(p.mels)[s * K + k] = mel;
            else
              reinterpret_cast<float *>(p.mels)[s * K + k] = (float)mel;
          }
        }
        pos += __popc(m);
      }
    }
    const int n_conn = pos;
    // padding: (sigma, 0)
    for (int k = min(n_conn, K) + lane; k < K; k += 32) {
      if (p.mel_dtype == NK_F64)
        reinterpret_cast<double *>(p.mels)[s * K + k] = 0.0;
      else
        reinterpret_cast<float *>(p.mels)[s * K + k] = 0.0f;
    }
    if (p.nconn != nullptr && lane == 0) p.nconn[s] = n_conn;
    __syncwarp();

    if (WORDS) {
      const uint32_t wpr = (uint32_t)N >> 2;
      const uint32_t total = (uint32_t)K * wpr;
      const uint32_t *sw = reinterpret_cast<const uint32_t *>(sig);
      uint32_t *out = reinterpret_cast<uint32_t *>(p.xp + (size_t)s * K * N);
      for (uint32_t w = lane; w < total; w += 32) {
        const uint32_t k = fastdiv(w, p.magic_row);
        const uint32_t wn = w - k * wpr;
        uint32_t v = sw[wn];
        const uint32_t d = rowdesc[k];
        if (d >> 31) {
          const uint32_t s0 = d & 0x3FFFu, s1 = (d >> 14) & 0x3FFFu;
          if ((s0 >> 2) == wn) {
            const uint32_t sh = (s0 & 3u) * 8u;
            v = (v & ~(0xFFu << sh)) | ((((d >> 28) & 1u) ? 0xFFu : 0x01u) << sh);
          }
          if ((s1 >> 2) == wn) {
            const uint32_t sh = (s1 & 3u) * 8u;
            v = (v & ~(0xFFu << sh)) | ((((d >> 29) & 1u) ? 0xFFu : 0x01u) << sh);
          }
        }
        out[w] = v;
      }
    } else {
      const uint32_t total = (uint32_t)K * (uint32_t)N;
      int8_t *out = p.xp + (size_t)s * K * N;
      for (uint32_t idx = lane; idx < total; idx += 32) {
        const uint32_t k = fastdiv(idx, p.magic_row);
        const uint32_t n = idx - k * (uint32_t)N;
        int8_t v = sig[n];
        const uint32_t d = rowdesc[k];
        if (d >> 31) {
          if ((d & 0x3FFFu) == n) v = ((d >> 28) & 1u) ? (int8_t)-1 : (int8_t)1;
          if (((d >> 14) & 0x3FFFu) == n) v = ((d >> 29) & 1u) ? (int8_t)-1 : (int8_t)1;
        }
        out[idx] = v;
      }
    }
    __syncwarp();
  }
}

int localop_conn(cudaStream_t stream, const nk_localop_t &op, const int8_t *x, int64_t B, int32_t N, int8_t *xp, void *mels,
                 int32_t mel_dtype, int32_t *nconn) {
  if (B == 0) return NK_OK;
  LocalopConnArgs a{};
  a.op = op;
  a.x = x;
  a.B = B;
  a.N = N;
  a.K = op.max_conn_size;
  a.xp = xp;
  a.mels = mels;
  a.mel_dtype = mel_dtype;
  a.nconn = nconn;
  a.n_pad = (N + 15) & ~15;
  if (N >= (1 << 14)) {
    set_error("nk_localop_conn: N=%d exceeds the 14-bit site field", N);
    return NK_EUNSUPPORTED;
  }
  if (a.K <= 0) return NK_OK;  // empty operator: outputs have a zero-sized K axis
  const bool words = (N % 4 == 0);
  const uint64_t div = words ? (uint64_t)N / 4 : (uint64_t)N;
  a.magic_row = make_magic((uint32_t)div);
  const size_t per_warp = (size_t)a.K * 4 + a.n_pad;
  int warps = 8;
  while (warps > 1 && warps * per_warp > 200 * 1024) warps >>= 1;
  if (warps * per_warp > 200 * 1024 || (uint64_t)a.K * div * div >= (1ull << 32)) {
    set_error("nk_localop_conn: K=%d too large", a.K);
    return NK_EUNSUPPORTED;
  }
  const size_t smem = warps * per_warp;
  auto kern = words ? localop_conn_kernel<true> : localop_conn_kernel<false>;
  NK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int64_t need = (B + warps - 1) / warps;
  const int64_t cap = (int64_t)num_sms() * 8;
  kern<<<(int)(need < cap ? need : cap), warps * 32, smem, stream>>>(a);
  NK_LAUNCH_OK();
  return NK_OK;
}

}  // namespace nk
