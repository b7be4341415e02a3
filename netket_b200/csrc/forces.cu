// Forces  F_k = < d log psi / d p_k  * (E_loc - <E_loc>) >  of an RBM over a batch of Monte-Carlo samples.
//
// Replaces the vjp of netket/vqs/mc/mc_state/expect_forces.py:69-112 (`forces_expect_hermitian`: O_loc -= mean, then
// vjp of the model with conj(O_loc) / n_samples) for nk.models.RBM with real parameters, whose log-derivatives are closed
// forms (netket/models/rbm.py:57-81):   d/dW_ij = sigma_i tanh(theta_j),   d/db_j = tanh(theta_j),   d/da_i = sigma_i.
// So  F_W = sigma^T (tanh(theta) * w) / n,  F_b = 1^T (tanh(theta) * w) / n,  F_a = sigma^T w / n  with  w = E_loc - mean:
// one (N x n_s)(n_s x M) contraction on theta, which the theta GEMM has already produced for the whole batch.
//
// forces_kernel (fp32 shapes outside the tcgen05 kernel's limits) / forces_dmma_kernel (fp64): a CTA owns a 128 (sites) x 64 (hidden units) tile of F_W and streams over its share of the samples,
// 32 at a time through shared memory; partial sums live in registers (32 per thread) and are flushed with double-precision
// atomics, so that only sums cross CTAs (and, between GPUs, only the N*M + M + N doubles all-reduced by the caller).
#include "kernels.cuh"

namespace nk {

constexpr int F_TI = 128, F_TJ = 64, F_THREADS = 256;
template <typename T>
struct ForcesTile {
  static constexpr int TS = sizeof(T) == 4 ? 32 : 16;  // samples staged per step (48 KB of static shared memory at most)
};

template <typename T>
__device__ __forceinline__ T tanh_t(T x);
template <>
__device__ __forceinline__ float tanh_t<float>(float x) { return tanhf(x); }
template <>
__device__ __forceinline__ double tanh_t<double>(double x) { return tanh(x); }

struct ForcesArgs {
  const int8_t *sigma;  // [Ns, N]
  const void *theta;    // [Ns, M] T: theta, or tanh(theta) if is_tanh
  int32_t is_tanh;
  const void *eloc;     // [Ns] eloc_dtype
  int32_t eloc_dtype;
  int64_t Ns;
  double mean;
  int32_t N, M;
  int32_t want_b, want_a;
  double *sums;  // [N*M | M | N], accumulated into (the caller zeroes it)
};

template <typename T>
__global__ void __launch_bounds__(F_THREADS) forces_kernel(const __grid_constant__ ForcesArgs p) {
  constexpr int F_TS = ForcesTile<T>::TS;
  __shared__ __align__(16) T sg[F_TS][F_TI];  // sigma as +-1 in T
  __shared__ __align__(16) T xs[F_TS][F_TJ];  // tanh(theta) * w
  __shared__ T ws[F_TS];
  const int t = threadIdx.x;
  const int j0 = blockIdx.y * F_TJ, i0 = blockIdx.z * F_TI;
  const int jg = t & 15, ig = t >> 4;  // this thread's 4 hidden units and 8 sites inside the tile
  const T *theta = reinterpret_cast<const T *>(p.theta);
  T acc[8][4];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = T(0);
  T fb = T(0), fa = T(0);  // thread t sums x for hidden unit j0 + (t & 63) and sigma * w for site i0 + (t & 127)
  const int64_t n_batches = (p.Ns + F_TS - 1) / F_TS;
  for (int64_t bt = blockIdx.x; bt < n_batches; bt += gridDim.x) {
    const int64_t s0 = bt * F_TS;
    if (t < F_TS) {
      const int64_t s = s0 + t;
      double w = 0.0;
      if (s < p.Ns) w = (p.eloc_dtype == NK_F64 ? reinterpret_cast<const double *>(p.eloc)[s] : (double)reinterpret_cast<const float *>(p.eloc)[s]) - p.mean;
      ws[t] = (T)w;
    }
    __syncthreads();
    // sigma tile: thread t -> site i0 + (t & 127), samples (t >> 7) + 2 k
    {
      const int i = i0 + (t & 127);
#pragma unroll 4
      for (int k = 0; k < F_TS / 2; ++k) {
        const int sl = (t >> 7) + 2 * k;
        const int64_t s = s0 + sl;
        T v = T(0);
        if (i < p.N && s < p.Ns) v = (T)p.sigma[s * p.N + i];
        sg[sl][t & 127] = v;
        fa += v * ws[sl];
      }
    }
    // x tile: thread t -> hidden unit j0 + (t & 63), samples (t >> 6) + 4 k
    {
      const int j = j0 + (t & 63);
#pragma unroll 4
      for (int k = 0; k < F_TS / 4; ++k) {
        const int sl = (t >> 6) + 4 * k;
        const int64_t s = s0 + sl;
        T v = T(0);
        if (j < p.M && s < p.Ns) v = (p.is_tanh ? theta[s * p.M + j] : tanh_t<T>(theta[s * p.M + j])) * ws[sl];
        xs[sl][t & 63] = v;
        fb += v;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int sl = 0; sl < F_TS; ++sl) {
      T a[8], b[4];
#pragma unroll
      for (int q = 0; q < 8; ++q) a[q] = sg[sl][ig * 8 + q];
#pragma unroll
      for (int q = 0; q < 4; ++q) b[q] = xs[sl][jg * 4 + q];
#pragma unroll
      for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[q][r] = Math<T>::fma(a[q], b[r], acc[q][r]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int i = i0 + ig * 8 + q;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int j = j0 + jg * 4 + r;
      if (i < p.N && j < p.M) atomicAdd(p.sums + (size_t)i * p.M + j, (double)acc[q][r]);
    }
  }
  if (p.want_b && blockIdx.z == 0 && j0 + (t & 63) < p.M) atomicAdd(p.sums + (size_t)p.N * p.M + j0 + (t & 63), (double)fb);
  if (p.want_a && blockIdx.y == 0 && i0 + (t & 127) < p.N)
    atomicAdd(p.sums + (size_t)p.N * p.M + p.M + i0 + (t & 127), (double)fa);
}

// ---- fp64: the same contraction on the FP64 tensor cores (DMMA m8n8k4).  Tile and staging as above (128 sites x 64
// hidden units per CTA, F_b / F_a accumulated while staging); a warp owns 16 sites x 64 hidden units = 2 x 8 DMMA tiles
// (32 accumulators per thread) and per K-step of 4 samples issues 16 DMMAs for 10 fragment loads.  sigma is staged as
// int8 and converted when the A fragment is read; the x rows are padded to 68 doubles (conflict-free B fragments).
constexpr int FD_TS = 64, FD_XS = 68;

__device__ __forceinline__ void dmma884_acc(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(F_THREADS, 2) forces_dmma_kernel(const __grid_constant__ ForcesArgs p) {
  __shared__ __align__(16) int8_t sg[FD_TS][F_TI + 4];   // sigma; row stride 132 bytes = 33 words (odd: conflict-free A fragments)
  __shared__ __align__(16) double xs[FD_TS][FD_XS];      // tanh(theta) * w
  __shared__ double ws[FD_TS];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int j0 = blockIdx.y * F_TJ, i0 = blockIdx.z * F_TI;
  const int fr = lane >> 2, fc = lane & 3;
  const double *theta = reinterpret_cast<const double *>(p.theta);
  double acc[2][8][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  double fb = 0.0, fa = 0.0;
  double fa4[4] = {0.0, 0.0, 0.0, 0.0};  // word path of the sigma tile: this thread's four sites
  // sigma rows of whole, aligned 32-bit words: 8 word loads per thread and batch, all in flight together (the byte path issues 32
  // loads in groups of 4: with ~16 warps per SM the staging sat on the latency of HBM eight times per batch)
  const bool sig_words = (p.N & 3) == 0 && (i0 & 3) == 0 && (reinterpret_cast<uintptr_t>(p.sigma) & 3) == 0;
  const int64_t n_batches = (p.Ns + FD_TS - 1) / FD_TS;
  for (int64_t bt = blockIdx.x; bt < n_batches; bt += gridDim.x) {
    const int64_t s0 = bt * FD_TS;
    __syncthreads();  // previous stage consumed
    if (t < FD_TS) {
      const int64_t s = s0 + t;
      double w = 0.0;
      if (s < p.Ns) w = (p.eloc_dtype == NK_F64 ? reinterpret_cast<const double *>(p.eloc)[s] : (double)reinterpret_cast<const float *>(p.eloc)[s]) - p.mean;
      ws[t] = w;
    }
    __syncthreads();
    if (sig_words) {  // sigma tile: thread t -> sites i0 + 4 (t & 31) .. + 3, samples (t >> 5) + 8 k
      const int wc = t & 31;
      uint32_t v[FD_TS / 8];
#pragma unroll
      for (int k = 0; k < FD_TS / 8; ++k) {
        const int64_t s = s0 + (t >> 5) + 8 * k;
        v[k] = (i0 + 4 * wc < p.N && s < p.Ns) ? *reinterpret_cast<const uint32_t *>(p.sigma + s * p.N + i0 + 4 * wc) : 0u;
      }
#pragma unroll
      for (int k = 0; k < FD_TS / 8; ++k) {
        const int sl = (t >> 5) + 8 * k;
        *reinterpret_cast<uint32_t *>(&sg[sl][4 * wc]) = v[k];
        const double w = ws[sl];
#pragma unroll
        for (int r = 0; r < 4; ++r) fa4[r] += (double)(int8_t)(v[k] >> (8 * r)) * w;
      }
    } else {  // sigma tile: thread t -> site i0 + (t & 127), samples (t >> 7) + 2 k
      const int i = i0 + (t & 127);
#pragma unroll 4
      for (int k = 0; k < FD_TS / 2; ++k) {
        const int sl = (t >> 7) + 2 * k;
        const int64_t s = s0 + sl;
        int8_t v = 0;
        if (i < p.N && s < p.Ns) v = p.sigma[s * p.N + i];
        sg[sl][t & 127] = v;
        fa += (double)v * ws[sl];
      }
    }
    {  // x tile: thread t -> hidden unit j0 + (t & 63), samples (t >> 6) + 4 k; 8 loads in flight at a time
      const int j = j0 + (t & 63);
#pragma unroll
      for (int kb = 0; kb < FD_TS / 4; kb += 8) {
        double th[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int64_t s = s0 + (t >> 6) + 4 * (kb + k);
          th[k] = (j < p.M && s < p.Ns) ? theta[s * p.M + j] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int sl = (t >> 6) + 4 * (kb + k);
          const double v = (p.is_tanh ? th[k] : tanh(th[k])) * ws[sl];
          xs[sl][t & 63] = v;
          fb += v;
        }
      }
    }
    __syncthreads();
    const int8_t *sw = &sg[fc][warp * 16 + fr];
    const double *xw = &xs[fc][fr];
#pragma unroll 4
    for (int ks = 0; ks < FD_TS; ks += 4) {
      const double a0 = (double)sw[ks * (F_TI + 4)], a1 = (double)sw[ks * (F_TI + 4) + 8];
      double bf[8];
#pragma unroll
      for (int rn = 0; rn < 8; ++rn) bf[rn] = xw[ks * FD_XS + 8 * rn];
#pragma unroll
      for (int rn = 0; rn < 8; ++rn) {
        dmma884_acc(acc[0][rn][0], acc[0][rn][1], a0, bf[rn]);
        dmma884_acc(acc[1][rn][0], acc[1][rn][1], a1, bf[rn]);
      }
    }
  }
#pragma unroll
  for (int rm = 0; rm < 2; ++rm) {
    const int i = i0 + warp * 16 + 8 * rm + fr;
#pragma unroll
    for (int rn = 0; rn < 8; ++rn) {
      const int j = j0 + 8 * rn + 2 * fc;
      if (i < p.N && j < p.M) atomicAdd(p.sums + (size_t)i * p.M + j, acc[rm][rn][0]);
      if (i < p.N && j + 1 < p.M) atomicAdd(p.sums + (size_t)i * p.M + j + 1, acc[rm][rn][1]);
    }
  }
  if (p.want_b && blockIdx.z == 0 && j0 + (t & 63) < p.M) atomicAdd(p.sums + (size_t)p.N * p.M + j0 + (t & 63), fb);
  if (p.want_a && blockIdx.y == 0) {
    if (sig_words) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (i0 + 4 * (t & 31) + r < p.N) atomicAdd(p.sums + (size_t)p.N * p.M + p.M + i0 + 4 * (t & 31) + r, fa4[r]);
    } else if (i0 + (t & 127) < p.N) {
      atomicAdd(p.sums + (size_t)p.N * p.M + p.M + i0 + (t & 127), fa);
    }
  }
}

template <typename T>
__global__ void forces_finalize_kernel(const double *__restrict__ sums, double scale, int64_t n, T *__restrict__ out) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) out[k] = (T)(sums[k] * scale);
}

int forces_sums(cudaStream_t stream, const nk_rbm_t &rbm, const int8_t *sigma, const void *theta, int is_tanh, const void *eloc,
                int32_t eloc_dtype, int64_t Ns, double mean, double *sums) {
  if (Ns == 0) return NK_OK;
  ForcesArgs a{};
  a.sigma = sigma;
  a.theta = theta;
  a.is_tanh = is_tanh;
  a.eloc = eloc;
  a.eloc_dtype = eloc_dtype;
  a.Ns = Ns;
  a.mean = mean;
  a.N = rbm.N;
  a.M = rbm.M;
  a.want_b = rbm.b != nullptr;
  a.want_a = rbm.a != nullptr;
  a.sums = sums;
  const int ty = (rbm.M + F_TJ - 1) / F_TJ, tz = (rbm.N + F_TI - 1) / F_TI;
  const int F_TS = rbm.dtype == NK_F32 ? ForcesTile<float>::TS : FD_TS;
  const int64_t n_batches = (Ns + F_TS - 1) / F_TS;
  int64_t gx = ((int64_t)num_sms() * 4 + ty * tz - 1) / (ty * tz);
  if (gx > n_batches) gx = n_batches;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)ty, (unsigned)tz);
  if (rbm.dtype == NK_F32)
    forces_kernel<float><<<grid, F_THREADS, 0, stream>>>(a);
  else
    forces_dmma_kernel<<<grid, F_THREADS, 0, stream>>>(a);  // fp64: FP64 tensor cores
  NK_LAUNCH_OK();
  return NK_OK;
}

int forces_finalize(cudaStream_t stream, const double *sums, double scale, int64_t n, void *out, int32_t dtype) {
  if (n == 0) return NK_OK;
  const int grid = (int)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024);
  if (dtype == NK_F32)
    forces_finalize_kernel<float><<<grid, 256, 0, stream>>>(sums, scale, n, reinterpret_cast<float *>(out));
  else
    forces_finalize_kernel<double><<<grid, 256, 0, stream>>>(sums, scale, n, reinterpret_cast<double *>(out));
  NK_LAUNCH_OK();
  return NK_OK;
}

}  // namespace nk
