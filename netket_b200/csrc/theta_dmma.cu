// theta[B, M] = sigma[B, N] W[N, M] + b in fp64 on the FP64 tensor cores (DMMA, mma.sync.m8n8k4.f64).
// (nn.Dense of netket/models/rbm.py:59-67 for param_dtype=float64, evaluated at `_reset`, netket/sampler/metropolis.py:399-403.)
//
// A CTA (8 warps) owns a tile of 64 hidden units: the W tile (N x 64 doubles, rows padded to 68 so that the B-fragment
// reads are bank-conflict free) is staged once in shared memory and reused for every block of 128 configurations the
// CTA streams over.  sigma (int8, +-1: exact in fp64) is staged per block and converted on the fly into A fragments.
// A warp computes 16 configurations x 64 hidden units: 2 x 8 m8n8k4 tiles, 32 accumulators per thread, bias folded
// into the initial accumulators; per K-step of 4 sites it issues 16 DMMAs for 10 fragment loads.
#include "kernels.cuh"

namespace nk {

constexpr int DM_TJ = 64, DM_ROWS = 128, DM_WSTRIDE = 68;

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) theta_dmma_kernel(const int8_t *__restrict__ sigma, const double *__restrict__ W,
                                                         const double *__restrict__ bias, double *__restrict__ theta, int64_t B, int N,
                                                         int M, int kpad, int sstride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *Ws = reinterpret_cast<double *>(smem_raw);                                   // [kpad][DM_WSTRIDE]
  int8_t *sg = reinterpret_cast<int8_t *>(smem_raw + (size_t)kpad * DM_WSTRIDE * 8);   // [DM_ROWS][sstride]
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int j0 = blockIdx.y * DM_TJ;
  for (int e = t; e < kpad * DM_TJ; e += 256) {
    const int k = e / DM_TJ, c = e - k * DM_TJ;
    Ws[k * DM_WSTRIDE + c] = (k < N && j0 + c < M) ? W[(size_t)k * M + j0 + c] : 0.0;
  }
  const int fr = lane >> 2, fc = lane & 3;  // fragment row / column of this thread
  const bool sig_words = (N & 3) == 0 && N <= 128 && (sstride & 3) == 0 && (reinterpret_cast<uintptr_t>(sigma) & 3) == 0;
  const int64_t n_blocks = (B + DM_ROWS - 1) / DM_ROWS;
  for (int64_t rb = blockIdx.x; rb < n_blocks; rb += gridDim.x) {
    __syncthreads();  // W tile ready / previous block's sigma consumed
    const int64_t row0 = rb * DM_ROWS;
    if (sig_words) {
      // rows of whole, aligned 32-bit words (N = kpad <= 128): a warp loads one row per instruction, DM_ROWS / 8 independent loads per
      // thread, all in flight together (the byte loop below issues ~50 dependent-address loads per thread and block)
      const int cw = t & 31, nw = N >> 2;
      uint32_t v[DM_ROWS / 8];
#pragma unroll
      for (int i = 0; i < DM_ROWS / 8; ++i) {
        const int r = (t >> 5) + 8 * i;
        v[i] = (cw < nw && row0 + r < B) ? *reinterpret_cast<const uint32_t *>(sigma + (row0 + r) * N + 4 * cw) : 0u;
      }
      if (cw < nw) {
#pragma unroll
        for (int i = 0; i < DM_ROWS / 8; ++i) *reinterpret_cast<uint32_t *>(sg + ((t >> 5) + 8 * i) * sstride + 4 * cw) = v[i];
      }
    } else {
      for (int e = t; e < DM_ROWS * kpad; e += 256) {
        const int r = e / kpad, k = e - r * kpad;
        sg[r * sstride + k] = (k < N && row0 + r < B) ? sigma[(row0 + r) * N + k] : (int8_t)0;
      }
    }
    __syncthreads();
    double acc[2][8][2];
#pragma unroll
    for (int rn = 0; rn < 8; ++rn) {
      const int j = j0 + 8 * rn + 2 * fc;
      const double b0 = (bias != nullptr && j < M) ? bias[j] : 0.0, b1 = (bias != nullptr && j + 1 < M) ? bias[j + 1] : 0.0;
#pragma unroll
      for (int rm = 0; rm < 2; ++rm) {
        acc[rm][rn][0] = b0;
        acc[rm][rn][1] = b1;
      }
    }
    const int8_t *sw = sg + (warp * 16 + fr) * sstride + fc;
    const double *wp = Ws + fc * DM_WSTRIDE + fr;
    for (int ks = 0; ks < kpad; ks += 4) {
      const double a0 = (double)sw[ks], a1 = (double)sw[8 * sstride + ks];
      double bf[8];
#pragma unroll
      for (int rn = 0; rn < 8; ++rn) bf[rn] = wp[ks * DM_WSTRIDE + 8 * rn];
#pragma unroll
      for (int rn = 0; rn < 8; ++rn) {
        dmma884(acc[0][rn][0], acc[0][rn][1], a0, bf[rn]);
        dmma884(acc[1][rn][0], acc[1][rn][1], a1, bf[rn]);
      }
    }
#pragma unroll
    for (int rm = 0; rm < 2; ++rm) {
      const int64_t row = row0 + warp * 16 + 8 * rm + fr;
      if (row >= B) continue;
      double *out = theta + row * M;
#pragma unroll
      for (int rn = 0; rn < 8; ++rn) {
        const int j = j0 + 8 * rn + 2 * fc;
        if (j + 1 < M && (M & 1) == 0) {
          *reinterpret_cast<double2 *>(out + j) = make_double2(acc[rm][rn][0], acc[rm][rn][1]);
        } else {
          if (j < M) out[j] = acc[rm][rn][0];
          if (j + 1 < M) out[j + 1] = acc[rm][rn][1];
        }
      }
    }
  }
}

bool theta_dmma_supported(const nk_rbm_t &rbm) {
  if (rbm.dtype != NK_F64) return false;
  const int kpad = (rbm.N + 3) & ~3;
  return (size_t)kpad * DM_WSTRIDE * 8 + (size_t)DM_ROWS * (kpad + 8) <= 200 * 1024;
}

int theta_dmma(cudaStream_t stream, const nk_rbm_t &rbm, const int8_t *sigma, int64_t B, void *theta_out) {
  const int kpad = (rbm.N + 3) & ~3;
  int sstride = kpad;  // words per row odd: the 8 rows of an A fragment fall into 8 different banks
  if (((sstride / 4) & 1) == 0) sstride += 4;
  const size_t smem = (size_t)kpad * DM_WSTRIDE * 8 + (size_t)DM_ROWS * sstride;
  NK_CUDA_OK(cudaFuncSetAttribute(theta_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int ty = (rbm.M + DM_TJ - 1) / DM_TJ;
  const int64_t n_blocks = (B + DM_ROWS - 1) / DM_ROWS;
  int occ = 1;
  NK_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, theta_dmma_kernel, 256, smem));
  if (occ < 1) occ = 1;
  int64_t gx = ((int64_t)num_sms() * occ + ty - 1) / ty;
  if (gx > n_blocks) gx = n_blocks;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)ty, 1);
  theta_dmma_kernel<<<grid, 256, smem, stream>>>(sigma, reinterpret_cast<const double *>(rbm.W), reinterpret_cast<const double *>(rbm.b),
                                                 reinterpret_cast<double *>(theta_out), B, rbm.N, rbm.M, kpad, sstride);
  NK_LAUNCH_OK();
  return NK_OK;
}

}  // namespace nk
