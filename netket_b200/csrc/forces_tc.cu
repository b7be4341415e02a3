// Forces contraction on the 5th-generation tensor cores (fp32 parameters):
//     F_W[i, j] = sum_s sigma[s, i] x[s, j],   x[s, j] = tanh(theta[s, j]) (E_loc[s] - mean)
// (the vjp of netket/vqs/mc/mc_state/expect_forces.py:69-112 for nk.models.RBM; see forces.cu for the CUDA-core version,
// which stays the path for fp64 and for shapes outside this kernel's limits).
//
// It is a (N x n_s)(n_s x M) GEMM whose reduction dimension is the sample index.  A CTA owns a tile of NT hidden units and
// streams over its share of the samples 128 at a time; per block of samples its 512 threads write two K-major operand
// tiles straight into shared memory in the tcgen05 core-matrix layout:
//     A'[i, s] = sigma[s, i]      (exact in bf16; row N is all ones, so that D[N, j] = sum_s x[s, j] = F_b[j])
//     B'[j, s] = x[s, j]          split exactly into three bf16 parts (three tiles); column NT holds w[s] = E_loc[s] - mean,
//                                 so that D[i, NT] = sum_s sigma[s, i] w[s] = F_a[i])
// then one thread issues 3 x 8 tcgen05.mma (128 x (NT+16) x 16, fp32 accumulation in TMEM, accumulating over all the
// sample blocks of the CTA) and a tcgen05.commit releases the tiles for the next block.  After the last block the four
// epilogue warps read the accumulator (tcgen05.ld, TMEM lane = site) and add it to the double-precision sums.
#include "kernels.cuh"
#include "tc_common.cuh"

namespace nk {

constexpr int FT_KS = 128;                      // samples per block = K extent of one round of MMAs
constexpr int FT_SBO = (FT_KS / 8) * 128;       // bytes between 8-row groups of a K-major tile: 2048
constexpr int FT_A_BYTES = 16 * FT_SBO;         // 128 rows: 32 KB

struct FtcArgs {
  const int8_t *sigma;
  const float *theta;  // theta, or tanh(theta) if is_tanh
  int32_t is_tanh;
  const void *eloc;
  int32_t eloc_dtype;
  int64_t Ns;
  double mean;
  int32_t N, M, nt, NT, NTX;  // tiles over the hidden units, tile width, NTX = NT + 16 (MMA N extent)
  int32_t tmem_cols;
  int32_t want_b, want_a;
  double *sums;
};

constexpr int FT_THREADS = 512;  // 16 warps: one group of 8 samples each per block (more theta loads in flight)
#ifndef NK_FT_JB
#define NK_FT_JB 4
#endif
constexpr int FT_JB = NK_FT_JB;  // row iterations of the B' fill whose loads are issued together

__global__ void __launch_bounds__(FT_THREADS, 1) forces_tc_kernel(const __grid_constant__ FtcArgs p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *a_tile = smem;                                   // 128 x 128 bf16
  const uint32_t b_part_bytes = (uint32_t)(p.NTX / 8) * FT_SBO;   // one bf16 part of B'
  unsigned char *b_tiles = smem + FT_A_BYTES;                     // 3 parts
  uint64_t *bar_mma = reinterpret_cast<uint64_t *>(b_tiles + 3 * (size_t)b_part_bytes);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_mma + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int tile = blockIdx.x % p.nt;
  const int j0 = tile * p.NT;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar_mma)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = make_idesc_bf16_f32(p.NTX);

  const int64_t n_blocks = (p.Ns + FT_KS - 1) / FT_KS;
  const int ctas_per_tile = gridDim.x / p.nt;
  // Tile fill: a warp owns groups of 8 consecutive samples (= one 16-byte K-chunk of the core-matrix layout); lanes own
  // rows (sites for A', hidden units for B').  Global reads are coalesced along the row index (theta[s, j0 + lane + 32 c],
  // sigma[s, lane + 32 c]) and every (row, 8 samples) chunk is written with one 128-bit shared-memory store per bf16 part.
  const int lane = tid & 31;
  const bool sig_words = (p.N & 3) == 0 && (reinterpret_cast<uintptr_t>(p.sigma) & 3) == 0;
  uint32_t mma_phase = 0;
  int n_done = 0;
  for (int64_t blk = blockIdx.x / p.nt; blk < n_blocks; blk += ctas_per_tile, ++n_done) {
    // the previous round of MMAs must have finished reading the tiles
    if (n_done > 0) {
      mbar_wait_parity(bar_mma, mma_phase);
      mma_phase ^= 1u;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int g = warp; g < FT_KS / 8; g += FT_THREADS / 32) {
      const int64_t s0 = blk * FT_KS + 8 * g;
      // w[q] = E_loc[s0 + q] - mean for the 8 samples of the group (lane q < 8 loads, everyone gets all 8 by shuffle)
      float wl = 0.0f;
      if (lane < 8 && s0 + lane < p.Ns)
        wl = (float)((p.eloc_dtype == NK_F64 ? reinterpret_cast<const double *>(p.eloc)[s0 + lane]
                                             : (double)reinterpret_cast<const float *>(p.eloc)[s0 + lane]) - p.mean);
      float w[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) w[q] = __shfl_sync(0xffffffffu, wl, q);
      // ---- A'[i, s] = sigma[s, i] (+1 = 0x3F80, -1 = 0xBF80), row N = ones
      if (sig_words) {
        // rows of whole, aligned 32-bit words: lane l reads sites 4l .. 4l+3 of the 8 samples with 8 loads (bit 7 of a byte is
        // the sign) and writes the four rows' chunks; the ones row N and the zero rows above it are written by the next lanes
        uint32_t v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q)
          v[q] = (4 * lane < p.N && s0 + q < p.Ns) ? *reinterpret_cast<const uint32_t *>(p.sigma + (s0 + q) * p.N + 4 * lane) : 0u;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int i = 4 * lane + r;
          uint32_t pk[4] = {0u, 0u, 0u, 0u};
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            uint32_t h = 0u;
            if (s0 + q < p.Ns) {
              if (i < p.N)
                h = ((v[q] >> (8 * r + 7)) & 1u) ? 0xBF80u : 0x3F80u;
              else if (i == p.N)
                h = 0x3F80u;
            }
            pk[q >> 1] |= h << (16 * (q & 1));
          }
          *reinterpret_cast<uint4 *>(a_tile + (uint32_t)(i >> 3) * FT_SBO + (uint32_t)(i & 7) * 16u + (uint32_t)g * 128u) =
              make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      } else {
        for (int i = lane; i < 128; i += 32) {
          uint32_t pk[4] = {0u, 0u, 0u, 0u};
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            uint32_t h = 0u;
            if (s0 + q < p.Ns) {
              if (i < p.N)
                h = p.sigma[(s0 + q) * p.N + i] < 0 ? 0xBF80u : 0x3F80u;
              else if (i == p.N)
                h = 0x3F80u;
            }
            pk[q >> 1] |= h << (16 * (q & 1));
          }
          *reinterpret_cast<uint4 *>(a_tile + (uint32_t)(i >> 3) * FT_SBO + (uint32_t)(i & 7) * 16u + (uint32_t)g * 128u) =
              make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      // ---- B'[j, s] = tanh(theta[s, j0 + j]) * w[s] in three bf16 parts; column NT = w (tile 0 only: F_a)
      // (FT_JB row iterations at a time: their 8 FT_JB loads are all in flight before the arithmetic starts - with one load batch
      // per iteration the warp sat on the latency of HBM seven times per group: 25 % of the kernel's stall samples)
      for (int jb = lane; jb < p.NTX; jb += 32 * FT_JB) {
        float th[FT_JB][8];
#pragma unroll
        for (int u = 0; u < FT_JB; ++u) {
          const int j = jb + 32 * u;
          const bool in_tile = j < p.NT && j0 + j < p.M;
#pragma unroll
          for (int q = 0; q < 8; ++q) th[u][q] = (in_tile && s0 + q < p.Ns) ? p.theta[(s0 + q) * p.M + j0 + j] : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < FT_JB; ++u) {
          const int j = jb + 32 * u;
          if (j >= p.NTX) break;
          const bool in_tile = j < p.NT && j0 + j < p.M;
          // exact three-way bf16 split of two values at a time: cvt.rn.bf16x2.f32 packs a pair (element q in the low half,
          // q + 1 in the high half: the K-major order of the tile), the residuals are exact fp32 subtractions
          uint32_t p1[4], p2[4], p3[4];
#pragma unroll
          for (int q = 0; q < 8; q += 2) {
            float x0 = 0.0f, x1 = 0.0f;
            if (in_tile) {
              x0 = (p.is_tanh ? th[u][q] : tanhf(th[u][q])) * w[q];
              x1 = (p.is_tanh ? th[u][q + 1] : tanhf(th[u][q + 1])) * w[q + 1];
            } else if (j == p.NT && tile == 0) {
              x0 = w[q];
              x1 = w[q + 1];
            }
            const uint32_t h1 = pack_bf16x2_rn(x0, x1);
            const float r0 = x0 - __uint_as_float(h1 << 16), r1 = x1 - __uint_as_float(h1 & 0xffff0000u);
            const uint32_t h2 = pack_bf16x2_rn(r0, r1);
            const float s0r = r0 - __uint_as_float(h2 << 16), s1r = r1 - __uint_as_float(h2 & 0xffff0000u);
            p1[q >> 1] = h1;
            p2[q >> 1] = h2;
            p3[q >> 1] = pack_bf16x2_rn(s0r, s1r);
          }
          unsigned char *dst = b_tiles + (uint32_t)(j >> 3) * FT_SBO + (uint32_t)(j & 7) * 16u + (uint32_t)g * 128u;
          *reinterpret_cast<uint4 *>(dst) = make_uint4(p1[0], p1[1], p1[2], p1[3]);
          *reinterpret_cast<uint4 *>(dst + b_part_bytes) = make_uint4(p2[0], p2[1], p2[2], p2[3]);
          *reinterpret_cast<uint4 *>(dst + 2 * (size_t)b_part_bytes) = make_uint4(p3[0], p3[1], p3[2], p3[3]);
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_s = s32(a_tile), b_s = s32(b_tiles);
      for (int part = 0; part < 3; ++part) {
        for (int ks = 0; ks < FT_KS / 16; ++ks) {
          const uint64_t adesc = make_smem_desc(a_s + ks * 256, 128, FT_SBO);
          const uint64_t bdesc = make_smem_desc(b_s + part * b_part_bytes + ks * 256, 128, FT_SBO);
          const uint32_t acc = (n_done == 0 && part == 0 && ks == 0) ? 0u : 1u;
          asm volatile(
              "{\n.reg .pred q;\nsetp.ne.b32 q, %4, 0;\n"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n}\n" ::"r"(tmem_base),
              "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
              : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar_mma)) : "memory");
    }
  }
  // ---- epilogue: accumulator -> double sums.  TMEM lane = row i (site, or the ones row N); column = hidden unit (or NT: F_a)
  if (n_done > 0) {
    mbar_wait_parity(bar_mma, mma_phase);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 4) {
      const int i = tid;  // 0..127
      const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
      for (int c0 = 0; c0 < p.NTX; c0 += 32) {
        uint32_t r[32];
        const int width = min(32, p.NTX - c0);
        if (width == 32)
          tmem_ld32(lane_addr + c0, r);
        else
          tmem_ld16(lane_addr + c0, r);  // NTX is a multiple of 16
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int j = c0 + e;
          if (e >= width) continue;
          const double v = (double)__uint_as_float(r[e]);
          if (j < p.NT) {
            if (j0 + j < p.M) {
              if (i < p.N)
                atomicAdd(p.sums + (size_t)i * p.M + j0 + j, v);
              else if (i == p.N && p.want_b)
                atomicAdd(p.sums + (size_t)p.N * p.M + j0 + j, v);
            }
          } else if (j == p.NT && tile == 0 && p.want_a && i < p.N) {
            atomicAdd(p.sums + (size_t)p.N * p.M + p.M + i, v);
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

struct FtcGeom {
  int nt, NT, NTX, tmem_cols;
  size_t smem;
};

static bool ftc_geometry(const nk_rbm_t &rbm, FtcGeom *g) {
  if (rbm.dtype != NK_F32 || rbm.N > 127 || rbm.M < 1) return false;  // row N of the A' tile is the ones row
  for (int nt = 1; nt <= 64; ++nt) {
    const int per = (rbm.M + nt - 1) / nt;
    const int NT = (per + 15) & ~15;
    const int NTX = NT + 16;
    if (NTX > 256) continue;
    const size_t smem = (size_t)FT_A_BYTES + 3 * (size_t)(NTX / 8) * FT_SBO + 64;
    if (smem > 220 * 1024) continue;
    g->nt = nt;
    g->NT = NT;
    g->NTX = NTX;
    g->tmem_cols = NTX <= 32 ? 32 : (NTX <= 64 ? 64 : (NTX <= 128 ? 128 : 256));
    g->smem = smem;
    return true;
  }
  return false;
}

bool forces_tc_supported(const nk_rbm_t &rbm) {
  FtcGeom g;
  return ftc_geometry(rbm, &g);
}

int forces_tc_sums(cudaStream_t stream, const nk_rbm_t &rbm, const int8_t *sigma, const void *theta, int is_tanh, const void *eloc,
                   int32_t eloc_dtype, int64_t Ns, double mean, double *sums) {
  if (Ns == 0) return NK_OK;
  FtcGeom g;
  if (!ftc_geometry(rbm, &g)) {
    set_error("forces_tc: unsupported shape");
    return NK_EUNSUPPORTED;
  }
  FtcArgs a{};
  a.sigma = sigma;
  a.theta = reinterpret_cast<const float *>(theta);
  a.is_tanh = is_tanh;
  a.eloc = eloc;
  a.eloc_dtype = eloc_dtype;
  a.Ns = Ns;
  a.mean = mean;
  a.N = rbm.N;
  a.M = rbm.M;
  a.nt = g.nt;
  a.NT = g.NT;
  a.NTX = g.NTX;
  a.tmem_cols = g.tmem_cols;
  a.want_b = rbm.b != nullptr;
  a.want_a = rbm.a != nullptr;
  a.sums = sums;
  NK_CUDA_OK(cudaFuncSetAttribute(forces_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
  const int64_t n_blocks = (Ns + FT_KS - 1) / FT_KS;
  int64_t ctas = (int64_t)(num_sms() / g.nt) * g.nt;
  if (ctas < g.nt) ctas = g.nt;
  if (ctas > n_blocks * g.nt) ctas = n_blocks * g.nt;
  forces_tc_kernel<<<(int)ctas, FT_THREADS, g.smem, stream>>>(a);
  NK_LAUNCH_OK();
  return NK_OK;
}

}  // namespace nk
