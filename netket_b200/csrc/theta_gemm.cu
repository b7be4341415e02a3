// theta[B, M] = sigma[B, N] W[N, M] + b  — the only dense contraction on the path
// (nn.Dense of netket/models/rbm.py:59-67, evaluated at `_reset`, netket/sampler/metropolis.py:399-403).
//
// fp32: 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).  The product is made EXACT in fp32 terms by
// splitting W into three bf16 parts, W = W1 + W2 + W3 (8+8+8 mantissa bits), while sigma in {+1,-1} is exact in bf16:
//     theta = sigma W1 + sigma W2 + sigma W3          (fp32 accumulation in TMEM, 3 x K/16 MMAs per tile)
// so the result differs from an fp32 FMA chain only by summation order.
//   * a prep kernel writes the three bf16 parts of W^T as ready-made shared-memory images (K-major, no-swizzle
//     core-matrix layout: 8 rows x 16 bytes per core matrix);
//   * each CTA (128 threads, persistent) bulk-copies its N-tile image once (cp.async.bulk + mbarrier), then per block of
//     128 configurations: threads convert sigma (int8) to bf16 straight into the A tile, one thread issues the MMAs
//     (UMMA 128 x NT x 16, cta_group::1), tcgen05.commit signals an mbarrier, and the four warps read their TMEM lane
//     quarter with tcgen05.ld, add the bias and store theta.
// fp64: FP64 tensor cores (DMMA), theta_dmma.cu.  Shapes outside both kernels' limits: the CUDA-core kernel rbm_logpsi_kernel.
#include "kernels.cuh"
#include "tc_common.cuh"

namespace nk {

int rbm_logpsi(cudaStream_t stream, const nk_rbm_t &rbm, const int8_t *sigma, int64_t B, void *out, void *theta_out);
bool theta_dmma_supported(const nk_rbm_t &rbm);
int theta_dmma(cudaStream_t stream, const nk_rbm_t &rbm, const int8_t *sigma, int64_t B, void *theta_out);

// ---------------------------------------------------------------------------------------------- geometry
struct TcGeom {
  int kpad;     // N rounded up to a multiple of 16
  int nt;       // number of N-tiles (1 or 2)
  int NT;       // tile width (multiple of 16, <= 256)
  int sbo;      // bytes between 8-row groups = (kpad / 8) * 128
  size_t part_bytes;   // one bf16 part of one N-tile image
  size_t img_bytes;    // 3 parts
  size_t a_bytes;      // A tile (128 rows)
  size_t smem_bytes;
  int tmem_cols, acc_cols;
};

static bool tc_geometry(const nk_rbm_t &rbm, TcGeom *g) {
  if (rbm.dtype != NK_F32 || rbm.N > 128 || rbm.M > 512 || rbm.M < 16) return false;
  g->kpad = (rbm.N + 15) & ~15;
  g->nt = rbm.M <= 256 ? 1 : 2;
  const int per = (rbm.M + g->nt - 1) / g->nt;
  g->NT = (per + 15) & ~15;
  g->sbo = (g->kpad / 8) * 128;
  g->part_bytes = (size_t)(g->NT / 8) * g->sbo;
  g->img_bytes = 3 * g->part_bytes;
  g->a_bytes = (size_t)16 * g->sbo;
  g->smem_bytes = g->img_bytes + 2 * g->a_bytes + 64;  // double-buffered A tile
  g->acc_cols = g->NT <= 32 ? 32 : (g->NT <= 64 ? 64 : (g->NT <= 128 ? 128 : 256));
  g->tmem_cols = 2 * g->acc_cols;  // two accumulators: MMA of block k+1 overlaps the epilogue of block k
  return g->smem_bytes <= 220 * 1024 && g->img_bytes < (1u << 20);
}

// ---------------------------------------------------------------------------------------------- prep: W -> 3 bf16 images
// image[tile][part][(n / 8) * sbo + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2],  n = column inside the tile, k = site
// `bias` != nullptr (only when kpad > N): row k = N of the image holds the bias, multiplied by a column of ones in the A tile, so
// that theta = sigma W + b comes out of the MMAs and the epilogue does no bias loads (they were ~45 % of its stall samples)
__global__ void theta_prep_kernel(const float *__restrict__ W, const float *__restrict__ bias, int N, int M, int kpad, int nt, int NT, int sbo,
                                  uint16_t *__restrict__ img) {
  const int total = nt * NT * kpad;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int k = idx % kpad;
    const int ncol = (idx / kpad) % NT;
    const int tile = idx / (kpad * NT);
    const int j = tile * NT + ncol;
    float w = (k < N && j < M) ? W[(size_t)k * M + j] : 0.0f;
    if (bias != nullptr && k == N && j < M) w = bias[j];
    const uint16_t h1 = f32_to_bf16_rn(w);
    const float r1 = w - bf16_to_f32(h1);
    const uint16_t h2 = f32_to_bf16_rn(r1);
    const float r2 = r1 - bf16_to_f32(h2);
    const uint16_t h3 = f32_to_bf16_rn(r2);
    const size_t part_elems = (size_t)(NT / 8) * sbo / 2;
    const size_t off = (size_t)(ncol / 8) * (sbo / 2) + (size_t)(k / 8) * 64 + (size_t)(ncol % 8) * 8 + (k % 8);
    uint16_t *base = img + (size_t)tile * 3 * part_elems;
    base[off] = h1;
    base[part_elems + off] = h2;
    base[2 * part_elems + off] = h3;
  }
}

// ---------------------------------------------------------------------------------------------- GEMM
struct TcArgs {
  const int8_t *sigma;
  const uint16_t *img;
  const float *bias;
  float *theta;
  int64_t B;
  int N, M, kpad, nt, NT, sbo, tmem_cols, acc_cols, vec_ok;
  int ones_col;  // 1: A carries a column of ones at k = N (the bias sits in the W image), `bias` is NULL
  uint32_t part_bytes, img_bytes;
};

#ifndef NK_TC_THREADS
#define NK_TC_THREADS 512
#endif
constexpr int TC_THREADS = NK_TC_THREADS;  // TC_PARTS warps per TMEM lane quarter: they split the A-tile chunks and the accumulator's columns
constexpr int TC_PARTS = TC_THREADS / 128;

__global__ void __launch_bounds__(TC_THREADS, 1) theta_tc_kernel(const __grid_constant__ TcArgs p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *b_img = smem;                          // 3 parts, K-major core-matrix layout
  unsigned char *a_tiles = smem + p.img_bytes;          // 2 x (128 x kpad bf16), same layout
  const size_t a_bytes = (size_t)16 * p.sbo;
  uint64_t *bars = reinterpret_cast<uint64_t *>(a_tiles + 2 * a_bytes);
  uint64_t *bar_load = bars, *bar_mma = bars + 1;       // bar_mma[0..1]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row_t = tid & 127;   // row of the 128-row block this thread works on (= TMEM lane: warp % 4 selects the lane quarter)
  const int half_t = tid >> 7;   // which share of the K chunks (A tile) / of the columns (epilogue) it takes
  const int tile = blockIdx.x % p.nt;
  const int n0 = tile * p.NT;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar_load)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar_mma)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar_mma + 1)));
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

  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar_load)), "r"(p.img_bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(b_img)),
                 "l"(reinterpret_cast<const unsigned char *>(p.img) + (size_t)tile * p.img_bytes), "r"(p.img_bytes), "r"(s32(bar_load))
                 : "memory");
  }

  auto wait_bar = [&](uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    do {
      asm volatile("{\n.reg .pred q;\nmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\nselp.u32 %0, 1, 0, q;\n}\n"
                   : "=r"(done)
                   : "r"(s32(bar)), "r"(parity)
                   : "memory");
    } while (!done);
  };
  // A tile of row block rb into buffer `buf`: row = tid % 128 (the two thread halves alternate over the 16-byte K chunks), sigma int8 -> bf16 (+1 = 0x3F80, -1 = 0xBF80), zero beyond N / B
  auto build_a = [&](int64_t rb, int buf) {
    const int64_t row = rb * 128 + row_t;
    const int8_t *src = p.sigma + row * p.N;
    unsigned char *dst = a_tiles + buf * a_bytes + (size_t)(row_t >> 3) * p.sbo + (size_t)(row_t & 7) * 16;
    // rows of whole, aligned 32-bit words (row * N is then a multiple of 4): 4 spins per load
    const bool words = (p.N & 3) == 0 && (reinterpret_cast<uintptr_t>(p.sigma) & 3) == 0;
    for (int c = half_t; c < p.kpad / 8; c += TC_PARTS) {
      uint32_t w[4] = {0u, 0u, 0u, 0u};
      if (row < p.B) {
        if (words) {
#pragma unroll
          for (int hw = 0; hw < 2; ++hw) {
            const int k0 = 8 * c + 4 * hw;
            if (k0 < p.N) {
              const uint32_t v = *reinterpret_cast<const uint32_t *>(src + k0);  // bytes +1 = 0x01, -1 = 0xFF: bit 7 is the sign
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const uint32_t h = ((v >> (8 * e + 7)) & 1u) ? 0xBF80u : 0x3F80u;
                w[2 * hw + (e >> 1)] |= h << (16 * (e & 1));
              }
            }
          }
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int k = 8 * c + e;
            const uint32_t h = (k < p.N) ? (src[k] < 0 ? 0xBF80u : 0x3F80u) : 0u;
            w[e >> 1] |= h << (16 * (e & 1));
          }
        }
        if (p.ones_col && (p.N >> 3) == c) w[(p.N & 7) >> 1] |= 0x3F80u << (16 * (p.N & 1));  // 1.0 at k = N
      }
      *reinterpret_cast<uint4 *>(dst + (size_t)c * 128) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  };
  // instruction descriptor, kind::f16: D = F32 (bit 4), A = B = BF16 (bits 7, 10), both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  // 3 bf16 parts x (kpad / 16) K-steps into accumulator `buf`, then commit -> bar_mma[buf]   (one thread)
  auto issue_mma = [&](int buf) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t a_s = s32(a_tiles + buf * a_bytes), b_s = s32(b_img);
    const uint32_t d_t = tmem_base + (uint32_t)(buf * p.acc_cols);
    int first = 1;
    for (int part = 0; part < 3; ++part) {
      for (int ks = 0; ks < p.kpad / 16; ++ks) {
        const uint64_t adesc = make_smem_desc(a_s + ks * 256, 128, (uint32_t)p.sbo);
        const uint64_t bdesc = make_smem_desc(b_s + part * p.part_bytes + ks * 256, 128, (uint32_t)p.sbo);
        const uint32_t acc = first ? 0u : 1u;
        asm volatile(
            "{\n.reg .pred q;\nsetp.ne.b32 q, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n}\n" ::"r"(d_t),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
            : "memory");
        first = 0;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar_mma + buf)) : "memory");
  };

  const int64_t n_blocks = (p.B + 127) / 128;
  const int ctas_per_tile = gridDim.x / p.nt;
  const int64_t rb0 = blockIdx.x / p.nt;
  // prologue: first A tile while the W image is still in flight
  if (rb0 < n_blocks) build_a(rb0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core (async proxy)
  wait_bar(bar_load, 0);
  __syncthreads();
  if (tid == 0 && rb0 < n_blocks) issue_mma(0);

  int it = 0;
  for (int64_t rb = rb0; rb < n_blocks; rb += ctas_per_tile, ++it) {
    const int buf = it & 1;
    const int64_t rb_next = rb + ctas_per_tile;
    // ---- next block: build its A tile and start its MMAs (they run while this block's accumulator is drained)
    if (rb_next < n_blocks) {
      build_a(rb_next, buf ^ 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0 && rb_next < n_blocks) issue_mma(buf ^ 1);
    // ---- this block: wait for its accumulator, epilogue: warp w owns TMEM lanes 32w..32w+31 = rows 32w..32w+31
    wait_bar(bar_mma + buf, (uint32_t)((it >> 1) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
      const int64_t row = rb * 128 + row_t;
      float *out = p.theta + row * p.M;
      const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(buf * p.acc_cols);
      for (int c0 = 32 * half_t; c0 < p.NT; c0 += 32 * TC_PARTS) {
        uint32_t r[32];
        const int width = min(32, p.NT - c0);
        if (width == 32)
          tmem_ld32(lane_addr + c0, r);
        else
          tmem_ld16(lane_addr + c0, r);  // NT is a multiple of 16
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < p.B) {
          const int jbase = n0 + c0;
          if (p.vec_ok == 2 && jbase + width <= p.M) {
            // 256-bit stores (sm_100): this thread owns `width` consecutive floats of its row and writes whole 32-byte
            // sectors, so that no partial-sector writes reach L2
#pragma unroll
            for (int e = 0; e < 32; e += 8) {
              if (e < width) {
                float v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = __uint_as_float(r[e + q]) + (p.bias != nullptr ? p.bias[jbase + e + q] : 0.0f);
                asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(out + jbase + e), "f"(v[0]), "f"(v[1]),
                             "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                             : "memory");
              }
            }
          } else if (p.vec_ok && jbase + width <= p.M) {
            // 128-bit stores
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              if (e < width) {
                float4 v = make_float4(__uint_as_float(r[e]), __uint_as_float(r[e + 1]), __uint_as_float(r[e + 2]), __uint_as_float(r[e + 3]));
                if (p.bias != nullptr) {
                  const float4 bb = *reinterpret_cast<const float4 *>(p.bias + jbase + e);
                  v.x += bb.x;
                  v.y += bb.y;
                  v.z += bb.z;
                  v.w += bb.w;
                }
                *reinterpret_cast<float4 *>(out + jbase + e) = v;
              }
            }
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const int j = jbase + e;
              if (e < width && j < p.M) out[j] = __uint_as_float(r[e]) + (p.bias != nullptr ? p.bias[j] : 0.0f);
            }
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    // the next iteration's __syncthreads orders these TMEM reads before accumulator `buf` is overwritten
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------- host
int64_t theta_gemm_workspace_bytes(const nk_rbm_t &rbm, int64_t B) {
  TcGeom g;
  if (tc_geometry(rbm, &g)) return (int64_t)(((size_t)g.nt * g.img_bytes + 255) & ~(size_t)255);
  return (int64_t)((((size_t)B * (rbm.dtype == NK_F32 ? 4 : 8)) + 255) & ~(size_t)255);  // CUDA-core kernel: logpsi scratch
}

int theta_gemm(cudaStream_t stream, const nk_rbm_t &rbm, const int8_t *sigma, int64_t B, void *theta_out, void *workspace) {
  if (B == 0) return NK_OK;
  if (workspace == nullptr) {
    set_error("nk_theta_gemm: workspace is NULL");
    return NK_EINVAL;
  }
  TcGeom g;
  if (theta_dmma_supported(rbm)) return theta_dmma(stream, rbm, sigma, B, theta_out);  // fp64: DMMA
  if (!tc_geometry(rbm, &g)) return rbm_logpsi(stream, rbm, sigma, B, workspace, theta_out);
  uint16_t *img = reinterpret_cast<uint16_t *>(workspace);
  const bool fold_bias = rbm.b != nullptr && g.kpad > rbm.N;  // a spare K slot: the bias rides along as row N of W
  {
    const int total = g.nt * g.NT * g.kpad;
    theta_prep_kernel<<<(total + 255) / 256, 256, 0, stream>>>(reinterpret_cast<const float *>(rbm.W),
                                                               fold_bias ? reinterpret_cast<const float *>(rbm.b) : nullptr, rbm.N, rbm.M,
                                                               g.kpad, g.nt, g.NT, g.sbo, img);
    NK_LAUNCH_OK();
  }
  TcArgs a{};
  a.sigma = sigma;
  a.img = img;
  a.bias = fold_bias ? nullptr : reinterpret_cast<const float *>(rbm.b);
  a.ones_col = fold_bias ? 1 : 0;
  a.theta = reinterpret_cast<float *>(theta_out);
  a.B = B;
  a.N = rbm.N;
  a.M = rbm.M;
  a.kpad = g.kpad;
  a.nt = g.nt;
  a.NT = g.NT;
  a.sbo = g.sbo;
  a.tmem_cols = g.tmem_cols;
  a.acc_cols = g.acc_cols;
  a.vec_ok = (rbm.M % 4 == 0) && ((reinterpret_cast<uintptr_t>(theta_out) & 15) == 0) &&
             (rbm.b == nullptr || (reinterpret_cast<uintptr_t>(rbm.b) & 15) == 0);
  // 256-bit stores need 32-byte aligned rows, tiles and 16-column chunks (NT is a multiple of 16: chunk widths are 32 or 16)
  if (a.vec_ok && rbm.M % 8 == 0 && g.NT % 8 == 0 && (reinterpret_cast<uintptr_t>(theta_out) & 31) == 0) a.vec_ok = 2;
  a.part_bytes = (uint32_t)g.part_bytes;
  a.img_bytes = (uint32_t)g.img_bytes;
  NK_CUDA_OK(cudaFuncSetAttribute(theta_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes));
  const int64_t n_blocks = (B + 127) / 128;
  int64_t ctas = (int64_t)(num_sms() / g.nt) * g.nt;
  if (ctas > n_blocks * g.nt) ctas = n_blocks * g.nt;
  theta_tc_kernel<<<(int)ctas, TC_THREADS, g.smem_bytes, stream>>>(a);
  NK_LAUNCH_OK();
  return NK_OK;
}

}  // namespace nk
