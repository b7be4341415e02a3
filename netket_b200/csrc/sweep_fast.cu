// Product-form fast path of the Metropolis sweep (fp32, LocalRule, tanh table resident in shared memory),
// optionally fused with the transverse-field-Ising local energy.
//
// Replaces the hot loop of netket/sampler/metropolis.py:427-462 (+ rules/local.py:40-49) and, when fused,
// netket/vqs/mc/kernels.py:62-71 with netket/operator/_ising/jax.py:125-165.
//
// Math.  With t_j = tanh(theta_j), tau_ij = tanh(2 W_ij), nu = -sigma_i:
//     cosh(theta_j + 2 nu W_ij) / cosh(theta_j) = cosh(2 W_ij) * (1 + nu tau_ij t_j)
// so a single-flip log-ratio is
//     delta_i = 2 nu a_i + sum_j log cosh(2 W_ij) + log prod_j (1 + nu tau_ij t_j).
// Each hidden unit is carried as an *unnormalised* pair (C_j, S_j) ~ (cosh theta_j, sinh theta_j):
//     proposal:  C'_j = C_j + nu tau_ij S_j            (1 FFMA)   and   prod_j C'_j   (1 FMUL)
//     accept:    S'_j = S_j + nu tau_ij C_j, C_j <- C'_j
// i.e. no transcendental per (proposal, hidden unit): one lg2 per lane per proposal, one division per hidden unit
// every `renorm` proposals (the period is chosen from max|tau| so that no lane product can leave the fp32 range).
// One warp owns one chain for the whole call; lanes own hidden units (packed as float2 -> FFMA2/FMUL2);
// W is brought in once per CTA by TMA bulk copies and turned into the tau table in place.
// The binding resource is the shared-memory read of one tau row (M floats) per proposal per chain.
//
// Validity: max|tau| <= TAU_LIMIT (|W| < ~0.97); otherwise the kernel raises a device flag and the theta-form
// generic kernel, always enqueued behind it, does the work (no host round trip).
#include "kernels.cuh"

namespace nk {

typedef unsigned long long u64;

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<u64 *>(&a)), "l"(*reinterpret_cast<u64 *>(&b)),
      "l"(*reinterpret_cast<u64 *>(&c)));
  return *reinterpret_cast<float2 *>(&d);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<u64 *>(&a)), "l"(*reinterpret_cast<u64 *>(&b)));
  return *reinterpret_cast<float2 *>(&d);
}
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }

// ---- mbarrier / TMA bulk copy (global -> shared), PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

constexpr int FAST_WARPS = 24;
constexpr int FAST_THREADS = FAST_WARPS * 32;
constexpr float TAU_LIMIT = 0.96f;

template <int NFULL, int TAIL>
struct Lanes {
  static constexpr int NE = 4 * NFULL + TAIL;        // hidden units per lane
  static constexpr int NP = (NE + 1) / 2;            // float2 pairs per lane
  static constexpr int MP = 128 * NFULL + 32 * TAIL; // padded row length of the tau table (floats)
  // hidden-unit index of element e of this lane (may be >= M: padding)
  static __device__ __forceinline__ int unit(int e, int lane) {
    return e < 4 * NFULL ? 128 * (e >> 2) + 4 * lane + (e & 3) : 128 * NFULL + TAIL * lane + (e - 4 * NFULL);
  }
  static __device__ __forceinline__ void load_row(const float *row, int lane, float2 (&t2)[NP]) {
#pragma unroll
    for (int q = 0; q < NFULL; ++q) {
      const float4 v = *reinterpret_cast<const float4 *>(row + 128 * q + 4 * lane);
      t2[2 * q] = make_float2(v.x, v.y);
      t2[2 * q + 1] = make_float2(v.z, v.w);
    }
    if (TAIL == 1) t2[2 * NFULL] = make_float2(row[128 * NFULL + lane], 0.0f);
    if (TAIL == 2) t2[2 * NFULL] = *reinterpret_cast<const float2 *>(row + 128 * NFULL + 2 * lane);
  }
};

struct FastSmem {
  float *tau;     // [N][MP]
  float *lcrow;   // [N]  sum_j log cosh(2 W_ij)
  float *a2;      // [N]  2 a_i (0 without visible bias)
  uint8_t *edges; // [E][2]
  uint64_t *bar;
  float *red;     // [32]
};

__host__ __device__ inline size_t fast_smem_bytes(int N, int MP, int E) {
  size_t s = (size_t)N * MP * 4;
  s += (size_t)N * 4 * 2;
  s += ((size_t)2 * E + 15) & ~(size_t)15;
  s += 16 + 32 * 4;
  return s;
}

template <int NFULL, int TAIL>
__global__ void __launch_bounds__(FAST_THREADS, 1)
    sweep_fast_kernel(const __grid_constant__ SweepKernelArgs p, const float *__restrict__ theta_ws, int *__restrict__ flags) {
  using LM = Lanes<NFULL, TAIL>;
  constexpr int NP = LM::NP, NE = LM::NE, MP = LM::MP;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int N = p.rbm.N, M = p.rbm.M, E = p.eloc_kind == 1 ? p.ising.n_edges : 0;
  FastSmem sm;
  sm.tau = reinterpret_cast<float *>(smem_raw);
  sm.lcrow = sm.tau + (size_t)N * MP;
  sm.a2 = sm.lcrow + N;
  sm.edges = reinterpret_cast<uint8_t *>(sm.a2 + N);
  sm.bar = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(sm.edges) + (((size_t)2 * E + 15) & ~(size_t)15));
  sm.red = reinterpret_cast<float *>(sm.bar + 2);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float *W = reinterpret_cast<const float *>(p.rbm.W);
  const float *avis = reinterpret_cast<const float *>(p.rbm.a);

  // ---------------- stage W into shared memory with TMA bulk copies (one per row: rows are padded to MP floats)
  if (tid == 0) {
    mbar_init(sm.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(sm.bar, (uint32_t)((size_t)N * M * 4));
    for (int i = 0; i < N; ++i) tma_bulk_g2s(sm.tau + (size_t)i * MP, W + (size_t)i * M, (uint32_t)(M * 4), sm.bar);
  }
  for (int i = tid; i < N; i += FAST_THREADS) sm.a2[i] = avis != nullptr ? 2.0f * avis[i] : 0.0f;
  for (int e = tid; e < 2 * E; e += FAST_THREADS) sm.edges[e] = (uint8_t)p.ising.edges[e];
  mbar_wait(sm.bar, 0);
  // ---------------- W -> tau = tanh(2W) in place; row constants; max|tau|
  float tmax = 0.0f;
  for (int i = warp; i < N; i += FAST_WARPS) {
    float *row = sm.tau + (size_t)i * MP;
    float lc = 0.0f;
    for (int j = lane; j < MP; j += 32) {
      float tv = 0.0f;
      if (j < M) {
        const float w2 = 2.0f * row[j];
        tv = tanhf(w2);
        lc += lncosh(w2);
        tmax = fmaxf(tmax, fabsf(tv));
      }
      row[j] = tv;
    }
    lc = warp_sum(lc);
    if (lane == 0) sm.lcrow[i] = lc;
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, m));
  if (lane == 0) sm.red[warp] = tmax;
  __syncthreads();
  tmax = 0.0f;
  for (int w = 0; w < FAST_WARPS; ++w) tmax = fmaxf(tmax, sm.red[w]);
  // renormalisation period: NE * (r + 1) * max(log2(1+tmax), -log2(1-tmax)) must stay below the fp32 exponent range
  int renorm = 0;
  if (tmax <= TAU_LIMIT) {
    const float per = (float)NE * fmaxf(log2f(1.0f + tmax), -log2f(1.0f - tmax));
    renorm = 32;
    while (renorm >= 1 && (float)(renorm + 1) * per > 120.0f) renorm >>= 1;
  }
  if (renorm < 1) {  // weights too large for the product form: hand over to the generic kernel queued behind us
    if (blockIdx.x == 0 && tid == 0) flags[0] = 1;
    return;
  }
  const int rmask = renorm - 1;

  const float pw = (float)p.machine_pow;
  const float inv_pw = pw > 0.0f ? 1.0f / pw : 0.0f;
  const int64_t T_total = (int64_t)(p.n_discard + p.chain_length) * p.sweep_size;
  const float LN2 = 0.69314718055994530942f;
  const float hh = (float)p.ising.h, JJ = (float)p.ising.J;

  for (int64_t chain = (int64_t)blockIdx.x * FAST_WARPS + warp; chain < p.B; chain += (int64_t)gridDim.x * FAST_WARPS) {
    // ---- sigma as a bit mask (bit = 1 <=> sigma = -1), replicated in every lane
    uint32_t sb[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int idx = 32 * w + lane;
      const bool neg = idx < N && p.sigma[chain * N + idx] < 0;
      sb[w] = __ballot_sync(0xffffffffu, neg);
    }
    // ---- theta (from the GEMM) -> (C, S) = (1, tanh theta); logpsi
    float2 C2[NP], S2[NP];
    float lc = 0.0f;
    {
      const float *th = theta_ws + (size_t)chain * M;
#pragma unroll
      for (int e = 0; e < 2 * NP; ++e) {
        const int j = e < NE ? LM::unit(e, lane) : M;
        float tv = 0.0f;
        if (j < M) {
          const float x = th[j];
          tv = tanhf(x);
          lc += lncosh(x);
        }
        if (e & 1)
          S2[e >> 1].y = tv;
        else
          S2[e >> 1].x = tv;
      }
#pragma unroll
      for (int q = 0; q < NP; ++q) C2[q] = make_float2(1.0f, 1.0f);
    }
    for (int i = lane; i < N; i += 32) {
      const float s = ((sb[i >> 5] >> (i & 31)) & 1u) ? -1.0f : 1.0f;
      lc = fmaf(0.5f * sm.a2[i], s, lc);
    }
    float logpsi = warp_sum(lc);
    float R = 0.0f;  // log2 prod_j C_j
    int64_t nacc = 0;
    int in_sweep = 0;
    int64_t sweep_idx = 0;
    const uint64_t gchain = p.chain_offset + (uint64_t)chain;

    auto renormalise = [&]() {
#pragma unroll
      for (int q = 0; q < NP; ++q) {
        S2[q].x = __fdividef(S2[q].x, C2[q].x);
        S2[q].y = __fdividef(S2[q].y, C2[q].y);
        C2[q] = make_float2(1.0f, 1.0f);
      }
      R = 0.0f;
    };
    // log2 prod_j (C_j + nu tau_ij S_j), reduced over the warp
    auto row_log2 = [&](const float2(&t2)[NP], bool nu_pos) -> float {
      float2 Pa = make_float2(1.0f, 1.0f), Pb = make_float2(1.0f, 1.0f);
      if (nu_pos) {
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          const float2 c = ffma2(t2[q], S2[q], C2[q]);
          if (q & 1)
            Pb = fmul2(Pb, c);
          else
            Pa = fmul2(Pa, c);
        }
      } else {
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          const float2 c = ffma2(neg2(t2[q]), S2[q], C2[q]);
          if (q & 1)
            Pb = fmul2(Pb, c);
          else
            Pa = fmul2(Pa, c);
        }
      }
      const float2 P = fmul2(Pa, Pb);
      return warp_sum(__log2f(P.x * P.y));
    };

    for (int64_t tt = 0; tt < T_total; tt += 32) {
      // ---- 32 proposals' worth of randomness, one Philox call per lane
      int site_l = 0;
      float thr_l = 0.0f;
      if (tt + lane < T_total) {
        uint32_t w0;
        float u;
        if (p.stream_w0 != nullptr) {
          w0 = p.stream_w0[(tt + lane) * p.B + chain];
          u = reinterpret_cast<const float *>(p.stream_u)[(tt + lane) * p.B + chain];
        } else {
          const uint4 w = philox_words(p.seed, p.t0 + (uint64_t)(tt + lane), gchain, STREAM_STEP);
          w0 = w.x;
          u = uniform_from_words<float>(w);
        }
        site_l = (int)__umulhi(w0, (uint32_t)N);
        // accept = u < exp(pw * delta)  <=>  log(u) / pw < delta      (metropolis.py:444-450)
        thr_l = pw > 0.0f ? logf(u) * inv_pw : -INFINITY;
      }
      const int nb = (int)min((int64_t)32, T_total - tt);
      for (int k = 0; k < nb; ++k) {
        if ((k & rmask) == 0) renormalise();
        const int i = __shfl_sync(0xffffffffu, site_l, k);
        const float thr = __shfl_sync(0xffffffffu, thr_l, k);
        const uint32_t word = (i < 64) ? ((i < 32) ? sb[0] : sb[1]) : ((i < 96) ? sb[2] : sb[3]);
        const bool neg = (word >> (i & 31)) & 1u;  // sigma_i = -1  =>  nu = +1
        float2 t2[NP];
        LM::load_row(sm.tau + (size_t)i * MP, lane, t2);
        const float Rp = row_log2(t2, neg);
        const float nu = neg ? 1.0f : -1.0f;
        const float delta = fmaf(LN2, Rp - R, fmaf(nu, sm.a2[i], sm.lcrow[i]));
        if (thr < delta) {
          if (neg) {
#pragma unroll
            for (int q = 0; q < NP; ++q) {
              const float2 cn = ffma2(t2[q], S2[q], C2[q]);
              S2[q] = ffma2(t2[q], C2[q], S2[q]);
              C2[q] = cn;
            }
          } else {
#pragma unroll
            for (int q = 0; q < NP; ++q) {
              const float2 nt = neg2(t2[q]);
              const float2 cn = ffma2(nt, S2[q], C2[q]);
              S2[q] = ffma2(nt, C2[q], S2[q]);
              C2[q] = cn;
            }
          }
          const uint32_t bit = 1u << (i & 31);
          sb[0] ^= (i < 32) ? bit : 0u;
          sb[1] ^= (i >= 32 && i < 64) ? bit : 0u;
          sb[2] ^= (i >= 64 && i < 96) ? bit : 0u;
          sb[3] ^= (i >= 96) ? bit : 0u;
          R = Rp;
          logpsi += delta;
          ++nacc;
        }
        if (++in_sweep == p.sweep_size) {
          in_sweep = 0;
          const int64_t sw = sweep_idx - p.n_discard;
          ++sweep_idx;
          if (sw >= 0) {
            const int64_t o = chain * p.chain_length + sw;
            if (p.samples_out != nullptr)
              for (int n = lane; n < N; n += 32) p.samples_out[o * N + n] = ((sb[n >> 5] >> (n & 31)) & 1u) ? (int8_t)-1 : (int8_t)1;
            if (p.logp_out != nullptr && lane == 0) reinterpret_cast<float *>(p.logp_out)[o] = pw * logpsi;
            if (p.eloc_kind == 1) {
              // E_loc = J sum_<ij> s_i s_j - h sum_i exp(delta_i)
              renormalise();
              int zz = 0;
              for (int e = lane; e < E; e += 32) {
                const int a = sm.edges[2 * e], b = sm.edges[2 * e + 1];
                const uint32_t x = ((sb[a >> 5] >> (a & 31)) ^ (sb[b >> 5] >> (b & 31))) & 1u;
                zz += 1 - 2 * (int)x;
              }
              zz = __reduce_add_sync(0xffffffffu, zz);
              float off = 0.0f;
              if (hh != 0.0f) {
                for (int s = 0; s < N; ++s) {
                  const bool ng = (sb[s >> 5] >> (s & 31)) & 1u;
                  float2 r2[NP];
                  LM::load_row(sm.tau + (size_t)s * MP, lane, r2);
                  const float Rs = row_log2(r2, ng);
                  off += __expf(fmaf(LN2, Rs, fmaf(ng ? 1.0f : -1.0f, sm.a2[s], sm.lcrow[s])));
                }
              }
              const float e_loc = JJ * (float)zz - hh * off;
              if (lane == 0) store_as<float>(p.eloc_out, o, e_loc, p.eloc_dtype);
            }
          }
        }
      }
    }
    // ---- write the chain state back
    for (int n = lane; n < N; n += 32) p.sigma[chain * N + n] = ((sb[n >> 5] >> (n & 31)) & 1u) ? (int8_t)-1 : (int8_t)1;
    if (lane == 0) {
      reinterpret_cast<float *>(p.log_prob)[chain] = pw * logpsi;
      p.n_accepted[chain] += nacc;
    }
  }
}

// ------------------------------------------------------------------------------------------ host side
struct FastShape {
  int nfull, tail, mp;
};

static bool fast_shape(int M, FastShape *fs) {
  if (M % 4 != 0 || M < 4) return false;
  int nfull = M / 128, rem = M % 128, tail = 0;
  if (rem == 0)
    tail = 0;
  else if (rem <= 32)
    tail = 1;
  else if (rem <= 64)
    tail = 2;
  else {
    nfull += 1;
    tail = 0;
  }
  if (nfull > 4 || (nfull == 4 && tail != 0)) return false;
  fs->nfull = nfull;
  fs->tail = tail;
  fs->mp = 128 * nfull + 32 * tail;
  return true;
}

bool sweep_fast_supported(const SweepKernelArgs &a) {
  FastShape fs;
  if (a.rbm.dtype != NK_F32 || a.rule != NK_RULE_LOCAL) return false;
  if (a.rbm.N > 128 || !fast_shape(a.rbm.M, &fs)) return false;
  if (a.eloc_kind == 2) return false;
  if (a.eloc_kind == 1 && a.rbm.N > 256) return false;
  const int E = a.eloc_kind == 1 ? a.ising.n_edges : 0;
  if (fast_smem_bytes(a.rbm.N, fs.mp, E) > 227 * 1024) return false;
  if ((size_t)a.rbm.N * a.rbm.M * 4 >= (1u << 20)) return false;  // mbarrier tx-count range
  return true;
}

template <int NFULL, int TAIL>
static int launch_fast(cudaStream_t stream, const SweepKernelArgs &a, const float *theta_ws, int *flags, size_t smem) {
  auto kern = sweep_fast_kernel<NFULL, TAIL>;
  NK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t need = (a.B + FAST_WARPS - 1) / FAST_WARPS;
  const int64_t cap = num_sms();
  kern<<<(int)(need < cap ? need : cap), FAST_THREADS, smem, stream>>>(a, theta_ws, flags);
  NK_LAUNCH_OK();
  return NK_OK;
}

int sweep_fast(cudaStream_t stream, const SweepKernelArgs &a, const float *theta_ws, int *flags) {
  FastShape fs;
  if (!fast_shape(a.rbm.M, &fs)) {
    set_error("sweep_fast: unsupported M=%d", a.rbm.M);
    return NK_EUNSUPPORTED;
  }
  const int E = a.eloc_kind == 1 ? a.ising.n_edges : 0;
  const size_t smem = fast_smem_bytes(a.rbm.N, fs.mp, E);
#define NK_FAST_CASE(NF, TL) \
  if (fs.nfull == NF && fs.tail == TL) return launch_fast<NF, TL>(stream, a, theta_ws, flags, smem);
  NK_FAST_CASE(0, 1)
  NK_FAST_CASE(0, 2)
  NK_FAST_CASE(1, 0)
  NK_FAST_CASE(1, 1)
  NK_FAST_CASE(1, 2)
  NK_FAST_CASE(2, 0)
  NK_FAST_CASE(2, 1)
  NK_FAST_CASE(2, 2)
  NK_FAST_CASE(3, 0)
  NK_FAST_CASE(3, 1)
  NK_FAST_CASE(3, 2)
  NK_FAST_CASE(4, 0)
#undef NK_FAST_CASE
  set_error("sweep_fast: no instantiation for M=%d", a.rbm.M);
  return NK_EUNSUPPORTED;
}

bool eloc_fast_supported(const nk_rbm_t &) { return false; }
int eloc_fast_ising(cudaStream_t, const nk_rbm_t &, const nk_ising_t &, const int8_t *, int64_t, void *, int32_t) {
  set_error("eloc_fast: not built");
  return NK_EUNSUPPORTED;
}

}  // namespace nk
