// Product-form fast path of the Metropolis sweep (fp32, LocalRule, weight table resident in shared memory),
// optionally fused with the transverse-field-Ising local energy and the MC statistics of those energies.
//
// Replaces the hot loop of netket/sampler/metropolis.py:427-462 (+ rules/local.py:40-49) and, when fused,
// netket/vqs/mc/kernels.py:62-71 with netket/operator/_ising/jax.py:125-165 and the sums of
// netket/stats/mc_stats_old.py:87-196.
//
// Math ("exponential form").  Each hidden unit is carried as an unnormalised positive pair
//     (A_j, B_j)  proportional to  (exp(theta_j), exp(-theta_j)),     cosh(theta_j) ~ (A_j + B_j) / 2,
// and the weight table is G_ij = exp(-4 W_ij).  Flipping site i changes theta_j by 2 nu W_ij, nu = -sigma_i:
//     nu = +1:  cosh(theta'_j) ~ exp(2 W_ij) (A_j + B_j G_ij) / 2      accept:  B_j <- B_j G_ij
//     nu = -1:  cosh(theta'_j) ~ exp(2 W_ij) (A_j G_ij + B_j) / 2      accept:  A_j <- A_j G_ij
// so  delta_i = 2 nu a_i + 2 sum_j W_ij + log prod_j (A_j + B_j G_ij | A_j G_ij + B_j) - log prod_j (A_j + B_j).
// Per (proposal, hidden unit) that is one FFMA and one FMUL, no transcendental and - every term being positive -
// no cancellation, however saturated the unit is; the accept costs one FMUL.  One lg2 per lane per proposal turns the
// lane products into a fixed-point sum (REDUX); (A, B) are renormalised to A + B = 1 every `renorm` accepted moves, a
// period chosen from max|W| so that no lane product can leave the fp32 range.
// One warp owns one chain for the whole call; lanes own hidden units (packed as float2 -> FFMA2/FMUL2, an odd unit as a
// scalar); W is brought in once per CTA by TMA bulk copies and turned into the G table in place.
// The binding resource is the shared-memory read of one G row (M floats) per proposal per chain.
//
// Proposal loop (v7).  The 32 proposals of a batch are prepared by the 32 lanes in parallel (Philox, site, threshold,
// per-site constants) and left in a per-warp shared-memory record {row address, spin address, threshold - fix(x_i),
// fix(y_i)}; a proposal then starts with ONE broadcast LDS.128 instead of three shuffles, the chain's spins live as
// bytes in shared memory (one LDS.U8 to read, one STS.U8 on accept) and the whole accept test is integer:
//   fix(log2 ratio) = (Rp - R) + fix(x_i) +- fix(y_i),   x_i = log2e 2 sum_j W_ij,  y_i = log2e 2 a_i,
//   accept = u < exp(machine_pow * delta)  <=>  fix(log2(u) / machine_pow) < fix(log2 ratio)   (metropolis.py:444-450).
//
// Validity: 2 NP * (4 max|W| log2 e) <= 120 (|W| <~ 1.5 at 14 units per lane); otherwise the kernel raises a device flag and
// the kernels enqueued behind it do the work without a host round trip: the general product-form kernel in its wide mode
// (two logarithms per lane product, |W| <~ 3), then the theta-form kernel.
#include "fast_common.cuh"

namespace nk {

using namespace fast;

#ifndef NK_FAST_WARPS
#define NK_FAST_WARPS 28
#endif
constexpr int FAST_WARPS = NK_FAST_WARPS;
constexpr int FAST_THREADS = FAST_WARPS * 32;
constexpr int SIG_STRIDE = 128;          // spin bytes per warp (N <= 128)
constexpr int WSTAT = 12;                // doubles of statistics scratch per warp

__host__ __device__ inline size_t fast_smem_bytes(int N, int MP, int E) {
  size_t s = (size_t)N * MP * 4;            // G table
  s += (size_t)N * 16;                      // per-site constants {x = log2e 2 sum_j W_ij, y = log2e 2 a_i, fix(x), fix(y)}
  s += ((size_t)2 * E + 15) & ~(size_t)15;  // edges (uint8 pairs)
  s += (size_t)FAST_WARPS * 512;            // proposal records of the current batch, per warp
  s += (size_t)FAST_WARPS * SIG_STRIDE;     // spins of the warp's chain (bytes, 1 = spin down)
  s += (size_t)FAST_WARPS * WSTAT * 8;      // statistics scratch per warp
  s += 16 + 32 * 4;                         // mbarrier, reduction scratch
  return s;
}

__device__ __forceinline__ uint32_t sw4sel(const uint32_t (&w)[4], int i) {
  return (i < 2) ? ((i == 0) ? w[0] : w[1]) : ((i == 2) ? w[2] : w[3]);
}

template <int NFULL, int TAIL>
__global__ void __launch_bounds__(FAST_THREADS, 1)
    sweep_fast_kernel(const __grid_constant__ SweepKernelArgs p, const float *__restrict__ theta_ws, int *__restrict__ flags) {
  using LM = Lanes<NFULL, TAIL>;
  constexpr int NP2 = LM::NP2, NPA = LM::NPA, NE = LM::NE, MP = LM::MP;
  constexpr bool HAS_T = LM::HAS_T;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int N = p.rbm.N, M = p.rbm.M, E = p.eloc_kind == 1 ? p.ising.n_edges : 0;
  float *Gtab = reinterpret_cast<float *>(smem_raw);
  float4 *rctab = reinterpret_cast<float4 *>(Gtab + (size_t)N * MP);
  uint8_t *edges = reinterpret_cast<uint8_t *>(rctab + N);
  uint4 *rectab = reinterpret_cast<uint4 *>(edges + (((size_t)2 * E + 15) & ~(size_t)15));
  uint8_t *sigtab = reinterpret_cast<uint8_t *>(rectab + FAST_WARPS * 32);
  double *wstat = reinterpret_cast<double *>(sigtab + FAST_WARPS * SIG_STRIDE);
  uint64_t *bar = reinterpret_cast<uint64_t *>(wstat + FAST_WARPS * WSTAT);
  float *red = reinterpret_cast<float *>(bar + 2);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform by construction: lets ptxas use uniform control flow
  const float *W = reinterpret_cast<const float *>(p.rbm.W);
  const float *avis = reinterpret_cast<const float *>(p.rbm.a);
  const float LOG2E = 1.4426950408889634f, LN2 = 0.69314718055994530942f;

  // ---------------- stage W into shared memory with TMA bulk copies (one per row: rows are padded to MP floats)
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(bar, (uint32_t)((size_t)N * M * 4));
    for (int i = 0; i < N; ++i) tma_bulk_g2s(Gtab + (size_t)i * MP, W + (size_t)i * M, (uint32_t)(M * 4), bar);
  }
  for (int e = tid; e < 2 * E; e += FAST_THREADS) edges[e] = (uint8_t)p.ising.edges[e];
  for (int e = lane; e < WSTAT; e += 32) wstat[warp * WSTAT + e] = 0.0;
  mbar_wait(bar, 0);
  // ---------------- W -> G = exp(-4W) in place; per-site constants; max|W|
  float wmax = 0.0f;
  for (int i = warp; i < N; i += FAST_WARPS) {
    float *row = Gtab + (size_t)i * MP;
    float rs = 0.0f;
    for (int j = lane; j < MP; j += 32) {
      float gv = 1.0f;  // padding: A + B*1 / A*1 + B leave the products unchanged up to the common factor (A+B)
      if (j < M) {
        const float w = row[j];
        gv = expf(-4.0f * w);
        rs += w;
        wmax = fmaxf(wmax, fabsf(w));
      }
      row[j] = gv;
    }
    rs = warp_sum(rs);
    if (lane == 0) {
      const float x = LOG2E * 2.0f * rs, y = avis != nullptr ? LOG2E * 2.0f * avis[i] : 0.0f;
      if (!(fabsf(y) < 1000.0f)) wmax = __int_as_float(0x7f800000);  // visible bias beyond the fixed-point range: hand over
      rctab[i] = make_float4(x, y, __int_as_float(__float2int_rn(x * FX_SCALE)), __int_as_float(__float2int_rn(y * FX_SCALE)));
    }
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, m));
  if (lane == 0) red[warp] = wmax;
  __syncthreads();
  wmax = 0.0f;
  for (int w = 0; w < FAST_WARPS; ++w) wmax = fmaxf(wmax, red[w]);
  // renormalisation period r: A + B = 1 is restored as soon as r accepted moves have piled up, so a product (2*NP factors)
  // sees at most r - 1 un-normalised accepts: factors within G^(+-r) of 1.  2 NP r 4 wmax log2(e) must stay below the fp32
  // exponent range (and 32 lanes of it, times FX_SCALE, inside int32)
  int renorm = 0;
  {
    const float per = (float)(2 * ((NE + 1) / 2)) * 4.0f * wmax * LOG2E;
    renorm = 32;
    while (renorm >= 1 && (float)renorm * per > EXP_RANGE) renorm >>= 1;
    if (!(wmax < 1.0e30f)) renorm = 0;  // NaN / Inf weights
  }
  if (renorm < 1) {  // weights too large for the product form: hand over to the generic kernel queued behind us
    if (blockIdx.x == 0 && tid == 0) {
      flags[0] = 1;
      if (p.no_handover && p.stats_out != nullptr) p.stats_out[0] = __longlong_as_double(0x7ff8000000000000ll);  // NaN: NK_SWEEP_NO_HANDOVER callers repeat the call
    }
    return;
  }

  const int T_total = (p.n_discard + p.chain_length) * p.sweep_size;
  const float hh = (float)p.ising.h, JJ = (float)p.ising.J;
  // loop-invariant values the hot loop needs, made opaque so that they stay in registers instead of being
  // rematerialised (S2R + LEA + constant-bank loads) on every proposal
  float pw = (float)p.machine_pow;
  const float inv_pw = pw > 0.0f ? 1.0f / pw : 0.0f;
  uint32_t rc_s = smem_u32(rctab);
  const uint32_t g_s = smem_u32(Gtab);
  uint32_t lane16 = 16u * lane;                             // + row address
  uint32_t tailoff = 512u * NFULL + 4u * TAIL * lane;       // + row address
  uint32_t rec_s = smem_u32(rectab) + 512u * warp;
  uint32_t sig_s = smem_u32(sigtab) + (uint32_t)SIG_STRIDE * warp;
  int sweep_size = p.sweep_size;
  int lane_o = lane;
  asm volatile("" : "+r"(rc_s), "+r"(lane16), "+r"(tailoff), "+r"(sweep_size), "+r"(lane_o), "+r"(rec_s), "+r"(sig_s));
  double *ws = wstat + warp * WSTAT;  // [0..7] partial sums of this warp's chains, [8] block sum, [9] / [10] half sums, [11] chain sum
  const bool want_stats = p.stats_out != nullptr && p.eloc_kind == 1;
  const int L = p.chain_length;
  const int l_block = (L / 32) > 1 ? (L / 32) : 1;  // mc_stats_old.py:100
  const int n_b = L / l_block, half = L / 2;

  // chains are dealt out warp-major (round r gives warp w of CTA b the chain (r * FAST_WARPS + w) * gridDim + b): the chains of
  // the last, partial round (2^16 chains = 15.8 rounds of 148 x 28 warps) then thin out every SM equally instead of leaving
  // whole SMs idle while the others run a full round
  for (int chain = warp * gridDim.x + blockIdx.x; chain < (int)p.B; chain += gridDim.x * FAST_WARPS) {
    ChainRegs<NPA> c;
    // ---- sigma as bytes in shared memory (1 = spin down); every lane performs every later store itself
    __syncwarp();
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int idx = 32 * b + lane;
      if (idx < N) sts_u8(sig_s + idx, p.sigma[(size_t)chain * N + idx] < 0 ? 1u : 0u);
    }
    // ---- theta (from the GEMM) -> (A, B) = (e^theta, e^-theta) / (2 cosh theta)
    {
      const float *th = theta_ws + (size_t)chain * M;
      c.At = 0.5f;
      c.Bt = 0.5f;
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        const int j = LM::unit(e, lane);
        float av = 0.5f, bv = 0.5f;  // padding units: theta = 0
        if (j < M) {
          const float x = th[j];
          const float ex = expf(-2.0f * fabsf(x));
          const float big = 1.0f / (1.0f + ex), small = ex * big;
          av = x >= 0.0f ? big : small;
          bv = x >= 0.0f ? small : big;
        }
        if (e < 2 * NP2) {
          if (e & 1) {
            c.A2[e >> 1].y = av;
            c.B2[e >> 1].y = bv;
          } else {
            c.A2[e >> 1].x = av;
            c.B2[e >> 1].x = bv;
          }
        } else {
          c.At = av;
          c.Bt = bv;
        }
      }
    }
    c.nacc = 0;
    c.next_renorm = (uint32_t)renorm;
    int in_sweep = 0, sweep_idx = 0;
    const uint64_t gchain = p.chain_offset + (uint64_t)chain;
    __syncwarp();

    auto renormalise = [&]() {  // A + B = 1 (approximately: R is re-measured, not assumed)
#pragma unroll
      for (int q = 0; q < NP2; ++q) {
        const float2 s2 = fadd2(c.A2[q], c.B2[q]);
        const float2 i2 = make_float2(rcp_fast(s2.x), rcp_fast(s2.y));
        c.A2[q] = fmul2(c.A2[q], i2);
        c.B2[q] = fmul2(c.B2[q], i2);
      }
      if (HAS_T) {
        const float it = rcp_fast(c.At + c.Bt);
        c.At *= it;
        c.Bt *= it;
      }
      c.R = __reduce_add_sync(0xffffffffu, __float2int_rn(lg2_fast(lane_norm<NP2, NPA, HAS_T>(c)) * FX_SCALE));
      c.next_renorm = c.nacc + (uint32_t)renorm;
    };
    c.R = __reduce_add_sync(0xffffffffu, __float2int_rn(lg2_fast(lane_norm<NP2, NPA, HAS_T>(c)) * FX_SCALE));
    // log psi of the current state from (A, B): lncosh(theta_j) = log((A_j + B_j) / (2 sqrt(A_j B_j)))
    auto logpsi_now = [&]() -> float {
      float acc2 = 0.0f;  // log2 units
#pragma unroll
      for (int q = 0; q < NP2; ++q) {
        acc2 += log2f(c.A2[q].x + c.B2[q].x) - 0.5f * (log2f(c.A2[q].x) + log2f(c.B2[q].x)) - 1.0f;
        acc2 += log2f(c.A2[q].y + c.B2[q].y) - 0.5f * (log2f(c.A2[q].y) + log2f(c.B2[q].y)) - 1.0f;
      }
      if (HAS_T) acc2 += log2f(c.At + c.Bt) - 0.5f * (log2f(c.At) + log2f(c.Bt)) - 1.0f;
      float vis2 = 0.0f;  // log2e * sum_i a_i sigma_i
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int idx = 32 * b + lane;
        if (idx < N) vis2 += lds_u8(sig_s + idx) ? -0.5f * rctab[idx].y : 0.5f * rctab[idx].y;
      }
      return LN2 * warp_sum(acc2 + vis2);
    };

    for (int tt = 0; tt < T_total; tt += 32) {
      // ---- 32 proposals' worth of randomness, one Philox call per lane; the lane leaves the proposal's record
      //      {row address, spin address, fix(log2(u) / machine_pow) - fix(x_site), fix(y_site)} in shared memory
      {
        int site_l = 0, thr_l = 0;
        if (tt + lane < T_total) {
          uint32_t w0;
          float u;
          if (p.stream_w0 != nullptr) {
            w0 = p.stream_w0[(size_t)(tt + lane) * p.B + chain];
            u = reinterpret_cast<const float *>(p.stream_u)[(size_t)(tt + lane) * p.B + chain];
          } else {
            const uint4 w = philox_words(p.seed, p.t0 + (uint64_t)(tt + lane), gchain, STREAM_STEP);
            w0 = w.x;
            u = uniform_from_words<float>(w);
          }
          site_l = (int)__umulhi(w0, (uint32_t)N);
          // u == 0 or machine_pow == 0: always accept
          const float t2 = (pw > 0.0f && u > 0.0f) ? log2f(u) * inv_pw * FX_SCALE : (float)THR_MIN;
          thr_l = __float2int_rn(fmaxf(t2, (float)THR_MIN));
        }
        const float4 rc = lds128(rc_s + 16u * site_l);
        uint4 rec;
        rec.x = g_s + (uint32_t)site_l * (uint32_t)(MP * 4);
        rec.y = sig_s + (uint32_t)site_l;
        rec.z = (uint32_t)thr_l - (uint32_t)__float_as_int(rc.z);
        rec.w = (uint32_t)__float_as_int(rc.w);
        __syncwarp();
        sts128u(rec_s + 16u * lane_o, rec);
        __syncwarp();
      }
      const int nb = min(32, T_total - tt);
      int k = 0;
      while (k < nb) {
        // a segment = proposals up to the end of the sweep / of this batch
        const int kend = k + min(nb - k, sweep_size - in_sweep);
        in_sweep += kend - k;
        for (; k < kend; ++k) {
          const uint4 rec = lds128u(rec_s + 16u * k);
          const uint32_t sdown = lds_u8(rec.y);
          float2 g2[NPA];
          float gt = 1.0f;
          LM::load_row(rec.x + lane16, rec.x + tailoff, g2, gt);
          // spin down (nu = +1): prod (B g + A), accept B <- B g;   spin up (nu = -1): prod (A g + B), accept A <- A g
          if (sdown) {
            const float P = lane_product<NP2, NPA, HAS_T>(c.B2, c.Bt, c.A2, c.At, g2, gt);
            const int Rp = __reduce_add_sync(0xffffffffu, __float2int_rn(lg2_fast(P) * FX_SCALE));
            if ((int)rec.z < (int)((uint32_t)Rp - (uint32_t)c.R + rec.w)) {
#pragma unroll
              for (int q = 0; q < NP2; ++q) c.B2[q] = fmul2(c.B2[q], g2[q]);
              if (HAS_T) c.Bt *= gt;
              c.R = Rp;
              sts_u8(rec.y, 0u);
              if (++c.nacc == c.next_renorm) renormalise();
            }
          } else {
            const float P = lane_product<NP2, NPA, HAS_T>(c.A2, c.At, c.B2, c.Bt, g2, gt);
            const int Rp = __reduce_add_sync(0xffffffffu, __float2int_rn(lg2_fast(P) * FX_SCALE));
            if ((int)rec.z < (int)((uint32_t)Rp - (uint32_t)c.R - rec.w)) {
#pragma unroll
              for (int q = 0; q < NP2; ++q) c.A2[q] = fmul2(c.A2[q], g2[q]);
              if (HAS_T) c.At *= gt;
              c.R = Rp;
              sts_u8(rec.y, 1u);
              if (++c.nacc == c.next_renorm) renormalise();
            }
          }
        }
        if (in_sweep == sweep_size) {
          in_sweep = 0;
          const int sw = sweep_idx - p.n_discard;
          ++sweep_idx;
          if (sw >= 0) {
            const size_t o = (size_t)chain * p.chain_length + sw;
            if (p.samples_out != nullptr) {
#pragma unroll
              for (int b = 0; b < 4; ++b) {
                const int idx = 32 * b + lane;
                if (idx < N) p.samples_out[o * N + idx] = lds_u8(sig_s + idx) ? (int8_t)-1 : (int8_t)1;
              }
            }
            if (p.logp_out != nullptr) {
              const float lp = logpsi_now();
              if (lane == 0) reinterpret_cast<float *>(p.logp_out)[o] = (float)p.machine_pow * lp;
            }
            if (p.tanh_out != nullptr) {
              // tanh(theta_j) = (A_j - B_j) / (A_j + B_j): what the forces need, while it is in registers
              float *to = reinterpret_cast<float *>(p.tanh_out) + o * M;
#pragma unroll
              for (int e = 0; e < NE; ++e) {
                const int j = LM::unit(e, lane);
                float av, bv;
                if (e < 2 * NP2) {
                  av = (e & 1) ? c.A2[e >> 1].y : c.A2[e >> 1].x;
                  bv = (e & 1) ? c.B2[e >> 1].y : c.B2[e >> 1].x;
                } else {
                  av = c.At;
                  bv = c.Bt;
                }
                if (j < M) to[j] = __fdividef(av - bv, av + bv);
              }
            }
            if (p.eloc_kind == 1) {
              // E_loc = J sum_<ij> s_i s_j - h sum_i exp(delta_i)
              uint32_t sw4[4];
#pragma unroll
              for (int b = 0; b < 4; ++b) {
                const int idx = 32 * b + lane;
                sw4[b] = __ballot_sync(0xffffffffu, idx < N && lds_u8(sig_s + idx) != 0u);
              }
              int zz = 0;
              for (int e = lane; e < E; e += 32) {
                const int a = edges[2 * e], bq = edges[2 * e + 1];
                const uint32_t wa = (a < 64) ? ((a < 32) ? sw4[0] : sw4[1]) : ((a < 96) ? sw4[2] : sw4[3]);
                const uint32_t wb = (bq < 64) ? ((bq < 32) ? sw4[0] : sw4[1]) : ((bq < 96) ? sw4[2] : sw4[3]);
                zz += 1 - 2 * (int)(((wa >> (a & 31)) ^ (wb >> (bq & 31))) & 1u);
              }
              zz = __reduce_add_sync(0xffffffffu, zz);
              float off = 0.0f;
              if (hh != 0.0f) {
                // accurate (float) normalisation log2 prod_j (A_j + B_j) for this sample
                const float Rf = warp_sum(lg2_fast(lane_norm<NP2, NPA, HAS_T>(c)));
                // sites in groups of 16: 16 independent lane products (ILP), then a transposed butterfly that needs
                // 16 shuffles for 16 sites instead of 80; lane l ends up with the total of site base + ((l >> 1) & 15)
                float off_l = 0.0f;
                const int myidx = (lane_o >> 1) & 15;
                auto group = [&](const int base, const bool full) {
                  const uint32_t word = sw4sel(sw4, base >> 5) >> (base & 16);
                  float v[16];
#pragma unroll
                  for (int jj = 0; jj < 16; ++jj) {
                    const int sidx = base + jj;
                    v[jj] = 0.0f;
                    if (full || sidx < N) {
                      const uint32_t so = g_s + (uint32_t)sidx * (uint32_t)(MP * 4);
                      float2 r2[NPA];
                      float rt = 1.0f;
                      LM::load_row(so + lane16, so + tailoff, r2, rt);
                      const float Ps = ((word >> jj) & 1u) ? lane_product<NP2, NPA, HAS_T>(c.B2, c.Bt, c.A2, c.At, r2, rt)
                                                            : lane_product<NP2, NPA, HAS_T>(c.A2, c.At, c.B2, c.Bt, r2, rt);
                      v[jj] = lg2_fast(Ps);
                    }
                  }
#pragma unroll
                  for (int h = 8; h >= 1; h >>= 1) {
                    const bool up = (lane_o & (2 * h)) != 0;
#pragma unroll
                    for (int jj = 0; jj < h; ++jj) {
                      const float send = up ? v[jj] : v[jj + h];
                      const float keep = up ? v[jj + h] : v[jj];
                      v[jj] = keep + __shfl_xor_sync(0xffffffffu, send, 2 * h);
                    }
                  }
                  const float tot = v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
                  const int mys = base + myidx;
                  if ((full || mys < N) && (lane_o & 1) == 0) {
                    const float2 rcs = lds64(rc_s + 16u * mys);
                    const float cst = ((word >> myidx) & 1u) ? rcs.x + rcs.y : rcs.x - rcs.y;
                    off_l += ex2_fast(tot - Rf + cst);
                  }
                };
                int base = 0;
#pragma unroll 1
                for (; base + 16 <= N; base += 16) group(base, true);
                if (base < N) group(base, false);
                off = warp_sum(off_l);
              }
              const float e_loc = JJ * (float)zz - hh * off;
              if (lane == 0) {
                store_as<float>(p.eloc_out, o, e_loc, p.eloc_dtype);
                if (want_stats) {
                  // shifted sums of mc_stats_old.py:87-196 (the layout of nk_stats_partial's phase 1)
                  const double d = (double)e_loc - p.stats_shift;
                  ws[0] += d * d;
                  ws[11] += d;
                  if (sw < n_b * l_block) {
                    ws[8] += d;
                    if ((sw + 1) % l_block == 0) {
                      const double m = ws[8] / (double)l_block;
                      ws[3] += m;
                      ws[4] += m * m;
                      ws[8] = 0.0;
                    }
                  }
                  if (sw < half)
                    ws[9] += d;
                  else if (sw < 2 * half)
                    ws[10] += d;
                }
              }
            }
          }
        }
      }
    }
    // ---- write the chain state back
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int idx = 32 * b + lane;
      if (idx < N) p.sigma[(size_t)chain * N + idx] = lds_u8(sig_s + idx) ? (int8_t)-1 : (int8_t)1;
    }
    const float lp = logpsi_now();
    if (lane == 0) {
      reinterpret_cast<float *>(p.log_prob)[chain] = (float)p.machine_pow * lp;
      p.n_accepted[chain] += (int64_t)c.nacc;
      if (want_stats && L > 0) {
        const double m = ws[11] / (double)L;
        ws[1] += m;
        ws[2] += m * m;
        ws[7] += ws[11];
        if (half > 0) {
          const double ha = ws[9] / (double)half, hb = ws[10] / (double)half;
          ws[5] += ha + hb;
          ws[6] += ha * ha + hb * hb;
        }
        ws[9] = ws[10] = ws[11] = 0.0;
      }
    }
  }
  if (want_stats) {  // one atomic per CTA and partial sum
    __syncthreads();
    if (tid < NK_STATS_NPARTIAL) {
      double s = 0.0;
      for (int w = 0; w < FAST_WARPS; ++w) s += wstat[w * WSTAT + tid];
      atomicAdd(p.stats_out + tid, s);
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
  const int E = a.eloc_kind == 1 ? a.ising.n_edges : 0;
  if (fast_smem_bytes(a.rbm.N, fs.mp, E) > 227 * 1024) return false;
  if ((size_t)a.rbm.N * a.rbm.M * 4 >= (1u << 20)) return false;  // mbarrier tx-count range
  if (a.B >= (1ll << 31) || (int64_t)(a.n_discard + a.chain_length) * a.sweep_size >= (1ll << 31)) return false;  // 32-bit counters
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

}  // namespace nk
