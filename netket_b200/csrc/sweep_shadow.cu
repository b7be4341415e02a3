// fp64 Metropolis sweep with fp32 "shadow" decisions (LocalRule, optional fused transverse-field-Ising local energy and MC
// statistics).  cfg-3 in fp64: the weight table in double (325 KB) does not fit shared memory, and at the benchmark's ~96 %
// acceptance every proposal of the one-table kernel (sweep_prod.cuh) touches a double row: it is bound by the latency of
// those rows (a third of them through L2) at 3 warps per scheduler.  Here the two jobs of the table are separated.
//
//   S  sweep.  The accept / reject decisions are taken on an fp32 shadow of the state, (A32, B32) ~ (A, B), against an
//      fp32 copy of the table G32 = float(exp(-4 W)) that IS resident (166 KB): exactly the proposal loop of sweep_fast.cu.
//      A decision whose fixed-point margin |fix(log2 ratio) - fix(log2(u) / machine_pow)| lies within BAND of zero - the
//      shadow's error is ~2 units typical, < 40 worst case - is re-decided in double precision, out of line (exact_accept:
//      the parked double state brought up to date with the rows of the sites flipped since, the proposal's product formed in
//      double, u < exp(machine_pow (x_i +- y_i + sum log(P / norm))): metropolis.py:441-450; ~1.6e-4 of the proposals at the
//      benchmark's weights).  So the chain is the fp64 chain, bit for bit, and no double is touched while sweeping.
//   U  update, once per sweep.  The double state only has to follow the NET change of the sweep: a site flipped an even number
//      of times multiplies A_j and B_j by the same G_ij, which cancels in every ratio.  The double table is streamed once
//      through a shared-memory ring by TMA bulk copies of 4 rows (warp 0 is a dedicated producer; full / empty mbarriers) and
//      every consumer warp applies the rows of its chain's net-flipped sites (~43 of 100) to its (A, B) registers; the state is
//      rescaled by exact powers of two and the shadow is refreshed from it (its rounding drift never outlives a sweep).
//   E  local energy of a recorded sample: the table is streamed a second time, every warp forms prod_j (X_j G_ij + Y_j) for
//      every site i in double (13 DFMA + 12 DMUL per lane), 8 sites per transposed multiplicative butterfly.
// L2 -> shared traffic is one table per CTA and pass instead of one row per proposal and chain.  What binds the kernel is neither
// a pipe nor a bandwidth but dependent-issue latency at 3 warps per scheduler (168 registers: 26 doubles of state per lane, two
// rows in flight); measurements, per-phase cycle counts and the rejected variants: profiles/r02_shadow_f64_experiments.md.
//
// Replaces netket/sampler/metropolis.py:427-462 + rules/local.py:40-49, vqs/mc/kernels.py:62-71 with
// operator/_ising/jax.py:125-165 and the sums of stats/mc_stats_old.py:87-196, like sweep_fast.cu / sweep_prod.cuh.
// Shapes: N <= 128, hidden units in whole 128-unit chunks plus at most 32 (M = 400: 3 + 16), so that a lane owns the SAME
// hidden units in the fp32 layout (4 per 128-bit load) and in the fp64 layout (2 per load): the fp32 table is stored permuted.
#include <stdlib.h>

#include "fast_common.cuh"
#include "sweep_prod.cuh"

namespace nk {
namespace shadow {

using namespace fast;

constexpr int WARPS = 12;         // warp 0 only feeds the ring of double rows (TMA); warps 1 .. 11 own one chain each
constexpr int CWARPS = WARPS - 1;  // consumer warps = chains in flight per CTA
constexpr int THREADS = WARPS * 32;
#ifndef NK_SH_GROUP
#define NK_SH_GROUP 4
#endif
#ifndef NK_SH_STAGES
#define NK_SH_STAGES 3
#endif
constexpr int GROUP = NK_SH_GROUP;  // rows per TMA bulk copy (one 3.3 KB row per copy ran at ~1 row / 1000 cycles: the copy engine
                                 // wants fewer, larger requests)
constexpr int STAGES = NK_SH_STAGES;  // groups of rows in the ring: one being read, the others in flight
constexpr int BAND = 64;         // fixed-point units (2^-19 in log2): decisions closer than this to the threshold are re-decided in fp64
                                 // (the shadow's error: ~2 units typical; PROD_FX_BAND of the one-table kernel is the same 64)
constexpr float SH_EXP_RANGE = 100.0f;
constexpr int SIG_STRIDE = 128;  // spin bytes per warp
constexpr int WSTAT = 12;
constexpr int REC_SLOT = 33 * 16;  // 32 proposal records per warp + one never-written entry: the record prefetch runs one past the last

struct Layout {
  int g32, rcx, rcd, edges, rec, sig, sigp, wstat, ring, bars, red, total;
};
__host__ __device__ inline Layout make_layout(int N, int MP32, int row64, int E) {
  Layout L;
  int o = 0;
  L.g32 = o;
  o += N * MP32 * 4;
  L.rcx = o;
  o += (N * 8 + 15) & ~15;  // fix(x), fix(y)
  L.rcd = o;
  o += N * 16;  // exp(x + y), exp(x - y) (double)
  L.edges = o;
  o += (2 * E + 15) & ~15;
  L.rec = o;
  o += WARPS * REC_SLOT;
  L.sig = o;
  o += WARPS * SIG_STRIDE;
  L.sigp = o;
  o += WARPS * SIG_STRIDE;
  L.wstat = o;
  o += WARPS * WSTAT * 8;
  L.ring = o;
  o += STAGES * GROUP * row64;
  L.bars = o;
  o += (2 * STAGES + 2) * 8;
  L.red = o;
  o += 32 * 4;
  L.total = o;
  return L;
}

__device__ __forceinline__ uint32_t sw4sel4(const uint32_t (&w)[4], int i) {
  return (i < 2) ? ((i == 0) ? w[0] : w[1]) : ((i == 2) ? w[2] : w[3]);
}
// all lanes poll, the exit is a warp vote: the loop is uniform for the compiler (no divergence to repair around the warp-wide
// reductions of the hot loops).  A poll is ~8 issued instructions; un-throttled, the polls of the producer (idle while the
// consumers sweep) and of consumers waiting for the stream were 29 % of all instructions the kernel issued (ncu source page,
// profiles/r02_shadow_f64_experiments.md), taken from the issue slots of the warps doing arithmetic: after FAST failed probes
// a consumer warp sleeps NS nanoseconds between probes (measured: no change in time either way).
#ifndef NK_SH_PSLEEP
#define NK_SH_PSLEEP 0  // the producer polls un-throttled: a sleeping producer issues the next copy late (U ring wait 6.5k -> 5.9k cycles)
#endif
#ifndef NK_SH_PFAST
#define NK_SH_PFAST 4
#endif
#ifndef NK_SH_CSLEEP
#define NK_SH_CSLEEP 40
#endif
template <int FAST, int NS>
__device__ __forceinline__ void mbar_wait_uniform(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  int tries = 0;
  for (;;) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
#ifdef NK_SH_TRYWAIT  // developer switch (compute-sanitizer racecheck experiment, profiles/README.md)
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
#else
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"  // non-blocking probe: try_wait suspends the thread for a
#endif
        "selp.u32 %0, 1, 0, p;\n"                                     // system-chosen time (measured: ~0.5 us per ring hand-over)
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (__all_sync(0xffffffffu, done != 0u)) break;
    if (NS > 0 && ++tries > FAST) __nanosleep(NS);
  }
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

#ifdef NK_SH_PROFILE
// build-time instrumentation (tools/build_variant.py prof -DNK_SH_PROFILE): cycles of warp 1 of CTA 0 per phase
__device__ unsigned long long nk_sh_prof[16];
#define NK_PROF_NOW() clock64()
#define NK_PROF_ADD(slot, dt)                                              \
  do {                                                                     \
    if (blockIdx.x == 0 && warp == 1 && lane == 0) nk_sh_prof[slot] += (unsigned long long)(dt); \
  } while (0)
#else
#define NK_PROF_NOW() 0ll
#define NK_PROF_ADD(slot, dt) \
  do {                        \
    (void)(dt);               \
  } while (0)
#endif

// prod_e (X_e g_e + Y_e) with four interleaved accumulators per row (dependency depth ~ NV / 4 + 2 instead of NV / 2 + 1),
// two rows at once: the multiply tree of one row fills the dependency stalls of the other (the kernel runs at < 3 warps per
// scheduler: the stalls are not hidden by other warps).  D0 / D1: the row's site is spin down (X = B, Y = A)
template <int NV, bool D0, bool D1>
__device__ __forceinline__ void lane_product64_x2(const double (&A)[NV], const double (&Bv)[NV], const double (&g0)[NV], const double (&g1)[NV],
                                                  double &p0, double &p1) {
  double a0[4] = {1.0, 1.0, 1.0, 1.0}, a1[4] = {1.0, 1.0, 1.0, 1.0};
#pragma unroll
  for (int e = 0; e < NV; ++e) {
    const double c0 = D0 ? fma(Bv[e], g0[e], A[e]) : fma(A[e], g0[e], Bv[e]);
    const double c1 = D1 ? fma(Bv[e], g1[e], A[e]) : fma(A[e], g1[e], Bv[e]);
    a0[e & 3] = e < 4 ? c0 : a0[e & 3] * c0;
    a1[e & 3] = e < 4 ? c1 : a1[e & 3] * c1;
  }
  if (NV == 1) {
    p0 = a0[0];
    p1 = a1[0];
  } else if (NV == 2) {
    p0 = a0[0] * a0[1];
    p1 = a1[0] * a1[1];
  } else if (NV == 3) {
    p0 = (a0[0] * a0[1]) * a0[2];
    p1 = (a1[0] * a1[1]) * a1[2];
  } else {
    p0 = (a0[0] * a0[1]) * (a0[2] * a0[3]);
    p1 = (a1[0] * a1[1]) * (a1[2] * a1[3]);
  }
}

// fp64 re-decision of one proposal (rare, ~3e-4 of the proposals).  The double state as of the last update is parked in the
// workspace; the sites flipped since then are the bytes where sigma differs from its copy of that moment.  The state is
// brought up to date in registers (one double row per such site, straight from the L2-resident table), the proposal's product
// is formed in double and  accept <=> u < exp(machine_pow * (x_i +- y_i + sum_lanes log(P / norm)))   (metropolis.py:441-450).
template <int NF64, int TL>
static __device__ __noinline__ bool exact_accept(const ProdArgs &p, const double *park, uint32_t sig_s, uint32_t sigp_s, int site, bool sdown,
                                                 double u, double pw, int lane) {
  using LD = prod::LaneMap<double, NF64, TL>;
  constexpr int NV = LD::NV;
  constexpr uint32_t FULL = 0xffffffffu;
  const int N = p.s.rbm.N;
  double A[NV], Bv[NV];
#pragma unroll
  for (int e = 0; e < NV; ++e) {
    A[e] = park[(size_t)(2 * e) * 32];
    Bv[e] = park[(size_t)(2 * e + 1) * 32];
  }
  auto load_row = [&](int i, double(&g)[NV]) {
    const unsigned char *row = p.gtab + (size_t)i * p.L.row_bytes;
    prod::load_row_p<NF64, TL>(row + 16 * lane, row + LD::TAIL_OFF + LD::TAIL_LANE * lane, g);
  };
  for (int b = 0; b < 4; ++b) {
    const int idx = 32 * b + lane;
    const uint32_t now = idx < N ? lds_u8(sig_s + idx) : 0u, was = idx < N ? lds_u8(sigp_s + idx) : 0u;
    uint32_t fl = __ballot_sync(FULL, now != was);
    const uint32_t dn = __ballot_sync(FULL, now != 0u);
    while (fl != 0u) {
      const int bit = __ffs(fl) - 1;
      fl &= fl - 1u;
      double g[NV];
      load_row(32 * b + bit, g);
      if ((dn >> bit) & 1u) {  // +1 -> -1: A <- A G
#pragma unroll
        for (int e = 0; e < NV; ++e) A[e] *= g[e];
      } else {
#pragma unroll
        for (int e = 0; e < NV; ++e) Bv[e] *= g[e];
      }
    }
  }
  double g[NV];
  load_row(site, g);
  double P = 1.0, nrm = 1.0;
#pragma unroll
  for (int e = 0; e < NV; ++e) {
    P *= sdown ? fma(Bv[e], g[e], A[e]) : fma(A[e], g[e], Bv[e]);
    nrm *= A[e] + Bv[e];
  }
  const double d = warp_sum(log(P / nrm));
  const RcD &rc = reinterpret_cast<const RcD *>(p.aux + p.L.rc_off)[site];
  return u < exp(pw * ((sdown ? rc.xn + rc.yn : rc.xn - rc.yn) + d));
}

template <int NF32, int TL>
__global__ void __launch_bounds__(THREADS, 1) sweep_shadow_kernel(const __grid_constant__ ProdArgs p, int give, double *__restrict__ park_all) {
  constexpr int NF64 = 2 * NF32;
  using LM = Lanes<NF32, TL>;                 // fp32 layout (permuted table)
  using LD = prod::LaneMap<double, NF64, TL>; // fp64 layout
  constexpr int NP2 = LM::NP2, NPA = LM::NPA, NE = LM::NE, MP = LM::MP;
  constexpr bool HAS_T = LM::HAS_T;
  constexpr int NV = LD::NV;  // = NE
  static_assert(NV == NE, "a lane owns the same hidden units in both layouts");
  constexpr uint32_t FULL = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char smem[];
  const SweepKernelArgs &s = p.s;
  const int N = s.rbm.N, M = s.rbm.M, E = s.eloc_kind == 1 ? s.ising.n_edges : 0;
  const int row64 = p.L.row_bytes;
  const Layout L = make_layout(N, MP, row64, E);
  float *G32 = reinterpret_cast<float *>(smem + L.g32);
  int2 *rcx = reinterpret_cast<int2 *>(smem + L.rcx);
  double2 *rcd = reinterpret_cast<double2 *>(smem + L.rcd);
  uint8_t *edges = smem + L.edges;
  double *wstat = reinterpret_cast<double *>(smem + L.wstat);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + L.bars);
  uint64_t *empty = full + STAGES;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(FULL, tid >> 5, 0);
  const float LOG2E = 1.4426950408889634f;

  if (p.run_if != nullptr && *p.run_if == 0) return;
  if (p.flags[p.giveup] != 0) return;  // the prep kernels found the weights outside the product form's range altogether
  // ---------------- range of the fp32 shadow (as sweep_fast.cu) and of the plain double local energy
  int renorm = 0;
  {
    const float wmax = __int_as_float(p.flags[2]);
    const float per = (float)(2 * ((NE + 1) / 2)) * 4.0f * wmax * LOG2E;
    renorm = 32;
    while (renorm >= 1 && (float)renorm * per > SH_EXP_RANGE) renorm >>= 1;
    if (!(wmax < 1.0e30f)) renorm = 0;
    if (p.flags[7] != 0) renorm = 0;  // row sums so large that the double products need the (mantissa, exponent) form
  }
  if (renorm < 1) {
    if (blockIdx.x == 0 && tid == 0) p.flags[give] = 1;  // the one-table kernel queued behind takes over
    return;
  }

  // ---------------- tables: fp32 copy of G (permuted), per-site constants, edges; barriers
  const RcD *rcg = reinterpret_cast<const RcD *>(p.aux + p.L.rc_off);
  for (int i = warp; i < N; i += WARPS) {
    const double *row = reinterpret_cast<const double *>(p.gtab + (size_t)i * row64);
    float *dst = G32 + (size_t)i * MP;
    for (int pos = lane; pos < MP; pos += 32) {
      int unit = pos;
      if (pos < 128 * NF32) {
        const int c = pos >> 7, l = (pos & 127) >> 2, sub = pos & 3;
        unit = 128 * c + 64 * (sub >> 1) + 2 * l + (sub & 1);
      }
      dst[pos] = (float)row[unit];
    }
    if (lane == 0) {
      rcx[i] = make_int2(rcg[i].fx, rcg[i].fy);
      rcd[i] = make_double2(rcg[i].ep, rcg[i].em);
    }
  }
  {
    const uint16_t *eg = reinterpret_cast<const uint16_t *>(p.aux + p.L.edges_off);
    for (int e = tid; e < 2 * E; e += THREADS) edges[e] = (uint8_t)eg[e];
  }
  for (int e = lane; e < WSTAT; e += 32) wstat[warp * WSTAT + e] = 0.0;
  if (tid == 0) {
    for (int st = 0; st < STAGES; ++st) {
      mbar_init(full + st, 1);        // the producer's expect_tx; completed by the bytes of the bulk copy
      mbar_init(empty + st, CWARPS);  // one arrival per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int n_sweeps = s.n_discard + s.chain_length;
  const int T_total = n_sweeps * s.sweep_size;
  const double pw = s.machine_pow;
  const double inv_pw = pw > 0.0 ? 1.0 / pw : 0.0;
  const double hh = s.ising.h, JJ = s.ising.J;
  const bool want_eloc = s.eloc_kind == 1;
  const bool want_stats = s.stats_out != nullptr && want_eloc;
  const int CL = s.chain_length;
  const int l_block = (CL / 32) > 1 ? (CL / 32) : 1;
  const int n_b = CL / l_block, half = CL / 2;
  uint32_t g_s = smem_u32(G32);
  uint32_t lane16 = 16u * lane, tailoff = 512u * NF32 + 4u * TL * lane;
  uint32_t rec_s = smem_u32(smem + L.rec) + (uint32_t)REC_SLOT * warp;
  uint32_t sig_s = smem_u32(smem + L.sig) + (uint32_t)SIG_STRIDE * warp;
  uint32_t sigp_s = smem_u32(smem + L.sigp) + (uint32_t)SIG_STRIDE * warp;
  const uint32_t ring_s = smem_u32(smem + L.ring);
  int sweep_size = s.sweep_size;
  int lane_o = lane;
  asm volatile("" : "+r"(lane16), "+r"(tailoff), "+r"(sweep_size), "+r"(lane_o), "+r"(rec_s), "+r"(sig_s), "+r"(sigp_s), "+r"(g_s));
  double *ws = wstat + warp * WSTAT;

  // ---------------- the stream of double rows: every pass (U of every sweep, E of every recorded sweep) is rows 0 .. N-1
  const int per_round = gridDim.x * CWARPS;
  const int n_rounds = (int)((s.B + per_round - 1) / per_round);
  const uint32_t passes_per_round = (uint32_t)n_sweeps + (want_eloc ? (uint32_t)CL : 0u);
  const int NG = (N + GROUP - 1) / GROUP;  // groups of rows per pass
  const uint32_t total_q = (uint32_t)n_rounds * passes_per_round * (uint32_t)NG;
  // every CTA walks the table from its own starting group: 148 CTAs asking L2 for the same rows at the same moment
  // serialise on its slices
  const int rotg = (int)((((unsigned long long)blockIdx.x * (unsigned)NG) / gridDim.x) % (unsigned)NG);
  const uint32_t stage_bytes = (uint32_t)(GROUP * row64);
  if (warp == 0) {
    // ---------------- producer: one TMA bulk copy per group, as soon as every consumer warp has released the stage.
    // (When the consumers also issued the copies - by TMA from warp 0 or by cp.async from all warps, both were built - every
    // group hand-over cost ~0.5 us of bookkeeping and exposed L2 latency in all warps at once.)
    int gi = rotg;
    uint32_t st = 0, par = 0;  // parity of the `empty` phase to wait for (the stage's previous use); first waited at t = STAGES
    for (uint32_t t = 0; t < total_q; ++t) {
      if (t >= (uint32_t)STAGES) mbar_wait_uniform<NK_SH_PFAST, NK_SH_PSLEEP>(empty + st, par);
      if (lane == 0) {
        const uint32_t bytes = (uint32_t)(min(GROUP, N - GROUP * gi) * row64);
        mbar_expect_tx(full + st, bytes);
        tma_bulk_g2s(smem + L.ring + (size_t)st * stage_bytes, p.gtab + (size_t)(GROUP * gi) * row64, bytes, full + st);
      }
      __syncwarp();
      if (++gi == NG) gi = 0;
      if (++st == (uint32_t)STAGES) {
        st = 0;
        if (t + 1 > (uint32_t)STAGES) par ^= 1u;  // uses 1, 2, 3, ... of a stage wait for phases 0, 1, 0, ...
      }
    }
  }
  // consumer side of the ring: stage, parity and table group of the next group in the stream, kept incrementally
  uint32_t cst = 0, cpar = 0;
  int cgi = rotg;
  long long prof_wait = 0;
  auto grp_wait = [&]() -> uint32_t {  // wait for the next group of the stream; returns the shared address of its first row
    const long long pw0 = NK_PROF_NOW();
    mbar_wait_uniform<2, NK_SH_CSLEEP>(full + cst, cpar);
    prof_wait += NK_PROF_NOW() - pw0;
    return ring_s + cst * stage_bytes;
  };
  auto grp_done = [&]() {  // this warp has read the group's rows into registers (or skipped them)
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + cst);
    if (++cgi == NG) cgi = 0;
    if (++cst == (uint32_t)STAGES) {
      cst = 0;
      cpar ^= 1u;
    }
  };

  for (int round = 0; round < n_rounds && warp != 0; ++round) {
    const long long chain_ll = ((long long)round * CWARPS + (warp - 1)) * gridDim.x + blockIdx.x;
    const bool active = chain_ll < s.B;
    const int chain = (int)chain_ll;
    if (!active) {  // keep the ring moving: consume every row of this round's passes
      for (uint32_t r = 0; r < passes_per_round * (uint32_t)NG; ++r) {
        (void)grp_wait();
        grp_done();
      }
      continue;
    }
    ChainRegs<NPA> c;
    // the double state lives in registers only while it is worked on (U, E); during the sweep it is parked in this warp's
    // slot of the workspace (L2-resident: 6.6 KB per warp) so that the proposal loop keeps the register budget of sweep_fast.cu
    double *park = park_all + ((size_t)(blockIdx.x * WARPS + warp) * (2 * NV)) * 32 + lane;
    auto park_store = [&](const double (&A)[NV], const double (&Bv)[NV]) {
#pragma unroll
      for (int e = 0; e < NV; ++e) {
        park[(size_t)(2 * e) * 32] = A[e];
        park[(size_t)(2 * e + 1) * 32] = Bv[e];
      }
    };
    auto park_load = [&](double (&A)[NV], double (&Bv)[NV]) {
#pragma unroll
      for (int e = 0; e < NV; ++e) {
        A[e] = park[(size_t)(2 * e) * 32];
        Bv[e] = park[(size_t)(2 * e + 1) * 32];
      }
    };
    // ---- sigma (bytes, 1 = spin down): current and as of the last update of the double state
    __syncwarp();
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int idx = 32 * b + lane;
      if (idx < N) {
        const uint32_t dn = s.sigma[(size_t)chain * N + idx] < 0 ? 1u : 0u;
        sts_u8(sig_s + idx, dn);
        sts_u8(sigp_s + idx, dn);
      }
    }
    c.nacc = 0;
    const uint64_t gchain = s.chain_offset + (uint64_t)chain;

    // shadow <- double state (element e of both layouts is the same hidden unit); R re-measured on the shadow
    auto refresh_shadow = [&](const double (&A)[NV], const double (&Bv)[NV]) {
#pragma unroll
      for (int e = 0; e < NV; ++e) {
        const float av = (float)A[e], bv = (float)Bv[e];
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
      if (!HAS_T) c.At = c.Bt = 0.5f;
      c.R = __reduce_add_sync(FULL, __float2int_rn(lg2_fast(lane_norm<NP2, NPA, HAS_T>(c)) * FX_SCALE));
      c.next_renorm = c.nacc + (uint32_t)renorm;
    };
    auto renormalise32 = [&]() {
#pragma unroll
      for (int qq = 0; qq < NP2; ++qq) {
        const float2 s2 = fadd2(c.A2[qq], c.B2[qq]);
        const float2 i2 = make_float2(rcp_fast(s2.x), rcp_fast(s2.y));
        c.A2[qq] = fmul2(c.A2[qq], i2);
        c.B2[qq] = fmul2(c.B2[qq], i2);
      }
      if (HAS_T) {
        const float it = rcp_fast(c.At + c.Bt);
        c.At *= it;
        c.Bt *= it;
      }
      c.R = __reduce_add_sync(FULL, __float2int_rn(lg2_fast(lane_norm<NP2, NPA, HAS_T>(c)) * FX_SCALE));
      c.next_renorm = c.nacc + (uint32_t)renorm;
    };
    // A + B into [1/2, 1) by an exact power of two (no rounding): keeps every product of M factors in range
    auto rescale64 = [&](double (&A)[NV], double (&Bv)[NV]) {
#pragma unroll
      for (int e = 0; e < NV; ++e) {
        const double sm = A[e] + Bv[e];
        const int ex = ((__double2hiint(sm) >> 20) & 0x7ff) - 1022;  // sm = m 2^ex, m in [1/2, 1)
        const double sc = __hiloint2double((1023 - ex) << 20, 0);
        A[e] *= sc;
        Bv[e] *= sc;
      }
    };
    auto logpsi64 = [&](const double (&A)[NV], const double (&Bv)[NV]) -> double {
      double acc = 0.0;
#pragma unroll
      for (int e = 0; e < NV; ++e) acc += prod::lncosh_from_pair(A[e], Bv[e]);
      // padded units carry theta = 0: lncosh = 0
      double vis = 0.0;
      if (s.rbm.a != nullptr) {
        const double *av = reinterpret_cast<const double *>(s.rbm.a);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int idx = 32 * b + lane;
          if (idx < N) vis += lds_u8(sig_s + idx) ? -av[idx] : av[idx];
        }
      }
      return warp_sum(acc + vis);
    };
    {  // ---- theta -> (A, B) = (e^theta, e^-theta) / (2 cosh theta), double
      double A[NV], Bv[NV];
      const double *th = reinterpret_cast<const double *>(p.theta) + (size_t)chain * M;
#pragma unroll
      for (int e = 0; e < NV; ++e) {
        const int j = LD::unit(e, lane);
        const double2 ab = prod::pair_from_theta(j < M ? th[j] : 0.0);
        A[e] = ab.x;
        Bv[e] = ab.y;
      }
      rescale64(A, Bv);
      park_store(A, Bv);
      refresh_shadow(A, Bv);
    }
    __syncwarp();

    // ---- end of a sweep: U (always), outputs and E (recorded sweeps)
    int sweep_idx = 0;
    long long prof_s0 = NK_PROF_NOW();
    auto end_of_sweep = [&]() {
      __syncwarp();
      const long long pt0 = NK_PROF_NOW();
      NK_PROF_ADD(0, pt0 - prof_s0);
      const long long pwt0 = prof_wait;
      double A[NV], Bv[NV];
      park_load(A, Bv);
      uint32_t fl[4], dn[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int idx = 32 * b + lane;
        const uint32_t now = idx < N ? lds_u8(sig_s + idx) : 0u, was = idx < N ? lds_u8(sigp_s + idx) : 0u;
        fl[b] = __ballot_sync(FULL, now != was);
        dn[b] = __ballot_sync(FULL, now != 0u);
        if (idx < N) sts_u8(sigp_s + idx, now);
      }
      // U: the double state follows the net change of the sweep
      for (int gp = 0; gp < NG; ++gp) {
        const int gi = cgi;
        const uint32_t gs = grp_wait();
        // the rows of a group sit in one 32-bit word of the masks (GROUP divides 32; bits past N are zero)
        const int i0 = GROUP * gi;
        const uint32_t flb = (sw4sel4(fl, i0 >> 5) >> (i0 & 31)) & ((1u << GROUP) - 1u);
        const uint32_t dnb = sw4sel4(dn, i0 >> 5) >> (i0 & 31);
        if (flb != 0u) {
#pragma unroll
          for (int sl = 0; sl < GROUP; ++sl) {
            if ((flb >> sl) & 1u) {
              const uint32_t rs = gs + (uint32_t)(sl * row64);
              double g[NV];
              prod::load_row_s<NF64, TL>(rs + 16u * lane_o, rs + (uint32_t)LD::TAIL_OFF + (uint32_t)LD::TAIL_LANE * lane_o, g);
              if ((dnb >> sl) & 1u) {  // +1 -> -1: A <- A G
#pragma unroll
                for (int e = 0; e < NV; ++e) A[e] *= g[e];
              } else {  // -1 -> +1: B <- B G
#pragma unroll
                for (int e = 0; e < NV; ++e) Bv[e] *= g[e];
              }
            }
          }
        }
        grp_done();
      }
      rescale64(A, Bv);
      park_store(A, Bv);
      const long long pt1 = NK_PROF_NOW();
      NK_PROF_ADD(2, pt1 - pt0);
      NK_PROF_ADD(8, prof_wait - pwt0);
      const long long pwt1 = prof_wait;
      const int sw = sweep_idx - s.n_discard;
      ++sweep_idx;
      if (sw >= 0) {
        const size_t o = (size_t)chain * CL + sw;
        if (s.samples_out != nullptr) {
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int idx = 32 * b + lane;
            if (idx < N) s.samples_out[o * N + idx] = ((dn[b] >> lane_o) & 1u) ? (int8_t)-1 : (int8_t)1;
          }
        }
        if (s.logp_out != nullptr) {
          const double lp = logpsi64(A, Bv);
          if (lane == 0) reinterpret_cast<double *>(s.logp_out)[o] = pw * lp;
        }
        if (s.tanh_out != nullptr) {
          double *to = reinterpret_cast<double *>(s.tanh_out) + o * M;
#pragma unroll
          for (int e = 0; e < NV; ++e) {
            const int j = LD::unit(e, lane);
            if (j < M) to[j] = (A[e] - Bv[e]) / (A[e] + Bv[e]);
          }
        }
        if (want_eloc) {
          // E_loc = J sum_<ij> s_i s_j - h sum_i psi(sigma^(i)) / psi(sigma)        (_ising/jax.py:125-165)
          int zz = 0;
          for (int e = lane; e < E; e += 32) {
            const int ea = edges[2 * e], eb = edges[2 * e + 1];
            zz += 1 - 2 * (int)(((sw4sel4(dn, ea >> 5) >> (ea & 31)) ^ (sw4sel4(dn, eb >> 5) >> (eb & 31))) & 1u);
          }
          zz = __reduce_add_sync(FULL, zz);
          double nl = 1.0;
#pragma unroll
          for (int e = 0; e < NV; ++e) nl *= A[e] + Bv[e];
          const double nrm = prod::warp_prod(nl);
          double off_l = 0.0;
          const int myidx = (lane_o >> 1) & 7;
          constexpr int GPB = 8 / GROUP;  // groups per butterfly of 8 sites
          static_assert(GPB * GROUP == 8, "GROUP must divide 8");
          for (int gp = 0; gp < NG; gp += GPB) {
            double v[8];
            int gis[GPB];
#pragma unroll
            for (int h = 0; h < GPB; ++h) {
              gis[h] = -1;
#pragma unroll
              for (int sl = 0; sl < GROUP; ++sl) v[GROUP * h + sl] = 1.0;
              if (gp + h < NG) {
                const int gi = cgi;
                gis[h] = gi;
                const uint32_t gs = grp_wait();
                const int i0 = GROUP * gi;
                const uint32_t dnb = sw4sel4(dn, i0 >> 5) >> (i0 & 31);  // spins of the group's rows (one word: GROUP divides 32)
                const int nval = N - i0;                                  // rows of the group inside the table
                // rows in pairs: both rows' loads first, then the two products interleaved
                static_assert(GROUP % 2 == 0, "rows are paired");
#pragma unroll
                for (int sl = 0; sl < GROUP; sl += 2) {
                  double g0[NV], g1[NV];  // (a row past the table's end is a harmless read of the ring)
                  const uint32_t r0 = gs + (uint32_t)(sl * row64), r1 = r0 + (uint32_t)row64;
                  prod::load_row_s<NF64, TL>(r0 + 16u * lane_o, r0 + (uint32_t)LD::TAIL_OFF + (uint32_t)LD::TAIL_LANE * lane_o, g0);
                  prod::load_row_s<NF64, TL>(r1 + 16u * lane_o, r1 + (uint32_t)LD::TAIL_OFF + (uint32_t)LD::TAIL_LANE * lane_o, g1);
                  double p0, p1;
                  switch ((dnb >> sl) & 3u) {
                    case 0u: lane_product64_x2<NV, false, false>(A, Bv, g0, g1, p0, p1); break;
                    case 1u: lane_product64_x2<NV, true, false>(A, Bv, g0, g1, p0, p1); break;
                    case 2u: lane_product64_x2<NV, false, true>(A, Bv, g0, g1, p0, p1); break;
                    default: lane_product64_x2<NV, true, true>(A, Bv, g0, g1, p0, p1); break;
                  }
                  if (sl < nval) v[GROUP * h + sl] = p0;
                  if (sl + 1 < nval) v[GROUP * h + sl + 1] = p1;
                }
                grp_done();
              }
            }
            const double tot = prod::bfly<double, 8>(v, lane_o);
            int mygi = gis[0];
#pragma unroll
            for (int h = 1; h < GPB; ++h) mygi = (myidx / GROUP == h) ? gis[h] : mygi;
            const int mys = GROUP * mygi + (myidx % GROUP);
            if ((lane_o & 1) == 0 && lane_o < 16 && mygi >= 0 && mys < N) {
              const double2 cst = rcd[mys];
              off_l += tot * (((sw4sel4(dn, mys >> 5) >> (mys & 31)) & 1u) ? cst.x : cst.y) / nrm;
            }
          }
          const double e_loc = JJ * (double)zz - hh * warp_sum(off_l);
          if (lane == 0) {
            store_as<double>(s.eloc_out, o, e_loc, s.eloc_dtype);
            if (want_stats) {
              const double d = e_loc - s.stats_shift;
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
      refresh_shadow(A, Bv);
      __syncwarp();
      prof_s0 = NK_PROF_NOW();
      NK_PROF_ADD(4, prof_s0 - pt1);
      NK_PROF_ADD(9, prof_wait - pwt1);
      NK_PROF_ADD(7, 1);
    };

    // ---- S: the proposal loop of sweep_fast.cu on the shadow, with the fp64 re-decision inside the band
    int in_sweep = 0;
    for (int tt = 0; tt < T_total; tt += 32) {
      {
        int site_l = 0, thr_l = 0;
        if (tt + lane < T_total) {
          uint32_t w0;
          double u;
          if (s.stream_w0 != nullptr) {
            w0 = s.stream_w0[(size_t)(tt + lane) * s.B + chain];
            u = reinterpret_cast<const double *>(s.stream_u)[(size_t)(tt + lane) * s.B + chain];
          } else {
            const uint4 w = philox_words(s.seed, s.t0 + (uint64_t)(tt + lane), gchain, STREAM_STEP);
            w0 = w.x;
            u = uniform_from_words<double>(w);
          }
          site_l = (int)__umulhi(w0, (uint32_t)N);
          thr_l = THR_MIN;  // u == 0 or machine_pow == 0: always accept
          if (pw > 0.0 && u > 0.0) {
            const double t2 = log2(u) * inv_pw * (double)FX_SCALE;
            thr_l = t2 > (double)THR_MIN ? __double2int_rn(t2) : THR_MIN;
          }
        }
        const int2 rc = rcx[site_l];
        uint4 rec;
        rec.x = g_s + (uint32_t)site_l * (uint32_t)(MP * 4);
        rec.y = sig_s + (uint32_t)site_l;
        rec.z = (uint32_t)thr_l - (uint32_t)rc.x;
        rec.w = (uint32_t)rc.y;
        __syncwarp();
        sts128u(rec_s + 16u * lane_o, rec);
        __syncwarp();
      }
      const int nb = min(32, T_total - tt);
      int k = 0;
      // forced: 0 none, 1 / 2 = the proposal at k was re-decided in double precision (reject / accept)
      int forced = 0;
      while (k < nb) {
        const int kend = k + min(nb - k, sweep_size - in_sweep);
        const int k0 = k;
        bool need_exact = false;
        uint4 nrec = lds128u(rec_s + 16u * k);
        for (; k < kend; ++k) {
          const uint4 rec = nrec;  // the next record is fetched a proposal ahead (past the last one: the slot's spare entry, unused)
          nrec = lds128u(rec_s + 16u * (k + 1));
          const uint32_t sdown = lds_u8(rec.y);
          float2 g2[NPA];
          float gt = 1.0f;
          LM::load_row(rec.x + lane16, rec.x + tailoff, g2, gt);
          float P;
          if (sdown)
            P = lane_product<NP2, NPA, HAS_T>(c.B2, c.Bt, c.A2, c.At, g2, gt);  // spin down (nu = +1): prod (B g + A)
          else
            P = lane_product<NP2, NPA, HAS_T>(c.A2, c.At, c.B2, c.Bt, g2, gt);  // spin up (nu = -1): prod (A g + B)
          const int Rp = __reduce_add_sync(FULL, __float2int_rn(lg2_fast(P) * FX_SCALE));
          const int margin = (int)((uint32_t)Rp - (uint32_t)c.R + (sdown ? rec.w : 0u - rec.w) - rec.z);  // > 0 <=> accept
          // the common case first (~96 % of the proposals are accepted well clear of the band): one compare and branch.
          // Inside the shadow's error band: leave the loop, decide in double precision (a function call: kept out of the hot
          // loop, whose registers it would otherwise push to the stack), come back to this proposal with the verdict
          bool acc = true;
          if (margin <= BAND) {
            if (margin < -BAND) {
              acc = false;
            } else {
              if (forced == 0) {
                need_exact = true;
                break;
              }
              acc = forced == 2;
              forced = 0;
            }
          }
          if (acc) {
            if (sdown) {
#pragma unroll
              for (int qq = 0; qq < NP2; ++qq) c.B2[qq] = fmul2(c.B2[qq], g2[qq]);
              if (HAS_T) c.Bt *= gt;
            } else {
#pragma unroll
              for (int qq = 0; qq < NP2; ++qq) c.A2[qq] = fmul2(c.A2[qq], g2[qq]);
              if (HAS_T) c.At *= gt;
            }
            c.R = Rp;
            sts_u8(rec.y, sdown ^ 1u);
            if (++c.nacc == c.next_renorm) renormalise32();
          }
        }
        in_sweep += k - k0;
        if (need_exact) {
          const uint4 rec = lds128u(rec_s + 16u * k);
          const uint32_t sdown = lds_u8(rec.y);
          const int site = (int)(rec.y - sig_s);
          double u;
          if (s.stream_w0 != nullptr)
            u = reinterpret_cast<const double *>(s.stream_u)[(size_t)(tt + k) * s.B + chain];
          else
            u = uniform_from_words<double>(philox_words(s.seed, s.t0 + (uint64_t)(tt + k), gchain, STREAM_STEP));
          const long long pe0 = NK_PROF_NOW();
          const bool ok = pw > 0.0 ? __any_sync(FULL, exact_accept<NF64, TL>(p, park, sig_s, sigp_s, site, sdown != 0u, u, pw, lane_o)) != 0 : true;
          forced = ok ? 2 : 1;
          NK_PROF_ADD(1, NK_PROF_NOW() - pe0);
          NK_PROF_ADD(6, 1);
          continue;
        }
        if (in_sweep == sweep_size) {
          in_sweep = 0;
          end_of_sweep();
        }
      }
    }
    // ---- write the chain state back (the double state is current: the last sweep ended with U)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int idx = 32 * b + lane;
      if (idx < N) s.sigma[(size_t)chain * N + idx] = lds_u8(sig_s + idx) ? (int8_t)-1 : (int8_t)1;
    }
    double lp;
    {
      double A[NV], Bv[NV];
      park_load(A, Bv);
      lp = logpsi64(A, Bv);
    }
    if (lane == 0) {
      reinterpret_cast<double *>(s.log_prob)[chain] = pw * lp;
      s.n_accepted[chain] += (int64_t)c.nacc;
      if (want_stats && CL > 0) {
        const double m = ws[11] / (double)CL;
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
    __syncwarp();
  }
  if (want_stats) {
    __syncthreads();
    if (tid < NK_STATS_NPARTIAL) {
      double sum = 0.0;
      for (int w = 0; w < WARPS; ++w) sum += wstat[w * WSTAT + tid];
      atomicAdd(s.stats_out + tid, sum);
    }
  }
}

}  // namespace shadow

// ------------------------------------------------------------------------------------------ host side
struct ShadowShape {
  int nf32, tl, mp32;
};

static bool shadow_shape(int M, ShadowShape *ss) {
  if (M < 1) return false;
  const int nf = M / 128, rem = M % 128;
  if (rem > 32 || nf > 4 || (nf == 4 && rem != 0) || (nf == 0 && rem == 0)) return false;
  ss->nf32 = nf;
  ss->tl = rem > 0 ? 1 : 0;
  ss->mp32 = 128 * nf + 32 * ss->tl;
  return true;
}

bool sweep_shadow_supported(const SweepKernelArgs &a, const ProdLayout &L) {
  ShadowShape ss;
  if (getenv("NKB200_NO_SHADOW") != nullptr) return false;  // developer switch: the one-table fp64 kernel
  if (a.rbm.dtype != NK_F64 || a.rule != NK_RULE_LOCAL || a.eloc_only || a.eloc_kind == 2) return false;
  if (a.rbm.N > 128 || L.kw != 1 || !shadow_shape(a.rbm.M, &ss)) return false;
  if (L.row_bytes != ss.mp32 * 8) return false;  // the fp64 rows the prep kernel builds are padded like the fp32 ones
  if (a.sweep_size < 1 || (int64_t)(a.n_discard + a.chain_length) * a.sweep_size >= (1ll << 31) || a.B >= (1ll << 31)) return false;
  const int E = a.eloc_kind == 1 ? a.ising.n_edges : 0;
  if (E > 0 && a.rbm.N > 256) return false;
  const shadow::Layout Ls = shadow::make_layout(a.rbm.N, ss.mp32, L.row_bytes, E);
  if (Ls.total > 227 * 1024) return false;
  // the stream position is a 32-bit counter
  const int64_t rounds = (a.B + (int64_t)num_sms() * shadow::CWARPS - 1) / ((int64_t)num_sms() * shadow::CWARPS);
  if (rounds * (a.n_discard + 2ll * a.chain_length) * a.rbm.N >= (1ll << 31)) return false;
  return true;
}

template <int NF32, int TL>
static int launch_shadow(cudaStream_t stream, const ProdArgs &pa, int give, int smem, double *park) {
  auto kern = shadow::sweep_shadow_kernel<NF32, TL>;
  NK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int64_t need = (pa.s.B + shadow::CWARPS - 1) / shadow::CWARPS;
  const int64_t cap = num_sms();
  kern<<<(int)(need < cap ? need : cap), shadow::THREADS, smem, stream>>>(pa, give, park);
  NK_LAUNCH_OK();
  return NK_OK;
}

size_t sweep_shadow_park_bytes() { return (size_t)(num_sms() + 8) * shadow::WARPS * 32 * 2 * 17 * sizeof(double); }

int sweep_shadow(cudaStream_t stream, const ProdArgs &pa, int give, double *park) {
  ShadowShape ss;
  if (!shadow_shape(pa.s.rbm.M, &ss)) {
    set_error("sweep_shadow: unsupported M=%d", pa.s.rbm.M);
    return NK_EUNSUPPORTED;
  }
  const int E = pa.s.eloc_kind == 1 ? pa.s.ising.n_edges : 0;
  const int smem = shadow::make_layout(pa.s.rbm.N, ss.mp32, pa.L.row_bytes, E).total;
#define NK_SH_CASE(NF, T) \
  if (ss.nf32 == NF && ss.tl == T) return launch_shadow<NF, T>(stream, pa, give, smem, park);
  NK_SH_CASE(0, 1)
  NK_SH_CASE(1, 0)
  NK_SH_CASE(1, 1)
  NK_SH_CASE(2, 0)
  NK_SH_CASE(2, 1)
  NK_SH_CASE(3, 0)
  NK_SH_CASE(3, 1)
  NK_SH_CASE(4, 0)
#undef NK_SH_CASE
  set_error("sweep_shadow: no instantiation for M=%d", pa.s.rbm.M);
  return NK_EUNSUPPORTED;
}

}  // namespace nk

#ifdef NK_SH_PROFILE
extern "C" int nk_debug_shadow_profile(unsigned long long *out16) {
  unsigned long long z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(out16, nk::shadow::nk_sh_prof, sizeof(z)) != cudaSuccess) return -1;
  return cudaMemcpyToSymbol(nk::shadow::nk_sh_prof, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
#endif
