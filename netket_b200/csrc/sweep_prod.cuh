// Product-form Metropolis sweep, general version: fp32 / fp64, LocalRule / ExchangeRule, optionally fused with the
// local energy of an Ising operator or of a LocalOperator made of 1- and 2-site terms.
//
// Replaces the hot loop of netket/sampler/metropolis.py:427-462 with rules/local.py:40-49 or rules/exchange.py:143-184
// and, when fused, netket/vqs/mc/kernels.py:62-71 with operator/_ising/jax.py:125-165 or
// operator/_local_operator/jax.py:74-201 (connected configurations are never materialised).
//
// Math ("exponential form", see sweep_fast.cu for the fp32 LocalRule specialisation).  Every hidden unit is an
// unnormalised positive pair (A_j, B_j) ~ (exp(theta_j), exp(-theta_j)) and the weight table is G_ij = exp(-4 W_ij).
// Setting sigma_i to +1 (nu_i = +1) multiplies B_j by G_ij, setting it to -1 multiplies A_j by G_ij, both up to the
// common factor exp(2 W_ij); so for any change of a set S of sites
//     psi(sigma') / psi(sigma) = exp( sum_{i in S} (2 sum_j W_ij + 2 nu_i a_i) )
//                                * prod_j (A_j GA_j + B_j GB_j) / prod_j (A_j + B_j),
//     GA_j = prod_{i in S, nu_i = -1} G_ij,    GB_j = prod_{i in S, nu_i = +1} G_ij.
// One proposal costs one FMA + one MUL per hidden unit and table row, no transcendental, and - all terms being
// positive - no cancellation.  The acceptance test is an integer comparison of fixed-point log2 values
// (one lg2.approx per lane + one REDUX per proposal); in fp64 a proposal whose test lands inside the error band of
// that approximation is re-decided in full double precision, so fp64 chains are decided in fp64 arithmetic.
//
// One warp owns one chain for the whole call; lanes own hidden units; the G table is built once per call by a prep
// kernel (global memory, padded rows) and staged into shared memory by one TMA bulk copy per CTA.  When the table
// does not fit (fp64 at N=100, M=400: 325 KB) the first `n_res` rows are resident and the others are read through L2.
// MULTI instantiations (M > 512, or an fp32 table larger than shared memory): kw warps cooperate on one chain, each
// owning a slice of the hidden units; per-proposal partial sums are combined through double-buffered shared-memory slots
// and one named barrier.  N <= 1024 sites (spins as bit words distributed over the lanes), <= 2048 exchange clusters.
// Optional outputs per recorded sample: sigma, log-probability, local energy, tanh(theta) (for the forces).
#pragma once

#include <type_traits>

#include "kernels.cuh"

namespace nk {

constexpr float PROD_FX_SCALE = 524288.0f;  // 2^19 fixed-point scale of log2 values
constexpr int PROD_FX_BAND = 64;            // fp64: |decision margin| below this many fixed-point units -> exact re-decision
constexpr float PROD_EXP_RANGE = 120.0f;    // log2 headroom allowed for a lane product
constexpr int PROD_ADJ_MAX = 32;            // clusters per site the exchange tables hold
constexpr int PROD_HOP_WORDS = 64;          // hoppable-cluster bit words per warp (C <= 2048)
constexpr int PROD_AUX_MAX = 72 * 1024;     // upper bound of the auxiliary table blob (workspace reservation)

// Byte layout of the per-CTA shared memory: [G rows | aux blob | hop words per warp | mbarrier].
// The aux blob is built in global memory by the prep kernels with exactly this layout and bulk-copied.
struct ProdLayout {
  int row_bytes, n_res, g_bytes, aux_bytes;
  int rc_off, rc_stride;           // per-site constants (RcF / RcD)
  int lg_off;                      // exchange: fix(log2(n) / machine_pow), n = 0..C
  int cl_off;                      // exchange: clusters as uint16 pairs
  int adjdeg_off, adj_off;         // exchange: clusters per site: degree (uint8), entries (cluster | partner << 16)
  int cp_off;                      // exchange with probabilities=: the clusters' weights (double), staged with the other tables
  int edges_off;                   // Ising: edges as uint16 pairs
  int lop_sites_off[2], lop_diag_off[2], lop_mel_off[2], lop_code_off[2];  // LocalOperator, compact tables
  int hop_off, bar_off, smem_bytes;
  int warps;            // warps per CTA actually launched (groups * kw)
  int kw, mw;           // warps cooperating on one chain, hidden units owned by each of them (kw == 1: mw == M)
  int seg_bytes;        // bytes of one warp's segment of a table row (row_bytes = kw * seg_bytes)
  int xs_off;           // MULTI kernels: cross-warp exchange slots, 2 * kw * 32 * 8 bytes per group
  int multi;            // 1: the MULTI instantiation is launched
};

struct ProdArgs {
  SweepKernelArgs s;
  const unsigned char *gtab;  // [N][row_bytes]
  const unsigned char *aux;   // aux blob (global)
  const void *theta;          // [B][M] T
  int *flags;                 // [0] hand over to the generic kernel, [1] renormalisation period, [2] max|W| bits, [3] max row sum |W| bits,
                              // [6] number of logarithms per lane product (1, or 2 = wide mode), [7] fp64 E_loc in (mantissa, exponent) form
  ProdLayout L;
  // chaining behind the tuned fp32 kernel: run only if *run_if != 0 (NULL: always); flags[giveup] = 1 tells the theta-form
  // kernel queued behind to take over
  const int *run_if;
  int giveup;
};

struct RcF {  // per-site constants, fp32 kernels (log2 units)
  float x2, y2;  // log2(e) * 2 sum_j W_ij,  log2(e) * 2 a_i
  int fx, fy;    // the same in fixed point
};
struct RcD {  // fp64 kernels
  double xn, yn;  // 2 sum_j W_ij, 2 a_i (natural-log units)
  double ep, em;  // exp(xn + yn), exp(xn - yn)
  int fx, fy, pad0, pad1;
};

// ---------------------------------------------------------------------------------------------- PTX helpers
namespace prod {

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done)
                 : "r"(s32(bar)), "r"(parity)
                 : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(s32(bar))
               : "memory");
}

typedef unsigned long long u64;
__device__ __forceinline__ float2 vfma(float2 a, float2 b, float2 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<u64 *>(&a)), "l"(*reinterpret_cast<u64 *>(&b)),
      "l"(*reinterpret_cast<u64 *>(&c)));
  return *reinterpret_cast<float2 *>(&d);
}
__device__ __forceinline__ float2 vmul(float2 a, float2 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<u64 *>(&a)), "l"(*reinterpret_cast<u64 *>(&b)));
  return *reinterpret_cast<float2 *>(&d);
}
__device__ __forceinline__ double vfma(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ double vmul(double a, double b) { return a * b; }
__device__ __forceinline__ float hprod(float2 v) { return v.x * v.y; }
__device__ __forceinline__ double hprod(double v) { return v; }
__device__ __forceinline__ float2 vsum2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double vsum2(double a, double b) { return a + b; }

__device__ __forceinline__ float lg2_fast(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// fixed-point log2 of a positive normal number
__device__ __forceinline__ int fxlog(float P) { return __float2int_rn(lg2_fast(P) * PROD_FX_SCALE); }
__device__ __forceinline__ int fxlog(double P) {
  // exponent exactly, mantissa in [1, 2) through lg2.approx (absolute error 2^-22): at most ~0.7 fixed-point units
  const int hi = __double2hiint(P), lo = __double2loint(P);
  const int e = ((hi >> 20) & 0x7ff) - 1023;
  const double m = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, lo);
  return (e << 19) + __float2int_rn(lg2_fast((float)m) * PROD_FX_SCALE);
}

// position of the (k+1)-th set bit of w (k < popc(w))
__device__ __forceinline__ int kth_set_bit(uint32_t w, int k) {
  int pos = 0;
  int c = __popc(w & 0xFFFFu);
  if (k >= c) { k -= c; pos += 16; w >>= 16; }
  c = __popc(w & 0xFFu);
  if (k >= c) { k -= c; pos += 8; w >>= 8; }
  c = __popc(w & 0xFu);
  if (k >= c) { k -= c; pos += 4; w >>= 4; }
  c = __popc(w & 0x3u);
  if (k >= c) { k -= c; pos += 2; w >>= 2; }
  c = (int)(w & 1u);
  if (k >= c) pos += 1;
  return pos;
}

template <typename T>
struct VecOf;
template <>
struct VecOf<float> {
  typedef float2 V;
  typedef RcF Rc;
};
template <>
struct VecOf<double> {
  typedef double V;
  typedef RcD Rc;
};

// Which hidden units a lane owns, and how a G row is read.  A row is NFULL chunks of 512 bytes (one 128-bit load per
// lane per chunk) followed by a tail of TAIL elements per lane.
template <typename T, int NFULL, int TAIL>
struct LaneMap {
  typedef typename VecOf<T>::V V;
  static constexpr int EPC = 16 / (int)sizeof(T);                       // elements per 128-bit load
  static constexpr int NE = EPC * NFULL + TAIL;                         // elements per lane
  static constexpr int NV = sizeof(T) == 4 ? (NE + 1) / 2 : NE;         // V registers per lane
  static constexpr int MP = 32 * EPC * NFULL + 32 * TAIL;               // padded row length
  static constexpr int ROW_BYTES = MP * (int)sizeof(T);
  static constexpr int TAIL_OFF = 512 * NFULL;                          // byte offset of the tail inside a row
  static constexpr int TAIL_LANE = TAIL * (int)sizeof(T);               // tail bytes per lane
  static __host__ __device__ __forceinline__ int unit(int e, int lane) {
    return e < EPC * NFULL ? 32 * EPC * (e / EPC) + EPC * lane + (e % EPC) : 32 * EPC * NFULL + TAIL * lane + (e - EPC * NFULL);
  }
};

template <int NFULL, int TAIL>
__device__ __forceinline__ void load_row_s(uint32_t row_lane, uint32_t tail_lane, float2 (&g)[LaneMap<float, NFULL, TAIL>::NV]) {
#pragma unroll
  for (int q = 0; q < NFULL; ++q) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(row_lane + 512 * q));
    g[2 * q] = make_float2(v.x, v.y);
    g[2 * q + 1] = make_float2(v.z, v.w);
  }
  if (TAIL == 1) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(tail_lane));
    g[2 * NFULL] = make_float2(v, 1.0f);  // odd element count: neutral partner (G = 1 with A = B = 1/2)
  }
  if (TAIL == 2) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(tail_lane));
    g[2 * NFULL] = v;
  }
}
template <int NFULL, int TAIL>
__device__ __forceinline__ void load_row_s(uint32_t row_lane, uint32_t tail_lane, double (&g)[LaneMap<double, NFULL, TAIL>::NV]) {
#pragma unroll
  for (int q = 0; q < NFULL; ++q)
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(g[2 * q]), "=d"(g[2 * q + 1]) : "r"(row_lane + 512 * q));
  if (TAIL == 1) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(g[2 * NFULL]) : "r"(tail_lane));
}
// generic-address loads: one code path for rows resident in shared memory and rows read through L2 (the address
// is selected, not the instruction, so nothing is issued twice under predication)
template <int NFULL, int TAIL>
__device__ __forceinline__ void load_row_p(const unsigned char *row_lane, const unsigned char *tail_lane,
                                           float2 (&g)[LaneMap<float, NFULL, TAIL>::NV]) {
#pragma unroll
  for (int q = 0; q < NFULL; ++q) {
    float4 v;
    asm volatile("ld.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(row_lane + 512 * q));
    g[2 * q] = make_float2(v.x, v.y);
    g[2 * q + 1] = make_float2(v.z, v.w);
  }
  if (TAIL == 1) {
    float v;
    asm volatile("ld.f32 %0, [%1];" : "=f"(v) : "l"(tail_lane));
    g[2 * NFULL] = make_float2(v, 1.0f);
  }
  if (TAIL == 2) {
    float2 v;
    asm volatile("ld.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(tail_lane));
    g[2 * NFULL] = v;
  }
}
template <int NFULL, int TAIL>
__device__ __forceinline__ void load_row_p(const unsigned char *row_lane, const unsigned char *tail_lane,
                                           double (&g)[LaneMap<double, NFULL, TAIL>::NV]) {
#pragma unroll
  for (int q = 0; q < NFULL; ++q)
    asm volatile("ld.v2.f64 {%0, %1}, [%2];" : "=d"(g[2 * q]), "=d"(g[2 * q + 1]) : "l"(row_lane + 512 * q));
  if (TAIL == 1) asm volatile("ld.f64 %0, [%1];" : "=d"(g[2 * NFULL]) : "l"(tail_lane));
}

// A lane product as its two interleaved accumulators (a * b is the product).  Keeping them apart lets the fp32 kernels take
// two logarithms instead of one when the weights are large ("wide" mode): each partial product then spans half the exponent
// range, which doubles the largest |W| the product form can handle.
template <typename T>
struct LanePair {
  T a, b;
};
// prod_q hprod(X_q * g_q + Y_q)
template <typename V, int NV>
__device__ __forceinline__ auto lane_product(const V (&X)[NV], const V (&Y)[NV], const V (&g)[NV]) -> LanePair<decltype(hprod(X[0]))> {
  typedef decltype(hprod(X[0])) T;
  V Pa = vfma(X[0], g[0], Y[0]);
  if (NV == 1) return LanePair<T>{hprod(Pa), T(1)};
  V Pb = vfma(X[NV > 1 ? 1 : 0], g[NV > 1 ? 1 : 0], Y[NV > 1 ? 1 : 0]);
#pragma unroll
  for (int q = 2; q < NV; ++q) {
    const V c = vfma(X[q], g[q], Y[q]);
    if (q & 1)
      Pb = vmul(Pb, c);
    else
      Pa = vmul(Pa, c);
  }
  return LanePair<T>{hprod(Pa), hprod(Pb)};
}
// prod_q hprod(c_q)
template <typename V, int NV>
__device__ __forceinline__ auto lane_product1(const V (&c)[NV]) -> LanePair<decltype(hprod(c[0]))> {
  typedef decltype(hprod(c[0])) T;
  V Pa = c[0];
  if (NV == 1) return LanePair<T>{hprod(Pa), T(1)};
  V Pb = c[NV > 1 ? 1 : 0];
#pragma unroll
  for (int q = 2; q < NV; ++q) {
    if (q & 1)
      Pb = vmul(Pb, c[q]);
    else
      Pa = vmul(Pa, c[q]);
  }
  return LanePair<T>{hprod(Pa), hprod(Pb)};
}
// fixed-point log2 / log2 / value of a lane product
__device__ __forceinline__ int lp_fx(const LanePair<float> &p, bool wide) { return wide ? fxlog(p.a) + fxlog(p.b) : fxlog(p.a * p.b); }
__device__ __forceinline__ int lp_fx(const LanePair<double> &p, bool) { return fxlog(p.a * p.b); }  // a double never leaves its range here
__device__ __forceinline__ float lp_lg2(const LanePair<float> &p, bool wide) { return wide ? lg2_fast(p.a) + lg2_fast(p.b) : lg2_fast(p.a * p.b); }
template <typename T>
__device__ __forceinline__ T lp_val(const LanePair<T> &p) { return p.a * p.b; }

__device__ __forceinline__ uint32_t sel4(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, int i) {
  return (i < 2) ? ((i == 0) ? w0 : w1) : ((i == 2) ? w2 : w3);
}

__device__ __forceinline__ double warp_prod(double v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

// fp64, once per chain and call: kept out of line so that the transcendental code is not replicated per hidden unit
// (the kernel's hot loops stay small for the instruction cache)
static __device__ __noinline__ double2 pair_from_theta(double x) {  // (e^x, e^-x) / (2 cosh x)
  const double ex = exp(-2.0 * fabs(x));
  const double big = 1.0 / (1.0 + ex), small = ex * big;
  return x >= 0.0 ? make_double2(big, small) : make_double2(small, big);
}
static __device__ __noinline__ double lncosh_from_pair(double a, double b) {  // log((a + b) / (2 sqrt(a b)))
  return log(a + b) - 0.5 * (log(a) + log(b)) - 0.69314718055994530942;
}
static __device__ __noinline__ double rcp_sum(double a, double b) { return 1.0 / (a + b); }

// fp64 only, rare: the decision of a proposal whose fixed-point test fell inside the approximation's error band.
//   accept  <=>  u < exp(machine_pow * (cst + sum_lanes log(Pprop / Pcur)) + corr)      (metropolis.py:444-450)
static __device__ __noinline__ double exact_logratio(double Pprop, double Pcur) {
  double d = log(Pprop / Pcur);
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) d += __shfl_xor_sync(0xffffffffu, d, m);
  return d;
}
static __device__ __noinline__ bool exact_decide(double d, double cst, double u, double pw, double corr) {
  return u < exp(pw * (cst + d) + corr);
}

// NB values per lane (partials of NB candidates, NB = 8 or 16) -> lane l holds the warp total of candidate
// (l >> 1) & (NB - 1).  fp32: additive (log2 partials); fp64: multiplicative (the lane products themselves).
__device__ __forceinline__ float bf_op(float a, float b) { return a + b; }
__device__ __forceinline__ double bf_op(double a, double b) { return a * b; }
__device__ __forceinline__ int bf_op(int a, int b) { return a + b; }  // exponents of split doubles
template <typename T, int NB>
__device__ __forceinline__ T bfly(T (&v)[NB], int lane) {
#pragma unroll
  for (int h = NB / 2; h >= 1; h >>= 1) {
    const bool up = (lane & (2 * h)) != 0;
#pragma unroll
    for (int jj = 0; jj < h; ++jj) {
      const T send = up ? v[jj] : v[jj + h];
      const T keep = up ? v[jj + h] : v[jj];
      v[jj] = bf_op(keep, __shfl_xor_sync(0xffffffffu, send, 2 * h));
    }
  }
  T tot = bf_op(v[0], __shfl_xor_sync(0xffffffffu, v[0], 1));
#pragma unroll
  for (int m = 2 * NB; m < 32; m <<= 1) tot = bf_op(tot, __shfl_xor_sync(0xffffffffu, tot, m));
  return tot;
}

// candidate descriptor (one connected configuration): sites, which of them change and to which sign
//   bits 0-9 s0, 10-19 s1, 20 chg0, 21 chg1, 22 pos0 (sigma'_{s0} = +1), 23 pos1, 31 valid
constexpr uint32_t CD_CHG0 = 1u << 20, CD_CHG1 = 1u << 21, CD_POS0 = 1u << 22, CD_POS1 = 1u << 23, CD_VALID = 1u << 31;
constexpr int CD_S1_SHIFT = 10;
constexpr uint32_t CD_SITE_MASK = 1023u;

#ifndef NK_PROD_WARPS_F32L
#define NK_PROD_WARPS_F32L 20
#endif
#ifndef NK_PROD_WARPS_F32X
#define NK_PROD_WARPS_F32X 20
#endif
#ifndef NK_PROD_WARPS_F64L
#define NK_PROD_WARPS_F64L 12
#endif
#ifndef NK_PROD_WARPS_F64X
#define NK_PROD_WARPS_F64X 12
#endif
// warps per CTA (one CTA per SM): set by the register budget of each variant, 65536 / (32 * warps) registers per thread
template <typename T, int RULE>
struct ProdWarps {
  static constexpr int value = sizeof(T) == 4 ? (RULE == NK_RULE_LOCAL ? NK_PROD_WARPS_F32L : NK_PROD_WARPS_F32X)
                                              : (RULE == NK_RULE_LOCAL ? NK_PROD_WARPS_F64L : NK_PROD_WARPS_F64X);
};

}  // namespace prod

// ============================================================================================== the kernel
// MULTI: kw > 1 warps cooperate on one chain (hidden units split between them; per-proposal partial sums are combined
// through shared memory and a named barrier) and the table rows are read with generic-address loads (mostly through L2).
template <typename T, int NFULL, int TAIL, int RULE, bool MULTI>
__global__ void __launch_bounds__(prod::ProdWarps<T, RULE>::value * 32, 1) sweep_prod_kernel(const __grid_constant__ ProdArgs p) {
  using namespace prod;
  typedef LaneMap<T, NFULL, TAIL> LM;
  typedef typename LM::V V;
  typedef typename VecOf<T>::Rc Rc;
  constexpr int NV = LM::NV, NE = LM::NE, ROW_BYTES = LM::ROW_BYTES;
  constexpr bool F64 = sizeof(T) == 8;
  constexpr bool GENERIC_ROWS = F64 || MULTI;  // generic-address row loads (rows partly or wholly outside shared memory)
  constexpr uint32_t FULL = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char smem[];
  const SweepKernelArgs &s = p.s;
  const ProdLayout &L = p.L;
  if (p.run_if != nullptr && *p.run_if == 0) return;  // the kernel in front of this one did the work
  if (p.flags[p.giveup] != 0) return;  // the prep kernels found this configuration outside the product form's range
  const int renorm = p.flags[1];
  const bool wide = p.flags[6] == 2;  // fp32: two logarithms per lane product (large weights)
  const bool wide_e = F64 && p.flags[7] != 0;  // fp64 local energy: products across the warp as (mantissa, exponent) pairs
  const int N = s.rbm.N, M = s.rbm.M;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(FULL, tid >> 5, 0);
  const int KW = MULTI ? L.kw : 1;                 // warps per chain
  const int grp = MULTI ? warp / KW : warp;        // chain slot of this warp inside the CTA
  const int wk = MULTI ? warp - grp * KW : 0;      // which part of the hidden layer this warp owns
  const int groups = MULTI ? L.warps / KW : L.warps;
  const int MW = MULTI ? L.mw : M;                 // hidden units of this warp: [wk * MW, min((wk + 1) * MW, M))
  const int nbw = (N + 31) >> 5;                   // sigma bit words (bits per lane)
  unsigned char *aux = smem + L.g_bytes;
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + L.bar_off);
  uint32_t *hopw = reinterpret_cast<uint32_t *>(smem + L.hop_off) + warp * PROD_HOP_WORDS;

  // ---------------- stage the G rows and the auxiliary tables (TMA bulk copies, one mbarrier)
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar, (uint32_t)(L.g_bytes + L.aux_bytes));
    // one bulk copy moves at most 2^20 - 16 bytes; chunks keep every size well inside that
    for (int off = 0; off < L.g_bytes; off += 65536) {
      const int n = min(65536, L.g_bytes - off);
      tma_bulk_g2s(s32(smem + off), p.gtab + off, (uint32_t)n, bar);
    }
    tma_bulk_g2s(s32(aux), p.aux, (uint32_t)L.aux_bytes, bar);
  }
  __syncthreads();  // the barrier is initialised before anyone waits on it
  mbar_wait(bar, 0);

  const Rc *rctab = reinterpret_cast<const Rc *>(aux + L.rc_off);
  const int *lgtab = reinterpret_cast<const int *>(aux + L.lg_off);
  const uint16_t *cl = reinterpret_cast<const uint16_t *>(aux + L.cl_off);
  // ExchangeRule(probabilities=) weights (shared-memory copy: the selection walks them bit by bit), or NULL
  const double *cprob = RULE == NK_RULE_EXCHANGE && s.cluster_probs != nullptr ? reinterpret_cast<const double *>(aux + L.cp_off) : nullptr;
  const uint8_t *adjdeg = aux + L.adjdeg_off;
  const uint32_t *adj = reinterpret_cast<const uint32_t *>(aux + L.adj_off);
  const uint16_t *edges = reinterpret_cast<const uint16_t *>(aux + L.edges_off);
  const int E = s.eloc_kind == 1 ? s.ising.n_edges : 0;
  const int C = RULE == NK_RULE_EXCHANGE ? s.n_clusters : 0;

  const int T_total = (s.n_discard + s.chain_length) * s.sweep_size;
  const T pw = (T)s.machine_pow;
  const T inv_pw = pw > T(0) ? T(1) / pw : T(0);
  const T LN2 = (T)0.69314718055994530942, LOG2E = (T)1.4426950408889634;
  // rows: the first n_res live in shared memory, the others are read from the global table (through L2).
  // fp32 single-warp kernels are only launched when the whole table is resident (explicit LDS); the others use
  // generic addresses so that one instruction stream serves both kinds of row.
  uint32_t lane_row = s32(smem) + 16u * lane;
  uint32_t lane_tail = s32(smem) + (uint32_t)LM::TAIL_OFF + (uint32_t)LM::TAIL_LANE * lane;
  // generic addresses of this lane's first element in row 0 (this warp's segment): resident copy / global table.
  // Kept opaque so that they stay in registers (no cvta / constant-bank reloads per row); offsets are 32-bit.
  const unsigned char *sbase = smem + (size_t)wk * ROW_BYTES + 16 * lane;
  const unsigned char *gbase = p.gtab + (size_t)wk * ROW_BYTES + 16 * lane;
  int tail_delta = LM::TAIL_OFF + LM::TAIL_LANE * lane - 16 * lane;  // tail element relative to the above (may be negative)
  uint32_t row_stride = (uint32_t)L.row_bytes;
  int n_res = L.n_res;
  int sweep_size = s.sweep_size;
  asm volatile("" : "+r"(lane_row), "+r"(lane_tail), "+r"(n_res), "+r"(sweep_size), "+r"(tail_delta), "+r"(row_stride));
  asm volatile("" : "+l"(sbase), "+l"(gbase));

  auto fetch = [&](int site, V(&g)[NV]) {
    if constexpr (!GENERIC_ROWS) {
      const uint32_t o = (uint32_t)site * (uint32_t)ROW_BYTES;
      load_row_s<NFULL, TAIL>(lane_row + o, lane_tail + o, g);
    } else {
      const unsigned char *row = (site < n_res ? sbase : gbase) + (uint32_t)site * row_stride;
      load_row_p<NFULL, TAIL>(row, row + (ptrdiff_t)tail_delta, g);
    }
  };

  // cross-warp combination (MULTI): every lane contributes one value and gets the combination over the KW warps of its
  // chain.  Two slot buffers alternate, so one named barrier per call suffices (a warp can only come back to the same
  // buffer after the next call's barrier, i.e. after every warp has read this call's slots).
  unsigned char *xs_group = smem + L.xs_off + (size_t)grp * (2 * KW * 256);
  int xphase = 0;
  auto gcomb = [&](auto v, auto op) {
    typedef decltype(v) U;
    if constexpr (!MULTI) {
      return v;
    } else {
      U *buf = reinterpret_cast<U *>(xs_group + xphase * (KW * 256));
      xphase ^= 1;
      buf[wk * 32 + lane] = v;
      asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(KW * 32) : "memory");
      U r = buf[lane];
      for (int k = 1; k < KW; ++k) r = op(r, buf[k * 32 + lane]);
      return r;
    }
  };
  auto op_add = [](auto a, auto b) { return a + b; };
  auto op_mul = [](auto a, auto b) { return a * b; };

  for (int chain = blockIdx.x * groups + grp; chain < (int)s.B; chain += gridDim.x * groups) {
    V A[NV], Bv[NV];
    // sigma, lane-distributed: bit b of lane l is set iff sigma of site 32 b + l is -1
    uint32_t mybits = 0;
    int R = 0, since = 0, n_hop = 0;
    uint32_t nacc = 0;
    {
      const int8_t *sg = s.sigma + (size_t)chain * N;
      for (int b = 0; b < nbw; ++b) {
        const int idx = 32 * b + lane;
        if (idx < N && sg[idx] < 0) mybits |= 1u << b;
      }
    }
    // spin bit of any site (the site may differ between lanes; all lanes must call): one shuffle
    auto sbit = [&](int site) -> uint32_t { return (__shfl_sync(FULL, mybits, site & 31) >> (site >> 5)) & 1u; };
    auto stoggle = [&](int site) {
      if (lane == (site & 31)) mybits ^= 1u << (site >> 5);
    };
    auto ownbit = [&](int b) -> uint32_t { return (mybits >> b) & 1u; };  // site 32 b + lane
    // ---- theta -> (A, B) = (e^theta, e^-theta) / (2 cosh theta)
    {
      const T *th = reinterpret_cast<const T *>(p.theta) + (size_t)chain * M;
      if constexpr (F64) {
#pragma unroll
        for (int q = 0; q < NV; ++q) {
          const int ju = LM::unit(q, lane), j = wk * MW + ju;
          const double2 ab = pair_from_theta(ju < MW && j < M ? th[j] : 0.0);  // padding units: theta = 0
          A[q] = ab.x;
          Bv[q] = ab.y;
        }
      } else {
        float av[2 * NV], bv[2 * NV];
#pragma unroll
        for (int e = 0; e < 2 * NV; ++e) {
          const int ju = e < NE ? LM::unit(e, lane) : MW, j = wk * MW + ju;
          av[e] = 0.5f;
          bv[e] = 0.5f;  // padding units: theta = 0
          if (ju < MW && j < M) {
            const float x = th[j];
            const float ex = expf(-2.0f * fabsf(x));
            const float big = 1.0f / (1.0f + ex), small = ex * big;
            av[e] = x >= 0.0f ? big : small;
            bv[e] = x >= 0.0f ? small : big;
          }
        }
#pragma unroll
        for (int q = 0; q < NV; ++q) {
          A[q] = make_float2(av[2 * q], av[2 * q + 1]);
          Bv[q] = make_float2(bv[2 * q], bv[2 * q + 1]);
        }
      }
    }
    // ---- exchange: hoppable-cluster bit words of this chain (rules/exchange.py:208-218)
    if (RULE == NK_RULE_EXCHANGE) {
      __syncwarp();
      for (int q = 0; q < PROD_HOP_WORDS; ++q) {
        const int c = 32 * q + lane;
        const int ci = c < C ? cl[2 * c] : 0, cj = c < C ? cl[2 * c + 1] : 0;
        const bool h = (sbit(ci) != sbit(cj)) && c < C;
        const uint32_t b = __ballot_sync(FULL, h);
        if (lane == 0) hopw[q] = b;
        n_hop += __popc(b);
      }
      __syncwarp();
    }
    const uint64_t gchain = s.chain_offset + (uint64_t)chain;

    auto renormalise = [&]() {  // A + B = 1
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        if constexpr (F64) {
          const double ix = rcp_sum(A[q], Bv[q]);
          A[q] *= ix;
          Bv[q] *= ix;
        } else {
          const float2 iv = make_float2(__frcp_rn(A[q].x + Bv[q].x), __frcp_rn(A[q].y + Bv[q].y));
          A[q] = vmul(A[q], iv);
          Bv[q] = vmul(Bv[q], iv);
        }
      }
      R = 0;
      since = 0;
    };
    // lane product of the current normalisation prod_j (A_j + B_j)
    auto lane_norm = [&]() -> T {
      V c[NV];
#pragma unroll
      for (int q = 0; q < NV; ++q) c[q] = vsum2(A[q], Bv[q]);
      return lp_val(lane_product1<V, NV>(c));
    };
    // log psi of the current state: lncosh(theta_j) = log((A_j + B_j) / (2 sqrt(A_j B_j)))
    auto logpsi_now = [&]() -> T {
      T acc = T(0);
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        if constexpr (F64) {
          acc += lncosh_from_pair(A[q], Bv[q]);
        } else {
          acc += LN2 * (log2f(A[q].x + Bv[q].x) - 0.5f * (log2f(A[q].x) + log2f(Bv[q].x)) - 1.0f);
          acc += LN2 * (log2f(A[q].y + Bv[q].y) - 0.5f * (log2f(A[q].y) + log2f(Bv[q].y)) - 1.0f);
        }
      }
      acc = gcomb(warp_sum(acc), op_add);  // hidden units: over the lanes and over the warps of the chain
      T vis = T(0);                         // visible bias: every warp holds all of sigma
      if (s.rbm.a != nullptr) {
        for (int b = 0; b < nbw; ++b) {
          const int idx = 32 * b + lane;
          if (idx < N) {
            T ai;
            if constexpr (F64)
              ai = 0.5 * rctab[idx].yn;
            else
              ai = 0.5f * LN2 * rctab[idx].y2;
            vis += ownbit(b) ? -ai : ai;
          }
        }
      }
      return acc + warp_sum(vis);
    };

    // ------------------------------------------------------------------ fused local energy of the current sample
    // lane product of one candidate (connected configuration) of kind KIND:
    //   0 one site flips, 1 two sites flip to opposite signs (exchange-like), 2 two sites flip to the same sign
    auto eval_cand = [&](auto kind_c, uint32_t d) -> LanePair<T> {
      constexpr int KIND = decltype(kind_c)::value;
      const int s0 = d & CD_SITE_MASK, s1 = (d >> CD_S1_SHIFT) & CD_SITE_MASK;
      const bool p0 = d & CD_POS0;
      V g[NV];
      if constexpr (KIND == 0) {
        const bool c0 = d & CD_CHG0;
        const int sa = c0 ? s0 : s1;
        const bool pa = c0 ? p0 : (d & CD_POS1) != 0;
        fetch(sa, g);
        return pa ? lane_product<V, NV>(Bv, A, g) : lane_product<V, NV>(A, Bv, g);
      } else if constexpr (KIND == 1) {
        const int sp = p0 ? s0 : s1, sm = p0 ? s1 : s0;
        V t[NV];
        fetch(sm, g);
#pragma unroll
        for (int q = 0; q < NV; ++q) t[q] = vmul(A[q], g[q]);
        fetch(sp, g);
#pragma unroll
        for (int q = 0; q < NV; ++q) t[q] = vfma(Bv[q], g[q], t[q]);
        return lane_product1<V, NV>(t);
      } else {
        V g2[NV];
        fetch(s0, g);
        fetch(s1, g2);
#pragma unroll
        for (int q = 0; q < NV; ++q) g[q] = vmul(g[q], g2[q]);
        return p0 ? lane_product<V, NV>(Bv, A, g) : lane_product<V, NV>(A, Bv, g);
      }
    };
    // constant of a candidate: fp32 log2 units (added before ex2), fp64 a multiplier
    auto cand_const = [&](uint32_t d) -> T {
      const int s0 = d & CD_SITE_MASK, s1 = (d >> CD_S1_SHIFT) & CD_SITE_MASK;
      if constexpr (F64) {
        double m = 1.0;
        if (d & CD_CHG0) m *= (d & CD_POS0) ? rctab[s0].ep : rctab[s0].em;
        if (d & CD_CHG1) m *= (d & CD_POS1) ? rctab[s1].ep : rctab[s1].em;
        return m;
      } else {
        float c = 0.0f;
        if (d & CD_CHG0) c += (d & CD_POS0) ? rctab[s0].x2 + rctab[s0].y2 : rctab[s0].x2 - rctab[s0].y2;
        if (d & CD_CHG1) c += (d & CD_POS1) ? rctab[s1].x2 + rctab[s1].y2 : rctab[s1].x2 - rctab[s1].y2;
        return c;
      }
    };
    // sum over the candidates of kind KIND held one per lane: mel * psi(sigma') / psi(sigma)   (per-lane partial sums).
    // NB candidates at a time: NB independent lane products, then one transposed butterfly.
    constexpr int NB = F64 ? 8 : 16;  // candidates per butterfly (register budget)
    const int myidx = (lane >> 1) & (NB - 1);
    // fp64: warp (and chain) total of candidate (lane >> 1) & (NB - 1).  Narrow mode: the plain product.  Wide mode (row sums
    // of |W| so large that a product of M factors could leave the double range): every lane product is split into mantissa
    // and exponent, mantissas are multiplied (32 x [1, 2) stays tiny), exponents are added in a second butterfly.
    auto reduce_cands = [&](T(&v)[NB], int &tot_e) -> T {
      tot_e = 0;
      if constexpr (F64) {
        if (wide_e) {
          int ve[NB];
#pragma unroll
          for (int jj = 0; jj < NB; ++jj) {
            const int hi = __double2hiint(v[jj]), lo = __double2loint(v[jj]);
            ve[jj] = ((hi >> 20) & 0x7ff) - 1023;
            v[jj] = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, lo);
          }
          tot_e = gcomb(bfly<int, NB>(ve, lane), op_add);
        }
      }
      return gcomb(bfly<T, NB>(v, lane), [&](T x, T y) { return bf_op(x, y); });
    };
    // psi(sigma') / psi(sigma) of one candidate from its total (fp32: log2 domain; fp64: product, or log2 domain in wide mode)
    auto cand_ratio = [&](T tot, int tot_e, T nrm, uint32_t d) -> T {
      if constexpr (F64) {
        if (wide_e) {
          const int s0 = d & CD_SITE_MASK, s1 = (d >> CD_S1_SHIFT) & CD_SITE_MASK;
          double c = 0.0;  // natural-log constant: sum over the changed sites of 2 sum_j W_ij +- 2 a_i
          if (d & CD_CHG0) c += (d & CD_POS0) ? rctab[s0].xn + rctab[s0].yn : rctab[s0].xn - rctab[s0].yn;
          if (d & CD_CHG1) c += (d & CD_POS1) ? rctab[s1].xn + rctab[s1].yn : rctab[s1].xn - rctab[s1].yn;
          return exp2(log2(tot) - log2(nrm) + (double)tot_e + 1.4426950408889634 * c);
        }
        return tot * cand_const(d) / nrm;
      } else {
        return ex2_fast(tot - nrm + cand_const(d));
      }
    };
    auto eval_round = [&](auto kind_c, uint32_t desc, T mel, T nrm, T &off_l) {
      constexpr int KIND = decltype(kind_c)::value;
      const bool both = (desc & CD_CHG0) && (desc & CD_CHG1);
      const bool opp = ((desc & CD_POS0) != 0) != ((desc & CD_POS1) != 0);
      const int kind = !both ? 0 : (opp ? 1 : 2);
      uint32_t mask = __ballot_sync(FULL, (desc & CD_VALID) != 0 && kind == KIND);
      while (mask != 0) {
        T v[NB];
        uint32_t mydesc = 0;
        T mymel = T(0);
#pragma unroll
        for (int jj = 0; jj < NB; ++jj) {
          v[jj] = F64 ? T(1) : T(0);
          if (mask != 0) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const uint32_t d = __shfl_sync(FULL, desc, src);
            const T m = __shfl_sync(FULL, mel, src);
            if (myidx == jj) {
              mydesc = d;
              mymel = m;
            }
            const LanePair<T> P = eval_cand(kind_c, d);
            if constexpr (F64)
              v[jj] = lp_val(P);
            else
              v[jj] = lp_lg2(P, wide);
          }
        }
        int tot_e;
        const T tot = reduce_cands(v, tot_e);
        if ((lane & 1) == 0 && lane < 2 * NB && (mydesc & CD_VALID)) off_l += mymel * cand_ratio(tot, tot_e, nrm, mydesc);
      }
    };
    auto local_energy = [&]() -> T {
      // fp64: A + B = 1 keeps every product of M factors inside the double range (see prep: row-sum bound); fp32: the
      // candidates' lane products must see at most renorm - 1 un-normalised accepts, like the proposals
      if (F64 || since >= renorm) renormalise();
      T nrm;                   // fp64: prod_j (A_j + B_j); fp32: its log2
      if constexpr (F64)
        nrm = gcomb(warp_prod(lane_norm()), op_mul);
      else
        nrm = gcomb(warp_sum(lg2_fast(lane_norm())), op_add);
      T off_l = T(0);
      double dl = 0.0;  // diagonal, per-lane partial
      // LocalOperator candidates are produced 32 at a time (one per lane); rounds enumerate the (term, entry) slots
      int rounds0, rounds1 = 0;
      if (s.eloc_kind == 1) {
        // E_loc = J sum_<ij> s_i s_j - h sum_i psi(sigma^(i)) / psi(sigma)        (_ising/jax.py:125-165)
        int zz = 0;
        for (int e0 = 0; e0 < E; e0 += 32) {
          const int e = e0 + lane;
          const int ea = e < E ? edges[2 * e] : 0, eb = e < E ? edges[2 * e + 1] : 0;
          const int par = (int)(sbit(ea) ^ sbit(eb));
          if (e < E) zz += 1 - 2 * par;
        }
        dl = s.ising.J * (double)zz;
        rounds0 = 0;
        // the N single flips are a static candidate list: no descriptors to shuffle.  NB sites at a time, spins from one
        // ballot per 32 sites, NB independent lane products, then the transposed butterfly.
        if (s.ising.h != 0.0) {
          const T mel = (T)(-s.ising.h);
          uint32_t word = 0;
          for (int base = 0; base < N; base += NB) {
            if ((base & 31) == 0) word = __ballot_sync(FULL, (mybits >> (base >> 5)) & 1u);
            const uint32_t bits = word >> (base & 31);
            T v[NB];
#pragma unroll
            for (int jj = 0; jj < NB; ++jj) {
              v[jj] = F64 ? T(1) : T(0);
              const int site = base + jj;
              if (site < N) {
                V g[NV];
                fetch(site, g);
                const LanePair<T> P = ((bits >> jj) & 1u) ? lane_product<V, NV>(Bv, A, g) : lane_product<V, NV>(A, Bv, g);
                if constexpr (F64)
                  v[jj] = lp_val(P);
                else
                  v[jj] = lp_lg2(P, wide);
              }
            }
            int tot_e;
            const T tot = reduce_cands(v, tot_e);
            const int mys = base + myidx;
            if ((lane & 1) == 0 && lane < 2 * NB && mys < N) {
              const uint32_t dd = CD_CHG0 | (uint32_t)mys | (((bits >> myidx) & 1u) ? CD_POS0 : 0u);
              off_l += mel * cand_ratio(tot, tot_e, nrm, dd);
            }
          }
        }
      } else {
        // LocalOperator: diagonal = constant + sum_terms diag_mels[row]; off-diagonal entries with |mel| > cutoff
        // (_local_operator/jax.py:104-199; the compaction only reorders, the sum runs over the same entries)
        const nk_localop_group_t &G0 = s.localop.groups[0];
        const nk_localop_group_t &G1 = s.localop.groups[1];
        rounds0 = s.localop.n_groups > 0 ? (G0.n_ops * max(G0.ncmax, 1) + 31) / 32 : 0;
        rounds1 = s.localop.n_groups > 1 ? (G1.n_ops * max(G1.ncmax, 1) + 31) / 32 : 0;
      }
      for (int r = 0; r < rounds0 + rounds1; ++r) {
        uint32_t d = 0;
        T mel = T(0);
        {  // LocalOperator only: the Ising candidates were handled above (rounds0 = rounds1 = 0)
          const int gi = r < rounds0 ? 0 : 1;
          const nk_localop_group_t &G = s.localop.groups[gi];
          const int rows = 1 << G.n_sites, ncm = G.ncmax;
          const uint16_t *sites = reinterpret_cast<const uint16_t *>(aux + L.lop_sites_off[gi]);
          const T *dg = reinterpret_cast<const T *>(aux + L.lop_diag_off[gi]);
          const T *ml = reinterpret_cast<const T *>(aux + L.lop_mel_off[gi]);
          const uint8_t *cd = aux + L.lop_code_off[gi];
          const int slots = G.n_ops * max(ncm, 1);
          const int q = 32 * (gi == 0 ? r : r - rounds0) + lane;
          const bool live = q < slots;
          const int o = live ? (ncm > 1 ? q / ncm : q) : 0, c = live && ncm > 1 ? q - o * ncm : 0;
          const int s0 = sites[2 * o], s1 = sites[2 * o + 1];
          const int x0 = (int)sbit(s0), x1b = (int)sbit(s1);  // all lanes shuffle
          if (live) {
            const int x1 = G.n_sites == 2 ? x1b : 0;
            const int row = G.n_sites == 2 ? 2 * x0 + x1 : x0;  // _state_to_number: first site most significant
            if (c == 0) dl += (double)dg[o * rows + row];
            if (ncm > 0) {
              const int code = cd[(o * rows + row) * ncm + c];  // bit 0 valid, bit 1 x'_0, bit 2 x'_1
              if (code & 1) {
                mel = ml[(o * rows + row) * ncm + c];
                const int xp0 = (code >> 1) & 1, xp1 = (code >> 2) & 1;
                const bool ch0 = xp0 != x0, ch1 = G.n_sites == 2 && xp1 != x1;
                if (!ch0 && !ch1) {
                  off_l += mel;  // an entry that maps sigma onto itself: ratio 1
                } else {
                  d = CD_VALID | (uint32_t)s0 | ((uint32_t)s1 << CD_S1_SHIFT) | (ch0 ? CD_CHG0 : 0u) | (ch1 ? CD_CHG1 : 0u) |
                      (xp0 == 0 ? CD_POS0 : 0u) | (xp1 == 0 ? CD_POS1 : 0u);
                }
              }
            }
          }
        }
        eval_round(std::integral_constant<int, 0>{}, d, mel, nrm, off_l);
        eval_round(std::integral_constant<int, 1>{}, d, mel, nrm, off_l);
        eval_round(std::integral_constant<int, 2>{}, d, mel, nrm, off_l);
      }
      const double diag = warp_sum(dl) + (s.eloc_kind == 2 ? s.localop.constant : 0.0);
      T acc = warp_sum(off_l);
      if (s.eloc_kind == 1 || (s.localop.nonzero_diagonal && fabs(diag) > s.localop.mel_cutoff)) acc += (T)diag;
      return acc;
    };

    int in_sweep = 0, sweep_idx = 0;
    auto end_of_sweep = [&]() {
      in_sweep = 0;
      const int sw = sweep_idx - s.n_discard;
      ++sweep_idx;
      if (sw < 0) return;
      const size_t o = (size_t)chain * s.chain_length + sw;
      if (s.samples_out != nullptr && wk == 0) {
        for (int b = 0; b < nbw; ++b) {
          const int idx = 32 * b + lane;
          if (idx < N) s.samples_out[o * N + idx] = ownbit(b) ? (int8_t)-1 : (int8_t)1;
        }
      }
      if (s.logp_out != nullptr) {
        const T lp = logpsi_now();
        if (lane == 0 && wk == 0) reinterpret_cast<T *>(s.logp_out)[o] = pw * lp;
      }
      if (s.tanh_out != nullptr) {
        // tanh(theta_j) = (A_j - B_j) / (A_j + B_j): what the forces need, while it is in registers
        T *to = reinterpret_cast<T *>(s.tanh_out) + o * M;
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          const int ju = LM::unit(e, lane), j = wk * MW + ju;
          T av, bv;
          if constexpr (F64) {
            av = A[e];
            bv = Bv[e];
          } else {
            av = (e & 1) ? A[e >> 1].y : A[e >> 1].x;
            bv = (e & 1) ? Bv[e >> 1].y : Bv[e >> 1].x;
          }
          if (ju < MW && j < M) to[j] = (av - bv) / (av + bv);
        }
      }
      if (s.eloc_kind != 0) {
        const T e = local_energy();
        if (lane == 0 && wk == 0) store_as<T>(s.eloc_out, o, e, s.eloc_dtype);
      }
    };

    for (int tt = 0; tt < T_total; tt += 32) {
      // ---- 32 proposals' worth of randomness, one Philox call per lane
      uint32_t w0_l = 0;
      int thr_l = 0;
      T u_l = T(0);
      if (tt + lane < T_total) {
        if (s.stream_w0 != nullptr) {
          w0_l = s.stream_w0[(size_t)(tt + lane) * s.B + chain];
          u_l = reinterpret_cast<const T *>(s.stream_u)[(size_t)(tt + lane) * s.B + chain];
        } else {
          const uint4 w = philox_words(s.seed, s.t0 + (uint64_t)(tt + lane), gchain, STREAM_STEP);
          w0_l = w.x;
          u_l = uniform_from_words<T>(w);
        }
        // thr = fix(log2(u) / machine_pow);  u == 0 or machine_pow == 0: always accept
        thr_l = -2000000000;
        if (pw > T(0) && u_l > T(0)) {
          if constexpr (F64) {
            const double t2 = log2(u_l) * inv_pw * (double)PROD_FX_SCALE;
            thr_l = t2 > -2.0e9 ? __double2int_rn(t2) : -2000000000;
          } else {
            thr_l = __float2int_rn(fmaxf(log2f(u_l) * inv_pw * PROD_FX_SCALE, -2.0e9f));
          }
        }
        if (RULE == NK_RULE_LOCAL) w0_l = __umulhi(w0_l, (uint32_t)N);  // the site
      }
      const int nb = min(32, T_total - tt);
      int k = 0;
      while (k < nb) {
        const int kend = k + min(nb - k, sweep_size - in_sweep);
        in_sweep += kend - k;
        for (; k < kend; ++k) {
          const int thr = __shfl_sync(FULL, thr_l, k);
          const uint32_t w0 = __shfl_sync(FULL, w0_l, k);
          if (s.eloc_only) continue;  // stand-alone local estimator: one empty "sweep", then the sample's E_loc
          if (since >= renorm) renormalise();
          if constexpr (RULE == NK_RULE_LOCAL) {
            const int site = (int)w0;
            V g[NV];
            fetch(site, g);
            const Rc &rc = rctab[site];
            const int fx = rc.fx, fy = rc.fy;
            const bool pos = sbit(site) != 0;  // sigma = -1 -> +1
            const LanePair<T> P = pos ? lane_product<V, NV>(Bv, A, g) : lane_product<V, NV>(A, Bv, g);
            const int Rp = gcomb(__reduce_add_sync(FULL, lp_fx(P, wide)), op_add);
            const int X = (int)((uint32_t)Rp - (uint32_t)R + (uint32_t)(pos ? fx + fy : fx - fy));
            bool acc;
            if constexpr (F64) {
              if (thr < X - PROD_FX_BAND)
                acc = true;
              else if (thr >= X + PROD_FX_BAND)
                acc = false;
              else
                acc = exact_decide(gcomb(exact_logratio(lp_val(P), lane_norm()), op_add), pos ? rc.xn + rc.yn : rc.xn - rc.yn,
                                   __shfl_sync(FULL, u_l, k), pw, 0.0);
            } else {
              acc = thr < X;
            }
            if (acc) {
              if (pos) {
#pragma unroll
                for (int q = 0; q < NV; ++q) Bv[q] = vmul(Bv[q], g[q]);
              } else {
#pragma unroll
                for (int q = 0; q < NV; ++q) A[q] = vmul(A[q], g[q]);
              }
              R = Rp;
              ++nacc;
              ++since;
              stoggle(site);
            }
          } else {
            // ExchangeRule.transition (rules/exchange.py:143-184): uniform over the hoppable clusters, or
            // (probabilities=, :86-123,155-160) by inverse CDF over their weights
            int csel = -1;
            double w_hop = 0.0;  // weighted rule: total weight of the hoppable clusters
            if (cprob != nullptr) {
              // lane l owns clusters 64 l .. 64 l + 63 (words 2l, 2l+1): its weight, an inclusive scan over the lanes, then the
              // first cluster (in cluster order) whose running weight reaches r = W (w0 + 1/2) / 2^32  (jax.random.choice:
              // searchsorted(cumsum(p), r))
              const uint2 hw = *reinterpret_cast<const uint2 *>(hopw + 2 * lane);
              double wl = 0.0;
              for (uint32_t bts = hw.x; bts != 0u; bts &= bts - 1u) wl += cprob[64 * lane + __ffs(bts) - 1];
              for (uint32_t bts = hw.y; bts != 0u; bts &= bts - 1u) wl += cprob[64 * lane + 32 + __ffs(bts) - 1];
              double incl = wl;
#pragma unroll
              for (int dd = 1; dd < 32; dd <<= 1) {
                const double t = __shfl_up_sync(FULL, incl, dd);
                if (lane >= dd) incl += t;
              }
              w_hop = __shfl_sync(FULL, incl, 31);
              if (w_hop > 0.0) {
                const double r = w_hop * (((double)w0 + 0.5) * 2.3283064365386963e-10);
                const uint32_t bal = __ballot_sync(FULL, (hw.x | hw.y) != 0u && incl >= r);
                // (rounding can leave r above the last running sum: the last hoppable cluster, as searchsorted clipped)
                const uint32_t any = __ballot_sync(FULL, (hw.x | hw.y) != 0u);
                const int src = bal != 0u ? __ffs(bal) - 1 : 31 - __clz(any);
                int cidx = 0;
                if (lane == src) {
                  double run = incl - wl;
                  int last = 0;
                  bool found = false;
                  for (uint32_t bts = hw.x; bts != 0u && !found; bts &= bts - 1u) {
                    last = 64 * lane + __ffs(bts) - 1;
                    run += cprob[last];
                    found = run >= r;
                  }
                  for (uint32_t bts = hw.y; bts != 0u && !found; bts &= bts - 1u) {
                    last = 64 * lane + 32 + __ffs(bts) - 1;
                    run += cprob[last];
                    found = run >= r;
                  }
                  cidx = last;
                }
                csel = __shfl_sync(FULL, cidx, src);
              }
            } else if (n_hop > 0) {
              const int kth = (int)__umulhi(w0, (uint32_t)n_hop);
              // ---- the kth hoppable cluster in cluster order: lane l owns words 2l, 2l+1
              {
                const uint2 hw = *reinterpret_cast<const uint2 *>(hopw + 2 * lane);
                const int c0 = __popc(hw.x), cnt = c0 + __popc(hw.y);
                int incl = cnt;
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) {
                  const int t = __shfl_up_sync(FULL, incl, dd);
                  if (lane >= dd) incl += t;
                }
                const uint32_t bal = __ballot_sync(FULL, kth < incl);
                const int src = __ffs(bal) - 1;
                const int kk = kth - (incl - cnt);
                int cidx = 0;
                if (lane == src) cidx = kk < c0 ? 64 * lane + kth_set_bit(hw.x, kk) : 64 * lane + 32 + kth_set_bit(hw.y, kk - c0);
                csel = __shfl_sync(FULL, cidx, src);
              }
            }
            if (csel >= 0) {
              const int si = cl[2 * csel], sj = cl[2 * csel + 1];
              const bool ineg = sbit(si) != 0;
              const int sp = ineg ? si : sj;  // sigma = -1 -> +1: multiplies B
              const int sm = ineg ? sj : si;  // sigma = +1 -> -1: multiplies A
              // ---- n_hop(sigma'): every other cluster containing si or sj toggles
              uint32_t ei = 0xffffffffu, ej = 0xffffffffu;
              int n1 = 0, n0 = 0;
              double w_hop_p = 0.0;
              {
                const bool vi = lane < (int)adjdeg[si];
                if (vi) ei = adj[si * PROD_ADJ_MAX + lane];
                const bool ui = vi && (int)(ei >> 16) != sj;
                if (!ui) ei = 0xffffffffu;
                const uint32_t bi = ui ? (hopw[(ei & 0xffffu) >> 5] >> (ei & 31u)) & 1u : 0u;
                const bool vj = lane < (int)adjdeg[sj];
                if (vj) ej = adj[sj * PROD_ADJ_MAX + lane];
                const bool uj = vj && (int)(ej >> 16) != si;
                if (!uj) ej = 0xffffffffu;
                const uint32_t bj = uj ? (hopw[(ej & 0xffffu) >> 5] >> (ej & 31u)) & 1u : 0u;
                n1 = __popc(__ballot_sync(FULL, bi != 0)) + __popc(__ballot_sync(FULL, bj != 0));
                n0 = __popc(__ballot_sync(FULL, ui && bi == 0)) + __popc(__ballot_sync(FULL, uj && bj == 0));
                if (cprob != nullptr) {  // W(sigma') = W(sigma) -+ the weights of the clusters that stop / start being hoppable
                  double dw = 0.0;
                  if (ui) dw += bi ? -cprob[ei & 0xffffu] : cprob[ei & 0xffffu];
                  if (uj) dw += bj ? -cprob[ej & 0xffffu] : cprob[ej & 0xffffu];
                  w_hop_p = w_hop + warp_sum(dw);
                }
              }
              const int nhp = n_hop - n1 + n0;
              const Rc &rp = rctab[sp];
              const Rc &rm = rctab[sm];
              // log_prob_corr = log n(sigma) - log n(sigma')  (:177-182), with the total weights in place of the counts
              const double corr_w = cprob != nullptr ? log(w_hop) - log(w_hop_p) : 0.0;
              const int corr_fx = cprob != nullptr ? __double2int_rn(corr_w * 1.4426950408889634 * (double)inv_pw * (double)PROD_FX_SCALE)
                                                   : lgtab[n_hop] - lgtab[nhp];
              const int cfix = rp.fx + rp.fy + rm.fx - rm.fy + corr_fx;
              V g[NV], t[NV];
              fetch(sm, g);
#pragma unroll
              for (int q = 0; q < NV; ++q) t[q] = vmul(A[q], g[q]);
              fetch(sp, g);
              V c[NV];
#pragma unroll
              for (int q = 0; q < NV; ++q) c[q] = vfma(Bv[q], g[q], t[q]);
              const LanePair<T> P = lane_product1<V, NV>(c);
              const int Rp = gcomb(__reduce_add_sync(FULL, lp_fx(P, wide)), op_add);
              const int X = (int)((uint32_t)Rp - (uint32_t)R + (uint32_t)cfix);
              bool acc;
              if constexpr (F64) {
                if (thr < X - PROD_FX_BAND)
                  acc = true;
                else if (thr >= X + PROD_FX_BAND)
                  acc = false;
                else
                  acc = exact_decide(gcomb(exact_logratio(lp_val(P), lane_norm()), op_add), rp.xn + rp.yn + rm.xn - rm.yn, __shfl_sync(FULL, u_l, k), pw,
                                     cprob != nullptr ? corr_w : log((double)n_hop) - log((double)nhp));
              } else {
                acc = thr < X;
              }
              if (acc) {
#pragma unroll
                for (int q = 0; q < NV; ++q) {
                  A[q] = t[q];
                  Bv[q] = vmul(Bv[q], g[q]);
                }
                R = Rp;
                ++nacc;
                ++since;
                stoggle(si);
                stoggle(sj);
                if (ei != 0xffffffffu) atomicXor(hopw + ((ei & 0xffffu) >> 5), 1u << (ei & 31u));
                if (ej != 0xffffffffu) atomicXor(hopw + ((ej & 0xffffu) >> 5), 1u << (ej & 31u));
                n_hop = nhp;
                __syncwarp();
              }
            }
          }
        }
        if (in_sweep == sweep_size) end_of_sweep();
      }
    }
    if (s.eloc_only) continue;
    // ---- write the chain state back
    if (wk == 0) {
      for (int b = 0; b < nbw; ++b) {
        const int idx = 32 * b + lane;
        if (idx < N) s.sigma[(size_t)chain * N + idx] = ownbit(b) ? (int8_t)-1 : (int8_t)1;
      }
    }
    const T lp = logpsi_now();
    if (lane == 0 && wk == 0) {
      reinterpret_cast<T *>(s.log_prob)[chain] = pw * lp;
      s.n_accepted[chain] += (int64_t)nacc;
    }
    __syncwarp();
  }
}

// host-side launcher of one instantiation (called from sweep_prod_inst_*.cu)
template <typename T, int NFULL, int TAIL, int RULE, bool MULTI>
int launch_prod(cudaStream_t stream, const ProdArgs &a) {
  auto kern = sweep_prod_kernel<T, NFULL, TAIL, RULE, MULTI>;
  NK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, a.L.smem_bytes));
  const int groups = a.L.warps / a.L.kw;  // chains in flight per CTA
  const int64_t need = (a.s.B + groups - 1) / groups;
  const int64_t cap = num_sms();
  kern<<<(int)(need < cap ? need : cap), a.L.warps * 32, a.L.smem_bytes, stream>>>(a);
  NK_LAUNCH_OK();
  return NK_OK;
}

int launch_prod_f32_local(cudaStream_t stream, const ProdArgs &a, int nfull, int tail);
int launch_prod_f32_exchange(cudaStream_t stream, const ProdArgs &a, int nfull, int tail);
int launch_prod_f64_local(cudaStream_t stream, const ProdArgs &a, int nfull, int tail);
int launch_prod_f64_exchange(cudaStream_t stream, const ProdArgs &a, int nfull, int tail);
// several warps per chain (M > 512): LocalRule only
int launch_prod_f32_local_multi(cudaStream_t stream, const ProdArgs &a, int nfull, int tail);
int launch_prod_f64_local_multi(cudaStream_t stream, const ProdArgs &a, int nfull, int tail);
int launch_prod_f32_exchange_multi(cudaStream_t stream, const ProdArgs &a, int nfull, int tail);
int launch_prod_f64_exchange_multi(cudaStream_t stream, const ProdArgs &a, int nfull, int tail);

}  // namespace nk
