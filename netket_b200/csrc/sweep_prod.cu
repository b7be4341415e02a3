// Host side and table preparation of the general product-form sweep kernel (sweep_prod.cuh).
//
// prod_prep_rows<T>    one CTA per site i: G_i. = exp(-4 W_i.) padded to the kernel's row length, the per-site
//                      constants (2 sum_j W_ij, 2 a_i and their fixed-point / exponential forms), max |W|, max_i sum_j |W_ij|
// prod_prep_tables<T>  one CTA: ExchangeRule tables (clusters as bytes, clusters-per-site lists, fix(log2 n / machine_pow)),
//                      Ising edges as bytes, LocalOperator tables compacted (netket/operator/_local_operator/
//                      compile_helpers.py:29-218 layout -> sites / diag / mel / code arrays), the renormalisation period and
//                      the decision whether the product form applies at all (flags[0] = 1 hands over to the generic kernel).
#include <string.h>

#include "sweep_prod.cuh"

namespace nk {

using namespace prod;

template <typename T>
__global__ void __launch_bounds__(128) prod_prep_rows(const __grid_constant__ ProdArgs p, int MP) {
  // row layout: kw segments (one per cooperating warp) of MP elements; segment k holds hidden units [k mw, (k+1) mw)
  typedef typename VecOf<T>::Rc Rc;
  __shared__ double red[3][4];
  if (p.run_if != nullptr && *p.run_if == 0) return;
  const int i = blockIdx.x, N = p.s.rbm.N, M = p.s.rbm.M;
  (void)N;
  const T *W = reinterpret_cast<const T *>(p.s.rbm.W) + (size_t)i * M;
  T *row = reinterpret_cast<T *>(const_cast<unsigned char *>(p.gtab) + (size_t)i * p.L.row_bytes);
  double rs = 0.0, ra = 0.0, wm = 0.0;
  const int kw = p.L.kw, mw = p.L.mw;
  for (int pos = threadIdx.x; pos < kw * MP; pos += blockDim.x) {
    const int seg = pos / MP, t = pos - seg * MP, j = seg * mw + t;
    T g = T(1);  // padding: leaves every product unchanged
    if (t < mw && j < M) {
      const T w = W[j];
      g = Math<T>::exp(T(-4) * w);
      rs += (double)w;
      ra += fabs((double)w);
      wm = fmax(wm, fabs((double)w));
      if (!(fabs((double)w) < 1.0e30)) wm = 1.0e300;  // NaN / Inf
    }
    row[pos] = g;
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    rs += __shfl_xor_sync(0xffffffffu, rs, m);
    ra += __shfl_xor_sync(0xffffffffu, ra, m);
    wm = fmax(wm, __shfl_xor_sync(0xffffffffu, wm, m));
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = rs;
    red[1][threadIdx.x >> 5] = ra;
    red[2][threadIdx.x >> 5] = wm;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    rs = red[0][0] + red[0][1] + red[0][2] + red[0][3];
    ra = red[1][0] + red[1][1] + red[1][2] + red[1][3];
    wm = fmax(fmax(red[2][0], red[2][1]), fmax(red[2][2], red[2][3]));
    const double LOG2E = 1.4426950408889634;
    const double xn = 2.0 * rs;
    const double yn = p.s.rbm.a != nullptr ? 2.0 * (double)reinterpret_cast<const T *>(p.s.rbm.a)[i] : 0.0;
    Rc *rc = reinterpret_cast<Rc *>(const_cast<unsigned char *>(p.aux) + p.L.rc_off) + i;
    if constexpr (sizeof(T) == 8) {
      rc->xn = xn;
      rc->yn = yn;
      rc->ep = exp(xn + yn);
      rc->em = exp(xn - yn);
      rc->fx = __double2int_rn(fmax(fmin(LOG2E * xn * (double)PROD_FX_SCALE, 1.0e9), -1.0e9));
      rc->fy = __double2int_rn(fmax(fmin(LOG2E * yn * (double)PROD_FX_SCALE, 1.0e9), -1.0e9));
      rc->pad0 = rc->pad1 = 0;
    } else {
      rc->x2 = (float)(LOG2E * xn);
      rc->y2 = (float)(LOG2E * yn);
      rc->fx = __double2int_rn(fmax(fmin(LOG2E * xn * (double)PROD_FX_SCALE, 1.0e9), -1.0e9));
      rc->fy = __double2int_rn(fmax(fmin(LOG2E * yn * (double)PROD_FX_SCALE, 1.0e9), -1.0e9));
    }
    // non-negative floats order like their bit patterns
    atomicMax(p.flags + 2, __float_as_int((float)fmin(wm, 3.0e38)));
    atomicMax(p.flags + 3, __float_as_int((float)fmin(fmax(ra, fabs(yn)), 3.0e38)));
  }
}

template <typename T>
__global__ void __launch_bounds__(256) prod_prep_tables(const __grid_constant__ ProdArgs p, int NE_pad) {
  __shared__ int deg[1024];
  __shared__ int bad;
  const SweepKernelArgs &s = p.s;
  const ProdLayout &L = p.L;
  unsigned char *aux = const_cast<unsigned char *>(p.aux);
  const int N = s.rbm.N, tid = threadIdx.x, nt = blockDim.x;
  if (p.run_if != nullptr && *p.run_if == 0) return;
  if (tid == 0) bad = 0;
  for (int i = tid; i < 1024; i += nt) deg[i] = 0;
  __syncthreads();
  // ---- ExchangeRule tables
  if (s.rule == NK_RULE_EXCHANGE) {
    const int C = s.n_clusters;
    uint16_t *cl = reinterpret_cast<uint16_t *>(aux + L.cl_off);
    uint32_t *adj = reinterpret_cast<uint32_t *>(aux + L.adj_off);
    int *lg = reinterpret_cast<int *>(aux + L.lg_off);
    if (s.cluster_probs != nullptr) {
      double *cp = reinterpret_cast<double *>(aux + L.cp_off);
      for (int c = tid; c < C; c += nt) cp[c] = s.cluster_probs[c];
    }
    for (int c = tid; c < C; c += nt) {
      const int i = s.clusters[2 * c], j = s.clusters[2 * c + 1];
      if (i < 0 || j < 0 || i >= N || j >= N || i == j) {
        bad = 1;
        continue;
      }
      cl[2 * c] = (uint16_t)i;
      cl[2 * c + 1] = (uint16_t)j;
      const int a = atomicAdd(&deg[i], 1), b = atomicAdd(&deg[j], 1);
      if (a < PROD_ADJ_MAX) adj[i * PROD_ADJ_MAX + a] = (uint32_t)c | ((uint32_t)j << 16);
      if (b < PROD_ADJ_MAX) adj[j * PROD_ADJ_MAX + b] = (uint32_t)c | ((uint32_t)i << 16);
      if (a >= PROD_ADJ_MAX || b >= PROD_ADJ_MAX) bad = 1;
    }
    const double pw = s.machine_pow;
    for (int n = tid; n <= C + 1; n += nt)
      lg[n] = (n >= 1 && pw > 0.0) ? __double2int_rn(log2((double)n) / pw * (double)PROD_FX_SCALE) : 0;
    __syncthreads();
    uint8_t *adjdeg = aux + L.adjdeg_off;
    for (int i = tid; i < N; i += nt) adjdeg[i] = (uint8_t)min(deg[i], PROD_ADJ_MAX);
  }
  // ---- Ising edges
  if (s.eloc_kind == 1) {
    uint16_t *edges = reinterpret_cast<uint16_t *>(aux + L.edges_off);
    for (int e = tid; e < 2 * s.ising.n_edges; e += nt) {
      const int v = s.ising.edges[e];
      if (v < 0 || v >= N) bad = 1;
      edges[e] = (uint16_t)v;
    }
  }
  // ---- LocalOperator: compact tables
  if (s.eloc_kind == 2) {
    for (int gi = 0; gi < s.localop.n_groups; ++gi) {
      const nk_localop_group_t &G = s.localop.groups[gi];
      const int rows = 1 << G.n_sites, ncm = G.ncmax;
      uint16_t *sites = reinterpret_cast<uint16_t *>(aux + L.lop_sites_off[gi]);
      T *dg = reinterpret_cast<T *>(aux + L.lop_diag_off[gi]);
      T *ml = reinterpret_cast<T *>(aux + L.lop_mel_off[gi]);
      uint8_t *cd = aux + L.lop_code_off[gi];
      for (int o = tid; o < G.n_ops; o += nt) {
        const int s0 = G.acting_on[o * G.n_sites], s1 = G.n_sites == 2 ? G.acting_on[o * G.n_sites + 1] : s0;
        if (s0 < 0 || s1 < 0 || s0 >= N || s1 >= N || (G.n_sites == 2 && s0 == s1)) bad = 1;
        sites[2 * o] = (uint16_t)s0;
        sites[2 * o + 1] = (uint16_t)s1;
      }
      for (int e = tid; e < G.n_ops * rows; e += nt) dg[e] = (T)G.diag_mels[e];
      for (int e = tid; e < G.n_ops * rows * ncm; e += nt) {
        const int c = e % ncm, orow = e / ncm;
        const double mel = c < G.n_conns[orow] ? G.mels[e] : 0.0;
        const bool valid = c < G.n_conns[orow] && fabs(mel) > s.localop.mel_cutoff;
        int code = valid ? 1 : 0;
        if (valid) {
          const int8_t *xp = G.x_prime + (size_t)e * G.n_sites;
          code |= (xp[0] != 0 ? 2 : 0);
          if (G.n_sites == 2) code |= (xp[1] != 0 ? 4 : 0);
        }
        ml[e] = valid ? (T)mel : T(0);
        cd[e] = (uint8_t)code;
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    const float wmax = __int_as_float(p.flags[2]), rowabs = __int_as_float(p.flags[3]);
    const float LOG2E = 1.4426950408889634f;
    // Renormalisation period r (A + B = 1 is restored once r accepted moves have piled up, before the next lane product):
    // a lane product then sees at most r - 1 un-normalised accepts, i.e. factors within G^(+-r) of 1.  Two ranges bind:
    //   (i)  fp32: a partial lane product (NE_pad / nsplit factors) must stay a normal float:
    //              (NE_pad / nsplit) r 4 max|W| log2(e) <= PROD_EXP_RANGE;   fp64: NE_pad factors inside the double range;
    //   (ii) the warp's fixed-point sum must stay inside int32 (2^31 / 2^19 = 4096): r 4 max_i sum_j |W_ij| log2(e) <= 1800.
    // A LocalOperator may hold same-sign double flips (G_0 G_1 per factor): (i) is then applied with twice the exponent.
    const float perW = 4.0f * wmax * LOG2E * (s.eloc_kind == 2 ? 2.0f : 1.0f), perRow = 4.0f * rowabs * LOG2E;
    int renorm = 0, nsplit = 1;
    if (wmax < 1.0e30f) {
      auto period = [&](int ns) {
        const float lim_i = sizeof(T) == 8 ? 900.0f : PROD_EXP_RANGE;
        float r = 32.0f;
        if ((float)(NE_pad / ns) * perW > 0.0f) r = fminf(r, lim_i / ((float)(NE_pad / ns) * perW));
        if (perRow > 0.0f) r = fminf(r, 1800.0f / perRow);
        return (int)r;
      };
      renorm = period(1);
      if (sizeof(T) == 4 && renorm < 4 && NE_pad >= 4) {  // wide mode: two logarithms per lane product
        const int r2 = period(2);
        if (r2 > renorm) {
          renorm = r2;
          nsplit = 2;
        }
      }
    }
    // fp64 local energy multiplies M factors of up to two rows across the warp: keep 2 x 4 sum_j |W_ij| (+ the visible term)
    // inside the double range with a wide margin; the per-site exponentials exp(xn +- yn) must be finite as well
    // (beyond that the fp64 local energy reduces (mantissa, exponent) pairs and works in the log2 domain: flags[7])
    p.flags[7] = (sizeof(T) == 8 && !(8.0f * rowabs * LOG2E < 900.0f)) ? 1 : 0;
    if (sizeof(T) == 4 && !(4.0f * rowabs * LOG2E < 2000.0f)) renorm = 0;  // fixed-point constants stay inside int32
    p.flags[6] = nsplit;
    p.flags[1] = renorm;
    if (renorm < 1 || bad) p.flags[p.giveup] = 1;
  }
}

// ------------------------------------------------------------------------------------------ host side
struct ProdShape {
  int nfull, tail, ne_pad, seg_bytes, mp;  // per-warp segment
  int kw, mw;                              // warps per chain, hidden units per warp
};

// shape of one warp's segment holding `m` hidden units
static bool seg_shape(int m, int dtype, ProdShape *ps, int max_full_f32 = 4) {
  const int esz = dtype == NK_F32 ? 4 : 8;
  const int chunk = 512 / esz;  // elements per 512-byte chunk
  int nfull = m / chunk, rem = m % chunk, tail = 0;
  if (rem == 0)
    tail = 0;
  else if (rem <= 32)
    tail = 1;
  else if (rem <= 64 && esz == 4)
    tail = 2;
  else {
    nfull += 1;
    tail = 0;
  }
  const int max_full = esz == 4 ? max_full_f32 : 8;
  if (m < 1 || nfull > max_full || (nfull == max_full && tail != 0)) return false;
  ps->nfull = nfull;
  ps->tail = tail;
  const int ne = (16 / esz) * nfull + tail;
  ps->ne_pad = esz == 4 ? 2 * ((ne + 1) / 2) : ne;
  ps->mp = 32 * (16 / esz) * nfull + 32 * tail;
  ps->seg_bytes = ps->mp * esz;
  return true;
}

static int prod_warps(int dtype, int rule) {
  return dtype == NK_F32 ? (rule == NK_RULE_LOCAL ? ProdWarps<float, NK_RULE_LOCAL>::value : ProdWarps<float, NK_RULE_EXCHANGE>::value)
                         : (rule == NK_RULE_LOCAL ? ProdWarps<double, NK_RULE_LOCAL>::value : ProdWarps<double, NK_RULE_EXCHANGE>::value);
}

// M <= 512: one warp per chain.  Larger hidden layers: kw warps per chain, chosen to waste the least padding
// among the instantiated segment shapes (at least half-full segments), then the fewest warps.
#ifndef NK_PROD_MULTI_MAX_FULL_F32
#define NK_PROD_MULTI_MAX_FULL_F32 5
#endif
static bool prod_shape(int M, int dtype, int rule, ProdShape *ps) {
  if (seg_shape(M, dtype, ps)) {
    ps->kw = 1;
    ps->mw = M;
    return true;
  }
  const int warps = prod_warps(dtype, rule);
  const int min_full = dtype == NK_F32 ? 2 : 4;
  long best = -1;
  for (int kw = 2; kw <= 16 && kw <= warps; ++kw) {
    ProdShape c;
    const int mw = (M + kw - 1) / kw;
    // fp32 segments of up to 5 x 128 units (20 per lane) exist for the multi-warp kernel only: fewer, fatter warps per chain
    // mean more chains in flight per SM (M = 3200: 5 warps x 640 units, 4 chains per SM instead of 2)
    if (!seg_shape(mw, dtype, &c, NK_PROD_MULTI_MAX_FULL_F32) || c.nfull < min_full) continue;
    if (dtype == NK_F32 && c.nfull >= 4 && c.tail != 0) continue;  // (4, 0) and (5, 0) are the instantiated fat segments
    const long cost = (long)kw * c.mp * 64 + kw;
    if (best < 0 || cost < best) {
      best = cost;
      *ps = c;
      ps->kw = kw;
      ps->mw = mw;
    }
  }
  return best >= 0;
}

static inline int align16(int x) { return (x + 15) & ~15; }

static bool prod_layout(const SweepKernelArgs &a, const ProdShape &ps, ProdLayout *L) {
  const int esz = a.rbm.dtype == NK_F32 ? 4 : 8;
  const int N = a.rbm.N;
  memset(L, 0, sizeof(*L));
  L->seg_bytes = ps.seg_bytes;
  L->row_bytes = ps.kw * ps.seg_bytes;
  L->kw = ps.kw;
  L->mw = ps.mw;
  const int max_warps = prod_warps(a.rbm.dtype, a.rule);
  L->warps = (max_warps / ps.kw) * ps.kw;
  int off = 0;
  L->rc_off = off;
  L->rc_stride = esz == 4 ? (int)sizeof(RcF) : (int)sizeof(RcD);
  off = align16(off + N * L->rc_stride);
  if (a.rule == NK_RULE_EXCHANGE) {
    const int C = a.n_clusters;
    L->lg_off = off;
    off = align16(off + (C + 2) * 4);
    L->cl_off = off;
    off = align16(off + 4 * C);
    L->adjdeg_off = off;
    off = align16(off + N);
    L->adj_off = off;
    off = align16(off + N * PROD_ADJ_MAX * 4);
    L->cp_off = off;
    if (a.cluster_probs != nullptr) off = align16(off + 8 * C);
  }
  if (a.eloc_kind == 1) {
    L->edges_off = off;
    off = align16(off + 4 * a.ising.n_edges);
  }
  if (a.eloc_kind == 2) {
    for (int gi = 0; gi < a.localop.n_groups; ++gi) {
      const nk_localop_group_t &G = a.localop.groups[gi];
      const int64_t rows = 1 << G.n_sites;
      const int64_t total = (int64_t)G.n_ops * rows * (G.ncmax + 1) * (esz + 1);
      if (total > PROD_AUX_MAX) return false;
      L->lop_sites_off[gi] = off;
      off = align16(off + 4 * G.n_ops);
      L->lop_diag_off[gi] = off;
      off = align16(off + (int)(G.n_ops * rows) * esz);
      L->lop_mel_off[gi] = off;
      off = align16(off + (int)(G.n_ops * rows * G.ncmax) * esz);
      L->lop_code_off[gi] = off;
      off = align16(off + (int)(G.n_ops * rows * G.ncmax));
    }
  }
  if (off > PROD_AUX_MAX) return false;
  L->aux_bytes = off;
  const int hop_bytes = a.rule == NK_RULE_EXCHANGE ? max_warps * PROD_HOP_WORDS * 4 : 0;
  // MULTI kernels (generic-address rows, cross-warp slots): several warps per chain, or an fp32 table that does not fit
  // shared memory (the one-warp fp32 kernels read the table with explicit shared-memory loads only)
  int multi = ps.kw > 1 ? 1 : 0;
  int n_res = 0, xs_bytes = 0;
  for (int pass = 0; pass < 2; ++pass) {
    if (multi) {  // one named barrier (ids 1..15) per chain slot of the CTA
      int groups = max_warps / ps.kw;
      if (groups > 15) groups = 15;
      L->warps = groups * ps.kw;
    }
    xs_bytes = multi ? (L->warps / ps.kw) * 2 * ps.kw * 256 : 0;
    const int budget = 227 * 1024 - (L->aux_bytes + hop_bytes + xs_bytes + 16);  // all offsets are 16-byte multiples
    if (budget < 0) return false;
    n_res = budget / L->row_bytes;
    if (n_res > N) n_res = N;
    if (esz == 4 && !multi && n_res < N) {
      if (ps.nfull < 2) return false;
      multi = 1;
      continue;
    }
    break;
  }
  L->multi = multi;
  L->n_res = n_res;
  L->g_bytes = n_res * L->row_bytes;
  L->hop_off = L->g_bytes + L->aux_bytes;
  L->xs_off = L->hop_off + hop_bytes;
  L->bar_off = L->xs_off + xs_bytes;
  L->smem_bytes = L->bar_off + 16;
  return true;
}

bool sweep_prod_supported(const SweepKernelArgs &a) {
  ProdShape ps;
  ProdLayout L;
  if (a.rbm.N > 1024 || !prod_shape(a.rbm.M, a.rbm.dtype, a.rule, &ps)) return false;
  if (a.rule == NK_RULE_EXCHANGE && (a.n_clusters < 1 || a.n_clusters > 32 * PROD_HOP_WORDS)) return false;
  if (a.eloc_kind == 1 && a.ising.n_edges > 8192) return false;
  if (a.B >= (1ll << 31) || (int64_t)(a.n_discard + a.chain_length) * a.sweep_size >= (1ll << 31)) return false;  // 32-bit counters
  return prod_layout(a, ps, &L);
}

// bytes of workspace behind theta and the flags: the G table and the aux blob
size_t sweep_shadow_park_bytes();  // sweep_shadow.cu

static size_t prod_tables_bytes(const nk_rbm_t &rbm) {
  ProdShape ps;
  if (rbm.N > 1024 || !prod_shape(rbm.M, rbm.dtype, NK_RULE_LOCAL, &ps)) return 0;
  return (((size_t)rbm.N * ps.kw * ps.seg_bytes + 255) & ~(size_t)255) + PROD_AUX_MAX;
}
size_t sweep_prod_workspace_bytes(const nk_rbm_t &rbm) {
  const size_t t = prod_tables_bytes(rbm);
  if (t == 0) return 0;
  // fp64: room to park the double state of the chains in flight while they sweep on the fp32 shadow (sweep_shadow.cu)
  return t + (rbm.dtype == NK_F64 ? sweep_shadow_park_bytes() : 0);
}

static size_t prod_tables_bytes(const nk_rbm_t &rbm);
// sweep_shadow.cu
bool sweep_shadow_supported(const SweepKernelArgs &a, const ProdLayout &L);
int sweep_shadow(cudaStream_t stream, const ProdArgs &pa, int give, double *park);
size_t sweep_shadow_park_bytes();

int sweep_prod(cudaStream_t stream, const SweepKernelArgs &a, const void *theta_ws, int *flags, void *tables_ws, const int *run_if,
               int giveup, const int **stats_guard) {
  if (stats_guard) *stats_guard = nullptr;
  ProdShape ps;
  ProdArgs pa{};
  if (!prod_shape(a.rbm.M, a.rbm.dtype, a.rule, &ps) || !prod_layout(a, ps, &pa.L)) {
    set_error("sweep_prod: unsupported configuration");
    return NK_EUNSUPPORTED;
  }
  pa.s = a;
  pa.gtab = reinterpret_cast<const unsigned char *>(tables_ws);
  pa.aux = pa.gtab + (((size_t)a.rbm.N * pa.L.row_bytes + 255) & ~(size_t)255);
  pa.theta = theta_ws;
  pa.flags = flags;
  pa.run_if = run_if;
  pa.giveup = giveup;
  const bool multi = pa.L.multi != 0;
  if (a.rbm.dtype == NK_F32) {
    prod_prep_rows<float><<<a.rbm.N, 128, 0, stream>>>(pa, ps.mp);
    NK_LAUNCH_OK();
    prod_prep_tables<float><<<1, 256, 0, stream>>>(pa, ps.ne_pad);
    NK_LAUNCH_OK();
    if (multi)
      return a.rule == NK_RULE_LOCAL ? launch_prod_f32_local_multi(stream, pa, ps.nfull, ps.tail)
                                     : launch_prod_f32_exchange_multi(stream, pa, ps.nfull, ps.tail);
    return a.rule == NK_RULE_LOCAL ? launch_prod_f32_local(stream, pa, ps.nfull, ps.tail)
                                   : launch_prod_f32_exchange(stream, pa, ps.nfull, ps.tail);
  }
  prod_prep_rows<double><<<a.rbm.N, 128, 0, stream>>>(pa, ps.mp);
  NK_LAUNCH_OK();
  prod_prep_tables<double><<<1, 256, 0, stream>>>(pa, ps.ne_pad);
  NK_LAUNCH_OK();
  if (multi)
    return a.rule == NK_RULE_LOCAL ? launch_prod_f64_local_multi(stream, pa, ps.nfull, ps.tail)
                                   : launch_prod_f64_exchange_multi(stream, pa, ps.nfull, ps.tail);
  if (run_if == nullptr && sweep_shadow_supported(a, pa.L)) {
    // fp32 shadow decisions + streamed double table (sweep_shadow.cu); if the weights are outside the shadow's range it raises
    // flags[8] and the one-table kernel, queued behind with that guard, does the work
    double *park = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(tables_ws) + prod_tables_bytes(a.rbm));
    int rc = sweep_shadow(stream, pa, 8, park);
    if (rc) return rc;
    pa.run_if = flags + 8;
    if (stats_guard) *stats_guard = flags + 8;  // the shadow kernel reduces its energies itself
  }
  return a.rule == NK_RULE_LOCAL ? launch_prod_f64_local(stream, pa, ps.nfull, ps.tail)
                                 : launch_prod_f64_exchange(stream, pa, ps.nfull, ps.tail);
}

}  // namespace nk
