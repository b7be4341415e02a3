// extern "C" entry points of libnkb200 (see include/nkb200.h): argument validation, path selection, launches.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "kernels.cuh"

namespace nk {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int num_sms() {
  static std::mutex mu;
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  std::lock_guard<std::mutex> lk(mu);
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev] = n;
  }
  return cache[dev];
}

// implemented in the other translation units
int rbm_logpsi(cudaStream_t stream, const nk_rbm_t &rbm, const int8_t *sigma, int64_t B, void *out, void *theta_out);
int eloc_generic(cudaStream_t stream, const nk_rbm_t &rbm, const nk_ising_t *ising, const nk_localop_t *localop,
                 const int8_t *sigma, int64_t B, void *eloc_out, int32_t eloc_dtype, const int *run_if_flag);
int ising_conn(cudaStream_t stream, const nk_ising_t &op, const int8_t *x, int64_t B, int32_t N, int8_t *xp, void *mels,
               int32_t mel_dtype);
int ising_n_conn(cudaStream_t stream, const nk_ising_t &op, const int8_t *x, int64_t B, int32_t N, int32_t *out);
int localop_conn(cudaStream_t stream, const nk_localop_t &op, const int8_t *x, int64_t B, int32_t N, int8_t *xp, void *mels,
                 int32_t mel_dtype, int32_t *nconn);
int stats_partial(cudaStream_t stream, const void *data, int32_t dtype, int64_t n_chains, int64_t L, int32_t phase, double shift,
                  double *out, const int *run_if = nullptr);
int stats_finalize(const double *p, double mean, int64_t n_chains, int64_t L, double *out);
int stats_tau(cudaStream_t stream, const void *data, int32_t dtype, int64_t n_chains, int64_t L, double c, double *out);
int rbm_tanh_inplace(cudaStream_t stream, void *x, int32_t dtype, int64_t n);
int rbm_jvp_dot(cudaStream_t stream, const nk_rbm_t &v, const int8_t *sigma, int64_t Ns, const void *t, const void *g, double *y,
                double *y_sum);
int online_stats_update(cudaStream_t stream, const nk_online_stats_t *in, const nk_online_stats_t *out, const void *data, int32_t dtype,
                        int64_t n, double decay);
int online_stats_summary(cudaStream_t stream, const nk_online_stats_t *s, int32_t phase, double gmean, double mbar, double *out);
int online_stats_finalize(const double *p0, const double *p1, int64_t n_chains, int64_t n_samples, int32_t L, double *o, double *acf);
int random_state(cudaStream_t stream, int8_t *sigma, int64_t B, int32_t N, int32_t n_down, uint64_t seed, uint64_t chain_offset);
int64_t theta_gemm_workspace_bytes(const nk_rbm_t &rbm, int64_t B);
int theta_gemm(cudaStream_t stream, const nk_rbm_t &rbm, const int8_t *sigma, int64_t B, void *theta_out, void *workspace);

int forces_sums(cudaStream_t stream, const nk_rbm_t &rbm, const int8_t *sigma, const void *theta, int is_tanh, const void *eloc,
                int32_t eloc_dtype, int64_t Ns, double mean, double *sums);
int forces_finalize(cudaStream_t stream, const double *sums, double scale, int64_t n, void *out, int32_t dtype);
bool forces_tc_supported(const nk_rbm_t &rbm);
int forces_tc_sums(cudaStream_t stream, const nk_rbm_t &rbm, const int8_t *sigma, const void *theta, int is_tanh, const void *eloc,
                   int32_t eloc_dtype, int64_t Ns, double mean, double *sums);

static int check_rbm(const nk_rbm_t *rbm, const char *who) {
  NK_CHECK_ARG(rbm != nullptr, "%s: rbm is NULL", who);
  NK_CHECK_ARG(rbm->W != nullptr, "%s: rbm.W is NULL", who);
  NK_CHECK_ARG(rbm->N > 0 && rbm->M > 0, "%s: bad RBM shape N=%d M=%d", who, rbm->N, rbm->M);
  NK_CHECK_ARG(rbm->dtype == NK_F32 || rbm->dtype == NK_F64, "%s: bad dtype %d", who, rbm->dtype);
  return NK_OK;
}

static int check_localop(const nk_localop_t *op, int N, const char *who) {
  NK_CHECK_ARG(op != nullptr, "%s: operator is NULL", who);
  NK_CHECK_ARG(op->n_groups >= 0 && op->n_groups <= 2, "%s: n_groups=%d (only 1- and 2-site terms are supported)", who,
               op->n_groups);
  for (int g = 0; g < op->n_groups; ++g) {
    const nk_localop_group_t &G = op->groups[g];
    NK_CHECK_ARG(G.n_sites == 1 || G.n_sites == 2, "%s: group %d acts on %d sites (supported: 1, 2)", who, g, G.n_sites);
    NK_CHECK_ARG(G.n_ops >= 0 && G.ncmax >= 0, "%s: group %d has bad sizes", who, g);
    if (G.n_ops > 0)
      NK_CHECK_ARG(G.acting_on && G.diag_mels && G.n_conns && (G.ncmax == 0 || (G.mels && G.x_prime)),
                   "%s: group %d has NULL tables", who, g);
  }
  (void)N;
  return NK_OK;
}

}  // namespace nk

using namespace nk;

extern "C" {

const char *nk_last_error(void) { return g_err; }
int nk_version(void) { return NK_VERSION; }
long long nk_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int nk_rbm_logpsi(void *stream, const nk_rbm_t *rbm, const int8_t *sigma, int64_t B, void *logpsi_out, void *theta_out) {
  int rc = check_rbm(rbm, "nk_rbm_logpsi");
  if (rc) return rc;
  NK_CHECK_ARG(B >= 0, "nk_rbm_logpsi: B=%lld", (long long)B);
  NK_CHECK_ARG(B == 0 || (sigma && logpsi_out), "nk_rbm_logpsi: NULL buffer");
  return rbm_logpsi((cudaStream_t)stream, *rbm, sigma, B, logpsi_out, theta_out);
}

int64_t nk_theta_gemm_workspace_bytes(const nk_rbm_t *rbm, int64_t B) {
  if (check_rbm(rbm, "nk_theta_gemm_workspace_bytes") || B < 0) return -1;
  return theta_gemm_workspace_bytes(*rbm, B);
}

int nk_theta_gemm(void *stream, const nk_rbm_t *rbm, const int8_t *sigma, int64_t B, void *theta_out, void *workspace) {
  int rc = check_rbm(rbm, "nk_theta_gemm");
  if (rc) return rc;
  NK_CHECK_ARG(B >= 0, "nk_theta_gemm: B=%lld", (long long)B);
  NK_CHECK_ARG(B == 0 || (sigma && theta_out), "nk_theta_gemm: NULL buffer");
  return theta_gemm((cudaStream_t)stream, *rbm, sigma, B, theta_out, workspace);
}

int nk_random_state(void *stream, int8_t *sigma, int64_t B, int32_t N, int32_t n_down, uint64_t seed, uint64_t chain_offset) {
  NK_CHECK_ARG(B >= 0 && N > 0, "nk_random_state: bad shape B=%lld N=%d", (long long)B, N);
  NK_CHECK_ARG(n_down <= N, "nk_random_state: n_down=%d > N=%d", n_down, N);
  NK_CHECK_ARG(B == 0 || sigma, "nk_random_state: NULL buffer");
  return random_state((cudaStream_t)stream, sigma, B, N, n_down, seed, chain_offset);
}

// workspace layout of the product-form paths:
//   [theta: B*M elements | pad to 256] [flags: 256 B] [G table + auxiliary tables (sweep_prod)] [theta-GEMM workspace]
static inline size_t ws_theta_bytes(const nk_rbm_t *rbm, int64_t B) {
  return (((size_t)B * rbm->M * (rbm->dtype == NK_F32 ? 4 : 8)) + 255) & ~(size_t)255;
}

int64_t nk_sweep_workspace_bytes(const nk_rbm_t *rbm, int64_t B) {
  if (check_rbm(rbm, "nk_sweep_workspace_bytes") || B < 0) return -1;
  const size_t tables = sweep_prod_workspace_bytes(*rbm);
  if (tables == 0 && rbm->dtype != NK_F32) return 0;  // generic path only: theta lives in shared memory for the whole call
  return (int64_t)(ws_theta_bytes(rbm, B) + 256 + tables + (size_t)theta_gemm_workspace_bytes(*rbm, B));
}

int nk_sweep(void *stream, const nk_rbm_t *rbm, nk_chains_t *ch, const nk_sweep_t *a) {
  int rc = check_rbm(rbm, "nk_sweep");
  if (rc) return rc;
  NK_CHECK_ARG(ch && a, "nk_sweep: NULL chains/args");
  NK_CHECK_ARG(ch->B >= 0, "nk_sweep: B=%lld", (long long)ch->B);
  NK_CHECK_ARG(ch->B == 0 || (ch->sigma && ch->log_prob && ch->n_accepted), "nk_sweep: NULL chain buffers");
  NK_CHECK_ARG(a->rule == NK_RULE_LOCAL || a->rule == NK_RULE_EXCHANGE, "nk_sweep: unknown rule %d (only LocalRule and "
               "ExchangeRule are implemented)", a->rule);
  NK_CHECK_ARG(a->chain_length >= 0 && a->n_discard >= 0 && a->sweep_size >= 1, "nk_sweep: bad chain_length/n_discard/sweep_size");
  NK_CHECK_ARG(a->machine_pow >= 0.0, "nk_sweep: machine_pow must be a non-negative real (sampler/base.py:139-150)");
  NK_CHECK_ARG((a->stream_w0 == nullptr) == (a->stream_u == nullptr), "nk_sweep: stream_w0 and stream_u go together");
  if (a->rule == NK_RULE_EXCHANGE)
    NK_CHECK_ARG(a->clusters != nullptr && a->n_clusters > 0, "nk_sweep: ExchangeRule needs clusters");
  NK_CHECK_ARG(!(a->ising && a->localop), "nk_sweep: pass at most one operator");
  if (a->ising || a->localop) {
    NK_CHECK_ARG(a->eloc_out != nullptr, "nk_sweep: eloc_out is NULL");
    NK_CHECK_ARG(a->eloc_dtype == NK_F32 || a->eloc_dtype == NK_F64, "nk_sweep: bad eloc_dtype");
  }
  if (a->ising) NK_CHECK_ARG(a->ising->n_edges == 0 || a->ising->edges, "nk_sweep: ising.edges is NULL");
  if (a->localop) {
    rc = check_localop(a->localop, rbm->N, "nk_sweep");
    if (rc) return rc;
  }
  NK_CHECK_ARG(a->path >= NK_PATH_AUTO && a->path <= NK_PATH_PROD, "nk_sweep: bad path %d", a->path);
  NK_CHECK_ARG(a->stats_out == nullptr || a->ising || a->localop, "nk_sweep: stats_out needs a fused operator (ising / localop)");
  NK_CHECK_ARG(!(a->flags & NK_SWEEP_NO_HANDOVER) || a->stats_out != nullptr, "nk_sweep: NK_SWEEP_NO_HANDOVER needs stats_out");
  if (a->stats_out) NK_CUDA_OK(cudaMemsetAsync(a->stats_out, 0, sizeof(double) * NK_STATS_NPARTIAL, (cudaStream_t)stream));
  if (ch->B == 0) return NK_OK;

  SweepKernelArgs k{};
  k.rbm = *rbm;
  k.sigma = ch->sigma;
  k.log_prob = ch->log_prob;
  k.n_accepted = ch->n_accepted;
  k.B = ch->B;
  k.seed = ch->seed;
  k.t0 = ch->t;
  k.chain_offset = ch->chain_offset;
  k.rule = a->rule;
  k.chain_length = a->chain_length;
  k.n_discard = a->n_discard;
  k.sweep_size = a->sweep_size;
  k.machine_pow = a->machine_pow;
  k.samples_out = a->samples_out;
  k.logp_out = a->logp_out;
  k.stream_w0 = a->stream_w0;
  k.stream_u = a->stream_u;
  k.clusters = a->clusters;
  k.n_clusters = a->n_clusters;
  k.eloc_kind = a->ising ? 1 : (a->localop ? 2 : 0);
  if (a->ising) k.ising = *a->ising;
  if (a->localop) k.localop = *a->localop;
  k.eloc_out = a->eloc_out;
  k.eloc_dtype = a->eloc_dtype;
  k.tanh_out = a->tanh_out;
  k.stats_out = a->stats_out;
  k.stats_shift = a->stats_shift;
  k.cluster_probs = a->rule == NK_RULE_EXCHANGE ? a->cluster_probs : nullptr;
  k.no_handover = (a->flags & NK_SWEEP_NO_HANDOVER) ? 1 : 0;
  const int *stats_guard = nullptr;  // the tuned fp32 kernel reduces its energies itself; the other kernels leave it to K6

  // path selection: the tuned fp32 LocalRule kernel (sweep_fast) where it applies, the general product-form kernel
  // (sweep_prod: fp32/fp64, both rules, Ising / LocalOperator) otherwise, the theta-form generic kernel as the last resort
  // and as the in-stream hand-over target when the weights leave the product form's range.
  const bool have_ws = ch->workspace != nullptr;
  const bool fast_ok = a->path != NK_PATH_PROD && sweep_fast_supported(k) && have_ws;
  const bool prod_ok = !fast_ok && sweep_prod_supported(k) && have_ws;
  if ((a->path == NK_PATH_FAST || a->path == NK_PATH_PROD) && !fast_ok && !prod_ok) {
    set_error("nk_sweep: no product-form kernel for this configuration (needs N<=1024, M<=512 per warp with at most 16 warps "
              "per chain (LocalRule) or M<=512 (ExchangeRule), at most 2048 exchange clusters with at most 32 per site, 1- and "
              "2-site operator terms, and a workspace of nk_sweep_workspace_bytes())");
    return NK_EUNSUPPORTED;
  }
  if (a->path != NK_PATH_GENERIC && (fast_ok || prod_ok)) {
    cudaStream_t st = (cudaStream_t)stream;
    char *wsb = reinterpret_cast<char *>(ch->workspace);
    void *theta = wsb;
    int *flags = reinterpret_cast<int *>(wsb + ws_theta_bytes(rbm, ch->B));
    void *tables = reinterpret_cast<char *>(flags) + 256;
    void *scratch = reinterpret_cast<char *>(tables) + sweep_prod_workspace_bytes(*rbm);
    NK_CUDA_OK(cudaMemsetAsync(flags, 0, 256, st));
    rc = theta_gemm(st, *rbm, ch->sigma, ch->B, theta, scratch);  // theta = sigma W + b  (the only dense contraction)
    if (rc) return rc;
    const int *guard = flags;  // the flag the theta-form kernel waits on
    if (fast_ok) {
      // tuned fp32 kernel first; if the weights are beyond its range (flags[0]) the general kernel's wide mode takes over
      // in-stream, and only if that gives up as well (flags[5]) the theta-form kernel runs
      rc = sweep_fast(st, k, reinterpret_cast<const float *>(theta), flags);
      stats_guard = flags;  // flags[0] != 0: the tuned kernel handed over without producing anything
      if (a->flags & NK_SWEEP_NO_HANDOVER) {
        // the caller takes the hand-over on itself (stats_out[0] = NaN tells it): no guarded launches behind the tuned kernel
        if (rc == NK_OK) ch->t += (uint64_t)(a->n_discard + a->chain_length) * (uint64_t)a->sweep_size;
        return rc;
      }
      if (rc == NK_OK && a->path != NK_PATH_FAST && sweep_prod_supported(k)) {
        rc = sweep_prod(st, k, theta, flags, tables, flags, 5);
        guard = flags + 5;
      }
    } else {
      rc = sweep_prod(st, k, theta, flags, tables, nullptr, 0, &stats_guard);
    }
    if (rc == NK_OK) {
      // weights beyond the product form's range (a property of the data, found by the prep kernels): the product kernel
      // raises flags[0] and exits; this one then runs.  Also enqueued when the path was forced, so that a forced path can
      // never return without having produced its outputs.
      k.run_if_flag = guard;
      rc = sweep_generic(st, k);
    }
  } else {
    rc = sweep_generic((cudaStream_t)stream, k);
  }
  // kernels without the fused reduction (and the hand-over targets of the tuned one, guarded by its flag): one pass of K6
  if (rc == NK_OK && a->stats_out != nullptr && a->chain_length > 0)
    rc = stats_partial((cudaStream_t)stream, a->eloc_out, a->eloc_dtype, ch->B, a->chain_length, 1, a->stats_shift, a->stats_out, stats_guard);
  if (rc == NK_OK) ch->t += (uint64_t)(a->n_discard + a->chain_length) * (uint64_t)a->sweep_size;
  return rc;
}

int nk_ising_conn(void *stream, const nk_ising_t *op, const int8_t *x, int64_t B, int32_t N, int8_t *xp_out, void *mels_out,
                  int32_t mel_dtype) {
  NK_CHECK_ARG(op != nullptr, "nk_ising_conn: operator is NULL");
  NK_CHECK_ARG(B >= 0 && N > 0, "nk_ising_conn: bad shape");
  NK_CHECK_ARG(op->n_edges >= 0 && (op->n_edges == 0 || op->edges), "nk_ising_conn: bad edges");
  NK_CHECK_ARG(mel_dtype == NK_F32 || mel_dtype == NK_F64, "nk_ising_conn: bad mel_dtype");
  NK_CHECK_ARG(B == 0 || (x && xp_out && mels_out), "nk_ising_conn: NULL buffer");
  return ising_conn((cudaStream_t)stream, *op, x, B, N, xp_out, mels_out, mel_dtype);
}

int nk_ising_n_conn(void *stream, const nk_ising_t *op, const int8_t *x, int64_t B, int32_t N, int32_t *nconn_out) {
  NK_CHECK_ARG(op != nullptr, "nk_ising_n_conn: operator is NULL");
  NK_CHECK_ARG(B >= 0 && N > 0, "nk_ising_n_conn: bad shape");
  NK_CHECK_ARG(B == 0 || (x && nconn_out), "nk_ising_n_conn: NULL buffer");
  return ising_n_conn((cudaStream_t)stream, *op, x, B, N, nconn_out);
}

int nk_localop_conn(void *stream, const nk_localop_t *op, const int8_t *x, int64_t B, int32_t N, int8_t *xp_out, void *mels_out,
                    int32_t mel_dtype, int32_t *nconn_out) {
  int rc = check_localop(op, N, "nk_localop_conn");
  if (rc) return rc;
  NK_CHECK_ARG(B >= 0 && N > 0, "nk_localop_conn: bad shape");
  NK_CHECK_ARG(mel_dtype == NK_F32 || mel_dtype == NK_F64, "nk_localop_conn: bad mel_dtype");
  NK_CHECK_ARG(B == 0 || op->max_conn_size == 0 || (x && xp_out && mels_out), "nk_localop_conn: NULL buffer");
  return localop_conn((cudaStream_t)stream, *op, x, B, N, xp_out, mels_out, mel_dtype, nconn_out);
}

// Stand-alone local estimator: the product-form kernel in its `eloc_only` mode (theta GEMM, (A, B) from theta, the fused
// local-energy code of the sweep kernel), or the theta-form generic kernel (no workspace, NK_PATH_GENERIC, or in-stream
// hand-over when the weights are outside the product form's range).
static int eloc_dispatch(cudaStream_t st, const nk_rbm_t *rbm, const nk_ising_t *ising, const nk_localop_t *localop,
                         const int8_t *sigma, int64_t B, void *eloc_out, int32_t eloc_dtype, int32_t path, void *workspace,
                         const char *who) {
  NK_CHECK_ARG(path >= NK_PATH_AUTO && path <= NK_PATH_PROD, "%s: bad path %d", who, path);
  if (B == 0) return NK_OK;
  SweepKernelArgs k{};
  k.rbm = *rbm;
  k.sigma = const_cast<int8_t *>(sigma);  // not written in eloc_only mode
  k.B = B;
  k.rule = NK_RULE_LOCAL;
  k.chain_length = 1;
  k.n_discard = 0;
  k.sweep_size = 1;
  k.machine_pow = 2.0;
  k.eloc_kind = ising ? 1 : 2;
  if (ising) k.ising = *ising;
  if (localop) k.localop = *localop;
  k.eloc_out = eloc_out;
  k.eloc_dtype = eloc_dtype;
  k.eloc_only = 1;
  const bool prod_ok = workspace != nullptr && sweep_prod_supported(k);
  if ((path == NK_PATH_FAST || path == NK_PATH_PROD) && !prod_ok) {
    set_error("%s: no product-form kernel for this configuration (needs N<=1024, M<=8192, 1- and 2-site operator terms and a "
              "workspace of nk_sweep_workspace_bytes())", who);
    return NK_EUNSUPPORTED;
  }
  if (path == NK_PATH_GENERIC || !prod_ok) return eloc_generic(st, *rbm, ising, localop, sigma, B, eloc_out, eloc_dtype, nullptr);
  char *wsb = reinterpret_cast<char *>(workspace);
  void *theta = wsb;
  int *flags = reinterpret_cast<int *>(wsb + ws_theta_bytes(rbm, B));
  void *tables = reinterpret_cast<char *>(flags) + 256;
  void *scratch = reinterpret_cast<char *>(tables) + sweep_prod_workspace_bytes(*rbm);
  NK_CUDA_OK(cudaMemsetAsync(flags, 0, 256, st));
  int rc = theta_gemm(st, *rbm, sigma, B, theta, scratch);
  if (rc) return rc;
  rc = sweep_prod(st, k, theta, flags, tables);
  if (rc == NK_OK) rc = eloc_generic(st, *rbm, ising, localop, sigma, B, eloc_out, eloc_dtype, flags);  // in-stream hand-over
  return rc;
}

int nk_eloc_ising_rbm(void *stream, const nk_rbm_t *rbm, const nk_ising_t *op, const int8_t *sigma, int64_t B, void *eloc_out,
                      int32_t eloc_dtype, int32_t path, void *workspace) {
  int rc = check_rbm(rbm, "nk_eloc_ising_rbm");
  if (rc) return rc;
  NK_CHECK_ARG(op != nullptr, "nk_eloc_ising_rbm: operator is NULL");
  NK_CHECK_ARG(op->n_edges >= 0 && (op->n_edges == 0 || op->edges), "nk_eloc_ising_rbm: bad edges");
  NK_CHECK_ARG(B >= 0 && (B == 0 || (sigma && eloc_out)), "nk_eloc_ising_rbm: bad batch / NULL buffer");
  NK_CHECK_ARG(eloc_dtype == NK_F32 || eloc_dtype == NK_F64, "nk_eloc_ising_rbm: bad eloc_dtype");
  return eloc_dispatch((cudaStream_t)stream, rbm, op, nullptr, sigma, B, eloc_out, eloc_dtype, path, workspace, "nk_eloc_ising_rbm");
}

int nk_eloc_localop_rbm(void *stream, const nk_rbm_t *rbm, const nk_localop_t *op, const int8_t *sigma, int64_t B, void *eloc_out,
                        int32_t eloc_dtype, int32_t path, void *workspace) {
  int rc = check_rbm(rbm, "nk_eloc_localop_rbm");
  if (rc) return rc;
  rc = check_localop(op, rbm->N, "nk_eloc_localop_rbm");
  if (rc) return rc;
  NK_CHECK_ARG(B >= 0 && (B == 0 || (sigma && eloc_out)), "nk_eloc_localop_rbm: bad batch / NULL buffer");
  NK_CHECK_ARG(eloc_dtype == NK_F32 || eloc_dtype == NK_F64, "nk_eloc_localop_rbm: bad eloc_dtype");
  return eloc_dispatch((cudaStream_t)stream, rbm, nullptr, op, sigma, B, eloc_out, eloc_dtype, path, workspace, "nk_eloc_localop_rbm");
}

int64_t nk_forces_workspace_bytes(const nk_rbm_t *rbm, int64_t Ns) {
  if (check_rbm(rbm, "nk_forces_workspace_bytes") || Ns < 0) return -1;
  return (int64_t)(ws_theta_bytes(rbm, Ns) + (size_t)theta_gemm_workspace_bytes(*rbm, Ns));
}

int nk_forces_rbm(void *stream, const nk_rbm_t *rbm, const int8_t *samples, int64_t Ns, const void *eloc, int32_t eloc_dtype,
                  double mean, double *sums, void *workspace, const void *tanh_theta) {
  int rc = check_rbm(rbm, "nk_forces_rbm");
  if (rc) return rc;
  NK_CHECK_ARG(Ns >= 0 && sums != nullptr, "nk_forces_rbm: bad arguments");
  NK_CHECK_ARG(Ns == 0 || (samples && eloc && (workspace || tanh_theta)), "nk_forces_rbm: NULL buffer");
  NK_CHECK_ARG(eloc_dtype == NK_F32 || eloc_dtype == NK_F64, "nk_forces_rbm: bad eloc_dtype");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)rbm->N * rbm->M + rbm->M + rbm->N;
  NK_CUDA_OK(cudaMemsetAsync(sums, 0, n * sizeof(double), st));
  if (Ns == 0) return NK_OK;
  const void *theta = tanh_theta;
  const int is_tanh = tanh_theta != nullptr ? 1 : 0;
  if (!is_tanh) {  // recompute theta for the whole batch
    void *scratch = reinterpret_cast<char *>(workspace) + ws_theta_bytes(rbm, Ns);
    rc = theta_gemm(st, *rbm, samples, Ns, workspace, scratch);
    if (rc) return rc;
    theta = workspace;
  }
  // fp32: the contraction over the samples runs on the tensor cores (forces_tc.cu); fp64 / large shapes: CUDA cores
  if (forces_tc_supported(*rbm) && getenv("NKB200_FORCES_CUDA_CORE") == nullptr)
    return forces_tc_sums(st, *rbm, samples, theta, is_tanh, eloc, eloc_dtype, Ns, mean, sums);
  return forces_sums(st, *rbm, samples, theta, is_tanh, eloc, eloc_dtype, Ns, mean, sums);
}

int nk_forces_finalize(void *stream, const double *sums, double scale, int64_t n, void *out, int32_t dtype) {
  NK_CHECK_ARG(n >= 0 && (n == 0 || (sums && out)), "nk_forces_finalize: bad arguments");
  NK_CHECK_ARG(dtype == NK_F32 || dtype == NK_F64, "nk_forces_finalize: bad dtype");
  return forces_finalize((cudaStream_t)stream, sums, scale, n, out, dtype);
}

int nk_stats_partial(void *stream, const void *data, int32_t dtype, int64_t n_chains, int64_t L, int32_t phase, double shift,
                     double *partials_out) {
  NK_CHECK_ARG(dtype == NK_F32 || dtype == NK_F64, "nk_stats_partial: bad dtype");
  NK_CHECK_ARG(n_chains >= 0 && L >= 0 && partials_out, "nk_stats_partial: bad arguments");
  NK_CHECK_ARG(phase == 0 || phase == 1, "nk_stats_partial: phase must be 0 or 1");
  NK_CHECK_ARG(n_chains * L == 0 || data, "nk_stats_partial: NULL data");
  return stats_partial((cudaStream_t)stream, data, dtype, n_chains, L, phase, shift, partials_out);
}

int nk_stats_tau(void *stream, const void *data, int32_t dtype, int64_t n_chains, int64_t L, double c, double *out) {
  NK_CHECK_ARG(dtype == NK_F32 || dtype == NK_F64, "nk_stats_tau: bad dtype");
  NK_CHECK_ARG(n_chains >= 0 && L >= 0 && out && c > 0.0, "nk_stats_tau: bad arguments");
  NK_CHECK_ARG(n_chains * L == 0 || data, "nk_stats_tau: NULL data");
  return stats_tau((cudaStream_t)stream, data, dtype, n_chains, L, c, out);
}

double nk_stats_tau_max_decode(double encoded) {
  unsigned long long e;
  memcpy(&e, &encoded, 8);
  if (e == 0ull) return NAN;  // no chain contributed
  const unsigned long long b = (e >> 63) ? (e & 0x7fffffffffffffffull) : ~e;
  double v;
  memcpy(&v, &b, 8);
  return v;
}

int nk_rbm_tanh_theta(void *stream, const nk_rbm_t *rbm, const int8_t *samples, int64_t Ns, void *out, void *workspace) {
  int rc = check_rbm(rbm, "nk_rbm_tanh_theta");
  if (rc) return rc;
  NK_CHECK_ARG(Ns >= 0, "nk_rbm_tanh_theta: Ns=%lld", (long long)Ns);
  NK_CHECK_ARG(Ns == 0 || (samples && out), "nk_rbm_tanh_theta: NULL buffer");
  rc = theta_gemm((cudaStream_t)stream, *rbm, samples, Ns, out, workspace);
  if (rc) return rc;
  return rbm_tanh_inplace((cudaStream_t)stream, out, rbm->dtype, Ns * (int64_t)rbm->M);
}

int nk_rbm_jvp(void *stream, const nk_rbm_t *v, const int8_t *samples, int64_t Ns, const void *tanh_theta, double *y_out,
               double *y_sum_out, void *scratch, void *workspace) {
  int rc = check_rbm(v, "nk_rbm_jvp");
  if (rc) return rc;
  NK_CHECK_ARG(Ns >= 0, "nk_rbm_jvp: Ns=%lld", (long long)Ns);
  NK_CHECK_ARG(Ns == 0 || (samples && tanh_theta && y_out && scratch), "nk_rbm_jvp: NULL buffer");
  rc = theta_gemm((cudaStream_t)stream, *v, samples, Ns, scratch, workspace);  // sigma V + v_b on the tensor cores
  if (rc) return rc;
  return rbm_jvp_dot((cudaStream_t)stream, *v, samples, Ns, tanh_theta, scratch, y_out, y_sum_out);
}

static int online_state_ok(const nk_online_stats_t *s, const char *who) {
  NK_CHECK_ARG(s, "%s: NULL state", who);
  NK_CHECK_ARG(s->n_chains >= 0 && s->max_lag >= 0 && s->max_lag <= NK_ONLINE_MAX_LAG, "%s: bad n_chains / max_lag (max_lag <= %d)", who,
               NK_ONLINE_MAX_LAG);
  NK_CHECK_ARG(s->buf_len >= 0 && s->buf_len <= s->max_lag, "%s: buf_len must lie in [0, max_lag]", who);
  if (s->n_chains > 0) {
    NK_CHECK_ARG(s->chain_count && s->chain_mean && s->chain_M2, "%s: NULL per-chain array", who);
    NK_CHECK_ARG(s->max_lag == 0 || (s->cross_sum && s->m1_sum && s->m2_sum && s->pair_count && s->chain_buf), "%s: NULL lag array", who);
  }
  return NK_OK;
}

int nk_online_stats_update(void *stream, const nk_online_stats_t *in, const nk_online_stats_t *out, const void *data, int32_t dtype,
                           int64_t n, double decay) {
  if (int rc = online_state_ok(in, "nk_online_stats_update")) return rc;
  if (int rc = online_state_ok(out, "nk_online_stats_update")) return rc;
  NK_CHECK_ARG(dtype == NK_F32 || dtype == NK_F64, "nk_online_stats_update: bad dtype");
  NK_CHECK_ARG(in->n_chains == out->n_chains && in->max_lag == out->max_lag, "nk_online_stats_update: in / out shapes differ");
  NK_CHECK_ARG(n >= 1, "nk_online_stats_update: the batch must hold at least one sample per chain");
  NK_CHECK_ARG(decay > 0.0 && decay <= 1.0, "nk_online_stats_update: decay must lie in (0, 1]");
  if (in->n_chains == 0) return NK_OK;
  NK_CHECK_ARG(data, "nk_online_stats_update: NULL data");
  return online_stats_update((cudaStream_t)stream, in, out, data, dtype, n, decay);
}

int nk_online_stats_summary(void *stream, const nk_online_stats_t *state, int32_t phase, double gmean, double mbar, double *sums_out) {
  if (int rc = online_state_ok(state, "nk_online_stats_summary")) return rc;
  NK_CHECK_ARG(phase == 0 || phase == 1, "nk_online_stats_summary: phase must be 0 or 1");
  NK_CHECK_ARG(sums_out, "nk_online_stats_summary: NULL output");
  return online_stats_summary((cudaStream_t)stream, state, phase, gmean, mbar, sums_out);
}

int nk_online_stats_finalize(const double *phase0_host, const double *phase1_host, int64_t n_chains_total, int64_t n_samples_total,
                             int32_t max_lag, double *out_host, double *acf_host) {
  NK_CHECK_ARG(phase0_host && phase1_host && out_host, "nk_online_stats_finalize: NULL argument");
  NK_CHECK_ARG(n_chains_total > 0 && n_samples_total >= 0 && max_lag >= 0 && max_lag <= NK_ONLINE_MAX_LAG,
               "nk_online_stats_finalize: bad arguments");
  return online_stats_finalize(phase0_host, phase1_host, n_chains_total, n_samples_total, max_lag, out_host, acf_host);
}

int nk_stats_finalize(const double *sums_host, double mean, int64_t n_chains_total, int64_t L, double *out_host) {
  NK_CHECK_ARG(sums_host && out_host && n_chains_total > 0 && L > 0, "nk_stats_finalize: bad arguments");
  return stats_finalize(sums_host, mean, n_chains_total, L, out_host);
}

// ------------------------------------------------------------------------------------------ host-buffer context
struct nk_ctx {
  int device;
  int32_t N, M, dtype, chain_length, sweep_size, rule, n_clusters, n_down, return_samples, op_kind;  // op_kind 1 ising, 2 localop
  int64_t B;
  size_t esz, eloc_esz;
  int32_t eloc_dtype;
  double machine_pow;
  cudaStream_t stream;
  bool own_stream;
  void *W, *b, *a;
  int8_t *sigma, *samples;
  void *log_prob;
  int64_t *n_accepted;
  int32_t *edges, *clusters;
  double *cluster_probs;
  void *eloc;
  double *partials;       // device, NK_CTX_NPARTIAL: [NK_STATS_NPARTIAL sums | n_chains | n_accepted]
  double *partials_host;  // pinned
  void *workspace;        // nk_sweep_workspace_bytes (theta scratch + hand-over flag)
  nk_ising_t ising;
  nk_localop_t localop;
  std::vector<void *> op_allocs;
  uint64_t seed, t, chain_offset;
  double shift;           // shift of the in-kernel statistics: the previous step's mean
  int32_t n_discard_last;
  bool has_b, has_a;
};

__global__ void ctx_tail_kernel(const int64_t *nacc, int64_t n, double *partials) {
  // partials[NPARTIAL] = n_chains, partials[NPARTIAL + 1] = sum of the acceptance counters (exact in double below 2^53)
  unsigned long long s = 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += (unsigned long long)nacc[i];
  for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  __shared__ unsigned long long sh[32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    partials[NK_STATS_NPARTIAL] = (double)n;
    partials[NK_STATS_NPARTIAL + 1] = (double)t;
  }
}

}  // extern "C"

template <typename T>
static int upload(void **dst, const T *src_host, size_t count, cudaStream_t st, std::vector<void *> *owned) {
  *dst = nullptr;
  if (count == 0 || src_host == nullptr) return NK_OK;
  NK_CUDA_OK(cudaMalloc(dst, count * sizeof(T)));
  if (owned) owned->push_back(*dst);
  NK_CUDA_OK(cudaMemcpyAsync(*dst, src_host, count * sizeof(T), cudaMemcpyHostToDevice, st));
  return NK_OK;
}

extern "C" {

int nk_ctx_create2(nk_ctx **out, const nk_ctx_desc_t *d) {
  NK_CHECK_ARG(out != nullptr && d != nullptr, "nk_ctx_create2: NULL argument");
  NK_CHECK_ARG(d->N > 0 && d->M > 0 && d->n_chains > 0 && d->chain_length > 0, "nk_ctx_create2: bad sizes");
  NK_CHECK_ARG(d->dtype == NK_F32 || d->dtype == NK_F64, "nk_ctx_create2: bad dtype");
  NK_CHECK_ARG(d->rule == NK_RULE_LOCAL || d->rule == NK_RULE_EXCHANGE, "nk_ctx_create2: unknown rule %d", d->rule);
  NK_CHECK_ARG(d->rule != NK_RULE_EXCHANGE || (d->clusters_host != nullptr && d->n_clusters > 0), "nk_ctx_create2: ExchangeRule needs clusters");
  NK_CHECK_ARG(d->machine_pow >= 0.0, "nk_ctx_create2: machine_pow must be a non-negative real");
  NK_CHECK_ARG(d->sweep_size >= 0, "nk_ctx_create2: bad sweep_size");
  NK_CHECK_ARG((d->ising_host != nullptr) != (d->localop_host != nullptr), "nk_ctx_create2: pass exactly one operator (ising_host / localop_host)");
  NK_CHECK_ARG(d->n_down <= d->N, "nk_ctx_create2: n_down > N");
  if (d->ising_host) NK_CHECK_ARG(d->ising_host->n_edges >= 0 && (d->ising_host->n_edges == 0 || d->ising_host->edges), "nk_ctx_create2: bad edges");
  if (d->localop_host) {
    int rc = check_localop(d->localop_host, d->N, "nk_ctx_create2");
    if (rc) return rc;
  }
  NK_CUDA_OK(cudaSetDevice(d->device));
  nk_ctx *c = new nk_ctx();
  c->device = d->device;
  c->N = d->N;
  c->M = d->M;
  c->dtype = d->dtype;
  c->B = d->n_chains;
  c->chain_length = d->chain_length;
  c->sweep_size = d->sweep_size > 0 ? d->sweep_size : d->N;
  c->rule = d->rule;
  c->n_clusters = d->rule == NK_RULE_EXCHANGE ? d->n_clusters : 0;
  c->n_down = d->n_down;
  c->return_samples = d->return_samples;
  c->machine_pow = d->machine_pow;
  c->esz = d->dtype == NK_F32 ? 4 : 8;
  c->seed = d->seed;
  c->t = 0;
  c->chain_offset = d->chain_offset;
  c->shift = 0.0;
  c->n_discard_last = 0;
  c->W = c->b = c->a = nullptr;
  c->sigma = c->samples = nullptr;
  c->log_prob = c->eloc = c->workspace = nullptr;
  c->n_accepted = nullptr;
  c->edges = c->clusters = nullptr;
  c->cluster_probs = nullptr;
  c->partials = c->partials_host = nullptr;
  if (d->stream != nullptr) {
    c->stream = (cudaStream_t)d->stream;
    c->own_stream = false;
  } else {
    NK_CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
  }
  const int64_t B = c->B;
  const int N = c->N, M = c->M;
  NK_CUDA_OK(cudaMalloc(&c->W, (size_t)N * M * c->esz));
  NK_CUDA_OK(cudaMalloc(&c->b, (size_t)M * c->esz));
  NK_CUDA_OK(cudaMalloc(&c->a, (size_t)N * c->esz));
  NK_CUDA_OK(cudaMalloc((void **)&c->sigma, (size_t)B * N));
  NK_CUDA_OK(cudaMalloc(&c->log_prob, (size_t)B * c->esz));
  NK_CUDA_OK(cudaMalloc((void **)&c->n_accepted, (size_t)B * 8));
  NK_CUDA_OK(cudaMalloc((void **)&c->partials, sizeof(double) * NK_CTX_NPARTIAL));
  NK_CUDA_OK(cudaMallocHost((void **)&c->partials_host, sizeof(double) * NK_CTX_NPARTIAL));
  if (c->return_samples) NK_CUDA_OK(cudaMalloc((void **)&c->samples, (size_t)B * c->chain_length * N));
  // operator tables: host -> device
  int rc = NK_OK;
  if (d->ising_host) {
    c->op_kind = 1;
    c->ising = *d->ising_host;
    void *p = nullptr;
    rc = upload<int32_t>(&p, d->ising_host->edges, (size_t)2 * d->ising_host->n_edges, c->stream, &c->op_allocs);
    if (rc) return rc;
    c->ising.edges = (const int32_t *)p;
    c->eloc_dtype = NK_F64;  // operator dtype float64 (netket/operator/_ising/base.py:75): E_loc promotes to double ...
  } else {
    c->op_kind = 2;
    c->localop = *d->localop_host;
    for (int g = 0; g < c->localop.n_groups; ++g) {
      const nk_localop_group_t &h = d->localop_host->groups[g];
      nk_localop_group_t &G = c->localop.groups[g];
      const size_t rows = (size_t)h.n_ops << h.n_sites;
      void *p = nullptr;
      if ((rc = upload<int32_t>(&p, h.acting_on, (size_t)h.n_ops * h.n_sites, c->stream, &c->op_allocs))) return rc;
      G.acting_on = (const int32_t *)p;
      if ((rc = upload<double>(&p, h.diag_mels, rows, c->stream, &c->op_allocs))) return rc;
      G.diag_mels = (const double *)p;
      if ((rc = upload<int32_t>(&p, h.n_conns, rows, c->stream, &c->op_allocs))) return rc;
      G.n_conns = (const int32_t *)p;
      if ((rc = upload<double>(&p, h.mels, rows * h.ncmax, c->stream, &c->op_allocs))) return rc;
      G.mels = (const double *)p;
      if ((rc = upload<int8_t>(&p, h.x_prime, rows * h.ncmax * h.n_sites, c->stream, &c->op_allocs))) return rc;
      G.x_prime = (const int8_t *)p;
    }
    c->eloc_dtype = NK_F64;
  }
  // ... unless the caller asked for E_loc in the parameter dtype (the C-ABI v1 behaviour, kept by nk_ctx_create)
  if (d->eloc_in_param_dtype) c->eloc_dtype = c->dtype;
  c->eloc_esz = c->eloc_dtype == NK_F32 ? 4 : 8;
  NK_CUDA_OK(cudaMalloc(&c->eloc, (size_t)B * c->chain_length * c->eloc_esz));
  if (c->rule == NK_RULE_EXCHANGE) {
    void *p = nullptr;
    if ((rc = upload<int32_t>(&p, d->clusters_host, (size_t)2 * d->n_clusters, c->stream, nullptr))) return rc;
    c->clusters = (int32_t *)p;
    if ((rc = upload<double>(&p, d->cluster_probs_host, (size_t)d->n_clusters, c->stream, nullptr))) return rc;
    c->cluster_probs = (double *)p;
  }
  {
    nk_rbm_t shape{};
    shape.W = c->W;
    shape.N = N;
    shape.M = M;
    shape.dtype = c->dtype;
    const int64_t wsb = nk_sweep_workspace_bytes(&shape, B);
    if (wsb > 0) NK_CUDA_OK(cudaMalloc(&c->workspace, (size_t)wsb));
  }
  rc = random_state(c->stream, c->sigma, B, N, c->n_down, c->seed, c->chain_offset);
  if (rc) return rc;
  NK_CUDA_OK(cudaStreamSynchronize(c->stream));
  *out = c;
  return NK_OK;
}

int nk_ctx_create(nk_ctx **out, int32_t device, int32_t N, int32_t M, int32_t dtype, int64_t n_chains, int32_t chain_length,
                  const int32_t *edges_host, int32_t n_edges, double h, double J, uint64_t seed, uint64_t chain_offset) {
  NK_CHECK_ARG(out != nullptr, "nk_ctx_create: out is NULL");
  NK_CHECK_ARG(N > 0 && M > 0 && n_chains > 0 && chain_length > 0, "nk_ctx_create: bad sizes");
  NK_CHECK_ARG(dtype == NK_F32 || dtype == NK_F64, "nk_ctx_create: bad dtype");
  NK_CHECK_ARG(n_edges >= 0 && (n_edges == 0 || edges_host), "nk_ctx_create: bad edges");
  nk_ising_t op{};
  op.edges = edges_host;
  op.n_edges = n_edges;
  op.h = h;
  op.J = J;
  nk_ctx_desc_t d{};
  d.device = device;
  d.N = N;
  d.M = M;
  d.dtype = dtype;
  d.n_chains = n_chains;
  d.chain_length = chain_length;
  d.rule = NK_RULE_LOCAL;
  d.machine_pow = 2.0;
  d.n_down = -1;
  d.ising_host = &op;
  d.seed = seed;
  d.chain_offset = chain_offset;
  d.eloc_in_param_dtype = 1;
  return nk_ctx_create2(out, &d);
}

void nk_ctx_destroy(nk_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  cudaFree(c->W);
  cudaFree(c->b);
  cudaFree(c->a);
  cudaFree(c->sigma);
  cudaFree(c->samples);
  cudaFree(c->log_prob);
  cudaFree(c->n_accepted);
  cudaFree(c->clusters);
  cudaFree(c->cluster_probs);
  cudaFree(c->eloc);
  cudaFree(c->partials);
  cudaFree(c->workspace);
  for (void *p : c->op_allocs) cudaFree(p);
  cudaFreeHost(c->partials_host);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
}

void *nk_ctx_stream(nk_ctx *c) { return c ? (void *)c->stream : nullptr; }
double *nk_ctx_partials_device(nk_ctx *c) { return c ? c->partials : nullptr; }

int nk_ctx_step_begin(nk_ctx *c, const void *W_host, const void *b_host, const void *a_host, int32_t n_discard) {
  NK_CHECK_ARG(c && W_host, "nk_ctx_step_begin: NULL argument");
  NK_CHECK_ARG(n_discard >= 0, "nk_ctx_step_begin: n_discard < 0");
  NK_CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  NK_CUDA_OK(cudaMemcpyAsync(c->W, W_host, (size_t)c->N * c->M * c->esz, cudaMemcpyHostToDevice, st));
  if (b_host) NK_CUDA_OK(cudaMemcpyAsync(c->b, b_host, (size_t)c->M * c->esz, cudaMemcpyHostToDevice, st));
  if (a_host) NK_CUDA_OK(cudaMemcpyAsync(c->a, a_host, (size_t)c->N * c->esz, cudaMemcpyHostToDevice, st));
  c->has_b = b_host != nullptr;
  c->has_a = a_host != nullptr;
  nk_rbm_t rbm{};
  rbm.W = c->W;
  rbm.b = b_host ? c->b : nullptr;
  rbm.a = a_host ? c->a : nullptr;
  rbm.N = c->N;
  rbm.M = c->M;
  rbm.dtype = c->dtype;
  // MCState.sample always resets first: acceptance counters are zeroed per call (mc_state/state.py:555-557)
  NK_CUDA_OK(cudaMemsetAsync(c->n_accepted, 0, (size_t)c->B * 8, st));
  nk_chains_t ch{};
  ch.sigma = c->sigma;
  ch.log_prob = c->log_prob;
  ch.n_accepted = c->n_accepted;
  ch.workspace = c->workspace;
  ch.B = c->B;
  ch.seed = c->seed;
  ch.t = c->t;
  ch.chain_offset = c->chain_offset;
  nk_sweep_t a{};
  a.rule = c->rule;
  a.chain_length = c->chain_length;
  a.n_discard = n_discard;
  a.sweep_size = c->sweep_size;
  a.machine_pow = c->machine_pow;
  a.samples_out = c->samples;
  a.clusters = c->clusters;
  a.n_clusters = c->n_clusters;
  a.cluster_probs = c->cluster_probs;
  if (c->op_kind == 1)
    a.ising = &c->ising;
  else
    a.localop = &c->localop;
  a.eloc_out = c->eloc;
  a.eloc_dtype = c->eloc_dtype;
  a.path = NK_PATH_AUTO;
  a.stats_out = c->partials;  // the statistics' sums come out of the sweep kernel, shifted by the previous step's mean
  a.stats_shift = c->shift;
  int rc = nk_sweep(st, &rbm, &ch, &a);
  if (rc) return rc;
  c->t = ch.t;
  c->n_discard_last = n_discard;
  ctx_tail_kernel<<<1, 1024, 0, st>>>(c->n_accepted, c->B, c->partials);
  NK_LAUNCH_OK();
  return NK_OK;
}

int nk_ctx_step_end(nk_ctx *c, void *eloc_host, int8_t *samples_host, double *stats_host) {
  NK_CHECK_ARG(c && eloc_host && stats_host, "nk_ctx_step_end: NULL argument");
  NK_CHECK_ARG(samples_host == nullptr || c->return_samples, "nk_ctx_step_end: the context was created without return_samples");
  NK_CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const int64_t L = c->chain_length;
  NK_CUDA_OK(cudaMemcpyAsync(eloc_host, c->eloc, (size_t)c->B * L * c->eloc_esz, cudaMemcpyDeviceToHost, st));
  if (samples_host) NK_CUDA_OK(cudaMemcpyAsync(samples_host, c->samples, (size_t)c->B * L * c->N, cudaMemcpyDeviceToHost, st));
  NK_CUDA_OK(cudaMemcpyAsync(c->partials_host, c->partials, sizeof(double) * NK_CTX_NPARTIAL, cudaMemcpyDeviceToHost, st));
  NK_CUDA_OK(cudaStreamSynchronize(st));  // the one host synchronisation of a step
  const double *p = c->partials_host;
  const int64_t n_chains_total = (int64_t)llround(p[NK_STATS_NPARTIAL]);
  NK_CHECK_ARG(n_chains_total > 0, "nk_ctx_step_end: the partial sums hold no chains (all-reduce gone wrong?)");
  const double ts = (double)n_chains_total * (double)L;
  const double dm = p[7] / ts, var = p[0] / ts - dm * dm;
  int status = NK_OK;
  if (!(dm * dm * (double)L <= 1.0e5 * var)) {
    // the shift was too far from the mean for the one-pass formulas (first step, or a low-variance state)
    if (n_chains_total == c->B) {  // single device: redo the sums in two passes around the mean just found
      int rc = stats_partial(st, c->eloc, c->eloc_dtype, c->B, L, 1, c->shift + dm, c->partials);
      if (rc) return rc;
      NK_CUDA_OK(cudaMemcpyAsync(c->partials_host, c->partials, sizeof(double) * NK_STATS_NPARTIAL, cudaMemcpyDeviceToHost, st));
      NK_CUDA_OK(cudaStreamSynchronize(st));
      c->shift += dm;
    } else {
      status = NK_RESHIFT;  // multi-device: statistics below are valid to reduced precision; the next step's shift is the mean found
    }
  }
  stats_finalize(c->partials_host, c->shift, n_chains_total, L, stats_host);
  const double n_steps = (double)n_chains_total * (double)(c->n_discard_last + c->chain_length) * (double)c->sweep_size;
  stats_host[5] = p[NK_STATS_NPARTIAL + 1] / n_steps;
  c->shift = stats_host[0];
  return status;
}

int nk_ctx_step_host(nk_ctx *c, const void *W_host, const void *b_host, const void *a_host, int32_t n_discard, void *eloc_host,
                     double *stats_host) {
  NK_CHECK_ARG(c && W_host && eloc_host && stats_host, "nk_ctx_step_host: NULL argument");
  int rc = nk_ctx_step_begin(c, W_host, b_host, a_host, n_discard);
  if (rc) return rc;
  rc = nk_ctx_step_end(c, eloc_host, nullptr, stats_host);
  return rc == NK_RESHIFT ? NK_OK : rc;
}

int nk_ctx_get_sigma_host(nk_ctx *c, int8_t *sigma_host) {
  NK_CHECK_ARG(c && sigma_host, "nk_ctx_get_sigma_host: NULL argument");
  NK_CUDA_OK(cudaSetDevice(c->device));
  NK_CUDA_OK(cudaMemcpyAsync(sigma_host, c->sigma, (size_t)c->B * c->N, cudaMemcpyDeviceToHost, c->stream));
  NK_CUDA_OK(cudaStreamSynchronize(c->stream));
  return NK_OK;
}

}  // extern "C"
