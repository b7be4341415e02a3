// extern "C" entry points of libnkb200 (see include/nkb200.h): argument validation, path selection, launches.
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
      if (rc == NK_OK && a->path != NK_PATH_FAST && sweep_prod_supported(k)) {
        rc = sweep_prod(st, k, theta, flags, tables, flags, 5);
        guard = flags + 5;
      }
    } else {
      rc = sweep_prod(st, k, theta, flags, tables);
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
  int32_t N, M, dtype, chain_length, n_edges;
  int64_t B;
  size_t esz;
  cudaStream_t stream;
  void *W, *b, *a;
  int8_t *sigma;
  void *log_prob;
  int64_t *n_accepted;
  int32_t *edges;
  void *eloc;
  double *partials;       // device, NK_STATS_NPARTIAL
  double *partials_host;  // pinned
  int64_t *nacc_host;     // pinned, 1
  int64_t *nacc_sum;      // device, 1
  void *workspace;        // nk_sweep_workspace_bytes (theta scratch + hand-over flag)
  nk_ising_t ising;
  uint64_t seed, t, chain_offset;
};

__global__ void acc_sum_kernel(const int64_t *x, int64_t n, unsigned long long *out) {
  unsigned long long s = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s += (unsigned long long)x[i];
  for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

int nk_ctx_create(nk_ctx **out, int32_t device, int32_t N, int32_t M, int32_t dtype, int64_t n_chains, int32_t chain_length,
                  const int32_t *edges_host, int32_t n_edges, double h, double J, uint64_t seed, uint64_t chain_offset) {
  NK_CHECK_ARG(out != nullptr, "nk_ctx_create: out is NULL");
  NK_CHECK_ARG(N > 0 && M > 0 && n_chains > 0 && chain_length > 0, "nk_ctx_create: bad sizes");
  NK_CHECK_ARG(dtype == NK_F32 || dtype == NK_F64, "nk_ctx_create: bad dtype");
  NK_CHECK_ARG(n_edges >= 0 && (n_edges == 0 || edges_host), "nk_ctx_create: bad edges");
  NK_CUDA_OK(cudaSetDevice(device));
  nk_ctx *c = new nk_ctx();
  memset(c, 0, sizeof(*c));
  c->device = device;
  c->N = N;
  c->M = M;
  c->dtype = dtype;
  c->B = n_chains;
  c->chain_length = chain_length;
  c->n_edges = n_edges;
  c->esz = dtype == NK_F32 ? 4 : 8;
  c->seed = seed;
  c->t = 0;
  c->chain_offset = chain_offset;
  NK_CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  NK_CUDA_OK(cudaMalloc(&c->W, (size_t)N * M * c->esz));
  NK_CUDA_OK(cudaMalloc(&c->b, (size_t)M * c->esz));
  NK_CUDA_OK(cudaMalloc(&c->a, (size_t)N * c->esz));
  NK_CUDA_OK(cudaMalloc((void **)&c->sigma, (size_t)n_chains * N));
  NK_CUDA_OK(cudaMalloc(&c->log_prob, (size_t)n_chains * c->esz));
  NK_CUDA_OK(cudaMalloc((void **)&c->n_accepted, (size_t)n_chains * 8));
  NK_CUDA_OK(cudaMalloc((void **)&c->edges, (size_t)(n_edges > 0 ? n_edges : 1) * 8));
  NK_CUDA_OK(cudaMalloc(&c->eloc, (size_t)n_chains * chain_length * c->esz));
  NK_CUDA_OK(cudaMalloc((void **)&c->partials, sizeof(double) * NK_STATS_NPARTIAL));
  NK_CUDA_OK(cudaMalloc((void **)&c->nacc_sum, 8));
  {
    nk_rbm_t shape{};
    shape.W = c->W;
    shape.N = N;
    shape.M = M;
    shape.dtype = dtype;
    const int64_t wsb = nk_sweep_workspace_bytes(&shape, n_chains);
    if (wsb > 0) NK_CUDA_OK(cudaMalloc(&c->workspace, (size_t)wsb));
  }
  NK_CUDA_OK(cudaMallocHost((void **)&c->partials_host, sizeof(double) * NK_STATS_NPARTIAL));
  NK_CUDA_OK(cudaMallocHost((void **)&c->nacc_host, 8));
  if (n_edges > 0) NK_CUDA_OK(cudaMemcpyAsync(c->edges, edges_host, (size_t)n_edges * 8, cudaMemcpyHostToDevice, c->stream));
  c->ising.edges = c->edges;
  c->ising.n_edges = n_edges;
  c->ising.h = h;
  c->ising.J = J;
  int rc = random_state(c->stream, c->sigma, n_chains, N, -1, seed, chain_offset);
  if (rc) return rc;
  NK_CUDA_OK(cudaStreamSynchronize(c->stream));
  *out = c;
  return NK_OK;
}

void nk_ctx_destroy(nk_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  cudaFree(c->W);
  cudaFree(c->b);
  cudaFree(c->a);
  cudaFree(c->sigma);
  cudaFree(c->log_prob);
  cudaFree(c->n_accepted);
  cudaFree(c->edges);
  cudaFree(c->eloc);
  cudaFree(c->partials);
  cudaFree(c->nacc_sum);
  cudaFree(c->workspace);
  cudaFreeHost(c->partials_host);
  cudaFreeHost(c->nacc_host);
  cudaStreamDestroy(c->stream);
  delete c;
}

int nk_ctx_step_host(nk_ctx *c, const void *W_host, const void *b_host, const void *a_host, int32_t n_discard, void *eloc_host,
                     double *stats_host) {
  NK_CHECK_ARG(c && W_host && eloc_host && stats_host, "nk_ctx_step_host: NULL argument");
  NK_CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  NK_CUDA_OK(cudaMemcpyAsync(c->W, W_host, (size_t)c->N * c->M * c->esz, cudaMemcpyHostToDevice, st));
  if (b_host) NK_CUDA_OK(cudaMemcpyAsync(c->b, b_host, (size_t)c->M * c->esz, cudaMemcpyHostToDevice, st));
  if (a_host) NK_CUDA_OK(cudaMemcpyAsync(c->a, a_host, (size_t)c->N * c->esz, cudaMemcpyHostToDevice, st));
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
  a.rule = NK_RULE_LOCAL;
  a.chain_length = c->chain_length;
  a.n_discard = n_discard;
  a.sweep_size = c->N;
  a.machine_pow = 2.0;
  a.ising = &c->ising;
  a.eloc_out = c->eloc;
  a.eloc_dtype = c->dtype;
  a.path = NK_PATH_AUTO;
  int rc = nk_sweep(st, &rbm, &ch, &a);
  if (rc) return rc;
  c->t = ch.t;
  const int64_t L = c->chain_length;
  NK_CUDA_OK(cudaMemcpyAsync(eloc_host, c->eloc, (size_t)c->B * L * c->esz, cudaMemcpyDeviceToHost, st));
  // statistics: phase 0 -> mean, phase 1 -> shifted moments (single device: no all-reduce)
  rc = stats_partial(st, c->eloc, c->dtype, c->B, L, 0, 0.0, c->partials);
  if (rc) return rc;
  NK_CUDA_OK(cudaMemcpyAsync(c->partials_host, c->partials, sizeof(double) * NK_STATS_NPARTIAL, cudaMemcpyDeviceToHost, st));
  NK_CUDA_OK(cudaStreamSynchronize(st));
  const double mean = c->partials_host[0] / ((double)c->B * (double)L);
  rc = stats_partial(st, c->eloc, c->dtype, c->B, L, 1, mean, c->partials);
  if (rc) return rc;
  NK_CUDA_OK(cudaMemsetAsync(c->nacc_sum, 0, 8, st));
  acc_sum_kernel<<<64, 256, 0, st>>>(c->n_accepted, c->B, (unsigned long long *)c->nacc_sum);
  NK_LAUNCH_OK();
  NK_CUDA_OK(cudaMemcpyAsync(c->partials_host, c->partials, sizeof(double) * NK_STATS_NPARTIAL, cudaMemcpyDeviceToHost, st));
  NK_CUDA_OK(cudaMemcpyAsync(c->nacc_host, c->nacc_sum, 8, cudaMemcpyDeviceToHost, st));
  NK_CUDA_OK(cudaStreamSynchronize(st));
  stats_finalize(c->partials_host, mean, c->B, L, stats_host);
  const double n_steps = (double)c->B * (double)(n_discard + c->chain_length) * (double)c->N;
  stats_host[5] = (double)(*c->nacc_host) / n_steps;
  return NK_OK;
}

int nk_ctx_get_sigma_host(nk_ctx *c, int8_t *sigma_host) {
  NK_CHECK_ARG(c && sigma_host, "nk_ctx_get_sigma_host: NULL argument");
  NK_CUDA_OK(cudaSetDevice(c->device));
  NK_CUDA_OK(cudaMemcpyAsync(sigma_host, c->sigma, (size_t)c->B * c->N, cudaMemcpyDeviceToHost, c->stream));
  NK_CUDA_OK(cudaStreamSynchronize(c->stream));
  return NK_OK;
}

}  // extern "C"
