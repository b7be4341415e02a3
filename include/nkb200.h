/*
 * nkb200 — C ABI of the B200-native VMC inner loop (Metropolis sweep + local energy for RBM).
 *
 * This header is the drop-in boundary.  The reference (NetKet) has no FFI of its own: its
 * extension points are Python-level.  Each entry point below states the
 * reference function it replaces (paths relative to /root/reference).  INTEGRATION.md shows
 * the reference-side binding (ctypes / jax.ffi) a maintainer would add.
 *
 * Conventions
 *   - plain C, no torch / jax types; every pointer is a DEVICE pointer unless the name ends
 *     in `_host`; the caller owns every buffer; nothing is allocated or freed inside except
 *     by the nk_ctx_* host-buffer API at the bottom of this file;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never
 *     synchronises the device, and is safe to call concurrently from different host threads
 *     on different devices (no mutable globals; the error string is thread-local);
 *   - returns 0 on success, a negative NK_E* code otherwise; nk_last_error() gives the text;
 *   - sigma is int8 with values +1 / -1 (netket/hilbert/spin.py:165-171: local index 0 <-> +1,
 *     1 <-> -1); parameters follow Flax's layout: W (N, M) row-major "in x out", b (M), a (N)
 *     (netket/models/rbm.py:59-79);
 *   - dtype codes: NK_F32 / NK_F64 (parameter and log-amplitude precision).
 */
#ifndef NKB200_H
#define NKB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NK_VERSION 100

#define NK_F32 0
#define NK_F64 1

#define NK_OK 0
#define NK_EINVAL (-1)   /* bad argument (shape, dtype, null pointer) */
#define NK_ECUDA (-2)    /* CUDA runtime error, text in nk_last_error() */
#define NK_EUNSUPPORTED (-3)

#define NK_RULE_LOCAL 0    /* netket/sampler/rules/local.py:23-52 */
#define NK_RULE_EXCHANGE 1 /* netket/sampler/rules/exchange.py:25-187 (nk_sweep_t.cluster_probs: probabilities=) */

/* kernel-path selection for the sweep / E_loc kernels (NK_PATH_AUTO picks the fastest valid one) */
#define NK_PATH_AUTO 0
#define NK_PATH_GENERIC 1 /* theta-form, lncosh differences; any shape, any |W| */
#define NK_PATH_FAST 2    /* product form on exp(-4W) tables; error if no such kernel covers the configuration (weights outside
                             the form's numerical range still hand over, in-stream, to the theta-form kernel) */
#define NK_PATH_PROD 3    /* like NK_PATH_FAST, but always the general product-form kernel (fp32/fp64, both rules) */

typedef struct nk_rbm_t {
  const void *W; /* [N, M] */
  const void *b; /* [M] or NULL   (use_hidden_bias=False) */
  const void *a; /* [N] or NULL   (use_visible_bias=False) */
  int32_t N;     /* visible units = sites */
  int32_t M;     /* hidden units = alpha * N */
  int32_t dtype; /* NK_F32 | NK_F64 */
  int32_t reserved;
} nk_rbm_t;

/* Transverse-field Ising  H = -h sum_i sx_i + J sum_<ij> sz_i sz_j  (netket/operator/_ising/jax.py:35-175) */
typedef struct nk_ising_t {
  const int32_t *edges; /* [E, 2] */
  int32_t n_edges;
  int32_t reserved;
  double h; /* h == 0 => StaticZero: K = 1 (jax.py:60-61,127-131,156-157) */
  double J;
} nk_ising_t;

/* One group of LocalOperator terms acting on the same number of sites (1 or 2), i.e. the packed
 * lookup tables of netket/operator/_local_operator/compile_helpers.py:29-218. */
typedef struct nk_localop_group_t {
  int32_t n_ops;
  int32_t n_sites; /* 1 or 2 */
  int32_t ncmax;   /* padded number of off-diagonal entries per row */
  int32_t reserved;
  const int32_t *acting_on; /* [n_ops, n_sites] */
  const double *diag_mels;  /* [n_ops, 2^n_sites] */
  const int32_t *n_conns;   /* [n_ops, 2^n_sites] */
  const double *mels;       /* [n_ops, 2^n_sites, ncmax]  (padding may be NaN, never read) */
  const int8_t *x_prime;    /* [n_ops, 2^n_sites, ncmax, n_sites] local indices 0/1 */
} nk_localop_group_t;

typedef struct nk_localop_t {
  nk_localop_group_t groups[2];
  int32_t n_groups;
  int32_t nonzero_diagonal;
  int32_t max_conn_size; /* K */
  int32_t reserved;
  double constant;
  double mel_cutoff; /* 1e-10 (netket/operator/_local_operator/base.py:90) */
} nk_localop_t;

/* Markov-chain state = MetropolisSamplerState (netket/sampler/metropolis.py:42-137). */
typedef struct nk_chains_t {
  int8_t *sigma;         /* [B, N] in/out */
  void *log_prob;        /* [B] dtype; out: machine_pow * logpsi(sigma) after the call */
  int64_t *n_accepted;   /* [B] in/out, incremented (n_accepted_proc) */
  void *workspace;       /* nk_sweep_workspace_bytes() bytes of scratch (theta); contents are derived state */
  int64_t B;             /* chains on this device */
  uint64_t seed;         /* Philox key */
  uint64_t t;            /* Metropolis steps already performed per chain (Philox counter) */
  uint64_t chain_offset; /* global index of chain 0 of this device (multi-GPU sharding) */
} nk_chains_t;

/* nk_sweep_t.flags */
#define NK_SWEEP_NO_HANDOVER 1 /* enqueue ONLY the tuned kernel NK_PATH_AUTO would try first (fewer launches per step).  If the
                                  weights turn out to be outside that kernel's numerical range nothing is sampled and, with
                                  stats_out given, stats_out[0] is NaN: the caller, who reads the sums anyway, repeats the call
                                  without this flag (the chain state was not touched).  Needs stats_out. */

typedef struct nk_sweep_t {
  int32_t rule;         /* NK_RULE_* */
  int32_t chain_length; /* recorded sweeps */
  int32_t n_discard;    /* burn-in sweeps run before, not recorded (mc_state/state.py:559-568) */
  int32_t sweep_size;   /* MH steps per sweep, default N (metropolis.py:283-284) */
  double machine_pow;   /* default 2 (sampler/base.py:139-150) */
  int8_t *samples_out;  /* [B, chain_length, N] or NULL */
  void *logp_out;       /* [B, chain_length] dtype or NULL (return_log_probabilities) */
  /* optional explicit proposal stream (fixed-proposal-stream mode), indexed [step, chain] with
   * step in [0, (n_discard+chain_length)*sweep_size); NULL => in-kernel Philox4x32-10 */
  const uint32_t *stream_w0;
  const void *stream_u; /* dtype */
  /* ExchangeRule */
  const int32_t *clusters; /* [C, 2] */
  int32_t n_clusters;
  int32_t path; /* NK_PATH_* */
  /* optional fused local energy of every recorded sample (sigma' never materialised) */
  const nk_ising_t *ising;     /* or NULL */
  const nk_localop_t *localop; /* or NULL */
  void *eloc_out;              /* [B, chain_length] */
  int32_t eloc_dtype;          /* NK_F32 | NK_F64 = promote(operator dtype, rbm dtype) */
  int32_t flags;               /* NK_SWEEP_* bits */
  /* optional: tanh(theta) of every recorded sample, [B, chain_length, M] in the rbm dtype.  The sweep kernels hold it in
   * registers anyway ((A - B) / (A + B)); nk_forces_rbm takes it instead of recomputing theta for the whole batch */
  void *tanh_out;
  /* optional: MC statistics of the fused local energies, reduced inside the sweep kernel (netket/stats/mc_stats_old.py:87-196).
   * stats_out: NK_STATS_NPARTIAL doubles (device, zeroed by the call) that receive the phase-1 sums of nk_stats_partial
   * over this launch's eloc_out [B, chain_length] with shift = stats_shift (any estimate of the mean: the previous step's
   * energy; the sums are exact in real arithmetic for every shift, accumulated in double).  The caller all-reduces them over
   * GPUs and calls nk_stats_finalize(sums, stats_shift, ...); nk_stats_finalize returns mean = stats_shift + sums[7] / n. */
  double *stats_out;
  double stats_shift;
  /* ExchangeRule(probabilities=...) (netket/sampler/rules/exchange.py:86-123,155-160,177-182): relative weight of every
   * cluster, [n_clusters] doubles > 0 (device), or NULL for the uniform rule.  The cluster is drawn with weight
   * hoppable * p by inverse CDF in cluster order (jax.random.choice's algorithm) at r = (w0 + 1/2) / 2^32, and the
   * log-ratio correction is log sum_c w_c(sigma) - log sum_c w_c(sigma'). */
  const double *cluster_probs;
} nk_sweep_t;

const char *nk_last_error(void);
int nk_version(void);
/* number of CUDA kernels this library has launched in this process (a statistics counter; bench.py's gpu_launches) */
long long nk_launch_count(void);

/* RBM.apply: logpsi[B] = sum_j lncosh(sum_i sigma_i W_ij + b_j) + sum_i a_i sigma_i.
 * Replaces netket/models/rbm.py:57-81 + netket/nn/activation.py:78-84 (seam S5).
 * theta_out (optional, [B, M]) receives the hidden pre-activations. */
int nk_rbm_logpsi(void *stream, const nk_rbm_t *rbm, const int8_t *sigma, int64_t B, void *logpsi_out, void *theta_out);

/* theta[B, M] = sigma W + b as a tensor-core GEMM (tcgen05 for fp32 via exact 3-way bf16 split, DMMA for
 * fp64).  Replaces the nn.Dense half of netket/models/rbm.py:59-67 at `_reset`
 * (netket/sampler/metropolis.py:399-403).  workspace: nk_theta_gemm_workspace_bytes(). */
int64_t nk_theta_gemm_workspace_bytes(const nk_rbm_t *rbm, int64_t B);
int nk_theta_gemm(void *stream, const nk_rbm_t *rbm, const int8_t *sigma, int64_t B, void *theta_out, void *workspace);

/* hilbert.random_state for Spin-1/2 (netket/hilbert/random/homogeneous.py:35-72, random/fock.py:77-97).
 * n_down < 0: unconstrained; otherwise exactly n_down sites are -1 (total_sz constraint). */
int nk_random_state(void *stream, int8_t *sigma, int64_t B, int32_t N, int32_t n_down, uint64_t seed, uint64_t chain_offset);

/* MetropolisSampler._reset + _sample_chain (netket/sampler/metropolis.py:382-505) for LocalRule /
 * ExchangeRule on an RBM, optionally fused with local_value_kernel_jax (netket/vqs/mc/kernels.py:62-71). */
int64_t nk_sweep_workspace_bytes(const nk_rbm_t *rbm, int64_t B);
int nk_sweep(void *stream, const nk_rbm_t *rbm, nk_chains_t *chains, const nk_sweep_t *args);

/* IsingJax.get_conn_padded (netket/operator/_ising/jax.py:82-88,125-165).
 * xp_out [B, K, N] int8, mels_out [B, K] (mel_dtype), K = N+1 (1 if h == 0). */
int nk_ising_conn(void *stream, const nk_ising_t *op, const int8_t *x, int64_t B, int32_t N, int8_t *xp_out, void *mels_out,
                  int32_t mel_dtype);
/* IsingJax.n_conn (jax.py:71-80,168-175) */
int nk_ising_n_conn(void *stream, const nk_ising_t *op, const int8_t *x, int64_t B, int32_t N, int32_t *nconn_out);

/* LocalOperatorJax._get_conn_padded (netket/operator/_local_operator/jax.py:74-201,256-284):
 * xp_out [B, K, N] int8, mels_out [B, K], nconn_out [B] (may be NULL). */
int nk_localop_conn(void *stream, const nk_localop_t *op, const int8_t *x, int64_t B, int32_t N, int8_t *xp_out, void *mels_out,
                    int32_t mel_dtype, int32_t *nconn_out);

/* local_value_kernel_jax for (RBM, Ising) and (RBM, LocalOperator) without materialising sigma'
 * (netket/vqs/mc/kernels.py:62-71; seam S4).  eloc_out [B] (eloc_dtype).
 * workspace: nk_sweep_workspace_bytes(rbm, B) bytes of scratch for the product-form kernel, or NULL (theta-form kernel). */
int nk_eloc_ising_rbm(void *stream, const nk_rbm_t *rbm, const nk_ising_t *op, const int8_t *sigma, int64_t B, void *eloc_out,
                      int32_t eloc_dtype, int32_t path, void *workspace);
int nk_eloc_localop_rbm(void *stream, const nk_rbm_t *rbm, const nk_localop_t *op, const int8_t *sigma, int64_t B, void *eloc_out,
                        int32_t eloc_dtype, int32_t path, void *workspace);

/* statistics() (netket/stats/mc_stats_old.py:52-196), split so that only scalars cross devices:
 *   phase 0: partials_out[0] = sum(x)                                    -> all-reduce -> mean
 *   phase 1: partials_out[0..6] = shifted second moments around `shift`  -> all-reduce
 * then nk_stats_finalize on the host.  data [n_chains, L] (dtype).  partials_out: 8 doubles (device). */
#define NK_STATS_NPARTIAL 8
int nk_stats_partial(void *stream, const void *data, int32_t dtype, int64_t n_chains, int64_t L, int32_t phase, double shift,
                     double *partials_out);
/* host-only arithmetic: sums_host = all-reduced phase-1 partials, shift = the `shift` they were taken around (the mean of
 * phase 0, or any estimate for the one-pass reduction of nk_sweep_t.stats_out): mean = shift + sums[7] / n.
 * out_host = {mean, error_of_mean, variance, tau_corr, R_hat} */
int nk_stats_finalize(const double *sums_host, double shift, int64_t n_chains_total, int64_t L, double *out_host);

/* Integrated autocorrelation time of every chain with Sokal's automatic window: the tau_corr / tau_corr_max of the opt-in FFT
 * variant of `statistics` (netket/stats/mc_stats.py:303-331, netket/stats/_autocorr.py:40-86; the autocorrelation function the
 * reference obtains by a zero-padded FFT is summed directly, lag by lag, up to the window).  data [n_chains, L] (dtype), c: the
 * window constant (5).  out (3 doubles, device, zeroed by the call): out[0] = sum of the chains' tau, out[1] = their maximum in an
 * order-preserving integer encoding (combine across GPUs with an integer / bitwise-ordered MAX, decode with
 * nk_stats_tau_max_decode on the host), out[2] = number of chains whose tau is NaN (zero variance). */
int nk_stats_tau(void *stream, const void *data, int32_t dtype, int64_t n_chains, int64_t L, double c, double *out);
double nk_stats_tau_max_decode(double encoded);

/* Streaming statistics: OnlineStats (netket/_src/stats/online_stats/accumulator.py:31-447), the accumulator behind
 * thermalise_mcmc / check_mc_convergence / expect_to_precision (netket/_src/vqs/check_mc_convergence.py,
 * expect_to_precision.py).  The state is the reference's pytree, field for field, as device arrays of doubles:
 *   chain_count, chain_mean, chain_M2          [n_chains]
 *   cross_sum, m1_sum, m2_sum, pair_count      [n_chains, max_lag + 1]   (absent when max_lag == 0)
 *   chain_buf                                  [n_chains, max_lag]       last samples of each chain, right-aligned
 * buf_len = number of valid samples in chain_buf; the caller advances it: min(buf_len + n, max_lag).
 *   nk_online_stats_update    `_update_arrays` (kernels.py:116-190): merges data [n_chains, n] (dtype) into `in`, writing
 *                             `out` (may be the same struct: in-place).  decay: 1.0 for the reference's None.
 *   nk_online_stats_summary   sums over this device's chains; phase 0 -> 3 doubles {sum count, sum count*mean, sum mean};
 *                             phase 1 (given gmean = [1]/[0] and mbar = [2]/n_chains of the all-reduced phase 0)
 *                             -> NK_ONLINE_NSUM + max_lag + 1 doubles.  Only these sums cross GPUs.
 *   nk_online_stats_finalize  host arithmetic on the all-reduced sums: out_host[NK_ONLINE_NOUT] = {mean, error_of_mean,
 *                             variance, tau_corr, R_hat, tau_corr_batch, tau_corr_acf (Geyer IPS + IMS), acf window
 *                             saturated (0/1), tau_corr reliable (0/1)}; acf_host [max_lag + 1] or NULL (NaN = no acf). */
typedef struct {
  double *chain_count, *chain_mean, *chain_M2;
  double *cross_sum, *m1_sum, *m2_sum, *pair_count;
  double *chain_buf;
  int64_t n_chains;
  int32_t max_lag;
  int32_t buf_len;
} nk_online_stats_t;
#define NK_ONLINE_NSUM 4
#define NK_ONLINE_NOUT 9
#define NK_ONLINE_MAX_LAG 4096
int nk_online_stats_update(void *stream, const nk_online_stats_t *in, const nk_online_stats_t *out, const void *data, int32_t dtype,
                           int64_t n, double decay);
int nk_online_stats_summary(void *stream, const nk_online_stats_t *state, int32_t phase, double gmean, double mbar, double *sums_out);
int nk_online_stats_finalize(const double *phase0_host, const double *phase1_host, int64_t n_chains_total, int64_t n_samples_total,
                             int32_t max_lag, double *out_host, double *acf_host);

/* Forces F_k = < d log psi / d p_k * (E_loc - mean) > of the RBM over a batch of samples: the vjp of
 * netket/vqs/mc/mc_state/expect_forces.py:69-112 (`forces_expect_hermitian`) in closed form
 * (d/dW_ij = sigma_i tanh theta_j, d/db_j = tanh theta_j, d/da_i = sigma_i).
 *   nk_forces_rbm       sums[N*M | M | N] (doubles, device; zeroed by the call) = sum_s dlogpsi(sigma_s) * (eloc_s - mean)
 *                       over this device's Ns samples; only these sums cross GPUs (all-reduce by the caller);
 *   nk_forces_finalize  out[k] = (dtype) (sums[k] * scale), scale = 1 / n_samples_total (x 2 for the gradient of a
 *                       real-parameter ansatz, netket/vqs/mc/common.py:103-118 `force_to_grad`).
 * tanh_theta: [Ns, M] (rbm dtype) as produced by nk_sweep_t.tanh_out, or NULL: theta is then recomputed for the batch
 * (theta GEMM) into workspace: nk_forces_workspace_bytes(rbm, Ns) bytes (may be NULL when tanh_theta is given). */
int64_t nk_forces_workspace_bytes(const nk_rbm_t *rbm, int64_t Ns);
int nk_forces_rbm(void *stream, const nk_rbm_t *rbm, const int8_t *samples, int64_t Ns, const void *eloc, int32_t eloc_dtype,
                  double mean, double *sums, void *workspace, const void *tanh_theta);
int nk_forces_finalize(void *stream, const double *sums, double scale, int64_t n, void *out, int32_t dtype);

/* Quantum geometric tensor, matrix-free (QGTOnTheFly: netket/optimizer/qgt/qgt_onthefly_logic.py:33-43,
 *   S v = O^H ((O v - mean(O v)) / n) + diag_shift v,   O[s, :] = d log psi(sigma_s) / d p  in closed form for the RBM).
 *   nk_rbm_tanh_theta  out[Ns, M] (rbm dtype) = tanh(sigma W + b): what nk_sweep_t.tanh_out records, for samples drawn without it
 *                      (theta GEMM + tanh in place).  workspace: nk_theta_gemm_workspace_bytes(rbm, Ns).
 *   nk_rbm_jvp         y[Ns] (double) = O v.  `v` carries the tangent vector in the layout of the parameters ({V, v_b, v_a};
 *                      b / a may be NULL like the parameters').  y_sum_out (1 double, device, optional) receives sum_s y[s].  scratch: Ns * M elements of v's dtype (receives sigma V + v_b);
 *                      workspace: nk_theta_gemm_workspace_bytes(v, Ns).
 *   O^H w is nk_forces_rbm with `eloc = y`, `mean = mean(y)` (all-reduced), scaled by 1 / n_samples in nk_forces_finalize. */
int nk_rbm_tanh_theta(void *stream, const nk_rbm_t *rbm, const int8_t *samples, int64_t Ns, void *out, void *workspace);
int nk_rbm_jvp(void *stream, const nk_rbm_t *v, const int8_t *samples, int64_t Ns, const void *tanh_theta, double *y_out,
               double *y_sum_out, void *scratch, void *workspace);

/* ---------------------------------------------------------------------------------------------
 * Host-buffer API: one VMC inner-loop step with HOST pointers (what bench.py's `e2e` times).
 * The context owns the device buffers (parameters, chains, operator tables, E_loc, samples).
 * Mirrors   vs.parameters = ...; vs.reset(); vs.expect(H)   of
 * netket/vqs/mc/mc_state/state.py:514-576,695-712 for (MetropolisLocal | MetropolisExchange, RBM, Ising | LocalOperator).
 *
 * A step is two calls, so that the caller's own collective can run between them on multi-GPU jobs:
 *   nk_ctx_step_begin   upload W, b, a; run n_discard + chain_length sweeps fused with E_loc and with the statistics'
 *                       partial sums; asynchronous on nk_ctx_stream(ctx);
 *   (all-reduce nk_ctx_partials_device(ctx), NK_CTX_NPARTIAL doubles, in place, ordered after nk_ctx_stream(ctx):
 *    NCCL / torch.distributed / jax.lax.psum - the library itself never communicates)
 *   nk_ctx_step_end     copy E_loc (and the samples) to the host, ONE synchronisation, the five statistics + acceptance.
 * nk_ctx_step_host = begin + end for a single device.
 * ------------------------------------------------------------------------------------------- */
typedef struct nk_ctx nk_ctx;
#define NK_CTX_NPARTIAL (NK_STATS_NPARTIAL + 2) /* [phase-1 sums | number of chains | sum of the acceptance counters] */
#define NK_RESHIFT 1 /* nk_ctx_step_end (multi-device only): the statistics' shift (previous mean) was too far from this
                        step's mean for full precision; results are valid to ~1e-8, the next step re-centres */
typedef struct nk_ctx_desc_t {
  int32_t device, N, M, dtype;
  int64_t n_chains;      /* chains on this device */
  int32_t chain_length;  /* recorded sweeps per step */
  int32_t sweep_size;    /* 0 => N (metropolis.py:283-284) */
  int32_t rule;          /* NK_RULE_* */
  int32_t n_clusters;
  const int32_t *clusters_host;     /* [n_clusters, 2]   (ExchangeRule) */
  const double *cluster_probs_host; /* [n_clusters] or NULL (ExchangeRule(probabilities=)) */
  double machine_pow;               /* default 2 */
  int32_t n_down;                   /* initial configurations: < 0 unconstrained, else exactly n_down spins down (total_sz) */
  int32_t return_samples;           /* keep the samples [n_chains, chain_length, N] of every step for nk_ctx_step_end */
  const nk_ising_t *ising_host;     /* exactly one of the two operators; every pointer inside is a HOST pointer */
  const nk_localop_t *localop_host;
  uint64_t seed, chain_offset;
  void *stream;                     /* cudaStream_t the context enqueues on, or NULL: a (non-blocking) stream of its own.  NULL is NOT
                                       the legacy default stream: pass cudaStreamLegacy ((cudaStream_t)0x1) to run on that one.  A
                                       collective between step_begin and step_end must be ordered after / before nk_ctx_stream(ctx) */
  int32_t eloc_in_param_dtype;      /* 0: E_loc in promote(operator, parameter) = float64; 1: in the parameter dtype */
  int32_t reserved;
} nk_ctx_desc_t;
int nk_ctx_create2(nk_ctx **out, const nk_ctx_desc_t *desc);
/* (MetropolisLocal, Ising, machine_pow 2, E_loc in the parameter dtype): the v1 signature, kept */
int nk_ctx_create(nk_ctx **out, int32_t device, int32_t N, int32_t M, int32_t dtype, int64_t n_chains, int32_t chain_length,
                  const int32_t *edges_host, int32_t n_edges, double h, double J, uint64_t seed, uint64_t chain_offset);
void nk_ctx_destroy(nk_ctx *ctx);
void *nk_ctx_stream(nk_ctx *ctx);
double *nk_ctx_partials_device(nk_ctx *ctx);
int nk_ctx_step_begin(nk_ctx *ctx, const void *W_host, const void *b_host, const void *a_host, int32_t n_discard);
/* eloc_host [n_chains, chain_length] (float64, or the parameter dtype); samples_host [n_chains, chain_length, N] or NULL;
 * stats_host [6]: mean, error_of_mean, variance, tau_corr, R_hat, acceptance - over ALL devices if the partials were reduced */
int nk_ctx_step_end(nk_ctx *ctx, void *eloc_host, int8_t *samples_host, double *stats_host);
/* upload parameters (pinned or pageable host memory), run n_discard + chain_length sweeps fused with E_loc,
 * copy E_loc [n_chains, chain_length] and the 5 statistics + acceptance back to host; synchronises once. */
int nk_ctx_step_host(nk_ctx *ctx, const void *W_host, const void *b_host, const void *a_host, int32_t n_discard,
                     void *eloc_host, double *stats_host /* [6]: mean, err, var, tau, rhat, acceptance */);
/* copy the current configurations sigma [n_chains, N] to host */
int nk_ctx_get_sigma_host(nk_ctx *ctx, int8_t *sigma_host);

/* Measurement utility (synchronises): on-chip peak rates of the current device, used as roofline denominators.
 * which: 0 shared-memory read GB/s (LDS.128), 1 L2 read GB/s, 2 fp32 FMA GFLOP/s, 3 MUFU Gop/s, 4 fp64 FMA GFLOP/s. */
int nk_microbench(int32_t which, double *result_host);

#ifdef __cplusplus
}
#endif
#endif /* NKB200_H */
