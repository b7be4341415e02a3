// Internal launcher declarations shared between the translation units of libnkb200.
#pragma once

#include "common.cuh"

namespace nk {

struct SweepKernelArgs {
  nk_rbm_t rbm;
  // chains
  int8_t *sigma;
  void *log_prob;
  int64_t *n_accepted;
  int64_t B;
  uint64_t seed, t0, chain_offset;
  // sweep
  int32_t rule, chain_length, n_discard, sweep_size;
  double machine_pow;
  int8_t *samples_out;
  void *logp_out;
  const uint32_t *stream_w0;
  const void *stream_u;
  const int32_t *clusters;
  int32_t n_clusters;
  // fused E_loc
  int32_t eloc_kind;  // 0 none, 1 ising, 2 localop
  nk_ising_t ising;
  nk_localop_t localop;
  void *eloc_out;
  int32_t eloc_dtype;
  int32_t n_pad;  // per-warp sigma stride in smem
  const int *run_if_flag;  // generic kernel only: run iff NULL or *run_if_flag != 0 (fast-path hand-over)
  void *tanh_out;          // optional [B, chain_length, M]: tanh(theta) of every recorded sample
  int32_t eloc_only;       // sweep_prod only: no proposals, eloc_out[chain] = E_loc(sigma[chain]) (stand-alone local estimator)
  // optional in-kernel statistics of the fused local energies (nk_sweep_t.stats_out): NK_STATS_NPARTIAL shifted sums of
  // this launch's [B, chain_length] values, laid out like the phase-1 output of nk_stats_partial with mu = stats_shift
  double *stats_out;
  double stats_shift;
  const double *cluster_probs;  // ExchangeRule(probabilities=): [n_clusters] weights, or NULL (uniform)
  int32_t no_handover;          // NK_SWEEP_NO_HANDOVER: a kernel that gives up writes NaN into stats_out[0]
};

// sweep_generic.cu — theta-form path (any shape / dtype / rule)
int sweep_generic(cudaStream_t stream, const SweepKernelArgs &a);
// sweep_fast.cu — product-form path (fp32 LocalRule, tanh table resident in shared memory)
bool sweep_fast_supported(const SweepKernelArgs &a);
int sweep_fast(cudaStream_t stream, const SweepKernelArgs &a, const float *theta_ws, int *flags);
// sweep_prod.cu — general product-form path (fp32 / fp64, LocalRule / ExchangeRule, Ising / LocalOperator E_loc)
bool sweep_prod_supported(const SweepKernelArgs &a);
size_t sweep_prod_workspace_bytes(const nk_rbm_t &rbm);
// run_if: launch guard (NULL: always run); flags[giveup] is raised when the weights are outside the product form's range
// stats_guard (out, optional): non-NULL afterwards iff the kernel launched first reduces the statistics of its energies itself
// (SweepKernelArgs.stats_out); *stats_guard != 0 on the device then says that it handed over and K6 must run
int sweep_prod(cudaStream_t stream, const SweepKernelArgs &a, const void *theta_ws, int *flags, void *tables_ws, const int *run_if = nullptr,
               int giveup = 0, const int **stats_guard = nullptr);

}  // namespace nk
