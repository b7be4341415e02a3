/* The drop-in boundary from plain C: a VMC inner-loop step (Metropolis sweeps of an RBM fused with the transverse-field-Ising
 * local energy and the MC statistics) through include/nkb200.h with host buffers only - no Python, no torch, no CUDA headers.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/c_abi_step.c -Lnetket_b200/lib -lnkb200 -Wl,-rpath,$PWD/netket_b200/lib -lm -o c_abi_step
 *   ./c_abi_step [L] [alpha] [chains] [steps]        (default: 6x6 lattice, alpha = 2, 4096 chains x 8 sweeps, 3 steps)
 *
 * Replaces, on the reference side, `vs.parameters = ...; vs.reset(); vs.expect(H)` (netket/vqs/mc/mc_state/state.py:514-576,
 * 695-712).  Exit code 0 iff every step returned finite statistics.  tests/test_gpu_c_abi_example.py builds and runs it. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "nkb200.h"

static uint64_t lcg_state = 0x9E3779B97F4A7C15ull;
static double uniform01(void) {
  lcg_state = lcg_state * 6364136223846793005ull + 1442695040888963407ull;
  return (double)(lcg_state >> 11) * (1.0 / 9007199254740992.0);
}
static float gauss(void) { /* Box-Muller */
  const double u1 = uniform01() + 1e-300, u2 = uniform01();
  return (float)(sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2));
}

int main(int argc, char **argv) {
  const int L = argc > 1 ? atoi(argv[1]) : 6, alpha = argc > 2 ? atoi(argv[2]) : 2;
  const long chains = argc > 3 ? atol(argv[3]) : 4096;
  const int steps = argc > 4 ? atoi(argv[4]) : 3, chain_length = 8;
  const int N = L * L, M = alpha * N;
  /* periodic square lattice: bonds to the right and down (L > 2) */
  const int n_edges = 2 * N;
  int32_t *edges = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)n_edges);
  int e = 0, x, y, i;
  for (y = 0; y < L; ++y)
    for (x = 0; x < L; ++x) {
      edges[2 * e] = y * L + x;
      edges[2 * e + 1] = y * L + (x + 1) % L;
      ++e;
      edges[2 * e] = y * L + x;
      edges[2 * e + 1] = ((y + 1) % L) * L + x;
      ++e;
    }
  float *W = (float *)malloc(sizeof(float) * (size_t)N * M), *b = (float *)malloc(sizeof(float) * M), *a = (float *)malloc(sizeof(float) * N);
  for (i = 0; i < N * M; ++i) W[i] = 0.01f * gauss();
  for (i = 0; i < M; ++i) b[i] = 0.01f * gauss();
  for (i = 0; i < N; ++i) a[i] = 0.01f * gauss();

  nk_ising_t op;
  op.edges = edges;
  op.n_edges = n_edges;
  op.reserved = 0;
  op.h = 3.0;
  op.J = 1.0;
  nk_ctx_desc_t d;
  {
    unsigned char *z = (unsigned char *)&d;
    size_t k;
    for (k = 0; k < sizeof(d); ++k) z[k] = 0;
  }
  d.device = 0;
  d.N = N;
  d.M = M;
  d.dtype = NK_F32;
  d.n_chains = chains;
  d.chain_length = chain_length;
  d.sweep_size = 0; /* = N */
  d.rule = NK_RULE_LOCAL;
  d.machine_pow = 2.0;
  d.n_down = -1;
  d.return_samples = 0;
  d.ising_host = &op;
  d.seed = 1234;
  d.chain_offset = 0;
  d.stream = NULL; /* a stream of the context's own */
  d.eloc_in_param_dtype = 1;
  nk_ctx *ctx = NULL;
  if (nk_ctx_create2(&ctx, &d) != NK_OK) {
    fprintf(stderr, "nk_ctx_create2: %s\n", nk_last_error());
    return 2;
  }
  float *eloc = (float *)malloc(sizeof(float) * (size_t)chains * chain_length);
  double stats[6];
  int s, ok = 1;
  for (s = 0; s < steps; ++s) {
    if (nk_ctx_step_host(ctx, W, b, a, s == 0 ? 5 : 0, eloc, stats) != NK_OK) {
      fprintf(stderr, "nk_ctx_step_host: %s\n", nk_last_error());
      return 3;
    }
    double m = 0.0;
    long k;
    for (k = 0; k < chains * chain_length; ++k) m += eloc[k];
    m /= (double)(chains * chain_length);
    printf("step %d: E = %.6f +- %.6f  var = %.4f  tau = %.4f  R_hat = %.4f  acceptance = %.4f  (mean of the E_loc copied back: %.6f)\n", s,
           stats[0], stats[1], stats[2], stats[3], stats[4], stats[5], m);
    if (!(isfinite(stats[0]) && isfinite(stats[1]) && stats[5] > 0.0 && stats[5] <= 1.0 && fabs(m - stats[0]) <= 1e-4 * fabs(stats[0])))
      ok = 0;
  }
  nk_ctx_destroy(ctx);
  free(eloc);
  free(W);
  free(b);
  free(a);
  free(edges);
  return ok ? 0 : 1;
}
