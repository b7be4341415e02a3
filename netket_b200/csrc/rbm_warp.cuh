// Warp-cooperative RBM primitives, "generic" (theta-form) path.
//
// One warp owns one configuration sigma (N int8 in shared memory) and its hidden pre-activations
// theta[M] (shared memory).  Lanes stride over hidden units j; W rows are read through L1/L2
// (coalesced: row W[i, :] is contiguous in Flax's (in, out) layout).
//
// Closed forms (SURVEY.md §8a "closed forms implied by (a)"):
//   single flip of site i:   delta_i = -2 a_i s_i + sum_j [lncosh(theta_j - 2 s_i W_ij) - lncosh(theta_j)]
//   generic 2-site change:   delta   = a_i d_i + a_k d_k + sum_j [lncosh(theta_j + d_i W_ij + d_k W_kj) - lncosh(theta_j)]
// which replace the reference's full forward pass per proposal (netket/sampler/metropolis.py:441)
// and per connected configuration (netket/vqs/mc/kernels.py:62-71).
#pragma once

#include "common.cuh"

namespace nk {

template <typename T>
struct RbmView {
  const T *__restrict__ W;
  const T *__restrict__ b;
  const T *__restrict__ a;
  int N, M;
};

template <typename T>
__device__ __forceinline__ RbmView<T> make_view(const nk_rbm_t &r) {
  RbmView<T> v;
  v.W = reinterpret_cast<const T *>(r.W);
  v.b = reinterpret_cast<const T *>(r.b);
  v.a = reinterpret_cast<const T *>(r.a);
  v.N = r.N;
  v.M = r.M;
  return v;
}

// theta_j = b_j + sum_i sigma_i W_ij ; returns logpsi = sum_j lncosh(theta_j) + sum_i a_i sigma_i (all lanes).
// STORE=false: theta is not written (pure RBM.apply).
template <typename T, bool STORE>
__device__ __forceinline__ T warp_theta_init(const RbmView<T> &r, const int8_t *sig, T *theta, int lane) {
  T lc = T(0);
  for (int j0 = 0; j0 < r.M; j0 += 128) {
    // 4 independent accumulators per lane for ILP on the W loads
    T acc[4];
    int j[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      j[q] = j0 + q * 32 + lane;
      acc[q] = (r.b != nullptr && j[q] < r.M) ? r.b[j[q]] : T(0);
    }
    for (int i = 0; i < r.N; ++i) {
      T s = (T)sig[i];
      const T *row = r.W + (size_t)i * r.M;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (j[q] < r.M) acc[q] = Math<T>::fma(s, __ldg(row + j[q]), acc[q]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (j[q] < r.M) {
        if (STORE) theta[j[q]] = acc[q];
        lc += lncosh(acc[q]);
      }
  }
  if (r.a != nullptr)
    for (int i = lane; i < r.N; i += 32) lc = Math<T>::fma((T)sig[i], r.a[i], lc);
  return warp_sum(lc);
}

// sum_j [lncosh(theta_j + d * W_ij) - lncosh(theta_j)]   (partial over this lane's j; caller reduces)
template <typename T>
__device__ __forceinline__ T lane_delta_one(const RbmView<T> &r, const T *theta, int i, T d, int lane) {
  const T *row = r.W + (size_t)i * r.M;
  T part = T(0);
  for (int j = lane; j < r.M; j += 32) {
    T th = theta[j];
    part += lncosh_diff(Math<T>::fma(d, __ldg(row + j), th), th);
  }
  return part;
}

template <typename T>
__device__ __forceinline__ T lane_delta_two(const RbmView<T> &r, const T *theta, int i, T di, int k, T dk, int lane) {
  const T *rowi = r.W + (size_t)i * r.M;
  const T *rowk = r.W + (size_t)k * r.M;
  T part = T(0);
  for (int j = lane; j < r.M; j += 32) {
    T th = theta[j];
    T y = Math<T>::fma(di, __ldg(rowi + j), Math<T>::fma(dk, __ldg(rowk + j), th));
    part += lncosh_diff(y, th);
  }
  return part;
}

template <typename T>
__device__ __forceinline__ void warp_theta_update_one(const RbmView<T> &r, T *theta, int i, T d, int lane) {
  const T *row = r.W + (size_t)i * r.M;
  for (int j = lane; j < r.M; j += 32) theta[j] = Math<T>::fma(d, __ldg(row + j), theta[j]);
}

template <typename T>
__device__ __forceinline__ void warp_theta_update_two(const RbmView<T> &r, T *theta, int i, T di, int k, T dk, int lane) {
  const T *rowi = r.W + (size_t)i * r.M;
  const T *rowk = r.W + (size_t)k * r.M;
  for (int j = lane; j < r.M; j += 32)
    theta[j] = Math<T>::fma(di, __ldg(rowi + j), Math<T>::fma(dk, __ldg(rowk + j), theta[j]));
}

// E_loc of the transverse-field Ising model for the configuration owned by this warp:
//   E_loc = J sum_<ij> s_i s_j - h sum_i exp(delta_i)
// (slot 0 / slots 1..N of IsingJax.get_conn_padded, netket/operator/_ising/jax.py:125-165, contracted with
//  local_value_kernel_jax, netket/vqs/mc/kernels.py:62-71).  h == 0: only the diagonal slot exists.
template <typename T>
__device__ __forceinline__ T warp_eloc_ising(const RbmView<T> &r, const T *theta, const int8_t *sig, const int32_t *edges,
                                             int n_edges, T h, T J, int lane) {
  int zz = 0;
  for (int e = lane; e < n_edges; e += 32) zz += (int)sig[edges[2 * e]] * (int)sig[edges[2 * e + 1]];
  zz = warp_sum(zz);
  T off = T(0);
  if (h != T(0)) {
    for (int i = 0; i < r.N; ++i) {
      T s = (T)sig[i];
      T d = T(-2) * s;
      T delta = warp_sum(lane_delta_one(r, theta, i, d, lane));
      if (r.a != nullptr) delta = Math<T>::fma(d, r.a[i], delta);
      off += Math<T>::exp(delta);
    }
  }
  return J * (T)zz - h * off;
}

// E_loc for a LocalOperator made of 1- and 2-site terms (packed tables).
// Entries with |mel| <= mel_cutoff are dropped, as the compaction of _local_operator_kernel_jax does
// (netket/operator/_local_operator/jax.py:177-199).
template <typename T>
__device__ __forceinline__ T warp_eloc_localop(const RbmView<T> &r, const T *theta, const int8_t *sig, const nk_localop_t &op,
                                               int lane) {
  double diag = op.constant;
  double acc = 0.0;
  for (int g = 0; g < op.n_groups; ++g) {
    const nk_localop_group_t &G = op.groups[g];
    const int rows = 1 << G.n_sites;
    for (int o = 0; o < G.n_ops; ++o) {
      int s0 = G.acting_on[o * G.n_sites];
      int s1 = G.n_sites == 2 ? G.acting_on[o * G.n_sites + 1] : s0;
      int x0 = sig[s0] > 0 ? 0 : 1;
      int x1 = sig[s1] > 0 ? 0 : 1;
      int row = G.n_sites == 2 ? (2 * x0 + x1) : x0;  // _state_to_number: first site most significant
      diag += G.diag_mels[o * rows + row];
      int nc = G.n_conns[o * rows + row];
      for (int c = 0; c < nc; ++c) {
        double mel = G.mels[((size_t)o * rows + row) * G.ncmax + c];
        if (!(fabs(mel) > op.mel_cutoff)) continue;
        const int8_t *xp = G.x_prime + (((size_t)o * rows + row) * G.ncmax + c) * G.n_sites;
        T d0 = (T)(2 * (x0 - (int)xp[0]));  // sigma' - sigma with sigma = 1 - 2 x
        T d1 = G.n_sites == 2 ? (T)(2 * (x1 - (int)xp[1])) : T(0);
        T part = (d1 == T(0)) ? lane_delta_one(r, theta, s0, d0, lane)
                              : ((d0 == T(0)) ? lane_delta_one(r, theta, s1, d1, lane)
                                              : lane_delta_two(r, theta, s0, d0, s1, d1, lane));
        T delta = warp_sum(part);
        if (r.a != nullptr) delta += d0 * r.a[s0] + d1 * r.a[s1];
        acc += mel * (double)Math<T>::exp(delta);
      }
    }
  }
  if (op.nonzero_diagonal && fabs(diag) > op.mel_cutoff) acc += diag;
  return (T)acc;
}

}  // namespace nk
