// hilbert.random_state for Spin-1/2 (netket/hilbert/random/homogeneous.py:35-72, random/fock.py:77-97),
// driven by the Philox STREAM_INIT stream defined in oracle/rng.py (must stay bit-identical to
// oracle/hilbert.py:random_state).
#include "kernels.cuh"

namespace nk {

__global__ void __launch_bounds__(128) random_state_kernel(int8_t *__restrict__ sigma, int64_t B, int N, int n_down, uint64_t seed,
                                                           uint64_t chain_offset) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= B) return;
  int8_t *row = sigma + c * N;
  const uint64_t gc = chain_offset + (uint64_t)c;
  if (n_down < 0) {
    // bit i of the 128-bit block i/128 is the local index of site i (0 -> +1, 1 -> -1)
    for (int blk = 0; blk * 128 < N; ++blk) {
      const uint4 w = philox_words(seed, (uint64_t)blk, gc, STREAM_INIT);
      const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
      for (int i = blk * 128; i < N && i < (blk + 1) * 128; ++i) {
        const uint32_t bit = (ww[(i >> 5) & 3] >> (i & 31)) & 1u;
        row[i] = bit ? (int8_t)-1 : (int8_t)1;
      }
    }
  } else {
    // n_down local-index-1 sites first, then Fisher-Yates with j = (word_i * (i+1)) >> 32
    for (int i = 0; i < N; ++i) row[i] = i < n_down ? (int8_t)-1 : (int8_t)1;
    uint4 w = make_uint4(0, 0, 0, 0);
    int cur_blk = -1;
    for (int i = N - 1; i > 0; --i) {
      if ((i >> 2) != cur_blk) {
        cur_blk = i >> 2;
        w = philox_words(seed, (uint64_t)cur_blk, gc, STREAM_INIT);
      }
      const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
      const int j = (int)__umulhi(ww[i & 3], (uint32_t)(i + 1));
      const int8_t tmp = row[i];
      row[i] = row[j];
      row[j] = tmp;
    }
  }
}

int random_state(cudaStream_t stream, int8_t *sigma, int64_t B, int32_t N, int32_t n_down, uint64_t seed, uint64_t chain_offset) {
  if (B == 0) return NK_OK;
  const int threads = 128;
  random_state_kernel<<<(unsigned)((B + threads - 1) / threads), threads, 0, stream>>>(sigma, B, N, n_down, seed, chain_offset);
  NK_LAUNCH_OK();
  return NK_OK;
}

}  // namespace nk
