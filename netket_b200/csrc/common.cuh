// Shared device/host helpers for the nkb200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "nkb200.h"

namespace nk {

// ------------------------------------------------------------------ errors (thread-local text)
void set_error(const char *fmt, ...);

#define NK_CHECK_ARG(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      nk::set_error(__VA_ARGS__);      \
      return NK_EINVAL;                \
    }                                  \
  } while (0)

#define NK_CUDA_OK(expr)                                                                   \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      nk::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return NK_ECUDA;                                                                     \
    }                                                                                      \
  } while (0)

#define NK_LAUNCH_OK()                                                                     \
  do {                                                                                     \
    nk::count_launch();                                                                    \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      nk::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return NK_ECUDA;                                                                     \
    }                                                                                      \
  } while (0)

int num_sms();  // SM count of the current device (cached per device)
void count_launch();  // statistics only: number of kernels this library has launched (nk_launch_count)

// ------------------------------------------------------------------ math in the working precision
template <typename T>
struct Math;

template <>
struct Math<float> {
  static __device__ __forceinline__ float exp(float x) { return expf(x); }
  static __device__ __forceinline__ float log(float x) { return logf(x); }
  static __device__ __forceinline__ float log1p(float x) { return log1pf(x); }
  static __device__ __forceinline__ float abs(float x) { return fabsf(x); }
  static __device__ __forceinline__ float fma(float a, float b, float c) { return fmaf(a, b, c); }
  static __device__ __forceinline__ float ln2() { return 0.69314718055994530942f; }
};
template <>
struct Math<double> {
  static __device__ __forceinline__ double exp(double x) { return ::exp(x); }
  static __device__ __forceinline__ double log(double x) { return ::log(x); }
  static __device__ __forceinline__ double log1p(double x) { return ::log1p(x); }
  static __device__ __forceinline__ double abs(double x) { return fabs(x); }
  static __device__ __forceinline__ double fma(double a, double b, double c) { return ::fma(a, b, c); }
  static __device__ __forceinline__ double ln2() { return 0.69314718055994530942; }
};

// log cosh(x) = |x| + log1p(exp(-2|x|)) - ln 2      (netket/nn/activation.py:78-84)
template <typename T>
__device__ __forceinline__ T lncosh(T x) {
  T ax = Math<T>::abs(x);
  return ax + Math<T>::log1p(Math<T>::exp(T(-2) * ax)) - Math<T>::ln2();
}

// log cosh(y) - log cosh(x), evaluated without subtracting the two ln2 / log1p tails separately
template <typename T>
__device__ __forceinline__ T lncosh_diff(T y, T x) {
  T ay = Math<T>::abs(y), ax = Math<T>::abs(x);
  T ey = Math<T>::exp(T(-2) * ay), ex = Math<T>::exp(T(-2) * ax);
  // log((1+ey)/(1+ex)) = log1p((ey-ex)/(1+ex))
  return (ay - ax) + Math<T>::log1p((ey - ex) / (T(1) + ex));
}

// ------------------------------------------------------------------ warp reductions (all lanes get the result)
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

// ------------------------------------------------------------------ Philox4x32-10 (must match oracle/rng.py)
struct Philox {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  static __host__ __device__ __forceinline__ void mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo) {
    uint64_t p = (uint64_t)a * (uint64_t)b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
  }
  static __host__ __device__ __forceinline__ uint4 run(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0, lo0, hi1, lo1;
      mulhilo(M0, c.x, hi0, lo0);
      mulhilo(M1, c.z, hi1, lo1);
      c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
      if (r != 9) {
        k.x += W0;
        k.y += W1;
      }
    }
    return c;
  }
};

constexpr uint32_t STREAM_STEP = 0u;
constexpr uint32_t STREAM_INIT = 1u;

__host__ __device__ __forceinline__ uint4 philox_words(uint64_t seed, uint64_t t, uint64_t chain, uint32_t stream) {
  uint4 c = make_uint4((uint32_t)t, (uint32_t)(t >> 32), (uint32_t)chain,
                       ((uint32_t)(chain >> 32) & 0x00FFFFFFu) | (stream << 24));
  uint2 k = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  return Philox::run(c, k);
}

template <typename T>
__device__ __forceinline__ T uniform_from_words(uint4 w);
template <>
__device__ __forceinline__ float uniform_from_words<float>(uint4 w) {
  return (float)(w.y >> 8) * 5.9604644775390625e-08f;  // 2^-24
}
template <>
__device__ __forceinline__ double uniform_from_words<double>(uint4 w) {
  return ((double)(w.y >> 5) * 67108864.0 + (double)(w.z >> 6)) * 1.1102230246251565e-16;  // 2^-53
}

template <typename T>
struct DType;
template <>
struct DType<float> {
  static constexpr int code = NK_F32;
};
template <>
struct DType<double> {
  static constexpr int code = NK_F64;
};

// store a value computed in T into an output array of dtype `code`
template <typename T>
__device__ __forceinline__ void store_as(void *out, int64_t idx, T v, int code) {
  if (code == NK_F64)
    reinterpret_cast<double *>(out)[idx] = (double)v;
  else
    reinterpret_cast<float *>(out)[idx] = (float)v;
}

}  // namespace nk
