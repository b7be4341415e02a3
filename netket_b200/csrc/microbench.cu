// Micro-benchmarks of the on-chip denominators the sweep / E_loc kernels are measured against
// ("No smem / L2 / SFU / fp64 peak is measured yet - the build must microbenchmark those").
// nk_microbench() is a measurement utility: it synchronises and times with CUDA events on its own stream.
#include "kernels.cuh"

namespace nk {

// ---- shared-memory read bandwidth: conflict-free LDS.128, 32 KB per warp-sweep
__global__ void __launch_bounds__(1024, 1) mb_smem_kernel(int iters, float *sink) {
  extern __shared__ __align__(16) float4 buf4[];
  const int n4 = 8192;  // 128 KB
  for (int i = threadIdx.x; i < n4; i += blockDim.x) buf4[i] = make_float4(i, 1.f, 2.f, 3.f);
  __syncthreads();
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int idx = threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float4 v = buf4[(idx + u * 1024) & (n4 - 1)];
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
    idx = (idx + 37 * 32) & (n4 - 1);
  }
  if (acc.x + acc.y + acc.z + acc.w == 12345.678f) sink[0] = acc.x;
}

// ---- L2 read bandwidth: every CTA streams a 32 MB (L2-resident) buffer with 128-bit loads, L1 bypassed
__global__ void __launch_bounds__(512) mb_l2_kernel(const float4 *__restrict__ buf, size_t n4, int iters, float *sink) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int it = 0; it < iters; ++it) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      const float4 v = __ldcg(buf + i);
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
  }
  if (acc.x + acc.y + acc.z + acc.w == 12345.678f) sink[0] = acc.x;
}

template <typename T>
__global__ void __launch_bounds__(1024, 1) mb_fma_kernel(int iters, T *sink, T seed) {
  T a0 = seed, a1 = seed + T(1), a2 = seed + T(2), a3 = seed + T(3), a4 = seed + T(4), a5 = seed + T(5), a6 = seed + T(6),
    a7 = seed + T(7);
  const T m = T(0.999), c = T(0.001);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = a0 * m + c;
      a1 = a1 * m + c;
      a2 = a2 * m + c;
      a3 = a3 * m + c;
      a4 = a4 * m + c;
      a5 = a5 * m + c;
      a6 = a6 * m + c;
      a7 = a7 * m + c;
    }
  }
  const T s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == T(12345.678)) sink[0] = s;
}

__global__ void __launch_bounds__(1024, 1) mb_mufu_kernel(int iters, float *sink, float seed) {
  float a0 = seed, a1 = seed + 0.1f, a2 = seed + 0.2f, a3 = seed + 0.3f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = __log2f(a0) + 2.0f;
      a1 = __log2f(a1) + 2.0f;
      a2 = __log2f(a2) + 2.0f;
      a3 = __log2f(a3) + 2.0f;
    }
  }
  if (a0 + a1 + a2 + a3 == 12345.678f) sink[0] = a0;
}

}  // namespace nk

using namespace nk;

// which: 0 smem read GB/s, 1 L2 read GB/s, 2 fp32 FMA GFLOP/s (2 flop/FMA), 3 MUFU Gop/s, 4 fp64 FMA GFLOP/s.
extern "C" int nk_microbench(int32_t which, double *result_host) {
  NK_CHECK_ARG(result_host != nullptr && which >= 0 && which <= 4, "nk_microbench: bad arguments");
  cudaStream_t st;
  NK_CUDA_OK(cudaStreamCreate(&st));
  cudaEvent_t e0, e1;
  NK_CUDA_OK(cudaEventCreate(&e0));
  NK_CUDA_OK(cudaEventCreate(&e1));
  float *sink = nullptr;
  NK_CUDA_OK(cudaMalloc(&sink, 64));
  const int sms = num_sms();
  double work = 0.0;
  float4 *buf = nullptr;
  const size_t n4 = (size_t)32 * 1024 * 1024 / 16;
  if (which == 1) {
    NK_CUDA_OK(cudaMalloc(&buf, n4 * 16));
    NK_CUDA_OK(cudaMemsetAsync(buf, 0, n4 * 16, st));
  }
  if (which == 0) NK_CUDA_OK(cudaFuncSetAttribute(mb_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
  float best_ms = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {  // first rep is the warm-up
    NK_CUDA_OK(cudaEventRecord(e0, st));
    switch (which) {
      case 0: {
        const int iters = 4000;
        mb_smem_kernel<<<sms, 1024, 128 * 1024, st>>>(iters, sink);
        work = (double)sms * 1024.0 * iters * 8.0 * 16.0;
        break;
      }
      case 1: {
        const int iters = 20;
        mb_l2_kernel<<<sms * 4, 512, 0, st>>>(buf, n4, iters, sink);
        work = (double)n4 * 16.0 * iters;
        break;
      }
      case 2: {
        const int iters = 4000;
        mb_fma_kernel<float><<<sms * 2, 1024, 0, st>>>(iters, sink, 1.0f);
        work = (double)sms * 2 * 1024.0 * iters * 64.0 * 2.0;
        break;
      }
      case 3: {
        const int iters = 2000;
        mb_mufu_kernel<<<sms * 2, 1024, 0, st>>>(iters, sink, 1.5f);
        work = (double)sms * 2 * 1024.0 * iters * 32.0;
        break;
      }
      case 4: {
        const int iters = 2000;
        mb_fma_kernel<double><<<sms * 2, 1024, 0, st>>>(iters, (double *)sink, 1.0);
        work = (double)sms * 2 * 1024.0 * iters * 64.0 * 2.0;
        break;
      }
    }
    NK_LAUNCH_OK();
    NK_CUDA_OK(cudaEventRecord(e1, st));
    NK_CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0.f;
    NK_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best_ms) best_ms = ms;
  }
  *result_host = work / (best_ms * 1e-3) / 1e9;
  if (buf) cudaFree(buf);
  cudaFree(sink);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaStreamDestroy(st);
  return NK_OK;
}
