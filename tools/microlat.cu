// Latency / throughput micro-measurements of the primitives in the sweep kernel's critical path (B200, sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microlat tools/microlat.cu && ./microlat
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096

__device__ __forceinline__ float lg2f_(float x) { float y; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2f_(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int OP>
__global__ void lat_kernel(long long *out, float seed, int iseed) {
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = (float)((i * 7 + 1) & 4095);
  __syncthreads();
  float x = seed + threadIdx.x * 1e-3f;
  int xi = iseed + threadIdx.x;
  unsigned long long xx = 0x3f8000003f800000ull;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    if (OP == 0) x = __shfl_sync(0xffffffffu, x, (xi + it) & 31);                       // SHFL.IDX dependent chain
    if (OP == 1) xi = __reduce_add_sync(0xffffffffu, xi) >> 5;                          // REDUX + shift dependent chain
    if (OP == 2) x = lg2f_(x) + 3.0f;                                                   // MUFU.LG2 + FADD
    if (OP == 3) { xi = __float2int_rn(x); x = (float)xi * 1.0001f + 0.5f; }            // F2I + I2F + FFMA
    if (OP == 4) { int a = (int)x & 4095; x = sm[a]; }                                  // LDS dependent (pointer chase) incl. F2I
    if (OP == 5) x = x * 1.0001f + 0.5f;                                                // FFMA
    if (OP == 6) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(xx));                // FFMA2 dependent
    if (OP == 7) x = ex2f_(x) * 0.25f;                                                  // MUFU.EX2 + FMUL
    if (OP == 8) { int a = xi & 4095; xi = __float_as_int(sm[a]) & 4095; }              // LDS pure int chase
    if (OP == 9) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0f;                         // SHFL.BFLY + FADD
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (x == 12345.f || xi == 12345 || xx == 1) out[1000] = 1;
}

// throughput: many warps, independent ops
template <int OP>
__global__ void tput_kernel(long long *out, float seed, int iseed) {
  float x0 = seed + threadIdx.x * 1e-3f, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
  int i0 = iseed + threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3;
  unsigned long long a0 = 0x3f8000003f800000ull, a1 = a0, a2 = a0, a3 = a0;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    if (OP == 0) { x0 = __shfl_sync(0xffffffffu, x0, it & 31); x1 = __shfl_sync(0xffffffffu, x1, it & 31); x2 = __shfl_sync(0xffffffffu, x2, it & 31); x3 = __shfl_sync(0xffffffffu, x3, it & 31); }
    if (OP == 1) { i0 = __reduce_add_sync(0xffffffffu, i0) >> 5; i1 = __reduce_add_sync(0xffffffffu, i1) >> 5; i2 = __reduce_add_sync(0xffffffffu, i2) >> 5; i3 = __reduce_add_sync(0xffffffffu, i3) >> 5; }
    if (OP == 2) { x0 = lg2f_(x0) + 3.f; x1 = lg2f_(x1) + 3.f; x2 = lg2f_(x2) + 3.f; x3 = lg2f_(x3) + 3.f; }
    if (OP == 3) { i0 = __float2int_rn(x0); i1 = __float2int_rn(x1); i2 = __float2int_rn(x2); i3 = __float2int_rn(x3); x0 += i0; x1 += i1; x2 += i2; x3 += i3; }
    if (OP == 6) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a0) : "l"(a1), "l"(a2)); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a1) : "l"(a2), "l"(a3));
                   asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a2) : "l"(a3), "l"(a0)); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a3) : "l"(a0), "l"(a1)); }
    if (OP == 10) { asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(a0) : "l"(a1)); asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(a1) : "l"(a2));
                    asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(a2) : "l"(a3)); asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(a3) : "l"(a0)); }
    if (OP == 5) { x0 = x0 * x1 + x2; x1 = x1 * x2 + x3; x2 = x2 * x3 + x0; x3 = x3 * x0 + x1; }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (x0 + x1 + x2 + x3 == 12345.f || i0 + i1 + i2 + i3 == 12345 || a0 + a1 + a2 + a3 == 1) out[1000] = 1;
}

int main() {
  long long *d; cudaMalloc(&d, 8192 * 8); long long h[4];
  const char *names[] = {"SHFL.IDX", "REDUX+SHF", "LG2+FADD", "F2I+I2F+FFMA", "F2I+LOP+LDS", "FFMA", "FFMA2", "EX2+FMUL", "LOP+LDS(int chase)", "SHFL.BFLY+FADD"};
#define LAT(OP) lat_kernel<OP><<<1, 32, 16384>>>(d, 1.5f, 3); cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost); printf("latency  %-22s %7.1f cycles/iter\n", names[OP], (double)h[0] / ITERS);
  LAT(0) LAT(1) LAT(2) LAT(3) LAT(4) LAT(5) LAT(6) LAT(7) LAT(8) LAT(9)
  // throughput with W warps on one SM: cycles per warp-instruction per SM
  const int Ws[] = {4, 8, 16, 32};
#define TP(OP, NAME) for (int w : Ws) { tput_kernel<OP><<<1, 32 * w>>>(d, 1.5f, 3); cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost); \
    printf("tput     %-12s warps=%2d  %6.2f cycles per warp-instr per SM\n", NAME, w, (double)h[0] / (ITERS * 4.0 * w)); }
  TP(0, "SHFL.IDX") TP(1, "REDUX") TP(2, "LG2(+FADD)") TP(3, "F2I(+IADD)") TP(5, "FFMA") TP(6, "FFMA2") TP(10, "FMUL2")
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
