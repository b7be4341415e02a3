// fp32 building blocks of the product-form proposal loop, shared by the tuned fp32 kernel (sweep_fast.cu) and by the fp32
// "shadow" decisions of the fp64 kernel (sweep_shadow.cu): packed FFMA2 / FMUL2, explicit shared-memory accesses, the lane
// <-> hidden-unit map of a table row and the lane product.
#pragma once

#include "kernels.cuh"

namespace nk {
namespace fast {

constexpr float EXP_RANGE = 120.0f;      // log2 headroom allowed for a lane product
constexpr float FX_SCALE = 524288.0f;    // 2^19: fixed-point scale of per-lane log2 partials (REDUX add)
constexpr int THR_MIN = -(1 << 29);      // "always accept" threshold (u == 0 or machine_pow == 0): -1024 in log2 units

typedef unsigned long long u64;

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<u64 *>(&a)), "l"(*reinterpret_cast<u64 *>(&b)),
      "l"(*reinterpret_cast<u64 *>(&c)));
  return *reinterpret_cast<float2 *>(&d);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<u64 *>(&a)), "l"(*reinterpret_cast<u64 *>(&b)));
  return *reinterpret_cast<float2 *>(&d);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<u64 *>(&a)), "l"(*reinterpret_cast<u64 *>(&b)));
  return *reinterpret_cast<float2 *>(&d);
}

// ---- mbarrier / TMA bulk copy (global -> shared), PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- explicit shared-space accesses (32-bit shared addresses kept in registers; no generic-address arithmetic)
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ uint4 lds128u(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128u(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float2 lds64(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ float lg2_fast(float x) {  // x is a positive normal number by construction
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_fast(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// A lane owns NE = 4 NFULL + TAIL hidden units: NP2 float2 pairs (FFMA2 / FMUL2) plus, for TAIL == 1, one scalar.
template <int NFULL, int TAIL>
struct Lanes {
  static constexpr int NE = 4 * NFULL + TAIL;                      // hidden units per lane
  static constexpr int NP2 = 2 * NFULL + (TAIL == 2 ? 1 : 0);      // float2 pairs per lane
  static constexpr int NPA = NP2 > 0 ? NP2 : 1;                    // array extent (no zero-sized arrays)
  static constexpr bool HAS_T = TAIL == 1;                         // odd unit carried as a scalar
  static constexpr int MP = 128 * NFULL + 32 * TAIL;               // padded row length of the G table (floats)
  // hidden-unit index of element e of this lane (may be >= M: padding)
  static __device__ __forceinline__ int unit(int e, int lane) {
    return e < 4 * NFULL ? 128 * (e >> 2) + 4 * lane + (e & 3) : 128 * NFULL + TAIL * lane + (e - 4 * NFULL);
  }
  // row_lane = shared address of G[i][0] + 16 * lane ; tail_lane = shared address of G[i][128*NFULL + TAIL*lane]
  static __device__ __forceinline__ void load_row(uint32_t row_lane, uint32_t tail_lane, float2 (&g2)[NPA], float &gt) {
#pragma unroll
    for (int q = 0; q < NFULL; ++q) {
      const float4 v = lds128(row_lane + 512 * q);
      g2[2 * q] = make_float2(v.x, v.y);
      g2[2 * q + 1] = make_float2(v.z, v.w);
    }
    if (TAIL == 1) gt = lds32(tail_lane);
    if (TAIL == 2) g2[2 * NFULL] = lds64(tail_lane);
  }
};


// Per-chain fp32 registers of one warp.
template <int NPA>
struct ChainRegs {
  float2 A2[NPA], B2[NPA];
  float At, Bt;          // the scalar unit (TAIL == 1)
  int R;                 // fixed-point log2 prod_j (A_j + B_j), summed over the warp
  uint32_t nacc;         // accepted moves of this call
  uint32_t next_renorm;  // renormalise when nacc reaches this
};

// lane product prod_j (X_j g_j + Y_j) over the lane's units
template <int NP2, int NPA, bool HAS_T>
__device__ __forceinline__ float lane_product(const float2 (&X)[NPA], float Xt, const float2 (&Y)[NPA], float Yt, const float2 (&g2)[NPA],
                                              float gt) {
  float P = 1.0f;
  if (NP2 > 0) {
    float2 Pa = ffma2(X[0], g2[0], Y[0]);
    float2 Pb = make_float2(1.0f, 1.0f);
    if (NP2 > 1) Pb = ffma2(X[1], g2[1], Y[1]);
#pragma unroll
    for (int q = 2; q < NP2; ++q) {
      const float2 c = ffma2(X[q], g2[q], Y[q]);
      if (q & 1)
        Pb = fmul2(Pb, c);
      else
        Pa = fmul2(Pa, c);
    }
    if (NP2 > 1) Pa = fmul2(Pa, Pb);
    P = Pa.x * Pa.y;
  }
  if (HAS_T) {
    const float ct = fmaf(Xt, gt, Yt);
    P = NP2 > 0 ? P * ct : ct;
  }
  return P;
}

// fixed-point log2 prod_j (A_j + B_j), summed over the warp
template <int NP2, int NPA, bool HAS_T>
__device__ __forceinline__ float lane_norm(const ChainRegs<NPA> &c) {
  float P = 1.0f;
  if (NP2 > 0) {
    float2 Pa = fadd2(c.A2[0], c.B2[0]);
#pragma unroll
    for (int q = 1; q < NP2; ++q) Pa = fmul2(Pa, fadd2(c.A2[q], c.B2[q]));
    P = Pa.x * Pa.y;
  }
  if (HAS_T) P *= c.At + c.Bt;
  return P;
}


}  // namespace fast
}  // namespace nk
