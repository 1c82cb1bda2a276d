// tx_common.cuh -- shared device helpers for the sm_100a kernels of libtacex_b200.
//
// Canonical float32 elementary functions: every operation is one IEEE-754 binary32 op in a fixed order
// (compile with -fmad=false so nvcc never contracts a*b+c on its own); the CPU checker under oracle/
// restates the same sequences, which is what makes bin indices bit-comparable between CPU and GPU.
#pragma once
#include "tx_kernels.h"
#include <cuda_runtime.h>
#include <stdint.h>

namespace tx {

// ---- canonical atan / atan2 (Cephes-style polynomial) ------------------------------
// Branch-free forms (one IEEE division each); bit-identical to the branchy CPU restatement:
//   |x| > tan(3pi/8): t = (-1)/x   == -(1/x)        |x| > tan(pi/8): t = (x-1)/(x+1)        else: t = x/1 == x
__device__ __forceinline__ float atanf_c(float xx)
{
    const float x = fabsf(xx);
    const bool big = x > 2.414213562373095f;
    const bool mid = x > 0.4142135623730950f;
    const float num = big ? -1.0f : (mid ? __fadd_rn(x, -1.0f) : x);
    const float den = big ? x : (mid ? __fadd_rn(x, 1.0f) : 1.0f);
    const float y0 = big ? 1.5707963267948966f : (mid ? 0.7853981633974483f : 0.0f);
    const float t = __fdiv_rn(num, den);
    const float z = __fmul_rn(t, t);
    float p = 8.05374449538e-2f;
    p = __fmaf_rn(p, z, -1.38776856032e-1f);
    p = __fmaf_rn(p, z, 1.99777106478e-1f);
    p = __fmaf_rn(p, z, -3.33329491539e-1f);
    p = __fmul_rn(p, z);
    p = __fmaf_rn(p, t, t);
    const float y = __fadd_rn(y0, p);
    return (xx < 0.0f) ? -y : y;
}

// atan2 for (y, x) not both zero. x == 0 gives y/x = +-inf and atanf_c(+-inf) = +-pi/2 exactly as the explicit case.
__device__ __forceinline__ float atan2f_c(float y, float x)
{
    const float PI_F = 3.14159265358979323846f;
    const float z = atanf_c(__fdiv_rn(y, x));
    const float adj = (y < 0.0f) ? -PI_F : PI_F;
    return (x < 0.0f) ? __fadd_rn(z, adj) : z;
}

// polynomial colour of one pixel: 3 channels x 6 coefficients (cf[6 ch + k], k over x^2, y^2, xy, x, y, 1), + background, clip
__device__ __forceinline__ void poly_rgb(const float* cf, float xf, float yf, float f0, float f1, float f2, const float* bgv,
                                         float* o)
{
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float* pc = cf + 6 * ch;
        float s = pc[5];
        s = __fmaf_rn(pc[4], yf, s);
        s = __fmaf_rn(pc[3], xf, s);
        s = __fmaf_rn(pc[2], f2, s);
        s = __fmaf_rn(pc[1], f1, s);
        s = __fmaf_rn(pc[0], f0, s);
        s = __fadd_rn(s, bgv[ch]);
        o[ch] = fminf(fmaxf(s, 0.0f), 1.0f);
    }
}

// ---- packed float32 x 2 arithmetic (Blackwell FFMA2 / FADD2 / FMUL2): two IEEE binary32 operations per instruction ------
// Each half is rounded exactly like the scalar instruction, so packing two independent outputs changes no bit of either.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// ---- mbarrier + bulk async copy (TMA, 1-D) -----------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy (SASS: UBLKCP), completion signalled on an mbarrier; size multiple of 16 B
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 16-byte store to a multicast address (NVLink SHARP / NVLS mapping of a symmetric allocation): the switch replicates it
__device__ __forceinline__ void multimem_st_f4(float* mc_addr, float4 v)
{
    asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// ---- warp reductions ---------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_min(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ unsigned warp_sum_u32(unsigned v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

} // namespace tx
