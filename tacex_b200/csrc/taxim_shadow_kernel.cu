// taxim_shadow_kernel.cu -- the SHADOW branch of the Taxim optical model (`with_shadow = True`; ref:
// /root/reference/source/tacex/tacex/simulation_approaches/gpu_taxim/sim/taxim_torch.py:260-346), a post-pass on the deformed
// gel + contact mask the fused kernel (taxim_kernel.cu) writes out:
//   1. shadow_cast_kernel     attachment boundary = (contact mask dilated by the two box kernels) minus the mask; every boundary
//                             pixel casts F rays of S samples along (direction bin + fan) and stores the shadow-table value of its
//                             (direction, height) bin with an atomic float minimum where the deformed gel is higher than at the
//                             pixel itself (the reference's scatter_min);
//   2. shadow_rawmin_kernel   polynomial colour WITHOUT background (same normals / bins / table as the colour stage), minimum with
//                             the shadow image;
//   3. blur_h / blur_v        shadow blur, + background, final blur, clip, NHWC store -- the canonical separable correlation
//                             (horizontal: acc = fma(w[k], x[reflect], acc), k ascending; vertical: centre-outward pairs).
// Correctness-first version: plain per-pixel kernels over global scratch planes [n][3][240][320] (bounded loops only, no
// inter-thread protocol); the operation order is that of oracle/taxim_canon.c::canon_taxim_render_shadow, which reproduces the
// executed reference's intermediate shadow image bit for bit. The ray trigonometry comes from host tables (the reference
// evaluates torch.cos / torch.sin of the same float32 angles once at init).
#include "tx_common.cuh"
#include "tx_kernels.h"
#include <math.h>

namespace tx {

constexpr int SH_THREADS = 256;
constexpr int HW = IMG_H * IMG_W;

__device__ __forceinline__ int reflect_i(int i, int n)
{
    i = i < 0 ? -i : i;
    return i >= n ? 2 * (n - 1) - i : i;
}

// minimum of IEEE floats with integer atomics: correct for any mix of signs (the stored value starts as +inf)
__device__ __forceinline__ void atomic_min_float(float* addr, float v)
{
    if (v >= 0.0f)
        atomicMin(reinterpret_cast<int*>(addr), __float_as_int(__fadd_rn(v, 0.0f))); // -0.0f + 0.0f = +0.0f: -0 must not sort below the negatives
    else
        atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void __launch_bounds__(SH_THREADS) shadow_fill_kernel(float* __restrict__ p, size_t n, float v)
{
    const size_t i = (size_t)blockIdx.x * SH_THREADS + threadIdx.x;
    if (i < n) p[i] = v;
}

// canonical gradient direction of pixel (y, x) of the deformed gel `b` (replicate-padded like the colour stage)
__device__ __forceinline__ void mag_dir(const float* __restrict__ b, int y, int x, float inv_pixmm, float sx, float sy, float& mag,
                                        float& dir)
{
    const int yy = min(max(y, 1), IMG_H - 2), xx = min(max(x, 1), IMG_W - 2);
    const float* ctr = b + yy * IMG_W + xx;
    const float top = __fmul_rn(ctr[-IMG_W], inv_pixmm), bot = __fmul_rn(ctr[IMG_W], inv_pixmm);
    const float lef = __fmul_rn(ctr[-1], inv_pixmm), rig = __fmul_rn(ctr[1], inv_pixmm);
    const float gx = __fmul_rn(__fmul_rn(__fadd_rn(top, -bot), 0.5f), sy);
    const float gy = __fmul_rn(__fmul_rn(__fadd_rn(lef, -rig), 0.5f), sx);
    const float tt = __fsqrt_rn(__fmaf_rn(gx, gx, __fmul_rn(gy, gy)));
    mag = atanf_c(tt);
    dir = (tt != 0.0f) ? atan2f_c(gx, gy) : 0.0f;
}

__global__ void __launch_bounds__(SH_THREADS) shadow_cast_kernel(const ShadowArgs a)
{
    const int n = blockIdx.y;
    const int i = blockIdx.x * SH_THREADS + threadIdx.x;
    if (i >= HW) return;
    const int y = i / IMG_W, x = i - y * IMG_W;
    const unsigned char* mk = a.mask + (size_t)n * HW;
    if (mk[i]) return;
    // conv2d(ones(ky, kx), padding = 'same') twice: torch pads the extra element of an even kernel on the bottom / right
    const int p0y = (a.dil[0] - 1) / 2, p0x = (a.dil[1] - 1) / 2, p1y = (a.dil[2] - 1) / 2, p1x = (a.dil[3] - 1) / 2;
    bool hit = false;
    for (int j1 = 0; j1 < a.dil[2] && !hit; ++j1) {
        const int y1 = y - p1y + j1;
        if (y1 < 0 || y1 >= IMG_H) continue;
        for (int i1 = 0; i1 < a.dil[3] && !hit; ++i1) {
            const int x1 = x - p1x + i1;
            if (x1 < 0 || x1 >= IMG_W) continue;
            for (int j0 = 0; j0 < a.dil[0] && !hit; ++j0) {
                const int y0 = y1 - p0y + j0;
                if (y0 < 0 || y0 >= IMG_H) continue;
                for (int i0 = 0; i0 < a.dil[1]; ++i0) {
                    const int x0 = x1 - p0x + i0;
                    if (x0 >= 0 && x0 < IMG_W && mk[y0 * IMG_W + x0]) { hit = true; break; }
                }
            }
        }
    }
    if (!hit) return;
    const float* b = a.deformed + (size_t)n * HW;
    const float PI_F = 3.14159265358979323846f;
    float mag, dir;
    mag_dir(b, y, x, a.inv_pixmm, a.sx, a.sy, mag, dir);
    (void)mag;
    int nidx = (int)floorf(__fdiv_rn(__fadd_rn(dir, PI_F), a.discretize_precision));
    nidx = min(max(nidx, 0), a.D - 1);
    const float g = a.gel ? __ldg(a.gel + i) : 0.0f;
    const float ch_px = __fdiv_rn(__fadd_rn(g, -b[i]), a.pixmm);
    int hidx = (int)floorf(__fdiv_rn(__fadd_rn(__fmul_rn(ch_px, a.pixmm), -a.depth_0), a.height_precision)) + 6;
    const int hmax = a.Hn - 1;
    if (hidx < 0 || hidx >= hmax) hidx = hmax; // the reference's table has no extra empty entry: the last real one (quirk kept)
    const float dsrc = __fdiv_rn(b[i], a.pixmm);
    float* sh = a.shadow + (size_t)n * 3 * HW;
    const float* row0 = a.table + ((size_t)(0 * a.D + nidx) * a.Hn + hidx) * a.S;
    const size_t cstride = (size_t)a.D * a.Hn * a.S;
    for (int f = 0; f < a.F; ++f) {
        const float cs = __ldg(a.fan_cos + nidx * a.F + f), sn = __ldg(a.fan_sin + nidx * a.F + f);
        for (int s = 0; s < a.S; ++s) {
            const float kx = __fmul_rn(a.step_x, (float)(s + 1)), ky = __fmul_rn(a.step_y, (float)(s + 1));
            const int cx = __float2int_rz(__fadd_rn((float)x, __fmul_rn(kx, cs)));
            const int cy = __float2int_rz(__fadd_rn((float)y, __fmul_rn(ky, sn)));
            if (cx < 0 || cx >= IMG_W || cy < 0 || cy >= IMG_H) continue;
            const int tgt = cy * IMG_W + cx;
            if (!(dsrc < __fdiv_rn(b[tgt], a.pixmm))) continue;
#pragma unroll
            for (int c = 0; c < 3; ++c) atomic_min_float(sh + (size_t)c * HW + tgt, __ldg(row0 + c * cstride + s));
        }
    }
}

// polynomial colour without background, minimum with the shadow image (in place in `shadow`)
__global__ void __launch_bounds__(SH_THREADS) shadow_rawmin_kernel(const ShadowArgs a)
{
    const int n = blockIdx.y;
    const int i = blockIdx.x * SH_THREADS + threadIdx.x;
    if (i >= HW) return;
    const int y = i / IMG_W, x = i - y * IMG_W;
    const float* b = a.deformed + (size_t)n * HW;
    const float PI_F = 3.14159265358979323846f;
    float mag, dir;
    mag_dir(b, y, x, a.inv_pixmm, a.sx, a.sy, mag, dir);
    int im = (int)floorf(__fmul_rn(mag, a.inv_xbin));
    int id = (int)floorf(__fmul_rn(__fadd_rn(dir, PI_F), a.inv_ybin));
    im = min(max(im, 0), a.nb - 1);
    id = min(max(id, 0), a.nb - 1);
    const float4* pf = a.poly + (size_t)(im * a.nb + id) * 5;
    const float4 a0 = __ldg(pf), a1 = __ldg(pf + 1), a2 = __ldg(pf + 2), a3 = __ldg(pf + 3), a4 = __ldg(pf + 4);
    const float cf[20] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y,
                          a2.z, a2.w, a3.x, a3.y, a3.z, a3.w, a4.x, a4.y, a4.z, a4.w};
    const float xf = __fmul_rn((float)x, a.fx), yf = __fmul_rn((float)y, a.fy);
    const float f0 = __fmul_rn(xf, xf), f1 = __fmul_rn(yf, yf), f2 = __fmul_rn(xf, yf);
    float* sh = a.shadow + (size_t)n * 3 * HW;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float* pc = cf + 6 * ch;
        float s = pc[5];
        s = __fmaf_rn(pc[4], yf, s);
        s = __fmaf_rn(pc[3], xf, s);
        s = __fmaf_rn(pc[2], f2, s);
        s = __fmaf_rn(pc[1], f1, s);
        s = __fmaf_rn(pc[0], f0, s);
        sh[(size_t)ch * HW + i] = fminf(s, sh[(size_t)ch * HW + i]);
    }
}

// horizontal pass of the canonical blur over [n][3] planes: dst = correlate(src, taps) with reflect padding
__global__ void __launch_bounds__(SH_THREADS) shadow_blur_h_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                                   const float* __restrict__ taps, int ks)
{
    const size_t plane = (size_t)blockIdx.y * HW;
    const int i = blockIdx.x * SH_THREADS + threadIdx.x;
    if (i >= HW) return;
    const int y = i / IMG_W, x = i - y * IMG_W, r = (ks - 1) / 2;
    const float* row = src + plane + (size_t)y * IMG_W;
    float acc = 0.0f;
    for (int k = 0; k < ks; ++k) acc = __fmaf_rn(__ldg(taps + k), row[reflect_i(x + k - r, IMG_W)], acc);
    dst[plane + i] = acc;
}

// vertical pass (centre-outward, symmetric pair summed first). MODE 0: dst plane = result + background (bg is [240][320][3]);
// MODE 1: final pass: clip to [0, 1] and store NHWC into rgb [n][240][320][3]
template <int MODE>
__global__ void __launch_bounds__(SH_THREADS) shadow_blur_v_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                                   const float* __restrict__ taps, int ks,
                                                                   const float* __restrict__ bg_hwc)
{
    const int pl = blockIdx.y; // n * 3 + channel
    const int i = blockIdx.x * SH_THREADS + threadIdx.x;
    if (i >= HW) return;
    const int y = i / IMG_W, x = i - y * IMG_W, r = (ks - 1) / 2;
    const float* p = src + (size_t)pl * HW;
    float acc = __fmul_rn(__ldg(taps + r), p[i]);
    for (int d = 1; d <= r; ++d)
        acc = __fmaf_rn(__ldg(taps + r + d), __fadd_rn(p[reflect_i(y - d, IMG_H) * IMG_W + x], p[reflect_i(y + d, IMG_H) * IMG_W + x]), acc);
    const int n = pl / 3, c = pl - n * 3;
    if (MODE == 0) {
        dst[(size_t)pl * HW + i] = __fadd_rn(acc, __ldg(bg_hwc + (size_t)i * 3 + c));
    } else {
        dst[((size_t)n * HW + i) * 3 + c] = fminf(fmaxf(acc, 0.0f), 1.0f);
    }
}

#ifndef TX_EMULATE // the host emulation of tools/emu drives the kernels itself (no <<< >>> there)
cudaError_t launch_shadow(const ShadowArgs& a, int n, float* t1, float* t2, float* rgb, const float* bg_hwc, const float* taps_sx,
                          int ks_sx, const float* taps_sy, int ks_sy, const float* taps_fx, int ks_fx, const float* taps_fy,
                          int ks_fy, cudaStream_t s)
{
    const size_t total = (size_t)n * 3 * HW;
    const dim3 gpix((HW + SH_THREADS - 1) / SH_THREADS, n), gpl((HW + SH_THREADS - 1) / SH_THREADS, 3 * n);
    shadow_fill_kernel<<<(unsigned)((total + SH_THREADS - 1) / SH_THREADS), SH_THREADS, 0, s>>>(a.shadow, total, INFINITY);
    shadow_cast_kernel<<<gpix, SH_THREADS, 0, s>>>(a);
    shadow_rawmin_kernel<<<gpix, SH_THREADS, 0, s>>>(a);
    shadow_blur_h_kernel<<<gpl, SH_THREADS, 0, s>>>(a.shadow, t1, taps_sx, ks_sx);
    shadow_blur_v_kernel<0><<<gpl, SH_THREADS, 0, s>>>(t1, t2, taps_sy, ks_sy, bg_hwc);
    shadow_blur_h_kernel<<<gpl, SH_THREADS, 0, s>>>(t2, t1, taps_fx, ks_fx);
    shadow_blur_v_kernel<1><<<gpl, SH_THREADS, 0, s>>>(t1, rgb, taps_fy, ks_fy, bg_hwc);
    return cudaGetLastError();
}
#endif

} // namespace tx
