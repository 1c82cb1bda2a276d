// overlay_kernel.cu -- marker image + marker overlay of the tactile RGB observation, batched over envs.
//
// Replaces the per-env host loop of the reference's RL task (ref: source/tacex_tasks/tacex_tasks/ball_rolling_tactile/
// ball_rolling_taxim_fots.py:918-937) around FOTSMarkerSimulator.draw_markers (ref: source/tacex/tacex/simulation_approaches/
// fots/fots_marker_sim.py:346-384):
//   marker image  (H + 24) x (W + 24) uint8 canvas of 255; for every marker IN ORDER paste the 12 x 12 anti-aliased dot
//                 patch [floor(frac(u) * 10)][floor(frac(v) * 10)] at (floor(v) - 6, floor(u) - 6) (u = x + 0.5 + 12, v = y + 0.5
//                 + 12 in float64 like NumPy; later markers overwrite earlier ones); crop the 12-pixel border;
//   overlay       rgb = ((rgb * 255) * (marker / 255)) / 255 per channel in float32, in the reference's operation order.
// One CTA per env; the canvas lives in shared memory; optional uint8 outputs (the marker image itself, and the overlaid RGB
// rounded to uint8 -- 4 x fewer bytes for the observation transport).
#include "tx_kernels.h"
#include <stdint.h>

namespace tx {

constexpr int OV_PAD = 12, OV_PATCH = 12, OV_SR = 10;
constexpr int OV_CW = IMG_W + 2 * OV_PAD, OV_CH = IMG_H + 2 * OV_PAD;
constexpr int OV_THREADS = 256;

__global__ void __launch_bounds__(OV_THREADS) marker_overlay_kernel(const OverlayArgs a)
{
    extern __shared__ unsigned char canvas[]; // [OV_CH][OV_CW]
    const int n = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < OV_CH * OV_CW / 4; i += OV_THREADS) reinterpret_cast<unsigned*>(canvas)[i] = 0xffffffffu;
    __syncthreads();
    const float* mk = a.markers + ((size_t)n * 2 + 1) * a.M * 2; // [:, 1] = current positions (x, y)
    for (int k = 0; k < a.M; ++k) { // in order: later markers overwrite earlier ones
        const double u = (double)mk[2 * k] + 0.5 + (double)OV_PAD;
        const double v = (double)mk[2 * k + 1] + 0.5 + (double)OV_PAD;
        const double fu = floor(u), fv = floor(v);
        const int pu = (int)floor((u - fu) * (double)OV_SR), pv = (int)floor((v - fv) * (double)OV_SR);
        const int cu = (int)fu - OV_PATCH / 2, cv = (int)fv - OV_PATCH / 2;
        if (cu >= 0 && cu < OV_CW - OV_PATCH && cv >= 0 && cv < OV_CH - OV_PATCH && pu >= 0 && pu < OV_SR && pv >= 0 && pv < OV_SR) {
            if (tid < OV_PATCH * OV_PATCH) {
                const int py = tid / OV_PATCH, px = tid - py * OV_PATCH;
                canvas[(cv + py) * OV_CW + cu + px] = __ldg(a.patch + ((size_t)(pu * OV_SR + pv) * OV_PATCH + py) * OV_PATCH + px);
            }
        }
        __syncthreads();
    }
    const size_t frame = (size_t)n * IMG_H * IMG_W;
    for (int i = tid; i < IMG_H * IMG_W; i += OV_THREADS) {
        const int y = i / IMG_W, x = i - y * IMG_W;
        const unsigned char m = canvas[(y + OV_PAD) * OV_CW + x + OV_PAD];
        if (a.marker_img) a.marker_img[frame + i] = m;
        if (a.rgb || a.rgb_u8) {
            const float w = __fdiv_rn((float)m, 255.0f);
            float o[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float r = a.rgb_in[(frame + i) * 3 + c];
                o[c] = a.apply ? __fdiv_rn(__fmul_rn(__fmul_rn(r, 255.0f), w), 255.0f) : r;
            }
            if (a.rgb) { a.rgb[(frame + i) * 3] = o[0]; a.rgb[(frame + i) * 3 + 1] = o[1]; a.rgb[(frame + i) * 3 + 2] = o[2]; }
            if (a.rgb_u8) {
#pragma unroll
                for (int c = 0; c < 3; ++c) // round to nearest (ties to even), clamp to [0, 255]
                    a.rgb_u8[(frame + i) * 3 + c] = (unsigned char)__float2int_rn(fminf(fmaxf(__fmul_rn(o[c], 255.0f), 0.0f), 255.0f));
            }
        }
    }
}

// plain float32 -> uint8 conversion of the observation (no marker image): elementwise, 16 floats per thread and iteration
__global__ void __launch_bounds__(256) rgb_to_u8_kernel(const float4* __restrict__ in, uint4* __restrict__ out, size_t n16)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        unsigned w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 v = __ldg(in + 4 * i + k);
            const unsigned b0 = (unsigned)__float2int_rn(fminf(fmaxf(__fmul_rn(v.x, 255.0f), 0.0f), 255.0f));
            const unsigned b1 = (unsigned)__float2int_rn(fminf(fmaxf(__fmul_rn(v.y, 255.0f), 0.0f), 255.0f));
            const unsigned b2 = (unsigned)__float2int_rn(fminf(fmaxf(__fmul_rn(v.z, 255.0f), 0.0f), 255.0f));
            const unsigned b3 = (unsigned)__float2int_rn(fminf(fmaxf(__fmul_rn(v.w, 255.0f), 0.0f), 255.0f));
            w[k] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
        }
        out[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

cudaError_t launch_marker_overlay(const OverlayArgs& a, int N, cudaStream_t s)
{
    if (a.M == 0 && !a.apply && !a.marker_img && !a.rgb && a.rgb_u8 && !((uintptr_t)a.rgb_in & 15u) && !((uintptr_t)a.rgb_u8 & 15u)) {
        const size_t n16 = (size_t)N * IMG_H * IMG_W * 3 / 16;
        rgb_to_u8_kernel<<<148 * 8, 256, 0, s>>>(reinterpret_cast<const float4*>(a.rgb_in), reinterpret_cast<uint4*>(a.rgb_u8), n16);
        return cudaGetLastError();
    }
    static bool attr_dev[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 63;
    if (!attr_dev[dev] || dev == 63) {
        cudaError_t e = cudaFuncSetAttribute(marker_overlay_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OV_CH * OV_CW);
        if (e != cudaSuccess) return e;
        attr_dev[dev] = true;
    }
    marker_overlay_kernel<<<N, OV_THREADS, OV_CH * OV_CW, s>>>(a);
    return cudaGetLastError();
}

// ---- antialiased bilinear resize of the height map (camera FINER than the tactile image) -----------------------------------------
// ref: TaximSimulator.optical_simulation, taxim_sim.py:88-89 (torchvision F.resize = aten's separable antialias kernel): float32
// weights (host, tx_api.cu), horizontal pass then vertical pass, each acc = w[0] x[0]; acc = fmaf(w[k], x[k], acc). One thread per
// output pixel evaluates the (at most 8 x 8) taps directly -- every intermediate row value is computed with exactly the arithmetic of
// the two-pass form, so the result is bit-identical to it (and to torch) without a temporary plane in HBM.
__global__ void __launch_bounds__(256) resize_aa_kernel(const ResizeArgs a)
{
    const int X = blockIdx.x * blockDim.x + threadIdx.x, Y = blockIdx.y, n = blockIdx.z;
    if (X >= a.Wo) return;
    const int fx = __ldg(a.fx + X), cx = __ldg(a.cx + X), fy = __ldg(a.fy + Y), cy = __ldg(a.cy + Y);
    const float* wx = a.wx + X * TX_RS_TAPS;
    const float* wy = a.wy + Y * TX_RS_TAPS;
    const float* s = a.src + ((size_t)n * a.Hi + fy) * a.Wi + fx;
    float acc = 0.0f;
    for (int j = 0; j < cy; ++j) {
        const float* sp = s + (size_t)j * a.Wi;
        float t = __fmul_rn(__ldg(wx), __ldg(sp));
        for (int k = 1; k < cx; ++k) t = __fmaf_rn(__ldg(wx + k), __ldg(sp + k), t);
        acc = j == 0 ? __fmul_rn(__ldg(wy), t) : __fmaf_rn(__ldg(wy + j), t, acc);
    }
    a.dst[((size_t)n * a.Ho + Y) * a.Wo + X] = acc;
}

cudaError_t launch_resize_aa(const ResizeArgs& a, int N, cudaStream_t s)
{
    resize_aa_kernel<<<dim3((a.Wo + 255) / 256, a.Ho, N), 256, 0, s>>>(a);
    return cudaGetLastError();
}

} // namespace tx
