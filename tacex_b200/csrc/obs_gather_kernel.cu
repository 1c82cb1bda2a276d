// obs_gather_kernel.cu -- the observation all-gather of the multi-GPU path, reduced to the pixels that carry information.
//
// envs shard contiguously over the GPUs and ONE all-gather of the final float32 RGB observation is the
// only cross-GPU traffic. A tactile frame differs from the precomputed flat image (zero gradient everywhere) only inside the
// rectangle the fused Taxim kernel reports per half frame (`TaximArgs::rect_out`), so every rank
//   1. pushes the pixels of its rectangles (+ the 16-byte rectangle descriptors) straight into the gathered buffers of all
//      peers with peer-to-peer stores over NVLink / NVSwitch (symmetric memory mapped by the host side), and
//   2. after a barrier, fills the complement of the rectangles of the REMOTE envs in its own gathered buffer from its local
//      flat image (local HBM traffic only).
// The result is bit-identical to gathering whole frames; the link traffic drops to the rectangle share (about half for the
// benchmark's sphere presses, 16 bytes for an env without contact).
#include "tx_kernels.h"

namespace tx {

constexpr int OG_THREADS = 512;
constexpr int QPR = IMG_W / 4; // 4-pixel quads per row (48 bytes = 3 float4 each)

// 16-byte store to a multicast address (NVLink SHARP / NVLS mapping of a symmetric allocation)
__device__ __forceinline__ void mc_store(float* mc_addr, float4 v)
{
    asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

__global__ void __launch_bounds__(OG_THREADS) obs_push_kernel(const ObsPushArgs a)
{
    constexpr int UNR = 4; // independent 16-byte loads in flight per thread (the grid is small: see tx_obs_push)
    // one (env, half) item per CTA iteration
    for (int item = blockIdx.x; item < 2 * a.N; item += gridDim.x) {
        const int4 r = reinterpret_cast<const int4*>(a.rect_local)[item]; // ry0, ry1 (local rows of the half), xa, xb
        if (threadIdx.x == 0) {
            if (a.mc_rect)
                mc_store(reinterpret_cast<float*>(a.mc_rect) + (size_t)item * 4,
                         make_float4(__int_as_float(r.x), __int_as_float(r.y), __int_as_float(r.z), __int_as_float(r.w)));
            else
                for (int p = 0; p < a.n_peers; ++p) reinterpret_cast<int4*>(a.peer_rect[p])[item] = r;
        }
        if (r.y < r.x || r.w < r.z) continue;
        const int nf4 = (r.w - r.z + 1) * 3 / 4;              // float4 per row segment (the width is a multiple of 4 pixels)
        const int total = (r.y - r.x + 1) * nf4;
        const float inv = 1.0f / (float)nf4;
        const size_t base = ((size_t)item * HALF_H * IMG_W) * 3; // float offset of this half frame
        for (int i0 = threadIdx.x; i0 < total; i0 += OG_THREADS * UNR) {
            float4 v[UNR];
            size_t off[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int i = i0 + u * OG_THREADS;
                const int rr = __float2int_rz(((float)i + 0.5f) * inv); // == i / nf4 (i < 2^17, nf4 <= 240)
                off[u] = base + ((size_t)(r.x + rr) * IMG_W + r.z) * 3 + (size_t)(i - rr * nf4) * 4;
                if (i < total) v[u] = __ldcs(reinterpret_cast<const float4*>(a.rgb_local + off[u]));
            }
            if (a.mc_rgb) { // one multicast store: the NVSwitch replicates it into every GPU's buffer (this GPU's included)
#pragma unroll
                for (int u = 0; u < UNR; ++u)
                    if (i0 + u * OG_THREADS < total) mc_store(a.mc_rgb + off[u], v[u]);
            } else {
#pragma unroll 1
                for (int p = 0; p < a.n_peers; ++p) {
                    float* dst = a.peer_rgb[p];
#pragma unroll
                    for (int u = 0; u < UNR; ++u)
                        if (i0 + u * OG_THREADS < total) *reinterpret_cast<float4*>(dst + off[u]) = v[u];
                }
            }
        }
    }
}

// Completes the REMOTE envs of this rank's gathered buffer. The senders stored the pixels of the new rectangles; everything else
// of a frame equals the flat image. `prev_rect` (local, one per gathered buffer) remembers the rectangle each half frame held the
// last time this buffer was filled, so only the pixels of the OLD rectangle that the NEW one does not cover have to be restored
// from the flat image (nothing at all while a contact does not move). prev_rect starts as the whole half frame (first fill =
// complete fill); without it every pixel outside the new rectangle is written.
__global__ void __launch_bounds__(OG_THREADS) obs_fill_kernel(const ObsFillArgs a)
{
    for (int item = blockIdx.x; item < 2 * a.N_total; item += gridDim.x) {
        const int env = item >> 1;
        if (env >= a.skip_lo && env < a.skip_hi) continue; // this rank's own envs were rendered in place
        const int4 r = reinterpret_cast<const int4*>(a.rect_all)[item];
        int4 o = make_int4(0, HALF_H - 1, 0, IMG_W - 1);
        if (a.prev_rect) {
            o = reinterpret_cast<const int4*>(a.prev_rect)[item];
            __syncthreads(); // every thread has read the old rectangle before thread 0 replaces it
            if (threadIdx.x == 0) reinterpret_cast<int4*>(a.prev_rect)[item] = r;
        }
        if (o.y < o.x || o.w < o.z) continue;
        const bool has = r.y >= r.x && r.w >= r.z;
        if (has && r.x <= o.x && r.y >= o.y && r.z <= o.z && r.w >= o.w) continue; // the new rectangle covers the old one
        const float* src = a.flat_rgb + (size_t)(item & 1) * HALF_H * IMG_W * 3;
        float* dst = a.rgb_all + (size_t)item * HALF_H * IMG_W * 3;
        const int q0 = o.z >> 2, nq = (o.w >> 2) - q0 + 1, total = (o.y - o.x + 1) * nq;
        const float inv = 1.0f / (float)nq;
        for (int i = threadIdx.x; i < total; i += OG_THREADS) {
            const int rr = __float2int_rz(((float)i + 0.5f) * inv); // == i / nq
            const int row = o.x + rr, x0 = (q0 + i - rr * nq) * 4;
            if (has && row >= r.x && row <= r.y && x0 >= r.z && x0 <= r.w) continue;
            const size_t qd = (size_t)row * QPR + (x0 >> 2);
            const float4* s4 = reinterpret_cast<const float4*>(src + qd * 12);
            float4* d4 = reinterpret_cast<float4*>(dst + qd * 12);
            const float4 t0 = __ldg(s4), t1 = __ldg(s4 + 1), t2 = __ldg(s4 + 2);
            d4[0] = t0; d4[1] = t1; d4[2] = t2;
        }
    }
}

cudaError_t launch_obs_push(const ObsPushArgs& a, int grid, cudaStream_t s)
{
    obs_push_kernel<<<grid, OG_THREADS, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_obs_fill(const ObsFillArgs& a, int grid, cudaStream_t s)
{
    obs_fill_kernel<<<grid, OG_THREADS, 0, s>>>(a);
    return cudaGetLastError();
}

} // namespace tx
