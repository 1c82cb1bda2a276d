// taxim_generic_kernel.cu -- the Taxim optical model at an ARBITRARY (small) tactile resolution, one CTA per frame.
//
// The reference scales every blur sigma with the image shape (taxim_impl.py:33-47), and its RL tasks render the tactile image at
// 32 x 24 / 32 x 32 (ref: tacex_tasks/.../ball_rolling_taxim_fots.py:306-321 `tactile_img_res=(32, 24)`,
// ball_rolling_tactile_rgb.py:303-318). The 2-CTA kernel of taxim_kernel.cu is specialised for the GelSight Mini's 240 x 320 /
// radii (30,16,8,4,2,1,2); this kernel covers every other shape of up to TXG_MAX_PIXELS pixels with run-time tap counts. Same
// operation sequence as the canonical restatement (oracle/taxim_canon.c) and as the specialised kernel, hence bit-identical:
//   horizontal pass   acc = 0; acc = fma(w[k], x[reflect(c + k - r)], acc), k ascending
//   vertical pass     acc = w[r] * x[row]; acc = fma(w[r + d], x[reflect(row - d)] + x[reflect(row + d)], acc), d = 1..r
//   re-imposition     plane[mask] = min(h, gel) recomputed from the input frame
// Replaces the same reference functions as taxim_kernel.cu (taxim_sim.py:80-131, taxim_torch.py:243-258,432-503).
// Two planes of the frame ping-pong in shared memory; the work per frame is tiny (768 pixels at 32 x 24), so the kernel is
// bound by launch / latency, not by HBM or the FP32 pipe.
#include "tx_common.cuh"
#include "tx_kernels.h"

namespace tx {

constexpr int TXG_THREADS = 256;

__device__ __forceinline__ int reflect_idx(int i, int n)
{
    i = i < 0 ? -i : i;
    return i >= n ? 2 * (n - 1) - i : i;
}

__global__ void __launch_bounds__(TXG_THREADS) taxim_generic_kernel(const TaximGenericArgs p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int H = p.H, W = p.W, HW = H * W;
    float* A = reinterpret_cast<float*>(smem);
    float* B = A + HW;
    unsigned char* mk = reinterpret_cast<unsigned char*>(B + HW);
    __shared__ float red[TXG_THREADS / 32];
    const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* hm = p.hm + (size_t)n * HW;
    const bool depth_in = p.input_is_depth != 0;

    // ---- load (+ optional GelSightSensor._get_height_map, ref: gelsight_sensor.py:581-593), frame minimum ----------------
    float mloc = __int_as_float(0x7f800000);
    for (int i = tid; i < HW; i += TXG_THREADS) {
        float v = __ldg(hm + i);
        if (depth_in) {
            v = __fmul_rn(isinf(v) ? p.clip_max_m : v, 1000.0f);
            if (p.hm_out) p.hm_out[(size_t)n * HW + i] = v;
        }
        A[i] = v;
        mloc = fminf(mloc, v);
    }
    mloc = warp_min(mloc);
    if (lane == 0) red[warp] = mloc;
    __syncthreads();
    float m = red[0];
#pragma unroll
    for (int w = 1; w < TXG_THREADS / 32; ++w) m = fminf(m, red[w]);

    // ---- indentation depth (explicit, or fused: ref taxim_sim.py:115-131) ----------------------------------------------
    float press;
    if (p.press_in) {
        press = __ldg(p.press_in + n);
    } else {
        float d = __fdiv_rn(m, 1000.0f);
        d = __fadd_rn(d, -p.gelpad_min);
        d = d < 0.0f ? 0.0f : d;
        press = (d <= p.gelpad_h) ? __fmul_rn(__fadd_rn(p.gelpad_h, -d), 1000.0f) : 0.0f;
    }
    if (p.depth_out && tid == 0) p.depth_out[n] = press;

    // ---- shifted height map, contact mask, joined map (ref: taxim_torch.py:441-461) ---------------------------------------
    const float thr = __fmul_rn(-press, p.contact_scale);
    auto joined = [&](float v, int i, bool& mask) -> float {
        const float h = __fadd_rn(__fadd_rn(v, -m), -press);
        const float g = p.gel ? __ldg(p.gel + i) : 0.0f;
        const float j = fminf(h, g);
        mask = (__fadd_rn(j, -g) < thr) && (h < 0.0f);
        return j;
    };
    for (int i = tid; i < HW; i += TXG_THREADS) {
        bool mask;
        A[i] = joined(A[i], i, mask);
        mk[i] = mask ? 1 : 0;
        if (p.mask_out) p.mask_out[(size_t)n * HW + i] = mask ? 1 : 0;
    }
    __syncthreads();

    // ---- Gaussian pyramid with masked re-imposition + final blur (ref: taxim_torch.py:463-471) -----------------------------
    for (int l = 0; l < p.n_blurs; ++l) {
        const float* tx = p.taps + (size_t)(l * 2 + 0) * TX_MAX_TAPS;
        const float* ty = p.taps + (size_t)(l * 2 + 1) * TX_MAX_TAPS;
        const int ksx = p.ksx[l], rx = (ksx - 1) / 2, ry = (p.ksy[l] - 1) / 2;
        for (int i = tid; i < HW; i += TXG_THREADS) {
            const int y = i / W, x = i - y * W;
            const float* row = A + y * W;
            float acc = 0.0f;
            for (int k = 0; k < ksx; ++k) acc = __fmaf_rn(__ldg(tx + k), row[reflect_idx(x + k - rx, W)], acc);
            B[i] = acc;
        }
        __syncthreads();
        const bool last = l == p.n_blurs - 1;
        for (int i = tid; i < HW; i += TXG_THREADS) {
            const int y = i / W, x = i - y * W;
            float acc = __fmul_rn(__ldg(ty + ry), B[i]);
            for (int d = 1; d <= ry; ++d)
                acc = __fmaf_rn(__ldg(ty + ry + d), __fadd_rn(B[reflect_idx(y - d, H) * W + x], B[reflect_idx(y + d, H) * W + x]), acc);
            if (!last && mk[i]) {
                float v = __ldg(hm + i);
                if (depth_in) v = __fmul_rn(isinf(v) ? p.clip_max_m : v, 1000.0f);
                bool mask;
                acc = joined(v, i, mask);
            }
            A[i] = acc;
        }
        __syncthreads();
    }
    if (p.deformed_out)
        for (int i = tid; i < HW; i += TXG_THREADS) p.deformed_out[(size_t)n * HW + i] = A[i];

    // ---- normals -> bins -> polynomial -> + background -> clip -> NHWC (ref: taxim_torch.py:475-503, 243-258) -------------
    const float PI_F = 3.14159265358979323846f;
    float* rgb = p.rgb + (size_t)n * HW * 3;
    for (int i = tid; i < HW; i += TXG_THREADS) {
        const int y = i / W, x = i - y * W;
        // replicate padding of the gradient maps: border pixels take the nearest interior pixel's (mag, dir)
        const int yy = min(max(y, 1), H - 2), xx = min(max(x, 1), W - 2);
        const float* ctr = A + yy * W + xx;
        const float top = __fmul_rn(ctr[-W], p.inv_pixmm), bot = __fmul_rn(ctr[W], p.inv_pixmm);
        const float lef = __fmul_rn(ctr[-1], p.inv_pixmm), rig = __fmul_rn(ctr[1], p.inv_pixmm);
        const float gx = __fmul_rn(__fmul_rn(__fadd_rn(top, -bot), 0.5f), p.sy);
        const float gy = __fmul_rn(__fmul_rn(__fadd_rn(lef, -rig), 0.5f), p.sx);
        const float tt = __fsqrt_rn(__fmaf_rn(gx, gx, __fmul_rn(gy, gy)));
        const float mag = atanf_c(tt);
        const float dir = (tt != 0.0f) ? atan2f_c(gx, gy) : 0.0f;
        int im = (int)floorf(__fmul_rn(mag, p.inv_xbin));
        int id = (int)floorf(__fmul_rn(__fadd_rn(dir, PI_F), p.inv_ybin));
        im = min(max(im, 0), p.nb - 1);
        id = min(max(id, 0), p.nb - 1);
        const float4* pf = p.poly + (size_t)(im * p.nb + id) * 5;
        const float4 a0 = __ldg(pf), a1 = __ldg(pf + 1), a2 = __ldg(pf + 2), a3 = __ldg(pf + 3), a4 = __ldg(pf + 4);
        const float cf[20] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y,
                              a2.z, a2.w, a3.x, a3.y, a3.z, a3.w, a4.x, a4.y, a4.z, a4.w};
        const float xf = __fmul_rn((float)x, p.fx), yf = __fmul_rn((float)y, p.fy);
        const float bgv[3] = {__ldg(p.bg_hwc + 3 * i), __ldg(p.bg_hwc + 3 * i + 1), __ldg(p.bg_hwc + 3 * i + 2)};
        float o[3];
        poly_rgb(cf, xf, yf, __fmul_rn(xf, xf), __fmul_rn(yf, yf), __fmul_rn(xf, yf), bgv, o);
        rgb[3 * i] = o[0]; rgb[3 * i + 1] = o[1]; rgb[3 * i + 2] = o[2];
    }
}

int taxim_generic_smem_bytes(int H, int W) { return H * W * 9 + 16; }

cudaError_t launch_taxim_generic(const TaximGenericArgs& a, int N, cudaStream_t s)
{
    static int smem_set_dev[64] = {}; // per (function, device)
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 63;
    int& smem_set = smem_set_dev[dev];
    const int smem = taxim_generic_smem_bytes(a.H, a.W);
    if (smem > smem_set || dev == 63) {
        cudaError_t e = cudaFuncSetAttribute(taxim_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        smem_set = smem;
    }
    taxim_generic_kernel<<<N, TXG_THREADS, smem, s>>>(a);
    return cudaGetLastError();
}

// ---- indentation depth for any frame size (ref: taxim_sim.py:115-131) ------------------------------------------------------
__global__ void __launch_bounds__(256) indentation_depth_generic_kernel(const float* __restrict__ hm, float* __restrict__ out,
                                                                        int npx, float gelpad_h, float gelpad_min)
{
    __shared__ float red[8];
    const float* p = hm + (size_t)blockIdx.x * npx;
    float m = __int_as_float(0x7f800000);
    for (int i = threadIdx.x; i < npx; i += 256) m = fminf(m, __ldg(p + i));
    m = warp_min(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        float v = red[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) v = fminf(v, red[w]);
        float d = __fdiv_rn(v, 1000.0f);
        d = __fadd_rn(d, -gelpad_min);
        d = d < 0.0f ? 0.0f : d;
        out[blockIdx.x] = (d <= gelpad_h) ? __fmul_rn(__fadd_rn(gelpad_h, -d), 1000.0f) : 0.0f;
    }
}

cudaError_t launch_indentation_depth_generic(const float* hm, float* out, int N, int npx, float gelpad_h, float gelpad_min,
                                             cudaStream_t s)
{
    indentation_depth_generic_kernel<<<N, 256, 0, s>>>(hm, out, npx, gelpad_h, gelpad_min);
    return cudaGetLastError();
}

} // namespace tx
