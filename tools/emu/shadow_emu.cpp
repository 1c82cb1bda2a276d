// shadow_emu.cpp -- TEST TOOLING: the kernels of csrc/taxim_shadow_kernel.cu compiled for the HOST (tools/emu/include stands in
// for the CUDA runtime) and driven thread by thread, so that tests/test_shadow_cpu.py can compare the kernel source's logic with
// the canonical restatement without a GPU. Not a product path (the product has no CPU path): built only by the test into
// tools/emu/_build/.
#include <cuda_runtime.h>
uint3 blockIdx, threadIdx;
#include "../../tacex_b200/csrc/taxim_shadow_kernel.cu"
#include <vector>

using namespace tx;

template <class K, class... A>
static void run(dim3 grid, int threads, K kernel, A... args)
{
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx)
            for (int t = 0; t < threads; ++t) {
                blockIdx = uint3{bx, by, 0};
                threadIdx = uint3{(unsigned)t, 0, 0};
                kernel(args...);
            }
}

// Same sequence as launch_shadow() in the .cu file. deformed / mask: what the fused kernel would hand over (the test takes them
// from the canonical restatement, with which the fused kernel is bit-identical on the GPU).
extern "C" void emu_shadow(const float* deformed, const unsigned char* mask, const float* gel, const float* poly20 /*[nb][nb][20]*/,
                           const float* bg_hwc, const float* table, const float* fan_cos, const float* fan_sin, int D, int Hn, int S,
                           int F, int nb, const int* dil, float pixmm, float calib_h, float calib_w, float depth_0,
                           float height_precision, float discretize_precision, float step_x, float step_y, const float* taps_sx,
                           int ks_sx, const float* taps_sy, int ks_sy, const float* taps_fx, int ks_fx, const float* taps_fy,
                           int ks_fy, int n, float* rgb)
{
    const size_t total = (size_t)n * 3 * HW;
    std::vector<float> sh(total), t1(total), t2(total);
    ShadowArgs a{};
    a.deformed = deformed; a.mask = mask; a.gel = gel; a.poly = reinterpret_cast<const float4*>(poly20); a.shadow = sh.data();
    a.table = table; a.fan_cos = fan_cos; a.fan_sin = fan_sin; a.D = D; a.Hn = Hn; a.S = S; a.F = F; a.nb = nb;
    for (int i = 0; i < 4; ++i) a.dil[i] = dil[i];
    // the constants exactly as fill_taxim_consts (tx_api.cu) derives them
    a.pixmm = pixmm; a.inv_pixmm = 1.0f / pixmm; a.sy = (float)IMG_H / calib_h; a.sx = (float)IMG_W / calib_w;
    a.fx = calib_w / (float)IMG_W; a.fy = calib_h / (float)IMG_H;
    a.inv_xbin = (float)(1.0 / (0.5 * M_PI / (nb - 1))); a.inv_ybin = (float)(1.0 / (2.0 * M_PI / (nb - 1)));
    a.depth_0 = depth_0; a.height_precision = height_precision; a.discretize_precision = discretize_precision;
    a.step_x = step_x; a.step_y = step_y;
    const dim3 gpix((HW + SH_THREADS - 1) / SH_THREADS, n), gpl((HW + SH_THREADS - 1) / SH_THREADS, 3 * n);
    run(dim3((unsigned)((total + SH_THREADS - 1) / SH_THREADS)), SH_THREADS, shadow_fill_kernel, sh.data(), total, INFINITY);
    run(gpix, SH_THREADS, shadow_cast_kernel, a);
    run(gpix, SH_THREADS, shadow_rawmin_kernel, a);
    run(gpl, SH_THREADS, shadow_blur_h_kernel, (const float*)sh.data(), t1.data(), taps_sx, ks_sx);
    run(gpl, SH_THREADS, shadow_blur_v_kernel<0>, (const float*)t1.data(), t2.data(), taps_sy, ks_sy, bg_hwc);
    run(gpl, SH_THREADS, shadow_blur_h_kernel, (const float*)t2.data(), t1.data(), taps_fx, ks_fx);
    run(gpl, SH_THREADS, shadow_blur_v_kernel<1>, (const float*)t1.data(), rgb, taps_fy, ks_fy, bg_hwc);
}
