// tx_api.cu -- the C ABI of libtacex_b200.so (see include/tacex_b200.h for the contract and reference citations).
#include "tx_kernels.h"
#include <math.h>
#include <new>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

using namespace tx;

struct tx_handle {
    tx_config cfg;
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    tx_counters ctr{};
    bool have_tables = false;
    // device tables
    float4* d_poly = nullptr; // [nb][nb][20]  (generic / shadow kernels)
    int rs_Hi = 0, rs_Wi = 0;          // tx_resize: tables of the last source shape
    int* d_rsa_i = nullptr;
    float* d_rsa_w = nullptr;
    unsigned char* d_patch = nullptr; // [10][10][12][12] marker dot patches (tx_set_marker_patches)
    float4* d_poly128 = nullptr; // [nb][nb][32] the same records padded to one 128-byte line each (fused 240 x 320 kernel)
    float* d_bg = nullptr;    // [H][W][3]
    float* d_gel = nullptr;   // [H][W] or nullptr
    float* d_flat = nullptr;  // [H][W][3] flat-pixel RGB
    // marker grid
    int M = 0;
    std::vector<int32_t> mx, my;
    int* d_mx = nullptr;
    int* d_my = nullptr;
    // FOTS inputs recorded by tx_render
    unsigned* d_aux_sums = nullptr;
    float* d_aux_bmax = nullptr;
    float* d_aux_b = nullptr;
    unsigned char* d_aux_m = nullptr;
    int aux_valid_n = 0;
    // tx_step_host staging
    float* d_hm = nullptr;
    float* d_rgb = nullptr;
    float* d_depth = nullptr;
    float* d_theta = nullptr;
    float* d_traj0 = nullptr;
    int* d_traj_len = nullptr;
    float* d_markers = nullptr;
    int host_cap = 0;
    cudaStream_t s_in = nullptr, s_out = nullptr; // copy streams of the pipelined host path
    std::vector<cudaEvent_t> ev_in, ev_done;
    long long* d_ticks = nullptr; // not owned
    int dbg = 0;
    // camera resolution != tactile resolution (tx_set_camera_resolution)
    int Hc = 0, Wc = 0;
    int *d_rs_x0 = nullptr, *d_rs_y0 = nullptr;
    float2 *d_rs_wx = nullptr, *d_rs_wy = nullptr;
    float* d_up = nullptr;
    int* d_rect = nullptr; // not owned: tx_set_rect_output
    float* mc_rgb = nullptr; // not owned: tx_set_multicast_output
    int* mc_rect = nullptr;
    // shadow branch (tx_upload_shadow_tables / tx_render_shadow)
    tx_shadow_config sh_cfg{};
    bool have_shadow = false;
    float *d_sh_table = nullptr, *d_sh_cos = nullptr, *d_sh_sin = nullptr, *d_sh_taps = nullptr; // taps: [4][TX_MAX_TAPS] sx, sy, fx, fy
    float *d_sh_def = nullptr, *d_sh_img = nullptr, *d_sh_t1 = nullptr, *d_sh_t2 = nullptr;
    unsigned char* d_sh_mask = nullptr;
    int sh_chunk = 0;
    bool generic = false;  // shape / radii other than the specialised 240 x 320 kernel: taxim_generic_kernel
    float* d_taps = nullptr; // generic: [n_blurs][2][TX_MAX_TAPS]
    TaximTaps taps{};        // specialised kernel: this handle's taps, passed as a kernel parameter with every launch
    int n_sm = 148;
};

static std::string g_create_err;

static int fail(tx_handle* h, int code, const std::string& msg)
{
    if (h)
        h->err = msg;
    else
        g_create_err = msg;
    return code;
}
#define TX_CUDA(h, expr)                                                                                              \
    do {                                                                                                              \
        cudaError_t _e = (expr);                                                                                      \
        if (_e != cudaSuccess)                                                                                        \
            return fail((h), TX_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                        \
    } while (0)

static void fill_taxim_consts(const tx_handle* h, TaximArgs& a);

extern "C" int tx_abi_version(void) { return TX_ABI_VERSION; }

// np.linspace(x0, W - x0, cols, dtype=int) (ref: marker_motion.py:58-60): float64 linspace, truncation toward zero
static void marker_axis(double lo, double hi, int n, std::vector<int32_t>& out)
{
    out.resize(n);
    const double step = n > 1 ? (hi - lo) / (double)(n - 1) : 0.0;
    for (int i = 0; i < n; ++i) {
        const double v = (i == n - 1 && n > 1) ? hi : lo + step * i;
        out[i] = (int32_t)v;
    }
}

extern "C" int tx_create(const tx_config* cfg, int device, void* cuda_stream, tx_handle** out)
{
    if (!cfg || !out) return fail(nullptr, TX_ERR_INVALID_ARG, "tx_create: null argument");
    *out = nullptr;
    if (cfg->abi_version != TX_ABI_VERSION) return fail(nullptr, TX_ERR_INVALID_ARG, "tx_create: ABI version mismatch");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, TX_ERR_NO_DEVICE, "tx_create: no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail(nullptr, TX_ERR_INVALID_ARG, "tx_create: bad device index");
    // The 2-CTA kernel is specialised for the GelSight Mini at 240 x 320 (kernel sizes 61,33,17,9,5,3,5); every other shape
    // (the reference's RL tasks render 32 x 24 / 32 x 32 tactile images) runs the arbitrary-resolution kernel.
    static const int want[7] = {61, 33, 17, 9, 5, 3, 5};
    bool generic = cfg->H != IMG_H || cfg->W != IMG_W || cfg->n_blurs != 7;
    if (cfg->n_blurs < 1 || cfg->n_blurs > TX_MAX_BLURS) return fail(nullptr, TX_ERR_INVALID_ARG, "tx_create: bad number of blur levels");
    for (int l = 0; l < cfg->n_blurs; ++l) {
        if (cfg->ksx[l] < 1 || cfg->ksy[l] < 1 || !(cfg->ksx[l] & 1) || !(cfg->ksy[l] & 1) || cfg->ksx[l] > TX_MAX_TAPS ||
            cfg->ksy[l] > TX_MAX_TAPS)
            return fail(nullptr, TX_ERR_INVALID_ARG, "tx_create: kernel sizes must be odd and at most TX_MAX_TAPS");
        if (!generic && (cfg->ksx[l] != want[l] || cfg->ksy[l] != want[l])) generic = true;
    }
    if (generic) {
        if (cfg->H < 3 || cfg->W < 3 || (long long)cfg->H * cfg->W > TXG_MAX_PIXELS)
            return fail(nullptr, TX_ERR_UNSUPPORTED, "tx_create: frames other than 240 x 320 must have 3 <= H, W and at most 19200 pixels");
        for (int l = 0; l < cfg->n_blurs; ++l)
            if ((cfg->ksx[l] - 1) / 2 >= cfg->W || (cfg->ksy[l] - 1) / 2 >= cfg->H)
                return fail(nullptr, TX_ERR_UNSUPPORTED, "tx_create: blur radius must be smaller than the frame (reflect padding)");
        if (cfg->marker_rows * cfg->marker_cols != 0)
            return fail(nullptr, TX_ERR_UNSUPPORTED, "tx_create: the FOTS marker model needs the 240 x 320 kernel");
    }
    if (cfg->max_envs <= 0) return fail(nullptr, TX_ERR_INVALID_ARG, "tx_create: max_envs must be positive");
    const int M = cfg->marker_rows * cfg->marker_cols;
    if (M < 0 || M > TX_MAX_MARKERS) return fail(nullptr, TX_ERR_INVALID_ARG, "tx_create: too many markers");

    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10)
        return fail(nullptr, TX_ERR_NO_DEVICE, "tx_create: an sm_100a (Blackwell) device is required");

    tx_handle* h = new (std::nothrow) tx_handle();
    if (!h) return fail(nullptr, TX_ERR_INVALID_ARG, "tx_create: out of host memory");
    h->cfg = *cfg;
    h->device = device;
    h->stream = (cudaStream_t)cuda_stream;
    h->M = M;
    h->n_sm = prop.multiProcessorCount;
    h->generic = generic;
#define TX_CUDA_C(expr)                                                                                               \
    do {                                                                                                              \
        cudaError_t _e = (expr);                                                                                      \
        if (_e != cudaSuccess) {                                                                                      \
            fail(nullptr, TX_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                           \
            tx_destroy(h);                                                                                            \
            return TX_ERR_CUDA;                                                                                       \
        }                                                                                                             \
    } while (0)
    TX_CUDA_C(cudaSetDevice(device));
    // taps, [blur][0 = x, 1 = y][tap]: per handle (global memory for the generic kernel, a kernel parameter otherwise)
    {
        std::vector<float> t((size_t)TX_MAX_BLURS * 2 * TX_MAX_TAPS, 0.0f);
        for (int l = 0; l < cfg->n_blurs; ++l)
            for (int k = 0; k < TX_MAX_TAPS; ++k) {
                t[((size_t)l * 2 + 0) * TX_MAX_TAPS + k] = k < cfg->ksx[l] ? cfg->taps_x[l][k] : 0.0f;
                t[((size_t)l * 2 + 1) * TX_MAX_TAPS + k] = k < cfg->ksy[l] ? cfg->taps_y[l][k] : 0.0f;
            }
        if (generic) { // per-handle taps in global memory (the __constant__ copy belongs to the specialised kernel)
            TX_CUDA_C(cudaMalloc(&h->d_taps, t.size() * sizeof(float)));
            TX_CUDA_C(cudaMemcpy(h->d_taps, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice));
        } else {
            for (int l = 0; l < TX_FUSED_BLURS; ++l)
                for (int ax = 0; ax < 2; ++ax)
                    memcpy(h->taps.t[l][ax], &t[((size_t)l * 2 + ax) * TX_MAX_TAPS], sizeof(float) * TX_MAX_TAPS);
        }
    }
    if (M > 0) {
        std::vector<int32_t> ax, ay;
        marker_axis(cfg->marker_x0, (double)cfg->W - cfg->marker_x0, cfg->marker_cols, ax);
        marker_axis(cfg->marker_y0, (double)cfg->H - cfg->marker_y0, cfg->marker_rows, ay);
        h->mx.resize(M);
        h->my.resize(M);
        for (int r = 0; r < cfg->marker_rows; ++r)
            for (int c = 0; c < cfg->marker_cols; ++c) {
                h->mx[r * cfg->marker_cols + c] = ax[c];
                h->my[r * cfg->marker_cols + c] = ay[r];
            }
        TX_CUDA_C(cudaMalloc(&h->d_mx, sizeof(int) * M));
        TX_CUDA_C(cudaMalloc(&h->d_my, sizeof(int) * M));
        TX_CUDA_C(cudaMemcpy(h->d_mx, h->mx.data(), sizeof(int) * M, cudaMemcpyHostToDevice));
        TX_CUDA_C(cudaMemcpy(h->d_my, h->my.data(), sizeof(int) * M, cudaMemcpyHostToDevice));
        const size_t N = (size_t)cfg->max_envs;
        TX_CUDA_C(cudaMalloc(&h->d_aux_sums, sizeof(unsigned) * 8 * N));
        TX_CUDA_C(cudaMalloc(&h->d_aux_bmax, sizeof(float) * 2 * N));
        TX_CUDA_C(cudaMalloc(&h->d_aux_b, sizeof(float) * M * N));
        TX_CUDA_C(cudaMalloc(&h->d_aux_m, M * N));
        TX_CUDA_C(cudaMemset(h->d_aux_m, 0, M * N));
        TX_CUDA_C(cudaMemset(h->d_aux_b, 0, sizeof(float) * M * N));
    }
#undef TX_CUDA_C
    *out = h;
    return TX_OK;
}

extern "C" void tx_destroy(tx_handle* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaFree(h->d_taps);
    cudaFree(h->d_sh_table); cudaFree(h->d_sh_cos); cudaFree(h->d_sh_sin); cudaFree(h->d_sh_taps); cudaFree(h->d_sh_def);
    cudaFree(h->d_sh_img); cudaFree(h->d_sh_t1); cudaFree(h->d_sh_t2); cudaFree(h->d_sh_mask);
    cudaFree(h->d_patch); cudaFree(h->d_rsa_i); cudaFree(h->d_rsa_w);
    cudaFree(h->d_poly); cudaFree(h->d_poly128); cudaFree(h->d_bg); cudaFree(h->d_gel); cudaFree(h->d_flat); cudaFree(h->d_mx); cudaFree(h->d_my);
    cudaFree(h->d_aux_sums); cudaFree(h->d_aux_bmax); cudaFree(h->d_aux_b); cudaFree(h->d_aux_m);
    cudaFree(h->d_rs_x0); cudaFree(h->d_rs_y0); cudaFree(h->d_rs_wx); cudaFree(h->d_rs_wy); cudaFree(h->d_up);
    cudaFree(h->d_hm); cudaFree(h->d_rgb); cudaFree(h->d_depth); cudaFree(h->d_theta); cudaFree(h->d_traj0);
    cudaFree(h->d_traj_len); cudaFree(h->d_markers);
    for (auto e : h->ev_in) cudaEventDestroy(e);
    for (auto e : h->ev_done) cudaEventDestroy(e);
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    delete h;
}

extern "C" const char* tx_last_error(const tx_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

extern "C" int tx_get_counters(const tx_handle* h, tx_counters* out)
{
    if (!h || !out) return TX_ERR_INVALID_ARG;
    *out = h->ctr;
    return TX_OK;
}

extern "C" int tx_upload_tables(tx_handle* h, const float* poly_grad, const float* background, const float* gel_map)
{
    if (!h || !poly_grad || !background) return fail(h, TX_ERR_INVALID_ARG, "tx_upload_tables: null argument");
    TX_CUDA(h, cudaSetDevice(h->device));
    const int nb = h->cfg.num_bins, H = h->cfg.H, W = h->cfg.W;
    // poly [3][nb][nb][6] -> [nb][nb][3*6 padded to 20]: one 80-byte record per bin
    std::vector<float> poly((size_t)nb * nb * 20, 0.0f);
    for (int c = 0; c < 3; ++c)
        for (int a = 0; a < nb; ++a)
            for (int b = 0; b < nb; ++b)
                for (int k = 0; k < 6; ++k)
                    poly[((size_t)a * nb + b) * 20 + c * 6 + k] = poly_grad[(((size_t)c * nb + a) * nb + b) * 6 + k];
    // background [3][H][W] -> [H][W][3] (the output layout)
    std::vector<float> bg((size_t)H * W * 3);
    for (int c = 0; c < 3; ++c)
        for (int i = 0; i < H * W; ++i) bg[(size_t)i * 3 + c] = background[(size_t)c * H * W + i];
    if (!h->d_poly) TX_CUDA(h, cudaMalloc(&h->d_poly, poly.size() * sizeof(float)));
    if (!h->d_bg) TX_CUDA(h, cudaMalloc(&h->d_bg, bg.size() * sizeof(float)));
    TX_CUDA(h, cudaMemcpy(h->d_poly, poly.data(), poly.size() * sizeof(float), cudaMemcpyHostToDevice));
    if (!h->generic) {
        std::vector<float> p128((size_t)nb * nb * 32, 0.0f);
        for (size_t r = 0; r < (size_t)nb * nb; ++r) memcpy(&p128[r * 32], &poly[r * 20], 20 * sizeof(float));
        if (!h->d_poly128) TX_CUDA(h, cudaMalloc(&h->d_poly128, p128.size() * sizeof(float)));
        TX_CUDA(h, cudaMemcpy(h->d_poly128, p128.data(), p128.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    TX_CUDA(h, cudaMemcpy(h->d_bg, bg.data(), bg.size() * sizeof(float), cudaMemcpyHostToDevice));
    if (gel_map) {
        if (!h->d_gel) TX_CUDA(h, cudaMalloc(&h->d_gel, sizeof(float) * H * W));
        TX_CUDA(h, cudaMemcpy(h->d_gel, gel_map, sizeof(float) * H * W, cudaMemcpyHostToDevice));
    } else if (h->d_gel) {
        cudaFree(h->d_gel);
        h->d_gel = nullptr;
    }
    if (!h->generic) {
        if (!h->d_flat) TX_CUDA(h, cudaMalloc(&h->d_flat, sizeof(float) * H * W * 3));
        TaximArgs a{};
        fill_taxim_consts(h, a);
        a.poly = taxim_record_f4() == 8 ? h->d_poly128 : h->d_poly;
        TX_CUDA(h, launch_flat_rgb(a, h->d_flat, h->stream));
        TX_CUDA(h, cudaStreamSynchronize(h->stream));
        h->ctr.kernels_launched++;
    }
    h->have_tables = true;
    return TX_OK;
}

static void fill_taxim_consts(const tx_handle* h, TaximArgs& a)
{
    const tx_config& c = h->cfg;
    a.gel = h->d_gel;
    a.poly = h->d_poly;
    a.bg_hwc = h->d_bg;
    a.flat_rgb = h->d_flat;
    a.inv_pixmm = 1.0f / c.pixmm;
    a.sy = (float)c.H / c.calib_h;
    a.sx = (float)c.W / c.calib_w;
    a.fx = c.calib_w / (float)c.W;
    a.fy = c.calib_h / (float)c.H;
    a.contact_scale = c.contact_scale;
    a.gelpad_h = c.gelpad_height_m;
    a.gelpad_min = c.gelpad_to_cam_min_m;
    a.inv_xbin = (float)(1.0 / (0.5 * M_PI / (c.num_bins - 1)));
    a.inv_ybin = (float)(1.0 / (2.0 * M_PI / (c.num_bins - 1)));
    a.nb = c.num_bins;
}

extern "C" int tx_indentation_depth(tx_handle* h, const float* height_mm, int N, float* depth_mm)
{
    if (!h || !height_mm || !depth_mm || N < 0) return fail(h, TX_ERR_INVALID_ARG, "tx_indentation_depth: bad argument");
    if (N == 0) return TX_OK;
    TX_CUDA(h, cudaSetDevice(h->device));
    if (h->generic)
        TX_CUDA(h, launch_indentation_depth_generic(height_mm, depth_mm, N, h->cfg.H * h->cfg.W, h->cfg.gelpad_height_m,
                                                    h->cfg.gelpad_to_cam_min_m, h->stream));
    else
        TX_CUDA(h, launch_indentation_depth(height_mm, depth_mm, N, h->cfg.gelpad_height_m, h->cfg.gelpad_to_cam_min_m,
                                            h->stream));
    h->ctr.depth_calls++;
    h->ctr.kernels_launched++;
    return TX_OK;
}

extern "C" int tx_indentation_depth_frames(tx_handle* h, const float* frames_mm, int N, int pixels_per_frame, float* depth_mm)
{
    if (!h || !frames_mm || !depth_mm || N < 0 || pixels_per_frame <= 0)
        return fail(h, TX_ERR_INVALID_ARG, "tx_indentation_depth_frames: bad argument");
    if (N == 0) return TX_OK;
    TX_CUDA(h, cudaSetDevice(h->device));
    TX_CUDA(h, launch_indentation_depth_generic(frames_mm, depth_mm, N, pixels_per_frame, h->cfg.gelpad_height_m,
                                                h->cfg.gelpad_to_cam_min_m, h->stream));
    h->ctr.depth_calls++;
    h->ctr.kernels_launched++;
    return TX_OK;
}

static int render_impl(tx_handle* h, const float* height_mm, const float* press_mm, int N, float* rgb, float* depth_out,
                       float* deformed, uint8_t* mask, int input_is_depth, float clip_max_m, float* hm_out, int lowres = 0,
                       int env0 = 0);

// first tap + two weights per output index of torch's antialiased bilinear filter for up-sampling (support 1): float32
// arithmetic like aten (UpSampleKernel.cpp, _compute_indices_min_size_weights_aa); returns false if more than two taps.
static bool resize_table(int in_size, int out_size, std::vector<int>& first, std::vector<float2>& w)
{
    const float scale = (float)in_size / (float)out_size;
    if (scale > 1.0f) return false; // down-sampling needs more taps than the fused load stage implements
    first.resize(out_size);
    w.resize(out_size);
    for (int i = 0; i < out_size; ++i) {
        const float center = scale * ((float)i + 0.5f);
        long xmin = (long)(center - 1.0f + 0.5f);
        if (xmin < 0) xmin = 0;
        long xmax = (long)(center + 1.0f + 0.5f);
        if (xmax > in_size) xmax = in_size;
        const int n = (int)(xmax - xmin);
        if (n < 1 || n > 2) return false;
        float wd[2] = {0.0f, 0.0f}, tot = 0.0f;
        for (int j = 0; j < n; ++j) {
            float x = ((float)(j + xmin) - center + 0.5f);
            if (x < 0) x = -x;
            wd[j] = x < 1.0f ? 1.0f - x : 0.0f;
            tot += wd[j];
        }
        first[i] = (int)xmin;
        w[i] = make_float2(wd[0] / tot, n > 1 ? wd[1] / tot : 0.0f);
    }
    return true;
}

extern "C" int tx_set_camera_resolution(tx_handle* h, int Hc, int Wc)
{
    if (!h || Hc <= 0 || Wc <= 0) return fail(h, TX_ERR_INVALID_ARG, "tx_set_camera_resolution: bad argument");
    if (h->generic) return fail(h, TX_ERR_UNSUPPORTED, "tx_set_camera_resolution: the fused resize belongs to the 240 x 320 kernel");
    if (Hc > h->cfg.H || Wc > h->cfg.W || Hc * Wc > taxim_lowres_max_pixels() || (Hc == h->cfg.H && Wc == h->cfg.W))
        return fail(h, TX_ERR_UNSUPPORTED, "tx_set_camera_resolution: only up-sampling of frames of at most 9600 pixels is fused");
    std::vector<int> fx, fy;
    std::vector<float2> wx, wy;
    if (!resize_table(Wc, h->cfg.W, fx, wx) || !resize_table(Hc, h->cfg.H, fy, wy))
        return fail(h, TX_ERR_UNSUPPORTED, "tx_set_camera_resolution: unsupported scale");
    TX_CUDA(h, cudaSetDevice(h->device));
    if (!h->d_rs_x0) {
        TX_CUDA(h, cudaMalloc(&h->d_rs_x0, sizeof(int) * h->cfg.W));
        TX_CUDA(h, cudaMalloc(&h->d_rs_y0, sizeof(int) * h->cfg.H));
        TX_CUDA(h, cudaMalloc(&h->d_rs_wx, sizeof(float2) * h->cfg.W));
        TX_CUDA(h, cudaMalloc(&h->d_rs_wy, sizeof(float2) * h->cfg.H));
        TX_CUDA(h, cudaMalloc(&h->d_up, sizeof(float) * (size_t)h->cfg.max_envs * h->cfg.H * h->cfg.W));
    }
    TX_CUDA(h, cudaStreamSynchronize(h->stream));
    TX_CUDA(h, cudaMemcpy(h->d_rs_x0, fx.data(), sizeof(int) * fx.size(), cudaMemcpyHostToDevice));
    TX_CUDA(h, cudaMemcpy(h->d_rs_y0, fy.data(), sizeof(int) * fy.size(), cudaMemcpyHostToDevice));
    TX_CUDA(h, cudaMemcpy(h->d_rs_wx, wx.data(), sizeof(float2) * wx.size(), cudaMemcpyHostToDevice));
    TX_CUDA(h, cudaMemcpy(h->d_rs_wy, wy.data(), sizeof(float2) * wy.size(), cudaMemcpyHostToDevice));
    h->Hc = Hc;
    h->Wc = Wc;
    return TX_OK;
}

extern "C" int tx_render_camera(tx_handle* h, const float* frames, int is_depth, float clip_max_m, const float* press_mm, int N,
                                float* rgb, float* depth_out, float* deformed, uint8_t* mask)
{
    if (!h) return TX_ERR_INVALID_ARG;
    if (h->Hc <= 0) return fail(h, TX_ERR_STATE, "tx_render_camera: call tx_set_camera_resolution first");
    if (is_depth && !(clip_max_m > 0.0f)) return fail(h, TX_ERR_INVALID_ARG, "tx_render_camera: clip_max_m must be positive");
    return render_impl(h, frames, press_mm, N, rgb, depth_out, deformed, mask, is_depth ? 1 : 0, clip_max_m, nullptr, 1);
}

extern "C" int tx_render(tx_handle* h, const float* height_mm, const float* press_mm, int N, float* rgb,
                         float* depth_out, float* deformed, uint8_t* mask)
{
    return render_impl(h, height_mm, press_mm, N, rgb, depth_out, deformed, mask, 0, 0.0f, nullptr);
}

extern "C" int tx_render_depth(tx_handle* h, const float* depth_m, float clip_max_m, int N, float* rgb, float* depth_out,
                               float* height_mm_out)
{
    if (!(clip_max_m > 0.0f)) return fail(h, TX_ERR_INVALID_ARG, "tx_render_depth: clip_max_m must be positive");
    if (height_mm_out && ((uintptr_t)height_mm_out & 15u))
        return fail(h, TX_ERR_INVALID_ARG, "tx_render_depth: device buffers must be 16-byte aligned");
    return render_impl(h, depth_m, nullptr, N, rgb, depth_out, nullptr, nullptr, 1, clip_max_m, height_mm_out);
}

static int render_impl(tx_handle* h, const float* height_mm, const float* press_mm, int N, float* rgb, float* depth_out,
                       float* deformed, uint8_t* mask, int input_is_depth, float clip_max_m, float* hm_out, int lowres, int env0)
{
    // env0: the frames are envs [env0, env0 + N) of a batch rendered in chunks (tx_render_shadow, tx_step_host): the per-env
    // records of the FOTS model and the gather rectangles land at their batch position, so the whole batch stays valid
    if (!h || !height_mm || !rgb || N < 0) return fail(h, TX_ERR_INVALID_ARG, "tx_render: bad argument");
    if (!h->have_tables) return fail(h, TX_ERR_NO_TABLES, "tx_render: call tx_upload_tables first");
    if (N > h->cfg.max_envs || env0 < 0 || env0 + N > h->cfg.max_envs) return fail(h, TX_ERR_INVALID_ARG, "tx_render: N exceeds max_envs");
    if (!h->generic &&
        ((!lowres && ((uintptr_t)height_mm & 15u)) || ((uintptr_t)rgb & 15u) || (deformed && ((uintptr_t)deformed & 15u)) ||
         (mask && ((uintptr_t)mask & 3u))))
        return fail(h, TX_ERR_INVALID_ARG, "tx_render: device buffers must be 16-byte aligned");
    if (N == 0) return TX_OK;
    TX_CUDA(h, cudaSetDevice(h->device));
    if (h->generic) {
        TaximArgs c{};
        fill_taxim_consts(h, c);
        TaximGenericArgs g{};
        g.hm = height_mm; g.press_in = press_mm; g.gel = c.gel; g.poly = c.poly; g.bg_hwc = c.bg_hwc; g.taps = h->d_taps;
        g.rgb = rgb; g.depth_out = depth_out; g.deformed_out = deformed; g.mask_out = mask; g.hm_out = hm_out;
        g.H = h->cfg.H; g.W = h->cfg.W; g.n_blurs = h->cfg.n_blurs; g.nb = c.nb; g.input_is_depth = input_is_depth;
        for (int l = 0; l < h->cfg.n_blurs; ++l) { g.ksx[l] = h->cfg.ksx[l]; g.ksy[l] = h->cfg.ksy[l]; }
        g.clip_max_m = clip_max_m; g.inv_pixmm = c.inv_pixmm; g.sx = c.sx; g.sy = c.sy; g.fx = c.fx; g.fy = c.fy;
        g.contact_scale = c.contact_scale; g.gelpad_h = c.gelpad_h; g.gelpad_min = c.gelpad_min;
        g.inv_xbin = c.inv_xbin; g.inv_ybin = c.inv_ybin;
        TX_CUDA(h, launch_taxim_generic(g, N, h->stream));
        h->aux_valid_n = 0;
        h->ctr.render_calls++;
        h->ctr.frames_rendered += (uint64_t)N;
        h->ctr.kernels_launched++;
        return TX_OK;
    }
    TaximArgs a{};
    a.hm = height_mm;
    a.press_in = press_mm;
    a.input_is_depth = input_is_depth;
    a.clip_max_m = clip_max_m;
    a.hm_out = hm_out;
    if (lowres) {
        a.Hc = h->Hc; a.Wc = h->Wc;
        a.rs_x0 = h->d_rs_x0; a.rs_y0 = h->d_rs_y0; a.rs_wx = h->d_rs_wx; a.rs_wy = h->d_rs_wy;
        a.up_scratch = h->d_up;
    }
    fill_taxim_consts(h, a);
    a.poly = taxim_record_f4() == 8 ? h->d_poly128 : h->d_poly; // one 128-byte line per record
    a.rgb = rgb;
    a.depth_out = depth_out;
    a.deformed_out = deformed;
    a.mask_out = mask;
    if (h->M > 0) {
        a.aux_sums = h->d_aux_sums + (size_t)env0 * 8;
        a.aux_bmax = h->d_aux_bmax + (size_t)env0 * 2;
        a.aux_b = h->d_aux_b + (size_t)env0 * h->M;
        a.aux_m = h->d_aux_m + (size_t)env0 * h->M;
        a.mk_x = h->d_mx;
        a.mk_y = h->d_my;
        a.M = h->M;
    }
    a.ticks = h->d_ticks;
    a.dbg = h->dbg;
    a.rect_out = h->d_rect ? h->d_rect + (size_t)env0 * 8 : nullptr;
    if (h->mc_rgb && h->d_rect && env0 == 0) { a.rgb_mc = h->mc_rgb; a.rect_mc = h->mc_rect; }
    TX_CUDA(h, launch_taxim(a, h->taps, N, h->stream));
    h->aux_valid_n = h->M > 0 ? env0 + N : 0;
    h->ctr.render_calls++;
    h->ctr.frames_rendered += (uint64_t)N;
    h->ctr.kernels_launched++;
    return TX_OK;
}

extern "C" int tx_upload_shadow_tables(tx_handle* h, const tx_shadow_config* c, const float* table, const float* fan_cos,
                                       const float* fan_sin)
{
    if (!h || !c || !table || !fan_cos || !fan_sin) return fail(h, TX_ERR_INVALID_ARG, "tx_upload_shadow_tables: null argument");
    if (h->generic) return fail(h, TX_ERR_UNSUPPORTED, "tx_upload_shadow_tables: the shadow branch belongs to the 240 x 320 kernel");
    if (c->D < 1 || c->Hn < 2 || c->S < 1 || c->F < 1 || c->ks_sx < 1 || c->ks_sy < 1 || !(c->ks_sx & 1) || !(c->ks_sy & 1) ||
        c->ks_sx > TX_MAX_TAPS || c->ks_sy > TX_MAX_TAPS || !(c->height_precision > 0.0f) || !(c->discretize_precision > 0.0f))
        return fail(h, TX_ERR_INVALID_ARG, "tx_upload_shadow_tables: bad configuration");
    for (int i = 0; i < 4; ++i)
        if (c->dil[i] < 1 || c->dil[i] > 16) return fail(h, TX_ERR_INVALID_ARG, "tx_upload_shadow_tables: bad dilation kernel");
    TX_CUDA(h, cudaSetDevice(h->device));
    TX_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(h->d_sh_table); cudaFree(h->d_sh_cos); cudaFree(h->d_sh_sin); cudaFree(h->d_sh_taps);
    h->d_sh_table = h->d_sh_cos = h->d_sh_sin = h->d_sh_taps = nullptr;
    h->have_shadow = false;
    const size_t nt = (size_t)3 * c->D * c->Hn * c->S, nf = (size_t)c->D * c->F;
    TX_CUDA(h, cudaMalloc(&h->d_sh_table, nt * sizeof(float)));
    TX_CUDA(h, cudaMalloc(&h->d_sh_cos, nf * sizeof(float)));
    TX_CUDA(h, cudaMalloc(&h->d_sh_sin, nf * sizeof(float)));
    TX_CUDA(h, cudaMalloc(&h->d_sh_taps, sizeof(float) * 4 * TX_MAX_TAPS));
    TX_CUDA(h, cudaMemcpy(h->d_sh_table, table, nt * sizeof(float), cudaMemcpyHostToDevice));
    TX_CUDA(h, cudaMemcpy(h->d_sh_cos, fan_cos, nf * sizeof(float), cudaMemcpyHostToDevice));
    TX_CUDA(h, cudaMemcpy(h->d_sh_sin, fan_sin, nf * sizeof(float), cudaMemcpyHostToDevice));
    std::vector<float> t((size_t)4 * TX_MAX_TAPS, 0.0f);
    const int lf = h->cfg.n_blurs - 1; // the final blur of the deformation pyramid is also the last blur of the shadow branch
    for (int k = 0; k < TX_MAX_TAPS; ++k) {
        t[0 * TX_MAX_TAPS + k] = k < c->ks_sx ? c->taps_sx[k] : 0.0f;
        t[1 * TX_MAX_TAPS + k] = k < c->ks_sy ? c->taps_sy[k] : 0.0f;
        t[2 * TX_MAX_TAPS + k] = k < h->cfg.ksx[lf] ? h->cfg.taps_x[lf][k] : 0.0f;
        t[3 * TX_MAX_TAPS + k] = k < h->cfg.ksy[lf] ? h->cfg.taps_y[lf][k] : 0.0f;
    }
    TX_CUDA(h, cudaMemcpy(h->d_sh_taps, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice));
    h->sh_cfg = *c;
    h->have_shadow = true;
    return TX_OK;
}

extern "C" int tx_render_shadow(tx_handle* h, const float* height_mm, const float* press_mm, int N, float* rgb, float* depth_out)
{
    if (!h || !height_mm || !rgb || N < 0) return fail(h, TX_ERR_INVALID_ARG, "tx_render_shadow: bad argument");
    if (h->generic) return fail(h, TX_ERR_UNSUPPORTED, "tx_render_shadow: the shadow branch belongs to the 240 x 320 kernel");
    if (!h->have_shadow) return fail(h, TX_ERR_NO_TABLES, "tx_render_shadow: call tx_upload_shadow_tables first");
    if (N > h->cfg.max_envs) return fail(h, TX_ERR_INVALID_ARG, "tx_render_shadow: N exceeds max_envs");
    if (N == 0) return TX_OK;
    TX_CUDA(h, cudaSetDevice(h->device));
    const size_t HWp = (size_t)IMG_H * IMG_W;
    if (!h->d_sh_def) { // scratch for one chunk of frames: deformed gel, mask, three [n][3][H][W] planes
        const int chunk = h->cfg.max_envs < 256 ? h->cfg.max_envs : 256;
        TX_CUDA(h, cudaMalloc(&h->d_sh_def, sizeof(float) * HWp * chunk));
        TX_CUDA(h, cudaMalloc(&h->d_sh_mask, HWp * chunk));
        TX_CUDA(h, cudaMalloc(&h->d_sh_img, sizeof(float) * 3 * HWp * chunk));
        TX_CUDA(h, cudaMalloc(&h->d_sh_t1, sizeof(float) * 3 * HWp * chunk));
        TX_CUDA(h, cudaMalloc(&h->d_sh_t2, sizeof(float) * 3 * HWp * chunk));
        h->sh_chunk = chunk;
    }
    TaximArgs c{};
    fill_taxim_consts(h, c);
    const tx_shadow_config& sc = h->sh_cfg;
    const int lf = h->cfg.n_blurs - 1;
    for (int n0 = 0; n0 < N; n0 += h->sh_chunk) {
        const int n = N - n0 < h->sh_chunk ? N - n0 : h->sh_chunk;
        int rc = render_impl(h, height_mm + HWp * n0, press_mm ? press_mm + n0 : nullptr, n, rgb + HWp * 3 * n0,
                             depth_out ? depth_out + n0 : nullptr, h->d_sh_def, h->d_sh_mask, 0, 0.0f, nullptr, 0, n0);
        if (rc != TX_OK) return rc;
        ShadowArgs a{};
        a.deformed = h->d_sh_def; a.mask = h->d_sh_mask; a.gel = c.gel; a.poly = c.poly; a.shadow = h->d_sh_img;
        a.table = h->d_sh_table; a.fan_cos = h->d_sh_cos; a.fan_sin = h->d_sh_sin;
        a.D = sc.D; a.Hn = sc.Hn; a.S = sc.S; a.F = sc.F; a.nb = c.nb;
        for (int i = 0; i < 4; ++i) a.dil[i] = sc.dil[i];
        a.pixmm = h->cfg.pixmm; a.inv_pixmm = c.inv_pixmm; a.sx = c.sx; a.sy = c.sy; a.fx = c.fx; a.fy = c.fy;
        a.inv_xbin = c.inv_xbin; a.inv_ybin = c.inv_ybin;
        a.depth_0 = sc.depth_0; a.height_precision = sc.height_precision; a.discretize_precision = sc.discretize_precision;
        a.step_x = sc.step_x; a.step_y = sc.step_y;
        TX_CUDA(h, launch_shadow(a, n, h->d_sh_t1, h->d_sh_t2, rgb + HWp * 3 * n0, c.bg_hwc, h->d_sh_taps, sc.ks_sx,
                                 h->d_sh_taps + TX_MAX_TAPS, sc.ks_sy, h->d_sh_taps + 2 * TX_MAX_TAPS, h->cfg.ksx[lf],
                                 h->d_sh_taps + 3 * TX_MAX_TAPS, h->cfg.ksy[lf], h->stream));
        h->ctr.kernels_launched += 7;
    }
    return TX_OK;
}

extern "C" int tx_set_rect_output(tx_handle* h, int32_t* rect)
{
    if (!h) return TX_ERR_INVALID_ARG;
    if (rect && ((uintptr_t)rect & 15u)) return fail(h, TX_ERR_INVALID_ARG, "tx_set_rect_output: buffer must be 16-byte aligned");
    h->d_rect = rect;
    return TX_OK;
}

extern "C" int tx_set_multicast_output(tx_handle* h, float* mc_rgb, int32_t* mc_rect)
{
    if (!h) return TX_ERR_INVALID_ARG;
    if ((mc_rgb != nullptr) != (mc_rect != nullptr)) return fail(h, TX_ERR_INVALID_ARG, "tx_set_multicast_output: both pointers or none");
    if (mc_rgb && (((uintptr_t)mc_rgb & 15u) || ((uintptr_t)mc_rect & 15u)))
        return fail(h, TX_ERR_INVALID_ARG, "tx_set_multicast_output: pointers must be 16-byte aligned");
    if (h->generic && mc_rgb) return fail(h, TX_ERR_UNSUPPORTED, "tx_set_multicast_output: belongs to the 240 x 320 kernel");
    h->mc_rgb = mc_rgb;
    h->mc_rect = mc_rect;
    return TX_OK;
}

extern "C" int tx_obs_push(tx_handle* h, const float* rgb_local, const int32_t* rect_local, int N, int n_peers,
                           float* const* peer_rgb, int32_t* const* peer_rect, float* mc_rgb, int32_t* mc_rect, void* cuda_stream)
{
    if (!h || !rgb_local || !rect_local || N < 0 || n_peers < 0 || n_peers > TX_MAX_PEERS || (n_peers > 0 && (!peer_rgb || !peer_rect)))
        return fail(h, TX_ERR_INVALID_ARG, "tx_obs_push: bad argument");
    if (N == 0 || n_peers == 0) return TX_OK;
    TX_CUDA(h, cudaSetDevice(h->device));
    ObsPushArgs a{};
    a.rgb_local = rgb_local; a.rect_local = rect_local; a.N = N; a.n_peers = n_peers;
    if ((mc_rgb != nullptr) != (mc_rect != nullptr)) return fail(h, TX_ERR_INVALID_ARG, "tx_obs_push: both multicast pointers or none");
    a.mc_rgb = mc_rgb; a.mc_rect = mc_rect;
    for (int p = 0; p < n_peers; ++p) {
        if (!peer_rgb[p] || !peer_rect[p]) return fail(h, TX_ERR_INVALID_ARG, "tx_obs_push: null peer pointer");
        a.peer_rgb[p] = peer_rgb[p];
        a.peer_rect[p] = peer_rect[p];
    }
    // The stores are posted writes bounded by the NVLink egress (900 GB/s per direction), not by the SMs: a few dozen CTAs keep
    // the links busy, and every SM that hosts one is lost to the render kernel of the next step (its 128-register threads leave
    // no room for a second CTA), so the grid stays small (64 CTAs). TX_OBS_PUSH_CTAS overrides it for experiments.
    int cap = 64; // measured: 4 GPUs 1.41 M (32) vs 1.43 M (64) frames/s, 8 GPUs 1.31 M vs 1.33 M; 16 CTAs starve the links
    if (const char* e = getenv("TX_OBS_PUSH_CTAS")) cap = atoi(e) > 0 ? atoi(e) : cap;
    const int grid = 2 * N < cap ? 2 * N : cap;
    TX_CUDA(h, launch_obs_push(a, grid, (cudaStream_t)cuda_stream));
    h->ctr.kernels_launched++;
    return TX_OK;
}

extern "C" int tx_obs_fill(tx_handle* h, float* rgb_all, const int32_t* rect_all, int32_t* prev_rect, int N_total, int skip_lo,
                           int skip_hi, void* cuda_stream)
{
    if (!h || !rgb_all || !rect_all || N_total < 0) return fail(h, TX_ERR_INVALID_ARG, "tx_obs_fill: bad argument");
    if (!h->have_tables) return fail(h, TX_ERR_NO_TABLES, "tx_obs_fill: call tx_upload_tables first");
    if (N_total == 0) return TX_OK;
    TX_CUDA(h, cudaSetDevice(h->device));
    ObsFillArgs a{};
    a.rgb_all = rgb_all; a.rect_all = rect_all; a.prev_rect = prev_rect; a.flat_rgb = h->d_flat; a.N_total = N_total; a.skip_lo = skip_lo; a.skip_hi = skip_hi;
    const int grid = 2 * N_total < 8 * h->n_sm ? 2 * N_total : 8 * h->n_sm;
    TX_CUDA(h, launch_obs_fill(a, grid, (cudaStream_t)cuda_stream));
    h->ctr.kernels_launched++;
    return TX_OK;
}

extern "C" int tx_fots_markers(tx_handle* h, const float* press_mm, const float* theta, int N, float* traj0,
                               int32_t* traj_len, float* markers)
{
    if (!h || !press_mm || !theta || !traj0 || !traj_len || !markers || N < 0)
        return fail(h, TX_ERR_INVALID_ARG, "tx_fots_markers: bad argument");
    if (h->M <= 0) return fail(h, TX_ERR_STATE, "tx_fots_markers: handle has no marker grid");
    if (N != h->aux_valid_n) return fail(h, TX_ERR_STATE, "tx_fots_markers: call tx_render on the same batch first");
    if (N == 0) return TX_OK;
    TX_CUDA(h, cudaSetDevice(h->device));
    FotsArgs f{};
    f.aux_sums = h->d_aux_sums;
    f.aux_bmax = h->d_aux_bmax;
    f.aux_b = h->d_aux_b;
    f.aux_m = h->d_aux_m;
    f.mk_x = h->d_mx;
    f.mk_y = h->d_my;
    f.press = press_mm;
    f.theta = theta;
    f.traj0 = traj0;
    f.traj_len = traj_len;
    f.markers = markers;
    f.M = h->M;
    f.rows = h->cfg.marker_rows;
    f.cols = h->cfg.marker_cols;
    f.lamb0 = h->cfg.fots_lambda[0];
    f.lamb1 = h->cfg.fots_lambda[1];
    f.lamb2 = h->cfg.fots_lambda[2];
    f.mm2pix = h->cfg.mm2pix;
    f.shear_max = h->cfg.shear_max_px;
    f.theta_max = h->cfg.theta_max_rad;
    TX_CUDA(h, launch_fots(f, N, h->stream));
    h->ctr.fots_calls++;
    h->ctr.kernels_launched++;
    return TX_OK;
}

// aten's antialias weights for one axis (UpSampleKernel.cpp, _compute_indices_min_size_weights_aa; float32 arithmetic), any scale
// with at most TX_RS_TAPS taps (down-sampling by up to 3.5 x)
static bool resize_table_aa(int in_size, int out_size, std::vector<int>& first, std::vector<int>& count, std::vector<float>& w)
{
    const float scale = (float)in_size / (float)out_size;
    const float support = scale >= 1.0f ? scale : 1.0f;
    const float invscale = scale >= 1.0f ? 1.0f / scale : 1.0f;
    first.resize(out_size); count.resize(out_size); w.assign((size_t)out_size * TX_RS_TAPS, 0.0f);
    for (int i = 0; i < out_size; ++i) {
        const float center = scale * ((float)i + 0.5f);
        long xmin = (long)(center - support + 0.5f);
        if (xmin < 0) xmin = 0;
        long xmax = (long)(center + support + 0.5f);
        if (xmax > in_size) xmax = in_size;
        const int n = (int)(xmax - xmin);
        if (n < 1 || n > TX_RS_TAPS) return false;
        float wd[TX_RS_TAPS], tot = 0.0f;
        for (int j = 0; j < n; ++j) {
            float x = ((float)(j + xmin) - center + 0.5f) * invscale;
            if (x < 0) x = -x;
            wd[j] = x < 1.0f ? 1.0f - x : 0.0f;
            tot += wd[j];
        }
        first[i] = (int)xmin;
        count[i] = n;
        for (int j = 0; j < n; ++j) w[(size_t)i * TX_RS_TAPS + j] = wd[j] / tot;
    }
    return true;
}

extern "C" int tx_resize(tx_handle* h, const float* src, int N, int Hi, int Wi, float* dst)
{
    if (!h || !src || !dst || N < 0 || Hi <= 0 || Wi <= 0) return fail(h, TX_ERR_INVALID_ARG, "tx_resize: bad argument");
    if (N == 0) return TX_OK;
    TX_CUDA(h, cudaSetDevice(h->device));
    const int Ho = h->cfg.H, Wo = h->cfg.W;
    if (h->rs_Hi != Hi || h->rs_Wi != Wi) {
        std::vector<int> fx, cx, fy, cy;
        std::vector<float> wx, wy;
        if (!resize_table_aa(Wi, Wo, fx, cx, wx) || !resize_table_aa(Hi, Ho, fy, cy, wy))
            return fail(h, TX_ERR_UNSUPPORTED, "tx_resize: scale needs more than 8 taps per axis");
        TX_CUDA(h, cudaStreamSynchronize(h->stream));
        cudaFree(h->d_rsa_i); cudaFree(h->d_rsa_w);
        h->d_rsa_i = nullptr; h->d_rsa_w = nullptr;
        TX_CUDA(h, cudaMalloc(&h->d_rsa_i, sizeof(int) * 2 * (Wo + Ho)));
        TX_CUDA(h, cudaMalloc(&h->d_rsa_w, sizeof(float) * TX_RS_TAPS * (Wo + Ho)));
        TX_CUDA(h, cudaMemcpy(h->d_rsa_i, fx.data(), sizeof(int) * Wo, cudaMemcpyHostToDevice));
        TX_CUDA(h, cudaMemcpy(h->d_rsa_i + Wo, cx.data(), sizeof(int) * Wo, cudaMemcpyHostToDevice));
        TX_CUDA(h, cudaMemcpy(h->d_rsa_i + 2 * Wo, fy.data(), sizeof(int) * Ho, cudaMemcpyHostToDevice));
        TX_CUDA(h, cudaMemcpy(h->d_rsa_i + 2 * Wo + Ho, cy.data(), sizeof(int) * Ho, cudaMemcpyHostToDevice));
        TX_CUDA(h, cudaMemcpy(h->d_rsa_w, wx.data(), sizeof(float) * TX_RS_TAPS * Wo, cudaMemcpyHostToDevice));
        TX_CUDA(h, cudaMemcpy(h->d_rsa_w + (size_t)TX_RS_TAPS * Wo, wy.data(), sizeof(float) * TX_RS_TAPS * Ho, cudaMemcpyHostToDevice));
        h->rs_Hi = Hi; h->rs_Wi = Wi;
    }
    ResizeArgs a{};
    a.src = src; a.dst = dst; a.Hi = Hi; a.Wi = Wi; a.Ho = Ho; a.Wo = Wo;
    a.fx = h->d_rsa_i; a.cx = h->d_rsa_i + Wo; a.fy = h->d_rsa_i + 2 * Wo; a.cy = h->d_rsa_i + 2 * Wo + Ho;
    a.wx = h->d_rsa_w; a.wy = h->d_rsa_w + (size_t)TX_RS_TAPS * Wo;
    TX_CUDA(h, launch_resize_aa(a, N, h->stream));
    h->ctr.kernels_launched++;
    return TX_OK;
}

extern "C" int tx_set_marker_patches(tx_handle* h, const uint8_t* patches)
{
    if (!h || !patches) return fail(h, TX_ERR_INVALID_ARG, "tx_set_marker_patches: null argument");
    if (h->generic) return fail(h, TX_ERR_UNSUPPORTED, "tx_set_marker_patches: the marker overlay belongs to the 240 x 320 kernel");
    TX_CUDA(h, cudaSetDevice(h->device));
    if (!h->d_patch) TX_CUDA(h, cudaMalloc(&h->d_patch, 10 * 10 * 12 * 12));
    TX_CUDA(h, cudaStreamSynchronize(h->stream));
    TX_CUDA(h, cudaMemcpy(h->d_patch, patches, 10 * 10 * 12 * 12, cudaMemcpyHostToDevice));
    return TX_OK;
}

extern "C" int tx_marker_overlay(tx_handle* h, const float* markers, int N, int M, const float* rgb_in, int apply, float* rgb_out,
                                 uint8_t* marker_img_out, uint8_t* rgb_u8_out)
{
    if (!h || N < 0 || M < 0 || M > TX_MAX_MARKERS) return fail(h, TX_ERR_INVALID_ARG, "tx_marker_overlay: bad argument");
    if (h->generic) return fail(h, TX_ERR_UNSUPPORTED, "tx_marker_overlay: the marker overlay belongs to the 240 x 320 kernel");
    if ((apply || marker_img_out) && (!markers || !h->d_patch))
        return fail(h, TX_ERR_STATE, "tx_marker_overlay: markers and tx_set_marker_patches are needed for the marker image");
    if ((rgb_out || rgb_u8_out) && !rgb_in) return fail(h, TX_ERR_INVALID_ARG, "tx_marker_overlay: rgb_in is needed for an RGB output");
    if (!rgb_out && !rgb_u8_out && !marker_img_out) return fail(h, TX_ERR_INVALID_ARG, "tx_marker_overlay: no output requested");
    if (N == 0) return TX_OK;
    TX_CUDA(h, cudaSetDevice(h->device));
    OverlayArgs a{};
    a.markers = markers; a.patch = h->d_patch; a.rgb_in = rgb_in; a.rgb = rgb_out; a.marker_img = marker_img_out; a.rgb_u8 = rgb_u8_out;
    a.M = (apply || marker_img_out) ? M : 0; a.apply = apply ? 1 : 0;
    TX_CUDA(h, launch_marker_overlay(a, N, h->stream));
    h->ctr.kernels_launched++;
    return TX_OK;
}

extern "C" int tx_debug_set_ticks(tx_handle* h, long long* ticks)
{
    if (!h) return TX_ERR_INVALID_ARG;
    h->d_ticks = ticks;
    const char* e = getenv("TX_DEBUG_FLAGS"); // profiling experiments (tools/phase_times.py); never set in production
    if (e) h->dbg = atoi(e);
    return TX_OK;
}

extern "C" int tx_debug_set_flags(tx_handle* h, int flags)
{
    if (!h) return TX_ERR_INVALID_ARG;
    h->dbg = flags;
    return TX_OK;
}

extern "C" int tx_marker_grid(const tx_handle* h, int32_t* mx, int32_t* my)
{
    if (!h || !mx || !my) return TX_ERR_INVALID_ARG;
    memcpy(mx, h->mx.data(), sizeof(int32_t) * h->M);
    memcpy(my, h->my.data(), sizeof(int32_t) * h->M);
    return TX_OK;
}

extern "C" int tx_step_host(tx_handle* h, const float* height_mm_host, const float* theta_host, int N, float* rgb_host,
                            float* depth_host, float* markers_host)
{
    if (!h || !height_mm_host || !rgb_host || N <= 0) return fail(h, TX_ERR_INVALID_ARG, "tx_step_host: bad argument");
    if (h->generic) return fail(h, TX_ERR_UNSUPPORTED, "tx_step_host: the pipelined host entry point belongs to the 240 x 320 kernel");
    if (markers_host && (!theta_host || h->M <= 0))
        return fail(h, TX_ERR_INVALID_ARG, "tx_step_host: markers need theta and a marker grid");
    if (N > h->cfg.max_envs) return fail(h, TX_ERR_INVALID_ARG, "tx_step_host: N exceeds max_envs");
    TX_CUDA(h, cudaSetDevice(h->device));
    const size_t HW = (size_t)h->cfg.H * h->cfg.W;
    if (h->host_cap < N) {
        cudaFree(h->d_hm); cudaFree(h->d_rgb); cudaFree(h->d_depth); cudaFree(h->d_theta); cudaFree(h->d_traj0);
        cudaFree(h->d_traj_len); cudaFree(h->d_markers);
        h->d_hm = h->d_rgb = h->d_depth = h->d_theta = h->d_traj0 = h->d_markers = nullptr;
        h->d_traj_len = nullptr;
        h->host_cap = 0;
        const size_t cap = (size_t)h->cfg.max_envs;
        TX_CUDA(h, cudaMalloc(&h->d_hm, sizeof(float) * HW * cap));
        TX_CUDA(h, cudaMalloc(&h->d_rgb, sizeof(float) * HW * 3 * cap));
        TX_CUDA(h, cudaMalloc(&h->d_depth, sizeof(float) * cap));
        TX_CUDA(h, cudaMalloc(&h->d_theta, sizeof(float) * cap));
        TX_CUDA(h, cudaMalloc(&h->d_traj0, sizeof(float) * 4 * cap));
        TX_CUDA(h, cudaMalloc(&h->d_traj_len, sizeof(int) * cap));
        TX_CUDA(h, cudaMemset(h->d_traj0, 0, sizeof(float) * 4 * cap));
        TX_CUDA(h, cudaMemset(h->d_traj_len, 0, sizeof(int) * cap));
        if (h->M > 0) TX_CUDA(h, cudaMalloc(&h->d_markers, sizeof(float) * 4 * h->M * cap));
        h->host_cap = (int)cap;
    }
    // Pipelined over chunks of envs: H2D of chunk c+1, kernels of chunk c and D2H of chunk c-1 overlap
    // (three streams, PCIe is full duplex). The chunks are disjoint slices of the same device buffers.
    const int CH = 256;
    const int nch = (N + CH - 1) / CH;
    if (!h->s_in) {
        TX_CUDA(h, cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
        TX_CUDA(h, cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
    }
    while ((int)h->ev_in.size() < nch) {
        cudaEvent_t a, b;
        TX_CUDA(h, cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        TX_CUDA(h, cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        h->ev_in.push_back(a);
        h->ev_done.push_back(b);
    }
    int* const saved_rect = h->d_rect; // the gather rectangles belong to device-resident batches, not to the chunked host path
    h->d_rect = nullptr;
    struct RestoreRect { tx_handle* h; int* r; ~RestoreRect() { h->d_rect = r; } } restore_rect{h, saved_rect};
    // (the fused multicast output needs the rectangles: it is off with them)
    // the copy streams must not run ahead of work already queued on the caller's stream
    TX_CUDA(h, cudaEventRecord(h->ev_done[0], h->stream));
    TX_CUDA(h, cudaStreamWaitEvent(h->s_in, h->ev_done[0], 0));
    TX_CUDA(h, cudaStreamWaitEvent(h->s_out, h->ev_done[0], 0));
    for (int c = 0; c < nch; ++c) {
        const int n0 = c * CH, n = (N - n0 < CH) ? N - n0 : CH;
        TX_CUDA(h, cudaMemcpyAsync(h->d_hm + HW * n0, height_mm_host + HW * n0, sizeof(float) * HW * n,
                                   cudaMemcpyHostToDevice, h->s_in));
        if (markers_host)
            TX_CUDA(h, cudaMemcpyAsync(h->d_theta + n0, theta_host + n0, sizeof(float) * n, cudaMemcpyHostToDevice, h->s_in));
        TX_CUDA(h, cudaEventRecord(h->ev_in[c], h->s_in));
        TX_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_in[c], 0));
        int rc = tx_render(h, h->d_hm + HW * n0, nullptr, n, h->d_rgb + HW * 3 * n0, h->d_depth + n0, nullptr, nullptr);
        if (rc != TX_OK) return rc;
        if (markers_host) {
            rc = tx_fots_markers(h, h->d_depth + n0, h->d_theta + n0, n, h->d_traj0 + 4 * n0, h->d_traj_len + n0,
                                 h->d_markers + (size_t)4 * h->M * n0);
            if (rc != TX_OK) return rc;
        }
        TX_CUDA(h, cudaEventRecord(h->ev_done[c], h->stream));
        TX_CUDA(h, cudaStreamWaitEvent(h->s_out, h->ev_done[c], 0));
        TX_CUDA(h, cudaMemcpyAsync(rgb_host + HW * 3 * n0, h->d_rgb + HW * 3 * n0, sizeof(float) * HW * 3 * n,
                                   cudaMemcpyDeviceToHost, h->s_out));
        if (markers_host)
            TX_CUDA(h, cudaMemcpyAsync(markers_host + (size_t)4 * h->M * n0, h->d_markers + (size_t)4 * h->M * n0,
                                       sizeof(float) * 4 * h->M * n, cudaMemcpyDeviceToHost, h->s_out));
        if (depth_host)
            TX_CUDA(h, cudaMemcpyAsync(depth_host + n0, h->d_depth + n0, sizeof(float) * n, cudaMemcpyDeviceToHost, h->s_out));
    }
    TX_CUDA(h, cudaStreamSynchronize(h->s_out));
    TX_CUDA(h, cudaStreamSynchronize(h->stream));
    return TX_OK;
}
