/*
 * oracle/taxim_canon.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Canonical CPU restatement (plain C, float32, fixed operation order) of the
 * reference's tactile hot path:
 *   - TaximSimulator.compute_indentation_depth   (ref: source/tacex/tacex/simulation_approaches/gpu_taxim/taxim_sim.py:115-131)
 *   - TaximTorch.__get_shifted_height_map        (ref: .../gpu_taxim/sim/taxim_torch.py:432-441)
 *   - TaximTorch.__compute_gel_pad_deformation   (ref: taxim_torch.py:443-473)
 *   - TaximTorch.__gaussian_blur / fast_conv2d   (ref: taxim_torch.py:382-412, 19-44)  -- restated as a direct separable
 *                                                  reflect-padded correlation (mathematically identical to the FFT path)
 *   - TaximTorch.__generate_normals              (ref: taxim_torch.py:475-503)
 *   - bin + polynomial lookup in __render        (ref: taxim_torch.py:243-258)
 *   - FOTS MarkerMotion._motion_callback etc.    (ref: .../fots/sim/marker_motion.py:78-120,144-219)
 *   - FOTSMarkerSimulator glue                   (ref: .../fots/fots_marker_sim.py:114-184)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load this library.
 * The product (tacex_b200/) never links or calls it.
 *
 * Parity status: PINNED against the executed reference (oracle/make_golden.py runs the reference's own
 * TaximTorch / MarkerMotion here and commits the vectors under tests/golden/); the reference's own tests
 * hold no golden vectors for this path.
 *
 * Canonical choices: blur = horizontal pass then vertical pass; horizontal outputs are
 * accumulated as acc = fmaf(w[k], x[k], acc) for k ascending from an initial 0.0f, vertical outputs centre-outward
 * with the symmetric pair summed first: acc = w[c]*x0; acc = fmaf(w[c+d], x[-d] + x[+d], acc), d = 1..r; gel map == 0 when the
 * caller passes gel == NULL; atanf / atan2f are the fixed polynomial below (identical text in the CUDA kernel).
 *
 * Build: gcc -O2 -ffp-contract=off -mfma -shared -fPIC (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define CANON_MAX_BLURS 8

typedef struct {
    int H, W;
    int num_bins;           /* 125 */
    float pixmm;            /* 0.0295 */
    float calib_h, calib_w; /* 480, 640 */
    float contact_scale;    /* 0.4 */
    int n_blurs;            /* pyramid levels + 1 final blur (7) */
    int ksx[CANON_MAX_BLURS];
    int ksy[CANON_MAX_BLURS];
    /* taps: for blur l, x taps at taps + off_x[l] (ksx[l] floats), y taps at taps + off_y[l] */
    int off_x[CANON_MAX_BLURS];
    int off_y[CANON_MAX_BLURS];
} canon_cfg;

/* ---------------- canonical elementary functions (float32, fixed op order) ---------------- */

/* Cephes-style atanf; every operation is a single IEEE float32 op (no contraction). */
static float canon_atanf(float xx)
{
    float x = fabsf(xx);
    float y;
    if (x > 2.414213562373095f) { /* tan(3 pi / 8) */
        y = 1.5707963267948966f;
        x = -(1.0f / x);
    } else if (x > 0.4142135623730950f) { /* tan(pi / 8) */
        y = 0.7853981633974483f;
        x = (x - 1.0f) / (x + 1.0f);
    } else {
        y = 0.0f;
    }
    float z = x * x;
    float p = 8.05374449538e-2f;
    p = fmaf(p, z, -1.38776856032e-1f);
    p = fmaf(p, z, 1.99777106478e-1f);
    p = fmaf(p, z, -3.33329491539e-1f);
    p = p * z;
    p = fmaf(p, x, x);
    y = y + p;
    return (xx < 0.0f) ? -y : y;
}

static float canon_atan2f(float y, float x)
{
    const float PI_F = 3.14159265358979323846f;
    const float PIO2_F = 1.5707963267948966f;
    if (x == 0.0f) {
        if (y > 0.0f) return PIO2_F;
        if (y < 0.0f) return -PIO2_F;
        return 0.0f;
    }
    float z = canon_atanf(y / x);
    if (x < 0.0f) {
        if (y < 0.0f) return z - PI_F;
        return z + PI_F;
    }
    return z;
}

static inline int reflect_idx(int i, int n)
{
    /* torch 'reflect' padding (no edge repeat); valid for |overhang| < n */
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

/* one separable blur: tmp = horizontal(src), dst = vertical(tmp) */
static void canon_blur(const float* src, float* tmp, float* dst, int H, int W, const float* kx, int ksx, const float* ky,
                       int ksy)
{
    int rx = (ksx - 1) / 2, ry = (ksy - 1) / 2;
    for (int y = 0; y < H; ++y) {
        const float* row = src + (size_t)y * W;
        for (int x = 0; x < W; ++x) {
            float acc = 0.0f;
            for (int k = 0; k < ksx; ++k) acc = fmaf(kx[k], row[reflect_idx(x + k - rx, W)], acc);
            tmp[(size_t)y * W + x] = acc;
        }
    }
    /* vertical: centre-outward, symmetric pairs summed first (direction independent):
       acc = w[c] * x[y]; acc = fma(w[c + d], x[y - d] + x[y + d], acc) for d = 1..ry */
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            float acc = ky[ry] * tmp[(size_t)y * W + x];
            for (int d = 1; d <= ry; ++d) {
                float a = tmp[(size_t)reflect_idx(y - d, H) * W + x];
                float b = tmp[(size_t)reflect_idx(y + d, H) * W + x];
                acc = fmaf(ky[ry + d], a + b, acc);
            }
            dst[(size_t)y * W + x] = acc;
        }
    }
}

/* ref: taxim_sim.py:115-131 */
void canon_indentation_depth(const float* hm_mm, int N, int H, int W, float gelpad_height_m, float gelpad_to_cam_min_m,
                             float* depth_mm)
{
    for (int n = 0; n < N; ++n) {
        const float* p = hm_mm + (size_t)n * H * W;
        float m = p[0];
        for (int i = 1; i < H * W; ++i) m = fminf(m, p[i]);
        float d = m / 1000.0f;
        d = d - gelpad_to_cam_min_m;
        if (d < 0.0f) d = 0.0f;
        depth_mm[n] = (d <= gelpad_height_m) ? (gelpad_height_m - d) * 1000.0f : 0.0f;
    }
}

/*
 * Full optical path for a batch.
 *   hm_mm   [N][H][W]   height map in mm (as GelSightSensor._get_height_map produces it)
 *   press   [N]         indentation depth in mm
 *   taps                concatenated 1-D Gaussian taps (see canon_cfg)
 *   poly    [3][nb][nb][6], bg [3][H][W], gel [H][W] or NULL (== 0)
 * outputs (any may be NULL): deformed [N][H][W], mask [N][H][W] u8, idx_mag/idx_dir [N][H][W] i32, rgb [N][H][W][3]
 */
int canon_taxim_render(const canon_cfg* cfg, const float* taps, const float* poly, const float* bg, const float* gel,
                       const float* hm_mm, const float* press, int N, float* deformed, uint8_t* mask_out,
                       int32_t* idx_mag_out, int32_t* idx_dir_out, float* rgb)
{
    const int H = cfg->H, W = cfg->W, HW = H * W, nb = cfg->num_bins;
    int status = 0;
#pragma omp parallel for schedule(dynamic)
    for (int n = 0; n < N; ++n) {
        float* h = (float*)malloc(sizeof(float) * HW);
        float* j = (float*)malloc(sizeof(float) * HW);
        float* b = (float*)malloc(sizeof(float) * HW);
        float* t1 = (float*)malloc(sizeof(float) * HW);
        float* t2 = (float*)malloc(sizeof(float) * HW);
        uint8_t* mk = (uint8_t*)malloc(HW);
        float* mag = (float*)malloc(sizeof(float) * HW);
        float* dir = (float*)malloc(sizeof(float) * HW);
        if (!h || !j || !b || !t1 || !t2 || !mk || !mag || !dir) {
            status = -1;
        } else {
            const float* hm = hm_mm + (size_t)n * HW;
            /* ref: taxim_torch.py:441  height_map - amin - press */
            float m = hm[0];
            for (int i = 1; i < HW; ++i) m = fminf(m, hm[i]);
            for (int i = 0; i < HW; ++i) h[i] = (hm[i] - m) - press[n];
            /* ref: taxim_torch.py:449-461 */
            float hmin = h[0];
            for (int i = 1; i < HW; ++i) hmin = fminf(hmin, h[i]);
            float pd = -hmin;
            float thr = (-pd) * cfg->contact_scale;
            for (int i = 0; i < HW; ++i) {
                float g = gel ? gel[i] : 0.0f;
                int contact = h[i] < 0.0f;
                j[i] = fminf(h[i], g);
                mk[i] = (uint8_t)(((j[i] - g) < thr) && contact);
            }
            /* ref: taxim_torch.py:464-471 */
            memcpy(b, j, sizeof(float) * HW);
            for (int l = 0; l < cfg->n_blurs; ++l) {
                canon_blur(b, t1, t2, H, W, taps + cfg->off_x[l], cfg->ksx[l], taps + cfg->off_y[l], cfg->ksy[l]);
                if (l < cfg->n_blurs - 1) {
                    for (int i = 0; i < HW; ++i) b[i] = mk[i] ? j[i] : t2[i];
                } else {
                    memcpy(b, t2, sizeof(float) * HW);
                }
            }
            if (deformed) memcpy(deformed + (size_t)n * HW, b, sizeof(float) * HW);
            if (mask_out) memcpy(mask_out + (size_t)n * HW, mk, HW);

            if (rgb || idx_mag_out || idx_dir_out) {
                /* ref: taxim_torch.py:237-238, 475-503. Canonical (GPU-friendly) operation order, each step one
                   float32 op: zs = b * (1/pixmm); central differences * 0.5; scale by H/calib_h, W/calib_w;
                   tt = sqrt(fma(gx, gx, gy*gy)); dir = atan2(gx, gy) (the reference normalises both by tt first,
                   which is the same angle). Differences to the literal reference sequence are <= 1-2 ulp, far inside
                   the reference's own FFT noise. */
                const float inv_pixmm = 1.0f / cfg->pixmm;
                const float sy = (float)H / cfg->calib_h, sx = (float)W / cfg->calib_w;
                for (int i = 0; i < HW; ++i) t1[i] = b[i] * inv_pixmm;
                for (int y = 1; y < H - 1; ++y) {
                    for (int x = 1; x < W - 1; ++x) {
                        float top = t1[(y - 1) * W + x], bot = t1[(y + 1) * W + x];
                        float left = t1[y * W + x - 1], right = t1[y * W + x + 1];
                        float dzdx = (top - bot) * 0.5f;
                        float dzdy = (left - right) * 0.5f;
                        float gx = dzdx * sy;
                        float gy = dzdy * sx;
                        float tt = sqrtf(fmaf(gx, gx, gy * gy));
                        mag[y * W + x] = canon_atanf(tt);
                        dir[y * W + x] = (tt != 0.0f) ? canon_atan2f(gx, gy) : 0.0f;
                    }
                }
                /* replicate pad by 1 (ref: taxim_torch.py:501-502) */
                for (int y = 0; y < H; ++y) {
                    int yy = y < 1 ? 1 : (y > H - 2 ? H - 2 : y);
                    for (int x = 0; x < W; ++x) {
                        int xx = x < 1 ? 1 : (x > W - 2 ? W - 2 : x);
                        if (yy != y || xx != x) {
                            mag[y * W + x] = mag[yy * W + xx];
                            dir[y * W + x] = dir[yy * W + xx];
                        }
                    }
                }
                /* ref: taxim_torch.py:241-258 */
                const float inv_xbin = (float)(1.0 / (0.5 * M_PI / (nb - 1)));
                const float inv_ybin = (float)(1.0 / (2.0 * M_PI / (nb - 1)));
                const float pi_f = (float)M_PI;
                const float fx = cfg->calib_w / (float)W, fy = cfg->calib_h / (float)H;
                for (int y = 0; y < H; ++y) {
                    for (int x = 0; x < W; ++x) {
                        int i = y * W + x;
                        int im = (int)floorf(mag[i] * inv_xbin);
                        int id = (int)floorf((dir[i] + pi_f) * inv_ybin);
                        if (idx_mag_out) idx_mag_out[(size_t)n * HW + i] = im;
                        if (idx_dir_out) idx_dir_out[(size_t)n * HW + i] = id;
                        if (rgb) {
                            if (im < 0) im = 0;
                            if (im > nb - 1) im = nb - 1;
                            if (id < 0) id = 0;
                            if (id > nb - 1) id = nb - 1;
                            /* features: x = col * calib_w / W, y = row * calib_h / H (ref: taxim_torch.py:139-157) */
                            float xf = (float)x * fx;
                            float yf = (float)y * fy;
                            float f0 = xf * xf, f1 = yf * yf, f2 = xf * yf;
                            for (int c = 0; c < 3; ++c) {
                                const float* p = poly + (((size_t)c * nb + im) * nb + id) * 6;
                                float s = p[5];
                                s = fmaf(p[4], yf, s);
                                s = fmaf(p[3], xf, s);
                                s = fmaf(p[2], f2, s);
                                s = fmaf(p[1], f1, s);
                                s = fmaf(p[0], f0, s);
                                s = s + bg[(size_t)c * HW + i];
                                s = fminf(fmaxf(s, 0.0f), 1.0f);
                                rgb[((size_t)n * HW + i) * 3 + c] = s;
                            }
                        }
                    }
                }
            }
        }
        free(h); free(j); free(b); free(t1); free(t2); free(mk); free(mag); free(dir);
    }
    return status;
}

/* ---------------- Taxim SHADOW branch (with_shadow = True; ref: taxim_torch.py:260-346) ----------------
 * Restated on top of the same canonical deformation / normals / polynomial image as canon_taxim_render (the code below repeats
 * that function's steps on purpose: the no-shadow checker stays untouched). Canonical choices: the ray-fan trigonometry comes in
 * as float32 tables (the reference evaluates torch.cos / torch.sin of the same float32 angles at init), every coordinate is
 * step * (s + 1) * cos -> + pixel -> truncation toward zero, all in float32 like the tensor expression; blurs are the direct
 * separable correlation of canon_blur. */
typedef struct {
    int D, Hn, S, F;          /* directions (63), height entries (24), table row length (51), fan rays (4) */
    float depth_0;            /* 0.4 */
    float height_precision;   /* 0.1 */
    float discretize_precision; /* 0.1 */
    float step_x, step_y;     /* 0.625 */
    int dil[2][2];            /* two dilation rounds: (ky, kx) */
    int ks_sx, ks_sy;         /* shadow blur tap counts */
} canon_shadow_cfg;

/* conv2d(ones(ky, kx), padding='same') != 0 of a 0/1 image, zero padded; torch pads the extra element of an even kernel on the
 * right / bottom: out[y][x] = OR over in[y - (ky-1)/2 .. + ky - 1][x - (kx-1)/2 .. + kx - 1] */
static void dilate_same(const uint8_t* in, uint8_t* out, int H, int W, int ky, int kx)
{
    const int py = (ky - 1) / 2, px = (kx - 1) / 2;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            uint8_t v = 0;
            for (int j = 0; j < ky && !v; ++j) {
                const int yy = y - py + j;
                if (yy < 0 || yy >= H) continue;
                for (int i = 0; i < kx; ++i) {
                    const int xx = x - px + i;
                    if (xx >= 0 && xx < W && in[yy * W + xx]) { v = 1; break; }
                }
            }
            out[y * W + x] = v;
        }
}

int canon_taxim_render_shadow(const canon_cfg* cfg, const float* taps, const float* poly, const float* bg, const float* gel,
                              const float* hm_mm, const float* press, int N, const canon_shadow_cfg* sc,
                              const float* table /*[3][D][Hn][S]*/, const float* fan_cos /*[D][F]*/, const float* fan_sin,
                              const float* staps_x, const float* staps_y, float* rgb /*[N][H][W][3]*/, uint8_t* boundary_out,
                              float* shadow_out /*[N][3][H][W] or NULL: the scatter-min image, +inf where no shadow sample landed*/)
{
    const int H = cfg->H, W = cfg->W, HW = H * W, nb = cfg->num_bins;
    int status = 0;
#pragma omp parallel for schedule(dynamic)
    for (int n = 0; n < N; ++n) {
        float* h = (float*)malloc(sizeof(float) * HW);
        float* j = (float*)malloc(sizeof(float) * HW);
        float* b = (float*)malloc(sizeof(float) * HW);
        float* t1 = (float*)malloc(sizeof(float) * HW);
        float* t2 = (float*)malloc(sizeof(float) * HW);
        uint8_t* mk = (uint8_t*)malloc(HW);
        uint8_t* d1 = (uint8_t*)malloc(HW);
        uint8_t* d2 = (uint8_t*)malloc(HW);
        float* mag = (float*)malloc(sizeof(float) * HW);
        float* dir = (float*)malloc(sizeof(float) * HW);
        float* raw = (float*)malloc(sizeof(float) * 3 * HW);  /* polynomial image without background, [3][H][W] */
        float* sh = (float*)malloc(sizeof(float) * 3 * HW);   /* shadow image */
        float* dpx = (float*)malloc(sizeof(float) * HW);
        if (!h || !j || !b || !t1 || !t2 || !mk || !d1 || !d2 || !mag || !dir || !raw || !sh || !dpx) {
            status = -1;
        } else {
            const float* hm = hm_mm + (size_t)n * HW;
            /* ---- deformation, identical to canon_taxim_render ---- */
            float m = hm[0];
            for (int i = 1; i < HW; ++i) m = fminf(m, hm[i]);
            for (int i = 0; i < HW; ++i) h[i] = (hm[i] - m) - press[n];
            float hmin = h[0];
            for (int i = 1; i < HW; ++i) hmin = fminf(hmin, h[i]);
            float thr = (-(-hmin)) * cfg->contact_scale;
            for (int i = 0; i < HW; ++i) {
                float g = gel ? gel[i] : 0.0f;
                int contact = h[i] < 0.0f;
                j[i] = fminf(h[i], g);
                mk[i] = (uint8_t)(((j[i] - g) < thr) && contact);
            }
            memcpy(b, j, sizeof(float) * HW);
            for (int l = 0; l < cfg->n_blurs; ++l) {
                canon_blur(b, t1, t2, H, W, taps + cfg->off_x[l], cfg->ksx[l], taps + cfg->off_y[l], cfg->ksy[l]);
                if (l < cfg->n_blurs - 1) {
                    for (int i = 0; i < HW; ++i) b[i] = mk[i] ? j[i] : t2[i];
                } else {
                    memcpy(b, t2, sizeof(float) * HW);
                }
            }
            /* ---- normals, identical to canon_taxim_render ---- */
            const float inv_pixmm = 1.0f / cfg->pixmm;
            const float sy = (float)H / cfg->calib_h, sx = (float)W / cfg->calib_w;
            for (int i = 0; i < HW; ++i) t1[i] = b[i] * inv_pixmm;
            for (int y = 1; y < H - 1; ++y)
                for (int x = 1; x < W - 1; ++x) {
                    float top = t1[(y - 1) * W + x], bot = t1[(y + 1) * W + x];
                    float left = t1[y * W + x - 1], right = t1[y * W + x + 1];
                    float gx = ((top - bot) * 0.5f) * sy;
                    float gy = ((left - right) * 0.5f) * sx;
                    float tt = sqrtf(fmaf(gx, gx, gy * gy));
                    mag[y * W + x] = canon_atanf(tt);
                    dir[y * W + x] = (tt != 0.0f) ? canon_atan2f(gx, gy) : 0.0f;
                }
            for (int y = 0; y < H; ++y) {
                int yy = y < 1 ? 1 : (y > H - 2 ? H - 2 : y);
                for (int x = 0; x < W; ++x) {
                    int xx = x < 1 ? 1 : (x > W - 2 ? W - 2 : x);
                    if (yy != y || xx != x) {
                        mag[y * W + x] = mag[yy * W + xx];
                        dir[y * W + x] = dir[yy * W + xx];
                    }
                }
            }
            /* ---- polynomial image WITHOUT background (ref: taxim_torch.py:248-253 sim_img_r) ---- */
            const float inv_xbin = (float)(1.0 / (0.5 * M_PI / (nb - 1)));
            const float inv_ybin = (float)(1.0 / (2.0 * M_PI / (nb - 1)));
            const float pi_f = (float)M_PI;
            const float fx = cfg->calib_w / (float)W, fy = cfg->calib_h / (float)H;
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    int i = y * W + x;
                    int im = (int)floorf(mag[i] * inv_xbin);
                    int id = (int)floorf((dir[i] + pi_f) * inv_ybin);
                    if (im < 0) im = 0;
                    if (im > nb - 1) im = nb - 1;
                    if (id < 0) id = 0;
                    if (id > nb - 1) id = nb - 1;
                    float xf = (float)x * fx, yf = (float)y * fy;
                    float f0 = xf * xf, f1 = yf * yf, f2 = xf * yf;
                    for (int c = 0; c < 3; ++c) {
                        const float* p = poly + (((size_t)c * nb + im) * nb + id) * 6;
                        float s = p[5];
                        s = fmaf(p[4], yf, s);
                        s = fmaf(p[3], xf, s);
                        s = fmaf(p[2], f2, s);
                        s = fmaf(p[1], f1, s);
                        s = fmaf(p[0], f0, s);
                        raw[(size_t)c * HW + i] = s;
                    }
                }
            /* ---- shadow attachment area: dilated mask minus mask (ref: taxim_torch.py:260-272) ---- */
            dilate_same(mk, d1, H, W, sc->dil[0][0], sc->dil[0][1]);
            dilate_same(d1, d2, H, W, sc->dil[1][0], sc->dil[1][1]);
            for (int i = 0; i < HW; ++i) d2[i] = (uint8_t)(d2[i] && !mk[i]);
            if (boundary_out) memcpy(boundary_out + (size_t)n * HW, d2, HW);
            /* ---- cast the shadows (ref: taxim_torch.py:274-336) ---- */
            for (int i = 0; i < 3 * HW; ++i) sh[i] = INFINITY;
            for (int i = 0; i < HW; ++i) dpx[i] = b[i] / cfg->pixmm;
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    const int i = y * W + x;
                    if (!d2[i]) continue;
                    int nidx = (int)floorf((dir[i] + pi_f) / sc->discretize_precision);
                    if (nidx < 0) nidx = 0;
                    if (nidx > sc->D - 1) nidx = sc->D - 1;
                    const float g = gel ? gel[i] : 0.0f;
                    const float ch_px = (g - b[i]) / cfg->pixmm;
                    int hidx = (int)floorf((ch_px * cfg->pixmm - sc->depth_0) / sc->height_precision) + 6;
                    const int hmax = sc->Hn - 1;
                    if (hidx < 0 || hidx >= hmax) hidx = hmax;
                    for (int f = 0; f < sc->F; ++f) {
                        const float cs = fan_cos[nidx * sc->F + f], sn = fan_sin[nidx * sc->F + f];
                        for (int s = 0; s < sc->S; ++s) {
                            const float kx_ = sc->step_x * (float)(s + 1), ky_ = sc->step_y * (float)(s + 1);
                            const long cx = (long)((float)x + kx_ * cs);
                            const long cy = (long)((float)y + ky_ * sn);
                            if (cx < 0 || cx >= W || cy < 0 || cy >= H) continue;
                            const int tgt = (int)cy * W + (int)cx;
                            if (!(dpx[i] < dpx[tgt])) continue;
                            for (int c = 0; c < 3; ++c) {
                                const float v = table[(((size_t)c * sc->D + nidx) * sc->Hn + hidx) * sc->S + s];
                                if (v < sh[(size_t)c * HW + tgt]) sh[(size_t)c * HW + tgt] = v;
                            }
                        }
                    }
                }
            if (shadow_out) memcpy(shadow_out + (size_t)n * 3 * HW, sh, sizeof(float) * 3 * HW);
            /* ---- min, shadow blur, + background, final blur, clip (ref: taxim_torch.py:337-346) ---- */
            const int lf = cfg->n_blurs - 1;
            for (int c = 0; c < 3; ++c) {
                float* r = raw + (size_t)c * HW;
                for (int i = 0; i < HW; ++i) r[i] = fminf(r[i], sh[(size_t)c * HW + i]);
                canon_blur(r, t1, t2, H, W, staps_x, sc->ks_sx, staps_y, sc->ks_sy);
                for (int i = 0; i < HW; ++i) t2[i] = t2[i] + bg[(size_t)c * HW + i];
                canon_blur(t2, t1, r, H, W, taps + cfg->off_x[lf], cfg->ksx[lf], taps + cfg->off_y[lf], cfg->ksy[lf]);
                for (int i = 0; i < HW; ++i) rgb[((size_t)n * HW + i) * 3 + c] = fminf(fmaxf(r[i], 0.0f), 1.0f);
            }
        }
        free(h); free(j); free(b); free(t1); free(t2); free(mk); free(d1); free(d2); free(mag); free(dir); free(raw); free(sh);
        free(dpx);
    }
    return status;
}

/* ---------------- FOTS marker motion (float64 like the NumPy reference) ---------------- */

typedef struct {
    int H, W;
    int rows, cols;        /* marker grid */
    double lamb[3];        /* dilate, shear, twist */
    double mm2pix;         /* 19.58 */
    double shear_max_px;   /* 10 */
    double theta_max_rad;  /* 60 deg */
} canon_fots_cfg;

/* ref: marker_motion.py:58-76 -- np.linspace(x0, W - x0, cols, dtype=int) */
void canon_fots_grid(const canon_fots_cfg* c, double x0, double y0, int32_t* mx /*[rows*cols]*/, int32_t* my)
{
    for (int r = 0; r < c->rows; ++r) {
        for (int q = 0; q < c->cols; ++q) {
            double stepx = c->cols > 1 ? ((double)(c->W - x0) - x0) / (double)(c->cols - 1) : 0.0;
            double stepy = c->rows > 1 ? ((double)(c->H - y0) - y0) / (double)(c->rows - 1) : 0.0;
            double vx = (q == c->cols - 1 && c->cols > 1) ? (double)(c->W - x0) : x0 + stepx * q;
            double vy = (r == c->rows - 1 && c->rows > 1) ? (double)(c->H - y0) : y0 + stepy * r;
            mx[r * c->cols + q] = (int32_t)vx; /* astype(int) truncates toward zero */
            my[r * c->cols + q] = (int32_t)vy;
        }
    }
}

/*
 * One FOTS step for a batch (restates fots_marker_sim.py:128-182 + marker_motion.py:144-219).
 *   deformed [N][H][W] f32, mask [N][H][W] u8, press [N], theta [N]
 *   traj0 [N][4] inout: (x0_mm, y0_mm, theta0, valid) -- first in-contact sample of the current contact episode
 *   traj_len [N] inout: number of samples in the episode (reference keeps a python list; only [0], [-1], len matter)
 *   markers [N][2][M][2] out, M = rows*cols; [:,0] initial, [:,1] current, last dim (x, y)
 */
void canon_fots_step(const canon_fots_cfg* c, const int32_t* mx, const int32_t* my, const float* deformed,
                     const uint8_t* mask, const float* press, const float* theta, int N, float* traj0,
                     int32_t* traj_len, float* markers)
{
    const int H = c->H, W = c->W, HW = H * W, M = c->rows * c->cols;
    for (int n = 0; n < N; ++n) {
        float* out0 = markers + (size_t)n * 2 * M * 2;
        float* out1 = out0 + (size_t)M * 2;
        for (int m = 0; m < M; ++m) {
            out0[2 * m] = (float)mx[m];
            out0[2 * m + 1] = (float)my[m];
        }
        if (!(press[n] > 0.0f)) {
            traj_len[n] = 0;
            traj0[4 * n + 3] = 0.0f;
            for (int m = 0; m < M; ++m) {
                out1[2 * m] = (float)mx[m];
                out1[2 * m + 1] = (float)my[m];
            }
            continue;
        }
        const float* b = deformed + (size_t)n * HW;
        const uint8_t* mk = mask + (size_t)n * HW;
        /* contact centroid: torch.mean(argwhere(mask).float(), dim=0) -- float32 mean; restated with exact
           integer sums then one float32 division (differs from torch's float32 summation by <= a few ulp) */
        double srow = 0.0, scol = 0.0;
        long cnt = 0;
        float bmax = b[0];
        for (int i = 0; i < HW; ++i) {
            bmax = fmaxf(bmax, b[i]);
            if (mk[i]) {
                srow += (double)(i / W);
                scol += (double)(i % W);
                ++cnt;
            }
        }
        float mrow = (float)(srow / (double)cnt); /* NaN when cnt == 0, like torch.mean of an empty tensor */
        float mcol = (float)(scol / (double)cnt);
        /* numpy float32 arithmetic with python-float scalars stays float32 (NumPy 2) */
        float cy = (mrow - (float)(H / 2.0)) / (float)c->mm2pix;
        float cx = (mcol - (float)(W / 2.0)) / (float)c->mm2pix;
        float th = theta[n];
        if (traj_len[n] == 0) {
            traj0[4 * n + 0] = cx;
            traj0[4 * n + 1] = cy;
            traj0[4 * n + 2] = th;
            traj0[4 * n + 3] = 1.0f;
        }
        traj_len[n] += 1;

        /* depth = (max(b) - b) - min(...) = max(b) - b ; /10 (float32, marker_motion.py:146-149).
           The reference subtracts the BATCH max and then the per-env min; that is per-env (max_b - b) up to
           one float32 rounding. */
        /* contact list: markers whose initial integer position lies on the mask */
        double dxs[1024], dys[1024];
        for (int m = 0; m < M; ++m) { dxs[m] = 0.0; dys[m] = 0.0; }
        int ncontact = 0;
        for (int q = 0; q < c->cols; ++q) {
            for (int r = 0; r < c->rows; ++r) {
                int m = r * c->cols + q;
                int yp = my[m], xp = mx[m];
                if (yp >= H || yp < 0 || xp >= W || xp < 0) continue;
                if (mk[yp * W + xp] == 1) {
                    float hgt = (bmax - b[yp * W + xp]) / 10.0f;
                    ++ncontact;
                    /* dilate contribution of this contact point on every marker (marker_motion.py:111-120) */
                    for (int k = 0; k < M; ++k) {
                        double ox = (double)(mx[k] - xp), oy = (double)(my[k] - yp);
                        double g = exp(-c->lamb[0] * (ox * ox + oy * oy));
                        dxs[k] += (double)hgt * ox * g;
                        dys[k] += (double)hgt * oy * g;
                    }
                }
            }
        }
        if (ncontact == 0) {
            for (int m = 0; m < M; ++m) {
                out1[2 * m] = (float)mx[m];
                out1[2 * m + 1] = (float)my[m];
            }
            continue;
        }
        int have_traj = traj_len[n] >= 2;
        double scx = 0, scy = 0, shx = 0, shy = 0, tcx = 0, tcy = 0, cm1 = 0, sn = 0;
        if (have_traj) {
            /* python: float32 scalar * python float -> float32 in NumPy 2 ; int() truncates toward zero */
            float x0 = traj0[4 * n + 0], y0 = traj0[4 * n + 1], t0 = traj0[4 * n + 2];
            scx = (double)(int)(x0 * (float)c->mm2pix + (float)(W / 2.0));
            scy = (double)(int)(y0 * (float)c->mm2pix + (float)(H / 2.0));
            shx = (double)(int)((cx - x0) * (float)c->mm2pix);
            shy = (double)(int)((cy - y0) * (float)c->mm2pix);
            if (shx > c->shear_max_px) shx = c->shear_max_px;
            if (shx < -c->shear_max_px) shx = -c->shear_max_px;
            if (shy > c->shear_max_px) shy = c->shear_max_px;
            if (shy < -c->shear_max_px) shy = -c->shear_max_px;
            tcx = (double)(int)(cx * (float)c->mm2pix + (float)(W / 2.0));
            tcy = (double)(int)(cy * (float)c->mm2pix + (float)(H / 2.0));
            /* np.clip / np.cos / np.sin on a float32 scalar stay float32 (NumPy 2 weak python scalars) */
            float thf = th - t0;
            float tmax = (float)c->theta_max_rad;
            if (thf > tmax) thf = tmax;
            if (thf < -tmax) thf = -tmax;
            cm1 = (double)cosf(thf - 1.0f); /* sic: cos(theta - 1), marker_motion.py:98-99 */
            sn = (double)sinf(thf);
        }
        for (int m = 0; m < M; ++m) {
            double px = (double)mx[m], py = (double)my[m];
            double nx = px + dxs[m], ny = py + dys[m];
            if (have_traj) {
                double ox = px - scx, oy = py - scy;
                double g = exp(-c->lamb[1] * (ox * ox + oy * oy));
                nx += shx * g;
                ny += shy * g;
                ox = px - tcx;
                oy = py - tcy;
                g = exp(-c->lamb[2] * (ox * ox + oy * oy));
                double rx = ox * cm1 - oy * sn;
                double ry = ox * sn + oy * cm1;
                nx += rx * g;
                ny += ry * g;
            }
            out1[2 * m] = (float)nx;
            out1[2 * m + 1] = (float)ny;
        }
    }
}

/* ---------------- camera resolution != tactile resolution: bilinear resize of the height map ----------------
 * ref: TaximSimulator.optical_simulation, .../gpu_taxim/taxim_sim.py:88-89 (torchvision F.resize, bilinear, antialias) and
 * FOTSMarkerSimulator, .../fots/fots_marker_sim.py:121-127. torch's antialiased bilinear filter (aten UpSampleKernel.cpp,
 * _compute_indices_min_size_weights_aa): scale = in / out, centre = scale (i + 0.5), support = 1 when up-sampling (scale when
 * down-sampling), taps xmin = max(int(centre - support + 0.5), 0) .. min(int(centre + support + 0.5), in) - 1 with triangle
 * weights normalised to sum 1. Canonical choices: float32 weight arithmetic (as aten does for float32 images); at most
 * CANON_RS_TAPS taps; each 1-D pass accumulates acc = w[0] x[0]; acc = fmaf(w[k], x[k], acc); horizontal pass first. */
#define CANON_RS_TAPS 8
int canon_resize_weights(int in_size, int out_size, int32_t* first /*[out]*/, int32_t* count /*[out]*/, float* w /*[out][CANON_RS_TAPS]*/)
{
    /* float32 arithmetic throughout, like aten computes the weights for float32 images */
    const float scale = (float)in_size / (float)out_size;
    const float support = scale >= 1.0f ? scale : 1.0f;
    const float invscale = scale >= 1.0f ? 1.0f / scale : 1.0f;
    for (int i = 0; i < out_size; ++i) {
        const float center = scale * ((float)i + 0.5f);
        long xmin = (long)(center - support + 0.5f);
        if (xmin < 0) xmin = 0;
        long xmax = (long)(center + support + 0.5f);
        if (xmax > in_size) xmax = in_size;
        const int n = (int)(xmax - xmin);
        if (n < 1 || n > CANON_RS_TAPS) return -1;
        float wd[CANON_RS_TAPS], tot = 0.0f;
        for (int j = 0; j < n; ++j) {
            float x = ((float)(j + xmin) - center + 0.5f) * invscale;
            if (x < 0) x = -x;
            wd[j] = x < 1.0f ? 1.0f - x : 0.0f;
            tot += wd[j];
        }
        first[i] = (int32_t)xmin;
        count[i] = n;
        for (int j = 0; j < CANON_RS_TAPS; ++j) w[i * CANON_RS_TAPS + j] = j < n ? wd[j] / tot : 0.0f;
    }
    return 0;
}

int canon_resize_bilinear(const float* src /*[N][Hi][Wi]*/, int N, int Hi, int Wi, int Ho, int Wo, float* dst /*[N][Ho][Wo]*/)
{
    int32_t *fx = malloc(sizeof(int32_t) * Wo), *cx = malloc(sizeof(int32_t) * Wo), *fy = malloc(sizeof(int32_t) * Ho), *cy = malloc(sizeof(int32_t) * Ho);
    float *wx = malloc(sizeof(float) * Wo * CANON_RS_TAPS), *wy = malloc(sizeof(float) * Ho * CANON_RS_TAPS);
    int rc = -1;
    if (fx && cx && fy && cy && wx && wy && canon_resize_weights(Wi, Wo, fx, cx, wx) == 0 && canon_resize_weights(Hi, Ho, fy, cy, wy) == 0) {
        rc = 0;
#pragma omp parallel for schedule(static)
        for (int n = 0; n < N; ++n) {
            float* tmp = malloc(sizeof(float) * (size_t)Hi * Wo);
            const float* s = src + (size_t)n * Hi * Wi;
            float* d = dst + (size_t)n * Ho * Wo;
            for (int y = 0; y < Hi; ++y)
                for (int X = 0; X < Wo; ++X) {
                    const float* sp = s + (size_t)y * Wi + fx[X];
                    const float* w = wx + X * CANON_RS_TAPS;
                    float acc = w[0] * sp[0];
                    for (int j = 1; j < cx[X]; ++j) acc = fmaf(w[j], sp[j], acc);
                    tmp[(size_t)y * Wo + X] = acc;
                }
            for (int Y = 0; Y < Ho; ++Y)
                for (int X = 0; X < Wo; ++X) {
                    const float* tp = tmp + (size_t)fy[Y] * Wo + X;
                    const float* w = wy + Y * CANON_RS_TAPS;
                    float acc = w[0] * tp[0];
                    for (int j = 1; j < cy[Y]; ++j) acc = fmaf(w[j], tp[(size_t)j * Wo], acc);
                    d[(size_t)Y * Wo + X] = acc;
                }
            free(tmp);
        }
    }
    free(fx); free(cx); free(fy); free(cy); free(wx); free(wy);
    return rc;
}

/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline legs ask for all host threads explicitly */
void canon_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#endif
}

int canon_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
