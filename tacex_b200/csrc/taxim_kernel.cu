// taxim_kernel.cu -- fused Taxim optical model for sm_100a: one 2-CTA cluster per 240x320 frame.
//
// Replaces (ref = /root/reference/source/tacex/tacex/simulation_approaches/gpu_taxim):
//   taxim_sim.py:115-131          compute_indentation_depth           (fused when press_in == NULL)
//   sim/taxim_torch.py:432-441    __get_shifted_height_map
//   sim/taxim_torch.py:443-473    __compute_gel_pad_deformation  (6-level Gaussian pyramid + masked re-imposition + final blur)
//   sim/taxim_torch.py:475-503    __generate_normals
//   sim/taxim_torch.py:243-258    bin -> polynomial table -> + background -> clip
//   taxim_sim.py:104-111          NCHW -> NHWC
//
// Data layout: each CTA of the cluster owns 120 rows of the frame as ONE float32 plane resident in shared memory
// (153,600 B, staged by a bulk-async TMA copy of the contiguous half frame). Every blur level runs in place:
//   horizontal pass  one warp per row, the row lives in registers (12 virtual columns per lane incl. reflect
//                    padding), neighbours reached with warp shuffles, taps from __constant__ memory;
//   halo exchange    the rows next to the CTA boundary are written straight from registers into the peer CTA's
//                    halo buffer through distributed shared memory, then ONE cluster barrier;
//   vertical pass    one thread per column marching from the image edge towards the CTA boundary with a sliding
//                    register window (in place, reads run ahead of writes);
//   re-imposition    plane[mask] = joined height map, mask kept as a bit plane (4.8 KB).
// The epilogue computes normals, bins, the polynomial lookup and stores NHWC RGB. Nothing but the input frame and
// the RGB frame touches HBM (algorithmic traffic: 307,200 B in + 921,600 B out per frame).
#include "tx_common.cuh"
#include "tx_kernels.h"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace tx {

// The blur taps travel as a __grid_constant__ kernel parameter (constant bank 0, read as immediate-offset operands of the
// FMAs exactly like a __constant__ array), so every handle -- every calibration -- brings its own and two handles of one
// process can never see each other's taps.

// ---- shared memory carve-up ---------------------------------------------------------------------------------
constexpr int HB0_ROWS = 30; // even levels (radius 30, 8, 2, 2)
constexpr int HB1_ROWS = 16; // odd levels (radius 16, 4, 1) and the 1-row epilogue halo
constexpr int SM_PLANE = 0;
constexpr int SM_HB0 = SM_PLANE + HALF_H * IMG_W * 4;
constexpr int SM_HB1 = SM_HB0 + HB0_ROWS * IMG_W * 4;
constexpr int SM_MASK = SM_HB1 + HB1_ROWS * IMG_W * 4;
constexpr int SM_MLIST = SM_MASK + HALF_H * (IMG_W / 32) * 4; // u16 indices of the non-empty mask words
constexpr int SM_MISC = SM_MLIST + HALF_H * (IMG_W / 32) * 2;
constexpr int SM_TOTAL = SM_MISC + 512;
static_assert(SM_TOTAL <= 227 * 1024, "shared memory budget");
// colour stage: the per-warp staging areas (640 floats each) start at hb0 and may run into the first rows of hb1 (contiguous),
// so the 1-row halo of the central differences sits behind them
constexpr int HALO1_OFF = 1024;
constexpr int IOST_OFF = HALO1_OFF + IMG_W; // per-warp background / RGB staging (192 floats each) behind the halo row
static_assert(NWARPS * 640 <= HB0_ROWS * IMG_W + HALO1_OFF && IOST_OFF + NWARPS * 192 <= HB1_ROWS * IMG_W, "staging area");
#ifndef TX_REC_F4
#define TX_REC_F4 8 // float4 per polynomial record in device memory: 8 = padded to one 128-byte line (5 are used)
#endif

struct Misc {
    uint64_t mbar;
    float xch_min[2];
    float red_f[NWARPS];
    unsigned red_u[3][NWARPS];
    int nlist;
    int bbox[4];      // contact bounding box of this half (image coordinates): row min, row max, col min, col max
    int bbox_peer[4]; // the peer CTA's box
};
static_assert(sizeof(Misc) <= 512, "misc");

int taxim_smem_bytes() { return SM_TOTAL; }

// ---- horizontal pass: warp per row, row in registers, shuffles ------------------------------------------------
// Lane i owns virtual columns v = 12*i + j - 32 (j = 0..11); virtual columns outside [0, 320) hold the reflected
// pixels (torch 'reflect'), so the correlation is uniform over the warp. PUSH: also store the result row into the
// peer CTA's halo buffer (row index = distance from the CTA boundary).
// Only the rows [la, lb] (local) can hold non-zero values (exact zeros elsewhere: a blur of zeros is +0.0f bit for bit),
// so only they are processed, interleaved over the warps; skipped rows inside the push range send zeros to the peer.
// LPR = lanes per row. 32: the warp holds the whole row (384 virtual columns incl. the reflected ones). 16: the non-zero
// columns [c0, c1] of the level's input, the blur radius on both sides and the alignment slack fit into 192 columns that do
// not reach the reflected border, so every warp processes TWO rows at once (one per half warp) on the window that starts at
// column `vbase`; the lanes wrap around inside their half, and everything a wrapped tap can reach is exactly zero, so the
// result is bit-identical to the full-width evaluation (same tap order, zero terms leave the accumulator unchanged).
// J = columns per lane (12 or 8), REFLECT = the window is the whole row incl. its reflected border (LPR = 32, J = 12 only).
template <int LPR, int J, bool REFLECT>
__device__ __forceinline__ void load_row(const float* rp, int v0, bool interior, float (&x)[J])
{
    if (interior) { // aligned 128-bit loads (conflict-free within each quarter warp)
#pragma unroll
        for (int e = 0; e < J / 4; ++e) {
            const float4 t = *reinterpret_cast<const float4*>(rp + v0 + 4 * e);
            x[4 * e + 0] = t.x; x[4 * e + 1] = t.y; x[4 * e + 2] = t.z; x[4 * e + 3] = t.w;
        }
    } else if (REFLECT) { // the six edge lanes hold the reflected columns
#pragma unroll
        for (int j = 0; j < J; ++j) {
            int v = v0 + j;
            v = v < 0 ? -v : (v > IMG_W - 1 ? 2 * (IMG_W - 1) - v : v);
            x[j] = rp[v];
        }
    } else { // window past the end of the row: those columns are zero / only feed outputs that are never stored
#pragma unroll
        for (int j = 0; j < J; ++j) x[j] = (v0 + j >= 0 && v0 + j < IMG_W) ? rp[v0 + j] : 0.0f;
    }
}

template <int L, int RAD, int LPR, int J, bool REFLECT>
__device__ __forceinline__ void hpass_rows(const TaximTaps& c_taps, float* plane, float* hb_remote, int warp, int lane, unsigned q, int la, int lb, int vbase)
{
    constexpr int D = (RAD + J - 1) / J;
    constexpr int RPW = 32 / LPR; // rows per warp
    const int sub = lane & (LPR - 1), half = lane / LPR;
    const int v0 = vbase + J * sub;
    const bool interior = v0 >= 0 && v0 + J - 1 < IMG_W;
    const int nact = lb - la + 1;
    const int ngrp = (nact + RPW - 1) / RPW; // groups of RPW consecutive rows
    if (warp >= ngrp) return;
    float xn[J];
    {
        const int r0 = min(la + warp * RPW + half, lb);
        load_row<LPR, J, REFLECT>(plane + r0 * IMG_W, v0, interior, xn);
    }
#pragma unroll 1
    for (int k = warp; k < ngrp; k += NWARPS) {
        const int rowu = la + k * RPW + half;
        const bool rvalid = rowu <= lb;
        const int row = min(rowu, lb);
        float* rp = plane + row * IMG_W;
        float x[J], acc[J];
#pragma unroll
        for (int j = 0; j < J; ++j) { x[j] = xn[j]; acc[j] = 0.0f; }
        if (k + NWARPS < ngrp) { // prefetch this warp's next row(s)
            const int rn = min(la + (k + NWARPS) * RPW + half, lb);
            load_row<LPR, J, REFLECT>(plane + rn * IMG_W, v0, interior, xn);
        }
#pragma unroll
        for (int d = -D; d <= D; ++d) {
#pragma unroll
            for (int j = 0; j < J; ++j) {
                if (J * d + j >= -RAD && J * d + j - (J - 1) <= RAD) {
                    const float y = (d == 0) ? x[j] : __shfl_sync(0xffffffffu, x[j], ((lane + d) & (LPR - 1)) | (lane & ~(LPR - 1)));
#pragma unroll
                    for (int m = 0; m < J; ++m) {
                        const int kk = J * d + j - m;
                        if (kk >= -RAD && kk <= RAD) acc[m] = __fmaf_rn(c_taps.t[L][0][kk + RAD], y, acc[m]);
                    }
                }
            }
        }
        // the shuffles above are warp-synchronous: every lane has consumed the old row before anyone overwrites it
        const int dist = q == 0 ? (HALF_H - 1 - row) : row; // distance from the CTA boundary
        const bool push = dist < RAD;
        if (rvalid) {
#pragma unroll
            for (int e = 0; e < J / 4; ++e) {
                const int vb = v0 + 4 * e;
                if (vb >= 0 && vb <= IMG_W - 4) {
                    const float4 t = make_float4(acc[4 * e + 0], acc[4 * e + 1], acc[4 * e + 2], acc[4 * e + 3]);
                    *reinterpret_cast<float4*>(rp + vb) = t;
                    if (push) *reinterpret_cast<float4*>(hb_remote + dist * IMG_W + vb) = t;
                }
            }
        }
    }
}

// c0 / c1: non-zero columns of the level's input (image coordinates)
template <int L, int RAD>
__device__ __forceinline__ void hpass(const TaximTaps& c_taps, float* plane, float* hb_remote, int tid, int warp, int lane, unsigned q, int la, int lb,
                                      int c0, int c1)
{
    // zero halo rows for the skipped rows of the push range (the halo buffers are reused across levels)
    for (int i = tid; i < RAD * (IMG_W / 4); i += NTHREADS) {
        const int dist = i / (IMG_W / 4);
        const int row = q == 0 ? (HALF_H - 1 - dist) : dist;
        if (row < la || row > lb)
            reinterpret_cast<float4*>(hb_remote + dist * IMG_W)[i - dist * (IMG_W / 4)] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // Windowed variants: when the non-zero columns [c0, c1] stay more than RAD away from both image borders, no reflected tap can
    // reach a non-zero value, so the row is evaluated on a window that starts at `vbase` with zeros outside the image -- no
    // reflected loads (which cost the six edge lanes a divergent scalar path), and fewer columns:
    //   192 columns (16 lanes x 12): two rows per warp;   256 columns (32 lanes x 8).
    // Halo rows written by a windowed variant keep stale columns outside the window; the vertical pass of this level only
    // reads the columns [c0 - RAD, c1 + RAD], which the window covers. Bit-identical to the full-width evaluation (same tap
    // order, zero terms leave the accumulator unchanged).
    const bool inner = c0 > RAD && c1 < IMG_W - 1 - RAD;
    const int need = (c1 - c0 + 1) + 2 * RAD + 4; // columns from the 4-aligned window start to c1 + RAD
    if (inner && need + 4 <= 192)
        hpass_rows<L, RAD, 16, 12, false>(c_taps, plane, hb_remote, warp, lane, q, la, lb, (c0 - RAD) & ~3);
    else if (inner && need <= 256)
        hpass_rows<L, RAD, 32, 8, false>(c_taps, plane, hb_remote, warp, lane, q, la, lb, (c0 - RAD) & ~3);
    else
        hpass_rows<L, RAD, 32, 12, true>(c_taps, plane, hb_remote, warp, lane, q, la, lb, -32);
}

// ---- vertical pass: thread per column, sliding register window, in place --------------------------------------
// Position t = 0..119 counts rows from the IMAGE edge of this CTA's half (q = 0: t = local row, marching down;
// q = 1: t = 119 - local row, marching up). Positions < 0 are the reflected rows, positions >= 120 come from
// the halo buffer (rows of the peer CTA by distance from the boundary).
template <int L, int RAD, int R>
__device__ __forceinline__ void vpass(const TaximTaps& c_taps, float* plane, const float* hb, int tid, unsigned q, int ca, int ncols, int ta, int tb)
{
    // only columns [ca, ca + ncols) and positions [ta, tb] can become non-zero; everything else stays exactly +0.0f
    if (tid >= ncols || tb < ta) return;
    constexpr int WN = R + 2 * RAD;
    static_assert(HALF_H % R == 0, "block size");
    const int col = ca + tid;
    float win[WN];
    float* p0 = plane + (q ? (HALF_H - 1) * IMG_W : 0) + col;
    const float* hbc = hb + col;
    const int S = q ? -IMG_W : IMG_W;
    const int b0 = ta / R, b1 = tb / R;
#pragma unroll
    for (int i = 0; i < WN; ++i) {
        const int t = b0 * R - RAD + i;
        const float* src = (t < 0) ? (p0 + (-t) * S) : ((t < HALF_H) ? (p0 + t * S) : (hbc + min(t - HALF_H, RAD - 1) * IMG_W));
        win[i] = *src;
    }
#pragma unroll 1
    for (int b = b0; b <= b1; ++b) {
        float acc[R];
        // centre-outward, symmetric pair summed first: independent of the marching direction
#pragma unroll
        for (int m = 0; m < R; ++m) acc[m] = __fmul_rn(c_taps.t[L][1][RAD], win[m + RAD]);
#pragma unroll
        for (int d = 1; d <= RAD; ++d) {
#pragma unroll
            for (int m = 0; m < R; ++m)
                acc[m] = __fmaf_rn(c_taps.t[L][1][RAD + d], __fadd_rn(win[m + RAD - d], win[m + RAD + d]), acc[m]);
        }
        float* po = p0 + (b * R) * S;
#pragma unroll
        for (int m = 0; m < R; ++m) po[m * S] = acc[m];
        if (b < b1) {
#pragma unroll
            for (int i = 0; i < 2 * RAD; ++i) win[i] = win[i + R];
            const int tbase = (b + 1) * R - RAD;
#pragma unroll
            for (int i = 2 * RAD; i < WN; ++i) {
                const int t = tbase + i;
                // t >= 120 -> halo row (t - 120); only rows < RAD are ever used by outputs < 120
                const float* src = (t < HALF_H) ? (p0 + t * S) : (hbc + min(t - HALF_H, RAD - 1) * IMG_W);
                win[i] = *src;
            }
        }
    }
}

// ---- masked re-imposition: plane[mask] = min(h, gel) recomputed from the input frame (L2 hit) ------------------
// `list` holds the indices of the mask words with at least one bit set (built once per frame).
__device__ __forceinline__ void reimpose(float* plane, const unsigned* maskbits, const unsigned short* list, int nlist,
                                         const float* __restrict__ hm_half, const float* __restrict__ gel_half, float m,
                                         float press, int warp, int lane, float depth_clip, bool coherent)
{
    for (int i0 = warp * 4; i0 < nlist; i0 += NWARPS * 4) {
        float v[4];
        int idx[4];
        bool on[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u;
            on[u] = false;
            if (i < nlist) {
                const int w = list[i];
                idx[u] = w * 32 + lane;
                on[u] = (maskbits[w] >> lane) & 1u;
                if (on[u]) {
                    v[u] = coherent ? hm_half[idx[u]] : __ldg(hm_half + idx[u]); // coherent: written by this kernel (resized map)
                    if (depth_clip > 0.0f) v[u] = __fmul_rn(isinf(v[u]) ? depth_clip : v[u], 1000.0f);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (on[u]) {
                const float h = __fadd_rn(__fadd_rn(v[u], -m), -press);
                const float g = gel_half ? __ldg(gel_half + idx[u]) : 0.0f;
                plane[idx[u]] = fminf(h, g);
            }
        }
    }
}

#define TX_TICK(slot)                                                                                                 \
    do {                                                                                                              \
        if (tk && tid == 0) tk[(slot)] = clock64();                                                                   \
    } while (0)

// Flat part of the colour stage: the pixels outside the rectangle [ry0, ry1] x [xa, xb] (local rows / image columns) of
// this CTA's half copy the precomputed flat RGB. The rectangle is known before the pyramid starts, so the copy is cut into
// one slice per blur level and done by the warps that have no column in that level's vertical pass (they would idle).
struct FlatCopy {
    const float* src; // flat RGB of this half
    float* dst;       // RGB output of this half
    int ry0, ry1, xa, xb, total;
    unsigned done;    // bit l: slice l has been copied
    bool coherent_src; // the re-imposition source was written by this kernel (resized camera frame): no read-only loads
};
constexpr int FC_SLICES = 7;
constexpr int FC_SLICE_QUADS = (HALF_H * (IMG_W / 4) + FC_SLICES - 1) / FC_SLICES;

__device__ __forceinline__ void flat_copy_slice(const FlatCopy& fc, int slice, int first, int nthr)
{
    constexpr int QPR = IMG_W / 4;
    const int q1 = min((slice + 1) * FC_SLICE_QUADS, HALF_H * QPR);
#pragma unroll 2
    for (int qd = slice * FC_SLICE_QUADS + first; qd < q1; qd += nthr) {
        const int row = qd / QPR;
        const int x0 = (qd - row * QPR) * 4;
        if (fc.total > 0 && row >= fc.ry0 && row <= fc.ry1 && x0 >= fc.xa && x0 <= fc.xb) continue;
        const size_t pix = (size_t)row * IMG_W + x0;
        const float4* f4 = reinterpret_cast<const float4*>(fc.src + pix * 3);
        float4* o4 = reinterpret_cast<float4*>(fc.dst + pix * 3);
        const float4 t0 = __ldg(f4), t1 = __ldg(f4 + 1), t2 = __ldg(f4 + 2);
        o4[0] = t0; o4[1] = t1; o4[2] = t2;
    }
}

// Region = bounding box (image coordinates) outside of which the plane is exactly zero at the start of the level.
struct Region {
    int r0, r1, c0, c1;
};

template <int L, int RAD, int R, bool FINAL>
__device__ __forceinline__ void blur_level(const TaximTaps& c_taps, float* plane, float* hb_local, float* hb_remote, const unsigned* maskbits,
                                           const unsigned short* mlist, int nlist, const float* hm_half, const float* gel_half,
                                           float m, float press, int tid, int warp, int lane, unsigned q, Region& rg,
                                           cg::cluster_group& cluster, long long* tk, float depth_clip, FlatCopy& fc)
{
    const int base = (int)q * HALF_H;
    hpass<L, RAD>(c_taps, plane, hb_remote, tid, warp, lane, q, max(rg.r0 - base, 0), min(rg.r1 - base, HALF_H - 1), rg.c0, rg.c1);
    TX_TICK(4 + 4 * L + 0);
    cluster.sync(); // rows + pushed halo rows visible in both CTAs
    TX_TICK(4 + 4 * L + 1);
    {
        const int ca = max(rg.c0 - RAD, 0), cb = min(rg.c1 + RAD, IMG_W - 1);
        const int ro0 = max(rg.r0 - RAD, 0), ro1 = min(rg.r1 + RAD, IMG_H - 1); // output rows (image)
        // position t counts from the image edge of this half: q = 0: t = image row; q = 1: t = 239 - image row
        const int ta = q == 0 ? ro0 : max(IMG_H - 1 - ro1, 0);
        const int tb = q == 0 ? min(ro1, HALF_H - 1) : min(IMG_H - 1 - ro0, HALF_H - 1);
        vpass<L, RAD, R>(c_taps, plane, hb_local, tid, q, ca, cb - ca + 1, ta, tb);
        // the warps without a column copy this level's slice of the flat RGB meanwhile
        const int nv = tb < ta ? 0 : ((cb - ca + 1 + 31) & ~31);
        if (nv < NTHREADS) {
            if (tid >= nv) flat_copy_slice(fc, L, tid - nv, NTHREADS - nv);
            fc.done |= 1u << L;
        }
        rg.r0 = ro0; rg.r1 = ro1; rg.c0 = ca; rg.c1 = cb;
    }
    __syncthreads();
    TX_TICK(4 + 4 * L + 2);
    if (!FINAL) {
        reimpose(plane, maskbits, mlist, nlist, hm_half, gel_half, m, press, warp, lane, depth_clip, fc.coherent_src);
        __syncthreads();
    }
    TX_TICK(4 + 4 * L + 3);
}

// RGB of a pixel with exactly zero gradient (mag = 0, dir = 0), for every pixel position: computed once per calibration
// with the same operation sequence the fused kernel uses, so copying it is bit-identical to evaluating it
__global__ void flat_rgb_kernel(const TaximArgs p, float* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= IMG_H * IMG_W) return;
    const float PI_F = 3.14159265358979323846f;
    const int id_flat = min(max((int)floorf(__fmul_rn(__fadd_rn(0.0f, PI_F), p.inv_ybin)), 0), p.nb - 1);
    const float4* pf = p.poly + (size_t)id_flat * TX_REC_F4;
    const float4 a0 = __ldg(pf), a1 = __ldg(pf + 1), a2 = __ldg(pf + 2), a3 = __ldg(pf + 3), a4 = __ldg(pf + 4);
    const float cf[20] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y,
                          a2.z, a2.w, a3.x, a3.y, a3.z, a3.w, a4.x, a4.y, a4.z, a4.w};
    const int y = i / IMG_W, x = i - y * IMG_W;
    const float xf = __fmul_rn((float)x, p.fx), yf = __fmul_rn((float)y, p.fy);
    const float bgv[3] = {p.bg_hwc[3 * i], p.bg_hwc[3 * i + 1], p.bg_hwc[3 * i + 2]};
    float o[3];
    poly_rgb(cf, xf, yf, __fmul_rn(xf, xf), __fmul_rn(yf, yf), __fmul_rn(xf, yf), bgv, o);
    out[3 * i] = o[0]; out[3 * i + 1] = o[1]; out[3 * i + 2] = o[2];
}

cudaError_t launch_flat_rgb(const TaximArgs& a, float* flat_rgb, cudaStream_t s)
{
    flat_rgb_kernel<<<(IMG_H * IMG_W + 255) / 256, 256, 0, s>>>(a, flat_rgb);
    return cudaGetLastError();
}

// ---- the fused kernel -------------------------------------------------------------------------------------------
// MODE bit 0: the input is the camera DEPTH image in metres (inf = no hit); bit 1: the input has the camera resolution
// [Hc][Wc] != 240 x 320 and is resized (bilinear, torchvision F.resize semantics) in the load stage.
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1) taxim_fused_kernel(const __grid_constant__ TaximArgs p, const __grid_constant__ TaximTaps c_taps)
{
    extern __shared__ __align__(128) unsigned char smem[];
    float* plane = reinterpret_cast<float*>(smem + SM_PLANE);
    float* hb0 = reinterpret_cast<float*>(smem + SM_HB0);
    float* hb1 = reinterpret_cast<float*>(smem + SM_HB1);
    unsigned* maskbits = reinterpret_cast<unsigned*>(smem + SM_MASK);
    unsigned short* mlist = reinterpret_cast<unsigned short*>(smem + SM_MLIST);
    Misc* misc = reinterpret_cast<Misc*>(smem + SM_MISC);

    cg::cluster_group cluster = cg::this_cluster();
    const unsigned q = cluster.block_rank(); // 0 = rows 0..119, 1 = rows 120..239
    const int n = blockIdx.x >> 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t half_off = (size_t)n * IMG_H * IMG_W + (size_t)q * HALF_H * IMG_W;
    constexpr bool DEPTH = (MODE & 1) != 0, LOWRES = (MODE & 2) != 0;
    const float* hm_half = (LOWRES ? p.up_scratch : p.hm) + half_off; // full-resolution height map of this half (global)
    const float* gel_half = p.gel ? p.gel + (size_t)q * HALF_H * IMG_W : nullptr;
    float* hb0_remote = cluster.map_shared_rank(hb0, q ^ 1u);
    float* hb1_remote = cluster.map_shared_rank(hb1, q ^ 1u);
    Misc* misc_remote = cluster.map_shared_rank(misc, q ^ 1u);
    long long* tk = p.ticks ? p.ticks + (size_t)blockIdx.x * 40 : nullptr; // optional per-CTA phase clock stamps
    TX_TICK(0);

    // ---- stage the half frame with bulk-async copies (TMA) ----------------------------------------------------
    if (tid == 0) {
        mbar_init(&misc->mbar, 1);
        fence_mbar_init();
        fence_proxy_async();
        misc->nlist = 0;
        misc->bbox[0] = IMG_H; misc->bbox[1] = -1; misc->bbox[2] = IMG_W; misc->bbox[3] = -1;
    }
    __syncthreads();
    float m_cam = 0.0f; // LOWRES: minimum of the camera-resolution height map (what compute_indentation_depth sees)
    if (!LOWRES) {
        if (tid == 0) {
            constexpr uint32_t CH = HALF_H * IMG_W * 4 / 4; // 4 chunks of 38,400 B
            mbar_expect_tx(&misc->mbar, HALF_H * IMG_W * 4);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                bulk_g2s(reinterpret_cast<unsigned char*>(plane) + c * CH,
                         reinterpret_cast<const unsigned char*>(hm_half) + c * CH, CH, &misc->mbar);
        }
        mbar_wait(&misc->mbar, 0);
        TX_TICK(1);

        // ---- optional fused GelSightSensor._get_height_map (ref: gelsight_sensor.py:581-593): depth [m] -> height map [mm] ----
        if (DEPTH) {
            float4* p4w = reinterpret_cast<float4*>(plane);
            float4* o4w = p.hm_out ? reinterpret_cast<float4*>(p.hm_out + half_off) : nullptr;
            for (int i = tid; i < HALF_H * IMG_W / 4; i += NTHREADS) {
                float4 t = p4w[i];
                t.x = __fmul_rn(isinf(t.x) ? p.clip_max_m : t.x, 1000.0f);
                t.y = __fmul_rn(isinf(t.y) ? p.clip_max_m : t.y, 1000.0f);
                t.z = __fmul_rn(isinf(t.z) ? p.clip_max_m : t.z, 1000.0f);
                t.w = __fmul_rn(isinf(t.w) ? p.clip_max_m : t.w, 1000.0f);
                p4w[i] = t;
                if (o4w) o4w[i] = t;
            }
            __syncthreads();
        }
    } else {
        // ---- camera resolution != tactile resolution (ref: taxim_sim.py:88-89, F.resize bilinear): the whole camera frame is
        // staged in shared memory (hb0 is free until the first blur level), converted like _get_height_map when it is a depth
        // image, and every pixel of this half is interpolated from it: horizontal taps first, then vertical, fmaf-accumulated
        // exactly like aten's separable antialias kernel (the tables hold first tap + two weights per output index).
        float* lo = hb0;
        const int npx = p.Hc * p.Wc;
        const float* src = p.hm + (size_t)n * npx;
        float lmin = __int_as_float(0x7f800000);
        for (int i = tid; i < npx; i += NTHREADS) {
            float v = __ldg(src + i);
            if (DEPTH) v = __fmul_rn(isinf(v) ? p.clip_max_m : v, 1000.0f);
            lo[i] = v;
            lmin = fminf(lmin, v);
        }
        lmin = warp_min(lmin);
        if (lane == 0) misc->red_f[warp] = lmin;
        __syncthreads();
        m_cam = misc->red_f[0];
#pragma unroll
        for (int w = 1; w < NWARPS; ++w) m_cam = fminf(m_cam, misc->red_f[w]);
        float* up_half = p.up_scratch + half_off;
        for (int idx = tid; idx < HALF_H * IMG_W; idx += NTHREADS) {
            const int row = idx / IMG_W, X = idx - row * IMG_W;
            const int Y = (int)q * HALF_H + row;
            const int y0 = __ldg(p.rs_y0 + Y), x0 = __ldg(p.rs_x0 + X);
            const float2 wy = __ldg(p.rs_wy + Y), wx = __ldg(p.rs_wx + X);
            const int y1 = min(y0 + 1, p.Hc - 1), x1 = min(x0 + 1, p.Wc - 1);
            const float t0 = __fmaf_rn(wx.y, lo[y0 * p.Wc + x1], __fmul_rn(wx.x, lo[y0 * p.Wc + x0]));
            const float t1 = __fmaf_rn(wx.y, lo[y1 * p.Wc + x1], __fmul_rn(wx.x, lo[y1 * p.Wc + x0]));
            const float v = __fmaf_rn(wy.y, t1, __fmul_rn(wy.x, t0));
            plane[idx] = v;
            up_half[idx] = v; // the masked re-imposition reads the resized map back (L2)
        }
        __syncthreads(); // also: red_f is reused by the frame minimum below
        TX_TICK(1);
    }

    // ---- frame minimum (ref: taxim_torch.py:441, taxim_sim.py:116-117) ----------------------------------------
    float mloc = __int_as_float(0x7f800000);
    {
        const float4* p4 = reinterpret_cast<const float4*>(plane);
        for (int i = tid; i < HALF_H * IMG_W / 4; i += NTHREADS) {
            const float4 t = p4[i];
            mloc = fminf(fminf(mloc, fminf(t.x, t.y)), fminf(t.z, t.w));
        }
    }
    mloc = warp_min(mloc);
    if (lane == 0) misc->red_f[warp] = mloc;
    __syncthreads();
    if (tid == 0) {
        float v = misc->red_f[0];
#pragma unroll
        for (int w = 1; w < NWARPS; ++w) v = fminf(v, misc->red_f[w]);
        misc->xch_min[q] = v;
        misc_remote->xch_min[q] = v;
    }
    cluster.sync();
    const float m = fminf(misc->xch_min[0], misc->xch_min[1]);
    TX_TICK(2);

    // ---- indentation depth (explicit, or fused: ref taxim_sim.py:115-131) --------------------------------------
    float press;
    if (p.press_in) {
        press = __ldg(p.press_in + n);
    } else {
        float d = __fdiv_rn(LOWRES ? m_cam : m, 1000.0f);
        d = __fadd_rn(d, -p.gelpad_min);
        d = d < 0.0f ? 0.0f : d;
        press = (d <= p.gelpad_h) ? __fmul_rn(__fadd_rn(p.gelpad_h, -d), 1000.0f) : 0.0f;
    }
    if (p.depth_out && q == 0 && tid == 0) p.depth_out[n] = press;

    // ---- shifted height map, contact mask, joined map (ref: taxim_torch.py:441-461) -----------------------------
    // min over the frame of ((hm - m) - press) is exactly -press, so pressing_depth_mm == press.
    const float thr = __fmul_rn(-press, p.contact_scale);
    unsigned cnt = 0, srow = 0, scol = 0;
    // One warp iteration = 128 consecutive pixels (4 mask words): every lane owns 4 pixels as one float4 (128-bit shared-memory
    // accesses), the 4-bit mask nibbles of the 8 lanes of a word are merged with three shuffles. The contact bounding box and
    // the FOTS sums are kept in per-lane registers and reduced once per warp at the end.
    constexpr int NSEG = HALF_H * IMG_W / 128;
    int bb_r0 = IMG_H, bb_r1 = -1, bb_c0 = IMG_W, bb_c1 = -1;
    {
        float4* plane4 = reinterpret_cast<float4*>(plane);
        const float4* gel4 = gel_half ? reinterpret_cast<const float4*>(gel_half) : nullptr;
        const int sub = lane & 7, grp = lane >> 3;
#pragma unroll 2
        for (int sg = warp; sg < NSEG; sg += NWARPS) {
            const int i4 = sg * 32 + lane;
            const float4 pv = plane4[i4];
            const float4 gv = gel4 ? __ldg(gel4 + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
            const int w = sg * 4 + grp;                 // mask word of this lane's pixels
            const int lrow = w / (IMG_W / 32);          // local row
            const int x0 = (w - lrow * (IMG_W / 32)) * 32 + sub * 4;
            const float hx = __fadd_rn(__fadd_rn(pv.x, -m), -press), hy = __fadd_rn(__fadd_rn(pv.y, -m), -press);
            const float hz = __fadd_rn(__fadd_rn(pv.z, -m), -press), hw = __fadd_rn(__fadd_rn(pv.w, -m), -press);
            const float4 j = make_float4(fminf(hx, gv.x), fminf(hy, gv.y), fminf(hz, gv.z), fminf(hw, gv.w));
            plane4[i4] = j;
            const unsigned c0 = hx < 0.0f, c1 = hy < 0.0f, c2 = hz < 0.0f, c3 = hw < 0.0f;
            const unsigned m0 = (__fadd_rn(j.x, -gv.x) < thr) & c0, m1 = (__fadd_rn(j.y, -gv.y) < thr) & c1;
            const unsigned m2 = (__fadd_rn(j.z, -gv.z) < thr) & c2, m3 = (__fadd_rn(j.w, -gv.w) < thr) & c3;
            const unsigned nb = m0 | (m1 << 1) | (m2 << 2) | (m3 << 3);
            const unsigned cn = c0 | (c1 << 1) | (c2 << 2) | (c3 << 3);
            unsigned wb = nb << (4 * sub);
            wb |= __shfl_xor_sync(0xffffffffu, wb, 1);
            wb |= __shfl_xor_sync(0xffffffffu, wb, 2);
            wb |= __shfl_xor_sync(0xffffffffu, wb, 4);
            const bool leader = sub == 0;
            if (leader) maskbits[w] = wb;
            const unsigned bal = __ballot_sync(0xffffffffu, leader && wb != 0u);
            if (bal) { // append the non-empty words to the list (one shared atomic per warp iteration)
                int pos = 0;
                if (lane == 0) pos = atomicAdd(&misc->nlist, __popc(bal));
                pos = __shfl_sync(0xffffffffu, pos, 0);
                if (leader && wb != 0u) mlist[pos + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)w;
            }
            if (cn) { // contact bounding box (the joined map is non-zero exactly where h < 0)
                const int rrow = (int)(q * HALF_H) + lrow;
                bb_r0 = min(bb_r0, rrow);
                bb_r1 = max(bb_r1, rrow);
                bb_c0 = min(bb_c0, x0 + __ffs(cn) - 1);
                bb_c1 = max(bb_c1, x0 + 31 - __clz(cn));
            }
            if (nb) {
                const unsigned k = __popc(nb);
                cnt += k;
                srow += k * ((unsigned)(q * HALF_H) + (unsigned)lrow);
                scol += k * (unsigned)x0 + m1 + 2u * m2 + 3u * m3;
            }
            if (p.mask_out)
                *reinterpret_cast<uchar4*>(p.mask_out + half_off + (size_t)i4 * 4) = make_uchar4(m0, m1, m2, m3);
        }
    }
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            bb_r0 = min(bb_r0, __shfl_xor_sync(0xffffffffu, bb_r0, o));
            bb_r1 = max(bb_r1, __shfl_xor_sync(0xffffffffu, bb_r1, o));
            bb_c0 = min(bb_c0, __shfl_xor_sync(0xffffffffu, bb_c0, o));
            bb_c1 = max(bb_c1, __shfl_xor_sync(0xffffffffu, bb_c1, o));
        }
    }
    if (lane == 0 && bb_r1 >= 0) {
        atomicMin(&misc->bbox[0], bb_r0);
        atomicMax(&misc->bbox[1], bb_r1);
        atomicMin(&misc->bbox[2], bb_c0);
        atomicMax(&misc->bbox[3], bb_c1);
    }
    if (p.aux_sums) {
        cnt = warp_sum_u32(cnt);
        srow = warp_sum_u32(srow);
        scol = warp_sum_u32(scol);
        if (lane == 0) {
            misc->red_u[0][warp] = cnt;
            misc->red_u[1][warp] = srow;
            misc->red_u[2][warp] = scol;
        }
    }
    __syncthreads();
    if (p.aux_sums && tid < 3) {
        unsigned s = 0;
#pragma unroll
        for (int w = 0; w < NWARPS; ++w) s += misc->red_u[tid][w];
        p.aux_sums[((size_t)n * 2 + q) * 4 + tid] = s;
    }
    const int nlist = misc->nlist;
    if (tid < 4) misc_remote->bbox_peer[tid] = misc->bbox[tid];
    cluster.sync();
    Region rg;
    rg.r0 = min(misc->bbox[0], misc->bbox_peer[0]);
    rg.r1 = max(misc->bbox[1], misc->bbox_peer[1]);
    rg.c0 = min(misc->bbox[2], misc->bbox_peer[2]);
    rg.c1 = max(misc->bbox[3], misc->bbox_peer[3]);
    if (p.gel != nullptr) { rg.r0 = 0; rg.r1 = IMG_H - 1; rg.c0 = 0; rg.c1 = IMG_W - 1; } // a gel map makes the plane dense

    TX_TICK(3);
    const float depth_clip = (DEPTH && !LOWRES) ? p.clip_max_m : 0.0f; // > 0: the re-imposition source is a depth image in metres
    const bool active = (p.gel != nullptr) || (press > 0.0f);
    // rectangle of the pixels that can have a non-zero gradient: the contact box grown by the sum of the blur radii (the
    // region after the last level) + 1 px for the central differences; the replicate-padded border rows / columns follow
    // their inner neighbour. LOCAL rows / image columns, columns aligned to the 4-pixel quads; empty without contact.
    FlatCopy fc;
    fc.src = p.flat_rgb + (size_t)q * HALF_H * IMG_W * 3;
    fc.dst = p.rgb + half_off * 3;
    fc.ry0 = 0; fc.ry1 = -1; fc.xa = 0; fc.xb = -1; fc.done = 0u;
    fc.coherent_src = LOWRES;
    int rect_a0 = 0, rect_a1 = -1; // IMAGE rows of the frame's rectangle (both halves): the colour stage is shared between the CTAs
    if (active && !(p.dbg & 4)) {
        constexpr int GROW = 30 + 16 + 8 + 4 + 2 + 1 + 2;
        const int base = (int)q * HALF_H;
        const int fr0 = max(rg.r0 - GROW, 0), fr1 = min(rg.r1 + GROW, IMG_H - 1);
        const int fc0 = max(rg.c0 - GROW, 0), fc1 = min(rg.c1 + GROW, IMG_W - 1);
        const int a0 = fr0 - 1 <= 1 ? 0 : fr0 - 1;                     // row 0 samples row 1
        const int a1 = fr1 + 1 >= IMG_H - 2 ? IMG_H - 1 : fr1 + 1;     // row 239 samples row 238
        fc.ry0 = max(a0, base) - base;
        fc.ry1 = min(a1, base + HALF_H - 1) - base;
        rect_a0 = a0; rect_a1 = a1;
        fc.xa = fc0 - 1 <= 1 ? 0 : ((fc0 - 1) & ~3);                   // column 0 samples column 1
        fc.xb = fc1 + 1 >= IMG_W - 2 ? IMG_W - 1 : ((fc1 + 1) | 3);    // column 319 samples column 318
    }
    {
        const int nq_ = fc.xb >= fc.xa ? (fc.xb - fc.xa + 1) >> 2 : 0;
        fc.total = fc.ry1 >= fc.ry0 ? nq_ * (fc.ry1 - fc.ry0 + 1) : 0;
    }
    if (p.rect_out && tid == 0) {
        const int4 rd = fc.total > 0 ? make_int4(fc.ry0, fc.ry1, fc.xa, fc.xb) : make_int4(0, -1, 0, -1);
        reinterpret_cast<int4*>(p.rect_out)[blockIdx.x] = rd;
        if (p.rect_mc) // the descriptor travels with the pixels: every GPU learns which part of the frame it has to complete itself
            multimem_st_f4(reinterpret_cast<float*>(p.rect_mc) + 4 * (size_t)blockIdx.x,
                           make_float4(__int_as_float(rd.x), __int_as_float(rd.y), __int_as_float(rd.z), __int_as_float(rd.w)));
    }
    // ---- Gaussian pyramid with masked re-imposition + final blur (ref: taxim_torch.py:463-471) -----------------
    // Without contact and with a flat gel map the joined map is identically zero: every blur returns exact zeros.
    if (active) {
        blur_level<0, 30, 12, false>(c_taps, plane, hb0, hb0_remote, maskbits, mlist, nlist, hm_half, gel_half, m, press, tid, warp, lane, q, rg, cluster, tk, depth_clip, fc);
        blur_level<1, 16, 24, false>(c_taps, plane, hb1, hb1_remote, maskbits, mlist, nlist, hm_half, gel_half, m, press, tid, warp, lane, q, rg, cluster, tk, depth_clip, fc);
        blur_level<2, 8, 24, false>(c_taps, plane, hb0, hb0_remote, maskbits, mlist, nlist, hm_half, gel_half, m, press, tid, warp, lane, q, rg, cluster, tk, depth_clip, fc);
        blur_level<3, 4, 30, false>(c_taps, plane, hb1, hb1_remote, maskbits, mlist, nlist, hm_half, gel_half, m, press, tid, warp, lane, q, rg, cluster, tk, depth_clip, fc);
        blur_level<4, 2, 30, false>(c_taps, plane, hb0, hb0_remote, maskbits, mlist, nlist, hm_half, gel_half, m, press, tid, warp, lane, q, rg, cluster, tk, depth_clip, fc);
        blur_level<5, 1, 30, false>(c_taps, plane, hb1, hb1_remote, maskbits, mlist, nlist, hm_half, gel_half, m, press, tid, warp, lane, q, rg, cluster, tk, depth_clip, fc);
        blur_level<6, 2, 30, true>(c_taps, plane, hb0, hb0_remote, maskbits, mlist, nlist, hm_half, gel_half, m, press, tid, warp, lane, q, rg, cluster, tk, depth_clip, fc);
    }

    // ---- 1-row halo for the central differences, optional outputs ---------------------------------------------
    {
        const int brow = q == 0 ? HALF_H - 1 : 0;
        for (int x = tid; x < IMG_W; x += NTHREADS) hb1_remote[HALO1_OFF + x] = plane[brow * IMG_W + x];
    }
    if (p.deformed_out) {
        const float4* s4 = reinterpret_cast<const float4*>(plane);
        float4* d4 = reinterpret_cast<float4*>(p.deformed_out + half_off);
        for (int i = tid; i < HALF_H * IMG_W / 4; i += NTHREADS) d4[i] = s4[i];
    }
    if (p.aux_bmax) {
        float bm = -__int_as_float(0x7f800000);
        for (int i = tid; i < HALF_H * IMG_W; i += NTHREADS) bm = fmaxf(bm, plane[i]);
        bm = warp_max(bm);
        if (lane == 0) misc->red_f[warp] = bm;
        __syncthreads();
        if (tid == 0) {
            float v = misc->red_f[0];
#pragma unroll
            for (int w = 1; w < NWARPS; ++w) v = fmaxf(v, misc->red_f[w]);
            p.aux_bmax[(size_t)n * 2 + q] = v;
        }
        for (int k = tid; k < p.M; k += NTHREADS) {
            const int my = p.mk_y[k], mx = p.mk_x[k];
            const int ly = my - (int)q * HALF_H;
            if (ly >= 0 && ly < HALF_H && mx >= 0 && mx < IMG_W) {
                p.aux_b[(size_t)n * p.M + k] = plane[ly * IMG_W + mx];
                p.aux_m[(size_t)n * p.M + k] = (unsigned char)((maskbits[ly * (IMG_W / 32) + (mx >> 5)] >> (mx & 31)) & 1u);
            }
        }
    }
    cluster.sync();
    TX_TICK(32);

    // ---- normals -> bins -> polynomial -> + background -> clip -> NHWC (ref: taxim_torch.py:475-503, 243-258) ---
    // The deformed gel is exactly zero outside the region `rg`, so only the pixels of the rectangle rg + 1 px (central
    // differences; the replicate-padded border rows / columns follow their inner neighbour) can have a non-zero gradient:
    //   (1) outside the rectangle: copy the precomputed flat RGB (bit-identical to evaluating it);
    //   (2) inside: one warp per block of 32 consecutive pixels of a rectangle row, one pixel per lane: canonical atan / atan2,
    //       bins, then the polynomial record of every pixel is gathered COOPERATIVELY (5 lanes read one record contiguously;
    //       the records are padded to one 128-byte line each) into a per-warp staging area; the background and the RGB output
    //       of the block -- 384 contiguous bytes each -- move as 24 x 128-bit accesses through a small staging area instead
    //       of 3 x 32 scalar accesses with a 12-byte stride; the gather of the NEXT block is requested before the current
    //       block is evaluated (software pipeline).
    const float PI_F = 3.14159265358979323846f;
    float* stage = hb0 + warp * (32 * 20); // 32 records x 20 floats per warp; hb0 is free now (16 x 2560 B = 40,960 B)
    float* iost = hb1 + IOST_OFF + warp * 192; // per warp: 96 floats of background + 96 floats of RGB
    const float* halo1 = hb1 + HALO1_OFF;
    const int xa = fc.xa, xb = fc.xb;
    // (1) flat copy of everything outside the rectangle: the slices no idle warp has taken during the pyramid
    for (int sl = 0; sl < FC_SLICES; ++sl)
        if (!((fc.done >> sl) & 1u)) flat_copy_slice(fc, sl, tid, NTHREADS);
    // (2) the rectangle. Its rows are split between the two halves as the contact happens to lie (typically 60 / 40), so the
    // 32-pixel blocks of BOTH halves are numbered through (half 0 first) and each CTA evaluates one half of the blocks: the CTA
    // whose half holds less of the rectangle takes rows of its peer, reading the peer's final plane through distributed shared
    // memory (4 loads per pixel against ~250 instructions). Same arithmetic per pixel whoever evaluates it: bit-identical.
    const float* plane_peer = cluster.map_shared_rank(plane, q ^ 1u);
    const int nw = (fc.total > 0 || rect_a1 >= rect_a0) && xb >= xa ? xb - xa + 1 : 0; // rectangle width in pixels (multiple of 4)
    const int nbr = (nw + 31) >> 5;                                                     // 32-pixel blocks per rectangle row
    const int r0_lo = rect_a0, r0_hi = min(rect_a1, HALF_H - 1);                         // image rows of the rectangle in half 0
    const int r1_lo = max(rect_a0, HALF_H), r1_hi = rect_a1;                             // ... in half 1
    const int rows0 = nw > 0 ? max(r0_hi - r0_lo + 1, 0) : 0, rows1 = nw > 0 ? max(r1_hi - r1_lo + 1, 0) : 0;
    const int nblk0 = nbr * rows0, nblk_all = nbr * (rows0 + rows1);
#ifdef TX_NO_COLOUR_BALANCE
    const int g_begin = q ? nblk0 : 0, g_end = q ? nblk_all : nblk0; // every CTA its own half (A/B reference)
#else
    // even split, optionally biased towards the CTA that owns the blocks (debug flags bits 8..11 = bias in sixteenths: the CTA that
    // takes blocks of its peer also has the larger flat-copy remainder and pays the DSMEM reads)
    const int g_half = (nblk_all + 1) >> 1;
    const int g_split = g_half + (((nblk0 - g_half) * (int)((p.dbg >> 8) & 15)) >> 4);
    const int g_begin = q ? g_split : 0, g_end = q ? nblk_all : g_split;
#endif
    const float inv_nbr = nbr > 0 ? 1.0f / (float)nbr : 0.0f;
    const float* bg_img = p.bg_hwc;
    float* rgb_img = p.rgb + (size_t)n * IMG_H * IMG_W * 3;
    float* rgb_mc_img = p.rgb_mc ? p.rgb_mc + (size_t)n * IMG_H * IMG_W * 3 : nullptr;
    float4 rec[5], bg4 = make_float4(0.f, 0.f, 0.f, 0.f);
    int g_rec[5], g_part[5]; // lane-constant gather pattern: float4 index e * 32 + lane = record (idx / 5), part (idx % 5)
#pragma unroll
    for (int e = 0; e < 5; ++e) {
        const int idx = e * 32 + lane;
        g_rec[e] = idx / 5;
        g_part[e] = idx - g_rec[e] * 5;
    }
    int row_n = 0, xs_n = 0;
    // `row` is the IMAGE row from here on
    auto block_bin = [&](int bi, int& row, int& xs) -> int {
        const int h = bi >= nblk0 ? 1 : 0;          // half the block lies in
        const int li = bi - h * nblk0;
        const int rr = __float2int_rz(__fmul_rn((float)li + 0.5f, inv_nbr)); // == li / nbr (li < 2^12, error << 0.5 / nbr)
        row = (h ? r1_lo : r0_lo) + rr;
        xs = xa + 32 * (li - rr * nbr);
        const int x = min(xs + lane, xb);
        // replicate padding of the gradient maps: border pixels take the nearest interior pixel's (mag, dir)
        const int yy = min(max(row, 1), IMG_H - 2) - h * HALF_H; // local row (in half h) of the sampled pixel
        const int xx = min(max(x, 1), IMG_W - 2);
        const bool mine = (unsigned)h == q;
        const float* ctr = (mine ? plane : plane_peer) + yy * IMG_W + xx;
        // the row across the boundary between the halves: the peer's boundary row (halo1) for a pixel of this CTA's half, this
        // CTA's own boundary row for a pixel of the peer's half
        const float* across = mine ? halo1 + xx : plane + (q == 0 ? HALF_H - 1 : 0) * IMG_W + xx;
        const float* up = (yy - 1 >= 0) ? ctr - IMG_W : across;
        const float* dn = (yy + 1 < HALF_H) ? ctr + IMG_W : across;
        const float top = __fmul_rn(*up, p.inv_pixmm), bot = __fmul_rn(*dn, p.inv_pixmm);
        const float lef = __fmul_rn(ctr[-1], p.inv_pixmm), rig = __fmul_rn(ctr[1], p.inv_pixmm);
        const float gx = __fmul_rn(__fmul_rn(__fadd_rn(top, -bot), 0.5f), p.sy);
        const float gy = __fmul_rn(__fmul_rn(__fadd_rn(lef, -rig), 0.5f), p.sx);
        const float s2 = __fmaf_rn(gx, gx, __fmul_rn(gy, gy));
        // tt = sqrt(s2) is zero iff s2 is zero: a flat pixel (exact zero gradient) lands in bin (0, dir = 0)
        const float tt = __fsqrt_rn(s2);
        const float mag = atanf_c(tt);
        const float dir = (tt != 0.0f) ? atan2f_c(gx, gy) : 0.0f;
        int im = (int)floorf(__fmul_rn(mag, p.inv_xbin));
        int id = (int)floorf(__fmul_rn(__fadd_rn(dir, PI_F), p.inv_ybin));
        im = min(max(im, 0), p.nb - 1);
        id = min(max(id, 0), p.nb - 1);
        return im * p.nb + id; // record index
    };
    auto request = [&](int bin, int row, int xs) {
#pragma unroll
        for (int e = 0; e < 5; ++e) {
            const int bsrc = __shfl_sync(0xffffffffu, bin, g_rec[e]);
            rec[e] = __ldg(p.poly + (size_t)bsrc * TX_REC_F4 + g_part[e]);
        }
        // background of the block: 3 * nval contiguous floats, 128-bit loads by the first lanes
        const int nval = min(32, xb - xs + 1);
        if (4 * lane < 3 * nval) bg4 = __ldg(reinterpret_cast<const float4*>(bg_img + ((size_t)row * IMG_W + xs) * 3) + lane);
    };
    // Software pipeline, three blocks deep: iteration i stages the records of block i (requested one iteration ago), requests
    // the records of block i + 1 (its bins were computed one iteration ago), computes the bins of block i + 2 -- the long
    // dependent sqrt / atan / atan2 chain, which covers the L2 round trip of the request -- and evaluates block i.
    int bi = g_begin + warp;
    int row_c = 0, xs_c = 0;     // block i (records in flight / staged)
    int bin_n = 0;               // block i + 1 (bins known)
    if (bi < g_end) {
        const int b = block_bin(bi, row_c, xs_c);
        request(b, row_c, xs_c);
        if (bi + NWARPS < g_end) bin_n = block_bin(bi + NWARPS, row_n, xs_n);
    }
#pragma unroll 1
    for (; bi < g_end; bi += NWARPS) {
        const int row = row_c, xs = xs_c;
        const int nval = min(32, xb - xs + 1);
        __syncwarp();
        // stage the current records (one 80-byte record per lane) and the background of the block
#pragma unroll
        for (int e = 0; e < 5; ++e) reinterpret_cast<float4*>(stage)[e * 32 + lane] = rec[e];
        if (4 * lane < 3 * nval) reinterpret_cast<float4*>(iost)[lane] = bg4;
        // block i + 1: its requests go out now; block i + 2: bins
        if (bi + NWARPS < g_end) {
            row_c = row_n; xs_c = xs_n;
            request(bin_n, row_c, xs_c);
            if (bi + 2 * NWARPS < g_end) bin_n = block_bin(bi + 2 * NWARPS, row_n, xs_n);
        }
        __syncwarp();
        float cf[20];
#pragma unroll
        for (int e = 0; e < 5; ++e) {
            const float4 t = reinterpret_cast<const float4*>(stage)[lane * 5 + e];
            cf[4 * e] = t.x; cf[4 * e + 1] = t.y; cf[4 * e + 2] = t.z; cf[4 * e + 3] = t.w;
        }
        const float bgv[3] = {iost[3 * lane], iost[3 * lane + 1], iost[3 * lane + 2]};
        const float yf = __fmul_rn((float)row, p.fy);
        const float xf = __fmul_rn((float)(xs + lane), p.fx);
        float o[3];
        poly_rgb(cf, xf, yf, __fmul_rn(xf, xf), __fmul_rn(yf, yf), __fmul_rn(xf, yf), bgv, o);
        iost[96 + 3 * lane] = o[0]; iost[96 + 3 * lane + 1] = o[1]; iost[96 + 3 * lane + 2] = o[2];
        __syncwarp();
        if (4 * lane < 3 * nval) {
            const float4 v = reinterpret_cast<const float4*>(iost + 96)[lane];
            const size_t off = ((size_t)row * IMG_W + xs) * 3 + 4 * lane;
            if (rgb_mc_img) multimem_st_f4(rgb_mc_img + off, v); // fused all-gather: into every GPU's buffer (this one included)
            else *reinterpret_cast<float4*>(rgb_img + off) = v;
        }
    }
    // a CTA that took blocks of the other half may still be reading its peer's plane: nobody leaves before both are done (the
    // decision is the same in both CTAs: it only depends on the frame's rectangle)
    if (g_end - g_begin != (q ? nblk_all - nblk0 : nblk0)) cluster.sync();
    else __syncthreads();
    TX_TICK(33);
}

cudaError_t launch_taxim(const TaximArgs& a, const TaximTaps& taps, int N, cudaStream_t s)
{
    static bool attr_set_dev[64] = {}; // the attribute belongs to the (function, device) pair: one process may drive several GPUs
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 63;
    bool& attr_set = attr_set_dev[dev];
    if (!attr_set || dev == 63) {
        cudaError_t e = cudaFuncSetAttribute(taxim_fused_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(taxim_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(taxim_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(taxim_fused_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int mode = (a.input_is_depth ? 1 : 0) | (a.Hc > 0 ? 2 : 0);
    const dim3 grid(2 * N), block(NTHREADS);
    switch (mode) {
    case 0: taxim_fused_kernel<0><<<grid, block, SM_TOTAL, s>>>(a, taps); break;
    case 1: taxim_fused_kernel<1><<<grid, block, SM_TOTAL, s>>>(a, taps); break;
    case 2: taxim_fused_kernel<2><<<grid, block, SM_TOTAL, s>>>(a, taps); break;
    default: taxim_fused_kernel<3><<<grid, block, SM_TOTAL, s>>>(a, taps); break;
    }
    return cudaGetLastError();
}

int taxim_lowres_max_pixels() { return HB0_ROWS * IMG_W; }
int taxim_record_f4() { return TX_REC_F4; }

// ---- stand-alone indentation depth (ref: taxim_sim.py:115-131): one CTA per frame, HBM-bound -------------------
__global__ void __launch_bounds__(256) indentation_depth_kernel(const float* __restrict__ hm, float* __restrict__ out,
                                                                float gelpad_h, float gelpad_min)
{
    __shared__ float red[8];
    const float4* p4 = reinterpret_cast<const float4*>(hm + (size_t)blockIdx.x * IMG_H * IMG_W);
    float m = __int_as_float(0x7f800000);
    for (int i = threadIdx.x; i < IMG_H * IMG_W / 4; i += 256) {
        const float4 t = __ldg(p4 + i);
        m = fminf(fminf(m, fminf(t.x, t.y)), fminf(t.z, t.w));
    }
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

cudaError_t launch_indentation_depth(const float* hm, float* out, int N, float gelpad_h, float gelpad_min, cudaStream_t s)
{
    indentation_depth_kernel<<<N, 256, 0, s>>>(hm, out, gelpad_h, gelpad_min);
    return cudaGetLastError();
}

} // namespace tx
