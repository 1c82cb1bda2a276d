// tx_kernels.h -- internal launch interface between the C-ABI layer (tx_api.cu) and the kernels.
#pragma once
#include "../../include/tacex_b200.h"
#include <cuda_runtime.h>
#include <stdint.h>

namespace tx {

constexpr int IMG_H = 240;
constexpr int IMG_W = 320;
constexpr int HALF_H = 120;   // rows owned by one CTA of the 2-CTA cluster
#ifndef TX_NTHREADS
#define TX_NTHREADS 512
#endif
constexpr int NTHREADS = TX_NTHREADS; // 16 warps x 128 registers = the whole register file of the SM (480 / 384 also build)
constexpr int NWARPS = NTHREADS / 32;

struct TaximArgs {
    const float* hm;       // [N][240][320] mm
    const float* press_in; // [N] or nullptr (fused indentation depth)
    int input_is_depth;    // 1: `hm` holds the camera depth image in metres (inf = no hit): mm = (isinf ? clip_max : d) * 1000
    float clip_max_m;      // far clipping plane of the sensor camera [m]
    float* hm_out;         // optional [N][240][320]: the height map in mm (what GelSightSensor publishes as 'height_map')
    // camera resolution != tactile resolution (ref: taxim_sim.py:88-89): `hm` holds [N][Hc][Wc] frames, resized in the load stage
    int Hc, Wc;            // 0, 0: `hm` is [N][240][320]
    const int* rs_x0;      // [320] first horizontal tap
    const float2* rs_wx;   // [320] the two horizontal weights
    const int* rs_y0;      // [240] first vertical tap
    const float2* rs_wy;   // [240] the two vertical weights
    float* up_scratch;     // [N][240][320] resized height maps (read back by the masked re-imposition)
    const float* gel;      // [240][320] or nullptr (flat)
    const float4* poly;    // [nb][nb][32]: 3 channels x 6 coefficients, every record padded to one 128-byte line
    const float* bg_hwc;   // [240][320][3]
    const float* flat_rgb; // [240][320][3] RGB of a flat (zero-gradient) pixel: clip(poly(bin(0, 0)) + background)
    float* rgb;            // [N][240][320][3]
    int* rect_out;         // optional [N][2][4]: per half frame the rectangle (ry0, ry1 local rows, xa, xb columns) outside of
                           // which the frame equals the flat RGB image (empty: ry1 < ry0); used by the multi-GPU gather
    // fused observation all-gather (multi-GPU): NVSwitch MULTICAST mappings of this rank's block of the gathered buffers. When set,
    // the evaluated rectangle and the rectangle descriptors are stored through them (multimem.st: one store leaves the GPU, the
    // switch replicates it into every GPU's buffer, this one included) instead of into rgb / rect_out; the flat part stays local.
    float* rgb_mc;         // [N][240][320][3] or nullptr
    int* rect_mc;          // [N][2][4] or nullptr
    float* depth_out;      // [N] or nullptr
    float* deformed_out;   // [N][240][320] or nullptr
    unsigned char* mask_out; // [N][240][320] or nullptr
    // FOTS inputs recorded per env (all nullptr when no marker grid is configured)
    unsigned* aux_sums; // [N][2][4]  (count, sum_row, sum_col, -) per CTA half
    float* aux_bmax;    // [N][2]
    float* aux_b;       // [N][M]
    unsigned char* aux_m; // [N][M]
    const int* mk_x;
    const int* mk_y;
    int M;
    float inv_pixmm, sy, sx, fx, fy, contact_scale, gelpad_h, gelpad_min, inv_xbin, inv_ybin;
    int nb;
    int dbg;          // profiling experiments only (0 in production): bit0 skip RGB stores, bit1 skip bg loads, bit2 force flat path
    long long* ticks; // optional [2N][40] phase clock stamps (profiling builds of the host call only)
};

// blur taps of the specialised 240 x 320 kernel: [level][0 = x, 1 = y][tap], a kernel parameter (per handle)
constexpr int TX_FUSED_BLURS = 7;
struct TaximTaps {
    float t[TX_FUSED_BLURS][2][TX_MAX_TAPS];
};

// arbitrary-resolution variant (taxim_generic_kernel.cu): one CTA per frame, two planes + a byte mask in shared memory
constexpr int TXG_MAX_PIXELS = 160 * 120;
struct TaximGenericArgs {
    const float* hm;         // [N][H][W] height map (mm) or depth image (m, input_is_depth)
    const float* press_in;   // [N] or nullptr (fused indentation depth)
    const float* gel;        // [H][W] or nullptr
    const float4* poly;      // [nb][nb][5 float4]
    const float* bg_hwc;     // [H][W][3]
    const float* taps;       // [n_blurs][2][TX_MAX_TAPS] (x, y) in global memory
    float* rgb;              // [N][H][W][3]
    float* depth_out;        // [N] or nullptr
    float* deformed_out;     // [N][H][W] or nullptr
    unsigned char* mask_out; // [N][H][W] or nullptr
    float* hm_out;           // [N][H][W] or nullptr (input_is_depth only)
    int H, W, n_blurs, nb, input_is_depth;
    int ksx[8], ksy[8];
    float clip_max_m, inv_pixmm, sx, sy, fx, fy, contact_scale, gelpad_h, gelpad_min, inv_xbin, inv_ybin;
};

// shadow branch (taxim_shadow_kernel.cu)
struct ShadowArgs {
    const float* deformed;      // [n][240][320] deformed gel (mm), written by the fused kernel
    const unsigned char* mask;  // [n][240][320] contact mask
    const float* gel;           // [240][320] or nullptr
    const float4* poly;         // [nb][nb][5 float4]
    float* shadow;              // [n][3][240][320] scratch: scatter-min image, then min(polynomial colour, shadow)
    const float* table;         // [3][D][Hn][S] shadow table (+inf padded)
    const float* fan_cos;       // [D][F]
    const float* fan_sin;       // [D][F]
    int D, Hn, S, F, nb;
    int dil[4];                 // ky0, kx0, ky1, kx1 of the two dilation rounds
    float pixmm, inv_pixmm, sx, sy, fx, fy, inv_xbin, inv_ybin;
    float depth_0, height_precision, discretize_precision, step_x, step_y;
};
cudaError_t launch_shadow(const ShadowArgs& a, int n, float* t1, float* t2, float* rgb, const float* bg_hwc, const float* taps_sx,
                          int ks_sx, const float* taps_sy, int ks_sy, const float* taps_fx, int ks_fx, const float* taps_fy,
                          int ks_fy, cudaStream_t s);

constexpr int TX_MAX_PEERS = 15;
struct ObsPushArgs {
    const float* rgb_local;  // [N][240][320][3] this rank's frames
    const int* rect_local;   // [N][2][4]
    int N, n_peers;
    float* peer_rgb[TX_MAX_PEERS];  // the same block inside every peer's gathered buffer (peer-mapped addresses)
    int* peer_rect[TX_MAX_PEERS];
    float* mc_rgb;  // optional: the same block through the NVSwitch MULTICAST mapping of the gathered buffers (multimem.st:
    int* mc_rect;   // one store leaves the GPU, the switch replicates it to every peer); nullptr = one store per peer
};
struct ObsFillArgs {
    float* rgb_all;          // [N_total][240][320][3] this rank's gathered buffer
    const int* rect_all;     // [N_total][2][4]
    int* prev_rect;          // optional [N_total][2][4] inout: the rectangles this buffer held after its last fill
    const float* flat_rgb;   // [240][320][3]
    int N_total, skip_lo, skip_hi; // envs [skip_lo, skip_hi) are this rank's own
};

struct FotsArgs {
    const unsigned* aux_sums;
    const float* aux_bmax;
    const float* aux_b;
    const unsigned char* aux_m;
    const int* mk_x;
    const int* mk_y;
    const float* press;
    const float* theta;
    float* traj0;     // [N][4]
    int* traj_len;    // [N]
    float* markers;   // [N][2][M][2]
    int M, rows, cols;
    double lamb0, lamb1, lamb2, mm2pix, shear_max, theta_max;
};

// marker image / marker overlay (overlay_kernel.cu)
struct OverlayArgs {
    const float* markers;      // [N][2][M][2]
    const unsigned char* patch; // [10][10][12][12] anti-aliased dot patches of the chosen marker size (device)
    const float* rgb_in;       // [N][240][320][3] or nullptr
    float* rgb;                // [N][240][320][3] or nullptr (may alias rgb_in)
    unsigned char* marker_img; // [N][240][320] or nullptr
    unsigned char* rgb_u8;     // [N][240][320][3] or nullptr
    int M, apply;              // apply = 0: rgb / rgb_u8 without the marker modulation (plain uint8 conversion)
};
cudaError_t launch_marker_overlay(const OverlayArgs& a, int N, cudaStream_t s);

// antialiased bilinear resize (overlay_kernel.cu)
constexpr int TX_RS_TAPS = 8;
struct ResizeArgs {
    const float* src; // [N][Hi][Wi]
    float* dst;       // [N][Ho][Wo]
    int Hi, Wi, Ho, Wo;
    const int *fx, *cx, *fy, *cy; // first tap / tap count per output column / row
    const float *wx, *wy;         // [Wo][8] / [Ho][8] normalised weights
};
cudaError_t launch_resize_aa(const ResizeArgs& a, int N, cudaStream_t s);

// ---- gel FEM ---------------------------------------------------------------------------------------------------------
typedef tx_fem_indenter FemIndenter;
typedef tx_fem_stats FemStats;

// Broad phase of the mesh indenter: a uniform grid in the INDENTER'S frame, built once on the host (the soup is rigid, so nothing is
// rebuilt per step as the reference's LBVH is). Every primitive (triangle / vertex / edge) is listed in exactly ONE cell, that of its
// reference point (centroid / position / mid-point); a query visits the cells its box -- inflated by the search radius and by rmax,
// the largest distance of a primitive's points from its reference point -- overlaps. A 1 x 1 x 1 grid is the brute-force loop.
struct MeshGrid {
    const int* start; // [nx * ny * nz + 1]
    const int* ids;   // primitive ids, ascending within a cell
    int nx, ny, nz;
    double lo[3], inv[3]; // cell = floor((x - lo) * inv), clamped
    double rmax;
};

struct FemArgs {
    int N, V, T, A, S;
    const int* tets;       // [T][4]
    const double* Dm_inv;  // [T][9]
    const double* vol;     // [T]   elastic rest "volume" (det Dm or det Dm / 6)
    const double* mass;    // [V]
    const int* attach;     // [A]
    const int* surf;       // [S]
    double* x; double* v; double* x_prev; // [N][V][3]
    const double* aim;     // [N][A][3]
    const FemIndenter* ind_prev; const FemIndenter* ind_next; // [N]
    FemStats* stats;       // [N] or nullptr
    double* tet_scratch;   // [grid][102][576] per-tet contributions of one chunk of tets (one tet per thread): 12 gradient, 4 diagonal and 6 off-diagonal 3x3 blocks
    double* val_scratch;   // [grid][9][nE - n_s] off-diagonal blocks that do not fit in shared memory (L2-resident)
    double* xt_scratch;    // [grid][3V] predicted positions
    const int* row_start;  // [nchunks+1][576] per chunk of 576 tets: first entry of `adj` (ascending tets) of every row
    const int* adj;        // [4T]
    int nE, n_s, nslots;   // edges (i < j) of the vertex graph, edges kept in shared memory, ELL width
    int chunk;             // tets per assembly chunk (<= FEM threads, multiple of 32): row_start / edge_start are built for it
    const int* edge_start; // [nchunks+1][nE] per chunk: first entry of `edge_adj` (ascending tets) of every edge
    const int* edge_adj;   // entries = tet << 4 | pair slot << 1 | transpose
    const int* ell;        // [nslots][FEM threads] row -> (neighbour j | transposed << 12 | edge << 13), -1 = empty
    const int* attach_of;  // [V] index into attach[] or -1
    const int* surf_of;    // [V] index into surf[] or -1
    const double* mesh_tri; // [mesh_n][9] triangles of the prescribed mesh indenter (type 2) in its local frame, or nullptr
    const double* mesh_box; // [mesh_n][6] their boxes (lo, hi)
    int mesh_n;
    // second half of the vertex-face contact (tx_fem_set_contact_surface): the indenter mesh's unique vertices against gel triangles
    const double* mesh_vert; // [mesh_nv][3] local frame
    int mesh_nv;
    const int* ctri;           // [n_ctri][3] gel contact triangles (one per thread: n_ctri <= FEM threads), or nullptr
    int n_ctri;
    const int* ctri_row_start; // [V + 1] entries of ctri_row_adj per vertex
    const int* ctri_row_adj;   // triangle << 2 | local vertex
    const int* ctri_edge_start; // [nE + 1] entries of ctri_edge_adj per mesh edge
    const int* ctri_edge_adj;  // triangle << 2 | pair (0: (0,1), 1: (0,2), 2: (1,2))
    // edge-edge candidates: the unique edges of the contact triangles (one per thread) against the unique edges of the indenter mesh
    const int* cedge;          // [n_cedge][2] gel vertex ids (i < j)
    const double* cedge_len2;  // [n_cedge] squared rest lengths (mollifier threshold)
    int n_cedge;
    const int* cedge_row_start; // [V + 1]
    const int* cedge_row_adj;  // contact edge << 1 | local vertex
    const int* edge_cedge;     // [nE] mesh edge -> contact edge or -1
    const int* mesh_edge;      // [mesh_ne][2] ids into mesh_vert
    int mesh_ne;
    MeshGrid grid_tri, grid_vert, grid_edge; // broad phase over the mesh's triangles / unique vertices / unique edges
    int dbg_mode;          // 0: cycles[3..5] = assembly sub-phases, 1: cycles[3] = SpMV, cycles[4] = rest of the PCG iteration
    long long* dbg_cycles; // optional [grid][6] phase cycle counters (grad_hess, pcg, line search, tets, vertices, edges)
    double dt, gravity[3], mu, lambda, attach_strength, d_hat, kappa, velocity_tol, pcg_tol_rate, friction_mu, eps_velocity;
    int newton_max_iter, pcg_max_iter_ratio, ls_max_iter, substep;
};

struct FemMarkerArgs {
    int V, M;
    const int* tri;        // [M][3] vertex ids of the surface triangle carrying marker k
    const double* weights; // [M][3] barycentric weights
    const double* x_rest;  // [V][3]
    const double* x;       // [N][V][3]
    float* out;            // [N][2][M][2]
    double cam_R[9], cam_t[3], fx, fy, cx, cy;
    int normalize;         // ret /= img_w / 2; ret -= 1 (gen_marker_flow's `normalize`)
    double half_w;
    int zero_all;          // no marker survived the reference's uv mask: the flow is all zeros
};

// FEM gel surface -> height map (raster_kernel.cu)
struct RasterArgs {
    const double* x;     // [N][V][3] FEM positions
    const int* tris;     // [n_tris][3] top-surface triangles
    float* hm;           // [N][H][W] height map, mm
    int V, n_tris, H, W;
    double pitch, ox, oy; // pixel pitch [m], pad-frame position of the image centre
    double cam_z;        // z of the camera plane in the pad frame [m]
    float far_mm;        // far clipping plane [mm] (pixels no triangle covers)
};
cudaError_t launch_heightmap(const RasterArgs& a, int N, cudaStream_t s);
cudaError_t launch_attachment_aim(const float* pose, const float* offsets, int N, int A, int per_env_offsets, double* aim, cudaStream_t s);

size_t fem_smem_bytes(int V, int n_s);
int fem_max_smem_edges(int V);
int fem_threads();
cudaError_t launch_fem_step(const FemArgs& a, int grid, cudaStream_t st);
cudaError_t launch_fem_markers(const FemMarkerArgs& m, int N, cudaStream_t st);

int taxim_smem_bytes();
int taxim_lowres_max_pixels();
int taxim_record_f4(); // float4 stride of the polynomial records the fused kernel expects (8 = one 128-byte line, 5 = packed)
cudaError_t launch_taxim(const TaximArgs& a, const TaximTaps& taps, int N, cudaStream_t s);
cudaError_t launch_flat_rgb(const TaximArgs& a, float* flat_rgb, cudaStream_t s);
cudaError_t launch_indentation_depth(const float* hm, float* out, int N, float gelpad_h, float gelpad_min, cudaStream_t s);
cudaError_t launch_fots(const FotsArgs& a, int N, cudaStream_t s);
cudaError_t launch_taxim_generic(const TaximGenericArgs& a, int N, cudaStream_t s);
int taxim_generic_smem_bytes(int H, int W);
cudaError_t launch_indentation_depth_generic(const float* hm, float* out, int N, int npx, float gelpad_h, float gelpad_min,
                                             cudaStream_t s);
cudaError_t launch_obs_push(const ObsPushArgs& a, int grid, cudaStream_t s);
cudaError_t launch_obs_fill(const ObsFillArgs& a, int grid, cudaStream_t s);

} // namespace tx
