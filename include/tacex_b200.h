/*
 * tacex_b200.h -- C ABI of the B200-native tactile-image synthesis engine (libtacex_b200.so).
 *
 * The reference (DH-Ng/TacEx) has no FFI on this path: its plug-in boundary is the Python ABC
 * GelSightSimulator (ref: source/tacex/tacex/simulation_approaches/gelsight_simulator.py:17-56) whose
 * implementations run torch / NumPy code in-process. This library is what a drop-in implementation of
 * that ABC binds (via ctypes): every entry point below replaces one reference
 * function of the hot path, cited at its declaration.
 *
 * Conventions
 *   - plain C types only; all data pointers are DEVICE pointers owned by the caller (e.g. torch tensors),
 *     except where a parameter is documented as a host pointer;
 *   - every call is asynchronous and ordered on the CUDA stream given at tx_create; no call synchronises
 *     the host except tx_create / tx_upload_tables / tx_destroy / the *_host convenience entry points;
 *   - return value: TX_OK (0) or a negative tx_status; tx_last_error() returns a message for the last
 *     failure on that handle; no C++ exception crosses the boundary;
 *   - a handle is bound to one (device, stream) and is not thread-safe; different handles are independent.
 *   - images are row-major [N][H][W]; RGB output is [N][H][W][3] float32 in [0, 1] (NHWC, as
 *     TaximSimulator.optical_simulation returns it, ref: .../gpu_taxim/taxim_sim.py:80-113).
 */
#ifndef TACEX_B200_H
#define TACEX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TX_ABI_VERSION 2
#define TX_MAX_BLURS 8
#define TX_MAX_TAPS 64
#define TX_MAX_MARKERS 256

typedef enum {
    TX_OK = 0,
    TX_ERR_INVALID_ARG = -1,
    TX_ERR_UNSUPPORTED = -2, /* shape / kernel radii this build has no specialisation for */
    TX_ERR_CUDA = -3,
    TX_ERR_NO_TABLES = -4,
    TX_ERR_NO_DEVICE = -5,
    TX_ERR_STATE = -6
} tx_status;

typedef struct tx_handle tx_handle;

/* Optical + marker configuration. Mirrors TaximSimulatorCfg / params.json / FOTSMarkerSimulatorCfg
 * (ref: .../gpu_taxim/taxim_sim_cfg.py:12-36, .../gpu_taxim/sim/taxim_impl.py:17-62,
 *       .../fots/fots_marker_sim_cfg.py:15-75, .../fots/fots_marker_sim.py:77). */
typedef struct {
    int abi_version;           /* TX_ABI_VERSION */
    int H, W;                  /* tactile image shape: (240, 320) runs the specialised 2-CTA kernel (with kernel sizes
                                  61,33,17,9,5,3,5); any other shape with 3 <= H, W and H * W <= 19200 -- the reference's RL
                                  tasks render 32 x 24 / 32 x 32 -- runs the arbitrary-resolution kernel: same results as the
                                  reference at that shape, no marker grid, no fused camera resize */
    int max_envs;              /* upper bound of N for scratch allocation */
    int num_bins;              /* 125 */
    float pixmm;               /* 0.0295 (NOT rescaled with the resolution) */
    float calib_h, calib_w;    /* 480, 640 */
    float contact_scale;       /* 0.4 */
    float gelpad_height_m;     /* 0.0045 */
    float gelpad_to_cam_min_m; /* 0.024 */
    int n_blurs;               /* pyramid levels + final blur (7) */
    int ksx[TX_MAX_BLURS];     /* odd tap counts along x (W) per blur */
    int ksy[TX_MAX_BLURS];     /* odd tap counts along y (H) per blur */
    float taps_x[TX_MAX_BLURS][TX_MAX_TAPS]; /* normalised float32 Gaussian taps (host computes them exactly */
    float taps_y[TX_MAX_BLURS][TX_MAX_TAPS]; /*  as taxim_torch.py:363-367 does)                              */
    /* FOTS marker model */
    int marker_rows, marker_cols; /* 9 x 11 (reference default) or 7 x 9 (63 markers) */
    float marker_x0, marker_y0;   /* 15, 26 */
    double fots_lambda[3];        /* dilate, shear, twist: 0.00125, 0.00021, 0.00038 */
    double mm2pix;                /* 19.58 */
    double shear_max_px;          /* 10 */
    double theta_max_rad;         /* 60 deg */
} tx_config;

typedef struct {
    uint64_t render_calls, frames_rendered, fots_calls, depth_calls;
    uint64_t kernels_launched; /* total kernels this handle has launched */
} tx_counters;

/* ---- lifetime ------------------------------------------------------------------------------------------- */

int tx_abi_version(void);

/* Creates a handle on CUDA device `device` bound to `cuda_stream` (a cudaStream_t cast to void*, NULL = the
 * legacy default stream). Fails with TX_ERR_NO_DEVICE when no CUDA device is usable -- there is no CPU path. */
int tx_create(const tx_config* cfg, int device, void* cuda_stream, tx_handle** out);
void tx_destroy(tx_handle* h);
const char* tx_last_error(const tx_handle* h); /* h may be NULL: returns the last creation error */
int tx_get_counters(const tx_handle* h, tx_counters* out);

/* Uploads the calibration tables (HOST pointers, copied and re-laid-out on the device):
 *   poly_grad  [3][nb][nb][6] float32, RGB channel order  (ref: taxim_torch.py:73-80 `__poly_grad`)
 *   background [3][H][W]      float32                     (ref: taxim_torch.py:136-137 background at (H, W))
 *   gel_map    [H][W] float32 or NULL for a flat gel map (== 0 after the shift; ref: taxim_torch.py:82-90,159-164) */
int tx_upload_tables(tx_handle* h, const float* poly_grad, const float* background, const float* gel_map);

/* ---- hot path --------------------------------------------------------------------------------------------- */

/* Replaces TaximSimulator.compute_indentation_depth (ref: .../gpu_taxim/taxim_sim.py:115-131).
 *   height_mm [N][H][W] -> depth_mm [N] */
int tx_indentation_depth(tx_handle* h, const float* height_mm, int N, float* depth_mm);
/* The same for height maps at the CAMERA resolution (any number of pixels per frame): the reference computes the indentation depth
 * from the camera-resolution map even when the optical model runs on a resized one (taxim_sim.py:115-131 reads
 * sensor._data.output["height_map"]). IEEE division by 1000, like the reference on the CPU. */
int tx_indentation_depth_frames(tx_handle* h, const float* frames_mm, int N, int pixels_per_frame, float* depth_mm);

/* Replaces TaximTorch.render_direct(with_shadow=False, press_depth=...) + the NHWC copy
 * (ref: .../gpu_taxim/sim/taxim_torch.py:174-195, 225-258, 432-503; .../gpu_taxim/taxim_sim.py:104-111).
 *   height_mm [N][H][W]
 *   press_mm  [N] or NULL: NULL = compute the indentation depth inside the same kernel (fused
 *             compute_indentation_depth + optical_simulation) and, if depth_out != NULL, store it there
 *   rgb       [N][H][W][3]
 *   deformed  optional [N][H][W]  deformed gel height (mm), the reference's `__compute_gel_pad_deformation` output
 *   mask      optional [N][H][W]  uint8 shrunken contact mask
 * When the handle has a marker grid, the per-env inputs of the FOTS model (mask centroid sums, max of the
 * deformed gel, gel height and mask at the marker positions) are recorded for a following tx_fots_markers. */
int tx_render(tx_handle* h, const float* height_mm, const float* press_mm, int N, float* rgb, float* depth_out,
              float* deformed, uint8_t* mask);

/* Same as tx_render(press_mm = NULL) but takes the sensor camera's DEPTH image in metres and fuses
 * GelSightSensor._get_height_map (ref: source/tacex/tacex/gelsight_sensor.py:581-593: inf -> far clipping plane,
 * metres -> millimetres) into the load stage. height_mm_out (optional, [N][H][W]) receives the height map the
 * sensor publishes as data.output["height_map"]. */
int tx_render_depth(tx_handle* h, const float* depth_m, float clip_max_m, int N, float* rgb, float* depth_out,
                    float* height_mm_out);

/* Camera resolution != tactile resolution. The reference resizes the sensor camera's height map with torchvision's
 * F.resize (bilinear, antialias) before rendering when the camera is coarser than the tactile image -- its RL tasks and
 * the FEM preset run 32x32 / 32x24 cameras (ref: .../gpu_taxim/taxim_sim.py:88-89, .../fots/fots_marker_sim.py:121-127,
 * tacex_assets/sensors/gelsight_mini/gsmini_taxim_fem_cfg.py:27,52). tx_set_camera_resolution precomputes the filter taps
 * for [Hc][Wc] -> [H][W] (up-sampling only, Hc * Wc <= 9600); tx_render_camera then takes frames [N][Hc][Wc] -- height maps
 * in mm, or (is_depth != 0) depth images in metres with inf -> clip_max_m as tx_render_depth -- resizes them in the load
 * stage of the fused kernel and renders as tx_render does. press_mm == NULL: the indentation depth is computed from the
 * CAMERA-resolution frame, as compute_indentation_depth does (ref: taxim_sim.py:115-131), and stored in depth_out. */
int tx_set_camera_resolution(tx_handle* h, int Hc, int Wc);
int tx_render_camera(tx_handle* h, const float* frames, int is_depth, float clip_max_m, const float* press_mm, int N,
                     float* rgb, float* depth_out, float* deformed, uint8_t* mask);

/* ---- shadow branch (TaximSimulatorCfg.with_shadow = True; ref: .../gpu_taxim/sim/taxim_torch.py:96-125, 260-346) ----------------
 * tx_upload_shadow_tables takes the init-time data the reference prepares in TaximTorch.__init__ (HOST pointers): the padded
 * shadow table [3][D][Hn][S] (RGB order, / 255, +inf padded), cos / sin of the ray-fan angles [D][F] (float32, evaluated by the
 * host exactly like the reference's torch tensors), the scalars of params.json scaled to the image shape, the box kernels of the
 * two mask-dilation rounds and the taps of the shadow blur. tx_render_shadow = tx_render + the shadow post-pass: the same
 * arguments and results as tx_render(..., deformed = NULL, mask = NULL), RGB with shadows. 240 x 320 handles only. */
typedef struct {
    int D, Hn, S, F;                 /* 63 directions, 24 heights, 51 samples per ray, 4 fan rays */
    float depth_0;                   /* 0.4 (taxim_torch.py:97) */
    float height_precision;          /* 0.1 */
    float discretize_precision;      /* 0.1 */
    float step_x, step_y;            /* shadow_step(shape)[1], [0] (the reference swaps the two, taxim_torch.py:300-305) */
    int dil[4];                      /* ky, kx of dilation round 0, then of round 1 (taxim_torch.py:261-270) */
    int ks_sx, ks_sy;                /* tap counts of the shadow blur */
    float taps_sx[TX_MAX_TAPS], taps_sy[TX_MAX_TAPS];
} tx_shadow_config;
int tx_upload_shadow_tables(tx_handle* h, const tx_shadow_config* cfg, const float* table, const float* fan_cos, const float* fan_sin);
int tx_render_shadow(tx_handle* h, const float* height_mm, const float* press_mm, int N, float* rgb, float* depth_out);

/* ---- multi-GPU: the single observation all-gather (the reference makes no collective call) ----------
 * A frame differs from the flat image only inside a rectangle per half frame. tx_set_rect_output makes the following
 * tx_render* calls record it: rect [N][2][4] int32 = (first row, last row -- local to the half --, first column, last column),
 * last < first when empty. tx_obs_push (on `cuda_stream`, e.g. a side stream) stores the pixels of this rank's rectangles and
 * the descriptors into the same block of every peer's gathered buffer: peer_rgb / peer_rect are HOST arrays of n_peers
 * peer-mapped DEVICE pointers (symmetric memory, NVLink peer-to-peer stores); mc_rgb / mc_rect (optional, both or none) are the
 * addresses of the same block in the NVSwitch MULTICAST mapping of the buffers: one multimem.st per 16 bytes leaves the GPU and the
 * switch replicates it to every peer (egress / (n_peers) of the unicast stores). After a cross-rank barrier tx_obs_fill completes
 * this rank's gathered buffer rgb_all [N_total][H][W][3]: everything outside the rectangles of the envs NOT in
 * [skip_lo, skip_hi) (this rank's own, rendered in place) is copied from the flat image. prev_rect [N_total][2][4] (optional,
 * inout, one per gathered buffer, initialised to (0, H/2-1, 0, W-1) per half) remembers what the buffer held after its last
 * fill, so that only the part of the old rectangle the new one does not cover is restored. Result == gathering whole frames. */
int tx_set_rect_output(tx_handle* h, int32_t* rect);
/* Fused observation all-gather: the following tx_render calls (with tx_set_rect_output active, whole batch, not the chunked host
 * path) store the evaluated rectangle of every frame and its descriptor through the NVSwitch MULTICAST mapping of this rank's block of
 * the gathered buffers (multimem.st from the render kernel's epilogue: compute and collective in one kernel, no second pass over
 * the pixels, no SM taken by a push kernel). mc_rgb / mc_rect: multicast addresses of [N][H][W][3] float32 / [N][2][4] int32, or NULL,
 * NULL to switch back. The flat remainder of each remote frame is completed locally by tx_obs_fill after a cross-rank barrier. */
int tx_set_multicast_output(tx_handle* h, float* mc_rgb, int32_t* mc_rect);
int tx_obs_push(tx_handle* h, const float* rgb_local, const int32_t* rect_local, int N, int n_peers, float* const* peer_rgb,
                int32_t* const* peer_rect, float* mc_rgb, int32_t* mc_rect, void* cuda_stream);
int tx_obs_fill(tx_handle* h, float* rgb_all, const int32_t* rect_all, int32_t* prev_rect, int N_total, int skip_lo, int skip_hi,
                void* cuda_stream);

/* Replaces FOTSMarkerSimulator.marker_motion_simulation + MarkerMotion.marker_sim
 * (ref: .../fots/fots_marker_sim.py:114-184, .../fots/sim/marker_motion.py:78-120,144-219) using the gel
 * deformation recorded by the preceding tx_render of the SAME batch (no second blur pyramid).
 *   press_mm [N]  indentation depth (> 0 = in contact), theta [N] relative yaw of the indenter (rad)
 *   traj0    [N][4] in/out  (x0_mm, y0_mm, theta0, valid): first in-contact sample of the contact episode
 *   traj_len [N]    in/out  samples in the current episode
 *   markers  [N][2][M][2]   [:,0] initial, [:,1] current marker (x, y) in pixels */
int tx_fots_markers(tx_handle* h, const float* press_mm, const float* theta, int N, float* traj0, int32_t* traj_len,
                    float* markers);

/* Replaces the F.resize of TaximSimulator.optical_simulation / FOTSMarkerSimulator.marker_motion_simulation when the sensor camera
 * is FINER than the tactile image (ref: .../gpu_taxim/taxim_sim.py:88-89, .../fots/fots_marker_sim.py:121-122): torchvision's
 * antialiased bilinear resize, bit-identical to torch.nn.functional.interpolate(mode="bilinear", antialias=True).
 *   src [N][Hi][Wi] -> dst [N][H][W] (the handle's tactile shape); scales that need more than 8 taps per axis: TX_ERR_UNSUPPORTED.
 * (A camera COARSER than the tactile image is resized inside the render kernel: tx_set_camera_resolution / tx_render_camera.) */
int tx_resize(tx_handle* h, const float* src, int N, int Hi, int Wi, float* dst);

/* Replaces FOTSMarkerSimulator.draw_markers and the per-env marker-overlay loop of the reference's RL task
 * (ref: .../fots/fots_marker_sim.py:346-384; source/tacex_tasks/tacex_tasks/ball_rolling_tactile/ball_rolling_taxim_fots.py:918-937),
 * batched over envs in one launch.
 *   tx_set_marker_patches: HOST pointer, [10][10][12][12] uint8 = patch_array[:, :, w] of the reference's generate_patch_array()
 *                          for the marker size in use (w = floor((marker_size - 1.5) * 10) = 15 for the default size 3)
 *   markers        [N][2][M][2] as tx_fots_markers returns them ([:,1] = current positions are drawn, in order)
 *   rgb_in         [N][H][W][3] float32 or NULL; apply != 0: rgb = ((rgb * 255) * (marker / 255)) / 255 (the reference's order)
 *   rgb_out        optional [N][H][W][3] float32 (may alias rgb_in)
 *   marker_img_out optional [N][H][W] uint8   the marker image draw_markers returns
 *   rgb_u8_out     optional [N][H][W][3] uint8 round(rgb * 255): the observation at a quarter of the bytes (extension) */
int tx_set_marker_patches(tx_handle* h, const uint8_t* patches);
int tx_marker_overlay(tx_handle* h, const float* markers, int N, int M, const float* rgb_in, int apply, float* rgb_out,
                      uint8_t* marker_img_out, uint8_t* rgb_u8_out);

/* Initial marker grid (host pointers, M ints each). ref: marker_motion.py:58-76 */
int tx_marker_grid(const tx_handle* h, int32_t* mx, int32_t* my);

/* Profiling hook: when `ticks` (device, [2*N][40] int64) is non-NULL, every CTA of the following tx_render launches
 * stores clock64() stamps at its phase boundaries there (used by tools/phase_times.py; NULL disables). */
int tx_debug_set_ticks(tx_handle* h, long long* ticks);
/* Test hook: bit 3 (8) forces the exact (slow) division / sqrt path of the epilogue; other bits are profiling ablations. */
int tx_debug_set_flags(tx_handle* h, int flags);

/* ---- host-buffer convenience (the end-to-end path the benchmark times) -------------------------------------- */

/* height_mm_host [N][H][W] pinned or pageable HOST memory -> rgb_host [N][H][W][3], depth_host [N] (optional),
 * markers_host [N][2][M][2] (optional, needs theta_host). Copies H2D, runs the fused path, copies D2H and
 * synchronises the stream before returning. */
int tx_step_host(tx_handle* h, const float* height_mm_host, const float* theta_host, int N, float* rgb_host,
                 float* depth_host, float* markers_host);

/* ==== gel FEM substep (UIPC-style) =================================================================================
 * Replaces UipcSim.step() = world.advance() + world.retrieve() for the gel pads of all sensors
 * (ref: source/tacex_uipc/tacex_uipc/sim/uipc_sim.py:250-252 -> libuipc SimEngine::do_advance,
 *  source/tacex_uipc/libuipc/src/backends/cuda/engine/sim_engine_do_advance.cu:16-387) and the FEM marker read-out
 * (ref: source/tacex/tacex/simulation_approaches/fem_based/mani_skill_sim.py:82-86). All state is float64 like libuipc. */

typedef struct {
    int type;    /* 0 sphere, 1 oriented box, 2 the triangle mesh of tx_fem_set_indenter_mesh (pose c, R; h unused) */
    double c[3]; /* centre, world frame [m] */
    double R[9]; /* rotation, row-major, world = R * local */
    double h[3]; /* box half extents, or h[0] = sphere radius */
} tx_fem_indenter;

typedef struct {
    int converged, newton_iters, pcg_iters, ls_halvings;
    double min_dist, last_res, energy;
} tx_fem_stats;

/* ref for the defaults: source/tacex_uipc/tacex_uipc/sim/uipc_sim.py:32-131 (dt 0.01, Newton max 1024, velocity_tol 0.05,
 * PCG tol 1e-3, line search 8, d_hat, contact resistance 10 GPa), objects/uipc_object.py:59-84 (E, nu, rho) */
typedef struct {
    int V, T, A, S;
    double dt, gravity[3];
    double mu, lambda, density;
    double attach_strength;
    double d_hat, kappa;
    int newton_max_iter;
    double velocity_tol, pcg_tol_rate;
    int pcg_max_iter_ratio, ls_max_iter, substep;
    int rest_volume_det; /* 1: elastic rest "volume" = det(Dm) as libuipc does; 0: det/6 */
    double friction_mu;  /* Coulomb coefficient gel / indenter (lagged IPC friction, ref: contact_system/contact_models/
                            ipc_vertex_half_plane_frictional_contact.cu; uipc_sim.py default friction ratio 0.5); 0 = off */
    double eps_velocity; /* friction eps_velocity [m/s] (0.01) */
} tx_fem_config;

typedef struct tx_fem tx_fem;

/* Mesh arrays are HOST pointers: X [V][3] rest positions, tets [T][4], attach [A], surf [S] vertex ids. */
int tx_fem_create(const tx_fem_config* cfg, const double* X, const int32_t* tets, const int32_t* attach, const int32_t* surf,
                  int device, void* cuda_stream, tx_fem** out);
void tx_fem_destroy(tx_fem* f);
const char* tx_fem_last_error(const tx_fem* f);
/* lumped vertex masses (HOST pointer, V doubles) -- ref: libuipc/src/geometry/compute_vertex_volume.cpp:20-52 */
int tx_fem_get_mass(const tx_fem* f, double* mass);

/* One implicit-Euler IPC step for N gels (DEVICE pointers): x, v, x_prev [N][V][3] in/out; aim [N][A][3] target positions
 * of the attached vertices; ind_prev / ind_next [N] pose of the prescribed indenter at the start / end of the step;
 * stats [N] or NULL. */
int tx_fem_step(tx_fem* f, double* x, double* v, double* x_prev, const double* aim, const tx_fem_indenter* ind_prev,
                const tx_fem_indenter* ind_next, int N, tx_fem_stats* stats);

/* Profiling hook: cycles (device, [#SMs][6] int64) receives the per-CTA clock64 totals of the phases of the last env each
 * CTA processed (assembly, PCG, line search, and within the assembly: tets, vertex rows, edges); NULL disables. */
/* Prescribed rigid TRIANGLE-MESH indenter (indenter type 2): tri_local HOST [n_tris][3][3] float64, triangle vertices in the indenter's
 * own frame [m]; one mesh per handle, posed per gel by tx_fem_indenter.c / .R. Every (gel surface vertex, triangle) candidate inside
 * d_hat gets the reference's point-triangle barrier with closest-feature classification (ref: libuipc
 * utils/distance/distance_flagged.h:248-350, contact_system/contact_models/ipc_simplex_normal_contact.cu:270-342,
 * collision_detection/filters/lbvh_simplex_trajectory_filter.cu:600-690), restricted to the gel vertex (the indenter is prescribed).
 * n_tris = 0 removes the mesh. Distances to the mesh are searched within 2 d_hat: tx_fem_stats.min_dist is exact below that radius
 * and reported as 2 d_hat beyond it. */
int tx_fem_set_indenter_mesh(tx_fem* f, int n_tris, const double* tri_local);
/* Second half of the vertex-face contact + the EDGE-EDGE candidates with a mesh indenter: tris HOST [n_tris][3] vertex ids of the gel
 * surface that the indenter's VERTICES (against the triangles) and EDGES (against the triangles' unique edges, mollified as the
 * reference: utils/distance/edge_edge_mollifier.h, distance_flagged.h:352-487, ccd.inl:267-354) can touch (every triangle edge must be an edge of the tet mesh; n_tris <= 576, one triangle per thread; 0
 * switches it off). Same classification / squared distance / barrier per candidate as above with the three gel vertices as the
 * unknowns: exact energy and gradient (the reference's 12-gradient), Gauss-Newton Hessian (closest point frozen), the reference's
 * additive CCD with the moving triangle (utils/distance/details/ccd.inl:200-262). */
int tx_fem_set_contact_surface(tx_fem* f, int n_tris, const int32_t* tris);
int tx_fem_debug_set_cycles(tx_fem* f, long long* cycles);

/* FEM marker read-out. set: HOST pointers, tri [M][3] surface-triangle vertex ids and barycentric weights [M][3] per marker,
 * camera pose (cam_R row-major world = R * cam, cam_t) and pinhole intrinsics. run: x DEVICE [N][V][3] ->
 * markers DEVICE [N][2][M][2] ([:,0] rest, [:,1] current, (u, v) pixels). */
int tx_fem_set_markers(tx_fem* f, int M, const int32_t* tri, const double* weights, const double* cam_R, const double* cam_t,
                       double fx, double fy, double cx, double cy);
int tx_fem_markers(tx_fem* f, const double* x, int N, float* markers);
/* Tail of gen_marker_flow (ref: tactile_sensor_sapienipc_modified.py:404-408): normalize != 0 maps the pixel coordinates to
 * uv / (img_w / 2) - 1; zero_all != 0: no marker survived the reference's uv mask (:382-387), the flow is all zeros. The mask
 * itself and the padding to num_markers only depend on the REST positions: the host applies them to (tri, weights) once
 * (tacex_b200/fem.py::reference_marker_tail) before tx_fem_set_markers. */
int tx_fem_set_marker_output(tx_fem* f, int normalize, double img_w, int zero_all);

/* FEM gel surface -> sensor height map for every env in one launch (the reference reads the height map of a UIPC scene from the RTX
 * depth camera and leaves the soft-body route as a TODO, ref: source/tacex/tacex/gelsight_sensor.py:581-598).
 *   tx_fem_set_surface: HOST pointer, tris [n_tris][3] vertex ids of the gel's top surface
 *   tx_fem_heightmap:   x DEVICE [N][V][3] -> height_mm DEVICE [N][H][W] float32: distance from the camera plane (z = cam_z_m in
 *                       the pad frame, below the pad: -0.024) up to the deformed surface along the optical axis, in mm, clipped to
 *                       [0, far_mm]; the image
 *                       is an orthographic grid of pitch_m centred at (origin_x, origin_y) of the pad frame (the optical
 *                       model's convention). Pixels no triangle covers hold far_mm. The result feeds tx_render directly. */
int tx_fem_set_surface(tx_fem* f, int n_tris, const int32_t* tris);

/* Isaac x UIPC attachment, per-step part (ref: source/tacex_uipc/tacex_uipc/sim/uipc_attachments.py:388-428 _compute_aim_positions +
 * the animator callback :364-386): aim positions of the attached vertices = R(quat) * offset + pos of the rigid body (sensor case) they
 * hang on, for every env in one launch. pose DEVICE [N][7] float32 (x y z, quaternion w x y z), offsets DEVICE float32 [A][3] (shared)
 * or [N][A][3] (per_env_offsets != 0): the attachment points in the body frame (host, once: tacex_b200/fem.py::
 * compute_attachment_data, ref :247-350), aim DEVICE [N][A][3] float64 = the `aim` argument of tx_fem_step. */
int tx_fem_attachment_aim(tx_fem* f, const float* pose, const float* offsets, int N, int per_env_offsets, double* aim);
int tx_fem_heightmap(tx_fem* f, const double* x, int N, float* height_mm, int H, int W, double pitch_m, double origin_x, double origin_y,
                     double cam_z_m, float far_mm);

#ifdef __cplusplus
}
#endif
#endif /* TACEX_B200_H */
