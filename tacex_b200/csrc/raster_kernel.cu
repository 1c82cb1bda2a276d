// raster_kernel.cu -- FEM gel surface -> sensor height map, batched over envs.
//
// The reference obtains the height map of a UIPC scene from the RTX depth camera and leaves "use soft body deformation as
// height map" as a TODO (ref: source/tacex/tacex/gelsight_sensor.py:581-598). Here the deformed top surface of every gel is
// rasterised straight from the FEM state: the tactile image is an orthographic grid over the pad (pixel pitch = pixmm * 640 / W,
// the convention of the optical model), the value of a pixel is the distance from the camera plane to
// the surface along the optical axis, in millimetres, clipped to the far plane -- the format GelSightSensor._get_height_map
// returns and tx_render consumes. One thread block per (env, triangle batch): every surface triangle covers the pixels of its
// bounding box that pass the barycentric inside test; pixels on a shared edge are written by both triangles with an atomic
// minimum on the (positive) float bits, so the result does not depend on the order.
// Compiled without FMA contraction: every double operation is one IEEE operation in the order written (the NumPy checker in
// tests/ reproduces it bit for bit).
#include "tx_kernels.h"

namespace tx {

__global__ void __launch_bounds__(256) heightmap_fill_kernel(float* __restrict__ hm, size_t n, float far_mm)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) hm[i] = far_mm;
}

__global__ void __launch_bounds__(128) heightmap_raster_kernel(const RasterArgs a)
{
    const int env = blockIdx.y;
    const double* X = a.x + (size_t)env * 3 * a.V;
    float* hm = a.hm + (size_t)env * a.H * a.W;
    for (int t = blockIdx.x; t < a.n_tris; t += gridDim.x) {
        const int i0 = a.tris[3 * t], i1 = a.tris[3 * t + 1], i2 = a.tris[3 * t + 2];
        const double ax = X[3 * i0], ay = X[3 * i0 + 1], az = X[3 * i0 + 2];
        const double bx = X[3 * i1], by = X[3 * i1 + 1], bz = X[3 * i1 + 2];
        const double cx = X[3 * i2], cy = X[3 * i2 + 1], cz = X[3 * i2 + 2];
        // pixel (col, row) -> pad frame: x = (col - W / 2 + 0.5) * pitch + ox, y = (row - H / 2 + 0.5) * pitch + oy
        const double x_lo = fmin(ax, fmin(bx, cx)), x_hi = fmax(ax, fmax(bx, cx));
        const double y_lo = fmin(ay, fmin(by, cy)), y_hi = fmax(ay, fmax(by, cy));
        int c0 = (int)floor((x_lo - a.ox) / a.pitch + a.W / 2.0 - 0.5), c1 = (int)ceil((x_hi - a.ox) / a.pitch + a.W / 2.0 - 0.5);
        int r0 = (int)floor((y_lo - a.oy) / a.pitch + a.H / 2.0 - 0.5), r1 = (int)ceil((y_hi - a.oy) / a.pitch + a.H / 2.0 - 0.5);
        c0 = max(c0, 0); r0 = max(r0, 0); c1 = min(c1, a.W - 1); r1 = min(r1, a.H - 1);
        if (c1 < c0 || r1 < r0) continue;
        const double v0x = bx - ax, v0y = by - ay, v1x = cx - ax, v1y = cy - ay;
        const double d00 = v0x * v0x + v0y * v0y, d01 = v0x * v1x + v0y * v1y, d11 = v1x * v1x + v1y * v1y;
        const double den = d00 * d11 - d01 * d01;
        if (!(fabs(den) > 0.0)) continue; // degenerate in projection
        const int bw = c1 - c0 + 1, npx = bw * (r1 - r0 + 1);
        for (int k = threadIdx.x; k < npx; k += blockDim.x) {
            const int row = r0 + k / bw, col = c0 + k % bw;
            const double px = ((double)col - a.W / 2.0 + 0.5) * a.pitch + a.ox, py = ((double)row - a.H / 2.0 + 0.5) * a.pitch + a.oy;
            const double v2x = px - ax, v2y = py - ay;
            const double d20 = v2x * v0x + v2y * v0y, d21 = v2x * v1x + v2y * v1y;
            const double b1 = (d11 * d20 - d01 * d21) / den, b2 = (d00 * d21 - d01 * d20) / den;
            const double b0 = 1.0 - b1 - b2;
            if (b0 >= -1e-9 && b1 >= -1e-9 && b2 >= -1e-9) {
                const double z = b0 * az + b1 * bz + b2 * cz;
                float d = (float)((z - a.cam_z) * 1000.0); // the camera sits below the pad and looks up
                d = fminf(fmaxf(d, 0.0f), a.far_mm);
                atomicMin(reinterpret_cast<int*>(hm + (size_t)row * a.W + col), __float_as_int(d)); // d >= 0: int order == float order
            }
        }
    }
}

// ---- Isaac x UIPC attachment: per-step aim positions of the attached gel vertices -----------------------------------------------
// ref: source/tacex_uipc/tacex_uipc/sim/uipc_attachments.py:388-428 (_compute_aim_positions): aim = transform_points(offsets, pos, quat)
// = R(quat) offset + pos in float32 (Isaac Lab's matrix_from_quat: two_s = 2 / |q|^2, the pytorch3d formula), handed to the solver
// as float64. The reference does this for ONE body with a torch -> NumPy -> python-callback round trip per step; here one launch
// serves every env. pose [N][7] = position xyz + quaternion wxyz (what the reference slices from the PhysX view).
__global__ void __launch_bounds__(128) attachment_aim_kernel(const float* __restrict__ pose, const float* __restrict__ offsets, int A,
                                                             int per_env_offsets, double* __restrict__ aim)
{
    const int env = blockIdx.x;
    const float* p = pose + (size_t)env * 7;
    const float r = p[3], i = p[4], j = p[5], k = p[6];
    const float two_s = __fdiv_rn(2.0f, __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r, r), __fmul_rn(i, i)), __fmul_rn(j, j)), __fmul_rn(k, k)));
    float R[9];
    R[0] = __fadd_rn(1.0f, -__fmul_rn(two_s, __fadd_rn(__fmul_rn(j, j), __fmul_rn(k, k))));
    R[1] = __fmul_rn(two_s, __fadd_rn(__fmul_rn(i, j), -__fmul_rn(k, r)));
    R[2] = __fmul_rn(two_s, __fadd_rn(__fmul_rn(i, k), __fmul_rn(j, r)));
    R[3] = __fmul_rn(two_s, __fadd_rn(__fmul_rn(i, j), __fmul_rn(k, r)));
    R[4] = __fadd_rn(1.0f, -__fmul_rn(two_s, __fadd_rn(__fmul_rn(i, i), __fmul_rn(k, k))));
    R[5] = __fmul_rn(two_s, __fadd_rn(__fmul_rn(j, k), -__fmul_rn(i, r)));
    R[6] = __fmul_rn(two_s, __fadd_rn(__fmul_rn(i, k), -__fmul_rn(j, r)));
    R[7] = __fmul_rn(two_s, __fadd_rn(__fmul_rn(j, k), __fmul_rn(i, r)));
    R[8] = __fadd_rn(1.0f, -__fmul_rn(two_s, __fadd_rn(__fmul_rn(i, i), __fmul_rn(j, j))));
    const float* off = offsets + (per_env_offsets ? (size_t)env * A * 3 : 0);
    for (int a = threadIdx.x; a < A; a += blockDim.x) {
        const float ox = off[3 * a], oy = off[3 * a + 1], oz = off[3 * a + 2];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[3 * c], ox), __fmul_rn(R[3 * c + 1], oy)), __fmul_rn(R[3 * c + 2], oz)), p[c]);
            aim[((size_t)env * A + a) * 3 + c] = (double)v;
        }
    }
}

cudaError_t launch_attachment_aim(const float* pose, const float* offsets, int N, int A, int per_env_offsets, double* aim, cudaStream_t s)
{
    attachment_aim_kernel<<<N, 128, 0, s>>>(pose, offsets, A, per_env_offsets, aim);
    return cudaGetLastError();
}

cudaError_t launch_heightmap(const RasterArgs& a, int N, cudaStream_t s)
{
    const size_t n = (size_t)N * a.H * a.W;
    heightmap_fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a.hm, n, a.far_mm);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int gx = a.n_tris < 60 ? a.n_tris : 60;
    heightmap_raster_kernel<<<dim3(gx, N), 128, 0, s>>>(a);
    return cudaGetLastError();
}

} // namespace tx
