// fots_kernel.cu -- FOTS marker-motion model, one CTA per env, fed by what the fused Taxim kernel recorded.
//
// Replaces (ref = /root/reference/source/tacex/tacex/simulation_approaches/fots):
//   fots_marker_sim.py:128-182   per-env python loop: contact centroid, trajectory bookkeeping, marker_sim call
//   sim/marker_motion.py:144-219 _motion_callback (contact markers, dilate, shear, twist)
//   sim/marker_motion.py:78-120  _shear / _twist / _dilate
// The reference recomputes the whole blur pyramid here and copies two planes per env to the host; this kernel
// reads 2 + 2*M small values per env instead. Arithmetic follows the NumPy reference: float32 for the trajectory
// scalars (NumPy 2 weak python scalars), float64 for the marker sums, including the `cos(theta - 1)` quirk
// (marker_motion.py:98-99).
#include "tx_common.cuh"
#include "tx_kernels.h"

namespace tx {

__global__ void __launch_bounds__(128) fots_kernel(const FotsArgs a)
{
    const int n = blockIdx.x, tid = threadIdx.x, M = a.M;
    __shared__ float s_h[TX_MAX_MARKERS];
    __shared__ unsigned char s_c[TX_MAX_MARKERS];
    __shared__ int s_x[TX_MAX_MARKERS], s_y[TX_MAX_MARKERS];
    __shared__ float s_tr[8]; // x0, y0, th0, cx, cy, th, have_traj, ncontact
    float* out0 = a.markers + (size_t)n * 4 * M;
    float* out1 = out0 + 2 * M;

    for (int k = tid; k < M; k += blockDim.x) {
        s_x[k] = a.mk_x[k];
        s_y[k] = a.mk_y[k];
        out0[2 * k] = (float)s_x[k];
        out0[2 * k + 1] = (float)s_y[k];
    }
    const float press = a.press[n];
    if (!(press > 0.0f)) { // contact episode over: clear the trajectory, markers at rest (fots_marker_sim.py:177-179)
        for (int k = tid; k < M; k += blockDim.x) {
            out1[2 * k] = (float)a.mk_x[k];
            out1[2 * k + 1] = (float)a.mk_y[k];
        }
        if (tid == 0) {
            a.traj_len[n] = 0;
            a.traj0[4 * n + 3] = 0.0f;
        }
        return;
    }
    const float bmax = fmaxf(a.aux_bmax[2 * n], a.aux_bmax[2 * n + 1]);
    for (int k = tid; k < M; k += blockDim.x) {
        const bool inb = s_x[k] >= 0 && s_x[k] < IMG_W && s_y[k] >= 0 && s_y[k] < IMG_H;
        const bool c = inb && a.aux_m[(size_t)n * M + k] == 1;
        s_c[k] = c ? 1 : 0;
        // depth = (max(b) - b) / 10 in float32 (marker_motion.py:146-149; fots_marker_sim.py:130)
        s_h[k] = c ? __fdiv_rn(__fadd_rn(bmax, -a.aux_b[(size_t)n * M + k]), 10.0f) : 0.0f;
    }
    if (tid == 0) {
        const unsigned* s0 = a.aux_sums + (size_t)n * 8;
        const double cnt = (double)s0[0] + (double)s0[4];
        const float mrow = (float)(((double)s0[1] + (double)s0[5]) / cnt);
        const float mcol = (float)(((double)s0[2] + (double)s0[6]) / cnt);
        const float cy = __fdiv_rn(__fadd_rn(mrow, -(float)(IMG_H / 2.0)), (float)a.mm2pix);
        const float cx = __fdiv_rn(__fadd_rn(mcol, -(float)(IMG_W / 2.0)), (float)a.mm2pix);
        const float th = a.theta[n];
        int len = a.traj_len[n];
        if (len == 0) {
            a.traj0[4 * n + 0] = cx;
            a.traj0[4 * n + 1] = cy;
            a.traj0[4 * n + 2] = th;
            a.traj0[4 * n + 3] = 1.0f;
        }
        len += 1;
        a.traj_len[n] = len;
        s_tr[0] = a.traj0[4 * n + 0];
        s_tr[1] = a.traj0[4 * n + 1];
        s_tr[2] = a.traj0[4 * n + 2];
        s_tr[3] = cx;
        s_tr[4] = cy;
        s_tr[5] = th;
        s_tr[6] = len >= 2 ? 1.0f : 0.0f;
    }
    __syncthreads();
    int ncontact = 0;
    for (int k = 0; k < M; ++k) ncontact += s_c[k];
    const bool have_traj = s_tr[6] != 0.0f;
    double scx = 0, scy = 0, shx = 0, shy = 0, tcx = 0, tcy = 0, cm1 = 0, sn = 0;
    if (have_traj) {
        const float x0 = s_tr[0], y0 = s_tr[1], t0 = s_tr[2], cx = s_tr[3], cy = s_tr[4], th = s_tr[5];
        const float mp = (float)a.mm2pix, hw = (float)(IMG_W / 2.0), hh = (float)(IMG_H / 2.0);
        scx = (double)(int)__fadd_rn(__fmul_rn(x0, mp), hw); // python int(): truncation toward zero
        scy = (double)(int)__fadd_rn(__fmul_rn(y0, mp), hh);
        shx = (double)(int)__fmul_rn(__fadd_rn(cx, -x0), mp);
        shy = (double)(int)__fmul_rn(__fadd_rn(cy, -y0), mp);
        shx = fmin(fmax(shx, -a.shear_max), a.shear_max);
        shy = fmin(fmax(shy, -a.shear_max), a.shear_max);
        tcx = (double)(int)__fadd_rn(__fmul_rn(cx, mp), hw);
        tcy = (double)(int)__fadd_rn(__fmul_rn(cy, mp), hh);
        float thf = __fadd_rn(th, -t0);
        const float tmax = (float)a.theta_max;
        thf = fminf(fmaxf(thf, -tmax), tmax);
        cm1 = (double)cosf(__fadd_rn(thf, -1.0f)); // sic: cos(theta - 1)
        sn = (double)sinf(thf);
    }
    for (int k = tid; k < M; k += blockDim.x) {
        const double px = (double)s_x[k], py = (double)s_y[k];
        double nx = px, ny = py;
        if (ncontact > 0) {
            double dx = 0.0, dy = 0.0;
            // contact list order of the reference: columns outer, rows inner (marker_motion.py:152-166)
            for (int q = 0; q < a.cols; ++q) {
                for (int r = 0; r < a.rows; ++r) {
                    const int c = r * a.cols + q;
                    if (s_c[c]) {
                        const double ox = (double)(s_x[k] - s_x[c]), oy = (double)(s_y[k] - s_y[c]);
                        const double g = exp(-a.lamb0 * (ox * ox + oy * oy));
                        dx += (double)s_h[c] * ox * g;
                        dy += (double)s_h[c] * oy * g;
                    }
                }
            }
            nx = px + dx;
            ny = py + dy;
            if (have_traj) {
                double ox = px - scx, oy = py - scy;
                double g = exp(-a.lamb1 * (ox * ox + oy * oy));
                nx += shx * g;
                ny += shy * g;
                ox = px - tcx;
                oy = py - tcy;
                g = exp(-a.lamb2 * (ox * ox + oy * oy));
                nx += (ox * cm1 - oy * sn) * g;
                ny += (ox * sn + oy * cm1) * g;
            }
        }
        out1[2 * k] = (float)nx;
        out1[2 * k + 1] = (float)ny;
    }
}

cudaError_t launch_fots(const FotsArgs& a, int N, cudaStream_t s)
{
    fots_kernel<<<N, 128, 0, s>>>(a);
    return cudaGetLastError();
}

} // namespace tx
