// fem_kernel.cu -- batched gel FEM substep (UIPC-style implicit Euler + IPC barrier), one persistent CTA per gel.
//
// Replaces, for the gel pad of every sensor in the batch (ref = /root/reference/source/tacex_uipc/libuipc/src/backends/cuda):
//   engine/sim_engine_do_advance.cu:200-360            Newton loop, convergence test, line search, velocity update
//   finite_element/bdf/*.cu                            BDF1 predict / kinetic energy, gradient, Hessian / update
//   finite_element/constitutions/stable_neo_hookean_3d.cu:68-165 (+ sym/stable_neo_hookean_3d.inl)   SNH energy / dE/dF / d2E/dF2
//   utils/make_spd.h:7-19                              SPD projection of the 9x9 / 3x3 Hessians
//   finite_element/constraints/soft_position_constraint.cu:99-179   attachment of the gel to the sensor case
//   contact_system/contact_models/ipc_vertex_half_plane_normal_contact.cu + sym/codim_ipc_contact.inl   barrier
//   linear_system/linear_pcg.cu:45-140, finite_element/fem_diag_preconditioner.cu:112-164              PCG + block Jacobi
//   newton_tolerance/max_translation_checker.cu:25-50  tolerance
// The reference runs ONE scene with dozens of kernel launches and host synchronisations per Newton iteration
// (cuBLAS dots, CUB reductions, radix-sorted triplet assembly); here the whole step of one gel runs inside one CTA
// without leaving the SM: vectors live in shared memory (float64), the operator is applied matrix-free from the per-tet
// projected 9x9 Hessians (symmetric packed, L2-resident scratch of the CTA), reductions are block reductions, and the
// element -> vertex assembly is a gather in fixed CSR order (no atomics), so results are bitwise reproducible.
// The indenter is a prescribed analytic rigid body (sphere / oriented box): no broad phase is needed.
#include "tx_kernels.h"
#include <cuda_runtime.h>
#include <math.h>

namespace tx {

constexpr int FEM_THREADS = 576; // one thread per vertex row (V <= 576), 18 warps
constexpr int FEM_VAL_STRIDE = 2836; // shared-memory blocks: val[k][e], k < 9, e < n_s <= FEM_VAL_STRIDE (compile-time stride: immediate offsets)

// ---- small dense helpers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double det3cm(const double* F)
{
    return F[0] * (F[4] * F[8] - F[7] * F[5]) - F[3] * (F[1] * F[8] - F[7] * F[2]) + F[6] * (F[1] * F[5] - F[4] * F[2]);
}

template <int n>
__device__ void jacobi_evd(double* A, double* w, double* Vv)
{
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) Vv[i * n + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                if (i != j) off += A[i * n + j] * A[i * n + j];
                else diag += A[i * n + j] * A[i * n + j];
            }
        if (off <= 1e-30 * (diag + 1e-300)) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = A[p * n + q];
                if (apq == 0.0) continue;
                const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) {
                    const double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - s * akq;
                    A[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {
                    const double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - s * aqk;
                    A[q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double vkp = Vv[k * n + p], vkq = Vv[k * n + q];
                    Vv[k * n + p] = c * vkp - s * vkq;
                    Vv[k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < n; ++i) w[i] = A[i * n + i];
}

// LDL^T with all pivots > 0 proves that H is positive definite (then make_spd is the identity)
template <int n>
__device__ bool is_pd(const double* H)
{
    double L[n * n], D[n];
    for (int j = 0; j < n; ++j) {
        double d = H[j * n + j];
        for (int k = 0; k < j; ++k) d -= L[j * n + k] * L[j * n + k] * D[k];
        if (!(d > 0.0)) return false;
        D[j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = H[i * n + j];
            for (int k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k] * D[k];
            L[i * n + j] = s / d;
        }
    }
    return true;
}

// make_spd with the LDL^T fast path (an all-positive-pivot factorisation proves H is already positive definite)
template <int n>
__device__ void spd_project(double* H)
{
    double L[n * n], D[n];
    bool pd = true;
    for (int j = 0; j < n && pd; ++j) {
        double d = H[j * n + j];
        for (int k = 0; k < j; ++k) d -= L[j * n + k] * L[j * n + k] * D[k];
        if (!(d > 0.0)) { pd = false; break; }
        D[j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = H[i * n + j];
            for (int k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k] * D[k];
            L[i * n + j] = s / d;
        }
    }
    if (pd) return;
    double A[n * n], w[n], Vv[n * n];
    for (int i = 0; i < n * n; ++i) A[i] = H[i];
    jacobi_evd<n>(A, w, Vv);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0.0;
            for (int k = 0; k < n; ++k) s += Vv[i * n + k] * (w[k] < 0.0 ? 0.0 : w[k]) * Vv[j * n + k];
            H[i * n + j] = s;
        }
}

// Stable Neo-Hookean (sym/stable_neo_hookean_3d.inl): energy, dPsi/dF (9), d2Psi/dF2 (81, row-major)
__device__ void snh(const double* F, double mu, double lambda, double* E, double* g, double* H)
{
    const double J = det3cm(F);
    if (E) {
        double IC = 0.0;
        for (int i = 0; i < 9; ++i) IC += F[i] * F[i];
        *E = 0.5 * lambda * (J - 1.0) * (J - 1.0) - mu * (J - 1.0) + 0.5 * mu * (IC - 3.0) + mu * mu / (lambda * lambda);
    }
    if (!g && !H) return;
    double gJ[9];
    const double *f0 = F, *f1 = F + 3, *f2 = F + 6;
    gJ[0] = f1[1] * f2[2] - f1[2] * f2[1]; gJ[1] = f1[2] * f2[0] - f1[0] * f2[2]; gJ[2] = f1[0] * f2[1] - f1[1] * f2[0];
    gJ[3] = f2[1] * f0[2] - f2[2] * f0[1]; gJ[4] = f2[2] * f0[0] - f2[0] * f0[2]; gJ[5] = f2[0] * f0[1] - f2[1] * f0[0];
    gJ[6] = f0[1] * f1[2] - f0[2] * f1[1]; gJ[7] = f0[2] * f1[0] - f0[0] * f1[2]; gJ[8] = f0[0] * f1[1] - f0[1] * f1[0];
    const double c = lambda * (J - 1.0) - mu;
    if (g)
        for (int i = 0; i < 9; ++i) g[i] = mu * F[i] + c * gJ[i];
    if (H) {
        for (int i = 0; i < 9; ++i)
            for (int j = 0; j < 9; ++j) H[i * 9 + j] = lambda * gJ[i] * gJ[j] + (i == j ? mu : 0.0);
        for (int a = 0; a < 3; ++a) {
            const int b1 = (a + 1) % 3, b2 = (a + 2) % 3;
            const double* u = F + 3 * b1;
            const double* v = F + 3 * b2;
            const double Ux[9] = {0, -u[2], u[1], u[2], 0, -u[0], -u[1], u[0], 0};
            const double Vx[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0};
            for (int r = 0; r < 3; ++r)
                for (int s = 0; s < 3; ++s) {
                    H[(3 * a + r) * 9 + (3 * b2 + s)] += c * Ux[r * 3 + s];
                    H[(3 * a + r) * 9 + (3 * b1 + s)] -= c * Vx[r * 3 + s];
                }
        }
    }
}

// ---- analytic SPD projection of the stable Neo-Hookean Hessian --------------------------------------------------------
// make_spd (utils/make_spd.h:7-19) clamps the negative eigenvalues of d2Psi/dF2. For Psi = mu/2 (I_C - 3) - mu (J - 1) +
// lambda/2 (J - 1)^2 the Hessian is H = mu I + c H_J + lambda g_J g_J^T with c = lambda (J - 1) - mu. In the frame of the
// rotation-variant SVD F = U S V^T (det U = det V = 1) it is block diagonal: a 3x3 "scaling" block on (d11, d22, d33),
//   A = mu I + c [[0, s3, s2], [s3, 0, s1], [s2, s1, 0]] + lambda w w^T,  w = (s2 s3, s1 s3, s1 s2),
// and, for every index pair (i, j) with third index k, a 2x2 block [[mu, -c s_k], [-c s_k, mu]] on (d_ij, d_ji) whose
// eigenpairs are the twist (1, -1)/sqrt2 : mu + c s_k and the flip (1, 1)/sqrt2 : mu - c s_k. Clamping those nine
// eigenvalues and rotating back gives the same matrix as the numeric 9x9 eigen-decomposition the reference (and the CPU
// checker) use, at a fraction of the cost and without spilling an 81-entry work matrix per thread.
// Reciprocal / reciprocal square root from the hardware approximation (about 20 bits) + two Newton steps: full double
// precision for normal arguments, a fraction of the cost of the IEEE division / sqrt subroutines (the solver is checked
// against the CPU restatement by tolerance, not bit for bit).
__device__ __forceinline__ double fast_rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}
__device__ __forceinline__ double fast_rsqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double h = 0.5 * x;
    y = y * fma(-h * y, y, 1.5);
    y = y * fma(-h * y, y, 1.5);
    return y * fma(-h * y, y, 1.5);
}

// one Jacobi rotation annihilating a_pq: t = sgn(theta) / (|theta| + sqrt(theta^2 + 1)), theta = (a_qq - a_pp) / (2 a_pq),
// evaluated as t = +-|b| / (|d| + sqrt(d^2 + b^2)) with d = a_qq - a_pp, b = 2 a_pq (no division by a_pq)
__device__ __forceinline__ void jrot(double& app, double& aqq, double& apq, double& arp, double& arq, double& vp0,
                                     double& vq0, double& vp1, double& vq1, double& vp2, double& vq2)
{
    const double d = aqq - app, b = 2.0 * apq;
    const double h2 = fma(d, d, b * b);
    if (!(h2 > 1e-280)) return; // a_pq == 0 (or negligible at any scale)
    const double h = h2 * fast_rsqrt(h2);
    const bool pos = (d == 0.0) || ((d > 0.0) == (b > 0.0));
    const double tm = fabs(b) * fast_rcp(fabs(d) + h);
    const double t = pos ? tm : -tm;
    const double c = fast_rsqrt(fma(t, t, 1.0)), s = t * c;
    const double app_n = app - t * apq, aqq_n = aqq + t * apq;
    const double arp_n = c * arp - s * arq, arq_n = s * arp + c * arq;
    app = app_n; aqq = aqq_n; apq = 0.0; arp = arp_n; arq = arq_n;
    double a, bq;
    a = vp0; bq = vq0; vp0 = c * a - s * bq; vq0 = s * a + c * bq;
    a = vp1; bq = vq1; vp1 = c * a - s * bq; vq1 = s * a + c * bq;
    a = vp2; bq = vq2; vp2 = c * a - s * bq; vq2 = s * a + c * bq;
}

// eigen-decomposition of a symmetric 3x3 (a00 a01 a02 a11 a12 a22); eigenvectors are the COLUMNS of V (row-major v[r][c])
__device__ __forceinline__ void jacobi3(double a00, double a01, double a02, double a11, double a12, double a22, double w[3],
                                        double v[3][3])
{
    double v00 = 1, v01 = 0, v02 = 0, v10 = 0, v11 = 1, v12 = 0, v20 = 0, v21 = 0, v22 = 1;
#pragma unroll 1
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = a01 * a01 + a02 * a02 + a12 * a12;
        if (off <= 1e-32 * (a00 * a00 + a11 * a11 + a22 * a22 + 1e-300)) break;
        jrot(a00, a11, a01, a02, a12, v00, v01, v10, v11, v20, v21); // (p, q) = (0, 1), r = 2
        jrot(a00, a22, a02, a01, a12, v00, v02, v10, v12, v20, v22); // (0, 2), r = 1
        jrot(a11, a22, a12, a01, a02, v01, v02, v11, v12, v21, v22); // (1, 2), r = 0
    }
    w[0] = a00; w[1] = a11; w[2] = a22;
    v[0][0] = v00; v[0][1] = v01; v[0][2] = v02; v[1][0] = v10; v[1][1] = v11; v[1][2] = v12; v[2][0] = v20; v[2][1] = v21; v[2][2] = v22;
}

// Per-tet gradient and 12x12 Hessian blocks straight from the analytic eigen-system (no 9x9 / 12x12 work matrices):
// with dF = sum_v dx_v W_v^T, y_v = U^T dx_v and z_v = V^T W_v the rotated increment is d-hat = sum_v y_v z_v^T, so the
// (va, vb) block of the projected Hessian is U S U^T with
//   S[i][j] = Ap[i][j] z_va[i] z_vb[j]                                        (scaling block, clamped)
//   S[i][i] += dd_ij z_va[j] z_vb[j],  S[j][j] += dd_ij z_va[i] z_vb[i]       (twist / flip pairs i < j, clamped)
//   S[i][j] += od_ij z_va[j] z_vb[i],  S[j][i] += od_ij z_va[i] z_vb[j]
// Output (structure of arrays over the tet index, stride T): rows 0..11 gradient, 12 + 9 v diagonal block of local vertex
// v, 48 + 9 s off-diagonal block of the local pair s in (0,1) (0,2) (0,3) (1,2) (1,3) (2,3).
// Scratch layout (one chunk of FEM_THREADS tets, tet tl of the chunk): 32-byte units u at tsc[(u * FEM_THREADS + tl) * 4],
// so that the tets of a warp write whole sectors side by side and a row / an edge reads whole sectors of one tet:
//   3 v + 0, 1, 2      local vertex v: (g0 g1 g2 -) (d00 d01 d02 d11) (d12 d22 - -)
//   12 + 3 s + 0, 1, 2 off-diagonal block of the local pair s: (b0 b1 b2 b3) (b4 b5 b6 b7) (b8 - - -)
constexpr int FEM_UNITS = 30;
__device__ __forceinline__ void st_unit(double* tsc_tl, int u, double a, double b, double c, double d)
{
    double2* p = reinterpret_cast<double2*>(tsc_tl + (size_t)u * FEM_THREADS * 4);
    p[0] = make_double2(a, b);
    p[1] = make_double2(c, d);
}
__device__ __forceinline__ void ld_unit(const double* tsc_tl, int u, double o[4])
{
    const double2* p = reinterpret_cast<const double2*>(tsc_tl + (size_t)u * FEM_THREADS * 4);
    const double2 a = p[0], b = p[1];
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}

__device__ __noinline__ void tet_contrib(const double* F, const double W[4][3], double mu, double lambda, double sc, double* to)
{
    double U[3][3], sg[3];
    double Z[4][3]; // z_v = V^T W_v (indexed dynamically by the pair loop below: lives in local memory on purpose)
    {
        double V[3][3];
        {
            const double *f0 = F, *f1 = F + 3, *f2 = F + 6;
            double w[3];
            jacobi3(f0[0] * f0[0] + f0[1] * f0[1] + f0[2] * f0[2], f0[0] * f1[0] + f0[1] * f1[1] + f0[2] * f1[2],
                    f0[0] * f2[0] + f0[1] * f2[1] + f0[2] * f2[2], f1[0] * f1[0] + f1[1] * f1[1] + f1[2] * f1[2],
                    f1[0] * f2[0] + f1[1] * f2[1] + f1[2] * f2[2], f2[0] * f2[0] + f2[1] * f2[1] + f2[2] * f2[2], w, V);
#define SWAPC(i, j)                                                                                                    \
    if (w[i] < w[j]) {                                                                                                \
        double t_ = w[i]; w[i] = w[j]; w[j] = t_;                                                                     \
        for (int r_ = 0; r_ < 3; ++r_) { t_ = V[r_][i]; V[r_][i] = V[r_][j]; V[r_][j] = t_; }                         \
    }
            SWAPC(0, 1) SWAPC(0, 2) SWAPC(1, 2)
#undef SWAPC
        }
        const double detV = V[0][0] * (V[1][1] * V[2][2] - V[1][2] * V[2][1]) - V[0][1] * (V[1][0] * V[2][2] - V[1][2] * V[2][0]) +
                            V[0][2] * (V[1][0] * V[2][1] - V[1][1] * V[2][0]);
        if (detV < 0.0)
            for (int r = 0; r < 3; ++r) V[r][2] = -V[r][2];
        {
            double Fv[3][3]; // Fv[i] = F v_i
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int r = 0; r < 3; ++r) Fv[i][r] = F[r] * V[0][i] + F[3 + r] * V[1][i] + F[6 + r] * V[2][i];
            const double n0 = fast_rsqrt(Fv[0][0] * Fv[0][0] + Fv[0][1] * Fv[0][1] + Fv[0][2] * Fv[0][2]);
            for (int r = 0; r < 3; ++r) U[r][0] = Fv[0][r] * n0;
            const double d01 = U[0][0] * Fv[1][0] + U[1][0] * Fv[1][1] + U[2][0] * Fv[1][2];
            const double t1[3] = {Fv[1][0] - d01 * U[0][0], Fv[1][1] - d01 * U[1][0], Fv[1][2] - d01 * U[2][0]};
            const double n1 = fast_rsqrt(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]);
            for (int r = 0; r < 3; ++r) U[r][1] = t1[r] * n1;
            U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
            U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
            U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
#pragma unroll
            for (int i = 0; i < 3; ++i) sg[i] = U[0][i] * Fv[i][0] + U[1][i] * Fv[i][1] + U[2][i] * Fv[i][2];
        }
#pragma unroll
        for (int v = 0; v < 4; ++v)
#pragma unroll
            for (int b = 0; b < 3; ++b) Z[v][b] = V[0][b] * W[v][0] + V[1][b] * W[v][1] + V[2][b] * W[v][2];
    }
    const double J = sg[0] * sg[1] * sg[2];
    const double c = lambda * (J - 1.0) - mu;
    // gradient: dPsi/dF = mu F + c cof(F) = U diag(mu s_i + c s_j s_k) V^T, g_v = dPsi/dF W_v = U (dg o z_v)
    {
        const double dg[3] = {sc * (mu * sg[0] + c * sg[1] * sg[2]), sc * (mu * sg[1] + c * sg[0] * sg[2]), sc * (mu * sg[2] + c * sg[0] * sg[1])};
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const double z0 = dg[0] * Z[v][0], z1 = dg[1] * Z[v][1], z2 = dg[2] * Z[v][2];
            st_unit(to, 3 * v, U[0][0] * z0 + U[0][1] * z1 + U[0][2] * z2, U[1][0] * z0 + U[1][1] * z1 + U[1][2] * z2,
                    U[2][0] * z0 + U[2][1] * z1 + U[2][2] * z2, 0.0);
        }
    }
    // clamped scaling block (symmetric: a00 a01 a02 a11 a12 a22) and twist / flip pairs
    double Ap[6], dd[3], od[3]; // dd / od indexed by the third index k of the pair (i, j)
    {
        const double sw[3] = {sg[1] * sg[2], sg[0] * sg[2], sg[0] * sg[1]};
        double aw[3], Q[3][3];
        jacobi3(mu + lambda * sw[0] * sw[0], c * sg[2] + lambda * sw[0] * sw[1], c * sg[1] + lambda * sw[0] * sw[2],
                mu + lambda * sw[1] * sw[1], c * sg[0] + lambda * sw[1] * sw[2], mu + lambda * sw[2] * sw[2], aw, Q);
#pragma unroll
        for (int k = 0; k < 3; ++k) aw[k] = aw[k] < 0.0 ? 0.0 : aw[k] * sc;
        Ap[0] = Q[0][0] * aw[0] * Q[0][0] + Q[0][1] * aw[1] * Q[0][1] + Q[0][2] * aw[2] * Q[0][2];
        Ap[1] = Q[0][0] * aw[0] * Q[1][0] + Q[0][1] * aw[1] * Q[1][1] + Q[0][2] * aw[2] * Q[1][2];
        Ap[2] = Q[0][0] * aw[0] * Q[2][0] + Q[0][1] * aw[1] * Q[2][1] + Q[0][2] * aw[2] * Q[2][2];
        Ap[3] = Q[1][0] * aw[0] * Q[1][0] + Q[1][1] * aw[1] * Q[1][1] + Q[1][2] * aw[2] * Q[1][2];
        Ap[4] = Q[1][0] * aw[0] * Q[2][0] + Q[1][1] * aw[1] * Q[2][1] + Q[1][2] * aw[2] * Q[2][2];
        Ap[5] = Q[2][0] * aw[0] * Q[2][0] + Q[2][1] * aw[1] * Q[2][1] + Q[2][2] * aw[2] * Q[2][2];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double lt = mu + c * sg[k], lf = mu - c * sg[k];
            const double ltp = lt < 0.0 ? 0.0 : lt, lfp = lf < 0.0 ? 0.0 : lf;
            dd[k] = 0.5 * (ltp + lfp) * sc;
            od[k] = 0.5 * (lfp - ltp) * sc;
        }
    }
    // the ten blocks, one at a time (a rolled loop keeps the live set small: U, Ap, dd, od + one block)
#pragma unroll 1
    for (int pr = 0; pr < 10; ++pr) {
        // pr 0..3: diagonal blocks (v, v); 4..9: pairs (0,1) (0,2) (0,3) (1,2) (1,3) (2,3)
        const int va = pr < 4 ? pr : (pr < 7 ? 0 : (pr < 9 ? 1 : 2));
        const int vb = pr < 4 ? pr : (pr < 7 ? pr - 3 : (pr < 9 ? pr - 5 : 3));
        const double a0 = Z[va][0], a1 = Z[va][1], a2 = Z[va][2];
        const double b0 = Z[vb][0], b1 = Z[vb][1], b2 = Z[vb][2];
        // S (pairs (0,1) k=2, (0,2) k=1, (1,2) k=0)
        const double S00 = Ap[0] * a0 * b0 + dd[2] * a1 * b1 + dd[1] * a2 * b2;
        const double S11 = Ap[3] * a1 * b1 + dd[2] * a0 * b0 + dd[0] * a2 * b2;
        const double S22 = Ap[5] * a2 * b2 + dd[1] * a0 * b0 + dd[0] * a1 * b1;
        const double S01 = Ap[1] * a0 * b1 + od[2] * a1 * b0, S10 = Ap[1] * a1 * b0 + od[2] * a0 * b1;
        const double S02 = Ap[2] * a0 * b2 + od[1] * a2 * b0, S20 = Ap[2] * a2 * b0 + od[1] * a0 * b2;
        const double S12 = Ap[4] * a1 * b2 + od[0] * a2 * b1, S21 = Ap[4] * a2 * b1 + od[0] * a1 * b2;
        double T[3][3]; // U S
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            T[i][0] = U[i][0] * S00 + U[i][1] * S10 + U[i][2] * S20;
            T[i][1] = U[i][0] * S01 + U[i][1] * S11 + U[i][2] * S21;
            T[i][2] = U[i][0] * S02 + U[i][1] * S12 + U[i][2] * S22;
        }
#define BIJ(i, j) (T[i][0] * U[j][0] + T[i][1] * U[j][1] + T[i][2] * U[j][2])
        if (pr < 4) {
            st_unit(to, 3 * pr + 1, BIJ(0, 0), BIJ(0, 1), BIJ(0, 2), BIJ(1, 1));
            st_unit(to, 3 * pr + 2, BIJ(1, 2), BIJ(2, 2), 0.0, 0.0);
        } else {
            const int u = 12 + 3 * (pr - 4);
            st_unit(to, u, BIJ(0, 0), BIJ(0, 1), BIJ(0, 2), BIJ(1, 0));
            st_unit(to, u + 1, BIJ(1, 1), BIJ(1, 2), BIJ(2, 0), BIJ(2, 1));
            st_unit(to, u + 2, BIJ(2, 2), 0.0, 0.0, 0.0);
        }
#undef BIJ
    }
}

__device__ __forceinline__ void barrier_fn(double D, double d_hat, double kappa, double* B, double* dB, double* ddB)
{
    const double D0 = d_hat * d_hat;
    if (!(D < D0)) { if (B) *B = 0; if (dB) *dB = 0; if (ddB) *ddB = 0; return; }
    const double t = D - D0, lg = log(D / D0);
    if (B) *B = -kappa * t * t * lg;
    if (dB) *dB = -kappa * (2.0 * t * lg + t * t / D);
    if (ddB) *ddB = -kappa * (2.0 * lg + 4.0 * t / D - t * t / (D * D));
}


// ---- prescribed triangle-mesh indenter (type 2) ----------------------------------------------------------------------------
// Every (gel surface vertex, indenter triangle) candidate is treated as the reference treats a point-triangle candidate: closest
// feature (distance_flagged.h:248-350, same decision order), squared distance of the PT / PE / PP case (details/point_*.inl), one
// barrier per candidate inside d_hat (ipc_simplex_normal_contact.cu:270-342; no de-duplication of shared edges / vertices,
// lbvh_simplex_trajectory_filter.cu:600-690), restricted to the gel vertex's degrees of freedom. dD/dp = 2 (p - closest point) in
// all three cases, and each candidate's make_spd'ed block has the closed form max(0, B'' + B' / (2 D)) g g^T (pinned against the
// reference's Hessian in tests/test_fem_ref_pin_cpu.py), so neither d2D/dp2 nor an eigen-decomposition is needed here.
// The triangles are static in the indenter's frame and shared by all gels (L1 / L2 resident, every lane of a warp reads the same
// triangle: a broadcast); the broad phase is a box test per triangle.
__device__ __forceinline__ double pt_distance2(const double* __restrict__ tr, const double p[3], double g[3])
{
    const double e01[3] = {tr[3] - tr[0], tr[4] - tr[1], tr[5] - tr[2]}, e02[3] = {tr[6] - tr[0], tr[7] - tr[1], tr[8] - tr[2]};
    const double n[3] = {e01[1] * e02[2] - e01[2] * e02[1], e01[2] * e02[0] - e01[0] * e02[2], e01[0] * e02[1] - e01[1] * e02[0]};
    double av[3];
    int kind = -1;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (kind >= 0) continue;
        const double* s = tr + 3 * k;
        const double* t = tr + 3 * ((k + 1) % 3);
        const double e[3] = {t[0] - s[0], t[1] - s[1], t[2] - s[2]}, q[3] = {p[0] - s[0], p[1] - s[1], p[2] - s[2]};
        const double m[3] = {e[1] * n[2] - e[2] * n[1], e[2] * n[0] - e[0] * n[2], e[0] * n[1] - e[1] * n[0]};
        av[k] = (e[0] * q[0] + e[1] * q[1] + e[2] * q[2]) / (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
        const double b = (m[0] * q[0] + m[1] * q[1] + m[2] * q[2]) / (m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
        if (av[k] > 0.0 && av[k] < 1.0 && b >= 0.0) kind = 1 + k;
    }
    if (kind < 0) {
        if (av[0] <= 0.0 && av[2] >= 1.0) kind = 4;
        else if (av[1] <= 0.0 && av[0] >= 1.0) kind = 5;
        else if (av[2] <= 0.0 && av[1] >= 1.0) kind = 6;
        else kind = 0;
    }
    double D;
    if (kind >= 4) {
        const double* v = tr + 3 * (kind - 4);
        const double r[3] = {p[0] - v[0], p[1] - v[1], p[2] - v[2]};
        D = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
        g[0] = 2.0 * r[0]; g[1] = 2.0 * r[1]; g[2] = 2.0 * r[2];
    } else if (kind >= 1) {
        const double* s = tr + 3 * (kind - 1);
        const double* t = tr + 3 * (kind % 3);
        const double u[3] = {t[0] - s[0], t[1] - s[1], t[2] - s[2]}, q[3] = {p[0] - s[0], p[1] - s[1], p[2] - s[2]};
        const double uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2], qu = q[0] * u[0] + q[1] * u[1] + q[2] * u[2];
        const double c[3] = {q[1] * u[2] - q[2] * u[1], q[2] * u[0] - q[0] * u[2], q[0] * u[1] - q[1] * u[0]};
        D = (c[0] * c[0] + c[1] * c[1] + c[2] * c[2]) / uu;
        const double f = qu / uu;
        g[0] = 2.0 * (q[0] - f * u[0]); g[1] = 2.0 * (q[1] - f * u[1]); g[2] = 2.0 * (q[2] - f * u[2]);
    } else {
        const double nn = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
        const double sd = n[0] * (p[0] - tr[0]) + n[1] * (p[1] - tr[1]) + n[2] * (p[2] - tr[2]);
        D = sd * sd / nn;
        const double f = 2.0 * sd / nn;
        g[0] = f * n[0]; g[1] = f * n[1]; g[2] = f * n[2];
    }
    return D;
}

// One pass over the triangles: nearest distance d + direction n (world), and -- when kdt2 > 0 -- the summed barrier energy, gradient
// (world) and projected Hessian block (world, symmetric 00 01 02 11 12 22) of the candidates inside d_hat. Returns whether any
// candidate is active. E / G / H6 are ACCUMULATED into.
// Out-of-line, and every input BY VALUE: a reference to the kernel's parameter struct or to an indenter held in registers would force
// both onto the local-memory stack of the (register-tight) step kernel for all indenter kinds.
// visits the primitives of the cells overlapped by the box [qlo, qhi] inflated by infl + rmax
template <class F>
__device__ __forceinline__ void grid_for(const MeshGrid& G, const double qlo[3], const double qhi[3], double infl, F&& body)
{
    int c0[3], c1[3];
    const int dims[3] = {G.nx, G.ny, G.nz};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double e = infl + G.rmax;
        const int i0 = (int)floor((qlo[a] - e - G.lo[a]) * G.inv[a]), i1 = (int)floor((qhi[a] + e - G.lo[a]) * G.inv[a]);
        if (i1 < 0 || i0 >= dims[a]) return;
        c0[a] = i0 < 0 ? 0 : i0;
        c1[a] = i1 >= dims[a] ? dims[a] - 1 : i1;
    }
    for (int cz = c0[2]; cz <= c1[2]; ++cz)
        for (int cy = c0[1]; cy <= c1[1]; ++cy)
            for (int cx = c0[0]; cx <= c1[0]; ++cx) {
                const int cell = (cz * G.ny + cy) * G.nx + cx;
                for (int q = G.start[cell]; q < G.start[cell + 1]; ++q) body(G.ids[q]);
            }
}

struct MeshRef { const double* tri; const double* box; int n; double d_hat; MeshGrid grid; };
struct MeshOut { double d, n[3], E, G[3], H6[6]; };
__device__ __noinline__ bool mesh_contact_impl(MeshRef m, double3 c, double3 r0, double3 r1, double3 r2, double3 xw, double kdt2, MeshOut* o)
{
    const double R[9] = {r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, r2.x, r2.y, r2.z};
    const double q[3] = {xw.x - c.x, xw.y - c.y, xw.z - c.z};
    double p[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) p[i] = R[0 * 3 + i] * q[0] + R[1 * 3 + i] * q[1] + R[2 * 3 + i] * q[2];
    const double D0 = m.d_hat * m.d_hat, Dcull = kdt2 > 0.0 ? D0 : 0.0;
    // search radius 2 d_hat: distances are exact below it and reported as the radius beyond (a candidate farther away can neither
    // carry a barrier nor limit a step of the sizes the solver takes), so far gel primitives skip every distance evaluation
    double best = 4.0 * D0, gb[3] = {0.0, 0.0, 1.0};
    double Es = 0.0, Gl[3] = {0.0, 0.0, 0.0}, Hl[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    bool active = false;
    grid_for(m.grid, p, p, 2.0 * m.d_hat, [&](int t) {
        const double* bx = m.box + 6 * t;
        double bd = 0.0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double lo = bx[c], hi = bx[3 + c];
            const double dd = p[c] < lo ? lo - p[c] : (p[c] > hi ? p[c] - hi : 0.0);
            bd += dd * dd;
        }
        if (!(bd < best) && !(bd < Dcull)) return;
        double g[3];
        const double D = pt_distance2(m.tri + 9 * t, p, g);
        if (D < best) { best = D; gb[0] = g[0]; gb[1] = g[1]; gb[2] = g[2]; }
        if (kdt2 > 0.0 && D < D0 && D > 0.0) {
            double B, dB, ddB;
            barrier_fn(D, m.d_hat, kdt2, &B, &dB, &ddB);
            active = true;
            Es += B;
            Gl[0] += dB * g[0]; Gl[1] += dB * g[1]; Gl[2] += dB * g[2];
            const double w = ddB + dB / (2.0 * D);
            if (w > 0.0) {
                Hl[0] += w * g[0] * g[0]; Hl[1] += w * g[0] * g[1]; Hl[2] += w * g[0] * g[2];
                Hl[3] += w * g[1] * g[1]; Hl[4] += w * g[1] * g[2]; Hl[5] += w * g[2] * g[2];
            }
        }
    });
    o->d = sqrt(best);
    {
        const double l = sqrt(gb[0] * gb[0] + gb[1] * gb[1] + gb[2] * gb[2]);
#pragma unroll
        for (int i = 0; i < 3; ++i)
            o->n[i] = l > 0.0 ? (R[i * 3 + 0] * gb[0] + R[i * 3 + 1] * gb[1] + R[i * 3 + 2] * gb[2]) / l : (i == 2 ? 1.0 : 0.0);
    }
    o->E = Es;
#pragma unroll
    for (int i = 0; i < 3; ++i) o->G[i] = R[i * 3 + 0] * Gl[0] + R[i * 3 + 1] * Gl[1] + R[i * 3 + 2] * Gl[2];
    { // R Hl R^T
        const double Hf[9] = {Hl[0], Hl[1], Hl[2], Hl[1], Hl[3], Hl[4], Hl[2], Hl[4], Hl[5]};
        double T[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) T[3 * i + j] = R[i * 3 + 0] * Hf[j] + R[i * 3 + 1] * Hf[3 + j] + R[i * 3 + 2] * Hf[6 + j];
        int k = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = i; j < 3; ++j, ++k) o->H6[k] = T[3 * i + 0] * R[j * 3 + 0] + T[3 * i + 1] * R[j * 3 + 1] + T[3 * i + 2] * R[j * 3 + 2];
    }
    return active;
}

// inlined front end: nearest distance d + direction n (world) and -- when kdt2 > 0 -- energy / gradient / projected block ACCUMULATED
// into E / G / H6 (any may be null). Returns whether a candidate is active.
__device__ __forceinline__ bool mesh_contact(const FemArgs& a, const FemIndenter& I, const double* x, double kdt2, double* d, double* n,
                                             double* E, double* G, double* H6)
{
    MeshOut o;
    const MeshRef m{a.mesh_tri, a.mesh_box, a.mesh_n, a.d_hat, a.grid_tri};
    const bool active = mesh_contact_impl(m, make_double3(I.c[0], I.c[1], I.c[2]), make_double3(I.R[0], I.R[1], I.R[2]),
                                          make_double3(I.R[3], I.R[4], I.R[5]), make_double3(I.R[6], I.R[7], I.R[8]),
                                          make_double3(x[0], x[1], x[2]), kdt2, &o);
    if (d) *d = o.d;
    if (n) { n[0] = o.n[0]; n[1] = o.n[1]; n[2] = o.n[2]; }
    if (!active) return false;
    if (E) *E += o.E;
    if (G) { G[0] += o.G[0]; G[1] += o.G[1]; G[2] += o.G[2]; }
    if (H6) {
#pragma unroll
        for (int k = 0; k < 6; ++k) H6[k] += o.H6[k];
    }
    return true;
}

// ---- indenter VERTEX against gel TRIANGLE candidates (second half of the vertex-face contact) ------------------------------------
// Same closest-feature classification and squared distance as above; the unknowns are the triangle's three vertices. By the envelope
// theorem dD/dt_j = -2 w_j r with r = p - c, c = sum w_j t_j the closest point (pinned against the reference's 12-gradient); the
// Hessian is Gauss-Newton (weights frozen), which after make_spd leaves max(0, B'' + B' / (2 D)) g g^T (see oracle/fem_canon.c).
__device__ __forceinline__ double pt_closest(const double* __restrict__ tr, const double p[3], double r[3], double w[3])
{
    const double e01[3] = {tr[3] - tr[0], tr[4] - tr[1], tr[5] - tr[2]}, e02[3] = {tr[6] - tr[0], tr[7] - tr[1], tr[8] - tr[2]};
    const double n[3] = {e01[1] * e02[2] - e01[2] * e02[1], e01[2] * e02[0] - e01[0] * e02[2], e01[0] * e02[1] - e01[1] * e02[0]};
    double av[3];
    int kind = -1;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (kind >= 0) continue;
        const double* s = tr + 3 * k;
        const double* t = tr + 3 * ((k + 1) % 3);
        const double e[3] = {t[0] - s[0], t[1] - s[1], t[2] - s[2]}, q[3] = {p[0] - s[0], p[1] - s[1], p[2] - s[2]};
        const double m[3] = {e[1] * n[2] - e[2] * n[1], e[2] * n[0] - e[0] * n[2], e[0] * n[1] - e[1] * n[0]};
        av[k] = (e[0] * q[0] + e[1] * q[1] + e[2] * q[2]) / (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
        const double b = (m[0] * q[0] + m[1] * q[1] + m[2] * q[2]) / (m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
        if (av[k] > 0.0 && av[k] < 1.0 && b >= 0.0) kind = 1 + k;
    }
    if (kind < 0) {
        if (av[0] <= 0.0 && av[2] >= 1.0) kind = 4;
        else if (av[1] <= 0.0 && av[0] >= 1.0) kind = 5;
        else if (av[2] <= 0.0 && av[1] >= 1.0) kind = 6;
        else kind = 0;
    }
    double D;
    w[0] = w[1] = w[2] = 0.0;
    if (kind >= 4) {
        const double* v = tr + 3 * (kind - 4);
        const double rr[3] = {p[0] - v[0], p[1] - v[1], p[2] - v[2]};
        D = rr[0] * rr[0] + rr[1] * rr[1] + rr[2] * rr[2];
        if (kind == 4) w[0] = 1.0; else if (kind == 5) w[1] = 1.0; else w[2] = 1.0;
    } else if (kind >= 1) {
        const double* s = tr + 3 * (kind - 1);
        const double* t = tr + 3 * (kind % 3);
        const double u[3] = {t[0] - s[0], t[1] - s[1], t[2] - s[2]}, q[3] = {p[0] - s[0], p[1] - s[1], p[2] - s[2]};
        const double uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2], qu = q[0] * u[0] + q[1] * u[1] + q[2] * u[2];
        const double c[3] = {q[1] * u[2] - q[2] * u[1], q[2] * u[0] - q[0] * u[2], q[0] * u[1] - q[1] * u[0]};
        D = (c[0] * c[0] + c[1] * c[1] + c[2] * c[2]) / uu;
        const double al = qu / uu;
        if (kind == 1) { w[0] = 1.0 - al; w[1] = al; }
        else if (kind == 2) { w[1] = 1.0 - al; w[2] = al; }
        else { w[2] = 1.0 - al; w[0] = al; }
    } else {
        const double nn = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
        const double sd = n[0] * (p[0] - tr[0]) + n[1] * (p[1] - tr[1]) + n[2] * (p[2] - tr[2]);
        D = sd * sd / nn;
        const double f = sd / nn;
        const double c[3] = {p[0] - f * n[0] - tr[0], p[1] - f * n[1] - tr[1], p[2] - f * n[2] - tr[2]};
        const double d00 = e01[0] * e01[0] + e01[1] * e01[1] + e01[2] * e01[2], d01 = e01[0] * e02[0] + e01[1] * e02[1] + e01[2] * e02[2],
                     d11 = e02[0] * e02[0] + e02[1] * e02[1] + e02[2] * e02[2], d20 = c[0] * e01[0] + c[1] * e01[1] + c[2] * e01[2],
                     d21 = c[0] * e02[0] + c[1] * e02[1] + c[2] * e02[2], den = d00 * d11 - d01 * d01;
        w[1] = (d11 * d20 - d01 * d21) / den;
        w[2] = (d00 * d21 - d01 * d20) / den;
        w[0] = 1.0 - w[1] - w[2];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) r[a] = p[a] - (w[0] * tr[a] + w[1] * tr[3 + a] + w[2] * tr[6 + a]);
    return D;
}

// The loops over the indenter's primitives run in the INDENTER'S frame: the gel primitive of the thread is transformed once
// (R^T (x - c)), the mesh's vertices are used as stored, and the gradients / symmetric blocks are rotated back at the end.
__device__ __forceinline__ void to_local3(const double R[9], double3 c, double3 x, double o[3])
{
    const double q[3] = {x.x - c.x, x.y - c.y, x.z - c.z};
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = R[0 * 3 + i] * q[0] + R[1 * 3 + i] * q[1] + R[2 * 3 + i] * q[2];
}
__device__ __forceinline__ void rot_dir3(const double R[9], double3 d, double o[3]) // R^T d
{
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = R[0 * 3 + i] * d.x + R[1 * 3 + i] * d.y + R[2 * 3 + i] * d.z;
}
__device__ __forceinline__ void rot_vec_back(const double R[9], double* v) // v <- R v
{
    const double a = v[0], b = v[1], c = v[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) v[i] = R[3 * i] * a + R[3 * i + 1] * b + R[3 * i + 2] * c;
}
__device__ __forceinline__ void rot_sym_back(const double R[9], double* h) // symmetric 00 01 02 11 12 22 <- R H R^T
{
    const double Hf[9] = {h[0], h[1], h[2], h[1], h[3], h[4], h[2], h[4], h[5]};
    double T[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) T[3 * i + j] = R[3 * i] * Hf[j] + R[3 * i + 1] * Hf[3 + j] + R[3 * i + 2] * Hf[6 + j];
    int k = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = i; j < 3; ++j, ++k) h[k] = T[3 * i] * R[3 * j] + T[3 * i + 1] * R[3 * j + 1] + T[3 * i + 2] * R[3 * j + 2];
}

struct MeshVerts { const double* vert; int nv; double d_hat; MeshGrid grid; };
// per gel triangle: k 0..8 gradient of its three vertices, 9..26 their diagonal blocks (symmetric 00 01 02 11 12 22), 27..44 the
// blocks of the pairs (0,1) (0,2) (1,2) (multiples of r r^T: symmetric)
struct TpOut { double E, dmin2, v[45]; int bad; };
__device__ __noinline__ void tp_terms_impl(MeshVerts mv, double3 c, double3 r0, double3 r1, double3 r2, double3 xa, double3 xb, double3 xc,
                                           double kdt2, int derivs, TpOut* o)
{
    const double R[9] = {r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, r2.x, r2.y, r2.z};
    double tr[9];
    to_local3(R, c, xa, tr); to_local3(R, c, xb, tr + 3); to_local3(R, c, xc, tr + 6);
    double lo[3], hi[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = fmin(tr[a], fmin(tr[3 + a], tr[6 + a]));
        hi[a] = fmax(tr[a], fmax(tr[3 + a], tr[6 + a]));
    }
    const double D0 = mv.d_hat * mv.d_hat;
    double E = 0.0, best = 4.0 * D0;
    int bad = 0;
    if (derivs)
        for (int k = 0; k < 45; ++k) o->v[k] = 0.0;
    grid_for(mv.grid, lo, hi, 2.0 * mv.d_hat, [&](int k) {
        const double* l = mv.vert + 3 * k;
        const double pw[3] = {l[0], l[1], l[2]};
        double bd = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double dd = pw[a] < lo[a] ? lo[a] - pw[a] : (pw[a] > hi[a] ? pw[a] - hi[a] : 0.0);
            bd += dd * dd;
        }
        if (!(bd < best) && !(bd < D0)) return;
        double r[3], w[3];
        const double D = pt_closest(tr, pw, r, w);
        if (D < best) best = D;
        if (!(D < D0)) return;
        if (!(D > 0.0)) { bad = 1; return; }
        double B, dB, ddB;
        barrier_fn(D, mv.d_hat, kdt2, &B, &dB, &ddB);
        E += B;
        if (derivs) {
            const double we = ddB + dB / (2.0 * D), w4 = we > 0.0 ? 4.0 * we : 0.0;
            const double rr[6] = {r[0] * r[0], r[0] * r[1], r[0] * r[2], r[1] * r[1], r[1] * r[2], r[2] * r[2]};
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const double gj = -2.0 * dB * w[j];
                o->v[3 * j] += gj * r[0]; o->v[3 * j + 1] += gj * r[1]; o->v[3 * j + 2] += gj * r[2];
                const double dj = w4 * w[j] * w[j];
#pragma unroll
                for (int q = 0; q < 6; ++q) o->v[9 + 6 * j + q] += dj * rr[q];
            }
            const double p01 = w4 * w[0] * w[1], p02 = w4 * w[0] * w[2], p12 = w4 * w[1] * w[2];
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                o->v[27 + q] += p01 * rr[q];
                o->v[33 + q] += p02 * rr[q];
                o->v[39 + q] += p12 * rr[q];
            }
        }
    });
    if (derivs) {
#pragma unroll
        for (int j = 0; j < 3; ++j) rot_vec_back(R, o->v + 3 * j);
#pragma unroll
        for (int j = 0; j < 6; ++j) rot_sym_back(R, o->v + 9 + 6 * j);
    }
    o->E = E; o->dmin2 = best; o->bad = bad;
}

// ACCD of a moving gel triangle (vertices + displacements) against the static vertices of the indenter: min(1, min toc)
__device__ __noinline__ double tp_ccd_impl(MeshVerts mv, double3 c, double3 r0, double3 r1, double3 r2, double3 xa, double3 xb, double3 xc,
                                           double3 da, double3 db, double3 dc)
{
    const double R[9] = {r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, r2.x, r2.y, r2.z};
    double t0[9], d0[9];
    to_local3(R, c, xa, t0); to_local3(R, c, xb, t0 + 3); to_local3(R, c, xc, t0 + 6);
    rot_dir3(R, da, d0); rot_dir3(R, db, d0 + 3); rot_dir3(R, dc, d0 + 6);
    double lo[3], hi[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = 1e300; hi[a] = -1e300;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double u = t0[3 * j + a], v = u + d0[3 * j + a];
            lo[a] = fmin(lo[a], fmin(u, v));
            hi[a] = fmax(hi[a], fmax(u, v));
        }
    }
    const double eta = 0.1;
    double alpha = 1.0;
    grid_for(mv.grid, lo, hi, mv.d_hat, [&](int k) {
        const double* l = mv.vert + 3 * k;
        double p[3] = {l[0], l[1], l[2]};
        bool far = false;
#pragma unroll
        for (int a = 0; a < 3; ++a)
            if (p[a] - hi[a] > mv.d_hat || lo[a] - p[a] > mv.d_hat) far = true;
        if (far) return;
        double tr[9], dt[9], dp[3];
        double mm = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double mov = (d0[a] + d0[3 + a] + d0[6 + a] + 0.0) / 4;
            dp[a] = 0.0 - mov;
            dt[a] = d0[a] - mov; dt[3 + a] = d0[3 + a] - mov; dt[6 + a] = d0[6 + a] - mov;
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) mm = fmax(mm, dt[3 * j] * dt[3 * j] + dt[3 * j + 1] * dt[3 * j + 1] + dt[3 * j + 2] * dt[3 * j + 2]);
#pragma unroll
        for (int q = 0; q < 9; ++q) tr[q] = t0[q];
        const double L = sqrt(dp[0] * dp[0] + dp[1] * dp[1] + dp[2] * dp[2]) + sqrt(mm);
        if (L <= 0.0) return;
        double g[3];
        double d2 = pt_distance2(tr, p, g), d = sqrt(d2);
        const double gap = eta * d2 / d, toc_prev = 1.1;
        double toc = 0.0;
        bool hit = true;
        for (int it = 1000;;) {
            if (--it < 0) break;
            const double lb = (1 - eta) * d2 / (d * L);
#pragma unroll
            for (int a = 0; a < 3; ++a) p[a] += lb * dp[a];
#pragma unroll
            for (int q = 0; q < 9; ++q) tr[q] += lb * dt[q];
            d2 = pt_distance2(tr, p, g);
            d = sqrt(d2);
            if (toc != 0.0 && d2 / d < gap) break;
            toc += lb;
            if (toc > toc_prev) { hit = false; break; }
        }
        if (hit && toc < alpha) alpha = toc;
    });
    return alpha;
}

// ---- edge-edge candidates: gel contact EDGE (two unknown vertices) against a static indenter EDGE --------------------------------
// Decision order of the reference's edge_edge_distance_flag (distance_flagged.h:352-487), squared distance of the resulting EE / PE /
// PP case, closest-point parameters s (gel edge) and t; gradient 2 r (x) [(1 - s), s] by the envelope theorem (pinned against the
// reference's 12-gradient). Returns the flag bits 8 a0 + 4 a1 + 2 b0 + b1 (15 = interior: the only mollified case).
__device__ __forceinline__ int ee_closest(const double a0[3], const double a1[3], const double b0[3], const double b1[3], double& D,
                                          double r[3], double& s_, double& t_)
{
    const double u[3] = {a1[0] - a0[0], a1[1] - a0[1], a1[2] - a0[2]}, v[3] = {b1[0] - b0[0], b1[1] - b0[1], b1[2] - b0[2]},
                 w[3] = {a0[0] - b0[0], a0[1] - b0[1], a0[2] - b0[2]};
    const double a = u[0] * u[0] + u[1] * u[1] + u[2] * u[2], b = u[0] * v[0] + u[1] * v[1] + u[2] * v[2],
                 c = v[0] * v[0] + v[1] * v[1] + v[2] * v[2], d = u[0] * w[0] + u[1] * w[1] + u[2] * w[2],
                 e = v[0] * w[0] + v[1] * w[1] + v[2] * w[2];
    const double Dn = a * c - b * b;
    double tD = Dn, tN;
    const double sN = b * e - c * d;
    const double x[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
    const double xx = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
    int F = 15;
    if (sN <= 0.0) { tN = e; tD = c; F = 8 | 2 | 1; }
    else if (sN >= Dn) { tN = e + b; tD = c; F = 4 | 2 | 1; }
    else {
        tN = a * e - b * d;
        if (tN > 0.0 && tN < tD && ((x[0] * w[0] + x[1] * w[1] + x[2] * w[2]) == 0.0 || xx < 1.0e-20 * a * c)) {
            if (sN < Dn / 2) { tN = e; tD = c; F = 8 | 2 | 1; }
            else { tN = e + b; tD = c; F = 4 | 2 | 1; }
        }
    }
    if (tN <= 0.0) {
        if (-d <= 0.0) F = 8 | 2;
        else if (-d >= a) F = 4 | 2;
        else F = 8 | 4 | 2;
    } else if (tN >= tD) {
        if ((-d + b) <= 0.0) F = 8 | 1;
        else if ((-d + b) >= a) F = 4 | 1;
        else F = 8 | 4 | 1;
    }
    switch (F) {
    case 15: s_ = sN / Dn; t_ = tN / Dn; break;
    case 8 | 2 | 1: s_ = 0.0; t_ = e / c; break;
    case 4 | 2 | 1: s_ = 1.0; t_ = (e + b) / c; break;
    case 8 | 4 | 2: t_ = 0.0; s_ = -d / a; break;
    case 8 | 4 | 1: t_ = 1.0; s_ = (-d + b) / a; break;
    case 8 | 2: s_ = 0.0; t_ = 0.0; break;
    case 4 | 2: s_ = 1.0; t_ = 0.0; break;
    case 8 | 1: s_ = 0.0; t_ = 1.0; break;
    default: s_ = 1.0; t_ = 1.0; break;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) r[k] = (a0[k] + s_ * u[k]) - (b0[k] + t_ * v[k]);
    if (F == 15) {
        const double q = -(w[0] * x[0] + w[1] * x[1] + w[2] * x[2]);
        D = q * q / xx;
    } else if (F == (8 | 2) || F == (4 | 2) || F == (8 | 1) || F == (4 | 1)) {
        D = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    } else {
        const double* P = F == (8 | 2 | 1) ? a0 : (F == (4 | 2 | 1) ? a1 : (F == (8 | 4 | 2) ? b0 : b1));
        const bool on_a = (F & 8) && (F & 4);
        const double* E0 = on_a ? a0 : b0;
        const double* E1 = on_a ? a1 : b1;
        const double p0[3] = {E0[0] - P[0], E0[1] - P[1], E0[2] - P[2]}, p1[3] = {E1[0] - P[0], E1[1] - P[1], E1[2] - P[2]},
                     ed[3] = {E1[0] - E0[0], E1[1] - E0[1], E1[2] - E0[2]};
        const double cr[3] = {p0[1] * p1[2] - p0[2] * p1[1], p0[2] * p1[0] - p0[0] * p1[2], p0[0] * p1[1] - p0[1] * p1[0]};
        D = (cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]) / (ed[0] * ed[0] + ed[1] * ed[1] + ed[2] * ed[2]);
    }
    return F;
}

struct MeshEdges { const double* vert; const int* edge; int ne; double d_hat; MeshGrid grid; };
// per gel contact edge: k 0..5 gradient of its two vertices, 6..17 their diagonal blocks (symmetric), 18..23 the pair block
struct EeOut { double E, dmin2, v[24]; int bad; };
__device__ __noinline__ void ee_terms_impl(MeshEdges me, double3 c, double3 r0, double3 r1, double3 r2, double3 xa, double3 xb, double len2,
                                           double kdt2, int derivs, EeOut* o)
{
    const double R[9] = {r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, r2.x, r2.y, r2.z};
    double a0[3], a1[3];
    to_local3(R, c, xa, a0); to_local3(R, c, xb, a1);
    const double D0 = me.d_hat * me.d_hat;
    double E = 0.0, best = 4.0 * D0;
    int bad = 0;
    if (derivs)
        for (int k = 0; k < 24; ++k) o->v[k] = 0.0;
    double elo[3], ehi[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { elo[k] = fmin(a0[k], a1[k]); ehi[k] = fmax(a0[k], a1[k]); }
    grid_for(me.grid, elo, ehi, 2.0 * me.d_hat, [&](int q) {
        const double* l0 = me.vert + 3 * me.edge[2 * q];
        const double* l1 = me.vert + 3 * me.edge[2 * q + 1];
        double b0[3] = {l0[0], l0[1], l0[2]}, b1[3] = {l1[0], l1[1], l1[2]};
        double bd = 0.0, vv = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double blo = fmin(b0[k], b1[k]), bhi = fmax(b0[k], b1[k]), alo = fmin(a0[k], a1[k]), ahi = fmax(a0[k], a1[k]);
            const double gap = blo > ahi ? blo - ahi : (alo > bhi ? alo - bhi : 0.0);
            bd += gap * gap;
            vv += (b1[k] - b0[k]) * (b1[k] - b0[k]);
        }
        if (!(bd < best) && !(bd < D0)) return;
        double D, r[3], s_, t_;
        const int F = ee_closest(a0, a1, b0, b1, D, r, s_, t_);
        if (D < best) best = D;
        if (!(D < D0)) return;
        if (!(D > 0.0)) { bad = 1; return; }
        double B, dB, ddB;
        barrier_fn(D, me.d_hat, kdt2, &B, &dB, &ddB);
        double ek = 1.0, dek = 0.0, du[3] = {0.0, 0.0, 0.0};
        if (F == 15) { // mollifier of nearly parallel edges: e_k(|u x v|^2), threshold 1e-3 |u_rest|^2 |v|^2
            const double u[3] = {a1[0] - a0[0], a1[1] - a0[1], a1[2] - a0[2]}, v[3] = {b1[0] - b0[0], b1[1] - b0[1], b1[2] - b0[2]};
            const double uv = u[0] * v[0] + u[1] * v[1] + u[2] * v[2];
            const double x[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
            const double cn = x[0] * x[0] + x[1] * x[1] + x[2] * x[2], eps = 1.0e-3 * len2 * vv;
            if (cn < eps) {
                const double qq = cn / eps;
                ek = (-qq + 2.0) * qq;
                dek = 2.0 / eps * (-qq + 1.0);
#pragma unroll
                for (int k = 0; k < 3; ++k) du[k] = 2.0 * (vv * u[k] - uv * v[k]);
            }
        }
        E += ek * B;
        if (derivs) {
            const double g0 = 2.0 * (1.0 - s_), g1 = 2.0 * s_;
            const double we = ek * (ddB + dB / (2.0 * D)), wp = we > 0.0 ? we : 0.0;
            const double rr[6] = {r[0] * r[0], r[0] * r[1], r[0] * r[2], r[1] * r[1], r[1] * r[2], r[2] * r[2]};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                o->v[k] += ek * dB * g0 * r[k] - B * dek * du[k];
                o->v[3 + k] += ek * dB * g1 * r[k] + B * dek * du[k];
            }
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                o->v[6 + k] += wp * g0 * g0 * rr[k];
                o->v[12 + k] += wp * g1 * g1 * rr[k];
                o->v[18 + k] += wp * g0 * g1 * rr[k];
            }
        }
    });
    if (derivs) {
        rot_vec_back(R, o->v); rot_vec_back(R, o->v + 3);
        rot_sym_back(R, o->v + 6); rot_sym_back(R, o->v + 12); rot_sym_back(R, o->v + 18);
    }
    o->E = E; o->dmin2 = best; o->bad = bad;
}

// ACCD of a moving gel edge against the static edges of the indenter (ccd.inl:267-354): min(1, min toc)
__device__ __forceinline__ double ee_dist2_ccd(const double a0[3], const double a1[3], const double b0[3], const double b1[3])
{
    double D, r[3], s_, t_;
    ee_closest(a0, a1, b0, b1, D, r, s_, t_);
    if (D <= 0.0) { // far away, nearly parallel: the smallest end-point distance stands in (as in the reference)
        D = 1e300;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const double* P = i ? a1 : a0;
                const double* Q = j ? b1 : b0;
                const double q = (P[0] - Q[0]) * (P[0] - Q[0]) + (P[1] - Q[1]) * (P[1] - Q[1]) + (P[2] - Q[2]) * (P[2] - Q[2]);
                D = fmin(D, q);
            }
    }
    return D;
}
__device__ __noinline__ double ee_ccd_impl(MeshEdges me, double3 c, double3 r0, double3 r1, double3 r2, double3 xa, double3 xb, double3 da,
                                           double3 db)
{
    const double R[9] = {r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, r2.x, r2.y, r2.z};
    double A0[3], A1[3], dA0[3], dA1[3];
    to_local3(R, c, xa, A0); to_local3(R, c, xb, A1);
    rot_dir3(R, da, dA0); rot_dir3(R, db, dA1);
    double lo[3], hi[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double u0 = A0[k], u1 = u0 + dA0[k], w0 = A1[k], w1 = w0 + dA1[k];
        lo[k] = fmin(fmin(u0, u1), fmin(w0, w1));
        hi[k] = fmax(fmax(u0, u1), fmax(w0, w1));
    }
    const double eta = 0.1;
    double alpha = 1.0;
    grid_for(me.grid, lo, hi, me.d_hat, [&](int q) {
        const double* l0 = me.vert + 3 * me.edge[2 * q];
        const double* l1 = me.vert + 3 * me.edge[2 * q + 1];
        double b0[3] = {l0[0], l0[1], l0[2]}, b1[3] = {l1[0], l1[1], l1[2]};
        bool far = false;
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (fmin(b0[k], b1[k]) - hi[k] > me.d_hat || lo[k] - fmax(b0[k], b1[k]) > me.d_hat) far = true;
        if (far) return;
        double a0[3], a1[3], da0[3], da1[3], db0[3], db1[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double mov = (dA0[k] + dA1[k] + 0.0 + 0.0) / 4;
            a0[k] = A0[k]; a1[k] = A1[k];
            da0[k] = dA0[k] - mov; da1[k] = dA1[k] - mov; db0[k] = 0.0 - mov; db1[k] = 0.0 - mov;
        }
        const double na0 = da0[0] * da0[0] + da0[1] * da0[1] + da0[2] * da0[2], na1 = da1[0] * da1[0] + da1[1] * da1[1] + da1[2] * da1[2],
                     nb0 = db0[0] * db0[0] + db0[1] * db0[1] + db0[2] * db0[2], nb1 = db1[0] * db1[0] + db1[1] * db1[1] + db1[2] * db1[2];
        const double L = sqrt(na0 > na1 ? na0 : na1) + sqrt(nb0 > nb1 ? nb0 : nb1);
        if (L == 0.0) return;
        double d2 = ee_dist2_ccd(a0, a1, b0, b1), d = sqrt(d2);
        const double gap = eta * d2 / d, toc_prev = 1.1;
        double toc = 0.0;
        bool hit = true;
        for (int it = 1000;;) {
            if (--it < 0) break;
            const double lb = (1 - eta) * d2 / (d * L);
#pragma unroll
            for (int k = 0; k < 3; ++k) { a0[k] += lb * da0[k]; a1[k] += lb * da1[k]; b0[k] += lb * db0[k]; b1[k] += lb * db1[k]; }
            d2 = ee_dist2_ccd(a0, a1, b0, b1);
            d = sqrt(d2);
            if (toc != 0.0 && d2 / d < gap) break;
            toc += lb;
            if (toc > toc_prev) { hit = false; break; }
        }
        if (hit && toc < alpha) alpha = toc;
    });
    return alpha;
}
__device__ __forceinline__ void ee_terms(const FemArgs& a, const FemIndenter& I, const double* xs, int ce, double kdt2, int derivs, EeOut* o)
{
    const MeshEdges me{a.mesh_vert, a.mesh_edge, a.mesh_ne, a.d_hat, a.grid_edge};
    const int i0 = a.cedge[2 * ce], i1 = a.cedge[2 * ce + 1];
    ee_terms_impl(me, make_double3(I.c[0], I.c[1], I.c[2]), make_double3(I.R[0], I.R[1], I.R[2]), make_double3(I.R[3], I.R[4], I.R[5]),
                  make_double3(I.R[6], I.R[7], I.R[8]), make_double3(xs[3 * i0], xs[3 * i0 + 1], xs[3 * i0 + 2]),
                  make_double3(xs[3 * i1], xs[3 * i1 + 1], xs[3 * i1 + 2]), a.cedge_len2[ce], kdt2, derivs, o);
}

__device__ __forceinline__ void tp_terms(const FemArgs& a, const FemIndenter& I, const double* xs, int f, double kdt2, int derivs, TpOut* o)
{
    const MeshVerts mv{a.mesh_vert, a.mesh_nv, a.d_hat, a.grid_vert};
    const int i0 = a.ctri[3 * f], i1 = a.ctri[3 * f + 1], i2 = a.ctri[3 * f + 2];
    tp_terms_impl(mv, make_double3(I.c[0], I.c[1], I.c[2]), make_double3(I.R[0], I.R[1], I.R[2]), make_double3(I.R[3], I.R[4], I.R[5]),
                  make_double3(I.R[6], I.R[7], I.R[8]), make_double3(xs[3 * i0], xs[3 * i0 + 1], xs[3 * i0 + 2]),
                  make_double3(xs[3 * i1], xs[3 * i1 + 1], xs[3 * i1 + 2]), make_double3(xs[3 * i2], xs[3 * i2 + 1], xs[3 * i2 + 2]), kdt2,
                  derivs, o);
}

// CCD step bound of ONE gel vertex against the (static during a Newton iteration) mesh indenter: the reference's additive CCD per
// (vertex, triangle) pair (utils/distance/details/ccd.inl:200-262: common translation removed, advance by the conservative bound until
// the gap has shrunk to eta = 0.1 of its initial value; at most 1000 iterations, horizon 1.1) behind its broad phase (box of the swept
// point vs the triangle's box inflated by d_hat, ccd.inl:90-122). Returns min(1, min toc). Same arithmetic as oracle/fem_canon.c
// fem_pt_accd, which is pinned against the compiled reference.
__device__ __noinline__ double mesh_ccd_impl(MeshRef m, double3 c, double3 r0, double3 r1, double3 r2, double3 xw, double3 dxw)
{
    const double R[9] = {r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, r2.x, r2.y, r2.z};
    const double q[3] = {xw.x - c.x, xw.y - c.y, xw.z - c.z}, dw[3] = {dxw.x, dxw.y, dxw.z};
    double p0[3], dp0[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        p0[i] = R[0 * 3 + i] * q[0] + R[1 * 3 + i] * q[1] + R[2 * 3 + i] * q[2];
        dp0[i] = R[0 * 3 + i] * dw[0] + R[1 * 3 + i] * dw[1] + R[2 * 3 + i] * dw[2];
    }
    const double eta = 0.1;
    double alpha = 1.0;
    double slo[3], shi[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { slo[a] = dp0[a] < 0 ? p0[a] + dp0[a] : p0[a]; shi[a] = dp0[a] < 0 ? p0[a] : p0[a] + dp0[a]; }
    grid_for(m.grid, slo, shi, m.d_hat, [&](int t) {
        const double* bx = m.box + 6 * t;
        bool far = false;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double lo = dp0[a] < 0 ? p0[a] + dp0[a] : p0[a], hi = dp0[a] < 0 ? p0[a] : p0[a] + dp0[a];
            if (lo - bx[3 + a] > m.d_hat || bx[a] - hi > m.d_hat) far = true;
        }
        if (far) return;
        double p[3], tr[9], dp[3], dt[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double mov = (0.0 + 0.0 + 0.0 + dp0[a]) / 4;
            p[a] = p0[a];
            dp[a] = dp0[a] - mov;
            dt[a] = 0.0 - mov;
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) tr[k] = m.tri[9 * t + k];
        const double mm = dt[0] * dt[0] + dt[1] * dt[1] + dt[2] * dt[2];
        const double L = sqrt(dp[0] * dp[0] + dp[1] * dp[1] + dp[2] * dp[2]) + sqrt(mm);
        if (L <= 0.0) return;
        double g[3];
        double d2 = pt_distance2(tr, p, g), d = sqrt(d2);
        const double gap = eta * d2 / d, toc_prev = 1.1;
        double toc = 0.0;
        bool hit = true;
        for (int it = 1000;;) {
            if (--it < 0) break;
            const double lb = (1 - eta) * d2 / (d * L);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                p[a] += lb * dp[a];
                tr[a] += lb * dt[a]; tr[3 + a] += lb * dt[a]; tr[6 + a] += lb * dt[a];
            }
            d2 = pt_distance2(tr, p, g);
            d = sqrt(d2);
            if (toc != 0.0 && d2 / d < gap) break;
            toc += lb;
            if (toc > toc_prev) { hit = false; break; }
        }
        if (hit && toc < alpha) alpha = toc;
    });
    return alpha;
}
__device__ __forceinline__ double mesh_ccd(const FemArgs& a, const FemIndenter& I, const double* x0, const double* dx)
{
    const MeshRef m{a.mesh_tri, a.mesh_box, a.mesh_n, a.d_hat, a.grid_tri};
    return mesh_ccd_impl(m, make_double3(I.c[0], I.c[1], I.c[2]), make_double3(I.R[0], I.R[1], I.R[2]), make_double3(I.R[3], I.R[4], I.R[5]),
                         make_double3(I.R[6], I.R[7], I.R[8]), make_double3(x0[0], x0[1], x0[2]), make_double3(dx[0], dx[1], dx[2]));
}

template <bool MESH>
__device__ void indenter_sdf(const FemArgs& a, const FemIndenter& I, const double* x, double* d, double* n, double* Hd)
{
    if (I.type == 2) {
        if (MESH) { // triangle mesh: UNSIGNED distance (Hd is not defined: the barrier goes through mesh_contact)
            mesh_contact(a, I, x, 0.0, d, n, nullptr, nullptr, nullptr);
        } else { // no mesh has been set: a type-2 indenter is no indenter
            *d = 1e150; n[0] = 0.0; n[1] = 0.0; n[2] = 1.0;
            if (Hd) for (int k = 0; k < 9; ++k) Hd[k] = 0.0;
        }
        return;
    }
    double p[3], q[3];
    for (int i = 0; i < 3; ++i) q[i] = x[i] - I.c[i];
    for (int i = 0; i < 3; ++i) p[i] = I.R[0 * 3 + i] * q[0] + I.R[1 * 3 + i] * q[1] + I.R[2 * 3 + i] * q[2];
    double nl[3] = {0, 0, 0}, Hl[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (I.type == 0) {
        const double r = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
        *d = r - I.h[0];
        for (int i = 0; i < 3; ++i) nl[i] = p[i] / r;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Hl[i * 3 + j] = ((i == j ? 1.0 : 0.0) - nl[i] * nl[j]) / r;
    } else {
        double u[3], sgn[3], qq[3];
        bool act[3];
        int nact = 0;
        for (int i = 0; i < 3; ++i) {
            sgn[i] = p[i] < 0 ? -1.0 : 1.0;
            qq[i] = fabs(p[i]) - I.h[i];
            act[i] = qq[i] > 0.0;
            u[i] = act[i] ? qq[i] : 0.0;
            nact += act[i] ? 1 : 0;
        }
        if (nact > 0) {
            const double r = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
            *d = r;
            for (int i = 0; i < 3; ++i) nl[i] = sgn[i] * u[i] / r;
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j)
                    Hl[i * 3 + j] = (act[i] && act[j]) ? ((i == j ? 1.0 : 0.0) - nl[i] * nl[j]) / r : 0.0;
        } else {
            int k = 0;
            for (int i = 1; i < 3; ++i)
                if (qq[i] > qq[k]) k = i;
            *d = qq[k];
            nl[k] = sgn[k];
        }
    }
    for (int i = 0; i < 3; ++i) n[i] = I.R[i * 3 + 0] * nl[0] + I.R[i * 3 + 1] * nl[1] + I.R[i * 3 + 2] * nl[2];
    if (Hd) {
        double T[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0;
                for (int k = 0; k < 3; ++k) s += I.R[i * 3 + k] * Hl[k * 3 + j];
                T[i * 3 + j] = s;
            }
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0;
                for (int k = 0; k < 3; ++k) s += T[i * 3 + k] * I.R[j * 3 + k];
                Hd[i * 3 + j] = s;
            }
    }
}

// Dm^-1 is stored structure-of-arrays ([9][T]) so that consecutive threads (tets) read consecutive addresses
__device__ __forceinline__ void tet_W(const double* __restrict__ Bsoa, int t, int T, double W[4][3])
{
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        const double b0 = Bsoa[(0 * 3 + b) * T + t], b1 = Bsoa[(1 * 3 + b) * T + t], b2 = Bsoa[(2 * 3 + b) * T + t];
        W[1][b] = b0;
        W[2][b] = b1;
        W[3][b] = b2;
        W[0][b] = -(b0 + b1 + b2);
    }
}

__device__ __forceinline__ void tet_F(const double* x, const int* e, const double W[4][3], double* F)
{
    double xe[4][3];
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
        for (int a = 0; a < 3; ++a) xe[v][a] = x[3 * e[v] + a];
#pragma unroll
    for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            double s = 0;
#pragma unroll
            for (int v = 0; v < 4; ++v) s += xe[v][a] * W[v][b];
            F[3 * b + a] = s;
        }
}

// ---- lagged Coulomb friction of a surface vertex against the prescribed indenter ------------------------------------------
// ref: contact_system/contact_models/ipc_vertex_half_plane_frictional_contact.cu:29-127, ipc_vertex_half_plane_contact_function.h
// :62-151, codim_ipc_contact_function.h:16-128. Normal force and tangent frame at the start-of-step position against the
// start-of-step indenter (lagged); slip measured relative to the indenter's prescribed translation. E / G / H may be null.
// Lagged contact of a vertex against the mesh indenter: ONE per vertex, the resultant of the candidates' normal forces at the
// start-of-step position. It does not change during the step, so the row's thread computes it once (a pass over the triangles).
struct FrLag { double fn, n[3]; };
// Glag: the vertex's share of the gradient of the candidate families whose unknowns are several vertices (indenter vertex vs gel
// triangle, edge-edge) at the same lagged state -- the resultant covers every candidate the vertex takes part in.
__device__ __forceinline__ FrLag mesh_friction_lag(const FemArgs& a, const FemIndenter& ind0, const double* xp, const double Glag[3])
{
    FrLag l{0.0, {0.0, 0.0, 1.0}};
    double Gb[3] = {0.0, 0.0, 0.0};
    mesh_contact(a, ind0, xp, a.kappa * a.dt * a.dt, nullptr, nullptr, nullptr, Gb, nullptr); // accumulates only when a candidate is active
    Gb[0] += Glag[0]; Gb[1] += Glag[1]; Gb[2] += Glag[2];
    const double fn = sqrt(Gb[0] * Gb[0] + Gb[1] * Gb[1] + Gb[2] * Gb[2]);
    if (!(fn > 0.0)) return l;
    l.fn = fn;
    l.n[0] = -Gb[0] / fn; l.n[1] = -Gb[1] / fn; l.n[2] = -Gb[2] / fn;
    return l;
}

template <bool MESH>
__device__ void friction_terms(const FemArgs& a, const FemIndenter& ind0, const FemIndenter& ind, const double* xp, const double* x,
                               double* E, double* G, double* H6 /* symmetric 00 01 02 11 12 22 */, const FrLag& lag)
{
    double d, n[3], dB;
    double fn;
    if (MESH && ind0.type == 2) {
        if (!(lag.fn > 0.0)) return;
        fn = lag.fn;
        n[0] = lag.n[0]; n[1] = lag.n[1]; n[2] = lag.n[2];
    } else {
        indenter_sdf<MESH>(a, ind0, xp, &d, n, nullptr);
        if (!(d > 0.0) || !(d < a.d_hat)) return;
        barrier_fn(d * d, a.d_hat, a.kappa * a.dt * a.dt, nullptr, &dB, nullptr);
        fn = -dB * 2.0 * d;
    }
    const double mf = a.friction_mu * fn;
    double t[3] = {1.0, 0.0, 0.0};
    if (n[0] > 0.9) { t[0] = 0.0; t[2] = 1.0; }
    const double c[3] = {t[1] * n[2] - t[2] * n[1], t[2] * n[0] - t[0] * n[2], t[0] * n[1] - t[1] * n[0]};
    const double il = fast_rsqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    const double e1[3] = {c[0] * il, c[1] * il, c[2] * il};
    const double e2[3] = {n[1] * e1[2] - n[2] * e1[1], n[2] * e1[0] - n[0] * e1[2], n[0] * e1[1] - n[1] * e1[0]};
    double rel[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) rel[k] = (x[k] - xp[k]) - (ind.c[k] - ind0.c[k]);
    const double u0 = e1[0] * rel[0] + e1[1] * rel[1] + e1[2] * rel[2];
    const double u1 = e2[0] * rel[0] + e2[1] * rel[1] + e2[2] * rel[2];
    const double x2 = u0 * u0 + u1 * u1, eps = a.eps_velocity * a.dt, y = sqrt(x2);
    const bool slip = x2 >= eps * eps;
    if (E) *E += mf * (slip ? y : x2 * (-y / 3.0 + eps) / (eps * eps) + eps / 3.0);
    const double f1 = slip ? 1.0 / y : (-y + 2.0 * eps) / (eps * eps);
    if (G) {
#pragma unroll
        for (int k = 0; k < 3; ++k) G[k] += mf * f1 * (u0 * e1[k] + u1 * e2[k]);
    }
    if (H6) {
        double h00, h01, h11;
        if (slip) {
            const double s = mf * f1 / x2;
            h00 = s * u1 * u1; h01 = -s * u1 * u0; h11 = s * u0 * u0;
        } else if (x2 == 0.0) {
            h00 = h11 = mf * f1; h01 = 0.0;
        } else { // both eigenvalues (f1 - y / eps^2 and f1) are positive: make_spd is the identity
            const double f2 = -1.0 / (eps * eps) / y;
            h00 = mf * (f2 * u0 * u0 + f1); h01 = mf * f2 * u0 * u1; h11 = mf * (f2 * u1 * u1 + f1);
        }
        // J^T H2 J is positive semi-definite by construction (H2 is): the reference's 3x3 make_spd changes nothing
        H6[0] += h00 * e1[0] * e1[0] + 2.0 * h01 * e1[0] * e2[0] + h11 * e2[0] * e2[0];
        H6[1] += h00 * e1[0] * e1[1] + h01 * (e1[0] * e2[1] + e2[0] * e1[1]) + h11 * e2[0] * e2[1];
        H6[2] += h00 * e1[0] * e1[2] + h01 * (e1[0] * e2[2] + e2[0] * e1[2]) + h11 * e2[0] * e2[2];
        H6[3] += h00 * e1[1] * e1[1] + 2.0 * h01 * e1[1] * e2[1] + h11 * e2[1] * e2[1];
        H6[4] += h00 * e1[1] * e1[2] + h01 * (e1[1] * e2[2] + e2[1] * e1[2]) + h11 * e2[1] * e2[2];
        H6[5] += h00 * e1[2] * e1[2] + 2.0 * h01 * e1[2] * e2[2] + h11 * e2[2] * e2[2];
    }
}

// ---- block reductions (deterministic: fixed tree over a fixed thread -> element mapping) -----------------------------
// One __syncthreads per reduction: the per-warp partials are double-buffered (consecutive calls alternate buffers) and every
// thread folds the NW partials itself in the same fixed order.
struct Red {
    double buf[2][FEM_THREADS / 32];
};

template <int OP> // 0 sum, 1 min, 2 max
__device__ __forceinline__ double block_reduce(double v, Red* red, int& phase)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double t = __shfl_xor_sync(0xffffffffu, v, o);
        v = OP == 0 ? v + t : (OP == 1 ? fmin(v, t) : fmax(v, t));
    }
    double* b = red->buf[phase & 1];
    phase ^= 1;
    if ((threadIdx.x & 31) == 0) b[threadIdx.x >> 5] = v;
    __syncthreads();
    // fixed pairwise tree over the NW partials (independent loads, depth log2 NW): identical in every thread
    constexpr int NW = FEM_THREADS / 32;
    double t[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) t[w] = b[w];
#pragma unroll
    for (int stride = 1; stride < NW; stride *= 2)
#pragma unroll
        for (int w = 0; w + stride < NW; w += 2 * stride)
            t[w] = OP == 0 ? t[w] + t[w + stride] : (OP == 1 ? fmin(t[w], t[w + stride]) : fmax(t[w], t[w + stride]));
    const double s = t[0];
    return s;
}

// symmetric packed index of a 9x9 matrix (upper triangle, row-major)
__device__ __forceinline__ int sym9(int i, int j)
{
    const int a = i < j ? i : j, b = i < j ? j : i;
    return a * 9 - a * (a - 1) / 2 + (b - a);
}

// symmetric 3x3 stored as (00, 01, 02, 11, 12, 22)
__device__ __forceinline__ void sym3_mul(const double* m, double p0, double p1, double p2, double& y0, double& y1, double& y2)
{
    y0 = m[0] * p0 + m[1] * p1 + m[2] * p2;
    y1 = m[1] * p0 + m[3] * p1 + m[4] * p2;
    y2 = m[2] * p0 + m[4] * p1 + m[5] * p2;
}

// Shared memory: positions x and the PCG direction p (both read by other rows), reduction scratch, and as many off-diagonal
// 3x3 blocks of the assembled Hessian as fit ([9][n_s], structure of arrays over the edge index); the remaining blocks
// live in an L2-resident global scratch of the CTA. Everything that only its own row touches (gradient, diagonal block,
// preconditioner, r / z / dx of PCG, line-search base) lives in the registers of the thread that owns the row.
struct FemShared {
    double* x;   // [3V]
    double* p;   // [3V]
    double* val; // [9][n_s]
    Red* red;
};

// total incremental potential at the positions in s.x (thread `row` owns vertex `row`)
template <bool MESH>
__device__ double total_energy(const FemArgs& a, const FemShared& s, const double* xt_g,
                               const double* __restrict__ xprev_g, const double* __restrict__ aim_g, const FemIndenter& ind,
                               const FemIndenter& ind0, double ratio, double* min_dist, int& ph, const FrLag& lag)
{
    const double dt2 = a.dt * a.dt;
    const int i = threadIdx.x;
    double E = 0.0, md = 1e300;
    bool bad = false;
    if (i < a.V) {
        const double xi[3] = {s.x[3 * i], s.x[3 * i + 1], s.x[3 * i + 2]};
        const double m = a.mass[i];
        double q = 0;
        for (int c = 0; c < 3; ++c) { const double d = xi[c] - xt_g[3 * i + c]; q += d * d; }
        E += 0.5 * m * q;
        const int ka = a.attach_of[i];
        if (ka >= 0) {
            double q2 = 0;
            for (int c = 0; c < 3; ++c) {
                const double xp = xprev_g[3 * i + c];
                const double aimx = xp + (aim_g[3 * ka + c] - xp) * ratio;
                const double d = xi[c] - aimx;
                q2 += d * d;
            }
            E += 0.5 * a.attach_strength * m * q2;
        }
        if (a.surf_of[i] >= 0) {
            double d, n[3], B;
            if (MESH && ind.type == 2) {
                mesh_contact(a, ind, xi, a.kappa * dt2, &d, nullptr, &E, nullptr, nullptr);
                md = d;
                if (d <= 0.0) bad = true;
            } else {
            indenter_sdf<MESH>(a, ind, xi, &d, n, nullptr);
            md = d;
            if (d <= 0.0) bad = true;
            else {
                barrier_fn(d * d, a.d_hat, a.kappa * dt2, &B, nullptr, nullptr);
                E += B;
            }
            }
            if (a.friction_mu > 0.0) {
                const double xp[3] = {xprev_g[3 * i], xprev_g[3 * i + 1], xprev_g[3 * i + 2]};
                friction_terms<MESH>(a, ind0, ind, xp, xi, &E, nullptr, nullptr, lag);
            }
        }
    }
    if (MESH && ind.type == 2 && a.n_ctri > 0 && threadIdx.x < a.n_ctri) { // indenter vertices against this thread's gel triangle
        TpOut o;
        tp_terms(a, ind, s.x, threadIdx.x, a.kappa * dt2, 0, &o);
        E += o.E;
        md = fmin(md, sqrt(o.dmin2));
        if (o.bad) bad = true;
    }
    if (MESH && ind.type == 2 && a.n_cedge > 0 && threadIdx.x < a.n_cedge) { // indenter edges against this thread's gel edge
        EeOut o;
        ee_terms(a, ind, s.x, threadIdx.x, a.kappa * dt2, 0, &o);
        E += o.E;
        md = fmin(md, sqrt(o.dmin2));
        if (o.bad) bad = true;
    }
    for (int t = threadIdx.x; t < a.T; t += FEM_THREADS) {
        double W[4][3], F[9], e;
        const int4 ev4 = reinterpret_cast<const int4*>(a.tets)[t];
        const int ev[4] = {ev4.x, ev4.y, ev4.z, ev4.w};
        tet_W(a.Dm_inv, t, a.T, W);
        tet_F(s.x, ev, W, F);
        snh(F, a.mu, a.lambda, &e, nullptr, nullptr);
        E += dt2 * a.vol[t] * e;
    }
    E = block_reduce<0>(E, s.red, ph);
    md = block_reduce<1>(md, s.red, ph);
    const double anybad = block_reduce<2>(bad ? 1.0 : 0.0, s.red, ph);
    if (min_dist) *min_dist = md;
    return anybad > 0.0 ? INFINITY : E;
}

// Gradient g3 (row-local), diagonal block d6 (row-local, symmetric) and the off-diagonal blocks of the Hessian (s.val / valg).
// The tets are processed in chunks of one tet per thread so that the per-tet scratch (102 doubles per tet) of all CTAs
// stays L2-resident; rows and edges accumulate their incident tets chunk by chunk, in ascending tet order (no atomics).
template <bool MESH>
__device__ void grad_hess(const FemArgs& a, const FemShared& s, const double* xt_g, const double* __restrict__ xprev_g,
                          const double* __restrict__ aim_g, const FemIndenter& ind, const FemIndenter& ind0, double ratio,
                          double* __restrict__ tsc, double* valg, double g3[3], double d6[6], long long* cyc, const FrLag& lag)
{
    long long tg0 = cyc ? clock64() : 0;
    const double dt2 = a.dt * a.dt;
    const int TC = a.chunk; // tets per chunk: one tet per thread of the first TC threads
    const int i = threadIdx.x;
    const bool on = i < a.V;
    double m = 0.0, xi[3] = {0, 0, 0};
    g3[0] = g3[1] = g3[2] = 0.0;
    for (int j = 0; j < 6; ++j) d6[j] = 0.0;
    if (on) {
        m = a.mass[i];
        for (int c = 0; c < 3; ++c) xi[c] = s.x[3 * i + c];
        for (int c = 0; c < 3; ++c) g3[c] = m * (xi[c] - xt_g[3 * i + c]);
        d6[0] = m; d6[3] = m; d6[5] = m;
    }
    for (int c0 = 0; c0 < a.T; c0 += TC) {
        // (1) per tet: gradient (12), the 4 diagonal and the 6 off-diagonal 3x3 blocks
        const int t = c0 + threadIdx.x;
        if (threadIdx.x < TC && t < a.T) {
            const int4 ev4 = reinterpret_cast<const int4*>(a.tets)[t];
            const int e[4] = {ev4.x, ev4.y, ev4.z, ev4.w};
            double W[4][3], F[9];
            tet_W(a.Dm_inv, t, a.T, W);
            tet_F(s.x, e, W, F);
            tet_contrib(F, W, a.mu, a.lambda, dt2 * a.vol[t], tsc + 4 * threadIdx.x);
        }
        __syncthreads();
        if (cyc) { cyc[0] += clock64() - tg0; tg0 = clock64(); }
        // (2) per vertex (row-local): elastic gradient and diagonal block of the incident tets of this chunk; the entry
        // lists are fetched four at a time so that the dependent index -> scratch loads overlap
        const int ck = c0 / TC;
        if (on) {
            const int q1 = a.row_start[(size_t)(ck + 1) * FEM_THREADS + i];
            for (int qb = a.row_start[(size_t)ck * FEM_THREADS + i]; qb < q1; qb += 4) {
                int tvs[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) tvs[u] = qb + u < q1 ? a.adj[qb + u] : -1;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int tv = tvs[u];
                    if (tv < 0) break;
                    const double* to = tsc + 4 * ((tv >> 2) - c0);
                    const int v = tv & 3;
                    double u0[4], u1[4], u2[4];
                    ld_unit(to, 3 * v, u0);
                    ld_unit(to, 3 * v + 1, u1);
                    ld_unit(to, 3 * v + 2, u2);
                    g3[0] += u0[0]; g3[1] += u0[1]; g3[2] += u0[2];
                    d6[0] += u1[0]; d6[1] += u1[1]; d6[2] += u1[2]; d6[3] += u1[3]; d6[4] += u2[0]; d6[5] += u2[1];
                }
            }
        }
        if (cyc) { cyc[1] += clock64() - tg0; tg0 = clock64(); }
        // (3) per edge (i < j): block A_ij += the incident tets of this chunk
        for (int e = threadIdx.x; e < a.nE; e += FEM_THREADS) {
            double blk[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            const int q0 = a.edge_start[(size_t)ck * a.nE + e], q1 = a.edge_start[(size_t)(ck + 1) * a.nE + e];
            if (q0 == q1 && c0 != 0) continue;
            for (int qb = q0; qb < q1; qb += 4) {
                int ens[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) ens[u] = qb + u < q1 ? a.edge_adj[qb + u] : -1; // tet << 4 | pair slot << 1 | transpose
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int en = ens[u];
                    if (en < 0) break;
                    const double* to = tsc + 4 * ((en >> 4) - c0);
                    const int ps = (en >> 1) & 7;
                    double b[9], u0[4], u1[4], u2[4];
                    ld_unit(to, 12 + 3 * ps, u0);
                    ld_unit(to, 13 + 3 * ps, u1);
                    ld_unit(to, 14 + 3 * ps, u2);
                    b[0] = u0[0]; b[1] = u0[1]; b[2] = u0[2]; b[3] = u0[3]; b[4] = u1[0]; b[5] = u1[1]; b[6] = u1[2]; b[7] = u1[3];
                    b[8] = u2[0];
                    if (en & 1) {
#pragma unroll
                        for (int c = 0; c < 3; ++c)
#pragma unroll
                            for (int bb = 0; bb < 3; ++bb) blk[3 * c + bb] += b[3 * bb + c];
                    } else {
#pragma unroll
                        for (int k = 0; k < 9; ++k) blk[k] += b[k];
                    }
                }
            }
            if (e < a.n_s) {
                double* dst = s.val + e;
#pragma unroll
                for (int k = 0; k < 9; ++k) dst[k * FEM_VAL_STRIDE] = (c0 == 0 ? 0.0 : dst[k * FEM_VAL_STRIDE]) + blk[k];
            } else {
                double* dst = valg + (e - a.n_s);
                const int nEg = a.nE - a.n_s;
#pragma unroll
                for (int k = 0; k < 9; ++k) dst[(size_t)k * nEg] = (c0 == 0 ? 0.0 : dst[(size_t)k * nEg]) + blk[k];
            }
        }
        __syncthreads(); // the next chunk overwrites the scratch
        if (cyc) { cyc[2] += clock64() - tg0; tg0 = clock64(); }
    }
    // (3b) indenter vertices against the gel's contact triangles: one triangle per thread writes its 45 numbers to the (now free)
    // scratch, rows and edges gather them through static incidence lists in a fixed order (deterministic, like the tets)
    if (MESH && ind.type == 2 && a.n_ctri > 0) {
        if (threadIdx.x < a.n_ctri) {
            TpOut o;
            tp_terms(a, ind, s.x, threadIdx.x, a.kappa * dt2, 1, &o);
#pragma unroll 5
            for (int k = 0; k < 45; ++k) tsc[(size_t)k * FEM_THREADS + threadIdx.x] = o.v[k];
        }
        if (threadIdx.x < a.n_cedge) { // edge-edge candidates of this thread's gel edge: rows 45..68 of the scratch
            EeOut o;
            ee_terms(a, ind, s.x, threadIdx.x, a.kappa * dt2, 1, &o);
#pragma unroll 4
            for (int k = 0; k < 24; ++k) tsc[(size_t)(45 + k) * FEM_THREADS + threadIdx.x] = o.v[k];
        }
        __syncthreads();
        if (on) {
            for (int q = a.ctri_row_start[i]; q < a.ctri_row_start[i + 1]; ++q) {
                const int ent = a.ctri_row_adj[q], f = ent >> 2, j = ent & 3;
                const double* tf = tsc + f;
                g3[0] += tf[(size_t)(3 * j) * FEM_THREADS]; g3[1] += tf[(size_t)(3 * j + 1) * FEM_THREADS]; g3[2] += tf[(size_t)(3 * j + 2) * FEM_THREADS];
#pragma unroll
                for (int k = 0; k < 6; ++k) d6[k] += tf[(size_t)(9 + 6 * j + k) * FEM_THREADS];
            }
            for (int q = a.cedge_row_start[i]; q < a.cedge_row_start[i + 1]; ++q) {
                const int ent = a.cedge_row_adj[q], ce = ent >> 1, j = ent & 1;
                const double* tf = tsc + (size_t)45 * FEM_THREADS + ce;
                g3[0] += tf[(size_t)(3 * j) * FEM_THREADS]; g3[1] += tf[(size_t)(3 * j + 1) * FEM_THREADS]; g3[2] += tf[(size_t)(3 * j + 2) * FEM_THREADS];
#pragma unroll
                for (int k = 0; k < 6; ++k) d6[k] += tf[(size_t)(6 + 6 * j + k) * FEM_THREADS];
            }
        }
        for (int e = threadIdx.x; e < a.nE; e += FEM_THREADS) {
            const int q0 = a.ctri_edge_start[e], q1 = a.ctri_edge_start[e + 1];
            const int ce = a.n_cedge > 0 ? a.edge_cedge[e] : -1;
            if (q0 == q1 && ce < 0) continue;
            double b6[6] = {0, 0, 0, 0, 0, 0};
            if (ce >= 0) {
#pragma unroll
                for (int k = 0; k < 6; ++k) b6[k] = tsc[(size_t)(45 + 18 + k) * FEM_THREADS + ce];
            }
            for (int q = q0; q < q1; ++q) {
                const int ent = a.ctri_edge_adj[q], f = ent >> 2, pr = ent & 3;
#pragma unroll
                for (int k = 0; k < 6; ++k) b6[k] += tsc[(size_t)(27 + 6 * pr + k) * FEM_THREADS + f];
            }
            const double blk[9] = {b6[0], b6[1], b6[2], b6[1], b6[3], b6[4], b6[2], b6[4], b6[5]};
            if (e < a.n_s) {
                double* dst = s.val + e;
#pragma unroll
                for (int k = 0; k < 9; ++k) dst[k * FEM_VAL_STRIDE] += blk[k];
            } else {
                double* dst = valg + (e - a.n_s);
                const int nEg = a.nE - a.n_s;
#pragma unroll
                for (int k = 0; k < 9; ++k) dst[(size_t)k * nEg] += blk[k];
            }
        }
        __syncthreads();
    }
    // (4) attachment + barrier (row-local)
    if (on) {
        const int ka = a.attach_of[i];
        if (ka >= 0) {
            const double sm = a.attach_strength * m;
            for (int c = 0; c < 3; ++c) {
                const double xp = xprev_g[3 * i + c];
                const double aimx = xp + (aim_g[3 * ka + c] - xp) * ratio;
                g3[c] += sm * (xi[c] - aimx);
            }
            d6[0] += sm; d6[3] += sm; d6[5] += sm;
        }
        if (a.surf_of[i] >= 0) {
            double d, n[3], Hd[9], dB, ddB, Hk[9];
            if (MESH && ind.type == 2) {
                mesh_contact(a, ind, xi, a.kappa * dt2, nullptr, nullptr, nullptr, g3, d6);
                d = -1.0;
            } else
                indenter_sdf<MESH>(a, ind, xi, &d, n, Hd);
            if ((d * d < a.d_hat * a.d_hat) && d > 0.0) {
                barrier_fn(d * d, a.d_hat, a.kappa * dt2, nullptr, &dB, &ddB);
                double dD[3];
                for (int c = 0; c < 3; ++c) dD[c] = 2.0 * d * n[c];
                for (int c = 0; c < 3; ++c) g3[c] += dB * dD[c];
                for (int c = 0; c < 3; ++c)
                    for (int b = 0; b < 3; ++b) Hk[3 * c + b] = ddB * dD[c] * dD[b] + dB * 2.0 * (n[c] * n[b] + d * Hd[3 * c + b]);
                spd_project<3>(Hk);
                d6[0] += Hk[0]; d6[1] += Hk[1]; d6[2] += Hk[2]; d6[3] += Hk[4]; d6[4] += Hk[5]; d6[5] += Hk[8];
            }
            if (a.friction_mu > 0.0) {
                const double xp[3] = {xprev_g[3 * i], xprev_g[3 * i + 1], xprev_g[3 * i + 2]};
                friction_terms<MESH>(a, ind0, ind, xp, xi, nullptr, g3, d6, lag);
            }
        }
    }
    if (cyc) cyc[1] += clock64() - tg0;
}

// y = A p for the row of this thread: diagonal block + the off-diagonal blocks in ELL order (symmetric storage: the block of
// edge (i, j), i < j, serves row i as is and row j transposed; both reads are coalesced when the edges are numbered by
// (j - i, i), which tx_fem_create does). The ELL entries of the row are fetched up front (one L2 round trip per product).
constexpr int FEM_SLOT_GROUP = 16;
__device__ __forceinline__ void spmv_row(const FemArgs& a, const FemShared& s, const double* valg, const double d6[6],
                                         double y[3])
{
    const int i = threadIdx.x;
    y[0] = y[1] = y[2] = 0.0;
    if (i >= a.V) return;
    sym3_mul(d6, s.p[3 * i], s.p[3 * i + 1], s.p[3 * i + 2], y[0], y[1], y[2]);
    const int nEg = a.nE - a.n_s;
    for (int sl0 = 0; sl0 < a.nslots; sl0 += FEM_SLOT_GROUP) {
        int pks[FEM_SLOT_GROUP];
#pragma unroll
        for (int u = 0; u < FEM_SLOT_GROUP; ++u)
            pks[u] = sl0 + u < a.nslots ? __ldg(a.ell + (size_t)(sl0 + u) * FEM_THREADS + i) : -1;
#pragma unroll
        for (int u = 0; u < FEM_SLOT_GROUP; ++u) {
            const int pk = pks[u];
            if (pk < 0) continue;
            const int j = pk & 0xfff, e = pk >> 13;
            const double p0 = s.p[3 * j], p1 = s.p[3 * j + 1], p2 = s.p[3 * j + 2];
            double m[9];
            if (e < a.n_s) {
#pragma unroll
                for (int k = 0; k < 9; ++k) m[k] = s.val[k * FEM_VAL_STRIDE + e];
            } else {
#pragma unroll
                for (int k = 0; k < 9; ++k) m[k] = valg[(size_t)k * nEg + (e - a.n_s)];
            }
            if (pk & 0x1000) { // this row is the j of edge (i', j): A^T
                y[0] += m[0] * p0 + m[3] * p1 + m[6] * p2;
                y[1] += m[1] * p0 + m[4] * p1 + m[7] * p2;
                y[2] += m[2] * p0 + m[5] * p1 + m[8] * p2;
            } else {
                y[0] += m[0] * p0 + m[1] * p1 + m[2] * p2;
                y[1] += m[3] * p0 + m[4] * p1 + m[5] * p2;
                y[2] += m[6] * p0 + m[7] * p1 + m[8] * p2;
            }
        }
    }
}

// PCG with the 3x3 block-Jacobi preconditioner (linear_pcg.cu:45-140, fem_diag_preconditioner.cu:112-164), x0 = 0,
// b = -g3; solution in dx (row-local). Every vector except p is row-local (registers).
__device__ int pcg(const FemArgs& a, const FemShared& s, const double* valg, const double g3[3], const double d6[6],
                   double dx[3], int& ph, long long* cyc)
{
    long long tp0 = 0;
    const int i = threadIdx.x;
    const bool on = i < a.V;
    double inv[6] = {0, 0, 0, 0, 0, 0};
    if (on) { // inverse of the symmetric diagonal block
        const double m0 = d6[0], m1 = d6[1], m2 = d6[2], m4 = d6[3], m5 = d6[4], m8 = d6[5];
        const double c00 = m4 * m8 - m5 * m5, c01 = m2 * m5 - m1 * m8, c02 = m1 * m5 - m2 * m4;
        const double det = m0 * c00 + m1 * c01 + m2 * c02;
        inv[0] = c00 / det; inv[1] = c01 / det; inv[2] = c02 / det;
        inv[3] = (m0 * m8 - m2 * m2) / det; inv[4] = (m1 * m2 - m0 * m5) / det; inv[5] = (m0 * m4 - m1 * m1) / det;
    }
    double r[3] = {-g3[0], -g3[1], -g3[2]}, z[3], pl[3], Ap[3];
    dx[0] = dx[1] = dx[2] = 0.0;
    sym3_mul(inv, r[0], r[1], r[2], z[0], z[1], z[2]);
    for (int c = 0; c < 3; ++c) pl[c] = z[c];
    if (on) { s.p[3 * i] = pl[0]; s.p[3 * i + 1] = pl[1]; s.p[3 * i + 2] = pl[2]; }
    double rz = block_reduce<0>(on ? r[0] * z[0] + r[1] * z[1] + r[2] * z[2] : 0.0, s.red, ph); // also publishes p
    const double rz0 = fabs(rz);
    if (rz0 == 0.0) return 0;
    int k;
    const int max_iter = a.pcg_max_iter_ratio * 3 * a.V;
    for (k = 1; k < max_iter; ++k) {
        if (cyc) tp0 = clock64();
        spmv_row(a, s, valg, d6, Ap);
        if (cyc) { cyc[0] += clock64() - tp0; tp0 = clock64(); }
        const double pAp = block_reduce<0>(on ? pl[0] * Ap[0] + pl[1] * Ap[1] + pl[2] * Ap[2] : 0.0, s.red, ph);
        const double alpha = rz / pAp;
        for (int c = 0; c < 3; ++c) { dx[c] += alpha * pl[c]; r[c] -= alpha * Ap[c]; }
        sym3_mul(inv, r[0], r[1], r[2], z[0], z[1], z[2]);
        const double rzn = block_reduce<0>(on ? r[0] * z[0] + r[1] * z[1] + r[2] * z[2] : 0.0, s.red, ph);
        if (fabs(rzn) <= a.pcg_tol_rate * rz0) break;
        const double beta = rzn / rz;
        for (int c = 0; c < 3; ++c) pl[c] = z[c] + beta * pl[c];
        // every thread is past its spmv_row (two reductions ago): p may be overwritten; the next reduction is one barrier
        // too late for the next spmv_row, hence the explicit one
        if (on) { s.p[3 * i] = pl[0]; s.p[3 * i + 1] = pl[1]; s.p[3 * i + 2] = pl[2]; }
        __syncthreads();
        rz = rzn;
        if (cyc) cyc[1] += clock64() - tp0;
    }
    return k;
}

__device__ __forceinline__ FemIndenter lerp_ind(const FemIndenter& a, const FemIndenter& b, double t)
{
    FemIndenter o = b;
    for (int i = 0; i < 3; ++i) o.c[i] = a.c[i] + (b.c[i] - a.c[i]) * t;
    return o;
}

// MESH = false is the kernel of the analytic indenters alone (the out-of-line mesh call would cost it registers and spills)
template <bool MESH>
__global__ void __launch_bounds__(FEM_THREADS, 1) fem_step_kernel(const FemArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = 3 * a.V;
    double* base = reinterpret_cast<double*>(smem_raw);
    FemShared s;
    s.x = base; s.p = s.x + n; s.val = s.p + n;
    s.red = reinterpret_cast<Red*>(s.val + (size_t)9 * FEM_VAL_STRIDE);
    double* tsc = a.tet_scratch + (size_t)blockIdx.x * FEM_THREADS * 4 * FEM_UNITS;
    double* valg = a.val_scratch + (size_t)blockIdx.x * 9 * (a.nE - a.n_s);
    double* xt_g = a.xt_scratch + (size_t)blockIdx.x * n;
    const int i = threadIdx.x;
    const bool on = i < a.V;
    int ph = 0;
    long long cyc[6] = {0, 0, 0, 0, 0, 0}, tc0 = 0;
#define FEM_TIC() do { if (a.dbg_cycles) tc0 = clock64(); } while (0)
#define FEM_TOC(k) do { if (a.dbg_cycles) cyc[k] += clock64() - tc0; } while (0)

    for (int env = blockIdx.x; env < a.N; env += gridDim.x) {
        double* xg = a.x + (size_t)env * n;
        double* vg = a.v + (size_t)env * n;
        double* xpg = a.x_prev + (size_t)env * n;
        const double* aim_g = a.aim + (size_t)env * 3 * a.A;
        const FemIndenter ind_prev = a.ind_prev[env], ind_next = a.ind_next[env];
        // predict (fem_bdf1_time_integrator.cu:19-55): every gel vertex is dynamic and not fixed
        for (int k = threadIdx.x; k < n; k += FEM_THREADS) {
            s.x[k] = xg[k];
            xt_g[k] = xpg[k] + a.gravity[k % 3] * a.dt * a.dt + vg[k] * a.dt;
        }
        __syncthreads();
        const double abs_tol = a.velocity_tol * a.dt;
        double res0 = 0.0, ccd_alpha = 1.0, ind_s = 0.0, last_res = 0.0, min_dist = 0.0, energy = 0.0;
        double umax = 0.0;
        for (int c = 0; c < 3; ++c) umax += (ind_next.c[c] - ind_prev.c[c]) * (ind_next.c[c] - ind_prev.c[c]);
        umax = sqrt(umax);
        const bool is_surf = on && a.surf_of[i] >= 0;
        FrLag lag{0.0, {0.0, 0.0, 1.0}};
        double Glag[3] = {0.0, 0.0, 0.0};
        if (MESH && ind_prev.type == 2 && a.friction_mu > 0.0 && a.n_ctri > 0) {
            // lagged normal forces of the multi-vertex candidate families at the start-of-step state: the triangle / edge threads
            // evaluate their candidates at x_prev against ind_prev, the rows gather their share through the static lists
            const double kdt2 = a.kappa * a.dt * a.dt;
            if (threadIdx.x < a.n_ctri) {
                TpOut o;
                tp_terms(a, ind_prev, xpg, threadIdx.x, kdt2, 1, &o);
#pragma unroll
                for (int k = 0; k < 9; ++k) tsc[(size_t)k * FEM_THREADS + threadIdx.x] = o.v[k];
            }
            if (threadIdx.x < a.n_cedge) {
                EeOut o;
                ee_terms(a, ind_prev, xpg, threadIdx.x, kdt2, 1, &o);
#pragma unroll
                for (int k = 0; k < 6; ++k) tsc[(size_t)(45 + k) * FEM_THREADS + threadIdx.x] = o.v[k];
            }
            __syncthreads();
            if (on) {
                for (int q = a.ctri_row_start[i]; q < a.ctri_row_start[i + 1]; ++q) {
                    const int ent = a.ctri_row_adj[q], f = ent >> 2, j = ent & 3;
#pragma unroll
                    for (int c = 0; c < 3; ++c) Glag[c] += tsc[(size_t)(3 * j + c) * FEM_THREADS + f];
                }
                for (int q = a.cedge_row_start[i]; q < a.cedge_row_start[i + 1]; ++q) {
                    const int ent = a.cedge_row_adj[q], ce = ent >> 1, j = ent & 1;
#pragma unroll
                    for (int c = 0; c < 3; ++c) Glag[c] += tsc[(size_t)(45 + 3 * j + c) * FEM_THREADS + ce];
                }
            }
            __syncthreads(); // the assembly reuses the scratch
        }
        if (MESH && ind_prev.type == 2 && is_surf && a.friction_mu > 0.0) {
            const double xp3[3] = {xpg[3 * i], xpg[3 * i + 1], xpg[3 * i + 2]};
            lag = mesh_friction_lag(a, ind_prev, xp3, Glag);
        }
        int it, pcg_total = 0, ls_total = 0, conv = 0;
        for (it = 0; it < a.newton_max_iter; ++it) {
            const double tt = ((double)it + 1.0) / (double)a.substep;
            const double ratio = tt < 1.0 ? tt : 1.0;
            if (ind_s < 1.0) { // advance the prescribed indenter by at most half of the current minimum gap
                const FemIndenter cur = lerp_ind(ind_prev, ind_next, ind_s);
                double md = 1e300;
                if (is_surf) {
                    double d, nn[3];
                    indenter_sdf<MESH>(a, cur, s.x + 3 * i, &d, nn, nullptr);
                    md = d;
                }
                if (MESH && cur.type == 2 && a.n_ctri > 0 && threadIdx.x < a.n_ctri) {
                    TpOut o;
                    tp_terms(a, cur, s.x, threadIdx.x, 0.0, 0, &o);
                    md = fmin(md, sqrt(o.dmin2));
                }
                if (MESH && cur.type == 2 && a.n_cedge > 0 && threadIdx.x < a.n_cedge) {
                    EeOut o;
                    ee_terms(a, cur, s.x, threadIdx.x, 0.0, 0, &o);
                    md = fmin(md, sqrt(o.dmin2));
                }
                md = block_reduce<1>(md, s.red, ph);
                double ds = umax > 0.0 ? 0.5 * md / umax : 1.0;
                if (ds < 0.0) ds = 0.0;
                ind_s = ind_s + ds < 1.0 ? ind_s + ds : 1.0;
            }
            const FemIndenter ind = lerp_ind(ind_prev, ind_next, ind_s);

            double g3[3], d6[6], dx[3];
            FEM_TIC();
            grad_hess<MESH>(a, s, xt_g, xpg, aim_g, ind, ind_prev, ratio, tsc, valg, g3, d6, (a.dbg_cycles && !a.dbg_mode) ? cyc + 3 : nullptr, lag);
            FEM_TOC(0);
            FEM_TIC();
            pcg_total += pcg(a, s, valg, g3, d6, dx, ph, (a.dbg_cycles && a.dbg_mode) ? cyc + 3 : nullptr);
            FEM_TOC(1);
            FEM_TIC();

            double res = on ? fmax(fmax(fabs(dx[0]), fabs(dx[1])), fabs(dx[2])) : 0.0;
            res = block_reduce<2>(res, s.red, ph);
            if (it == 0) res0 = res;
            const double rel = res == 0.0 ? 0.0 : res / res0;
            const bool converged = (res <= abs_tol) || (rel <= 0.001);
            last_res = res;
            if (it > 0 && converged && ccd_alpha >= 1.0 && ratio >= 1.0 && ind_s >= 1.0) { conv = 1; break; }

            // line search (sim_engine_do_advance.cu:276-347); x0 is row-local
            double x0[3] = {0, 0, 0};
            if (on) { x0[0] = s.x[3 * i]; x0[1] = s.x[3 * i + 1]; x0[2] = s.x[3 * i + 2]; }
            double alpha = 1.0;
            if (MESH && ind.type == 2) {
                if (is_surf) alpha = mesh_ccd(a, ind, x0, dx); // the reference's ACCD per (vertex, triangle) pair
                if (a.n_ctri > 0) { // moving gel triangles against the indenter's vertices: the step of the other rows through s.p
                    __syncthreads();
                    if (on) { s.p[3 * i] = dx[0]; s.p[3 * i + 1] = dx[1]; s.p[3 * i + 2] = dx[2]; }
                    __syncthreads();
                    if (threadIdx.x < a.n_ctri) {
                        const MeshVerts mv{a.mesh_vert, a.mesh_nv, a.d_hat, a.grid_vert};
                        const int f = threadIdx.x, i0 = a.ctri[3 * f], i1 = a.ctri[3 * f + 1], i2 = a.ctri[3 * f + 2];
                        const double at = tp_ccd_impl(
                            mv, make_double3(ind.c[0], ind.c[1], ind.c[2]), make_double3(ind.R[0], ind.R[1], ind.R[2]),
                            make_double3(ind.R[3], ind.R[4], ind.R[5]), make_double3(ind.R[6], ind.R[7], ind.R[8]),
                            make_double3(s.x[3 * i0], s.x[3 * i0 + 1], s.x[3 * i0 + 2]), make_double3(s.x[3 * i1], s.x[3 * i1 + 1], s.x[3 * i1 + 2]),
                            make_double3(s.x[3 * i2], s.x[3 * i2 + 1], s.x[3 * i2 + 2]), make_double3(s.p[3 * i0], s.p[3 * i0 + 1], s.p[3 * i0 + 2]),
                            make_double3(s.p[3 * i1], s.p[3 * i1 + 1], s.p[3 * i1 + 2]), make_double3(s.p[3 * i2], s.p[3 * i2 + 1], s.p[3 * i2 + 2]));
                        alpha = fmin(alpha, at);
                    }
                    if (threadIdx.x < a.n_cedge) {
                        const MeshEdges me{a.mesh_vert, a.mesh_edge, a.mesh_ne, a.d_hat, a.grid_edge};
                        const int i0 = a.cedge[2 * threadIdx.x], i1 = a.cedge[2 * threadIdx.x + 1];
                        const double ae = ee_ccd_impl(
                            me, make_double3(ind.c[0], ind.c[1], ind.c[2]), make_double3(ind.R[0], ind.R[1], ind.R[2]),
                            make_double3(ind.R[3], ind.R[4], ind.R[5]), make_double3(ind.R[6], ind.R[7], ind.R[8]),
                            make_double3(s.x[3 * i0], s.x[3 * i0 + 1], s.x[3 * i0 + 2]), make_double3(s.x[3 * i1], s.x[3 * i1 + 1], s.x[3 * i1 + 2]),
                            make_double3(s.p[3 * i0], s.p[3 * i0 + 1], s.p[3 * i0 + 2]), make_double3(s.p[3 * i1], s.p[3 * i1 + 1], s.p[3 * i1 + 2]));
                        alpha = fmin(alpha, ae);
                    }
                }
            } else if (is_surf) {
                double d, nn[3];
                indenter_sdf<MESH>(a, ind, x0, &d, nn, nullptr);
                const double len = sqrt(dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2]);
                if (len > 0.0 && d < 2.0 * len + a.d_hat) alpha = fmin(alpha, 0.8 * d / len);
            }
            alpha = block_reduce<1>(alpha, s.red, ph);
            ccd_alpha = alpha;
            const double E0 = total_energy<MESH>(a, s, xt_g, xpg, aim_g, ind, ind_prev, ratio, nullptr, ph, lag);
            __syncthreads(); // every thread has read s.x
            if (on) for (int c = 0; c < 3; ++c) s.x[3 * i + c] = x0[c] + alpha * dx[c];
            __syncthreads();
            double E = total_energy<MESH>(a, s, xt_g, xpg, aim_g, ind, ind_prev, ratio, &min_dist, ph, lag);
            if (!converged) {
                int ls = 0;
                while (ls < a.ls_max_iter) {
                    if (E <= E0) break;
                    alpha *= 0.5;
                    __syncthreads();
                    if (on) for (int c = 0; c < 3; ++c) s.x[3 * i + c] = x0[c] + alpha * dx[c];
                    __syncthreads();
                    E = total_energy<MESH>(a, s, xt_g, xpg, aim_g, ind, ind_prev, ratio, &min_dist, ph, lag);
                    ++ls;
                    ++ls_total;
                }
            }
            energy = E;
            __syncthreads();
            FEM_TOC(2);
        }
        // update velocity (fem_bdf1_time_integrator.cu:58-77), write back
        __syncthreads();
        for (int k = threadIdx.x; k < n; k += FEM_THREADS) {
            const double xn = s.x[k];
            vg[k] = (xn - xpg[k]) * (1.0 / a.dt);
            xpg[k] = xn;
            xg[k] = xn;
        }
        if (threadIdx.x == 0 && a.dbg_cycles)
            for (int k = 0; k < 6; ++k) a.dbg_cycles[(size_t)blockIdx.x * 6 + k] = cyc[k];
        if (threadIdx.x == 0 && a.stats) {
            FemStats st;
            st.converged = conv; st.newton_iters = it; st.pcg_iters = pcg_total; st.ls_halvings = ls_total;
            st.min_dist = min_dist; st.last_res = last_res; st.energy = energy;
            a.stats[env] = st;
        }
        __syncthreads();
    }
}

int fem_threads() { return FEM_THREADS; }

// n_s = number of off-diagonal blocks kept in shared memory
size_t fem_smem_bytes(int V, int n_s) { (void)n_s; return sizeof(double) * (2 * 3 * (size_t)V + 9 * (size_t)FEM_VAL_STRIDE) + sizeof(Red) + 64; }

int fem_max_smem_edges(int V) { (void)V; return FEM_VAL_STRIDE; }

cudaError_t launch_fem_step(const FemArgs& a, int grid, cudaStream_t st)
{
    const size_t smem = fem_smem_bytes(a.V, a.n_s);
    if (a.mesh_n > 0) { // a triangle-mesh indenter is set (tx_fem_set_indenter_mesh): the kernel that knows indenter type 2
        cudaError_t e = cudaFuncSetAttribute(fem_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        fem_step_kernel<true><<<grid, FEM_THREADS, smem, st>>>(a);
    } else {
        cudaError_t e = cudaFuncSetAttribute(fem_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        fem_step_kernel<false><<<grid, FEM_THREADS, smem, st>>>(a);
    }
    return cudaGetLastError();
}

// ---- FEM marker read-out: barycentric interpolation on the gel's top surface + pinhole projection ----------------------
// ref: source/tacex/tacex/simulation_approaches/fem_based/sim/tactile_sensor_sapienipc_modified.py:354-413 (gen_marker_flow;
// the weights are precomputed once here, the reference recomputes them every step and only for env 0)
__global__ void fem_marker_kernel(const FemMarkerArgs m)
{
    const int env = blockIdx.x;
    for (int k = threadIdx.x; k < m.M; k += blockDim.x) {
        const int* tri = m.tri + 3 * k;
        const double* w = m.weights + 3 * k;
        float* out = m.out + ((size_t)env * 2 * m.M + k) * 2;
        for (int which = 0; which < 2; ++which) {
            const double* X = which == 0 ? m.x_rest : m.x + (size_t)env * 3 * m.V;
            double pw[3] = {0, 0, 0};
            for (int c = 0; c < 3; ++c) pw[c] = w[0] * X[3 * tri[0] + c] + w[1] * X[3 * tri[1] + c] + w[2] * X[3 * tri[2] + c];
            // world -> camera frame: p_cam = Rc^T (p - tc)
            double pc[3];
            for (int c = 0; c < 3; ++c)
                pc[c] = m.cam_R[0 * 3 + c] * (pw[0] - m.cam_t[0]) + m.cam_R[1 * 3 + c] * (pw[1] - m.cam_t[1]) + m.cam_R[2 * 3 + c] * (pw[2] - m.cam_t[2]);
            const double u = m.fx * pc[0] / pc[2] + m.cx, vv = m.fy * pc[1] / pc[2] + m.cy;
            float* o = out + (size_t)which * m.M * 2;
            // the reference projects in float32 (Isaac Lab project_points) and post-processes in float64 (NumPy)
            float uf = (float)u, vf = (float)vv;
            if (m.normalize) {
                uf = (float)((double)uf / m.half_w - 1.0);
                vf = (float)((double)vf / m.half_w - 1.0);
            }
            o[0] = m.zero_all ? 0.0f : uf;
            o[1] = m.zero_all ? 0.0f : vf;
        }
    }
}

cudaError_t launch_fem_markers(const FemMarkerArgs& m, int N, cudaStream_t st)
{
    fem_marker_kernel<<<N, 128, 0, st>>>(m);
    return cudaGetLastError();
}

} // namespace tx
