// ref_sym.cpp -- TEST INFRASTRUCTURE: C entry points around reference source files compiled UNMODIFIED from where they lie under
// /root/reference (libuipc, cuda backend; the files are host-compilable: generated scalar code + small fixed-size algebra):
//   finite_element/constitutions/sym/stable_neo_hookean_3d.inl          E, dEdVecF, ddEddVecF of the stable Neo-Hookean energy
//   contact_system/contact_models/codim_ipc_contact_function.h          KappaBarrier (+ derivatives, sym/codim_ipc_contact.inl),
//                                                                       C1-clamped friction f0 / f1 / f2, 2 x 2 Hessian, normal_force
//   contact_system/contact_models/ipc_vertex_half_plane_contact_function.h   PH_barrier_* / PH_friction_* (vertex vs half-plane,
//                                                                       sym/vertex_half_plane_distance.inl)
//   finite_element/fem_utils.cu                                         Ds, Dm_inv, F = Ds Dm^-1, dFdx (9 x 12) of a tetrahedron
//   (libuipc/include) uipc/constitution/conversion.h                    EP_to_lame, what ElasticModuli::youngs_poisson calls
//                                                                       (src/constitution/elastic_moduli.cpp:20-27)
//   (libuipc/external/muda) muda/ext/eigen/inverse/analytic_inverse.h   the 3 x 3 inverse of the block-Jacobi preconditioner
//                                                                       (finite_element/fem_diag_preconditioner.cu:142-148 calls
//                                                                       muda::eigen::inverse -> AnalyticalInverse for 3 x 3)
// Eigen / muda are not in this image: oracle/ref_shim/ supplies a minimal stand-in (type_define.h, mini_eigen.h, a 2 x 2 evd) and
// empty headers for the includes the compiled subset does not use. Built by oracle/Makefile into oracle/_ref/libuipc_sym.so (only
// where /root/reference exists); used by tests/test_fem_ref_pin_cpu.py to pin oracle/fem_canon.c.
#include <type_define.h>
#include <contact_system/contact_models/ipc_vertex_half_plane_contact_function.h>
#include <uipc/constitution/conversion.h> // frontend: Young's modulus / Poisson ratio -> Lame parameters (header-only)

#include <finite_element/fem_utils.cu> // Ds, Dm_inv, F, dFdx (9 x 12): plain functions, host-compilable
#include <cstdlib>
#include <muda/ext/eigen/inverse/analytic_inverse.h> // the real muda file (found through -I$(MUDA_SRC)), host-compilable
namespace uipc::backend::cuda { // mentioned by fem_utils.cu's invariant helpers, never called by the pin tests
Float ddot(const Matrix3x3&, const Matrix3x3&) { std::abort(); }
void svd(const Matrix3x3&, Matrix3x3&, Vector3&, Matrix3x3&) noexcept { std::abort(); }
} // namespace uipc::backend::cuda

namespace ref_snh_ns {
using namespace uipc;
#include <finite_element/constitutions/sym/stable_neo_hookean_3d.inl>
} // namespace ref_snh_ns

using namespace uipc;
namespace PH = uipc::backend::cuda::sym::ipc_vertex_half_contact;
namespace CI = uipc::backend::cuda::sym::codim_ipc_contact;

static Vector3 v3(const double* p) { return Vector3(p[0], p[1], p[2]); }

extern "C" {

// F as column-major vec (VecF(3*b + a) = F(a, b)); H row-major 9 x 9
void ref_snh(const double* F, double mu, double lambda, double* E, double* g, double* H)
{
    Eigen::Vector<double, 9> vf, dg;
    Eigen::Matrix<double, 9, 9> dh;
    for (int i = 0; i < 9; ++i) vf(i) = F[i];
    ref_snh_ns::E(*E, mu, lambda, vf);
    ref_snh_ns::dEdVecF(dg, mu, lambda, vf);
    ref_snh_ns::ddEddVecF(dh, mu, lambda, vf);
    for (int i = 0; i < 9; ++i) {
        g[i] = dg(i);
        for (int j = 0; j < 9; ++j) H[9 * i + j] = dh(i, j);
    }
}

void ref_kappa_barrier(double kappa, double D, double d_hat, double xi, double* B, double* dB, double* ddB)
{
    CI::KappaBarrier(*B, kappa, D, d_hat, xi);
    CI::dKappaBarrierdD(*dB, kappa, D, d_hat, xi);
    CI::ddKappaBarrierddD(*ddB, kappa, D, d_hat, xi);
}

// vertex v against the half-plane (P, N); H row-major 3 x 3
void ref_ph_barrier(double kappa, double d_hat, double thickness, const double* v, const double* P, const double* N, double* E,
                    double* G, double* H)
{
    *E = PH::PH_barrier_energy(kappa, d_hat, thickness, v3(v), v3(P), v3(N));
    Vector3 g;
    Matrix3x3 h;
    PH::PH_barrier_gradient_hessian(g, h, kappa, d_hat, thickness, v3(v), v3(P), v3(N));
    for (int i = 0; i < 3; ++i) {
        G[i] = g(i);
        for (int j = 0; j < 3; ++j) H[3 * i + j] = h(i, j);
    }
}

void ref_ph_friction(double kappa, double d_hat, double thickness, double mu, double eps_vh, const double* prev_v, const double* v,
                     const double* P, const double* N, double* E, double* G, double* H)
{
    *E = PH::PH_friction_energy(kappa, d_hat, thickness, mu, eps_vh, v3(prev_v), v3(v), v3(P), v3(N));
    // PH_friction_gradient_hessian itself is declared Float but flows off its end without a return value: undefined behaviour
    // that g++ turns into a trap in host code. Its body (ipc_vertex_half_plane_contact_function.h:110-145) is therefore
    // composed here from the very same reference functions: normal_force, compute_tan_basis, TR, friction_gradient, dTRdx,
    // friction_hessian; only the two products G = J^T G2 and H = J^T H2 J are this file's.
    Vector3 g;
    Matrix3x3 h;
    {
        Float prev_D;
        PH::HalfPlaneD(prev_D, v3(prev_v), v3(P), v3(N));
        const Float f = CI::normal_force(kappa, d_hat, thickness, prev_D);
        Vector3 e1, e2;
        PH::compute_tan_basis(e1, e2, v3(N));
        Vector2 tan_dV;
        PH::TR(tan_dV, v3(v), v3(prev_v), e1, e2);
        Vector2 G2;
        CI::friction_gradient(G2, mu, f, eps_vh, tan_dV);
        Matrix<Float, 2, 3> J;
        PH::dTRdx(J, v3(v), v3(prev_v), e1, e2);
        g = J.transpose() * G2;
        Matrix2x2 H2;
        CI::friction_hessian(H2, mu, f, eps_vh, tan_dV);
        h = J.transpose() * H2 * J;
    }
    for (int i = 0; i < 3; ++i) {
        G[i] = g(i);
        for (int j = 0; j < 3; ++j) H[3 * i + j] = h(i, j);
    }
}

void ref_ep_to_lame(double E, double nu, double* lambda, double* mu) { uipc::constitution::EP_to_lame(E, nu, *lambda, *mu); }

// one tetrahedron: X / x = 4 rest / current vertices (12 doubles each); Dm^-1 row-major 3 x 3, F as column-major vec (the VecF
// of the constitution), dFdx row-major 9 x 12 (rows: VecF, columns: the 12 vertex coordinates)
void ref_tet(const double* X, const double* x, double* DmInv, double* vecF, double* dFdx)
{
    namespace fem = uipc::backend::cuda::fem;
    const Matrix3x3 Di = fem::Dm_inv(v3(X), v3(X + 3), v3(X + 6), v3(X + 9));
    const Matrix3x3 F = fem::F(v3(x), v3(x + 3), v3(x + 6), v3(x + 9), Di);
    const Matrix9x12 P = fem::dFdx(Di);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            DmInv[3 * i + j] = Di(i, j);
            vecF[3 * j + i] = F(i, j);
        }
    for (int i = 0; i < 9; ++i)
        for (int j = 0; j < 12; ++j) dFdx[12 * i + j] = P(i, j);
}

void ref_tan_basis(const double* N, double* e1, double* e2)
{
    Vector3 a, b;
    PH::compute_tan_basis(a, b, v3(N));
    for (int i = 0; i < 3; ++i) { e1[i] = a(i); e2[i] = b(i); }
}

} // extern "C"

extern "C" void ref_inverse3(const double* m_colmajor, double* out_colmajor)
{
    Eigen::Matrix<double, 3, 3> a;
    for (int i = 0; i < 9; ++i) a.data()[i] = m_colmajor[i];
    const Eigen::Matrix<double, 3, 3> r = muda::eigen::AnalyticalInverse{}(a);
    for (int i = 0; i < 9; ++i) out_colmajor[i] = r.data()[i];
}
