// ref_dist.cpp -- TEST INFRASTRUCTURE: C entry points around the reference's closest-feature distance headers, compiled UNMODIFIED
// from where they lie under /root/reference (libuipc cuda backend, host-compilable: generated scalar code + small fixed-size algebra):
//   utils/distance/distance_flagged.h        point_triangle_distance_flag / point_edge_distance_flag / edge_edge_distance_flag and the
//                                            flagged squared distances, gradients (12) and Hessians (12 x 12)
//   utils/distance/{point_point,point_edge,point_triangle,edge_edge}.h + details/*.inl
//   utils/distance/ccd.h + details/ccd.inl   point_triangle_ccd (additive CCD)
// They are what the reference's narrow phase and barrier call for every point-triangle / edge-edge candidate
// (collision_detection/filters/lbvh_simplex_trajectory_filter.cu:600-690, contact_system/contact_models/ipc_simplex_normal_contact.cu:270-342).
// Eigen / muda are not in this image: oracle/ref_shim/ supplies the minimal stand-in. Built by oracle/Makefile into
// oracle/_ref/libuipc_dist.so (only where /root/reference exists); used by tests/test_fem_ref_pin_cpu.py to pin the closest-feature
// classification and the distance derivatives of oracle/fem_canon.c (fem_pt_distance, fem_ee_distance).
#include <type_define.h>
#include <muda/muda_def.h>
// The reference writes `G.segment<3>(i)` on a dependent type without the `template` disambiguator (nvcc's front end accepts that, g++
// does not). oracle/Makefile therefore pipes distance_flagged.h through ONE sed expression that inserts the keyword into a temporary
// file outside the repo (deleted after the compile) -- every other reference file is read where it lies.
#include REF_DISTANCE_FLAGGED_H
#include <algorithm>
#include <utils/distance/ccd.h> // + details/ccd.inl (ACCD); instantiated below: point_triangle_ccd, edge_edge_ccd
#include <utils/distance/edge_edge_mollifier.h>

using namespace uipc;
namespace D = uipc::backend::cuda::distance;
static Vector3 v3(const double* p) { return Vector3(p[0], p[1], p[2]); }

extern "C" {

// point p vs triangle (t0, t1, t2): flag[4] (1 = the vertex takes part in the closest feature), squared distance, gradient w.r.t.
// (p, t0, t1, t2) and Hessian (row-major 12 x 12)
void ref_pt(const double* p, const double* t0, const double* t1, const double* t2, int* flag, double* D2, double* G, double* H)
{
    const Vector4i F = D::point_triangle_distance_flag(v3(p), v3(t0), v3(t1), v3(t2));
    for (int i = 0; i < 4; ++i) flag[i] = F[i];
    D::point_triangle_distance2(F, v3(p), v3(t0), v3(t1), v3(t2), *D2);
    Vector12 g;
    Matrix12x12 h;
    D::point_triangle_distance2_gradient(F, v3(p), v3(t0), v3(t1), v3(t2), g);
    D::point_triangle_distance2_hessian(F, v3(p), v3(t0), v3(t1), v3(t2), h);
    for (int i = 0; i < 12; ++i) {
        G[i] = g(i);
        for (int j = 0; j < 12; ++j) H[12 * i + j] = h(i, j);
    }
}

void ref_pe(const double* p, const double* e0, const double* e1, int* flag, double* D2, double* G, double* H)
{
    const Vector3i F = D::point_edge_distance_flag(v3(p), v3(e0), v3(e1));
    for (int i = 0; i < 3; ++i) flag[i] = F[i];
    D::point_edge_distance2(F, v3(p), v3(e0), v3(e1), *D2);
    Eigen::Vector<double, 9> g;
    Eigen::Matrix<double, 9, 9> h;
    D::point_edge_distance2_gradient(F, v3(p), v3(e0), v3(e1), g);
    D::point_edge_distance2_hessian(F, v3(p), v3(e0), v3(e1), h);
    for (int i = 0; i < 9; ++i) {
        G[i] = g(i);
        for (int j = 0; j < 9; ++j) H[9 * i + j] = h(i, j);
    }
}

void ref_ee(const double* a0, const double* a1, const double* b0, const double* b1, int* flag, double* D2, double* G, double* H)
{
    const Vector4i F = D::edge_edge_distance_flag(v3(a0), v3(a1), v3(b0), v3(b1));
    for (int i = 0; i < 4; ++i) flag[i] = F[i];
    D::edge_edge_distance2(F, v3(a0), v3(a1), v3(b0), v3(b1), *D2);
    Vector12 g;
    Matrix12x12 h;
    D::edge_edge_distance2_gradient(F, v3(a0), v3(a1), v3(b0), v3(b1), g);
    D::edge_edge_distance2_hessian(F, v3(a0), v3(a1), v3(b0), v3(b1), h);
    for (int i = 0; i < 12; ++i) {
        G[i] = g(i);
        for (int j = 0; j < 12; ++j) H[12 * i + j] = h(i, j);
    }
}

// additive CCD of a point against a triangle, all four moving (positions + displacements over the step); returns hit, *toc in/out
int ref_pt_ccd(const double* p, const double* t0, const double* t1, const double* t2, const double* dp, const double* dt0,
               const double* dt1, const double* dt2, double eta, double thickness, int max_iter, double* toc)
{
    return D::point_triangle_ccd(v3(p), v3(t0), v3(t1), v3(t2), v3(dp), v3(dt0), v3(dt1), v3(dt2), eta, thickness, max_iter, *toc) ? 1 : 0;
}

int ref_ee_ccd(const double* a0, const double* a1, const double* b0, const double* b1, const double* da0, const double* da1,
               const double* db0, const double* db1, double eta, double thickness, int max_iter, double* toc)
{
    return D::edge_edge_ccd(v3(a0), v3(a1), v3(b0), v3(b1), v3(da0), v3(da1), v3(db0), v3(db1), eta, thickness, max_iter, *toc) ? 1 : 0;
}

// mollifier of nearly parallel edges: threshold from the rest edges, value and 12-gradient
void ref_ee_mollifier(const double* ra0, const double* ra1, const double* rb0, const double* rb1, const double* a0, const double* a1,
                      const double* b0, const double* b1, double* eps_x, double* e, double* g)
{
    D::edge_edge_mollifier_threshold(v3(ra0), v3(ra1), v3(rb0), v3(rb1), *eps_x);
    D::edge_edge_mollifier(v3(a0), v3(a1), v3(b0), v3(b1), *eps_x, *e);
    Vector12 gg;
    D::edge_edge_mollifier_gradient(v3(a0), v3(a1), v3(b0), v3(b1), *eps_x, gg);
    for (int i = 0; i < 12; ++i) g[i] = gg(i);
}

} // extern "C"
