// Stand-in for muda/ext/eigen/evd.h (TEST INFRASTRUCTURE): eigen-decomposition of a symmetric 2 x 2 matrix, all the compiled
// reference headers need (make_spd of the 2 x 2 friction Hessian).
#pragma once
#include "../../../mini_eigen.h"
namespace muda::eigen {
template <class T>
inline void evd(const Eigen::Matrix<T, 2, 2>& A, Eigen::Matrix<T, 2, 1>& w, Eigen::Matrix<T, 2, 2>& V)
{
    const T a = A(0, 0), b = (A(0, 1) + A(1, 0)) * T(0.5), c = A(1, 1);
    const T th = T(0.5) * std::atan2(T(2) * b, a - c), cs = std::cos(th), sn = std::sin(th);
    V(0, 0) = cs; V(1, 0) = sn; V(0, 1) = -sn; V(1, 1) = cs;
    w(0) = cs * cs * a + T(2) * cs * sn * b + sn * sn * c;
    w(1) = sn * sn * a - T(2) * cs * sn * b + cs * cs * c;
}
} // namespace muda::eigen
