// Stand-in for muda/ext/eigen/inverse.h (TEST INFRASTRUCTURE): inverse of a 3 x 3 / 2 x 2 matrix by the adjugate.
#pragma once
#include "../../../mini_eigen.h"
namespace muda::eigen {
template <class T>
inline Eigen::Matrix<T, 3, 3> inverse(const Eigen::Matrix<T, 3, 3>& a)
{
    Eigen::Matrix<T, 3, 3> r;
    const T det = a.determinant();
    r(0, 0) = (a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) / det;
    r(0, 1) = (a(0, 2) * a(2, 1) - a(0, 1) * a(2, 2)) / det;
    r(0, 2) = (a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1)) / det;
    r(1, 0) = (a(1, 2) * a(2, 0) - a(1, 0) * a(2, 2)) / det;
    r(1, 1) = (a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0)) / det;
    r(1, 2) = (a(0, 2) * a(1, 0) - a(0, 0) * a(1, 2)) / det;
    r(2, 0) = (a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0)) / det;
    r(2, 1) = (a(0, 1) * a(2, 0) - a(0, 0) * a(2, 1)) / det;
    r(2, 2) = (a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0)) / det;
    return r;
}
template <class T>
inline Eigen::Matrix<T, 2, 2> inverse(const Eigen::Matrix<T, 2, 2>& a)
{
    Eigen::Matrix<T, 2, 2> r;
    const T det = a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0);
    r(0, 0) = a(1, 1) / det;
    r(0, 1) = -a(0, 1) / det;
    r(1, 0) = -a(1, 0) / det;
    r(1, 1) = a(0, 0) / det;
    return r;
}
} // namespace muda::eigen
