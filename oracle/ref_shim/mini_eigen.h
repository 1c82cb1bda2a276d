// mini_eigen.h -- TEST INFRASTRUCTURE: the small fixed-size subset of the Eigen API that the reference's self-contained contact /
// constitution headers use (element access, products, transpose, dot / cross / norms, Identity / UnitX / UnitZ, diagonal().array(),
// asDiagonal()). Eigen itself is not in this image; the reference headers are compiled UNMODIFIED from /root/reference against
// this stand-in (oracle/Makefile, target _ref/libuipc_sym.so). Value semantics, column-major storage, no expression templates.
#pragma once
#include <cmath>
#include <cstring>

namespace Eigen {

template <class T, int R, int C>
struct Matrix;

template <class T, int N>
struct DiagRef {
    Matrix<T, N, N>* m;
    struct Arr {
        Matrix<T, N, N>* m;
        Arr& operator+=(T s) { for (int i = 0; i < N; ++i) (*m)(i, i) += s; return *this; }
    };
    Arr array() { return Arr{m}; }
};

template <class T, int R, int C>
struct ColRef { // assignable column of a matrix (m.col(j) = v)
    Matrix<T, R, C>* m;
    int j;
    ColRef& operator=(const Matrix<T, R, 1>& v) { for (int i = 0; i < R; ++i) (*m)(i, j) = v.d[i]; return *this; }
};

template <class T, int R, int C>
struct Matrix {
    T d[R * C];
    Matrix() { for (int i = 0; i < R * C; ++i) d[i] = T(0); }
    Matrix(T a, T b) { static_assert(R * C == 2, "size"); d[0] = a; d[1] = b; }
    Matrix(T a, T b, T c) { static_assert(R * C == 3, "size"); d[0] = a; d[1] = b; d[2] = c; }
    T* data() { return d; }
    const T* data() const { return d; }
    T& operator()(int i, int j) { return d[j * R + i]; }
    const T& operator()(int i, int j) const { return d[j * R + i]; }
    T& operator()(int i) { return d[i]; }
    const T& operator()(int i) const { return d[i]; }
    T& operator[](int i) { return d[i]; }
    const T& operator[](int i) const { return d[i]; }
    static Matrix Zero() { return Matrix(); }
    static Matrix Identity() { Matrix m; for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = T(1); return m; }
    static Matrix UnitX() { Matrix m; m.d[0] = T(1); return m; }
    static Matrix UnitY() { Matrix m; m.d[1] = T(1); return m; }
    static Matrix UnitZ() { Matrix m; m.d[2] = T(1); return m; }
    Matrix<T, C, R> transpose() const { Matrix<T, C, R> t; for (int i = 0; i < R; ++i) for (int j = 0; j < C; ++j) t(j, i) = (*this)(i, j); return t; }
    T dot(const Matrix& o) const { T s = T(0); for (int i = 0; i < R * C; ++i) s += d[i] * o.d[i]; return s; }
    T squaredNorm() const { return dot(*this); }
    T norm() const { return std::sqrt(squaredNorm()); }
    Matrix normalized() const { Matrix m = *this; const T n = norm(); for (int i = 0; i < R * C; ++i) m.d[i] /= n; return m; }
    Matrix cross(const Matrix& o) const
    {
        static_assert(R * C == 3, "cross");
        return Matrix(d[1] * o.d[2] - d[2] * o.d[1], d[2] * o.d[0] - d[0] * o.d[2], d[0] * o.d[1] - d[1] * o.d[0]);
    }
    Matrix operator+(const Matrix& o) const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = d[i] + o.d[i]; return m; }
    Matrix operator-(const Matrix& o) const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = d[i] - o.d[i]; return m; }
    Matrix operator-() const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = -d[i]; return m; }
    Matrix operator*(T s) const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = d[i] * s; return m; }
    Matrix& operator*=(T s) { for (int i = 0; i < R * C; ++i) d[i] *= s; return *this; }
    Matrix& operator+=(const Matrix& o) { for (int i = 0; i < R * C; ++i) d[i] += o.d[i]; return *this; }
    template <int K>
    Matrix<T, R, K> operator*(const Matrix<T, C, K>& o) const
    {
        Matrix<T, R, K> m;
        for (int i = 0; i < R; ++i)
            for (int j = 0; j < K; ++j) {
                T s = T(0);
                for (int k = 0; k < C; ++k) s += (*this)(i, k) * o(k, j);
                m(i, j) = s;
            }
        return m;
    }
    ColRef<T, R, C> col(int j) { return ColRef<T, R, C>{this, j}; }
    Matrix<T, R, 1> col(int j) const { Matrix<T, R, 1> v; for (int i = 0; i < R; ++i) v.d[i] = (*this)(i, j); return v; }
    T determinant() const
    {
        static_assert(R == 3 && C == 3, "3 x 3");
        const Matrix& a = *this;
        return a(0, 0) * (a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) - a(0, 1) * (a(1, 0) * a(2, 2) - a(1, 2) * a(2, 0)) +
               a(0, 2) * (a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0));
    }
    DiagRef<T, R> diagonal() { static_assert(R == C, "square"); return DiagRef<T, R>{this}; }
    Matrix<T, R, R> asDiagonal() const { static_assert(C == 1, "vector"); Matrix<T, R, R> m; for (int i = 0; i < R; ++i) m(i, i) = d[i]; return m; }
};

template <class T, int R, int C>
inline Matrix<T, R, C> operator*(T s, const Matrix<T, R, C>& m) { return m * s; }

template <class T, int N>
using Vector = Matrix<T, N, 1>;

} // namespace Eigen
