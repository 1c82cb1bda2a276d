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
    void setConstant(T s) { for (int i = 0; i < N; ++i) (*m)(i, i) = s; }
    struct Unused {
    };
    Arr array() { return Arr{m}; }
};

template <class T, int R, int C>
struct ColRef { // assignable column of a matrix (m.col(j) = v)
    Matrix<T, R, C>* m;
    int j;
    ColRef& operator=(const Matrix<T, R, 1>& v) { for (int i = 0; i < R; ++i) (*m)(i, j) = v.d[i]; return *this; }
};


// ---- assignable views used by the reference's distance headers (utils/distance/distance_flagged.h) -------------------------------
template <class T, int R, int C, int N>
struct SegRef { // v.segment<N>(i) / v.head<N>() of a vector
    Matrix<T, R, C>* m;
    int i0;
    SegRef& operator=(const SegRef& o) { for (int i = 0; i < N; ++i) m->d[i0 + i] = o.m->d[o.i0 + i]; return *this; }
    template <int R2, int C2>
    SegRef& operator=(const SegRef<T, R2, C2, N>& o) { for (int i = 0; i < N; ++i) m->d[i0 + i] = o.m->d[o.i0 + i]; return *this; }
    SegRef& operator=(const Matrix<T, N, 1>& v) { for (int i = 0; i < N; ++i) m->d[i0 + i] = v.d[i]; return *this; }
    Matrix<T, N, 1> operator-() const { Matrix<T, N, 1> v; for (int i = 0; i < N; ++i) v.d[i] = -m->d[i0 + i]; return v; }
    operator Matrix<T, N, 1>() const { Matrix<T, N, 1> v; for (int i = 0; i < N; ++i) v.d[i] = m->d[i0 + i]; return v; }
};
template <class T, int R, int C, int BR, int BC>
struct BlockRef { // m.block<BR, BC>(i, j)
    Matrix<T, R, C>* m;
    int i0, j0;
    template <int R2, int C2>
    BlockRef& operator=(const BlockRef<T, R2, C2, BR, BC>& o)
    {
        for (int i = 0; i < BR; ++i) for (int j = 0; j < BC; ++j) (*m)(i0 + i, j0 + j) = (*o.m)(o.i0 + i, o.j0 + j);
        return *this;
    }
    BlockRef& operator=(const BlockRef& o)
    {
        for (int i = 0; i < BR; ++i) for (int j = 0; j < BC; ++j) (*m)(i0 + i, j0 + j) = (*o.m)(o.i0 + i, o.j0 + j);
        return *this;
    }
};
template <class T, int R, int C>
struct RowRef { // m.row(i): assignable from a row or a column vector, cross product with a 3-vector
    Matrix<T, R, C>* m;
    int i;
    RowRef& operator=(const Matrix<T, 1, C>& v) { for (int j = 0; j < C; ++j) (*m)(i, j) = v.d[j]; return *this; }
    RowRef& operator=(const Matrix<T, C, 1>& v) { for (int j = 0; j < C; ++j) (*m)(i, j) = v.d[j]; return *this; }
    operator Matrix<T, C, 1>() const { Matrix<T, C, 1> v; for (int j = 0; j < C; ++j) v.d[j] = (*m)(i, j); return v; }
    Matrix<T, C, 1> cross(const Matrix<T, C, 1>& o) const { return Matrix<T, C, 1>(*this).cross(o); }
    Matrix<T, C, 1> cross(const RowRef& o) const { return Matrix<T, C, 1>(*this).cross(Matrix<T, C, 1>(o)); }
};

template <class T, int R, int C>
struct Matrix {
    T d[R * C];
    Matrix() { for (int i = 0; i < R * C; ++i) d[i] = T(0); }
    Matrix(T a, T b) { static_assert(R * C == 2, "size"); d[0] = a; d[1] = b; }
    Matrix(T a, T b, T c) { static_assert(R * C == 3, "size"); d[0] = a; d[1] = b; d[2] = c; }
    Matrix(T a, T b, T c, T e) { static_assert(R * C == 4, "size"); d[0] = a; d[1] = b; d[2] = c; d[3] = e; }
    void setZero() { for (int i = 0; i < R * C; ++i) d[i] = T(0); }
    void setConstant(T v) { for (int i = 0; i < R * C; ++i) d[i] = v; }
    template <int N> SegRef<T, R, C, N> segment(int i0) { return SegRef<T, R, C, N>{this, i0}; }
    template <int N> SegRef<T, R, C, N> head() { return SegRef<T, R, C, N>{this, 0}; }
    template <int BR, int BC> BlockRef<T, R, C, BR, BC> block(int i0, int j0) { return BlockRef<T, R, C, BR, BC>{this, i0, j0}; }
    RowRef<T, R, C> row(int i) { return RowRef<T, R, C>{this, i}; }
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
    Matrix operator/(T s) const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = d[i] / s; return m; }
    Matrix operator-(const Matrix& o) const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = d[i] - o.d[i]; return m; }
    Matrix operator-() const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = -d[i]; return m; }
    Matrix operator*(T s) const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = d[i] * s; return m; }
    Matrix& operator*=(T s) { for (int i = 0; i < R * C; ++i) d[i] *= s; return *this; }
    Matrix& operator+=(const Matrix& o) { for (int i = 0; i < R * C; ++i) d[i] += o.d[i]; return *this; }
    Matrix& operator-=(const Matrix& o) { for (int i = 0; i < R * C; ++i) d[i] -= o.d[i]; return *this; }
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

// the little of Eigen::Array that utils/distance/details/ccd.inl touches (coefficient list + maxCoeff)
template <class T, int R, int C>
struct Array {
    T d[R * C];
    Array(T a, T b, T c) : d{a, b, c} { static_assert(R * C == 3, "size"); }
    Array(T a, T b, T c, T e) : d{a, b, c, e} { static_assert(R * C == 4, "size"); }
    T maxCoeff() const { T m = d[0]; for (int i = 1; i < R * C; ++i) m = d[i] > m ? d[i] : m; return m; }
    T minCoeff() const { T m = d[0]; for (int i = 1; i < R * C; ++i) m = d[i] < m ? d[i] : m; return m; }
};
template <class T> using Array3 = Array<T, 3, 1>;
template <class T> using Array4 = Array<T, 4, 1>;

} // namespace Eigen
