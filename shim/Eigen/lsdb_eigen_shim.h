// Minimal fixed-size stand-in for the Eigen3 features the LSD-SLAM host code
// uses (this image has no Eigen): Matrix<double,R,C> with comma initialiser
// (scalars and blocks), element access, + - *, scalar*, transpose, asDiagonal,
// llt().matrixL(), inverse(), Zero().  Eager evaluation, no expression templates.
#ifndef LSDB_EIGEN_SHIM_H
#define LSDB_EIGEN_SHIM_H
#include <math.h>
#include <assert.h>

namespace Eigen {

template <typename T, int R, int C> class Matrix;

template <int N> struct DiagonalWrapperShim { double d[N]; };

template <int R, int C> class CommaInitShim {
public:
    Matrix<double, R, C>& m;
    int row, col, blockRows;
    CommaInitShim(Matrix<double, R, C>& mm) : m(mm), row(0), col(0), blockRows(1) {}
    void advance(int r, int c) {
        col += c;
        blockRows = r;
        if (col >= C) { col = 0; row += blockRows; }
    }
    CommaInitShim& put(double v) { m(row, col) = v; advance(1, 1); return *this; }
    template <int R2, int C2> CommaInitShim& put(const Matrix<double, R2, C2>& b) {
        for (int i = 0; i < R2; i++) for (int j = 0; j < C2; j++) m(row + i, col + j) = b(i, j);
        advance(R2, C2);
        return *this;
    }
    CommaInitShim& operator,(double v) { return put(v); }
    template <int R2, int C2> CommaInitShim& operator,(const Matrix<double, R2, C2>& b) { return put(b); }
};

template <int N> class LLTShim {
public:
    Matrix<double, N, N> L;
    Matrix<double, N, N> matrixL() const { return L; }
};

template <typename T, int R, int C> class Matrix {
public:
    T v[R * C];  // row-major
    Matrix() { for (int i = 0; i < R * C; i++) v[i] = 0; }
    static Matrix Zero() { return Matrix(); }
    T& operator()(int i, int j) { return v[i * C + j]; }
    const T& operator()(int i, int j) const { return v[i * C + j]; }
    T& operator()(int i) { return v[i]; }
    const T& operator()(int i) const { return v[i]; }

    CommaInitShim<R, C> operator<<(double x) { CommaInitShim<R, C> ci(*this); ci.put(x); return ci; }
    template <int R2, int C2> CommaInitShim<R, C> operator<<(const Matrix<double, R2, C2>& b) {
        CommaInitShim<R, C> ci(*this); ci.put(b); return ci;
    }

    Matrix operator+(const Matrix& o) const { Matrix r; for (int i = 0; i < R * C; i++) r.v[i] = v[i] + o.v[i]; return r; }
    Matrix operator-(const Matrix& o) const { Matrix r; for (int i = 0; i < R * C; i++) r.v[i] = v[i] - o.v[i]; return r; }
    template <int C2> Matrix<T, R, C2> operator*(const Matrix<T, C, C2>& o) const {
        Matrix<T, R, C2> r;
        for (int i = 0; i < R; i++) for (int j = 0; j < C2; j++) {
            T s = 0;
            for (int k = 0; k < C; k++) s += (*this)(i, k) * o(k, j);
            r(i, j) = s;
        }
        return r;
    }
    Matrix operator*(const DiagonalWrapperShim<C>& d) const {
        Matrix r;
        for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) r(i, j) = (*this)(i, j) * d.d[j];
        return r;
    }
    Matrix operator*(double s) const { Matrix r; for (int i = 0; i < R * C; i++) r.v[i] = v[i] * s; return r; }
    Matrix<T, C, R> transpose() const {
        Matrix<T, C, R> r;
        for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) r(j, i) = (*this)(i, j);
        return r;
    }
    DiagonalWrapperShim<R * C> asDiagonal() const {
        DiagonalWrapperShim<R * C> d;
        for (int i = 0; i < R * C; i++) d.d[i] = v[i];
        return d;
    }
    LLTShim<R> llt() const {
        LLTShim<R> f;
        for (int j = 0; j < R; j++) {
            double s = (*this)(j, j);
            for (int k = 0; k < j; k++) s -= f.L(j, k) * f.L(j, k);
            double d = sqrt(s);
            f.L(j, j) = d;
            for (int i = j + 1; i < R; i++) {
                double t = (*this)(i, j);
                for (int k = 0; k < j; k++) t -= f.L(i, k) * f.L(j, k);
                f.L(i, j) = t / d;
            }
        }
        return f;
    }
    Matrix inverse() const {  // Gauss-Jordan with partial pivoting (R == C)
        Matrix a = *this, inv;
        for (int i = 0; i < R; i++) inv(i, i) = 1;
        for (int c = 0; c < R; c++) {
            int p = c;
            for (int r = c + 1; r < R; r++) if (fabs(a(r, c)) > fabs(a(p, c))) p = r;
            if (p != c) for (int j = 0; j < R; j++) { T t = a(c, j); a(c, j) = a(p, j); a(p, j) = t; t = inv(c, j); inv(c, j) = inv(p, j); inv(p, j) = t; }
            T d = a(c, c);
            for (int j = 0; j < R; j++) { a(c, j) /= d; inv(c, j) /= d; }
            for (int r = 0; r < R; r++) if (r != c) {
                T f = a(r, c);
                for (int j = 0; j < R; j++) { a(r, j) -= f * a(c, j); inv(r, j) -= f * inv(c, j); }
            }
        }
        return inv;
    }
};

template <typename T, int R, int C> Matrix<T, R, C> operator*(double s, const Matrix<T, R, C>& m) { return m * s; }

}  // namespace Eigen
#endif
