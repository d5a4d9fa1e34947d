// oracle/standin/l3d_standin_eigen.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A stand-in for the small part of Eigen 3 that the reference's Line3D++ sources touch
// (src/line3D.cc, src/view.cc, include/view.h, include/segment3D.h: Vector2d/3d/4d/4f, Matrix3d,
// Matrix<double,3,4>, MatrixXd, VectorXd, JacobiSVD of a symmetric 3x3, AngleAxisd), so that those
// sources compile UNMODIFIED, from where they lie under /root/reference, into oracle/_ref
// (oracle/Makefile, target ref_line3d).  Eigen itself is not installed in this image.
//
// Everything is evaluated eagerly in the canonical order of SURVEY.md Appendix A: left-to-right sums,
// row-times-column products, true divisions, normalized(v) = v / sqrt(v.v), adjugate/determinant inverse of
// a 3x3.  The real Eigen may differ from this in the last bit of some results (its version is not pinned
// by the reference either); the control flow, thresholds, containers and list handling that run on top
// of it are the reference's own.
#pragma once
#include <cassert>
#include <cmath>
#include <cstddef>
#include <iostream>
#include <type_traits>
#include <vector>

namespace Eigen {

const int Dynamic = -1;
enum { ComputeThinU = 1, ComputeThinV = 2, ComputeFullU = 4, ComputeFullV = 8 };

template <typename T, int R, int C>
class Matrix;

namespace detail {
template <typename T, int R, int C>
struct Store {
    T d[R * C];
    int rows() const { return R; }
    int cols() const { return C; }
    void resize(int, int) {}
};
template <typename T, int C>
struct Store<T, Dynamic, C> {
    std::vector<T> d;
    int r = 0;
    int rows() const { return r; }
    int cols() const { return C; }
    void resize(int rr, int) { r = rr; d.assign((size_t)rr * C, T(0)); }
};
template <typename T>
struct Store<T, Dynamic, Dynamic> {
    std::vector<T> d;
    int r = 0, c = 0;
    int rows() const { return r; }
    int cols() const { return c; }
    void resize(int rr, int cc) { r = rr; c = cc; d.assign((size_t)rr * cc, T(0)); }
};
}  // namespace detail

// comma initialiser: M << a, b, c, ...   (row-major fill)
template <typename M>
struct CommaInit {
    M& m;
    int k;
    CommaInit(M& mm, typename M::Scalar v) : m(mm), k(0) { put(v); }
    void put(typename M::Scalar v)
    {
        const int c = m.cols();
        m(k / c, k % c) = v;
        ++k;
    }
    CommaInit& operator,(typename M::Scalar v)
    {
        put(v);
        return *this;
    }
};

template <typename T, int R, int C>
class Matrix {
  public:
    typedef T Scalar;
    detail::Store<T, R, C> s;

    Matrix()
    {
        if (R != Dynamic && C != Dynamic)
            for (int i = 0; i < R * C; ++i) s.d[i] = T(0);
    }
    Matrix(int rows, int cols) { s.resize(rows, cols); zero(); }
    explicit Matrix(int n) { s.resize(n, 1); zero(); }
    Matrix(T a, T b) { s.resize(2, 1); s.d[0] = a; s.d[1] = b; }
    Matrix(T a, T b, T c) { s.resize(3, 1); s.d[0] = a; s.d[1] = b; s.d[2] = c; }
    Matrix(T a, T b, T c, T d) { s.resize(4, 1); s.d[0] = a; s.d[1] = b; s.d[2] = c; s.d[3] = d; }
    template <int R2, int C2>
    Matrix(const Matrix<T, R2, C2>& o)
    {
        s.resize(o.rows(), o.cols());
        assert(rows() == o.rows() && cols() == o.cols());
        for (int i = 0; i < rows(); ++i)
            for (int j = 0; j < cols(); ++j) (*this)(i, j) = o(i, j);
    }
    template <int R2, int C2>
    Matrix& operator=(const Matrix<T, R2, C2>& o)
    {
        s.resize(o.rows(), o.cols());
        assert(rows() == o.rows() && cols() == o.cols());
        for (int i = 0; i < rows(); ++i)
            for (int j = 0; j < cols(); ++j) (*this)(i, j) = o(i, j);
        return *this;
    }

    int rows() const { return s.rows(); }
    int cols() const { return s.cols(); }
    int size() const { return rows() * cols(); }
    void resize(int r, int c) { s.resize(r, c); }
    void zero()
    {
        for (int i = 0; i < size(); ++i) s.d[i] = T(0);
    }
    void setZero() { zero(); }
    // row-major storage (the layout is private to this stand-in)
    T& operator()(int i, int j) { return s.d[(size_t)i * cols() + j]; }
    const T& operator()(int i, int j) const { return s.d[(size_t)i * cols() + j]; }
    T& operator()(int i) { return s.d[i]; }
    const T& operator()(int i) const { return s.d[i]; }
    T& operator[](int i) { return s.d[i]; }
    const T& operator[](int i) const { return s.d[i]; }
    T& x() { return s.d[0]; }
    T& y() { return s.d[1]; }
    T& z() { return s.d[2]; }
    T& w() { return s.d[3]; }
    const T& x() const { return s.d[0]; }
    const T& y() const { return s.d[1]; }
    const T& z() const { return s.d[2]; }
    const T& w() const { return s.d[3]; }

    CommaInit<Matrix> operator<<(T v) { return CommaInit<Matrix>(*this, v); }

    static Matrix Zero() { return Matrix(); }
    static Matrix Zero(int r, int c) { return Matrix(r, c); }
    static Matrix Identity()
    {
        Matrix m;
        for (int i = 0; i < m.rows() && i < m.cols(); ++i) m(i, i) = T(1);
        return m;
    }
    static Matrix Identity(int r, int c)
    {
        Matrix m(r, c);
        for (int i = 0; i < r && i < c; ++i) m(i, i) = T(1);
        return m;
    }
    static Matrix Constant(int r, int c, T v)
    {
        Matrix m(r, c);
        for (int i = 0; i < m.size(); ++i) m.s.d[i] = v;
        return m;
    }

    // ---- vector operations (canonical order) ----
    template <int R2, int C2>
    T dot(const Matrix<T, R2, C2>& o) const
    {
        T acc = s.d[0] * o.s.d[0];
        for (int i = 1; i < size(); ++i) acc = acc + s.d[i] * o.s.d[i];
        return acc;
    }
    T squaredNorm() const { return dot(*this); }
    T norm() const { return std::sqrt(dot(*this)); }
    Matrix normalized() const
    {
        const T n = norm();
        Matrix m(*this);
        for (int i = 0; i < size(); ++i) m.s.d[i] = s.d[i] / n;
        return m;
    }
    void normalize() { *this = normalized(); }
    Matrix cross(const Matrix& b) const
    {
        const Matrix& a = *this;
        return Matrix(a.s.d[1] * b.s.d[2] - a.s.d[2] * b.s.d[1], a.s.d[2] * b.s.d[0] - a.s.d[0] * b.s.d[2],
                      a.s.d[0] * b.s.d[1] - a.s.d[1] * b.s.d[0]);
    }
    template <typename I>
    T maxCoeff(I* index) const
    {
        int at = 0;
        for (int i = 1; i < size(); ++i)
            if (s.d[i] > s.d[at]) at = i;
        *index = (I)at;
        return s.d[at];
    }
    static Matrix UnitX() { Matrix m; m.s.d[0] = T(1); return m; }
    static Matrix UnitY() { Matrix m; m.s.d[1] = T(1); return m; }
    static Matrix UnitZ() { Matrix m; m.s.d[2] = T(1); return m; }
    void transposeInPlace() { *this = Matrix(transpose()); }
    T maxCoeff() const
    {
        T m = s.d[0];
        for (int i = 1; i < size(); ++i)
            if (s.d[i] > m) m = s.d[i];
        return m;
    }
    T determinant() const
    {
        const Matrix& A = *this;
        return A(0, 0) * (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1)) - A(0, 1) * (A(1, 0) * A(2, 2) - A(1, 2) * A(2, 0)) +
               A(0, 2) * (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0));
    }

    Matrix<T, C, R> transpose() const
    {
        Matrix<T, C, R> t;
        t.resize(cols(), rows());
        for (int i = 0; i < rows(); ++i)
            for (int j = 0; j < cols(); ++j) t(j, i) = (*this)(i, j);
        return t;
    }
    // 3x3: adjugate / determinant, the closed form Eigen uses for fixed sizes up to 4
    Matrix inverse() const
    {
        assert(rows() == 3 && cols() == 3);
        const Matrix& A = *this;
        auto cof = [&](int i, int j) {
            return A((i + 1) % 3, (j + 1) % 3) * A((i + 2) % 3, (j + 2) % 3) -
                   A((i + 1) % 3, (j + 2) % 3) * A((i + 2) % 3, (j + 1) % 3);
        };
        const T det = cof(0, 0) * A(0, 0) + cof(1, 0) * A(1, 0) + cof(2, 0) * A(2, 0);
        const T invdet = T(1) / det;
        Matrix Rm;
        Rm.resize(3, 3);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Rm(i, j) = cof(j, i) * invdet;
        return Rm;
    }

    // ---- views returned by value; writable block access through a proxy ----
    struct BlockRef {
        Matrix& m;
        int r0, c0, nr, nc;
        template <int R2, int C2>
        BlockRef& operator=(const Matrix<T, R2, C2>& o)
        {
            for (int i = 0; i < nr; ++i)
                for (int j = 0; j < nc; ++j) m(r0 + i, c0 + j) = o(i, j);
            return *this;
        }
        BlockRef& operator=(const BlockRef& o)
        {
            for (int i = 0; i < nr; ++i)
                for (int j = 0; j < nc; ++j) m(r0 + i, c0 + j) = o.m(o.r0 + i, o.c0 + j);
            return *this;
        }
        operator Matrix<T, Dynamic, Dynamic>() const { return eval(); }
        Matrix<T, Dynamic, Dynamic> eval() const
        {
            Matrix<T, Dynamic, Dynamic> o(nr, nc);
            for (int i = 0; i < nr; ++i)
                for (int j = 0; j < nc; ++j) o(i, j) = m(r0 + i, c0 + j);
            return o;
        }
        template <int R2, int C2>
        T dot(const Matrix<T, R2, C2>& o) const
        {
            return eval().dot(o);
        }
        T operator()(int i, int j) const { return m(r0 + i, c0 + j); }
        T operator()(int i) const { return nr == 1 ? m(r0, c0 + i) : m(r0 + i, c0); }
        T x() const { return (*this)(0); }
        T y() const { return (*this)(1); }
        T z() const { return (*this)(2); }
        Matrix<T, Dynamic, Dynamic> transpose() const { return eval().transpose(); }
        T norm() const { return eval().norm(); }
    };
    BlockRef block(int r0, int c0, int nr, int nc) { return BlockRef{*this, r0, c0, nr, nc}; }
    BlockRef row(int i) { return BlockRef{*this, i, 0, 1, cols()}; }
    BlockRef col(int j) { return BlockRef{*this, 0, j, rows(), 1}; }
    template <int NR, int NC>
    BlockRef block(int r0, int c0) { return BlockRef{*this, r0, c0, NR, NC}; }
    Matrix<T, Dynamic, Dynamic> block(int r0, int c0, int nr, int nc) const
    {
        return BlockRef{const_cast<Matrix&>(*this), r0, c0, nr, nc}.eval();
    }
    template <int NR, int NC>
    Matrix<T, NR, NC> block(int r0, int c0) const
    {
        Matrix<T, NR, NC> o;
        for (int i = 0; i < NR; ++i)
            for (int j = 0; j < NC; ++j) o(i, j) = (*this)(r0 + i, c0 + j);
        return o;
    }
    Matrix<T, 1, C> row(int i) const
    {
        Matrix<T, 1, C> o;
        o.resize(1, cols());
        for (int j = 0; j < cols(); ++j) o(0, j) = (*this)(i, j);
        return o;
    }
    Matrix<T, R, 1> col(int j) const
    {
        Matrix<T, R, 1> o;
        o.resize(rows(), 1);
        for (int i = 0; i < rows(); ++i) o(i, 0) = (*this)(i, j);
        return o;
    }

    // ---- arithmetic ----
    Matrix operator-() const
    {
        Matrix m(*this);
        for (int i = 0; i < size(); ++i) m.s.d[i] = -s.d[i];
        return m;
    }
    Matrix& operator+=(const Matrix& o)
    {
        for (int i = 0; i < size(); ++i) s.d[i] = s.d[i] + o.s.d[i];
        return *this;
    }
    Matrix& operator-=(const Matrix& o)
    {
        for (int i = 0; i < size(); ++i) s.d[i] = s.d[i] - o.s.d[i];
        return *this;
    }
    Matrix& operator*=(T v)
    {
        for (int i = 0; i < size(); ++i) s.d[i] = s.d[i] * v;
        return *this;
    }
    Matrix& operator/=(T v)
    {
        for (int i = 0; i < size(); ++i) s.d[i] = s.d[i] / v;
        return *this;
    }
};

template <typename T, int R, int C>
Matrix<T, R, C> operator+(const Matrix<T, R, C>& a, const Matrix<T, R, C>& b)
{
    Matrix<T, R, C> m(a);
    m += b;
    return m;
}
template <typename T, int R, int C>
Matrix<T, R, C> operator-(const Matrix<T, R, C>& a, const Matrix<T, R, C>& b)
{
    Matrix<T, R, C> m(a);
    m -= b;
    return m;
}
// scalar operands of another arithmetic type are converted to the matrix scalar first, as Eigen's
// operator*(const Scalar&) does
template <typename T, int R, int C, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
Matrix<T, R, C> operator*(const Matrix<T, R, C>& a, S v)
{
    Matrix<T, R, C> m(a);
    m *= (T)v;
    return m;
}
template <typename T, int R, int C, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
Matrix<T, R, C> operator*(S v, const Matrix<T, R, C>& a)
{
    Matrix<T, R, C> m(a);
    for (int i = 0; i < m.size(); ++i) m.s.d[i] = (T)v * a.s.d[i];
    return m;
}
template <typename T, int R, int C, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
Matrix<T, R, C> operator/(const Matrix<T, R, C>& a, S v)
{
    Matrix<T, R, C> m(a);
    m /= (T)v;
    return m;
}
// row-times-column products, terms added left to right
template <typename T, int R, int K, int C>
Matrix<T, R, C> operator*(const Matrix<T, R, K>& a, const Matrix<T, K, C>& b)
{
    Matrix<T, R, C> m;
    m.resize(a.rows(), b.cols());
    assert(a.cols() == b.rows());
    for (int i = 0; i < a.rows(); ++i)
        for (int j = 0; j < b.cols(); ++j) {
            T acc = a(i, 0) * b(0, j);
            for (int k = 1; k < a.cols(); ++k) acc = acc + a(i, k) * b(k, j);
            m(i, j) = acc;
        }
    return m;
}
template <typename T, int R, int C>
std::ostream& operator<<(std::ostream& os, const Matrix<T, R, C>& m)
{
    for (int i = 0; i < m.rows(); ++i) {
        for (int j = 0; j < m.cols(); ++j) os << (j ? " " : "") << m(i, j);
        if (i + 1 < m.rows()) os << "\n";
    }
    return os;
}

typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<float, 2, 1> Vector2f;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 4, 1> Vector4f;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<double, Dynamic, 1> VectorXd;

// Singular value decomposition of a SYMMETRIC positive semi-definite 3x3 (the only use on the path:
// the scatter matrix of Line3D::get3DlineFromCluster, src/line3D.cc:2619-2633): cyclic Jacobi rotations,
// U = eigenvectors sorted by descending eigenvalue.  The sign of a column of U is arbitrary, as with the
// real JacobiSVD; the caller uses U.col(0) as a direction only.
template <typename M>
class JacobiSVD {
  public:
    JacobiSVD(const M& A, unsigned int = 0)
    {
        const int n = A.rows();
        assert(n == A.cols());
        MatrixXd a(n, n), v = MatrixXd::Identity(n, n);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) a(i, j) = 0.5 * (A(i, j) + A(j, i));
        for (int sweep = 0; sweep < 64; ++sweep) {
            double off = 0.0;
            for (int p = 0; p < n; ++p)
                for (int q = p + 1; q < n; ++q) off += a(p, q) * a(p, q);
            if (off < 1e-300) break;
            for (int p = 0; p < n; ++p)
                for (int q = p + 1; q < n; ++q) {
                    if (std::fabs(a(p, q)) < 1e-300) continue;
                    const double theta = (a(q, q) - a(p, p)) / (2.0 * a(p, q));
                    const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                    const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
                    for (int k = 0; k < n; ++k) {
                        const double akp = a(k, p), akq = a(k, q);
                        a(k, p) = c * akp - sn * akq;
                        a(k, q) = sn * akp + c * akq;
                    }
                    for (int k = 0; k < n; ++k) {
                        const double apk = a(p, k), aqk = a(q, k);
                        a(p, k) = c * apk - sn * aqk;
                        a(q, k) = sn * apk + c * aqk;
                    }
                    for (int k = 0; k < n; ++k) {
                        const double vkp = v(k, p), vkq = v(k, q);
                        v(k, p) = c * vkp - sn * vkq;
                        v(k, q) = sn * vkp + c * vkq;
                    }
                }
        }
        std::vector<int> order(n);
        for (int i = 0; i < n; ++i) order[i] = i;
        for (int i = 0; i < n; ++i)
            for (int j = i + 1; j < n; ++j)
                if (std::fabs(a(order[j], order[j])) > std::fabs(a(order[i], order[i]))) std::swap(order[i], order[j]);
        U_ = MatrixXd(n, n);
        S_ = VectorXd(n);
        for (int j = 0; j < n; ++j) {
            S_(j) = std::fabs(a(order[j], order[j]));
            for (int i = 0; i < n; ++i) U_(i, j) = v(i, order[j]);
        }
    }
    const MatrixXd& matrixU() const { return U_; }
    const MatrixXd& matrixV() const { return U_; }
    const VectorXd& singularValues() const { return S_; }

  private:
    MatrixXd U_;
    VectorXd S_;
};

// rotation about an axis (only Line3D::rotationFromRPY-style helpers use it, off the hot path)
class AngleAxisd {
  public:
    AngleAxisd(double angle, const Vector3d& axis) : a_(angle), ax_(axis) {}
    Matrix3d toRotationMatrix() const
    {
        const double c = std::cos(a_), s = std::sin(a_), t = 1.0 - c;
        const double x = ax_.x(), y = ax_.y(), z = ax_.z();
        Matrix3d R;
        R << t * x * x + c, t * x * y - s * z, t * x * z + s * y, t * x * y + s * z, t * y * y + c, t * y * z - s * x,
            t * x * z - s * y, t * y * z + s * x, t * z * z + c;
        return R;
    }
    operator Matrix3d() const { return toRotationMatrix(); }

  private:
    double a_;
    Vector3d ax_;
};

}  // namespace Eigen
