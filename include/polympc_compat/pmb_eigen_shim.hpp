// pmb_eigen_shim.hpp — the slice of the Eigen API that PolyMPC *problem classes* use, for host and device.
//
// Why it exists: the reference's user-facing functor concept (dynamics_impl / lagrange_term_impl / mayer_term_impl /
// inequality_constraints_impl, reference src/control/continuous_ocp.hpp:191-288) is written against fixed-size Eigen types
// (Eigen::Matrix<T,N,1>, Eigen::Ref, Eigen::DiagonalMatrix ...).  Real Eigen evaluates those expressions with libm and
// with an evaluation order that depends on the vector ISA; the GPU engine needs every functor to round identically on the
// host and on the device.  This header provides the same *spelling* (so that problem classes compile unchanged) on top of
// plain fixed-size arrays:
//   * no expression templates — every operator returns a small Matrix by value (sizes here are NX, NU <= ~16);
//   * reductions (dot, matrix * vector, sum) follow Eigen's unrolled order for fixed sizes — first half + second half,
//     recursively (Eigen/src/Core/Redux.h, redux_novec_unroller) — for every scalar type;
//   * mixed scalar types (double matrix times dual-number vector) promote like Eigen's ScalarBinaryOpTraits.
// Not a general linear-algebra library: dynamic sizes exist only as the block views / replicate() results that the
// reference's bound-setting idiom uses on the host (`lower_bound_x().segment(40, 4) = v`, `.tail(22) = u.replicate(11,1)`).
#pragma once
#include <cmath>
#include <cstddef>
#include <initializer_list>
#include <ostream>
#include <type_traits>
#include <utility>
#include <vector>

#if defined(__CUDACC__)
#define PMB_EHD __host__ __device__ inline
#else
#define PMB_EHD inline
#endif
#ifndef EIGEN_STRONG_INLINE
#define EIGEN_STRONG_INLINE PMB_EHD
#endif
#ifndef EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#endif

namespace Eigen {

enum { Dynamic = -1, ColMajor = 0, RowMajor = 1 };
using Index = std::ptrdiff_t;

template <class T, int R, int C = 1> struct Matrix;
template <class M> struct Ref;
template <class T> struct VecBlock;
template <class T> struct VecDyn;

template <class A, class B> using promote_t = decltype(std::declval<A>() * std::declval<B>());

namespace internal {
/** Eigen's fixed-size unrolled reduction: sum(start, len) = sum(first len/2) + sum(rest) */
template <class T, int START, int LEN> struct Halving {
    template <class F> PMB_EHD static T run(const F& f) { return Halving<T, START, LEN / 2>::run(f) + Halving<T, START + LEN / 2, LEN - LEN / 2>::run(f); }
};
template <class T, int START> struct Halving<T, START, 1> {
    template <class F> PMB_EHD static T run(const F& f) { return f(START); }
};
template <class T, int START> struct Halving<T, START, 0> {
    template <class F> PMB_EHD static T run(const F&) { return T(0.0); }
};
template <class D> struct traits;
} // namespace internal

/** CRTP root of everything that has compile-time sizes and coefficient access */
template <class D>
struct DenseBase {
    using Scalar = typename internal::traits<D>::Scalar;
    enum { Rows = internal::traits<D>::Rows, Cols = internal::traits<D>::Cols, Size = Rows * Cols,
           RowsAtCompileTime = Rows, ColsAtCompileTime = Cols, SizeAtCompileTime = Size };
    using Plain = Matrix<typename std::remove_const<Scalar>::type, Rows, Cols>;
    PMB_EHD const D& derived() const { return *static_cast<const D*>(this); }
    PMB_EHD D& derived() { return *static_cast<D*>(this); }
    PMB_EHD const Scalar& coeff(int i, int j) const { return derived().data()[i + j * Rows]; }
    PMB_EHD const Scalar& coeff(int i) const { return derived().data()[i]; }
    PMB_EHD const Scalar& operator()(int i) const { return derived().data()[i]; }
    PMB_EHD const Scalar& operator[](int i) const { return derived().data()[i]; }
    PMB_EHD const Scalar& operator()(int i, int j) const { return derived().data()[i + j * Rows]; }
    PMB_EHD const Scalar& x() const { return derived().data()[0]; }
    PMB_EHD const Scalar& y() const { return derived().data()[1]; }
    PMB_EHD const Scalar& z() const { return derived().data()[2]; }
    PMB_EHD static constexpr int rows() { return Rows; }
    PMB_EHD static constexpr int cols() { return Cols; }
    PMB_EHD static constexpr int size() { return Size; }

    PMB_EHD Plain eval() const { Plain r; for (int i = 0; i < Size; ++i) r.m[i] = coeff(i); return r; }
    template <class U> PMB_EHD Matrix<U, Rows, Cols> cast() const
    { Matrix<U, Rows, Cols> r; for (int i = 0; i < Size; ++i) r.m[i] = U(coeff(i)); return r; }
    PMB_EHD Matrix<typename std::remove_const<Scalar>::type, Cols, Rows> transpose() const
    {
        Matrix<typename std::remove_const<Scalar>::type, Cols, Rows> r;
        for (int j = 0; j < Cols; ++j) for (int i = 0; i < Rows; ++i) r.m[j + i * Cols] = coeff(i, j);
        return r;
    }
    /** a.dot(b) = sum_i a_i * b_i in halving order (Eigen/src/Core/Dot.h -> cwiseProduct().sum()) */
    template <class O> PMB_EHD promote_t<typename std::remove_const<Scalar>::type, typename std::remove_const<typename O::Scalar>::type>
    dot(const DenseBase<O>& o) const
    {
        using RT = promote_t<typename std::remove_const<Scalar>::type, typename std::remove_const<typename O::Scalar>::type>;
        static_assert((int)Size == (int)O::Size, "dot: size mismatch");
        const D& a = derived(); const O& b = o.derived();
        return internal::Halving<RT, 0, Size>::run([&](int i) -> RT { return a.data()[i] * b.data()[i]; });
    }
    PMB_EHD typename std::remove_const<Scalar>::type sum() const
    {
        using RT = typename std::remove_const<Scalar>::type;
        const D& a = derived();
        return internal::Halving<RT, 0, Size>::run([&](int i) -> RT { return a.data()[i]; });
    }
    PMB_EHD typename std::remove_const<Scalar>::type squaredNorm() const
    {
        using RT = typename std::remove_const<Scalar>::type;
        const D& a = derived();
        return internal::Halving<RT, 0, Size>::run([&](int i) -> RT { return a.data()[i] * a.data()[i]; });
    }
    PMB_EHD typename std::remove_const<Scalar>::type norm() const { using std::sqrt; return sqrt(squaredNorm()); }
    template <class O> PMB_EHD Matrix<promote_t<typename std::remove_const<Scalar>::type, typename std::remove_const<typename O::Scalar>::type>, Rows, Cols>
    cwiseProduct(const DenseBase<O>& o) const
    {
        Matrix<promote_t<typename std::remove_const<Scalar>::type, typename std::remove_const<typename O::Scalar>::type>, Rows, Cols> r;
        for (int i = 0; i < Size; ++i) r.m[i] = coeff(i) * o.coeff(i);
        return r;
    }
    PMB_EHD Plain cwiseAbs() const
    { using std::fabs; Plain r; for (int i = 0; i < Size; ++i) r.m[i] = coeff(i) < Scalar(0.0) ? -coeff(i) : coeff(i); return r; }
    /** column j / row i as a plain copy (read access: `H.col(i).cwiseAbs().sum()`) */
    PMB_EHD Matrix<typename std::remove_const<Scalar>::type, Rows, 1> col(int j) const
    { Matrix<typename std::remove_const<Scalar>::type, Rows, 1> r; for (int i = 0; i < Rows; ++i) r.m[i] = coeff(i, j); return r; }
    PMB_EHD Matrix<typename std::remove_const<Scalar>::type, 1, Cols> row(int i) const
    { Matrix<typename std::remove_const<Scalar>::type, 1, Cols> r; for (int j = 0; j < Cols; ++j) r.m[j] = coeff(i, j); return r; }
    /** Eigen's isApprox (Eigen/src/Core/Fuzzy.h): ||a - b||^2 <= prec^2 * min(||a||^2, ||b||^2) */
    template <class O> bool isApprox(const DenseBase<O>& o, double prec = 1e-12) const
    {
        double d2 = 0, a2 = 0, b2 = 0;
        for (int i = 0; i < Size; ++i) {
            const double a = (double)coeff(i), b = (double)o.coeff(i);
            d2 += (a - b) * (a - b); a2 += a * a; b2 += b * b;
        }
        return d2 <= prec * prec * (a2 < b2 ? a2 : b2);
    }
    /** host-side dynamic views (bound-setting idiom) */
    VecBlock<const Scalar> segment(int start, int len) const { return VecBlock<const Scalar>{derived().data() + start, len}; }
    VecBlock<const Scalar> head(int len) const { return segment(0, len); }
    VecBlock<const Scalar> tail(int len) const { return segment(Size - len, len); }
    template <int N> PMB_EHD Matrix<typename std::remove_const<Scalar>::type, N, 1> segment(int start) const
    { Matrix<typename std::remove_const<Scalar>::type, N, 1> r; for (int i = 0; i < N; ++i) r.m[i] = coeff(start + i); return r; }
    template <int N> PMB_EHD Matrix<typename std::remove_const<Scalar>::type, N, 1> head() const { return segment<N>(0); }
    template <int N> PMB_EHD Matrix<typename std::remove_const<Scalar>::type, N, 1> tail() const { return segment<N>(Size - N); }
    VecDyn<typename std::remove_const<Scalar>::type> replicate(int rfac, int cfac) const
    {
        VecDyn<typename std::remove_const<Scalar>::type> r;
        (void)cfac;
        for (int k = 0; k < rfac; ++k) for (int i = 0; i < Size; ++i) r.v.push_back(coeff(i));
        return r;
    }
};

/** `m << a, b, c;` — row by row, like Eigen's CommaInitializer */
template <class T, int R, int C>
struct CommaInit {
    T* m; int k;
    PMB_EHD CommaInit& operator,(const T& v) { put(v); return *this; }
    template <class U, class = typename std::enable_if<std::is_arithmetic<U>::value && !std::is_same<U, T>::value>::type>
    PMB_EHD CommaInit& operator,(U v) { put(T(v)); return *this; }
    PMB_EHD void put(const T& v) { const int r = k / C, c = k % C; m[r + c * R] = v; ++k; }
};

template <class T, int R, int C>
struct Matrix : DenseBase<Matrix<T, R, C>> {
    using Base = DenseBase<Matrix<T, R, C>>;
    using Scalar = T;
    T m[R * C > 0 ? R * C : 1];

    PMB_EHD Matrix() {}
    template <class O, class = typename std::enable_if<(int)O::Size == R * C>::type>
    PMB_EHD Matrix(const DenseBase<O>& o) { for (int i = 0; i < R * C; ++i) m[i] = T(o.derived().data()[i]); }
    /** fixed-size vectors from 2..4 coefficients (Eigen's Vector3d(x, y, z)) */
    PMB_EHD Matrix(const T& a, const T& b) { static_assert(R * C == 2, "size"); m[0] = a; m[1] = b; }
    PMB_EHD Matrix(const T& a, const T& b, const T& c) { static_assert(R * C == 3, "size"); m[0] = a; m[1] = b; m[2] = c; }
    PMB_EHD Matrix(const T& a, const T& b, const T& c, const T& d) { static_assert(R * C == 4, "size"); m[0] = a; m[1] = b; m[2] = c; m[3] = d; }
    Matrix(const VecBlock<const T>& b);
    Matrix(const VecBlock<T>& b);

    using Base::operator();
    using Base::operator[];
    using Base::x; using Base::y; using Base::z;
    using Base::segment; using Base::head; using Base::tail;
    PMB_EHD T* data() { return m; }
    PMB_EHD const T* data() const { return m; }
    PMB_EHD T& operator()(int i) { return m[i]; }
    PMB_EHD T& operator[](int i) { return m[i]; }
    PMB_EHD T& operator()(int i, int j) { return m[i + j * R]; }
    PMB_EHD T& coeffRef(int i, int j) { return m[i + j * R]; }
    PMB_EHD T& coeffRef(int i) { return m[i]; }
    PMB_EHD T& x() { return m[0]; }
    PMB_EHD T& y() { return m[1]; }
    PMB_EHD T& z() { return m[2]; }

    template <class O> PMB_EHD Matrix& operator=(const DenseBase<O>& o)
    { static_assert((int)O::Size == R * C, "size mismatch"); for (int i = 0; i < R * C; ++i) m[i] = T(o.derived().data()[i]); return *this; }
    Matrix& operator=(const VecDyn<T>& o) { for (int i = 0; i < R * C; ++i) m[i] = o.v[i]; return *this; }

    PMB_EHD Matrix& setZero() { for (int i = 0; i < R * C; ++i) m[i] = T(0.0); return *this; }
    PMB_EHD Matrix& setOnes() { for (int i = 0; i < R * C; ++i) m[i] = T(1.0); return *this; }
    PMB_EHD Matrix& setConstant(const T& v) { for (int i = 0; i < R * C; ++i) m[i] = v; return *this; }
    PMB_EHD Matrix& setIdentity() { for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) m[i + j * R] = T(i == j ? 1.0 : 0.0); return *this; }
    PMB_EHD static Matrix Zero() { Matrix r; r.setZero(); return r; }
    PMB_EHD static Matrix Ones() { Matrix r; r.setOnes(); return r; }
    PMB_EHD static Matrix Constant(const T& v) { Matrix r; r.setConstant(v); return r; }
    PMB_EHD static Matrix Identity() { Matrix r; r.setIdentity(); return r; }

    PMB_EHD CommaInit<T, R, C> operator<<(const T& v) { CommaInit<T, R, C> ci{m, 0}; ci.put(v); return ci; }
    template <class U, class = typename std::enable_if<std::is_arithmetic<U>::value && !std::is_same<U, T>::value>::type>
    PMB_EHD CommaInit<T, R, C> operator<<(U v) { CommaInit<T, R, C> ci{m, 0}; ci.put(T(v)); return ci; }

    /** view of the main diagonal: `Q.diagonal() << a, b, c;`, `Q.diagonal()(i)` */
    struct Diag {
        T* m;
        PMB_EHD T& operator()(int i) { return m[i * (R + 1)]; }
        struct DiagComma { T* m; int k; PMB_EHD DiagComma& operator,(const T& v) { m[k * (R + 1)] = v; ++k; return *this; } };
        PMB_EHD DiagComma operator<<(const T& v) { m[0] = v; return DiagComma{m, 1}; }
    };
    PMB_EHD Diag diagonal() { return Diag{m}; }
    PMB_EHD Matrix<T, (R < C ? R : C), 1> diagonal() const { Matrix<T, (R < C ? R : C), 1> r; for (int i = 0; i < (R < C ? R : C); ++i) r.m[i] = m[i * (R + 1)]; return r; }

    VecBlock<T> segment(int start, int len) { return VecBlock<T>{m + start, len}; }
    VecBlock<T> head(int len) { return segment(0, len); }
    VecBlock<T> tail(int len) { return segment(R * C - len, len); }
    PMB_EHD Matrix& noalias() { return *this; }              // every shim product is evaluated into a temporary anyway

    template <class O> PMB_EHD Matrix& operator+=(const DenseBase<O>& o) { for (int i = 0; i < R * C; ++i) m[i] = m[i] + o.derived().data()[i]; return *this; }
    template <class O> PMB_EHD Matrix& operator-=(const DenseBase<O>& o) { for (int i = 0; i < R * C; ++i) m[i] = m[i] - o.derived().data()[i]; return *this; }
    PMB_EHD Matrix& operator*=(const T& s) { for (int i = 0; i < R * C; ++i) m[i] = m[i] * s; return *this; }
    PMB_EHD Matrix& operator/=(const T& s) { for (int i = 0; i < R * C; ++i) m[i] = m[i] / s; return *this; }
};
namespace internal { template <class T, int R, int C> struct traits<Matrix<T, R, C>> { using Scalar = T; enum { Rows = R, Cols = C }; }; }

/** Eigen::Ref of a fixed-size plain matrix: a pointer view (contiguous, column-major) */
template <class T, int R, int C>
struct Ref<const Matrix<T, R, C>> : DenseBase<Ref<const Matrix<T, R, C>>> {
    using Scalar = T;
    const T* p;
    PMB_EHD explicit Ref(const T* ptr) : p(ptr) {}
    PMB_EHD Ref(const Matrix<T, R, C>& mat) : p(mat.m) {}
    PMB_EHD Ref(const Ref<Matrix<T, R, C>>& o);
    PMB_EHD const T* data() const { return p; }
};
template <class T, int R, int C>
struct Ref<Matrix<T, R, C>> : DenseBase<Ref<Matrix<T, R, C>>> {
    using Base = DenseBase<Ref<Matrix<T, R, C>>>;
    using Scalar = T;
    T* p;
    PMB_EHD explicit Ref(T* ptr) : p(ptr) {}
    PMB_EHD Ref(Matrix<T, R, C>& mat) : p(mat.m) {}
    PMB_EHD T* data() const { return p; }
    using Base::operator();
    using Base::operator[];
    using Base::segment; using Base::head; using Base::tail;
    PMB_EHD T& operator()(int i) { return p[i]; }
    PMB_EHD T& operator[](int i) { return p[i]; }
    PMB_EHD T& operator()(int i, int j) { return p[i + j * R]; }
    PMB_EHD T& coeffRef(int i, int j) { return p[i + j * R]; }
    template <class O> PMB_EHD Ref& operator=(const DenseBase<O>& o) { for (int i = 0; i < R * C; ++i) p[i] = T(o.derived().data()[i]); return *this; }
    PMB_EHD Ref& operator=(const Ref& o) { for (int i = 0; i < R * C; ++i) p[i] = o.p[i]; return *this; }
    PMB_EHD Ref(const Ref& o) : p(o.p) {}
    PMB_EHD Ref& setZero() { for (int i = 0; i < R * C; ++i) p[i] = T(0.0); return *this; }
    PMB_EHD CommaInit<T, R, C> operator<<(const T& v) { CommaInit<T, R, C> ci{p, 0}; ci.put(v); return ci; }
    VecBlock<T> segment(int start, int len) { return VecBlock<T>{p + start, len}; }
    VecBlock<T> head(int len) { return segment(0, len); }
    VecBlock<T> tail(int len) { return segment(R * C - len, len); }
};
template <class T, int R, int C> PMB_EHD Ref<const Matrix<T, R, C>>::Ref(const Ref<Matrix<T, R, C>>& o) : p(o.p) {}
namespace internal {
template <class T, int R, int C> struct traits<Ref<const Matrix<T, R, C>>> { using Scalar = T; enum { Rows = R, Cols = C }; };
template <class T, int R, int C> struct traits<Ref<Matrix<T, R, C>>> { using Scalar = T; enum { Rows = R, Cols = C }; };
}

/** host-only dynamic pieces: a writable run of coefficients and an owned dynamic vector */
template <class T> struct VecDyn { std::vector<T> v; int size() const { return (int)v.size(); } const T& operator()(int i) const { return v[i]; } };
template <class T>
struct VecBlock {
    T* p; int n;
    int size() const { return n; }
    T& operator()(int i) const { return p[i]; }
    template <class O> const VecBlock& operator=(const DenseBase<O>& o) const { for (int i = 0; i < n; ++i) p[i] = o.derived().data()[i]; return *this; }
    const VecBlock& operator=(const VecDyn<typename std::remove_const<T>::type>& o) const { for (int i = 0; i < n; ++i) p[i] = o.v[i]; return *this; }
    const VecBlock& operator=(const VecBlock<const typename std::remove_const<T>::type>& o) const { for (int i = 0; i < n; ++i) p[i] = o.p[i]; return *this; }
    const VecBlock& setConstant(const typename std::remove_const<T>::type& v) const { for (int i = 0; i < n; ++i) p[i] = v; return *this; }
    const VecBlock& setZero() const { return setConstant(0); }
};
template <class T, int R, int C> Matrix<T, R, C>::Matrix(const VecBlock<const T>& b) { for (int i = 0; i < R * C; ++i) m[i] = b.p[i]; }
template <class T, int R, int C> Matrix<T, R, C>::Matrix(const VecBlock<T>& b) { for (int i = 0; i < R * C; ++i) m[i] = b.p[i]; }

// ---- element-wise binary operators (any two dense operands of equal size) -------------------------------------------
#define PMB_EIGEN_RS(A) typename std::remove_const<typename A::Scalar>::type
template <class A, class B>
PMB_EHD Matrix<promote_t<PMB_EIGEN_RS(A), PMB_EIGEN_RS(B)>, A::Rows, A::Cols> operator+(const DenseBase<A>& a, const DenseBase<B>& b)
{
    static_assert((int)A::Rows == (int)B::Rows && (int)A::Cols == (int)B::Cols, "operator+: size mismatch");
    Matrix<promote_t<PMB_EIGEN_RS(A), PMB_EIGEN_RS(B)>, A::Rows, A::Cols> r;
    for (int i = 0; i < A::Size; ++i) r.m[i] = a.coeff(i) + b.coeff(i);
    return r;
}
template <class A, class B>
PMB_EHD Matrix<promote_t<PMB_EIGEN_RS(A), PMB_EIGEN_RS(B)>, A::Rows, A::Cols> operator-(const DenseBase<A>& a, const DenseBase<B>& b)
{
    static_assert((int)A::Rows == (int)B::Rows && (int)A::Cols == (int)B::Cols, "operator-: size mismatch");
    Matrix<promote_t<PMB_EIGEN_RS(A), PMB_EIGEN_RS(B)>, A::Rows, A::Cols> r;
    for (int i = 0; i < A::Size; ++i) r.m[i] = a.coeff(i) - b.coeff(i);
    return r;
}
template <class A> PMB_EHD Matrix<PMB_EIGEN_RS(A), A::Rows, A::Cols> operator-(const DenseBase<A>& a)
{ Matrix<PMB_EIGEN_RS(A), A::Rows, A::Cols> r; for (int i = 0; i < A::Size; ++i) r.m[i] = -a.coeff(i); return r; }

/** matrix product, coefficient (i, j) = sum_k a(i,k) * b(k,j) in halving order (Eigen's lazy coefficient-based product,
 *  which is what fixed sizes this small compile to) */
template <class A, class B>
PMB_EHD Matrix<promote_t<PMB_EIGEN_RS(A), PMB_EIGEN_RS(B)>, A::Rows, B::Cols> operator*(const DenseBase<A>& a, const DenseBase<B>& b)
{
    static_assert((int)A::Cols == (int)B::Rows, "operator*: inner sizes differ");
    using RT = promote_t<PMB_EIGEN_RS(A), PMB_EIGEN_RS(B)>;
    Matrix<RT, A::Rows, B::Cols> r;
    const A& ad = a.derived(); const B& bd = b.derived();
    for (int j = 0; j < B::Cols; ++j)
        for (int i = 0; i < A::Rows; ++i)
            r.m[i + j * A::Rows] = internal::Halving<RT, 0, A::Cols>::run(
                [&](int k) -> RT { return ad.data()[i + k * A::Rows] * bd.data()[k + j * B::Rows]; });
    return r;
}
/** scalar * matrix, matrix * scalar, matrix / scalar; the scalar is the operand's own scalar type or a plain double */
template <class A> PMB_EHD Matrix<PMB_EIGEN_RS(A), A::Rows, A::Cols> operator*(const DenseBase<A>& a, const PMB_EIGEN_RS(A)& s)
{ Matrix<PMB_EIGEN_RS(A), A::Rows, A::Cols> r; for (int i = 0; i < A::Size; ++i) r.m[i] = a.coeff(i) * s; return r; }
template <class A> PMB_EHD Matrix<PMB_EIGEN_RS(A), A::Rows, A::Cols> operator*(const PMB_EIGEN_RS(A)& s, const DenseBase<A>& a)
{ Matrix<PMB_EIGEN_RS(A), A::Rows, A::Cols> r; for (int i = 0; i < A::Size; ++i) r.m[i] = s * a.coeff(i); return r; }
template <class A> PMB_EHD Matrix<PMB_EIGEN_RS(A), A::Rows, A::Cols> operator/(const DenseBase<A>& a, const PMB_EIGEN_RS(A)& s)
{ Matrix<PMB_EIGEN_RS(A), A::Rows, A::Cols> r; for (int i = 0; i < A::Size; ++i) r.m[i] = a.coeff(i) / s; return r; }
template <class A, class = typename std::enable_if<!std::is_same<PMB_EIGEN_RS(A), double>::value>::type>
PMB_EHD Matrix<PMB_EIGEN_RS(A), A::Rows, A::Cols> operator*(const DenseBase<A>& a, double s)
{ Matrix<PMB_EIGEN_RS(A), A::Rows, A::Cols> r; for (int i = 0; i < A::Size; ++i) r.m[i] = a.coeff(i) * s; return r; }
template <class A, class = typename std::enable_if<!std::is_same<PMB_EIGEN_RS(A), double>::value>::type>
PMB_EHD Matrix<PMB_EIGEN_RS(A), A::Rows, A::Cols> operator*(double s, const DenseBase<A>& a)
{ Matrix<PMB_EIGEN_RS(A), A::Rows, A::Cols> r; for (int i = 0; i < A::Size; ++i) r.m[i] = s * a.coeff(i); return r; }
template <class A, class = typename std::enable_if<!std::is_same<PMB_EIGEN_RS(A), double>::value>::type>
PMB_EHD Matrix<PMB_EIGEN_RS(A), A::Rows, A::Cols> operator/(const DenseBase<A>& a, double s)
{ Matrix<PMB_EIGEN_RS(A), A::Rows, A::Cols> r; for (int i = 0; i < A::Size; ++i) r.m[i] = a.coeff(i) / s; return r; }

/** Eigen::DiagonalMatrix<T, N>: `DiagonalMatrix<double,3> Q{1,1,1}`, Q.diagonal() << ..., Q.toDenseMatrix(), Q * v */
template <class T, int N>
struct DiagonalMatrix {
    Matrix<T, N, 1> d;
    PMB_EHD DiagonalMatrix() {}
    PMB_EHD DiagonalMatrix(const T& a, const T& b) { static_assert(N == 2, "size"); d.m[0] = a; d.m[1] = b; }
    PMB_EHD DiagonalMatrix(const T& a, const T& b, const T& c) { static_assert(N == 3, "size"); d.m[0] = a; d.m[1] = b; d.m[2] = c; }
    DiagonalMatrix(std::initializer_list<T> l) { int i = 0; for (const T& v : l) if (i < N) d.m[i++] = v; }
    PMB_EHD Matrix<T, N, 1>& diagonal() { return d; }
    PMB_EHD const Matrix<T, N, 1>& diagonal() const { return d; }
    PMB_EHD Matrix<T, N, N> toDenseMatrix() const
    { Matrix<T, N, N> r; for (int j = 0; j < N; ++j) for (int i = 0; i < N; ++i) r.m[i + j * N] = (i == j) ? d.m[i] : T(0.0); return r; }
    PMB_EHD void setZero() { d.setZero(); }
    PMB_EHD void setIdentity() { d.setOnes(); }
};
template <class T, int N, class B>
PMB_EHD Matrix<promote_t<T, PMB_EIGEN_RS(B)>, N, B::Cols> operator*(const DiagonalMatrix<T, N>& a, const DenseBase<B>& b)
{
    Matrix<promote_t<T, PMB_EIGEN_RS(B)>, N, B::Cols> r;
    for (int j = 0; j < B::Cols; ++j) for (int i = 0; i < N; ++i) r.m[i + j * N] = a.d.m[i] * b.coeff(i, j);
    return r;
}
#undef PMB_EIGEN_RS

template <class D> std::ostream& operator<<(std::ostream& os, const DenseBase<D>& a)
{
    for (int i = 0; i < D::Rows; ++i) { for (int j = 0; j < D::Cols; ++j) os << (j ? " " : "") << (double)a.coeff(i, j); if (i + 1 < D::Rows) os << "\n"; }
    return os;
}

using Vector2d = Matrix<double, 2, 1>; using Vector3d = Matrix<double, 3, 1>; using Vector4d = Matrix<double, 4, 1>;
using Matrix2d = Matrix<double, 2, 2>; using Matrix3d = Matrix<double, 3, 3>; using Matrix4d = Matrix<double, 4, 4>;
template <class T, int N> using Vector = Matrix<T, N, 1>;

} // namespace Eigen
