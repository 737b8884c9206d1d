// oracle/shim/RcppEigen.h -- a minimal stand-in for <RcppEigen.h> / <Rcpp.h> so that the reference's OWN solver sources
// (/root/reference/src/*.cpp, unmodified, compiled where they lie) build into oracle/_ref/libbwgr_ref.so without R, Rcpp or
// Eigen (none of which exist in this image).  TEST INFRASTRUCTURE ONLY (see oracle/bwgr_oracle.hpp): it pins the
// hand-written oracle against the reference's text; nothing under bwgr_b200/ may include or link it.
//
// What it is: eager (no expression templates) dense Matrix / Vector / Array types with the subset of the Eigen API those
// sources use, LLT, a Jacobi SelfAdjointEigenSolver, a Jacobi-SVD pseudo-inverse behind completeOrthogonalDecomposition(),
// Rcpp::List / Named / Nullable / NumericVector, and R::rnorm / rchisq / rbinom on a seedable std::mt19937_64.
// What it is not: Eigen.  Reductions are plain loops (8 interleaved partial sums on contiguous data), so float results
// agree with an RcppEigen build up to reassociation -- the same caveat as any two builds of the reference with different
// compiler flags.  Statement order, operand types, integer / float casts and every quirk of the sources are the sources'.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <limits>
#include <memory>
#include <random>
#include <sstream>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>
// Eigen/Core includes <emmintrin.h> -> <mm_malloc.h> -> <stdlib.h> on x86-64; libstdc++'s <stdlib.h> wrapper puts the std::abs
// overloads into the global namespace, so the sources' unqualified abs(double) (RcppEigen20230423.cpp:609, :638) is the
// floating-point one in an RcppEigen build.  Same include here, same overload set (without it abs(double) would truncate).
#include <stdlib.h>
#include <math.h>

typedef void* SEXP;
#define R_NilValue ((SEXP)0)

namespace Eigen {

typedef std::ptrdiff_t Index;
enum ComputationInfo { Success = 0, NumericalIssue = 1, NoConvergence = 2, InvalidInput = 3 };
enum { ComputeFullU = 0x04, ComputeThinU = 0x08, ComputeFullV = 0x10, ComputeThinV = 0x20 };
inline void setNbThreads(int) {}
inline void initParallel() {}

template <class T> struct Mat;
template <class T> struct Vec;
template <class T> struct View;
template <class T> struct Scaled;
template <class T> struct Prod;
template <class T> struct Arr;
template <class T> struct AView;

// kind: 1 = matrix expression, 2 = array expression
template <class X> struct kind { static const int v = 0; };
template <class T> struct kind<Mat<T>> { static const int v = 1; typedef T S; };
template <class T> struct kind<Vec<T>> { static const int v = 1; typedef T S; };
template <class T> struct kind<View<T>> { static const int v = 1; typedef T S; };
template <class T> struct kind<Scaled<T>> { static const int v = 1; typedef T S; };
template <class T> struct kind<Prod<T>> { static const int v = 1; typedef T S; };
template <class T> struct kind<Arr<T>> { static const int v = 2; typedef T S; };
template <class T> struct kind<AView<T>> { static const int v = 2; typedef T S; };
#define SHIM_IF(...) typename std::enable_if<(__VA_ARGS__), int>::type = 0

// 8 interleaved partial sums on contiguous data (the compiler may vectorise each lane independently)
template <class T, class F> inline T reduce_sum(Index n, F f) {
  T a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  Index i = 0;
  for (; i + 8 <= n; i += 8)
    for (int l = 0; l < 8; l++) a[l] += f(i + l);
  T s = ((a[0] + a[4]) + (a[2] + a[6])) + ((a[1] + a[5]) + (a[3] + a[7]));
  for (; i < n; i++) s += f(i);
  return s;
}

template <class T> struct LLT;
template <class T> struct COD;
template <class D, class T> struct ColwiseOp;
template <class D, class T> struct RowwiseOp;

// ---------------------------------------------------------------------------------------------------------------------
// read-only API shared by every matrix-like and array-like type (Derived provides rows(), cols(), operator()(i,j))
// ---------------------------------------------------------------------------------------------------------------------
template <class D, class T> struct DenseRO {
  typedef T Scalar;
  const D& self() const { return *static_cast<const D*>(this); }
  Index size() const { return (Index)self().rows() * self().cols(); }
  T lin(Index i) const { const D& s = self(); return s.cols() == 1 ? s(i, 0) : s.rows() == 1 ? s(0, i) : s(i % s.rows(), i / s.rows()); }
  T sum() const {
    const D& s = self();
    if (s.rows() == 1 || s.cols() == 1) return reduce_sum<T>(size(), [&](Index i) { return lin(i); });
    T t = 0;
    for (Index j = 0; j < s.cols(); j++) t += reduce_sum<T>(s.rows(), [&](Index i) { return s(i, j); });
    return t;
  }
  T mean() const { return sum() / (T)size(); }
  T prod() const { T t = 1; for (Index i = 0; i < size(); i++) t *= lin(i); return t; }
  T squaredNorm() const {
    const D& s = self();
    if (s.rows() == 1 || s.cols() == 1) return reduce_sum<T>(size(), [&](Index i) { const T v = lin(i); return v * v; });
    T t = 0;
    for (Index j = 0; j < s.cols(); j++) t += reduce_sum<T>(s.rows(), [&](Index i) { const T v = s(i, j); return v * v; });
    return t;
  }
  T norm() const { return std::sqrt(squaredNorm()); }
  T minCoeff() const { T m = lin(0); for (Index i = 1; i < size(); i++) m = std::min(m, lin(i)); return m; }
  T maxCoeff() const { T m = lin(0); for (Index i = 1; i < size(); i++) m = std::max(m, lin(i)); return m; }
  T trace() const { T t = 0; for (Index i = 0; i < std::min(self().rows(), self().cols()); i++) t += self()(i, i); return t; }
  template <class O> T dot(const O& o) const {
    assert(o.size() == size());
    return reduce_sum<T>(size(), [&](Index i) { return lin(i) * o.lin(i); });
  }
  bool allFinite() const { for (Index i = 0; i < size(); i++) if (!std::isfinite(lin(i))) return false; return true; }
  bool hasNaN() const { for (Index i = 0; i < size(); i++) if (std::isnan(lin(i))) return true; return false; }
  ColwiseOp<D, T> colwise() const { return ColwiseOp<D, T>{self()}; }
  RowwiseOp<D, T> rowwise() const { return RowwiseOp<D, T>{self()}; }
};

// ---------------------------------------------------------------------------------------------------------------------
// strided view onto matrix storage (column, row, diagonal, transpose, block) -- assignable
// ---------------------------------------------------------------------------------------------------------------------
template <class T> struct View : DenseRO<View<T>, T> {
  T* p; Index r, c, rs, cs;
  View(T* p_, Index r_, Index c_, Index rs_, Index cs_) : p(p_), r(r_), c(c_), rs(rs_), cs(cs_) {}
  View(const View&) = default;
  Index rows() const { return r; }
  Index cols() const { return c; }
  T& operator()(Index i, Index j) const { return p[i * rs + j * cs]; }
  T& operator()(Index i) const { return c == 1 ? p[i * rs] : p[i * cs]; }
  T& operator[](Index i) const { return (*this)(i); }
  bool contiguous() const { return (c == 1 && rs == 1) || (r == 1 && cs == 1); }
  T* data() const { return p; }
  template <class X> void assign_from(const X& x) const {
    if (x.rows() == r && x.cols() == c) { for (Index j = 0; j < c; j++) for (Index i = 0; i < r; i++) (*this)(i, j) = x(i, j); }
    else { assert((r == 1 || c == 1) && x.size() == this->size()); for (Index i = 0; i < this->size(); i++) (*this)(i) = x.lin(i); }
  }
  const View& operator=(const View& o) const { Mat<T> t(o); assign_from(t); return *this; }
  template <class X, SHIM_IF(kind<X>::v != 0)> const View& operator=(const X& x) const { Mat<T> t(x); assign_from(t); return *this; }
  template <class X, SHIM_IF(kind<X>::v != 0)> const View& operator+=(const X& x) const { Mat<T> t(x); for (Index i = 0; i < this->size(); i++) lref(i) += t.lin(i); return *this; }
  template <class X, SHIM_IF(kind<X>::v != 0)> const View& operator-=(const X& x) const { Mat<T> t(x); for (Index i = 0; i < this->size(); i++) lref(i) -= t.lin(i); return *this; }
  template <class S, SHIM_IF(std::is_arithmetic<S>::value)> const View& operator*=(S s) const { for (Index i = 0; i < this->size(); i++) lref(i) *= (T)s; return *this; }
  template <class S, SHIM_IF(std::is_arithmetic<S>::value)> const View& operator/=(S s) const { for (Index i = 0; i < this->size(); i++) lref(i) /= (T)s; return *this; }
  T& lref(Index i) const { return c == 1 ? p[i * rs] : r == 1 ? p[i * cs] : p[(i % r) * rs + (i / r) * cs]; }
  void setZero() const { for (Index i = 0; i < this->size(); i++) lref(i) = 0; }
  void setConstant(T v) const { for (Index i = 0; i < this->size(); i++) lref(i) = v; }
  void fill(T v) const { setConstant(v); }
  View col(Index j) const { return View(p + j * cs, r, 1, rs, cs); }
  View row(Index i) const { return View(p + i * rs, 1, c, rs, cs); }
  View transpose() const { return View(p, c, r, cs, rs); }
  View adjoint() const { return transpose(); }
  View diagonal() const { return View(p, std::min(r, c), 1, rs + cs, rs + cs); }
  View block(Index i0, Index j0, Index nr, Index nc) const { return View(p + i0 * rs + j0 * cs, nr, nc, rs, cs); }
  View head(Index n) const { return c == 1 ? block(0, 0, n, 1) : block(0, 0, 1, n); }
  View tail(Index n) const { return c == 1 ? block(r - n, 0, n, 1) : block(0, c - n, 1, n); }
  View segment(Index i0, Index n) const { return c == 1 ? block(i0, 0, n, 1) : block(0, i0, 1, n); }
  AView<T> array() const;
  View matrix() const { return *this; }
  Mat<T> eval() const { return Mat<T>(*this); }
  Mat<T> cwiseAbs() const; Mat<T> cwiseAbs2() const; Mat<T> cwiseInverse() const; Mat<T> cwiseSqrt() const;
  template <class O> Mat<T> cwiseProduct(const O& o) const; template <class O> Mat<T> cwiseQuotient(const O& o) const;
  Mat<T> asDiagonal() const; Mat<T> inverse() const; LLT<T> llt() const; COD<T> completeOrthogonalDecomposition() const;
  template <class U> Mat<U> cast() const;
};

// ---------------------------------------------------------------------------------------------------------------------
// owning column-major matrix
// ---------------------------------------------------------------------------------------------------------------------
template <class T> struct Mat : DenseRO<Mat<T>, T> {
  Index r, c;
  std::vector<T> d;
  Mat() : r(0), c(0) {}
  Mat(Index r_, Index c_) : r(r_), c(c_), d((size_t)r_ * c_) {}
  Mat(const Mat&) = default;
  Mat(Mat&&) = default;
  template <class X, SHIM_IF(kind<X>::v != 0)> Mat(const X& x) : r(x.rows()), c(x.cols()), d((size_t)x.rows() * x.cols()) {
    for (Index j = 0; j < c; j++) for (Index i = 0; i < r; i++) d[(size_t)j * r + i] = x(i, j);
  }
  Mat(const Scaled<T>& x);
  Index rows() const { return r; }
  Index cols() const { return c; }
  T& operator()(Index i, Index j) { return d[(size_t)j * r + i]; }
  const T& operator()(Index i, Index j) const { return d[(size_t)j * r + i]; }
  T& operator()(Index i) { return d[i]; }
  const T& operator()(Index i) const { return d[i]; }
  T& operator[](Index i) { return d[i]; }
  const T& operator[](Index i) const { return d[i]; }
  T* data() { return d.data(); }
  const T* data() const { return d.data(); }
  View<T> v() const { return View<T>(const_cast<T*>(d.data()), r, c, 1, r); }
  void resize(Index r_, Index c_) { r = r_; c = c_; d.assign((size_t)r_ * c_, T(0)); }
  void resize(Index n) { if (c == 1 || r == 0) { r = n; c = 1; } else { c = n; } d.assign((size_t)r * c, T(0)); }
  void conservativeResize(Index r_, Index c_) { Mat t(r_, c_); for (Index j = 0; j < std::min(c, c_); j++) for (Index i = 0; i < std::min(r, r_); i++) t(i, j) = (*this)(i, j); *this = std::move(t); }
  Mat& operator=(const Mat&) = default;
  Mat& operator=(Mat&&) = default;
  template <class X, SHIM_IF(kind<X>::v != 0)> Mat& operator=(const X& x) { Mat t(x); r = t.r; c = t.c; d = std::move(t.d); return *this; }
  template <class X, SHIM_IF(kind<X>::v != 0)> Mat& operator+=(const X& x) { Mat t(x); assert(t.size() == this->size()); for (size_t i = 0; i < d.size(); i++) d[i] += t.lin(i); return *this; }
  template <class X, SHIM_IF(kind<X>::v != 0)> Mat& operator-=(const X& x) { Mat t(x); assert(t.size() == this->size()); for (size_t i = 0; i < d.size(); i++) d[i] -= t.lin(i); return *this; }
  Mat& operator-=(const Scaled<T>& x);   // e -= X.col(j) * s : one fused pass, like Eigen's
  Mat& operator+=(const Scaled<T>& x);
  template <class S, SHIM_IF(std::is_arithmetic<S>::value)> Mat& operator*=(S s) { for (auto& x : d) x *= (T)s; return *this; }
  template <class S, SHIM_IF(std::is_arithmetic<S>::value)> Mat& operator/=(S s) { for (auto& x : d) x /= (T)s; return *this; }
  Mat& setZero() { std::fill(d.begin(), d.end(), T(0)); return *this; }
  Mat& setZero(Index r_, Index c_) { resize(r_, c_); return *this; }
  Mat& setOnes() { std::fill(d.begin(), d.end(), T(1)); return *this; }
  Mat& setConstant(T v_) { std::fill(d.begin(), d.end(), v_); return *this; }
  Mat& setIdentity() { setZero(); for (Index i = 0; i < std::min(r, c); i++) (*this)(i, i) = 1; return *this; }
  void fill(T v_) { setConstant(v_); }
  static Mat Zero(Index r_, Index c_) { return Mat(r_, c_); }
  static Mat Ones(Index r_, Index c_) { Mat m(r_, c_); m.setOnes(); return m; }
  static Mat Constant(Index r_, Index c_, T v_) { Mat m(r_, c_); m.setConstant(v_); return m; }
  static Mat Identity(Index r_, Index c_) { Mat m(r_, c_); m.setIdentity(); return m; }
  View<T> col(Index j) const { return v().col(j); }
  View<T> row(Index i) const { return v().row(i); }
  View<T> transpose() const { return v().transpose(); }
  View<T> adjoint() const { return v().transpose(); }
  View<T> diagonal() const { return v().diagonal(); }
  View<T> block(Index i0, Index j0, Index nr, Index nc) const { return v().block(i0, j0, nr, nc); }
  View<T> leftCols(Index n) const { return block(0, 0, r, n); }
  View<T> rightCols(Index n) const { return block(0, c - n, r, n); }
  View<T> topRows(Index n) const { return block(0, 0, n, c); }
  View<T> bottomRows(Index n) const { return block(r - n, 0, n, c); }
  View<T> head(Index n) const { return v().head(n); }
  View<T> tail(Index n) const { return v().tail(n); }
  View<T> segment(Index i0, Index n) const { return v().segment(i0, n); }
  AView<T> array() const;
  const Mat& matrix() const { return *this; }
  const Mat& eval() const { return *this; }
  Mat cwiseAbs() const { Mat m(r, c); for (size_t i = 0; i < d.size(); i++) m.d[i] = std::abs(d[i]); return m; }
  Mat cwiseAbs2() const { Mat m(r, c); for (size_t i = 0; i < d.size(); i++) m.d[i] = d[i] * d[i]; return m; }
  Mat cwiseInverse() const { Mat m(r, c); for (size_t i = 0; i < d.size(); i++) m.d[i] = T(1) / d[i]; return m; }
  Mat cwiseSqrt() const { Mat m(r, c); for (size_t i = 0; i < d.size(); i++) m.d[i] = std::sqrt(d[i]); return m; }
  template <class O> Mat cwiseProduct(const O& o) const { Mat t(o); assert(t.size() == this->size()); Mat m(r, c); for (size_t i = 0; i < d.size(); i++) m.d[i] = d[i] * t.lin(i); return m; }
  template <class O> Mat cwiseQuotient(const O& o) const { Mat t(o); Mat m(r, c); for (size_t i = 0; i < d.size(); i++) m.d[i] = d[i] / t.lin(i); return m; }
  template <class O> Mat cwiseMax(const O& o) const { Mat t(o); Mat m(r, c); for (size_t i = 0; i < d.size(); i++) m.d[i] = std::max(d[i], t.lin(i)); return m; }
  template <class O> Mat cwiseMin(const O& o) const { Mat t(o); Mat m(r, c); for (size_t i = 0; i < d.size(); i++) m.d[i] = std::min(d[i], t.lin(i)); return m; }
  Mat asDiagonal() const { const Index n = this->size(); Mat m(n, n); for (Index i = 0; i < n; i++) m(i, i) = d[i]; return m; }
  Mat inverse() const;   // Gauss-Jordan with partial pivoting
  T determinant() const;
  LLT<T> llt() const;
  COD<T> completeOrthogonalDecomposition() const;
  template <class U> Mat<U> cast() const { Mat<U> m(r, c); for (size_t i = 0; i < d.size(); i++) m.d[i] = (U)d[i]; return m; }
};

template <class T> struct Vec : Mat<T> {
  Vec() : Mat<T>() { this->c = 1; }
  explicit Vec(Index n) : Mat<T>(n, 1) {}
  Vec(const Vec&) = default;
  Vec(Vec&&) = default;
  Vec(const Mat<T>& m) : Mat<T>(m) { fix(); }
  Vec(Mat<T>&& m) : Mat<T>(std::move(m)) { fix(); }
  template <class X, SHIM_IF(kind<X>::v != 0 && !std::is_base_of<Mat<T>, X>::value)> Vec(const X& x) : Mat<T>(x) { fix(); }
  void fix() { if (this->c != 1) { assert(this->r == 1); this->r = this->c; this->c = 1; } }
  Vec& operator=(const Vec&) = default;
  Vec& operator=(Vec&&) = default;
  template <class X, SHIM_IF(kind<X>::v != 0)> Vec& operator=(const X& x) { Mat<T>::operator=(x); fix(); return *this; }
  static Vec Zero(Index n) { return Vec(n); }
  static Vec Ones(Index n) { Vec m(n); m.setOnes(); return m; }
  static Vec Constant(Index n, T v_) { Vec m(n); m.setConstant(v_); return m; }
  static Vec LinSpaced(Index n, T lo, T hi) { Vec m(n); for (Index i = 0; i < n; i++) m[i] = n > 1 ? lo + (hi - lo) * (T)i / (T)(n - 1) : hi; return m; }
};

// lazy X.col(j) * s -- only so that `e -= X.col(j) * s` is one fused pass; everywhere else it materialises
template <class T> struct Scaled : DenseRO<Scaled<T>, T> {
  View<T> v; T s;
  Scaled(const View<T>& v_, T s_) : v(v_), s(s_) {}
  Index rows() const { return v.rows(); }
  Index cols() const { return v.cols(); }
  T operator()(Index i, Index j) const { return v(i, j) * s; }
};
template <class T> Mat<T>::Mat(const Scaled<T>& x) : r(x.rows()), c(x.cols()), d((size_t)x.rows() * x.cols()) {
  for (Index j = 0; j < c; j++) for (Index i = 0; i < r; i++) d[(size_t)j * r + i] = x(i, j);
}
template <class T> Mat<T>& Mat<T>::operator-=(const Scaled<T>& x) {
  assert(x.size() == this->size());
  if (x.v.contiguous()) { const T* q = x.v.p; const T s = x.s; T* e = d.data(); const Index n = this->size(); for (Index i = 0; i < n; i++) e[i] -= q[i] * s; }
  else for (Index i = 0; i < this->size(); i++) d[i] -= x.lin(i);
  return *this;
}
template <class T> Mat<T>& Mat<T>::operator+=(const Scaled<T>& x) {
  assert(x.size() == this->size());
  if (x.v.contiguous()) { const T* q = x.v.p; const T s = x.s; T* e = d.data(); const Index n = this->size(); for (Index i = 0; i < n; i++) e[i] += q[i] * s; }
  else for (Index i = 0; i < this->size(); i++) d[i] += x.lin(i);
  return *this;
}

// matrix product result: a 1 x 1 product converts to its scalar (Eigen's inner product)
template <class T> struct Prod : Mat<T> {
  Prod(Index r_, Index c_) : Mat<T>(r_, c_) {}
  operator T() const { assert(this->r == 1 && this->c == 1); return this->d[0]; }
};
template <class T> struct kind<const Prod<T>> { static const int v = 1; typedef T S; };

// ---------------------------------------------------------------------------------------------------------------------
// arrays (coefficient-wise semantics)
// ---------------------------------------------------------------------------------------------------------------------
template <class D, class T> struct ArrOps : DenseRO<D, T> {
  template <class F> Arr<T> map(F f) const;
  Arr<T> square() const { return map([](T x) { return x * x; }); }
  Arr<T> cube() const { return map([](T x) { return x * x * x; }); }
  Arr<T> sqrt() const { return map([](T x) { return (T)std::sqrt(x); }); }
  Arr<T> rsqrt() const { return map([](T x) { return T(1) / (T)std::sqrt(x); }); }
  Arr<T> inverse() const { return map([](T x) { return T(1) / x; }); }
  Arr<T> abs() const { return map([](T x) { return (T)std::abs(x); }); }
  Arr<T> abs2() const { return map([](T x) { return x * x; }); }
  Arr<T> log() const { return map([](T x) { return (T)std::log(x); }); }
  Arr<T> log10() const { return map([](T x) { return (T)std::log10(x); }); }
  Arr<T> exp() const { return map([](T x) { return (T)std::exp(x); }); }
  Arr<T> tanh() const { return map([](T x) { return (T)std::tanh(x); }); }
  template <class S> Arr<T> pow(S e) const { return map([e](T x) { return (T)std::pow(x, (T)e); }); }
  template <class S> Arr<T> max(S m) const { return map([m](T x) { return std::max(x, (T)m); }); }
  template <class S> Arr<T> min(S m) const { return map([m](T x) { return std::min(x, (T)m); }); }
  Arr<T> isNaN() const { return map([](T x) { return (T)(std::isnan(x) ? 1 : 0); }); }
};

template <class T> struct Arr : ArrOps<Arr<T>, T> {
  Index r, c;
  std::vector<T> d;
  Arr() : r(0), c(0) {}
  Arr(Index r_, Index c_) : r(r_), c(c_), d((size_t)r_ * c_) {}
  template <class X, SHIM_IF(kind<X>::v != 0)> Arr(const X& x) : r(x.rows()), c(x.cols()), d((size_t)x.rows() * x.cols()) {
    for (Index j = 0; j < c; j++) for (Index i = 0; i < r; i++) d[(size_t)j * r + i] = x(i, j);
  }
  Index rows() const { return r; }
  Index cols() const { return c; }
  T& operator()(Index i, Index j) { return d[(size_t)j * r + i]; }
  const T& operator()(Index i, Index j) const { return d[(size_t)j * r + i]; }
  T& operator()(Index i) { return d[i]; }
  const T& operator()(Index i) const { return d[i]; }
  T& operator[](Index i) { return d[i]; }
  const T& operator[](Index i) const { return d[i]; }
  Mat<T> matrix() const { Mat<T> m(r, c); m.d = d; return m; }
  const Arr& array() const { return *this; }
  Arr<T> transpose() const { Arr<T> t(c, r); for (Index j = 0; j < c; j++) for (Index i = 0; i < r; i++) t(j, i) = (*this)(i, j); return t; }
  AView<T> col(Index j) const;
  AView<T> row(Index i) const;
  template <class U> Arr<U> cast() const { Arr<U> m(r, c); for (size_t i = 0; i < d.size(); i++) m.d[i] = (U)d[i]; return m; }
};

template <class T> struct AView : ArrOps<AView<T>, T> {
  View<T> v;
  explicit AView(const View<T>& v_) : v(v_) {}
  AView(const AView&) = default;
  Index rows() const { return v.r; }
  Index cols() const { return v.c; }
  T& operator()(Index i, Index j) const { return v(i, j); }
  T& operator()(Index i) const { return v(i); }
  T& operator[](Index i) const { return v(i); }
  View<T> matrix() const { return v; }
  const AView& array() const { return *this; }
  AView transpose() const { return AView(v.transpose()); }
  AView col(Index j) const { return AView(v.col(j)); }
  AView row(Index i) const { return AView(v.row(i)); }
  const AView& operator=(const AView& o) const { Mat<T> t(o); v.assign_from(t); return *this; }
  template <class X, SHIM_IF(kind<X>::v != 0)> const AView& operator=(const X& x) const { Mat<T> t(x); v.assign_from(t); return *this; }
  template <class S, SHIM_IF(std::is_arithmetic<S>::value)> const AView& operator=(S s) const { v.setConstant((T)s); return *this; }
  template <class X, SHIM_IF(kind<X>::v != 0)> const AView& operator+=(const X& x) const { Mat<T> t(x); for (Index i = 0; i < this->size(); i++) v.lref(i) += t.lin(i); return *this; }
  template <class X, SHIM_IF(kind<X>::v != 0)> const AView& operator-=(const X& x) const { Mat<T> t(x); for (Index i = 0; i < this->size(); i++) v.lref(i) -= t.lin(i); return *this; }
  template <class X, SHIM_IF(kind<X>::v != 0)> const AView& operator*=(const X& x) const { Mat<T> t(x); for (Index i = 0; i < this->size(); i++) v.lref(i) *= t.lin(i); return *this; }
  template <class X, SHIM_IF(kind<X>::v != 0)> const AView& operator/=(const X& x) const { Mat<T> t(x); for (Index i = 0; i < this->size(); i++) v.lref(i) /= t.lin(i); return *this; }
  template <class S, SHIM_IF(std::is_arithmetic<S>::value)> const AView& operator+=(S s) const { for (Index i = 0; i < this->size(); i++) v.lref(i) += (T)s; return *this; }
  template <class S, SHIM_IF(std::is_arithmetic<S>::value)> const AView& operator-=(S s) const { for (Index i = 0; i < this->size(); i++) v.lref(i) -= (T)s; return *this; }
  template <class S, SHIM_IF(std::is_arithmetic<S>::value)> const AView& operator*=(S s) const { for (Index i = 0; i < this->size(); i++) v.lref(i) *= (T)s; return *this; }
  template <class S, SHIM_IF(std::is_arithmetic<S>::value)> const AView& operator/=(S s) const { for (Index i = 0; i < this->size(); i++) v.lref(i) /= (T)s; return *this; }
};
template <class T> AView<T> View<T>::array() const { return AView<T>(*this); }
template <class T> AView<T> Mat<T>::array() const { return AView<T>(v()); }
template <class T> AView<T> Arr<T>::col(Index j) const { return AView<T>(View<T>(const_cast<T*>(d.data()) + (size_t)j * r, r, 1, 1, r)); }
template <class T> AView<T> Arr<T>::row(Index i) const { return AView<T>(View<T>(const_cast<T*>(d.data()) + i, 1, c, 1, r)); }
template <class D, class T> template <class F> Arr<T> ArrOps<D, T>::map(F f) const {
  const D& s = this->self();
  Arr<T> a(s.rows(), s.cols());
  for (Index j = 0; j < s.cols(); j++) for (Index i = 0; i < s.rows(); i++) a(i, j) = f(s(i, j));
  return a;
}

template <class T> Mat<T> View<T>::cwiseAbs() const { return Mat<T>(*this).cwiseAbs(); }
template <class T> Mat<T> View<T>::cwiseAbs2() const { return Mat<T>(*this).cwiseAbs2(); }
template <class T> Mat<T> View<T>::cwiseInverse() const { return Mat<T>(*this).cwiseInverse(); }
template <class T> Mat<T> View<T>::cwiseSqrt() const { return Mat<T>(*this).cwiseSqrt(); }
template <class T> template <class O> Mat<T> View<T>::cwiseProduct(const O& o) const { return Mat<T>(*this).cwiseProduct(o); }
template <class T> template <class O> Mat<T> View<T>::cwiseQuotient(const O& o) const { return Mat<T>(*this).cwiseQuotient(o); }
template <class T> Mat<T> View<T>::asDiagonal() const { return Mat<T>(*this).asDiagonal(); }
template <class T> Mat<T> View<T>::inverse() const { return Mat<T>(*this).inverse(); }
template <class T> template <class U> Mat<U> View<T>::cast() const { return Mat<T>(*this).template cast<U>(); }

// ---------------------------------------------------------------------------------------------------------------------
// colwise / rowwise reductions (results are row / column vectors)
// ---------------------------------------------------------------------------------------------------------------------
template <class D, class T> struct ColwiseOp {
  const D& m;
  template <class F> Mat<T> red(F f) const { Mat<T> o(1, m.cols()); Mat<T> t(m); for (Index j = 0; j < m.cols(); j++) o(0, j) = f(t.col(j)); return o; }
  Mat<T> sum() const { return red([](const View<T>& c) { return c.sum(); }); }
  Mat<T> mean() const { return red([](const View<T>& c) { return c.mean(); }); }
  Mat<T> squaredNorm() const { return red([](const View<T>& c) { return c.squaredNorm(); }); }
  Mat<T> norm() const { return red([](const View<T>& c) { return c.norm(); }); }
  Mat<T> maxCoeff() const { return red([](const View<T>& c) { return c.maxCoeff(); }); }
  Mat<T> minCoeff() const { return red([](const View<T>& c) { return c.minCoeff(); }); }
};
template <class D, class T> struct RowwiseOp {
  const D& m;
  template <class F> Mat<T> red(F f) const { Mat<T> o(m.rows(), 1); Mat<T> t(m); for (Index i = 0; i < m.rows(); i++) o(i, 0) = f(t.row(i)); return o; }
  Mat<T> sum() const { return red([](const View<T>& c) { return c.sum(); }); }
  Mat<T> mean() const { return red([](const View<T>& c) { return c.mean(); }); }
  Mat<T> squaredNorm() const { return red([](const View<T>& c) { return c.squaredNorm(); }); }
  Mat<T> norm() const { return red([](const View<T>& c) { return c.norm(); }); }
  Mat<T> maxCoeff() const { return red([](const View<T>& c) { return c.maxCoeff(); }); }
  Mat<T> minCoeff() const { return red([](const View<T>& c) { return c.minCoeff(); }); }
  // X.rowwise() - v.transpose()
  template <class X, SHIM_IF(kind<X>::v == 1)> Mat<T> operator-(const X& x) const { Mat<T> t(m), o(t.r, t.c), w(x); for (Index j = 0; j < t.c; j++) for (Index i = 0; i < t.r; i++) o(i, j) = t(i, j) - w.lin(j); return o; }
  template <class X, SHIM_IF(kind<X>::v == 1)> Mat<T> operator+(const X& x) const { Mat<T> t(m), o(t.r, t.c), w(x); for (Index j = 0; j < t.c; j++) for (Index i = 0; i < t.r; i++) o(i, j) = t(i, j) + w.lin(j); return o; }
};

// ---------------------------------------------------------------------------------------------------------------------
// operators.  Vectors of equal length combine element by element whatever their orientation (Eigen transposes on assignment).
// ---------------------------------------------------------------------------------------------------------------------
template <class T, class A, class B, class F> inline void zip(const A& a, const B& b, T* out, F f) {
  if (a.rows() == b.rows() && a.cols() == b.cols()) { Index k = 0; for (Index j = 0; j < a.cols(); j++) for (Index i = 0; i < a.rows(); i++) out[k++] = f(a(i, j), b(i, j)); }
  else { assert(a.size() == b.size() && (a.rows() == 1 || a.cols() == 1) && (b.rows() == 1 || b.cols() == 1)); for (Index i = 0; i < a.size(); i++) out[i] = f(a.lin(i), b.lin(i)); }
}
#define SHIM_MM(op) \
  template <class A, class B, SHIM_IF(kind<A>::v == 1 && kind<B>::v == 1)> Mat<typename kind<A>::S> operator op(const A& a, const B& b) { \
    typedef typename kind<A>::S T; Mat<T> o(a.rows(), a.cols()); zip<T>(a, b, o.d.data(), [](T x, T y) { return x op y; }); return o; }
SHIM_MM(+)
SHIM_MM(-)
#define SHIM_AA(op) \
  template <class A, class B, SHIM_IF(kind<A>::v == 2 && kind<B>::v == 2)> Arr<typename kind<A>::S> operator op(const A& a, const B& b) { \
    typedef typename kind<A>::S T; Arr<T> o(a.rows(), a.cols()); zip<T>(a, b, o.d.data(), [](T x, T y) { return x op y; }); return o; } \
  template <class A, class S, SHIM_IF(kind<A>::v == 2 && std::is_arithmetic<S>::value)> Arr<typename kind<A>::S> operator op(const A& a, S s) { \
    typedef typename kind<A>::S T; const T t = (T)s; Arr<T> o(a.rows(), a.cols()); for (Index j = 0; j < a.cols(); j++) for (Index i = 0; i < a.rows(); i++) o(i, j) = a(i, j) op t; return o; } \
  template <class A, class S, SHIM_IF(kind<A>::v == 2 && std::is_arithmetic<S>::value)> Arr<typename kind<A>::S> operator op(S s, const A& a) { \
    typedef typename kind<A>::S T; const T t = (T)s; Arr<T> o(a.rows(), a.cols()); for (Index j = 0; j < a.cols(); j++) for (Index i = 0; i < a.rows(); i++) o(i, j) = t op a(i, j); return o; }
SHIM_AA(+)
SHIM_AA(-)
SHIM_AA(*)
SHIM_AA(/)
#define SHIM_ACMP(op) \
  template <class A, class B, SHIM_IF(kind<A>::v == 2 && kind<B>::v == 2)> Arr<typename kind<A>::S> operator op(const A& a, const B& b) { \
    typedef typename kind<A>::S T; Arr<T> o(a.rows(), a.cols()); zip<T>(a, b, o.d.data(), [](T x, T y) { return (T)(x op y ? 1 : 0); }); return o; } \
  template <class A, class S, SHIM_IF(kind<A>::v == 2 && std::is_arithmetic<S>::value)> Arr<typename kind<A>::S> operator op(const A& a, S s) { \
    typedef typename kind<A>::S T; const T t = (T)s; Arr<T> o(a.rows(), a.cols()); for (Index j = 0; j < a.cols(); j++) for (Index i = 0; i < a.rows(); i++) o(i, j) = (T)(a(i, j) op t ? 1 : 0); return o; }
SHIM_ACMP(>)
SHIM_ACMP(<)
SHIM_ACMP(>=)
SHIM_ACMP(<=)
template <class A, SHIM_IF(kind<A>::v == 2)> Arr<typename kind<A>::S> operator-(const A& a) { typedef typename kind<A>::S T; return Arr<T>(a).map([](T x) { return -x; }); }
template <class A, SHIM_IF(kind<A>::v == 1)> Mat<typename kind<A>::S> operator-(const A& a) { typedef typename kind<A>::S T; Mat<T> o(a); for (auto& x : o.d) x = -x; return o; }
// matrix (op) scalar
template <class A, class S, SHIM_IF(kind<A>::v == 1 && std::is_arithmetic<S>::value)> Mat<typename kind<A>::S> operator*(const A& a, S s) {
  typedef typename kind<A>::S T; Mat<T> o(a); const T t = (T)s; for (auto& x : o.d) x *= t; return o; }
template <class A, class S, SHIM_IF(kind<A>::v == 1 && std::is_arithmetic<S>::value)> Mat<typename kind<A>::S> operator*(S s, const A& a) {
  typedef typename kind<A>::S T; Mat<T> o(a); const T t = (T)s; for (auto& x : o.d) x = t * x; return o; }
template <class A, class S, SHIM_IF(kind<A>::v == 1 && std::is_arithmetic<S>::value)> Mat<typename kind<A>::S> operator/(const A& a, S s) {
  typedef typename kind<A>::S T; Mat<T> o(a); const T t = (T)s; for (auto& x : o.d) x /= t; return o; }
// the hot spot of every solver: X.col(j) * (b1 - b0), kept lazy
template <class T, class S, SHIM_IF(std::is_arithmetic<S>::value)> Scaled<T> operator*(const View<T>& a, S s) { return Scaled<T>(a, (T)s); }
template <class T, class S, SHIM_IF(std::is_arithmetic<S>::value)> Scaled<T> operator*(S s, const View<T>& a) { return Scaled<T>(a, (T)s); }
// matrix product
template <class A, class B, SHIM_IF(kind<A>::v == 1 && kind<B>::v == 1)> Prod<typename kind<A>::S> operator*(const A& a_, const B& b_) {
  typedef typename kind<A>::S T;
  const Mat<T> a(a_), b(b_);
  assert(a.c == b.r);
  Prod<T> o(a.r, b.c);
  if (a.r == 1) { for (Index j = 0; j < b.c; j++) { const T* bj = b.d.data() + (size_t)j * b.r; o.d[j] = reduce_sum<T>(a.c, [&](Index k) { return a.d[k] * bj[k]; }); } return o; }
  for (Index j = 0; j < b.c; j++)
    for (Index k = 0; k < a.c; k++) { const T bkj = b(k, j); const T* ak = a.d.data() + (size_t)k * a.r; T* oj = o.d.data() + (size_t)j * a.r; for (Index i = 0; i < a.r; i++) oj[i] += ak[i] * bkj; }
  return o;
}

// ---------------------------------------------------------------------------------------------------------------------
// dense solvers
// ---------------------------------------------------------------------------------------------------------------------
template <class T> Mat<T> Mat<T>::inverse() const {
  assert(r == c);
  const Index n = r;
  Mat<T> a(*this), inv = Mat<T>::Identity(n, n);
  for (Index k = 0; k < n; k++) {
    Index piv = k;
    for (Index i = k + 1; i < n; i++) if (std::abs(a(i, k)) > std::abs(a(piv, k))) piv = i;
    if (piv != k) for (Index j = 0; j < n; j++) { std::swap(a(k, j), a(piv, j)); std::swap(inv(k, j), inv(piv, j)); }
    const T dkk = a(k, k);
    for (Index j = 0; j < n; j++) { a(k, j) /= dkk; inv(k, j) /= dkk; }
    for (Index i = 0; i < n; i++) if (i != k) { const T f = a(i, k); if (f != T(0)) for (Index j = 0; j < n; j++) { a(i, j) -= f * a(k, j); inv(i, j) -= f * inv(k, j); } }
  }
  return inv;
}
template <class T> T Mat<T>::determinant() const {
  Mat<T> a(*this); T det = 1;
  for (Index k = 0; k < r; k++) {
    Index piv = k;
    for (Index i = k + 1; i < r; i++) if (std::abs(a(i, k)) > std::abs(a(piv, k))) piv = i;
    if (piv != k) { for (Index j = 0; j < r; j++) std::swap(a(k, j), a(piv, j)); det = -det; }
    det *= a(k, k);
    if (a(k, k) == T(0)) return 0;
    for (Index i = k + 1; i < r; i++) { const T f = a(i, k) / a(k, k); for (Index j = k; j < r; j++) a(i, j) -= f * a(k, j); }
  }
  return det;
}
template <class T> struct LLT {
  Mat<T> L; ComputationInfo inf = Success;
  explicit LLT(const Mat<T>& A) { compute(A); }
  void compute(const Mat<T>& A) {
    const Index n = A.r; L = Mat<T>(n, n); inf = Success;
    for (Index j = 0; j < n; j++) {
      T s = A(j, j);
      for (Index k = 0; k < j; k++) s -= L(j, k) * L(j, k);
      if (!(s > T(0))) { inf = NumericalIssue; return; }
      const T ljj = std::sqrt(s); L(j, j) = ljj;
      for (Index i = j + 1; i < n; i++) { T t = A(i, j); for (Index k = 0; k < j; k++) t -= L(i, k) * L(j, k); L(i, j) = t / ljj; }
    }
  }
  ComputationInfo info() const { return inf; }
  template <class B> Mat<T> solve(const B& b_) const {
    Mat<T> x(b_); const Index n = L.r; bool flip = false;
    if (x.r != n && x.c == n && x.r == 1) { x.r = n; x.c = 1; flip = true; }
    (void)flip;
    for (Index col = 0; col < x.c; col++) {
      for (Index i = 0; i < n; i++) { T t = x(i, col); for (Index k = 0; k < i; k++) t -= L(i, k) * x(k, col); x(i, col) = t / L(i, i); }
      for (Index i = n - 1; i >= 0; i--) { T t = x(i, col); for (Index k = i + 1; k < n; k++) t -= L(k, i) * x(k, col); x(i, col) = t / L(i, i); }
    }
    return x;
  }
  const Mat<T>& matrixL() const { return L; }
};
template <class T> LLT<T> Mat<T>::llt() const { return LLT<T>(*this); }
template <class T> LLT<T> View<T>::llt() const { return LLT<T>(Mat<T>(*this)); }

// symmetric eigen-decomposition, cyclic Jacobi; eigenvalues ascending, eigenvectors in columns
template <class M> struct SelfAdjointEigenSolver {
  typedef typename M::Scalar T;
  Vec<T> vals; Mat<T> vecs; ComputationInfo inf = Success;
  SelfAdjointEigenSolver() {}
  template <class X> explicit SelfAdjointEigenSolver(const X& A) { compute(A); }
  template <class X> SelfAdjointEigenSolver& compute(const X& A_) {
    Mat<T> A(A_); const Index n = A.r;
    Mat<T> V = Mat<T>::Identity(n, n);
    for (int sweep = 0; sweep < 100; sweep++) {
      T off = 0, dia = 0;
      for (Index i = 0; i < n; i++) { dia += A(i, i) * A(i, i); for (Index j = 0; j < i; j++) off += A(i, j) * A(i, j); }
      if (!(off > std::numeric_limits<T>::epsilon() * std::numeric_limits<T>::epsilon() * (dia + off))) break;
      for (Index p = 0; p < n - 1; p++)
        for (Index q = p + 1; q < n; q++) {
          const T apq = A(p, q);
          if (apq == T(0)) continue;
          const T theta = (A(q, q) - A(p, p)) / (2 * apq);
          const T t = (theta >= 0 ? T(1) : T(-1)) / (std::abs(theta) + std::sqrt(theta * theta + 1));
          const T cs = 1 / std::sqrt(t * t + 1), sn = t * cs;
          for (Index k = 0; k < n; k++) { const T akp = A(k, p), akq = A(k, q); A(k, p) = cs * akp - sn * akq; A(k, q) = sn * akp + cs * akq; }
          for (Index k = 0; k < n; k++) { const T apk = A(p, k), aqk = A(q, k); A(p, k) = cs * apk - sn * aqk; A(q, k) = sn * apk + cs * aqk; }
          for (Index k = 0; k < n; k++) { const T vkp = V(k, p), vkq = V(k, q); V(k, p) = cs * vkp - sn * vkq; V(k, q) = sn * vkp + cs * vkq; }
        }
    }
    std::vector<Index> idx(n);
    for (Index i = 0; i < n; i++) idx[i] = i;
    std::sort(idx.begin(), idx.end(), [&](Index a, Index b) { return A(a, a) < A(b, b); });
    vals = Vec<T>(n); vecs = Mat<T>(n, n);
    for (Index k = 0; k < n; k++) { vals[k] = A(idx[k], idx[k]); for (Index i = 0; i < n; i++) vecs(i, k) = V(i, idx[k]); }
    return *this;
  }
  const Vec<T>& eigenvalues() const { return vals; }
  const Mat<T>& eigenvectors() const { return vecs; }
  ComputationInfo info() const { return inf; }
};

// one-sided Jacobi SVD (thin), singular values descending
template <class T> inline void jacobi_svd(const Mat<T>& A, Mat<T>& U, Vec<T>& S, Mat<T>& V) {
  const Index m = A.r, n = A.c;
  if (m < n) { Mat<T> At(A.transpose()); jacobi_svd(At, V, S, U); return; }
  Mat<T> W(A); V = Mat<T>::Identity(n, n);
  for (int sweep = 0; sweep < 60; sweep++) {
    bool rotated = false;
    for (Index p = 0; p < n - 1; p++)
      for (Index q = p + 1; q < n; q++) {
        T a = 0, b = 0, g = 0;
        for (Index i = 0; i < m; i++) { a += W(i, p) * W(i, p); b += W(i, q) * W(i, q); g += W(i, p) * W(i, q); }
        if (std::abs(g) <= std::numeric_limits<T>::epsilon() * std::sqrt(a * b) || g == T(0)) continue;
        rotated = true;
        const T zeta = (b - a) / (2 * g);
        const T t = (zeta >= 0 ? T(1) : T(-1)) / (std::abs(zeta) + std::sqrt(1 + zeta * zeta));
        const T cs = 1 / std::sqrt(1 + t * t), sn = cs * t;
        for (Index i = 0; i < m; i++) { const T wp = W(i, p), wq = W(i, q); W(i, p) = cs * wp - sn * wq; W(i, q) = sn * wp + cs * wq; }
        for (Index i = 0; i < n; i++) { const T vp = V(i, p), vq = V(i, q); V(i, p) = cs * vp - sn * vq; V(i, q) = sn * vp + cs * vq; }
      }
    if (!rotated) break;
  }
  std::vector<T> sv(n);
  for (Index j = 0; j < n; j++) sv[j] = W.col(j).norm();
  std::vector<Index> idx(n);
  for (Index i = 0; i < n; i++) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&](Index a, Index b) { return sv[a] > sv[b]; });
  U = Mat<T>(m, n); S = Vec<T>(n); Mat<T> V2(n, n);
  for (Index k = 0; k < n; k++) {
    const Index j = idx[k]; S[k] = sv[j];
    for (Index i = 0; i < m; i++) U(i, k) = sv[j] > T(0) ? W(i, j) / sv[j] : T(0);
    for (Index i = 0; i < n; i++) V2(i, k) = V(i, j);
  }
  V = V2;
}
template <class M> struct SVDBase {
  typedef typename M::Scalar T;
  Mat<T> U, V; Vec<T> S;
  SVDBase() {}
  template <class X> SVDBase(const X& A, unsigned = 0) { compute(A); }
  template <class X> SVDBase& compute(const X& A, unsigned = 0) { jacobi_svd<T>(Mat<T>(A), U, S, V); return *this; }
  const Mat<T>& matrixU() const { return U; }
  const Mat<T>& matrixV() const { return V; }
  const Vec<T>& singularValues() const { return S; }
  template <class B> Mat<T> solve(const B& b) const { Mat<T> t = U.transpose() * Mat<T>(b); for (Index j = 0; j < t.c; j++) for (Index i = 0; i < t.r; i++) t(i, j) = S[i] > S[0] * std::numeric_limits<T>::epsilon() * (T)std::max(U.r, V.r) ? t(i, j) / S[i] : T(0); return V * t; }
};
template <class M> struct BDCSVD : SVDBase<M> { using SVDBase<M>::SVDBase; };
template <class M> struct JacobiSVD : SVDBase<M> { using SVDBase<M>::SVDBase; };
// completeOrthogonalDecomposition().pseudoInverse(): V S^+ U' with Eigen's rank threshold (epsilon * min(rows, cols) * s_max)
template <class T> struct COD {
  Mat<T> A;
  explicit COD(const Mat<T>& A_) : A(A_) {}
  Mat<T> pseudoInverse() const {
    Mat<T> U, V; Vec<T> S;
    jacobi_svd<T>(A, U, S, V);
    const T thr = std::numeric_limits<T>::epsilon() * (T)std::min(A.r, A.c) * (S.size() ? S[0] : T(0));
    Mat<T> Vs(V);
    for (Index k = 0; k < S.size(); k++) { const T inv = S[k] > thr ? T(1) / S[k] : T(0); for (Index i = 0; i < V.r; i++) Vs(i, k) = V(i, k) * inv; }
    return Vs * U.transpose();
  }
  template <class B> Mat<T> solve(const B& b) const { return pseudoInverse() * Mat<T>(b); }
};
template <class T> COD<T> Mat<T>::completeOrthogonalDecomposition() const { return COD<T>(*this); }
template <class T> COD<T> View<T>::completeOrthogonalDecomposition() const { return COD<T>(Mat<T>(*this)); }

typedef Mat<float> MatrixXf;
typedef Mat<double> MatrixXd;
typedef Mat<int> MatrixXi;
typedef Vec<float> VectorXf;
typedef Vec<double> VectorXd;
typedef Vec<int> VectorXi;
typedef Arr<float> ArrayXXf;
typedef Arr<double> ArrayXXd;

template <class A, SHIM_IF(kind<A>::v != 0)> std::ostream& operator<<(std::ostream& os, const A& a) {
  for (Index i = 0; i < a.rows(); i++) { for (Index j = 0; j < a.cols(); j++) os << a(i, j) << ' '; os << '\n'; }
  return os;
}

}  // namespace Eigen

// =====================================================================================================================
// Rcpp / R
// =====================================================================================================================
namespace Rcpp {

struct Value {
  std::vector<double> v;
  long r = -1, c = -1;  // -1/-1: scalar
};
template <class X, typename std::enable_if<std::is_arithmetic<X>::value, int>::type = 0> inline Value to_value(const X& x) { Value o; o.v.push_back((double)x); return o; }
template <class X, typename std::enable_if<Eigen::kind<X>::v != 0, int>::type = 0> inline Value to_value(const X& x) {
  Value o; o.r = (long)x.rows(); o.c = (long)x.cols();
  for (long j = 0; j < o.c; j++) for (long i = 0; i < o.r; i++) o.v.push_back((double)x(i, j));
  return o;
}
template <class T> inline Value to_value(const std::vector<T>& x) { Value o; o.r = (long)x.size(); o.c = 1; for (const T& t : x) o.v.push_back((double)t); return o; }

struct NamedArg { std::string name; Value val; };
struct Named {
  std::string n;
  explicit Named(const char* s) : n(s) {}
  explicit Named(const std::string& s) : n(s) {}
  template <class X> NamedArg operator=(const X& x) const { NamedArg a; a.name = n; a.val = to_value(x); return a; }
};
struct List {
  std::vector<NamedArg> items;
  template <class... A> static List create(const A&... a) { List l; (l.items.push_back(a), ...); return l; }
  operator SEXP() const { return (SEXP) new List(*this); }
  const Value* get(const char* name) const { for (const NamedArg& a : items) if (a.name == name) return &a.val; return nullptr; }
};

class NumericVector {
 public:
  std::vector<double> d;
  NumericVector() {}
  explicit NumericVector(int n) : d((size_t)n, 0.0) {}
  NumericVector(const double* p, size_t n) : d(p, p + n) {}
  double& operator[](long i) { return d[(size_t)i]; }
  const double& operator[](long i) const { return d[(size_t)i]; }
  double& operator()(long i) { return d[(size_t)i]; }
  long size() const { return (long)d.size(); }
  long length() const { return (long)d.size(); }
  std::vector<double>::iterator begin() { return d.begin(); }
  std::vector<double>::iterator end() { return d.end(); }
};
inline Value to_value(const NumericVector& x) { return to_value(x.d); }
template <class T> struct Nullable {
  std::shared_ptr<T> p;
  Nullable() {}
  Nullable(SEXP s) { (void)s; }  // only ever R_NilValue
  Nullable(const T& t) : p(new T(t)) {}
  bool isNotNull() const { return (bool)p; }
  bool isNull() const { return !p; }
  operator T() const { return *p; }
  T get() const { return *p; }
};
inline NumericVector operator-(double a, const NumericVector& x) { NumericVector o(x); for (double& v : o.d) v = a - v; return o; }
inline NumericVector operator-(const NumericVector& x) { NumericVector o(x); for (double& v : o.d) v = -v; return o; }
inline NumericVector log10(const NumericVector& x) { NumericVector o(x); for (double& v : o.d) v = std::log10(v); return o; }
// regularised lower incomplete gamma P(a, x) (series / continued fraction), for pchisq
inline double gamma_p(double a, double x) {
  if (!(x > 0)) return 0.0;
  const double gln = std::lgamma(a);
  if (x < a + 1) { double ap = a, sum = 1.0 / a, del = sum; for (int n = 0; n < 1000; n++) { ap += 1; del *= x / ap; sum += del; if (std::fabs(del) < std::fabs(sum) * 1e-16) break; } return sum * std::exp(-x + a * std::log(x) - gln); }
  double b = x + 1 - a, c = 1e300, dd = 1 / b, h = dd;
  for (int i = 1; i < 1000; i++) { const double an = -i * (i - a); b += 2; dd = an * dd + b; if (std::fabs(dd) < 1e-300) dd = 1e-300; c = b + an / c; if (std::fabs(c) < 1e-300) c = 1e-300; dd = 1 / dd; const double del = dd * c; h *= del; if (std::fabs(del - 1) < 1e-16) break; }
  return 1.0 - std::exp(-x + a * std::log(x) - gln) * h;
}
inline NumericVector pchisq(const NumericVector& x, double df, bool lower = true, bool lg = false) {
  NumericVector o(x);
  for (double& v : o.d) { double pr = gamma_p(0.5 * df, 0.5 * v); if (!lower) pr = 1 - pr; v = lg ? std::log(pr) : pr; }
  return o;
}
struct NullStream { template <class X> NullStream& operator<<(const X&) { return *this; } };
static NullStream Rcout;
static NullStream Rcerr;
inline void checkUserInterrupt() {}

}  // namespace Rcpp

// R's RNG entry points on one seedable generator (R's own Mersenne-Twister + inversion stream cannot be reproduced here)
namespace R {
inline std::mt19937_64& engine() { static std::mt19937_64 e(1); return e; }
inline void set_seed(uint64_t s) { engine().seed(s); }
inline double rnorm(double mu, double sd) { std::normal_distribution<double> d(0.0, 1.0); return mu + sd * d(engine()); }
inline double runif(double a, double b) { std::uniform_real_distribution<double> d(a, b); return d(engine()); }
inline double rchisq(double df) { std::chi_squared_distribution<double> d(df); return d(engine()); }
inline double rgamma(double shape, double scale) { std::gamma_distribution<double> d(shape, scale); return d(engine()); }
inline double rbinom(double n, double p) {
  if (!(p >= 0.0 && p <= 1.0)) return std::numeric_limits<double>::quiet_NaN();  // R: NaN (with a warning)
  if (n == 1.0) { std::uniform_real_distribution<double> u(0.0, 1.0); return u(engine()) < p ? 1.0 : 0.0; }  // one Bernoulli draw
  std::binomial_distribution<long> d((long)n, p);
  return (double)d(engine());
}
inline double rbeta(double a, double b) { const double x = rgamma(a, 1.0), y = rgamma(b, 1.0); return x / (x + y); }
inline double rexp(double scale) { std::exponential_distribution<double> d(1.0 / scale); return d(engine()); }
}  // namespace R
