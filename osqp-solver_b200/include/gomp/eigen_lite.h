// eigen_lite.h -- the sliver of the Eigen API that the GOMP host facade and its callers use.
//
// The reference expresses its public surface in Eigen types (reference src/utils.h:12-22,
// src/constraints/constraint-builder.h:16).  Eigen is not installed on the build box, so when
// <Eigen/Sparse> is absent this header supplies the handful of classes the drop-in surface
// needs (SURVEY.md 8b lists them).  With real Eigen on the include path it is skipped and the
// facade compiles against Eigen itself: the facade only uses members present in both.
#ifndef GOMP_EIGEN_LITE_H
#define GOMP_EIGEN_LITE_H

#if defined(__has_include) && !defined(GOMP_FORCE_EIGEN_LITE)
#if __has_include(<Eigen/Sparse>)
#define GOMP_HAVE_EIGEN 1
#endif
#endif

#ifdef GOMP_HAVE_EIGEN
#include <Eigen/Core>
#include <Eigen/Sparse>
#else

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <initializer_list>
#include <type_traits>
#include <vector>

namespace Eigen {

enum { ColMajor = 0, RowMajor = 1 };
typedef std::ptrdiff_t Index;

class VectorXd;

// ---------------------------------------------------------------- fixed-size column vector
template <typename T, int N>
class Vector {
public:
    Vector() { for (int i = 0; i < N; ++i) v_[i] = T(0); }
    Vector(std::initializer_list<T> il) {
        assert((int)il.size() == N);
        int i = 0;
        for (T x : il) v_[i++] = x;
    }
    template <typename... Rest, typename = typename std::enable_if<sizeof...(Rest) + 2 == N>::type>
    Vector(T a, T b, Rest... rest) {
        T tmp[] = {a, b, static_cast<T>(rest)...};
        for (int i = 0; i < N; ++i) v_[i] = tmp[i];
    }
    inline Vector(const VectorXd &d);

    T &operator[](Index i) { return v_[i]; }
    const T &operator[](Index i) const { return v_[i]; }
    T &operator()(Index i) { return v_[i]; }
    const T &operator()(Index i) const { return v_[i]; }
    T *data() { return v_; }
    const T *data() const { return v_; }
    Index size() const { return N; }
    Vector &setConstant(T c) { for (int i = 0; i < N; ++i) v_[i] = c; return *this; }
    Vector &setZero() { return setConstant(T(0)); }
    static Vector Zero() { return Vector(); }
    static Vector Constant(T c) { Vector r; r.setConstant(c); return r; }

    Vector operator+(const Vector &o) const { Vector r; for (int i = 0; i < N; ++i) r.v_[i] = v_[i] + o.v_[i]; return r; }
    Vector operator-(const Vector &o) const { Vector r; for (int i = 0; i < N; ++i) r.v_[i] = v_[i] - o.v_[i]; return r; }
    Vector operator-() const { Vector r; for (int i = 0; i < N; ++i) r.v_[i] = -v_[i]; return r; }
    Vector operator*(T c) const { Vector r; for (int i = 0; i < N; ++i) r.v_[i] = v_[i] * c; return r; }
    Vector operator/(T c) const { Vector r; for (int i = 0; i < N; ++i) r.v_[i] = v_[i] / c; return r; }
    Vector &operator+=(const Vector &o) { for (int i = 0; i < N; ++i) v_[i] += o.v_[i]; return *this; }
    Vector &operator-=(const Vector &o) { for (int i = 0; i < N; ++i) v_[i] -= o.v_[i]; return *this; }
    Vector &operator*=(T c) { for (int i = 0; i < N; ++i) v_[i] *= c; return *this; }
    Vector &operator/=(T c) { for (int i = 0; i < N; ++i) v_[i] /= c; return *this; }
    T dot(const Vector &o) const { T s = T(0); for (int i = 0; i < N; ++i) s += v_[i] * o.v_[i]; return s; }
    T squaredNorm() const { return dot(*this); }
    T norm() const { return std::sqrt(squaredNorm()); }
    bool isApprox(const Vector &o, T prec = T(1e-12)) const {
        return (*this - o).squaredNorm() <= prec * prec * std::min(squaredNorm(), o.squaredNorm());
    }

private:
    T v_[N];
};
template <typename T, int N>
inline Vector<T, N> operator*(T c, const Vector<T, N> &v) { return v * c; }

typedef Vector<double, 2> Vector2d;
typedef Vector<double, 3> Vector3d;

// ---------------------------------------------------------------- dynamic column vector
class VectorXd {
public:
    // `seg = fixed/dynamic vector` and `VectorXd v = seg`
    class Segment {
    public:
        Segment(double *p, Index n) : p_(p), n_(n) {}
        template <int N>
        Segment &operator=(const Vector<double, N> &v) { assert(N == n_); for (int i = 0; i < N; ++i) p_[i] = v[i]; return *this; }
        inline Segment &operator=(const VectorXd &v);
        Segment &operator=(const Segment &o) { assert(o.n_ == n_); std::copy(o.p_, o.p_ + n_, p_); return *this; }
        Segment &operator/=(double c) { for (Index i = 0; i < n_; ++i) p_[i] /= c; return *this; }
        Segment &operator*=(double c) { for (Index i = 0; i < n_; ++i) p_[i] *= c; return *this; }
        double &operator[](Index i) { return p_[i]; }
        double operator[](Index i) const { return p_[i]; }
        Index size() const { return n_; }
        const double *data() const { return p_; }
        inline bool isApprox(const VectorXd &o, double prec = 1e-12) const;
    private:
        double *p_;
        Index n_;
    };
    // `v << a, b, c;` with scalars, fixed vectors, dynamic vectors or segments
    class CommaInit {
    public:
        CommaInit(VectorXd &v) : v_(v), at_(0) {}
        CommaInit &operator,(double x) { v_[at_++] = x; return *this; }
        CommaInit &operator,(const VectorXd &o) { return put(o.data(), o.size()); }
        CommaInit &operator,(const Segment &o) { return put(o.data(), o.size()); }
        template <int N>
        CommaInit &operator,(const Vector<double, N> &o) { return put(o.data(), N); }
        CommaInit &put(const double *p, Index n) { for (Index i = 0; i < n; ++i) v_[at_++] = p[i]; return *this; }
        ~CommaInit() { assert(at_ == v_.size()); }
    private:
        VectorXd &v_;
        Index at_;
    };

    VectorXd() {}
    explicit VectorXd(Index n) : v_((size_t)n, 0.0) {}
    VectorXd(const Segment &s) : v_(s.data(), s.data() + s.size()) {}
    template <int N>
    VectorXd(const Vector<double, N> &f) : v_(f.data(), f.data() + N) {}

    double &operator[](Index i) { return v_[(size_t)i]; }
    const double &operator[](Index i) const { return v_[(size_t)i]; }
    double &operator()(Index i) { return v_[(size_t)i]; }
    const double &operator()(Index i) const { return v_[(size_t)i]; }
    double *data() { return v_.data(); }
    const double *data() const { return v_.data(); }
    Index size() const { return (Index)v_.size(); }
    Index rows() const { return size(); }
    void resize(Index n) { v_.resize((size_t)n); }
    VectorXd &setZero(Index n) { v_.assign((size_t)n, 0.0); return *this; }
    VectorXd &setZero() { std::fill(v_.begin(), v_.end(), 0.0); return *this; }
    VectorXd &setConstant(double c) { std::fill(v_.begin(), v_.end(), c); return *this; }
    static VectorXd Zero(Index n) { return VectorXd(n); }
    static VectorXd Constant(Index n, double c) { VectorXd r(n); r.setConstant(c); return r; }

    Segment segment(Index start, Index len) { assert(start + len <= size()); return Segment(data() + start, len); }
    const Segment segment(Index start, Index len) const {
        assert(start + len <= size());
        return Segment(const_cast<double *>(data()) + start, len);
    }
    Segment head(Index len) { return segment(0, len); }
    Segment tail(Index len) { return segment(size() - len, len); }

    CommaInit operator<<(double x) { CommaInit c(*this); c, x; return c; }
    CommaInit operator<<(const VectorXd &o) { CommaInit c(*this); c, o; return c; }
    CommaInit operator<<(const Segment &o) { CommaInit c(*this); c, o; return c; }
    template <int N>
    CommaInit operator<<(const Vector<double, N> &o) { CommaInit c(*this); c, o; return c; }

    VectorXd operator+(const VectorXd &o) const { VectorXd r(size()); for (Index i = 0; i < size(); ++i) r[i] = v_[i] + o[i]; return r; }
    VectorXd operator-(const VectorXd &o) const { VectorXd r(size()); for (Index i = 0; i < size(); ++i) r[i] = v_[i] - o[i]; return r; }
    VectorXd operator*(double c) const { VectorXd r(size()); for (Index i = 0; i < size(); ++i) r[i] = v_[i] * c; return r; }
    double dot(const VectorXd &o) const { double s = 0; for (Index i = 0; i < size(); ++i) s += v_[i] * o[i]; return s; }
    double squaredNorm() const { return dot(*this); }
    double norm() const { return std::sqrt(squaredNorm()); }
    double lpNormInf() const { double r = 0; for (double x : v_) r = std::max(r, std::fabs(x)); return r; }
    bool isApprox(const VectorXd &o, double prec = 1e-12) const {
        if (o.size() != size()) return false;
        return (*this - o).squaredNorm() <= prec * prec * std::min(squaredNorm(), o.squaredNorm());
    }

private:
    std::vector<double> v_;
};

inline VectorXd::Segment &VectorXd::Segment::operator=(const VectorXd &v) {
    assert(v.size() == n_);
    std::copy(v.data(), v.data() + n_, p_);
    return *this;
}
inline bool VectorXd::Segment::isApprox(const VectorXd &o, double prec) const { return VectorXd(*this).isApprox(o, prec); }
template <typename T, int N>
inline Vector<T, N>::Vector(const VectorXd &d) {
    assert(d.size() == N);
    for (int i = 0; i < N; ++i) v_[i] = d[i];
}

// `Map<VectorXd>(ptr, n)`: a view that converts to an owning VectorXd.
template <typename V>
class Map;
template <>
class Map<VectorXd> {
public:
    Map(double *p, Index n) : p_(p), n_(n) {}
    Map(const double *p, Index n) : p_(const_cast<double *>(p)), n_(n) {}
    operator VectorXd() const { VectorXd r(n_); std::copy(p_, p_ + n_, r.data()); return r; }
    double &operator[](Index i) { return p_[i]; }
    Index size() const { return n_; }
private:
    double *p_;
    Index n_;
};

// ---------------------------------------------------------------- small dense fixed matrix
template <typename T, int R, int C, int Opt = ColMajor>
class Matrix {
public:
    class ConstRow {
    public:
        ConstRow(const Matrix &m, Index r) : m_(m), r_(r) {}
        T operator*(const Vector<T, C> &v) const { T s = T(0); for (int j = 0; j < C; ++j) s += m_(r_, j) * v[j]; return s; }
        T operator[](Index j) const { return m_(r_, j); }
    private:
        const Matrix &m_;
        Index r_;
    };
    Matrix() { for (int i = 0; i < R * C; ++i) v_[i] = T(0); }
    T &operator()(Index i, Index j) { return v_[Opt == RowMajor ? i * C + j : j * R + i]; }
    const T &operator()(Index i, Index j) const { return v_[Opt == RowMajor ? i * C + j : j * R + i]; }
    T *data() { return v_; }
    const T *data() const { return v_; }
    Index rows() const { return R; }
    Index cols() const { return C; }
    ConstRow row(Index r) const { return ConstRow(*this, r); }
    Matrix &setZero() { for (int i = 0; i < R * C; ++i) v_[i] = T(0); return *this; }
private:
    T v_[R * C];
};

// ---------------------------------------------------------------- sparse (CSC) matrix
template <typename T, typename StorageIndex = int>
class Triplet {
public:
    Triplet() : r_(0), c_(0), v_(0) {}
    Triplet(StorageIndex r, StorageIndex c, T v = T(0)) : r_(r), c_(c), v_(v) {}
    StorageIndex row() const { return r_; }
    StorageIndex col() const { return c_; }
    T value() const { return v_; }
private:
    StorageIndex r_, c_;
    T v_;
};

template <typename T, int Opt = ColMajor, typename StorageIndex = int>
class SparseMatrix {
    static_assert(Opt == ColMajor, "eigen_lite only provides column-major sparse storage");
public:
    SparseMatrix() : rows_(0), cols_(0), outer_(1, 0) {}
    SparseMatrix(Index r, Index c) : rows_(r), cols_(c), outer_((size_t)c + 1, 0) {}
    void resize(Index r, Index c) { rows_ = r; cols_ = c; outer_.assign((size_t)c + 1, 0); inner_.clear(); val_.clear(); }
    Index rows() const { return rows_; }
    Index cols() const { return cols_; }
    Index nonZeros() const { return (Index)val_.size(); }
    Index outerSize() const { return cols_; }
    bool isCompressed() const { return true; }
    void makeCompressed() {}
    const StorageIndex *outerIndexPtr() const { return outer_.data(); }
    const StorageIndex *innerIndexPtr() const { return inner_.data(); }
    const T *valuePtr() const { return val_.data(); }
    StorageIndex *outerIndexPtr() { return outer_.data(); }
    StorageIndex *innerIndexPtr() { return inner_.data(); }
    T *valuePtr() { return val_.data(); }

    // Duplicates are combined with `dup(old, new)`; explicit zeros are kept (as Eigen does).
    template <typename It, typename Dup>
    void setFromTriplets(It first, It last, Dup dup) {
        struct E { StorageIndex r, c; T v; size_t seq; };
        std::vector<E> e;
        size_t seq = 0;
        for (It it = first; it != last; ++it) e.push_back(E{(StorageIndex)it->row(), (StorageIndex)it->col(), (T)it->value(), seq++});
        std::stable_sort(e.begin(), e.end(), [](const E &a, const E &b) { return a.c != b.c ? a.c < b.c : a.r < b.r; });
        outer_.assign((size_t)cols_ + 1, 0);
        inner_.clear();
        val_.clear();
        for (size_t k = 0; k < e.size(); ++k) {
            assert(e[k].r >= 0 && (Index)e[k].r < rows_ && e[k].c >= 0 && (Index)e[k].c < cols_);
            if (k > 0 && e[k].r == e[k - 1].r && e[k].c == e[k - 1].c) {
                val_.back() = dup(val_.back(), e[k].v);
            } else {
                inner_.push_back(e[k].r);
                val_.push_back(e[k].v);
                outer_[(size_t)e[k].c + 1]++;
            }
        }
        for (Index j = 0; j < cols_; ++j) outer_[(size_t)j + 1] += outer_[(size_t)j];
    }
    template <typename It>
    void setFromTriplets(It first, It last) { setFromTriplets(first, last, [](const T &a, const T &b) { return a + b; }); }

    T coeff(Index i, Index j) const {
        for (StorageIndex k = outer_[(size_t)j]; k < outer_[(size_t)j + 1]; ++k)
            if ((Index)inner_[(size_t)k] == i) return val_[(size_t)k];
        return T(0);
    }

private:
    Index rows_, cols_;
    std::vector<StorageIndex> outer_, inner_;
    std::vector<T> val_;
};

} // namespace Eigen
#endif // !GOMP_HAVE_EIGEN
#endif // GOMP_EIGEN_LITE_H
