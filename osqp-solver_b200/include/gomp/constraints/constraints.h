// constraints.h -- box-constraint vocabulary of the GOMP facade.
//
// Drop-in for reference src/constraints/constraints.h:9-69 (same names, same meaning):
// an N-dimensional Bound is an optional vector, a Constraint is a (lower, upper) pair of
// Bounds, an absent side means "unbounded" and is written into the QP as -/+INF.
#ifndef GOMP_CONSTRAINTS_H
#define GOMP_CONSTRAINTS_H

#include <cstddef>
#include <optional>
#include <utility>

#include "../eigen_lite.h"

namespace constraints {

// Value the builder writes for a missing side.  Equal to OSQP_INFTY (reference
// constraints.h:11), so OSQP classifies such rows as loose.
const double INF = 1e30;

template <size_t N>
using Bound = std::optional<Eigen::Vector<double, N>>;

template <size_t N>
using Constraint = std::pair<Bound<N>, Bound<N>>;

// N copies of `val`.
template <size_t N>
Eigen::Vector<double, N> of(double val) {
    Eigen::Vector<double, N> filled;
    for (size_t k = 0; k < N; ++k) filled[k] = val;
    return filled;
}

template <size_t N>
Constraint<N> inRange(Bound<N> low, Bound<N> upp) {
    return Constraint<N>(std::move(low), std::move(upp));
}

template <size_t N>
Constraint<N> equal(const Eigen::Vector<double, N> &vals) {
    return Constraint<N>(Bound<N>(vals), Bound<N>(vals));
}

template <size_t N>
Constraint<N> greaterEq(const Eigen::Vector<double, N> &vals) {
    return Constraint<N>(Bound<N>(vals), Bound<N>());
}

template <size_t N>
Constraint<N> lessEq(const Eigen::Vector<double, N> &vals) {
    return Constraint<N>(Bound<N>(), Bound<N>(vals));
}

// No restriction at all.
template <size_t N>
const Constraint<N> ANY = Constraint<N>(Bound<N>(), Bound<N>());

// All components pinned to zero.
template <size_t N>
const Constraint<N> EQ_ZERO = equal<N>(of<N>(0.0));

// Multiply whichever sides are present by `factor` (used to turn per-second limits into
// per-step limits, reference src/gomp-solver.h:29-30).
template <size_t N>
const Bound<N> scaled(const Bound<N> &b, double factor) {
    if (!b.has_value()) return Bound<N>();
    Eigen::Vector<double, N> r = *b;
    for (size_t k = 0; k < N; ++k) r[k] *= factor;
    return Bound<N>(r);
}

template <size_t N>
const Constraint<N> scaled(const Constraint<N> &c, double factor) {
    return Constraint<N>(scaled<N>(c.first, factor), scaled<N>(c.second, factor));
}

} // namespace constraints

#endif // GOMP_CONSTRAINTS_H
