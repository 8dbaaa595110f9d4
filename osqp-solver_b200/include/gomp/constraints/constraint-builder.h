// constraint-builder.h -- assembles the (l, A, u) of a GOMP trajectory QP.
//
// Drop-in for reference src/constraints/constraint-builder.h:18-283 (same fluent API, same
// row layout, same numbers), re-designed around a per-row cell store instead of an
// ever-growing triplet list: re-linearising the obstacle rows every SQP round
// (reference src/gomp-solver.h:83-87) overwrites cells in place, the sparsity pattern is
// computed once, and `exportValues` fills caller arrays in the fixed CSC order that the
// batched solver (include/gomp_cuda.h) shares across a whole batch.
//
// Variables (n = 2*N*W):  x = [q_0 .. q_{W-1} | v_0 .. v_{W-1}],  v_t = per-step displacement.
// Rows, in order (reference constraint-builder.h:30-45, :92):
//   link          N(W-1)   v_t + q_t - q_{t+1} = 0
//   position      N*W      q_t          (identity)
//   velocity      N(W-1)   v_t
//   acceleration  N(W-2)   v_{t+1} - v_t
//   3-D/obstacle  N*W*(3 + O*B) allocated; per ball, per waypoint: 3 rows J_axis q_t if the
//                 ball is the gripper, then one row J_z q_t per obstacle (active or dummy).
// Rows that were never written stay empty with bounds -/+INF.
#ifndef GOMP_CONSTRAINT_BUILDER_H
#define GOMP_CONSTRAINT_BUILDER_H

#include <algorithm>
#include <cassert>
#include <numeric>
#include <optional>
#include <tuple>
#include <utility>
#include <vector>

#include "../utils.h"
#include "constraints.h"
#include "../horizontal-line.h"

using namespace constraints;

// <lower_bounds, constraint_matrix, upper_bounds>
using QPConstraints = std::tuple<QPVector, QPMatrixSparse, QPVector>;

template <size_t N_DIM>
class ConstraintBuilder {
    using Jacobian = QPMatrix<3, N_DIM>;
    using Side = std::optional<double>;

    // position + velocity + acceleration
    static constexpr const size_t DYNAMICS_DERIVATIVES = 3;

    struct Cell {
        size_t col;
        double coeff;
    };

public:
    ConstraintBuilder(size_t waypoints, const std::vector<RobotBall> &m, const std::vector<HorizontalLine> &obstacles)
        : waypoints(waypoints), mappers(m), obstacles(obstacles) {
        // The v <-> q link rows come first; everything the user can constrain follows.
        for (size_t t = 0; t + 1 < waypoints; ++t) {
            for (size_t j = 0; j < N_DIM; ++j) {
                const size_t row = appendRow();
                const Cell link[3] = {{nthVelocity(t) + j, 1.0}, {nthPos(t + 1) + j, -1.0}, {nthPos(t) + j, 1.0}};
                writeRow(row, link, 3, Side(0.0), Side(0.0));
            }
        }
        userConstraintOffset = rows.size();
        const size_t per_joint = waypoints + (waypoints - 1) + (waypoints - 2) +
                                 waypoints * (3 + obstacles.size() * mappers.size());
        const size_t total = userConstraintOffset + N_DIM * per_joint;
        rows.resize(total);
        lowerBounds.resize(total, -INF);
        upperBounds.resize(total, INF);
    }

    ConstraintBuilder &position(size_t i, const Constraint<N_DIM> &c) { return positions(i, i, c); }

    ConstraintBuilder &positions(size_t first, size_t last, const Constraint<N_DIM> &c) {
        return boxVariables(nthPos(first), nthPos(last), c);
    }

    ConstraintBuilder &velocity(size_t i, const Constraint<N_DIM> &c) {
        assert(i < waypoints - 1);
        return velocities(i, i, c);
    }

    ConstraintBuilder &velocities(size_t first, size_t last, const Constraint<N_DIM> &c) {
        assert(first <= last && last < waypoints - 1);
        return boxVariables(nthVelocity(first), nthVelocity(last), c);
    }

    ConstraintBuilder &accelerations(size_t first, size_t last, const Constraint<N_DIM> &c) {
        for (size_t t = first; t <= last; ++t) acceleration(t, c);
        return *this;
    }

    // l <= v_{i+1} - v_i <= u
    ConstraintBuilder &acceleration(size_t i, const Constraint<N_DIM> &c) {
        assert(i + 2 < waypoints);
        const size_t row0 = userConstraintOffset + nthAcceleration(i);
        for (size_t j = 0; j < N_DIM; ++j) {
            const Cell diff[2] = {{nthVelocity(i + 1) + j, 1.0}, {nthVelocity(i) + j, -1.0}};
            writeRow(row0 + j, diff, 2, side(c.first, j), side(c.second, j));
        }
        return *this;
    }

    // (Re-)linearise the Cartesian rows around `trajectory` ([positions | velocities]; only
    // the position half is read).  Per ball: FK of every waypoint first, then per waypoint
    // one Jacobian, the three workspace-box rows if the ball is the gripper, and one
    // vertical row per obstacle -- active where that waypoint collides, otherwise a dummy
    // row with the same cells and infinite bounds so that the pattern of A never changes.
    ConstraintBuilder &withObstacles(const Constraint<3> &con_3d, const QPVector &trajectory) {
        size_t row = userConstraintOffset + N_DIM * (waypoints + waypoints - 1 + waypoints - 2);
        for (const RobotBall &ball : mappers) {
            const QPVector xyz = mapJointTrajectoryToXYZ<N_DIM>(trajectory, ball.fk);
            for (size_t w = 0; w < waypoints; ++w) {
                Ctrl<N_DIM> q;
                for (size_t j = 0; j < N_DIM; ++j) q[j] = trajectory[w * N_DIM + j];
                const Point p{xyz[3 * w], xyz[3 * w + 1], xyz[3 * w + 2]};
                Jacobian jac;
                ball.jacobian(jac.data(), q.data());

                if (ball.is_gripper) {
                    for (Axis axis : XYZ_AXES) {
                        double lo = -INF, hi = INF;
                        if (con_3d.first.has_value()) lo = linearised((*con_3d.first)[axis], p[axis], jac, axis, q);
                        if (con_3d.second.has_value()) hi = linearised((*con_3d.second)[axis], p[axis], jac, axis, q);
                        writeJacobianRow(row++, ball, axis, jac, w, lo, hi);
                    }
                }
                for (const HorizontalLine &line : obstacles) {
                    double lo = -INF, hi = INF;
                    if (line.hasCollision((int)w, xyz, ball)) {
                        const double level = linearised(line[p][Axis::Z], p[Axis::Z], jac, Axis::Z, q);
                        (line.bypassFromBelow() ? hi : lo) = level;
                    }
                    writeJacobianRow(row++, ball, Axis::Z, jac, w, lo, hi);
                }
            }
        }
        return *this;
    }

    QPConstraints build() {
        std::vector<Eigen::Triplet<double, long long>> cells;
        cells.reserve(cellCount());
        for (size_t r = 0; r < rows.size(); ++r)
            for (const Cell &c : rows[r]) cells.emplace_back((long long)r, (long long)c.col, c.coeff);
        QPMatrixSparse A;
        A.resize((long long)rows.size(), (long long)numVariables());
        A.setFromTriplets(cells.begin(), cells.end(), [](const double &, const double &newest) { return newest; });
        QPVector l(rows.size()), u(rows.size());
        for (size_t r = 0; r < rows.size(); ++r) {
            l[r] = lowerBounds[r];
            u[r] = upperBounds[r];
        }
        return QPConstraints{l, A, u};
    }

    [[nodiscard]] size_t nthVelocity(size_t i) const {
        assert(i < waypoints - 1);
        return waypoints * N_DIM + i * N_DIM;
    }

    [[nodiscard]] size_t nthPos(size_t i) const {
        assert(i < waypoints);
        return i * N_DIM;
    }

    [[nodiscard]] size_t nthAcceleration(size_t i) const {
        assert(i < waypoints - 2);
        return waypoints * N_DIM + (waypoints - 1) * N_DIM + i * N_DIM;
    }

    // ---- batch-oriented extras (not in the reference) -------------------------------------
    [[nodiscard]] size_t numVariables() const { return N_DIM * 2 * waypoints; } // positions + velocities
    [[nodiscard]] size_t numRows() const { return rows.size(); }
    [[nodiscard]] size_t cellCount() const {
        size_t k = 0;
        for (const auto &r : rows) k += r.size();
        return k;
    }

    // CSC pattern of A (sorted rows within each column), identical for every QP built with
    // the same sequence of calls.
    void exportPattern(std::vector<long long> &colptr, std::vector<long long> &rowidx) {
        const std::vector<std::pair<size_t, size_t>> &order = cscOrder();
        colptr.assign(numVariables() + 1, 0);
        rowidx.resize(order.size());
        for (size_t k = 0; k < order.size(); ++k) {
            const Cell &c = rows[order[k].first][order[k].second];
            rowidx[k] = (long long)order[k].first;
            colptr[c.col + 1]++;
        }
        for (size_t j = 0; j < numVariables(); ++j) colptr[j + 1] += colptr[j];
    }

    // Values of A in exportPattern order plus the bounds, written into caller storage
    // (A_val: cellCount(), l/u: numRows()).
    void exportValues(double *A_val, double *l, double *u) {
        const std::vector<std::pair<size_t, size_t>> &order = cscOrder();
        for (size_t k = 0; k < order.size(); ++k) A_val[k] = rows[order[k].first][order[k].second].coeff;
        std::copy(lowerBounds.begin(), lowerBounds.end(), l);
        std::copy(upperBounds.begin(), upperBounds.end(), u);
    }

private:
    const size_t waypoints;
    size_t userConstraintOffset = 0;
    std::vector<RobotBall> mappers;
    std::vector<HorizontalLine> obstacles;
    std::vector<std::vector<Cell>> rows; // rows[r] = cells of constraint r, newest value per column
    std::vector<double> lowerBounds;
    std::vector<double> upperBounds;
    std::vector<std::pair<size_t, size_t>> csc_order; // (row, slot) sorted by (col, row)
    bool csc_order_valid = false;

    static Side side(const Bound<N_DIM> &b, size_t dim) {
        if (b.has_value()) return Side((*b)[dim]);
        return Side();
    }

    size_t appendRow() {
        rows.emplace_back();
        lowerBounds.push_back(-INF);
        upperBounds.push_back(INF);
        return rows.size() - 1;
    }

    // Newest coefficient per (row, column) wins; an absent side keeps the stored bound.
    void writeRow(size_t row, const Cell *cells, size_t count, Side lo, Side hi) {
        std::vector<Cell> &dst = rows[row];
        for (size_t k = 0; k < count; ++k) {
            auto hit = std::find_if(dst.begin(), dst.end(), [&](const Cell &c) { return c.col == cells[k].col; });
            if (hit != dst.end()) {
                hit->coeff = cells[k].coeff;
            } else {
                dst.push_back(cells[k]);
                csc_order_valid = false;
            }
        }
        if (lo) lowerBounds[row] = *lo;
        if (hi) upperBounds[row] = *hi;
        assert(lowerBounds[row] <= upperBounds[row]);
    }

    // Identity rows on N_DIM-wide variable blocks starting at first_start, ..., last_start.
    ConstraintBuilder &boxVariables(size_t first_start, size_t last_start, const Constraint<N_DIM> &c) {
        for (size_t start = first_start; start <= last_start; start += N_DIM) {
            for (size_t j = 0; j < N_DIM; ++j) {
                const Cell one[1] = {{start + j, 1.0}};
                writeRow(userConstraintOffset + start + j, one, 1, side(c.first, j), side(c.second, j));
            }
        }
        return *this;
    }

    // First-order model of a Cartesian limit around (q, p = fk(q)):
    //   limit <= p + J (q' - q)   <=>   limit - p + J q <= J q'
    static double linearised(double limit, double p_axis, const Jacobian &jac, Axis axis, const Ctrl<N_DIM> &q) {
        double value = limit;
        value -= p_axis;
        value += jac.row(axis) * q;
        return value;
    }

    // Row `J_axis(q_w) . q_w` with the ball's radius taken off both sides.
    void writeJacobianRow(size_t row, const RobotBall &ball, Axis axis, const Jacobian &jac, size_t waypoint,
                          double lo, double hi) {
        Cell cells[N_DIM];
        for (size_t j = 0; j < N_DIM; ++j) cells[j] = Cell{nthPos(waypoint) + j, jac(axis, j)};
        writeRow(row, cells, N_DIM, Side(lo + ball.radius), Side(hi - ball.radius));
    }

    const std::vector<std::pair<size_t, size_t>> &cscOrder() {
        if (!csc_order_valid) {
            csc_order.clear();
            for (size_t r = 0; r < rows.size(); ++r)
                for (size_t s = 0; s < rows[r].size(); ++s) csc_order.emplace_back(r, s);
            std::stable_sort(csc_order.begin(), csc_order.end(),
                             [&](const std::pair<size_t, size_t> &a, const std::pair<size_t, size_t> &b) {
                                 const size_t ca = rows[a.first][a.second].col, cb = rows[b.first][b.second].col;
                                 return ca != cb ? ca < cb : a.first < b.first;
                             });
            csc_order_valid = true;
        }
        return csc_order;
    }
};

#endif // GOMP_CONSTRAINT_BUILDER_H
