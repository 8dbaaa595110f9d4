// utils.h -- shared types of the GOMP facade.
//
// Drop-in for reference src/utils.h:11-96: the QP boundary types (CSC fp64 with int64
// indices because OSQP's c_int is long long, dense fp64 vectors), the FK / Jacobian callback
// types, RobotBall, triDiagonalMatrix (the objective P), linspace and
// mapJointTrajectoryToXYZ.  osqp::OsqpExitCode is supplied here because osqp-cpp is not a
// dependency of this implementation.
#ifndef GOMP_UTILS_H
#define GOMP_UTILS_H

#include <array>
#include <cstdio>
#include <functional>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "eigen_lite.h"
#include "constraints/constraints.h"

namespace osqp {
// Same enumerators, same order as osqp-cpp's OsqpExitCode (numerically equal to the
// GOMP_EXIT_* codes of include/gomp_cuda.h).
enum class OsqpExitCode {
    kOptimal,
    kPrimalInfeasible,
    kDualInfeasible,
    kOptimalInaccurate,
    kPrimalInfeasibleInaccurate,
    kDualInfeasibleInaccurate,
    kMaxIterations,
    kInterrupted,
    kTimeLimitReached,
    kNonConvex,
    kUnknown,
};
inline std::string ToString(OsqpExitCode c) {
    static const char *names[] = {"kOptimal", "kPrimalInfeasible", "kDualInfeasible", "kOptimalInaccurate",
                                  "kPrimalInfeasibleInaccurate", "kDualInfeasibleInaccurate", "kMaxIterations",
                                  "kInterrupted", "kTimeLimitReached", "kNonConvex", "kUnknown"};
    int k = static_cast<int>(c);
    return (k >= 0 && k <= 10) ? names[k] : "kUnknown";
}
} // namespace osqp

using ExitCode = osqp::OsqpExitCode;
using QPMatrixSparse = Eigen::SparseMatrix<double, Eigen::ColMajor, long long>;
template <size_t Row, size_t Col>
using QPMatrix = Eigen::Matrix<double, Row, Col, Eigen::RowMajor>;
using QPVector = Eigen::VectorXd;
using QPVector2d = Eigen::Vector2d;
using QPVector3d = Eigen::Vector3d;
using Point = Eigen::Vector3d;
template <size_t N>
using Ctrl = Eigen::Vector<double, N>;
// (x, y, z) = fk(q);  jacobian(out /* 3 x N row-major */, q)
using ForwardKinematicsFun = std::function<std::tuple<double, double, double>(double *)>;
using JacobianFun = std::function<void(double *, double *)>;

enum Axis : size_t { X, Y, Z };
constexpr const std::array<Axis, 3> XYZ_AXES = {X, Y, Z};

constexpr const double CENTIMETER = 0.01;
const double ERROR = 1e-3; // slack used by the feasibility re-check (reference utils.h:31)

namespace gomp {
// Progress chatter of the reference (per-waypoint printf, reference utils.h:88,92) is off by
// default here; set to true to get it back.
inline bool &verbose() {
    static bool flag = false;
    return flag;
}
} // namespace gomp

// A collision sphere riding on the arm: where its centre is (fk), how it moves (jacobian),
// its radius, and whether the 3-D workspace box applies to it (is_gripper).
struct RobotBall {
    RobotBall(const ForwardKinematicsFun &fk, const JacobianFun &jac, double radius, bool is_gripper = false)
        : fk(fk), jacobian(jac), radius(radius), is_gripper(is_gripper) {}

    ForwardKinematicsFun fk;
    JacobianFun jacobian;
    double radius;
    bool is_gripper;
};

// n x n sparse M with, for every row i >= offset: M(i,i) = a and M(i, i -/+ diagonal_num) = b
// where that column stays inside [offset, n).  Rows above `offset` are empty.
// GOMP calls it as (2, -1, 2NW, NW, N): the sum of squared velocity differences per joint
// (reference src/gomp-solver.h:63).
inline QPMatrixSparse triDiagonalMatrix(double a, double b, int n, int offset = 0, int diagonal_num = 1) {
    std::vector<Eigen::Triplet<double>> cells;
    cells.reserve(3 * (size_t)(n - offset));
    for (int row = offset; row < n; ++row) {
        const int right = row + diagonal_num, left = row - diagonal_num;
        cells.emplace_back(row, row, a);
        if (right < n) cells.emplace_back(row, right, b);
        if (left >= offset) cells.emplace_back(row, left, b);
    }
    QPMatrixSparse M(n, n);
    M.setFromTriplets(cells.begin(), cells.end());
    return M;
}

class NoInverseKinematicSolution : public std::runtime_error {
public:
    explicit NoInverseKinematicSolution(Point) : std::runtime_error("No inverse kinematic solution exists for point ") {}
};

// n_steps points from a to b inclusive, concatenated waypoint-major.
template <size_t N>
QPVector linspace(const Eigen::Vector<double, N> &a, const Eigen::Vector<double, N> &b, size_t n_steps) {
    QPVector out(N * n_steps);
    Eigen::Vector<double, N> step = (b - a) / (double)(n_steps - 1);
    for (size_t s = 0; s < n_steps; ++s) {
        Eigen::Vector<double, N> at = step * (double)s + a;
        for (size_t k = 0; k < N; ++k) out[s * N + k] = at[k];
    }
    return out;
}

// Cartesian image of the position half of a [positions | velocities] trajectory: one FK call
// per waypoint, in waypoint order (the reference's test position3d_2 pins that order).
template <size_t N>
QPVector mapJointTrajectoryToXYZ(const QPVector &trajectory, const ForwardKinematicsFun &mapper) {
    const int waypoints = (int)(trajectory.size() / 2 / N);
    QPVector xyz(3 * waypoints);
    if (gomp::verbose()) std::printf("mapping trajectory Q to XYZ. Waypoints XYZ: %d\n", waypoints);
    for (int w = 0; w < waypoints; ++w) {
        double q[N];
        for (size_t k = 0; k < N; ++k) q[k] = trajectory[w * N + k];
        std::tuple<double, double, double> p = mapper(q);
        xyz[3 * w + 0] = std::get<0>(p);
        xyz[3 * w + 1] = std::get<1>(p);
        xyz[3 * w + 2] = std::get<2>(p);
        if (gomp::verbose()) std::printf("(%f, %f, %f)\n", xyz[3 * w], xyz[3 * w + 1], xyz[3 * w + 2]);
    }
    return xyz;
}

#endif // GOMP_UTILS_H
