// trajectory-constraints.h -- the constraint recipe of one GOMP horizon.
//
// What reference src/gomp-solver.h:118-139 (GOMPSolver::initConstraints) writes into a
// ConstraintBuilder: start pinned at waypoint 0, goal pinned at waypoint W-3 (the last
// three waypoints rest at the goal: velocity W-3 and acceleration W-3 are zero), joint
// limits in between, and the Cartesian rows linearised around the warm start.  Shared by
// GOMPSolver<N> and by the batched problem generator so that both produce identical QPs.
#ifndef GOMP_TRAJECTORY_CONSTRAINTS_H
#define GOMP_TRAJECTORY_CONSTRAINTS_H

#include <vector>

#include "constraints/constraint-builder.h"

namespace gomp {

template <size_t N_DIM>
ConstraintBuilder<N_DIM> makeTrajectoryConstraints(size_t waypoints, const Ctrl<N_DIM> &start_pos,
                                                   const Ctrl<N_DIM> &end_pos, const Constraint<N_DIM> &pos_con,
                                                   const Constraint<N_DIM> &vel_con_per_step,
                                                   const Constraint<N_DIM> &acc_con_per_step,
                                                   const Constraint<3> &con_3d,
                                                   const std::vector<HorizontalLine> &obstacles,
                                                   const std::vector<RobotBall> &balls, const QPVector &linearise_around) {
    assert(waypoints >= 4);
    ConstraintBuilder<N_DIM> b{waypoints, balls, obstacles};
    b.position(0, constraints::equal<N_DIM>(start_pos));
    b.positions(1, waypoints - 2, pos_con);
    b.position(waypoints - 3, constraints::equal<N_DIM>(end_pos));
    b.velocities(0, waypoints - 4, vel_con_per_step);
    b.velocity(waypoints - 3, EQ_ZERO<N_DIM>);
    b.accelerations(0, waypoints - 4, acc_con_per_step);
    b.acceleration(waypoints - 3, EQ_ZERO<N_DIM>);
    b.withObstacles(con_3d, linearise_around);
    return b;
}

// [linspace(start, end, W) | zeros]: the warm start of reference src/gomp-solver.h:105-116.
template <size_t N_DIM>
QPVector straightLineWarmStart(const Ctrl<N_DIM> &start_pos, const Ctrl<N_DIM> &end_pos, size_t waypoints) {
    QPVector positions = linspace<N_DIM>(start_pos, end_pos, waypoints);
    QPVector warm(2 * waypoints * N_DIM);
    warm.setZero();
    for (size_t k = 0; k < waypoints * N_DIM; ++k) warm[k] = positions[k];
    return warm;
}

// Variable order that makes the reduced KKT matrix block tridiagonal with 2N x 2N blocks:
// (q_0, v_0, q_1, v_1, ...).  order[p] = index of the variable placed at position p.
inline std::vector<long long> stageInterleavedOrder(size_t n_dim, size_t waypoints) {
    std::vector<long long> order(2 * n_dim * waypoints);
    size_t p = 0;
    for (size_t t = 0; t < waypoints; ++t) {
        for (size_t j = 0; j < n_dim; ++j) order[p++] = (long long)(t * n_dim + j);
        for (size_t j = 0; j < n_dim; ++j) order[p++] = (long long)(waypoints * n_dim + t * n_dim + j);
    }
    return order;
}

} // namespace gomp
#endif // GOMP_TRAJECTORY_CONSTRAINTS_H
