// gomp-solver.h -- the GOMP planner loops on top of QPSolver.
//
// Drop-in for reference src/gomp-solver.h:13-201: GOMPSolver<N>(waypoints, time_step, pos/vel/acc
// limits, 3-D box, obstacles, balls, ik), run(start, end) = the time-shrink loop over
// W*i/SEGMENTS horizons, run(start, end, waypoints, warm) = the SQP loop
//     QPSolver ctor -> setWarmStart -> { solve ; isSolutionOK ? done : update(withObstacles(x)) }
// with the same constants (MAX_ITERATIONS, SEGMENTS) and the same status semantics
// (kOptimal when the re-checked trajectory is feasible, kUnknown otherwise).  The QPs are solved by
// the CUDA library (osqp-wrapper.h); the per-waypoint printing of the reference is behind
// gomp::verbose().
#ifndef GOMP_SOLVER_H
#define GOMP_SOLVER_H

#include <utility>

#include "constraints/constraint-builder.h"
#include "osqp-wrapper.h"
#include "trajectory-constraints.h"

using InverseKinematics = std::function<int(double *, double, double, double)>;
const int MAX_ITERATIONS = 100;
const int SEGMENTS = 10;

template <size_t N_DIM>
class GOMPSolver {
public:
    GOMPSolver(size_t waypoints, double time_step, const Constraint<N_DIM> &pos_con, const Constraint<N_DIM> &vel_con,
               const Constraint<N_DIM> &acc_con, const Constraint<3> &con_3d, const std::vector<HorizontalLine> &obstacles,
               const std::vector<RobotBall> &m, const InverseKinematics &gripper_ik)
        : max_waypoints(waypoints), time_step(time_step), pos_con(pos_con),
          vel_con(scaled<N_DIM>(vel_con, time_step)),            // per-step displacement limits
          acc_con(scaled<N_DIM>(acc_con, time_step * time_step)), // per-step^2
          con_3d(con_3d), obstacles(obstacles), mappers(m), gripper_ik(gripper_ik) {
        assert(max_waypoints >= 4);
    }

    // Shortest feasible horizon: solve for W*i/SEGMENTS waypoints, i = SEGMENTS..1, warm-starting each
    // horizon from the last optimal trajectory; stop at the first horizon that is not solvable.
    std::pair<ExitCode, QPVector> run(Ctrl<N_DIM> start_pos, Ctrl<N_DIM> end_pos) {
        QPVector last_solution = gomp::straightLineWarmStart<N_DIM>(start_pos, end_pos, max_waypoints);
        ExitCode last_code = ExitCode::kUnknown;
        for (int i = SEGMENTS; i >= 1; --i) {
            const size_t waypoints = max_waypoints * i / SEGMENTS;
            if (waypoints < 4) break;
            // (the reference's slicing of the previous solution, kept verbatim: positions from the
            //  front, "velocities" from offset waypoints*N -- gomp-solver.h:43-44)
            QPVector warm_start(waypoints * N_DIM * 2);
            for (size_t k = 0; k < waypoints * N_DIM; ++k) {
                warm_start[k] = last_solution[k];
                warm_start[waypoints * N_DIM + k] = last_solution[waypoints * N_DIM + k];
            }
            auto [exit_code, solution] = run(start_pos, end_pos, waypoints, warm_start);
            if (exit_code != ExitCode::kOptimal && exit_code != ExitCode::kUnknown) {
                break;
            } else if (exit_code == ExitCode::kOptimal) {
                last_code = ExitCode::kOptimal;
                last_solution = solution;
            }
        }
        const long long half = last_solution.size() / 2;
        for (long long k = half; k < last_solution.size(); ++k) last_solution[k] /= time_step;
        return {last_code, last_solution};
    }

    // One horizon: sequential convex programming over the linearised obstacle / workspace rows.
    std::pair<ExitCode, QPVector> run(Ctrl<N_DIM> start_pos, Ctrl<N_DIM> end_pos, size_t waypoints, const QPVector &warm_start) {
        ConstraintBuilder<N_DIM> constraint_builder = gomp::makeTrajectoryConstraints<N_DIM>(
            waypoints, start_pos, end_pos, pos_con, vel_con, acc_con, con_3d, obstacles, mappers, warm_start);

        QPSolver qp_solver{constraint_builder.build(), triDiagonalMatrix(2, -1, N_DIM * 2 * waypoints, waypoints * N_DIM, N_DIM),
                           gomp::stageInterleavedOrder(N_DIM, waypoints), (int)(2 * N_DIM)};
        qp_solver.setWarmStart(warm_start);

        QPVector last_solution = warm_start;
        ExitCode last_code = ExitCode::kUnknown;
        for (int round = 0; round < MAX_ITERATIONS; ++round) {
            auto [exit_code, solution] = qp_solver.solve();
            if (exit_code != ExitCode::kOptimal) {
                last_solution = solution; // no solution for this horizon
                break;
            }
            if (isSolutionOK(solution)) {
                last_solution = solution;
                last_code = ExitCode::kOptimal;
                break;
            }
            qp_solver.update(constraint_builder.withObstacles(con_3d, solution).build());
        }
        return {last_code, last_solution};
    }

    // Feasibility re-check on the true (non-linearised) kinematics: workspace box for the gripper
    // ball and clearance over every obstacle the trajectory touches, both with ERROR slack.
    bool isSolutionOK(const QPVector &q_trajectory) const {
        bool ok = true;
        for (const RobotBall &ball : mappers) {
            const QPVector xyz = mapJointTrajectoryToXYZ<N_DIM>(q_trajectory, ball.fk);
            const int count = (int)(xyz.size() / 3);
            for (int w = 0; w < count; ++w) {
                const Point p{xyz[3 * w], xyz[3 * w + 1], xyz[3 * w + 2]};
                if (ball.is_gripper) {
                    for (Axis axis : XYZ_AXES) {
                        const double lo = con_3d.first.has_value() ? (*con_3d.first)[axis] : -INF;
                        const double hi = con_3d.second.has_value() ? (*con_3d.second)[axis] : INF;
                        if (!(lo - ERROR <= p[axis] - ball.radius && p[axis] + ball.radius <= hi + ERROR)) ok = false;
                    }
                }
                for (const HorizontalLine &line : obstacles)
                    if (line.hasCollision(w, xyz, ball) && !line.isAbove(p, ball)) ok = false;
            }
        }
        return ok;
    }

private:
    const size_t max_waypoints;
    const double time_step;
    const Constraint<N_DIM> pos_con;
    const Constraint<N_DIM> vel_con;
    const Constraint<N_DIM> acc_con;
    const Constraint<3> con_3d;
    const std::vector<HorizontalLine> obstacles;
    const std::vector<RobotBall> mappers;
    const InverseKinematics gripper_ik; // stored, never called (as in the reference)
};

#endif // GOMP_SOLVER_H
