// solver-example.cpp -- the reference's example (reference examples/solver-example.cpp:35-101) on the
// CUDA-backed GOMPSolver: one 6-DoF arm, a long horizon, gripper + forearm balls, workspace box, no
// obstacles.  The UR5e kinematics library of the reference is not vendored; the repo's synthetic chain
// (first 6 joints) stands in.  Build:  make -C osqp-solver_b200 example
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>

#include "gomp-solver.h"
#include "synthetic-arm.h"

constexpr double TIME_STEP = 0.1;
constexpr size_t DIMS = 6;

int main(int argc, char **argv) {
    // the reference's horizon (examples/solver-example.cpp:12): 802 waypoints -> the streamed long-horizon back
    // end; horizons up to ~60 waypoints run on the shared-memory resident kernel
    const size_t WAYPOINTS = argc > 1 ? (size_t)std::atoi(argv[1]) : 802;
    const gomp::ArmGeometry geo = gomp::pandaLikeGeometry();
    std::vector<RobotBall> mappers{
        RobotBall([geo](double *q) { double p[3]; gomp::armPoint<DIMS>(geo, q, 4, 0.0, p); return std::tuple<double, double, double>{p[0], p[1], p[2]}; },
                  [geo](double *out, double *q) { gomp::armJacobian<DIMS>(geo, q, 4, 0.0, out); }, 0.15, false),
        RobotBall([geo](double *q) { double p[3]; gomp::armPoint<DIMS>(geo, q, DIMS, geo.flange, p); return std::tuple<double, double, double>{p[0], p[1], p[2]}; },
                  [geo](double *out, double *q) { gomp::armJacobian<DIMS>(geo, q, DIMS, geo.flange, out); }, 0.05, true),
    };
    GOMPSolver<DIMS> planner(WAYPOINTS, TIME_STEP, constraints::inRange<DIMS>(of<DIMS>(-2 * M_PI), of<DIMS>(2 * M_PI)),
                             constraints::inRange<DIMS>(of<DIMS>(-M_PI), of<DIMS>(M_PI)),
                             constraints::inRange<DIMS>(of<DIMS>(-M_PI * 800 / 180), of<DIMS>(M_PI * 800 / 180)),
                             constraints::inRange<3>({{-INF, -0.4, -INF}}, {{INF, INF, INF}}), {}, mappers,
                             [](double *, double, double, double) { return 0; });
    // argv[2]: how many times to plan (the first call also pays for the CUDA context and the library's kernels being loaded)
    const int repeats = argc > 2 ? std::atoi(argv[2]) : 1;
    ExitCode code = ExitCode::kUnknown;
    QPVector traj;
    for (int rep = 0; rep < repeats; ++rep) {
        const auto t0 = std::chrono::steady_clock::now();
        auto result = planner.run({0, 0, 0, 0, 0, 0}, {M_PI, 0, 0, 0, 0, 0});
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        code = result.first;
        traj = result.second;
        std::printf("%s  (%.1f ms, trajectory of %lld values)\n", osqp::ToString(code).c_str(), ms, (long long)traj.size());
    }
    std::ofstream ctrl("output_trajectory_ctrl.data");
    for (long long i = 0; i < traj.size() / (long long)DIMS / 2; ++i) {
        for (size_t j = 0; j < DIMS; ++j) ctrl << traj[i * DIMS + j] << (j + 1 < DIMS ? " " : "\n");
    }
    return code == ExitCode::kOptimal ? 0 : 1;
}
