// gomp_host.cpp -- libgomp_host: batch generator of synthetic GOMP QPs (see
// include/gomp_problem.h).  Everything here goes through the facade's ConstraintBuilder<N>,
// HorizontalLine and makeTrajectoryConstraints, i.e. the same code a GOMPSolver<N> user runs.
#include "gomp_problem.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

#include "horizontal-line.h"
#include "synthetic-arm.h"
#include "trajectory-constraints.h"

namespace {

constexpr double kGripperRadius = 0.05;
constexpr double kForearmRadius = 0.15;
constexpr double kLineDrop = 0.15;      // line sits this far below the higher end point
constexpr double kMinEndpointZ = 0.05;  // start/goal gripper height is resampled above this
constexpr double kSampleFraction = 0.6; // start/goal ~ U(+-0.6 * pos_limit)
constexpr double kMinChordXY = 0.3;     // goal is resampled until the horizontal start->goal distance exceeds this,
                                        // so that the obstacle line never passes within a ball radius of an end point

template <int N>
struct Scene {
    gomp::ArmGeometry geo = gomp::pandaLikeGeometry();
    std::vector<RobotBall> balls;
    Scene(const gomp_problem_config &cfg) {
        const gomp::ArmGeometry g = geo;
        const int forearm_upto = N >= 5 ? 4 : (N > 1 ? N - 1 : 1);
        if (cfg.n_balls >= 2) {
            balls.emplace_back(
                [g, forearm_upto](double *q) {
                    double p[3];
                    gomp::armPoint<N>(g, q, forearm_upto, 0.0, p);
                    return std::tuple<double, double, double>{p[0], p[1], p[2]};
                },
                [g, forearm_upto](double *out, double *q) { gomp::armJacobian<N>(g, q, forearm_upto, 0.0, out); },
                kForearmRadius, false);
        }
        balls.emplace_back(
            [g](double *q) {
                double p[3];
                gomp::armPoint<N>(g, q, N, g.flange, p);
                return std::tuple<double, double, double>{p[0], p[1], p[2]};
            },
            [g](double *out, double *q) { gomp::armJacobian<N>(g, q, N, g.flange, out); }, kGripperRadius, true);
    }
};

template <int N>
struct Instance {
    Ctrl<N> start, goal;
    std::vector<HorizontalLine> lines;
};

template <int N>
Instance<N> drawInstance(const gomp_problem_config &cfg, const gomp::ArmGeometry &geo, int64_t id) {
    std::mt19937_64 rng(cfg.seed + (uint64_t)id);
    auto uniform = [&rng]() { return (double)(rng() >> 11) * (1.0 / 9007199254740992.0); }; // [0,1)
    Instance<N> in;
    double ps[3], pg[3];
    auto draw = [&](Ctrl<N> &q, double p[3]) {
        for (int tries = 0; tries < 1000; ++tries) {
            for (int j = 0; j < N; ++j) q[j] = (2.0 * uniform() - 1.0) * kSampleFraction * cfg.pos_limit;
            gomp::armPoint<N>(geo, q.data(), N, geo.flange, p);
            if (p[2] >= kMinEndpointZ) return;
        }
    };
    draw(in.start, ps);
    for (int tries = 0; tries < 1000; ++tries) {
        draw(in.goal, pg);
        const double cx = pg[0] - ps[0], cy = pg[1] - ps[1];
        if (std::sqrt(cx * cx + cy * cy) >= kMinChordXY) break;
    }
    // obstacle k: horizontal line perpendicular to the start->goal chord (XY), crossing it at
    // fraction (k+1)/(O+1), kLineDrop below the higher end point; passed from above.
    double dx = pg[0] - ps[0], dy = pg[1] - ps[1];
    if (std::sqrt(dx * dx + dy * dy) < 1e-9) {
        dx = 1.0;
        dy = 0.0;
    }
    for (int k = 0; k < cfg.n_obstacles; ++k) {
        const double f = (double)(k + 1) / (double)(cfg.n_obstacles + 1);
        Point at{ps[0] + f * (pg[0] - ps[0]), ps[1] + f * (pg[1] - ps[1]), std::max(ps[2], pg[2]) - kLineDrop};
        in.lines.emplace_back(QPVector2d{-dy, dx}, at, false);
    }
    return in;
}

template <int N>
ConstraintBuilder<N> builderFor(const gomp_problem_config &cfg, const Scene<N> &scene, const Instance<N> &in,
                                const QPVector &around) {
    const double dt = cfg.time_step;
    const Constraint<N> pos = constraints::inRange<N>(of<N>(-cfg.pos_limit), of<N>(cfg.pos_limit));
    // per-second limits become per-step limits (reference src/gomp-solver.h:29-30)
    const Constraint<N> vel = scaled<N>(constraints::inRange<N>(of<N>(-cfg.vel_limit), of<N>(cfg.vel_limit)), dt);
    const Constraint<N> acc = scaled<N>(constraints::inRange<N>(of<N>(-cfg.acc_limit), of<N>(cfg.acc_limit)), dt * dt);
    const Constraint<3> box = constraints::inRange<3>({{-INF, -INF, cfg.z_floor}}, {{INF, INF, INF}});
    return gomp::makeTrajectoryConstraints<N>((size_t)cfg.waypoints, in.start, in.goal, pos, vel, acc, box, in.lines,
                                              scene.balls, around);
}

template <typename F>
void parallelFor(int64_t count, int32_t nthreads, F body) {
    int nt = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
    nt = (int)std::max<int64_t>(1, std::min<int64_t>(nt, count));
    if (nt == 1) {
        for (int64_t i = 0; i < count; ++i) body(i);
        return;
    }
    std::atomic<int64_t> next{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t)
        pool.emplace_back([&]() {
            for (;;) {
                int64_t i = next.fetch_add(1);
                if (i >= count) break;
                body(i);
            }
        });
    for (auto &th : pool) th.join();
}

template <int N>
struct Ops {
    static int dims(const gomp_problem_config &cfg, int64_t *n, int64_t *m, int64_t *nnzA, int64_t *nnzP) {
        Scene<N> scene(cfg);
        Instance<N> in = drawInstance<N>(cfg, scene.geo, 0);
        QPVector warm = gomp::straightLineWarmStart<N>(in.start, in.goal, (size_t)cfg.waypoints);
        ConstraintBuilder<N> b = builderFor<N>(cfg, scene, in, warm);
        if (n) *n = (int64_t)b.numVariables();
        if (m) *m = (int64_t)b.numRows();
        if (nnzA) *nnzA = (int64_t)b.cellCount();
        if (nnzP) {
            QPMatrixSparse P = triDiagonalMatrix(2, -1, N * 2 * cfg.waypoints, cfg.waypoints * N, N);
            *nnzP = (int64_t)P.nonZeros();
        }
        return 0;
    }
    static int pattern(const gomp_problem_config &cfg, int64_t *Ap, int64_t *Ai, int64_t *Pp, int64_t *Pi, double *Px,
                       int64_t *order) {
        Scene<N> scene(cfg);
        Instance<N> in = drawInstance<N>(cfg, scene.geo, 0);
        QPVector warm = gomp::straightLineWarmStart<N>(in.start, in.goal, (size_t)cfg.waypoints);
        ConstraintBuilder<N> b = builderFor<N>(cfg, scene, in, warm);
        std::vector<long long> cp, ri;
        b.exportPattern(cp, ri);
        if (Ap) std::copy(cp.begin(), cp.end(), Ap);
        if (Ai) std::copy(ri.begin(), ri.end(), Ai);
        QPMatrixSparse P = triDiagonalMatrix(2, -1, N * 2 * cfg.waypoints, cfg.waypoints * N, N);
        if (Pp) std::copy(P.outerIndexPtr(), P.outerIndexPtr() + P.cols() + 1, Pp);
        if (Pi) std::copy(P.innerIndexPtr(), P.innerIndexPtr() + P.nonZeros(), Pi);
        if (Px) std::copy(P.valuePtr(), P.valuePtr() + P.nonZeros(), Px);
        if (order) {
            std::vector<long long> o = gomp::stageInterleavedOrder(N, (size_t)cfg.waypoints);
            std::copy(o.begin(), o.end(), order);
        }
        return 0;
    }
    static int generate(const gomp_problem_config &cfg, int64_t first, int64_t count, double *Ax, double *l, double *u,
                        double *x0, double *sg, int32_t nthreads, const double *around) {
        Scene<N> scene(cfg);
        int64_t n, m, nnzA;
        dims(cfg, &n, &m, &nnzA, nullptr);
        parallelFor(count, nthreads, [&](int64_t i) {
            Instance<N> in = drawInstance<N>(cfg, scene.geo, first + i);
            QPVector warm = gomp::straightLineWarmStart<N>(in.start, in.goal, (size_t)cfg.waypoints);
            if (x0) std::copy(warm.data(), warm.data() + n, x0 + i * n);
            if (sg) {
                std::copy(in.start.data(), in.start.data() + N, sg + i * 2 * N);
                std::copy(in.goal.data(), in.goal.data() + N, sg + i * 2 * N + N);
            }
            if (!Ax) return;
            ConstraintBuilder<N> b = builderFor<N>(cfg, scene, in, warm);
            if (around) {
                // the SQP update of reference src/gomp-solver.h:83-87 on the same builder
                QPVector traj((long long)n);
                std::copy(around + i * n, around + (i + 1) * n, traj.data());
                const Constraint<3> box = constraints::inRange<3>({{-INF, -INF, cfg.z_floor}}, {{INF, INF, INF}});
                b.withObstacles(box, traj);
            }
            b.exportValues(Ax + i * nnzA, l + i * m, u + i * m);
        });
        return 0;
    }
    // reference src/gomp-solver.h:141-199 without the printing
    static int check(const gomp_problem_config &cfg, int64_t first, int64_t count, const double *x, int32_t *ok,
                     int32_t nthreads) {
        Scene<N> scene(cfg);
        const int64_t n = 2 * (int64_t)N * cfg.waypoints;
        parallelFor(count, nthreads, [&](int64_t i) {
            Instance<N> in = drawInstance<N>(cfg, scene.geo, first + i);
            QPVector traj((long long)n);
            std::copy(x + i * n, x + (i + 1) * n, traj.data());
            bool good = true;
            for (const RobotBall &ball : scene.balls) {
                QPVector xyz = mapJointTrajectoryToXYZ<N>(traj, ball.fk);
                for (int w = 0; w < cfg.waypoints; ++w) {
                    Point p{xyz[3 * w], xyz[3 * w + 1], xyz[3 * w + 2]};
                    if (ball.is_gripper) {
                        const double low[3] = {-INF, -INF, cfg.z_floor}, upp[3] = {INF, INF, INF};
                        for (int a = 0; a < 3; ++a)
                            if (!(low[a] - ERROR <= p[a] - ball.radius && p[a] + ball.radius <= upp[a] + ERROR)) good = false;
                    }
                    for (const HorizontalLine &line : in.lines)
                        if (line.hasCollision(w, xyz, ball) && !line.isAbove(p, ball)) good = false;
                }
            }
            ok[i] = good ? 1 : 0;
        });
        return 0;
    }
    // per-problem obstacle lines for the device-side re-linearisation: [count][O][5] = unit direction (x, y), point (x, y, z)
    static int scene(const gomp_problem_config &cfg, int64_t first, int64_t count, double *lines) {
        Scene<N> sc(cfg);
        parallelFor(count, 0, [&](int64_t i) {
            Instance<N> in = drawInstance<N>(cfg, sc.geo, first + i);
            for (int k = 0; k < cfg.n_obstacles; ++k) {
                double *o = lines + ((size_t)i * cfg.n_obstacles + k) * 5;
                o[0] = in.lines[k].direction()[0];
                o[1] = in.lines[k].direction()[1];
                o[2] = in.lines[k].point()[0];
                o[3] = in.lines[k].point()[1];
                o[4] = in.lines[k].point()[2];
            }
        });
        return 0;
    }
    // Cartesian rows in the order withObstacles writes them (constraint-builder.h: ball-major, waypoint, [x, y, z box
    // rows of the gripper], one row per obstacle): row index and CSC value index of each of the N cells
    static int cartRows(const gomp_problem_config &cfg, int64_t *n_rows, int32_t *rows, int32_t *kidx) {
        Scene<N> sc(cfg);
        const int W = cfg.waypoints;
        int64_t count = 0;
        for (const RobotBall &b : sc.balls) count += (int64_t)W * ((b.is_gripper ? 3 : 0) + cfg.n_obstacles);
        if (n_rows) *n_rows = count;
        if (!rows || !kidx) return 0;
        int64_t n, m, nnzA;
        dims(cfg, &n, &m, &nnzA, nullptr);
        std::vector<int64_t> Ap((size_t)n + 1), Ai((size_t)nnzA);
        pattern(cfg, Ap.data(), Ai.data(), nullptr, nullptr, nullptr, nullptr);
        // first Cartesian row: after the dynamics equalities and the N*(W + W-1 + W-2) box rows of q, v, a
        const int64_t dyn_rows = (int64_t)N * (W - 1);
        int64_t row = dyn_rows + (int64_t)N * (W + W - 1 + W - 2);
        int64_t r = 0;
        for (const RobotBall &b : sc.balls)
            for (int w = 0; w < W; ++w)
                for (int k = 0; k < (b.is_gripper ? 3 : 0) + cfg.n_obstacles; ++k, ++row, ++r) {
                    rows[r] = (int32_t)row;
                    for (int j = 0; j < N; ++j) {
                        const int64_t col = (int64_t)w * N + j;
                        int64_t hit = -1;
                        for (int64_t q = Ap[(size_t)col]; q < Ap[(size_t)col + 1]; ++q)
                            if (Ai[(size_t)q] == row) hit = q;
                        if (hit < 0) return 2; // the builder's row layout changed: refuse rather than scribble
                        kidx[r * N + j] = (int32_t)hit;
                    }
                }
        return 0;
    }
    static int fk(const gomp_problem_config &cfg, int32_t ball, int64_t count, const double *q, double *xyz, double *jac) {
        gomp::ArmGeometry g = gomp::pandaLikeGeometry();
        const int upto = ball == 0 ? N : (N >= 5 ? 4 : (N > 1 ? N - 1 : 1));
        const double off = ball == 0 ? g.flange : 0.0;
        (void)cfg;
        for (int64_t i = 0; i < count; ++i) {
            if (xyz) gomp::armPoint<N>(g, q + i * N, upto, off, xyz + 3 * i);
            if (jac) gomp::armJacobian<N>(g, q + i * N, upto, off, jac + 3 * N * i);
        }
        return 0;
    }
};

bool valid(const gomp_problem_config *c) {
    return c && c->n_dim >= 2 && c->n_dim <= 7 && c->waypoints >= 4 && c->n_balls >= 1 && c->n_balls <= 2 &&
           c->n_obstacles >= 0 && c->n_obstacles <= 2 && c->time_step > 0;
}

#define DISPATCH(cfg, call)             \
    switch ((cfg)->n_dim) {             \
    case 2: return Ops<2>::call;        \
    case 3: return Ops<3>::call;        \
    case 4: return Ops<4>::call;        \
    case 5: return Ops<5>::call;        \
    case 6: return Ops<6>::call;        \
    case 7: return Ops<7>::call;        \
    default: return 1;                  \
    }

} // namespace

extern "C" {

int gomp_problem_default_config(gomp_problem_config *cfg) {
    if (!cfg) return 1;
    cfg->n_dim = 7;
    cfg->waypoints = 50;
    cfg->time_step = 0.1;
    cfg->pos_limit = 2.0 * M_PI;
    cfg->vel_limit = M_PI;
    cfg->acc_limit = M_PI * 800.0 / 180.0;
    cfg->n_balls = 1;
    cfg->n_obstacles = 1;
    cfg->z_floor = -0.2;
    cfg->seed = 1234;
    return 0;
}

int gomp_problem_dims(const gomp_problem_config *cfg, int64_t *n, int64_t *m, int64_t *nnzA, int64_t *nnzP) {
    if (!valid(cfg)) return 1;
    DISPATCH(cfg, dims(*cfg, n, m, nnzA, nnzP));
}

int gomp_problem_pattern(const gomp_problem_config *cfg, int64_t *A_colptr, int64_t *A_rowidx, int64_t *P_colptr,
                         int64_t *P_rowidx, double *P_val, int64_t *var_order) {
    if (!valid(cfg)) return 1;
    DISPATCH(cfg, pattern(*cfg, A_colptr, A_rowidx, P_colptr, P_rowidx, P_val, var_order));
}

int gomp_problem_generate(const gomp_problem_config *cfg, int64_t first_id, int64_t count, double *A_val, double *l,
                          double *u, double *x0, double *start_goal, int32_t nthreads) {
    if (!valid(cfg) || count < 0) return 1;
    DISPATCH(cfg, generate(*cfg, first_id, count, A_val, l, u, x0, start_goal, nthreads, nullptr));
}

int gomp_problem_relinearize(const gomp_problem_config *cfg, int64_t first_id, int64_t count, const double *x,
                             double *A_val, double *l, double *u, int32_t nthreads) {
    if (!valid(cfg) || count < 0 || !x || !A_val) return 1;
    DISPATCH(cfg, generate(*cfg, first_id, count, A_val, l, u, nullptr, nullptr, nthreads, x));
}

int gomp_problem_check(const gomp_problem_config *cfg, int64_t first_id, int64_t count, const double *x, int32_t *ok,
                       int32_t nthreads) {
    if (!valid(cfg) || count < 0 || !x || !ok) return 1;
    DISPATCH(cfg, check(*cfg, first_id, count, x, ok, nthreads));
}

int gomp_problem_scene(const gomp_problem_config *cfg, int64_t first_id, int64_t count, double *lines) {
    if (!valid(cfg) || count < 0 || (!lines && cfg->n_obstacles > 0)) return 1;
    DISPATCH(cfg, scene(*cfg, first_id, count, lines));
}

int gomp_problem_cart_rows(const gomp_problem_config *cfg, int64_t *n_rows, int32_t *rows, int32_t *kidx) {
    if (!valid(cfg)) return 1;
    DISPATCH(cfg, cartRows(*cfg, n_rows, rows, kidx));
}

int gomp_problem_fk(const gomp_problem_config *cfg, int32_t ball, int64_t count, const double *q, double *xyz,
                    double *jac) {
    if (!valid(cfg) || !q) return 1;
    DISPATCH(cfg, fk(*cfg, ball, count, q, xyz, jac));
}

} // extern "C"
