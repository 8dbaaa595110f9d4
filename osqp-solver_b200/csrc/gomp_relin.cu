// gomp_relin.cu -- device-side ConstraintBuilder::withObstacles and isSolutionOK for whole batches (sm_100a).
//
// One SQP round of the reference (src/gomp-solver.h:83-87: withObstacles(x) -> build -> update) re-linearises
// the Cartesian rows of A, l, u around the last trajectory.  On the host path that is a D2H copy of x, a CPU
// pass over every waypoint and an H2D copy of A_val/l/u per round -- as long as the solve itself.  Here one
// thread per (problem, ball, waypoint) does what constraint-builder.h:90-122 does for that waypoint:
//   forward kinematics of the waypoint and its two neighbours (src/utils.h:84-96), one Jacobian,
//   the three workspace rows of the gripper ball, one vertical row per obstacle line -- active when
//   HorizontalLine::hasCollision (src/horizontal-line.h:78-98) holds, otherwise the same cells with
//   infinite bounds so the pattern of A never changes (constraint-builder.h:113-115) --
// and writes the N coefficients and the two bounds of each row straight into the device arrays the solver
// borrows.  The kinematics are the repo's synthetic arm (include/gomp/synthetic-arm.h, host/device code);
// the arithmetic mirrors the facade expression by expression and is compiled without FMA contraction, so
// host and device differ only through sin/cos (<= 2 ulp).
// isSolutionOK (src/gomp-solver.h:141-199) is the second kernel: per-problem feasibility flag.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "gomp_problem.h"
#include "synthetic-arm.h"

namespace {

constexpr double kInf = 1e30;       // constraints.h:19
constexpr double kSlack = 1e-3;     // ERROR, utils.h:65 (reference utils.h:31)
constexpr double kGripperRadius = 0.05, kForearmRadius = 0.15; // gomp_host.cpp

struct Ball {
    int upto;      // joints that move the ball centre
    double offset; // tool offset along the last of them
    double radius;
    int gripper;
    int row_base; // first entry of the Cartesian row table that belongs to this ball
};

struct SceneDev {
    int N, W, O, n_balls;
    long long n, m, nnzA, count;
    Ball ball[2];
    double box_lo[3], box_hi[3];
    gomp::ArmGeometry geo;
    double *lines;        // [count][O][5]
    const double *tA, *tl, *tu; // template of one problem (gomp_device_scene_set_template): everything but the start / goal rows
                                // and the Cartesian rows is the same for every problem of a configuration
    int pos_row0;         // row of (waypoint 0, joint 0) of the position box rows
    const int32_t *rows;  // [n_rows]
    const int32_t *kidx;  // [n_rows][N]
};

struct Line {
    double dx, dy, dz, ax, ay, az;
    // HorizontalLine::getDistanceVec (horizontal-line.h): foot of P on the line minus P
    __device__ void distVec(const double P[3], double out[3]) const {
        const double along = (P[0] - ax) * dx + (P[1] - ay) * dy + (P[2] - az) * dz;
        out[0] = (ax + along * dx) - P[0];
        out[1] = (ay + along * dy) - P[1];
        out[2] = (az + along * dz) - P[2];
    }
    __device__ bool opposite(const double P[3], const double Q[3]) const {
        double a[3], b[3];
        distVec(P, a);
        distVec(Q, b);
        return a[0] * b[0] + a[1] * b[1] < 0;
    }
    __device__ bool close(const double P[3], double radius) const {
        double a[3];
        distVec(P, a);
        return sqrt(a[0] * a[0] + a[1] * a[1]) < radius;
    }
};

__device__ inline Line load_line(const SceneDev &S, long long prob, int k) {
    const double *o = S.lines + ((size_t)prob * S.O + k) * 5;
    Line L;
    L.dx = o[0]; L.dy = o[1]; L.dz = 0.0;
    L.ax = o[2]; L.ay = o[3]; L.az = o[4];
    return L;
}

// (problem, ball, waypoint) of this thread and the trajectory it reads (NaN -> fallback)
template <int N>
__device__ inline bool locate(const SceneDev &S, const double *x, const double *x_fallback, long long &prob, int &b, int &w,
                              const double *&traj) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per_problem = (long long)S.n_balls * S.W;
    if (idx >= S.count * per_problem) return false;
    prob = idx / per_problem;
    const int rem = (int)(idx - prob * per_problem);
    b = rem / S.W;
    w = rem - b * S.W;
    traj = x + (size_t)prob * S.n;
    if (isnan(traj[0])) {
        if (!x_fallback) return false;
        traj = x_fallback + (size_t)prob * S.n;
    }
    return true;
}

template <int N>
__global__ void relin_kernel(const SceneDev S, const double *__restrict__ x, const double *__restrict__ x_fallback, double *A_val,
                             double *l, double *u) {
    long long prob;
    int b, w;
    const double *traj;
    if (!locate<N>(S, x, x_fallback, prob, b, w, traj)) return;
    const Ball ball = S.ball[b];
    double q[N], p[3], pm[3], pp[3], J[3 * N];
    for (int j = 0; j < N; ++j) q[j] = traj[w * N + j];
    gomp::armPoint<N>(S.geo, q, ball.upto, ball.offset, p);
    gomp::armJacobian<N>(S.geo, q, ball.upto, ball.offset, J);
    if (w > 0) gomp::armPoint<N>(S.geo, traj + (w - 1) * N, ball.upto, ball.offset, pm);
    if (w + 1 < S.W) gomp::armPoint<N>(S.geo, traj + (w + 1) * N, ball.upto, ball.offset, pp);

    const int rows_per_w = (ball.gripper ? 3 : 0) + S.O;
    int r = ball.row_base + w * rows_per_w;
    double *Av = A_val + (size_t)prob * S.nnzA, *lo_out = l + (size_t)prob * S.m, *hi_out = u + (size_t)prob * S.m;
    // ConstraintBuilder::linearised + writeJacobianRow: bound - p_axis + J_axis . q, ball radius taken off both sides
    auto write_row = [&](int axis, double lo, double hi) {
        const int32_t *k = S.kidx + (size_t)r * N;
        for (int j = 0; j < N; ++j) Av[k[j]] = J[axis * N + j];
        const int row = S.rows[r];
        lo_out[row] = lo + ball.radius;
        hi_out[row] = hi - ball.radius;
        ++r;
    };
    auto linearised = [&](double limit, int axis) {
        double value = limit;
        value -= p[axis];
        double dot = 0.0;
        for (int j = 0; j < N; ++j) dot += J[axis * N + j] * q[j];
        value += dot;
        return value;
    };
    if (ball.gripper) {
        for (int axis = 0; axis < 3; ++axis) {
            // both sides of con_3d are present in GOMPSolver (inRange): infinite sides are linearised too
            write_row(axis, linearised(S.box_lo[axis], axis), linearised(S.box_hi[axis], axis));
        }
    }
    for (int k = 0; k < S.O; ++k) {
        const Line L = load_line(S, prob, k);
        double lo = -kInf, hi = kInf;
        bool hit = L.close(p, ball.radius);
        if (!hit && w > 0) hit = L.opposite(pm, p);
        if (!hit && w + 1 < S.W) hit = L.opposite(p, pp);
        if (hit) {
            double dv[3];
            L.distVec(p, dv);
            lo = linearised(p[2] + dv[2], 2); // line[p][Z]; every generated line is passed from above
        }
        write_row(2, lo, hi);
    }
}

// ---- gomp_problem_generate on the device (reference src/gomp-solver.h:57-64 initConstraints + build, :105-116 warm start) ----
// What makeTrajectoryConstraints (trajectory-constraints.h) writes depends on the problem in three places only: the rows that pin
// waypoint 0 to the start and waypoint W-3 to the goal, the Cartesian rows (relin_kernel around the warm start) and the warm
// start itself.  One thread per output element: A_val / l / u from the template with the pinned rows overwritten, x0 = [linspace
// (start, goal, W) | 0] with the arithmetic of utils.h linspace.
__global__ void assemble_fill_kernel(const SceneDev S, const double *__restrict__ sg, double *A_val, double *l, double *u, double *x0) {
    const long long per = S.nnzA + 2 * S.m + S.n;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S.count * per) return;
    const long long prob = idx / per;
    long long e = idx - prob * per;
    const double *start = sg + (size_t)prob * 2 * S.N, *goal = start + S.N;
    if (e < S.nnzA) {
        A_val[(size_t)prob * S.nnzA + e] = S.tA[e];
        return;
    }
    e -= S.nnzA;
    if (e < 2 * S.m) {
        const bool upper = e >= S.m;
        const long long row = upper ? e - S.m : e;
        double v = upper ? S.tu[row] : S.tl[row];
        const long long r0 = row - S.pos_row0, rg = r0 - (long long)(S.W - 3) * S.N;
        if (r0 >= 0 && r0 < S.N) v = start[r0];      // position(0, equal(start))
        else if (rg >= 0 && rg < S.N) v = goal[rg];  // position(W - 3, equal(goal))
        (upper ? u : l)[(size_t)prob * S.m + row] = v;
        return;
    }
    e -= 2 * S.m;
    double v = 0.0;
    if (e < (long long)S.W * S.N) {
        const int w = (int)(e / S.N), j = (int)(e - (long long)w * S.N);
        const double step = (goal[j] - start[j]) / (double)(S.W - 1);
        v = step * (double)w + start[j];
    }
    x0[(size_t)prob * S.n + e] = v;
}

// obstacle lines of a problem from its start and goal (gomp_host.cpp drawInstance): line k is perpendicular to the start->goal
// chord of the gripper (XY), crosses it at (k+1)/(O+1), kLineDrop below the higher end point
template <int N>
__global__ void assemble_lines_kernel(const SceneDev S, const double *__restrict__ sg) {
    const long long prob = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (prob >= S.count) return;
    const Ball g = S.ball[S.n_balls - 1]; // the gripper is listed last
    double ps[3], pg[3];
    gomp::armPoint<N>(S.geo, sg + (size_t)prob * 2 * N, g.upto, g.offset, ps);
    gomp::armPoint<N>(S.geo, sg + (size_t)prob * 2 * N + N, g.upto, g.offset, pg);
    double dx = pg[0] - ps[0], dy = pg[1] - ps[1];
    if (sqrt(dx * dx + dy * dy) < 1e-9) {
        dx = 1.0;
        dy = 0.0;
    }
    const double vx = -dy, vy = dx, len = sqrt(vx * vx + vy * vy);
    for (int k = 0; k < S.O; ++k) {
        const double f = (double)(k + 1) / (double)(S.O + 1);
        double *o = S.lines + ((size_t)prob * S.O + k) * 5;
        o[0] = vx / len;
        o[1] = vy / len;
        o[2] = ps[0] + f * (pg[0] - ps[0]);
        o[3] = ps[1] + f * (pg[1] - ps[1]);
        o[4] = fmax(ps[2], pg[2]) - 0.15; // kLineDrop
    }
}

__global__ void fill_kernel(int32_t *ok, long long count, int32_t v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) ok[i] = v;
}

template <int N>
__global__ void check_kernel(const SceneDev S, const double *__restrict__ x, int32_t *ok) {
    long long prob;
    int b, w;
    const double *traj;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per_problem = (long long)S.n_balls * S.W;
    if (idx >= S.count * per_problem) return;
    prob = idx / per_problem;
    const int rem = (int)(idx - prob * per_problem);
    b = rem / S.W;
    w = rem - b * S.W;
    traj = x + (size_t)prob * S.n;
    if (isnan(traj[0])) { // no trajectory: not a solution
        ok[prob] = 0;
        return;
    }
    const Ball ball = S.ball[b];
    double p[3], pm[3], pp[3];
    gomp::armPoint<N>(S.geo, traj + w * N, ball.upto, ball.offset, p);
    if (w > 0) gomp::armPoint<N>(S.geo, traj + (w - 1) * N, ball.upto, ball.offset, pm);
    if (w + 1 < S.W) gomp::armPoint<N>(S.geo, traj + (w + 1) * N, ball.upto, ball.offset, pp);
    bool good = true;
    if (ball.gripper)
        for (int a = 0; a < 3; ++a)
            if (!(S.box_lo[a] - kSlack <= p[a] - ball.radius && p[a] + ball.radius <= S.box_hi[a] + kSlack)) good = false;
    for (int k = 0; k < S.O; ++k) {
        const Line L = load_line(S, prob, k);
        bool hit = L.close(p, ball.radius);
        if (!hit && w > 0) hit = L.opposite(pm, p);
        if (!hit && w + 1 < S.W) hit = L.opposite(p, pp);
        if (hit && !(p[2] - L.az >= ball.radius - kSlack)) good = false; // HorizontalLine::isAbove, bypass from above
    }
    if (!good) ok[prob] = 0;
}

thread_local std::string g_relin_error;

} // namespace

struct gomp_device_scene {
    SceneDev d;
    int device = 0;
    double *lines = nullptr;
    int32_t *rows = nullptr, *kidx = nullptr;
    double *tmpl = nullptr; // [nnzA + 2 m]
};

extern "C" {

int gomp_device_scene_destroy(gomp_device_scene *s) {
    if (!s) return 0;
    cudaSetDevice(s->device);
    if (s->lines) cudaFree(s->lines);
    if (s->rows) cudaFree(s->rows);
    if (s->kidx) cudaFree(s->kidx);
    if (s->tmpl) cudaFree(s->tmpl);
    delete s;
    return 0;
}

int gomp_device_scene_create(gomp_device_scene **out, const gomp_problem_config *cfg, int64_t count, int64_t n, int64_t m, int64_t nnzA,
                             const double *lines, int64_t n_rows, const int32_t *rows, const int32_t *kidx, int32_t device) {
    if (!out || !cfg || count <= 0 || !rows || !kidx) return 1; // lines == NULL: filled in by gomp_device_assemble
    if (cfg->n_dim < 2 || cfg->n_dim > 7 || cfg->n_balls < 1 || cfg->n_balls > 2) return 1;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return 3; // no CPU fallback
    gomp_device_scene *s = new gomp_device_scene();
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) {
        delete s;
        return 4;
    }
    cudaGetDevice(&s->device);
    SceneDev &d = s->d;
    const int N = cfg->n_dim;
    d.N = N; d.W = cfg->waypoints; d.O = cfg->n_obstacles; d.n_balls = cfg->n_balls;
    d.n = n; d.m = m; d.nnzA = nnzA; d.count = count;
    d.geo = gomp::pandaLikeGeometry();
    // ball list in the order libgomp_host builds it (gomp_host.cpp Scene): forearm first when present, gripper last
    int nb = 0, base = 0;
    if (cfg->n_balls >= 2) {
        d.ball[nb] = Ball{N >= 5 ? 4 : (N > 1 ? N - 1 : 1), 0.0, kForearmRadius, 0, base};
        base += d.W * d.O;
        ++nb;
    }
    d.ball[nb] = Ball{N, d.geo.flange, kGripperRadius, 1, base};
    base += d.W * (3 + d.O);
    if (base != n_rows) {
        delete s;
        return 2;
    }
    d.box_lo[0] = d.box_lo[1] = -kInf; d.box_lo[2] = cfg->z_floor;
    d.box_hi[0] = d.box_hi[1] = d.box_hi[2] = kInf;
    const size_t lb = (size_t)count * std::max(1, d.O) * 5 * sizeof(double);
    bool ok = cudaMalloc((void **)&s->lines, lb) == cudaSuccess && cudaMalloc((void **)&s->rows, (size_t)n_rows * 4) == cudaSuccess &&
              cudaMalloc((void **)&s->kidx, (size_t)n_rows * N * 4) == cudaSuccess;
    if (ok && d.O > 0 && lines) ok = cudaMemcpy(s->lines, lines, (size_t)count * d.O * 5 * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess;
    if (ok) ok = cudaMemcpy(s->rows, rows, (size_t)n_rows * 4, cudaMemcpyHostToDevice) == cudaSuccess;
    if (ok) ok = cudaMemcpy(s->kidx, kidx, (size_t)n_rows * N * 4, cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) {
        gomp_device_scene_destroy(s);
        return 4;
    }
    d.lines = s->lines; d.rows = s->rows; d.kidx = s->kidx;
    d.tA = d.tl = d.tu = nullptr;
    d.pos_row0 = N * (d.W - 1); // after the N (W - 1) dynamics rows (constraint-builder.h row layout; gomp_host.cpp cartRows)
    *out = s;
    return 0;
}

#define RELIN_DISPATCH(KERNEL, ...)                                                   \
    switch (s->d.N) {                                                                 \
    case 2: KERNEL<2><<<grid, block, 0, st>>>(__VA_ARGS__); break;                    \
    case 3: KERNEL<3><<<grid, block, 0, st>>>(__VA_ARGS__); break;                    \
    case 4: KERNEL<4><<<grid, block, 0, st>>>(__VA_ARGS__); break;                    \
    case 5: KERNEL<5><<<grid, block, 0, st>>>(__VA_ARGS__); break;                    \
    case 6: KERNEL<6><<<grid, block, 0, st>>>(__VA_ARGS__); break;                    \
    default: KERNEL<7><<<grid, block, 0, st>>>(__VA_ARGS__); break;                   \
    }

int gomp_device_relinearize(gomp_device_scene *s, const double *x, const double *x_fallback, double *A_val, double *l, double *u,
                            void *stream) {
    if (!s || !x || !A_val || !l || !u) return 1;
    if (cudaSetDevice(s->device) != cudaSuccess) return 4;
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = s->d.count * s->d.n_balls * s->d.W;
    const int block = 128;
    const unsigned grid = (unsigned)((total + block - 1) / block);
    RELIN_DISPATCH(relin_kernel, s->d, x, x_fallback, A_val, l, u)
    return cudaGetLastError() == cudaSuccess ? 0 : 4;
}

int gomp_device_scene_set_template(gomp_device_scene *s, const double *A_val, const double *l, const double *u) {
    if (!s || !A_val || !l || !u) return 1;
    if (cudaSetDevice(s->device) != cudaSuccess) return 4;
    const size_t na = (size_t)s->d.nnzA, m = (size_t)s->d.m;
    if (!s->tmpl && cudaMalloc((void **)&s->tmpl, (na + 2 * m) * sizeof(double)) != cudaSuccess) return 4;
    if (cudaMemcpy(s->tmpl, A_val, na * sizeof(double), cudaMemcpyDefault) != cudaSuccess ||
        cudaMemcpy(s->tmpl + na, l, m * sizeof(double), cudaMemcpyDefault) != cudaSuccess ||
        cudaMemcpy(s->tmpl + na + m, u, m * sizeof(double), cudaMemcpyDefault) != cudaSuccess)
        return 4;
    s->d.tA = s->tmpl; s->d.tl = s->tmpl + na; s->d.tu = s->tmpl + na + m;
    return 0;
}

int gomp_device_assemble(gomp_device_scene *s, const double *start_goal, double *A_val, double *l, double *u, double *x0, void *stream) {
    if (!s || !start_goal || !A_val || !l || !u || !x0) return 1;
    if (!s->d.tA) return 1; // gomp_device_scene_set_template first
    if (cudaSetDevice(s->device) != cudaSuccess) return 4;
    cudaStream_t st = (cudaStream_t)stream;
    const int block = 256;
    {
        const long long total = s->d.count * (s->d.nnzA + 2 * s->d.m + s->d.n);
        assemble_fill_kernel<<<(unsigned)((total + block - 1) / block), block, 0, st>>>(s->d, start_goal, A_val, l, u, x0);
    }
    if (s->d.O > 0) {
        const unsigned grid = (unsigned)((s->d.count + block - 1) / block);
        RELIN_DISPATCH(assemble_lines_kernel, s->d, start_goal)
    }
    if (cudaGetLastError() != cudaSuccess) return 4;
    return gomp_device_relinearize(s, x0, nullptr, A_val, l, u, stream); // withObstacles around the warm start
}

int gomp_device_check(gomp_device_scene *s, const double *x, int32_t *ok, void *stream) {
    if (!s || !x || !ok) return 1;
    if (cudaSetDevice(s->device) != cudaSuccess) return 4;
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = s->d.count * s->d.n_balls * s->d.W;
    const int block = 128;
    fill_kernel<<<(unsigned)((s->d.count + block - 1) / block), block, 0, st>>>(ok, s->d.count, 1);
    const unsigned grid = (unsigned)((total + block - 1) / block);
    RELIN_DISPATCH(check_kernel, s->d, x, ok)
    return cudaGetLastError() == cudaSuccess ? 0 : 4;
}

} // extern "C"
