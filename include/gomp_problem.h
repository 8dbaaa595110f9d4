/*
 * gomp_problem.h -- C ABI of libgomp_host: synthetic GOMP trajectory QPs for a whole batch.
 *
 * The generator is host-only (no CUDA); the device-side re-linearisation at the end of this header lives in
 * libgomp_cuda.  The generator drives the facade's ConstraintBuilder<N> exactly as the
 * reference's GOMPSolver does (reference src/gomp-solver.h:57-64 initConstraints + build,
 * :83-87 withObstacles + build, :105-116 warm start) with the repo's synthetic arm standing
 * in for the un-vendored UR5e kinematics, and hands out what the solver boundary consumes:
 * one shared CSC pattern, per-problem A values, l, u and the warm start.
 * The concrete instance is documented in DESIGN.md.
 */
#ifndef GOMP_PROBLEM_H
#define GOMP_PROBLEM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gomp_problem_config {
    int32_t n_dim;       /* joints N: 2..7 */
    int32_t waypoints;   /* W >= 4 */
    double time_step;    /* dt */
    double pos_limit;    /* |q|   <= pos_limit            (rad) */
    double vel_limit;    /* |dq|  <= vel_limit  per second (rad/s)  */
    double acc_limit;    /* |ddq| <= acc_limit  per second^2 */
    int32_t n_balls;     /* 1: gripper ball (r = 0.05); 2: + forearm ball (r = 0.15, not gripper) listed first */
    int32_t n_obstacles; /* 0..2 horizontal lines across the start->goal chord */
    double z_floor;      /* workspace box: z >= z_floor for the gripper ball */
    uint64_t seed;       /* problem b draws start/goal from mt19937_64(seed + b) */
} gomp_problem_config;

int gomp_problem_default_config(gomp_problem_config *cfg); /* N=7, W=50: BASELINE.json configs[1] */

int gomp_problem_dims(const gomp_problem_config *cfg, int64_t *n, int64_t *m, int64_t *nnzA, int64_t *nnzP);

/* Shared structure: A pattern (CSC), P = triDiagonalMatrix(2,-1,2NW,NW,N) (full symmetric
 * CSC, reference src/gomp-solver.h:63) and the stage-interleaved variable order. */
int gomp_problem_pattern(const gomp_problem_config *cfg, int64_t *A_colptr, int64_t *A_rowidx, int64_t *P_colptr,
                         int64_t *P_rowidx, double *P_val, int64_t *var_order);

/* Problems first_id .. first_id+count-1: A_val [count, nnzA], l/u [count, m], x0 [count, n]
 * (straight-line warm start), start_goal [count, 2N] (nullable).  nthreads <= 0: all cores. */
int gomp_problem_generate(const gomp_problem_config *cfg, int64_t first_id, int64_t count, double *A_val, double *l,
                          double *u, double *x0, double *start_goal, int32_t nthreads);

/* One SQP round of reference src/gomp-solver.h:83-87: re-linearise the Cartesian rows of the
 * same problems around trajectories x [count, n] and export the new A_val, l, u. */
int gomp_problem_relinearize(const gomp_problem_config *cfg, int64_t first_id, int64_t count, const double *x,
                             double *A_val, double *l, double *u, int32_t nthreads);

/* isSolutionOK of reference src/gomp-solver.h:141-199 for each trajectory: ok[count] = 1/0. */
int gomp_problem_check(const gomp_problem_config *cfg, int64_t first_id, int64_t count, const double *x, int32_t *ok,
                       int32_t nthreads);

/* Forward kinematics / Jacobian of the synthetic arm (ball 0 = gripper, 1 = forearm):
 * q [count, N] -> xyz [count, 3], jac [count, 3, N] (nullable). */
int gomp_problem_fk(const gomp_problem_config *cfg, int32_t ball, int64_t count, const double *q, double *xyz,
                    double *jac);

/* ---- inputs of the device-side re-linearisation (host side, libgomp_host) -------------------------------- */

/* Obstacle lines of problems first_id .. first_id+count-1: lines [count, n_obstacles, 5] =
 * unit direction (x, y), a point of the line (x, y, z); all lines are passed from above. */
int gomp_problem_scene(const gomp_problem_config *cfg, int64_t first_id, int64_t count, double *lines);

/* Cartesian rows in the order ConstraintBuilder::withObstacles writes them (reference
 * src/constraints/constraint-builder.h:90-122: ball-major, per waypoint the x/y/z workspace rows of the
 * gripper ball, then one row per obstacle).  rows [n_rows] = row index in A, kidx [n_rows, N] = position of
 * each cell (columns w*N .. w*N+N-1) in the CSC value array.  Call with rows = kidx = NULL for n_rows. */
int gomp_problem_cart_rows(const gomp_problem_config *cfg, int64_t *n_rows, int32_t *rows, int32_t *kidx);

/* ---- device-side withObstacles / isSolutionOK (libgomp_cuda, sm_100a) ------------------------------------------
 * The SQP round of reference src/gomp-solver.h:83-87 without leaving the GPU: batched forward kinematics +
 * Jacobian of the synthetic arm, collision classification against the horizontal lines
 * (src/horizontal-line.h:78-98) and the linearised bounds, written in place into the device arrays the solver
 * reads (gomp_batch_update with borrow_inputs).  The kinematic model must be known to the device, so this
 * covers the repo's synthetic arm only; user-supplied FK callbacks stay on the host path. */
typedef struct gomp_device_scene gomp_device_scene;

int gomp_device_scene_create(gomp_device_scene **out, const gomp_problem_config *cfg, int64_t count, int64_t n, int64_t m,
                             int64_t nnzA, const double *lines /* host, gomp_problem_scene */, int64_t n_rows,
                             const int32_t *rows, const int32_t *kidx /* host, gomp_problem_cart_rows */, int32_t device);
/* gomp_problem_generate on the device (BASELINE.json configs[4]: inputs built on the GPU so that only start / goal cross
 * PCIe, 16 N bytes per QP instead of 8 (nnzA + 2 m + n)).  set_template: (A_val, l, u) of ANY one problem of the
 * configuration (host or device arrays of one problem) -- every row that does not depend on the problem is taken from it.
 * assemble: start_goal [count, 2N] (device) -> the scene's obstacle lines (`lines` may be NULL at scene creation), the
 * straight-line warm start x0 [count, n], and A_val / l / u [count, .] exactly as reference src/gomp-solver.h:57-64 +
 * :105-116 build them: start pinned at waypoint 0, goal at waypoint W-3, Cartesian rows linearised around x0.  All device
 * arrays, asynchronous on `stream`. */
int gomp_device_scene_set_template(gomp_device_scene *s, const double *A_val, const double *l, const double *u);
int gomp_device_assemble(gomp_device_scene *s, const double *start_goal, double *A_val, double *l, double *u, double *x0,
                         void *stream);
/* x [count, n] (device): trajectories to linearise around; a problem whose x is NaN (infeasible QP) uses its
 * row of x_fallback (nullable: then that problem is left untouched).  A_val [count, nnzA], l/u [count, m]
 * (device) are updated in place; only the Cartesian rows are written.  Asynchronous on `stream`. */
int gomp_device_relinearize(gomp_device_scene *s, const double *x, const double *x_fallback, double *A_val, double *l,
                            double *u, void *stream);
/* isSolutionOK (reference src/gomp-solver.h:141-199): ok [count] (device int32) = 1 / 0. */
int gomp_device_check(gomp_device_scene *s, const double *x, int32_t *ok, void *stream);
int gomp_device_scene_destroy(gomp_device_scene *s);

#ifdef __cplusplus
}
#endif
#endif /* GOMP_PROBLEM_H */
