/*
 * gomp_problem.h -- C ABI of libgomp_host: synthetic GOMP trajectory QPs for a whole batch.
 *
 * Host-only (no CUDA).  It drives the facade's ConstraintBuilder<N> exactly as the
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

#ifdef __cplusplus
}
#endif
#endif /* GOMP_PROBLEM_H */
