/*
 * oracle/osqp_ref.h -- CPU restatement of the OSQP (0.6.x) ADMM algorithm.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / timed CPU baseline.
 *
 * PARITY UNPINNED.  The reference's solver arithmetic lives in third-party code
 * that is not vendored under /root/reference and is not on this box:
 *   google/osqp-cpp @ unpinned default branch (reference src/CMakeLists.txt:25-30)
 *     -> OSQP 0.6.x (C) + QDLDL + AMD.
 * The reference calls it at src/osqp-wrapper.h:16-31 (Init), :33-43
 * (UpdateConstraintMatrix + SetBounds), :45-49 (SetPrimalWarmStart), :51-54 (Solve).
 * None of the reference's own tests (tests/test.cpp) constructs a QPSolver, so no golden
 * vector pins any solve() result.  This file restates the published algorithm
 * (Stellato et al., "OSQP: an operator splitting solver for quadratic programs", 2020,
 * Alg. 1 + sec. 5) and the behaviour of the OSQP 0.6 C sources as recollected in
 * SURVEY.md 3.4.  It is cross-checked in tests/ against closed-form QPs, KKT optimality
 * conditions and scipy, not against genuine OSQP output.
 */
#ifndef OSQP_REF_H
#define OSQP_REF_H

#ifdef __cplusplus
extern "C" {
#endif

typedef long long oref_int; /* OSQP c_int is long long in osqp-cpp builds (reference src/utils.h:12) */

/* OSQP status values (osqp constants.h, recollected) */
#define OREF_SOLVED 1
#define OREF_SOLVED_INACCURATE 2
#define OREF_PRIMAL_INFEASIBLE_INACCURATE 3
#define OREF_DUAL_INFEASIBLE_INACCURATE 4
#define OREF_MAX_ITER_REACHED (-2)
#define OREF_PRIMAL_INFEASIBLE (-3)
#define OREF_DUAL_INFEASIBLE (-4)
#define OREF_NON_CVX (-7)
#define OREF_UNSOLVED (-10)

/* osqp-cpp OsqpExitCode values (what reference src/osqp-wrapper.h:51-54 returns) */
#define OREF_EXIT_OPTIMAL 0
#define OREF_EXIT_PRIMAL_INFEASIBLE 1
#define OREF_EXIT_DUAL_INFEASIBLE 2
#define OREF_EXIT_OPTIMAL_INACCURATE 3
#define OREF_EXIT_PRIMAL_INFEASIBLE_INACCURATE 4
#define OREF_EXIT_DUAL_INFEASIBLE_INACCURATE 5
#define OREF_EXIT_MAX_ITERATIONS 6
#define OREF_EXIT_INTERRUPTED 7
#define OREF_EXIT_TIME_LIMIT 8
#define OREF_EXIT_NON_CONVEX 9
#define OREF_EXIT_UNKNOWN 10

typedef struct {
    double rho;                    /* 0.1 */
    double sigma;                  /* 1e-6 */
    double alpha;                  /* 1.6 */
    int scaling;                   /* 10 Ruiz passes */
    int adaptive_rho;              /* 1 */
    int adaptive_rho_interval;     /* OSQP default 0 = wall-clock driven; PINNED here (default 25) */
    double adaptive_rho_tolerance; /* 5 */
    int max_iter;                  /* 4000 */
    double eps_abs, eps_rel;       /* 1e-3 */
    double eps_prim_inf, eps_dual_inf; /* 1e-4 */
    int scaled_termination;        /* 0 */
    int check_termination;         /* 25 */
    int warm_start;                /* 1 */
} oref_settings;

typedef struct oref_solver oref_solver;

void oref_default_settings(oref_settings *s);

/* Minimum-degree ordering of the KKT pattern for (P, A).  perm has n+m entries.  The real
 * OSQP runs AMD inside every osqp_setup; the oracle lets callers compute it once per
 * sparsity pattern and share it (pass to oref_setup) -- this only ever favours the CPU
 * baseline's timing. */
int oref_kkt_ordering(oref_int n, oref_int m, const oref_int *Pp, const oref_int *Pi,
                      const oref_int *Ap, const oref_int *Ai, oref_int *perm);

/* osqp_setup.  P may be full symmetric or upper triangular (only triu is kept, as osqp-cpp
 * does).  q may be NULL (== 0, reference src/osqp-wrapper.h:22).  perm may be NULL. */
oref_solver *oref_setup(oref_int n, oref_int m, const oref_int *Pp, const oref_int *Pi,
                        const double *Px, const double *q, const oref_int *Ap,
                        const oref_int *Ai, const double *Ax, const double *l,
                        const double *u, const oref_settings *settings, const oref_int *perm);
void oref_cleanup(oref_solver *s);

int oref_warm_start_x(oref_solver *s, const double *x);   /* osqp_warm_start_x */
int oref_update_A(oref_solver *s, const double *Ax_new);  /* osqp_update_A, all values */
int oref_update_bounds(oref_solver *s, const double *l, const double *u); /* osqp_update_bounds */
int oref_solve(oref_solver *s);                           /* osqp_solve; returns status_val */

/* results of the last solve */
const double *oref_x(const oref_solver *s);  /* unscaled primal (NaN if infeasible) */
const double *oref_y(const oref_solver *s);
const double *oref_delta_y(const oref_solver *s); /* infeasibility certificates (after an infeasible status) */
const double *oref_delta_x(const oref_solver *s);  /* unscaled dual */
int oref_status(const oref_solver *s);
int oref_exit_code(const oref_solver *s);    /* osqp-cpp OsqpExitCode value */
int oref_iter(const oref_solver *s);
int oref_rho_updates(const oref_solver *s);
double oref_rho(const oref_solver *s);
double oref_pri_res(const oref_solver *s);
double oref_dua_res(const oref_solver *s);
double oref_obj_val(const oref_solver *s);
oref_int oref_nnz_L(const oref_solver *s);
int oref_num_factorizations(const oref_solver *s);
/* scaling vectors (for white-box tests of the GPU Ruiz pass) */
const double *oref_D(const oref_solver *s);
const double *oref_E(const oref_solver *s);
double oref_c(const oref_solver *s);

/* Batched driver used as the timed CPU baseline and for batch parity checks:
 * for every problem b: setup(P, q=0, A pattern + Ax[b], l[b], u[b]) -> warm_start_x(x0[b])
 * (if x0) -> solve, then `rounds` times { update_A(Ax_upd) ; update_bounds ; solve } when
 * update arrays are given.  One problem per OpenMP thread; ordering shared.
 * Outputs: x[batch*n], y[batch*m] (nullable), status/iters/rho_updates[batch] (nullable).
 * Returns 0 on success. */
int oref_solve_batch(oref_int batch, oref_int n, oref_int m, const oref_int *Pp,
                     const oref_int *Pi, const double *Px, const oref_int *Ap,
                     const oref_int *Ai, const double *Ax, const double *l, const double *u,
                     const double *x0, const oref_settings *settings, int nthreads,
                     double *x, double *y, int *exit_code, int *iters, int *rho_updates);
/* the same, `passes` passes over the batch as one continuous work queue (item w solves QP w % batch) */
int oref_solve_batch_passes(oref_int batch, oref_int n, oref_int m, const oref_int *Pp,
                     const oref_int *Pi, const double *Px, const oref_int *Ap,
                     const oref_int *Ai, const double *Ax, const double *l, const double *u,
                     const double *x0, const oref_settings *settings, int nthreads, int passes,
                     double *x, double *y, int *exit_code, int *iters, int *rho_updates);

int oref_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
