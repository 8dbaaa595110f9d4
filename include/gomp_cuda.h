/*
 * gomp_cuda.h -- C ABI of libgomp_cuda, the B200-native batched OSQP-style ADMM solver for
 * GOMP trajectory QPs.
 *
 * The reference has no FFI: its seam is the C++ class QPSolver (reference
 * src/osqp-wrapper.h:12-60), a shim over osqp::OsqpSolver.  Every entry point below
 * replaces one of that class's members for a whole batch of QPs that share a sparsity
 * pattern; the header-only facade in osqp-solver_b200/include/gomp/osqp-wrapper.h binds
 * them back to the reference's signatures.
 *
 *   gomp_batch_create        <- QPSolver::QPSolver(c, P)      osqp-wrapper.h:16-31
 *   gomp_batch_setup         <- constructing a fresh QPSolver on the same pattern
 *                               (reference src/gomp-solver.h:61-64, once per horizon)
 *   gomp_batch_warm_start_x  <- QPSolver::setWarmStart        osqp-wrapper.h:45-49
 *   gomp_batch_solve         <- QPSolver::solve               osqp-wrapper.h:51-54
 *   gomp_batch_update        <- QPSolver::update              osqp-wrapper.h:33-43
 *   gomp_batch_get           <- solver.primal_solution()      osqp-wrapper.h:53
 *   gomp_settings_default    <- osqp::OsqpSettings defaults   osqp-wrapper.h:26-27
 *
 * Conventions: plain pointers and sizes, int return codes (0 = OK), no exceptions, no
 * C++/torch types.  Every data pointer may be a HOST or a DEVICE pointer (detected with
 * cudaPointerGetAttributes).  Pageable host arrays are copied during the call and need not
 * outlive it (as with osqp-cpp, SURVEY.md 8b "Ownership"); copies from PINNED host memory are
 * truly asynchronous on the handle's stream, so a pinned input must stay unchanged until the next
 * gomp_batch_sync / gomp_batch_get on that handle; device inputs with settings.borrow_inputs are
 * used in place until the next setup / update.  Host bounds are validated (l <= u) during the
 * call; device-resident bounds are validated on the device and a violation is reported by the
 * next gomp_batch_sync / gomp_batch_get (GOMP_ERR_INVALID_ARGUMENT).
 * One handle is not thread-safe; distinct handles are.  All arithmetic is fp64.
 * There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef GOMP_CUDA_H
#define GOMP_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* return codes */
#define GOMP_OK 0
#define GOMP_ERR_INVALID_ARGUMENT 1 /* bad sizes / l > u / pattern mismatch (update throws in the reference) */
#define GOMP_ERR_CUDA 2             /* a CUDA runtime call failed; see gomp_last_error_string() */
#define GOMP_ERR_NO_DEVICE 3        /* no usable CUDA device: there is no CPU fallback */
#define GOMP_ERR_UNSUPPORTED 4      /* structure outside what the kernels handle (e.g. block too large) */
#define GOMP_ERR_ALLOC 5

/* per-problem exit codes: numerically equal to osqp-cpp's OsqpExitCode
 * (reference src/utils.h:11 `using ExitCode = osqp::OsqpExitCode`) */
#define GOMP_EXIT_OPTIMAL 0
#define GOMP_EXIT_PRIMAL_INFEASIBLE 1
#define GOMP_EXIT_DUAL_INFEASIBLE 2
#define GOMP_EXIT_OPTIMAL_INACCURATE 3
#define GOMP_EXIT_PRIMAL_INFEASIBLE_INACCURATE 4
#define GOMP_EXIT_DUAL_INFEASIBLE_INACCURATE 5
#define GOMP_EXIT_MAX_ITERATIONS 6
#define GOMP_EXIT_INTERRUPTED 7
#define GOMP_EXIT_TIME_LIMIT 8
#define GOMP_EXIT_NON_CONVEX 9
#define GOMP_EXIT_UNKNOWN 10

/* linear-system back ends */
#define GOMP_LINSYS_AUTO 0            /* resident kernel when the QP fits one SM, else the streamed one */
#define GOMP_LINSYS_BLOCKTRI_SMEM 1   /* block-tridiagonal reduced KKT, QP + factor resident in registers / shared memory */
#define GOMP_LINSYS_BLOCKTRI_GLOBAL 2 /* same factorisation, QP + factor in an L2-resident workspace, factor streamed
                                         through shared memory with TMA (long horizons: H = 400, the reference example's 802) */
#define GOMP_LINSYS_PCG 3             /* reserved: not built (an inexact solve would change OSQP's iteration counts);
                                         gomp_batch_create returns GOMP_ERR_UNSUPPORTED */
#define GOMP_LINSYS_BLOCKTRI_TMEM 4   /* same factorisation, TWO QPs resident per SM: one 512-thread CTA with two independent
                                         slots, the inverted pivot blocks of both QPs kept in tensor memory (tcgen05.st / ld,
                                         lane-private), the block factorisation inside the substitution warps.  Block size
                                         14 or 12 with the uniform 2-wide off-diagonal pattern, 4..63 stages; AUTO picks it
                                         whenever the QP qualifies */

typedef struct gomp_batch gomp_batch;

/* OSQP settings (defaults = OSQP 0.6 defaults as exposed by osqp::OsqpSettings, except
 * adaptive_rho_interval, which OSQP derives from wall-clock time when 0 and which is
 * therefore pinned to 25 here) plus device-side knobs. */
typedef struct gomp_settings {
    double rho;                    /* 0.1 */
    double sigma;                  /* 1e-6 */
    double alpha;                  /* 1.6 */
    int32_t scaling;               /* 10 */
    int32_t adaptive_rho;          /* 1 */
    int32_t adaptive_rho_interval; /* 25 (pinned; must be a multiple of check_termination, >0) */
    double adaptive_rho_tolerance; /* 5 */
    int32_t max_iter;              /* 4000 */
    double eps_abs;                /* 1e-3 */
    double eps_rel;                /* 1e-3 */
    double eps_prim_inf;           /* 1e-4 */
    double eps_dual_inf;           /* 1e-4 */
    int32_t scaled_termination;    /* 0 */
    int32_t check_termination;     /* 25 */
    int32_t warm_start;            /* 1 (0 is rejected: iterates always persist across solves of a handle, as the reference's
                                      QPSolver relies on; a cold start is gomp_batch_setup) */
    int32_t verbose;               /* 0 (the reference sets true; output only) */
    /* device side */
    int32_t device;                /* CUDA device ordinal, -1 = current */
    void *stream;                  /* cudaStream_t to launch on, NULL = a stream owned by the handle */
    int32_t linsys;                /* GOMP_LINSYS_* */
    int32_t block_size;            /* stage-block size of the variable ordering, 0 = detect */
    const int64_t *var_order;      /* n entries: position p holds variable var_order[p]; NULL = detect */
    int32_t threads_per_cta;       /* must be 0 (fixed per back end; reported by gomp_batch_info) */
    int32_t borrow_inputs;         /* 1: device-pointer inputs are used in place (caller keeps them alive until the next setup/update) */
    int32_t prefer_throughput;     /* 1 (with linsys = AUTO): when the QP also fits the two-CTAs-per-SM mode of the streamed back end
                                      (block size 14 with the uniform 2-wide off-diagonal pattern, < 113 KB of shared memory per
                                      QP, at most 4096 non-empty rows, columns of at most 15 entries starting below entry 4096,
                                      batch >= 2 QPs per SM), use it -- +40 % QPs/s on 4096 x (N=7, H=50) when several batches are
                                      kept in flight, but one QP alone takes 1.4x longer, so a single batch with a few max_iter
                                      problems (an SQP round sequence) finishes later.  0: resident kernel. */
} gomp_settings;

typedef struct gomp_info {
    int64_t batch, n, m, m_used, nnzA, nnzP;
    int32_t block_size, num_blocks, offdiag_width;
    int32_t linsys, threads_per_cta, ctas_per_sm, num_sms, grid;
    int32_t qps_per_sm;            /* QPs resident per SM (two slots per CTA on the TMEM back end) */
    int32_t reserved0;
    int64_t smem_bytes;            /* dynamic shared memory per CTA */
    int64_t workspace_bytes;       /* device memory owned by the handle */
    int64_t kernel_launches;       /* launches issued by this handle so far */
    double last_solve_ms;          /* device time of the last gomp_batch_solve (CUDA events on its stream) */
    double last_h2d_ms;            /* device time of input copies in the last setup/update */
    int64_t flops_per_iter;        /* algorithmic fp64 flops of one ADMM iteration of one QP */
    int64_t smem_bytes_per_iter;   /* algorithmic shared-memory bytes moved by one ADMM iteration */
    int64_t flops_per_factor, smem_bytes_per_factor;
    int64_t flops_per_scaling, smem_bytes_per_scaling;
    int64_t flops_per_check, smem_bytes_per_check;
    double phase_cycles[8];        /* with env GOMP_PROFILE_PHASES=1 at create: SM cycles summed over all QPs per phase
                                      (load, scaling, factor, rhs gather, block solve, x/z/y update, residuals+termination, store) */
    double chain_cycles;           /* cycles warp 0 spends in the block substitutions proper (inside phase_cycles[4]) */
    double assemble_cycles;        /* part of phase_cycles[2] spent assembling S (all threads) */
} gomp_info;

int gomp_settings_default(gomp_settings *s);

/* Build a batch of `batch` QPs  min 1/2 x'Px + q'x  s.t.  l <= Ax <= u  that share P and the
 * CSC pattern of A.  P: CSC, full symmetric or upper triangular (only the upper triangle is
 * used, as osqp-cpp does).  A_val: [batch, nnzA] (CSC order), l,u: [batch, m], q: [batch, n]
 * or NULL (the reference always passes q = 0, osqp-wrapper.h:22).  Index arrays are host
 * pointers.  Does the symbolic analysis (ordering, block structure, assembly maps) on the
 * host and loads the data; numeric setup (Ruiz scaling, rho_vec, factorisation) runs on the
 * device inside the first solve. */
int gomp_batch_create(gomp_batch **out, const gomp_settings *settings, int64_t batch, int64_t n,
                      int64_t m, const int64_t *P_colptr, const int64_t *P_rowidx,
                      const double *P_val, const int64_t *A_colptr, const int64_t *A_rowidx,
                      const double *A_val, const double *l, const double *u, const double *q);

/* Re-initialise every QP of the batch with new data as if freshly constructed (iterates and
 * rho reset: a cold start) while re-using the symbolic analysis. */
int gomp_batch_setup(gomp_batch *h, const double *A_val, const double *l, const double *u,
                     const double *q);

/* x0: [batch, n]; x <- Dinv x0, z <- A x, y untouched (OSQP osqp_warm_start_x). */
int gomp_batch_warm_start_x(gomp_batch *h, const double *x0);

/* Run ADMM on every QP until its own termination test fires.  Iterates, rho and the
 * factorisation state persist on the device across calls.  Asynchronous on the handle's stream.
 * A solve that follows gomp_batch_update (same problems, new data) pulls the problems from the
 * device work queue in descending order of the iterations their previous solve took, so that
 * the longest ones start first; gomp_batch_setup (new problems) resets that order. */
int gomp_batch_solve(gomp_batch *h);

/* New A values (same pattern), bounds and/or q; each pointer may be NULL = unchanged.
 * Iterates and rho are kept (warm start), scaling is recomputed, the KKT refactored. */
int gomp_batch_update(gomp_batch *h, const double *A_val, const double *l, const double *u,
                      const double *q);

/* Results of the last solve.  Any pointer may be NULL.  x: [batch, n], y: [batch, m],
 * exit_code/iters/rho_updates: [batch] int32.  x, y are NaN for infeasible problems.  Waits for the handle's stream;
 * with a host exit_code array OSQP's status values are copied and mapped to OsqpExitCode on the host (no kernel). */
int gomp_batch_get(gomp_batch *h, double *x, double *y, int32_t *exit_code, int32_t *iters,
                   int32_t *rho_updates);

/* The same copies queued on the handle's stream WITHOUT waiting for them: for device destinations (a collective
 * that follows on another stream orders itself after the handle's stream) or pinned host buffers read after
 * gomp_batch_sync.  Exit codes are mapped to OsqpExitCode values on the device. */
int gomp_batch_get_async(gomp_batch *h, double *x, double *y, int32_t *exit_code, int32_t *iters,
                         int32_t *rho_updates);

/* Extra per-problem diagnostics: pri_res, dua_res, rho, obj_val: [batch] or NULL. */
int gomp_batch_get_info(gomp_batch *h, double *pri_res, double *dua_res, double *rho,
                        double *obj_val);

/* Infeasibility certificates of the last solve, as osqp-cpp hands them out (OsqpSolver::primal_infeasibility_certificate /
 * dual_infeasibility_certificate; the reference never reads them, src/osqp-wrapper.h:53).  prim_inf_cert [batch, m]: for
 * a problem that ended (inaccurately) primal infeasible, delta_y of the terminating iteration projected onto the polar of
 * the recession cone of [l, u], un-scaled and normalised to unit infinity norm (OSQP check_termination + store_solution):
 * A' v ~ 0 and u' max(v, 0) + l' min(v, 0) < 0.  dual_inf_cert [batch, n]: delta_x likewise for dual infeasible problems
 * (only possible with q != 0).  Rows of problems with any other status are NaN.  Either pointer may be NULL; host or
 * device pointers; waits for the handle's stream. */
int gomp_batch_get_certificates(gomp_batch *h, double *prim_inf_cert, double *dual_inf_cert);

/* Per-problem mask for the following solves: active [batch] int32 (host or device; borrowed when it is a
 * device pointer and settings.borrow_inputs is set), NULL = every problem.  A problem with active == 0 is
 * skipped: its iterates, rho and last results stay untouched.  This is what a batched SQP loop
 * (reference src/gomp-solver.h:70-88, one QPSolver per problem there) needs once some problems have
 * converged or failed while others still iterate. */
int gomp_batch_set_active(gomp_batch *h, const int32_t *active);

/* Cap on the CTAs the following solves of this handle launch (0 = default: every resident slot of the device; ignored
 * by the cluster mode of the streamed back end).  No counterpart in the reference (one QPSolver = one CPU thread); it is
 * the knob a caller who keeps SEVERAL batches in flight on one GPU uses to share the SMs between them.  A batch's CTAs are
 * persistent -- they leave their SM only when the batch's work queue is empty -- so with the default, batches run one
 * after the other and each ends with its longest solves (a max_iter QP: ~45 ms on 4096 x (N=7, H=50)) alone on their SMs.
 * With a cap of num_sms / L, L batches are co-resident, a batch lasts L times longer (about as long as its longest solve),
 * and the stragglers of one batch overlap with the bulk of the others: 4096 x (N=7, H=50), 12 handles, L = 4 (37 CTAs):
 * 315 k -> 337 k QP/s steady state.  Needs more handles in flight than L.  A caller that feeds host buffers and knows its
 * own queue widens the share of the first and last batches (bench.py, lane_ctas: L = min(4, batches already in flight + 1,
 * batches still to come + 1) in the host-buffer leg; a constant L where every step is queued up front). */
int gomp_batch_set_max_ctas(gomp_batch *h, int32_t max_ctas);

/* Device-side views of a handle, for callers that keep an SQP round on the GPU without linking the CUDA runtime
 * themselves (gomp_problem.h, device section): the handle's own copies of A_val / l / u (NULL while that input is
 * borrowed from the caller), the primal results of the last solve [batch, n] and the stream all of the handle's
 * work is ordered on.  Any out-pointer may be NULL.  After rewriting A_val / l / u in place on that stream, call
 * gomp_batch_update with the same pointers: nothing is copied, the next solve re-scales and re-factors. */
int gomp_batch_device_views(gomp_batch *h, double **A_val, double **l, double **u, const double **x, void **stream);

/* Plain device-memory helpers for such callers: allocate / free / copy (cudaMemcpyDefault on `stream`, then wait). */
int gomp_device_malloc(void **ptr, int64_t bytes, int32_t device);
int gomp_device_free(void *ptr);
int gomp_device_memcpy(void *dst, const void *src, int64_t bytes, void *stream);

int gomp_batch_info(gomp_batch *h, gomp_info *info);
int gomp_batch_sync(gomp_batch *h);
int gomp_batch_destroy(gomp_batch *h);

/* ---- one batch over several GPUs of one process (the QPs are independent, so the batch shards
 * with no exchange during the solve; the only data movement between devices is the gather of the results).
 * Problem b lives on shard b % n_shards ("interleaved": consecutive problems of a GOMP batch are similar, so
 * every device sees the same mix of short and long solves and none of them holds the tail).  Every entry point
 * takes and returns whole-batch arrays in the caller's order, host or device pointers (a device pointer may
 * live on any GPU: the copies go through peer access / NVLink when the devices allow it); inputs are staged
 * onto their shard's device with strided copies and the shards' results are scattered back the same way.
 * Calls mirror gomp_batch_*: what QPSolver's members are for one QP (reference src/osqp-wrapper.h:16-54),
 * these are for a batch spread over `n_devices` GPUs. */
typedef struct gomp_mg gomp_mg;

/* devices: CUDA ordinals [n_devices] (NULL: 0 .. n_devices-1; n_devices <= 0: every visible device).
 * settings->device and ->stream are ignored (one stream per shard, owned by the handle). */
int gomp_mg_create(gomp_mg **out, const gomp_settings *settings, const int32_t *devices, int32_t n_devices,
                   int64_t batch, int64_t n, int64_t m, const int64_t *P_colptr, const int64_t *P_rowidx,
                   const double *P_val, const int64_t *A_colptr, const int64_t *A_rowidx);
int gomp_mg_setup(gomp_mg *g, const double *A_val, const double *l, const double *u, const double *q);
int gomp_mg_warm_start_x(gomp_mg *g, const double *x0);
int gomp_mg_update(gomp_mg *g, const double *A_val, const double *l, const double *u, const double *q);
int gomp_mg_solve(gomp_mg *g); /* launches every shard's solve; asynchronous */
/* waits for every shard and gathers: x [batch, n], y [batch, m], exit_code / iters / rho_updates [batch] */
int gomp_mg_get(gomp_mg *g, double *x, double *y, int32_t *exit_code, int32_t *iters, int32_t *rho_updates);
/* shard k: its gomp_batch handle (for gomp_batch_info etc.), device ordinal and number of problems */
int gomp_mg_shard(gomp_mg *g, int32_t k, gomp_batch **handle, int32_t *device, int64_t *count);
int gomp_mg_num_shards(gomp_mg *g);
int gomp_mg_destroy(gomp_mg *g);

/* Measured roofline denominators of the current device for this kernel class: FP64 FMA throughput
 * (TFLOP/s) and aggregate shared-memory load bandwidth (GB/s).  Either pointer may be NULL. */
int gomp_measure_peaks(double *fp64_tflops, double *smem_gbs);

const char *gomp_last_error_string(void);
int gomp_device_count(void);
const char *gomp_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GOMP_CUDA_H */
