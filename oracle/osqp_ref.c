/*
 * oracle/osqp_ref.c -- CPU restatement of OSQP 0.6.x (ADMM + quasi-definite LDL^T).
 *
 * TEST INFRASTRUCTURE ONLY (see osqp_ref.h).  PARITY UNPINNED: the arithmetic restated
 * here is third-party code that the reference fetches at configure time
 * (reference src/CMakeLists.txt:25-30: google/osqp-cpp, no GIT_TAG -> OSQP 0.6.x + QDLDL
 * + AMD) and that is absent from /root/reference and from this machine.  What is pinned
 * by the reference itself is only the call sequence at the QPSolver seam:
 *   src/osqp-wrapper.h:16-31  ctor  -> q = 0, default settings, Init        (oref_setup)
 *   src/osqp-wrapper.h:45-49  setWarmStart -> SetPrimalWarmStart             (oref_warm_start_x)
 *   src/osqp-wrapper.h:51-54  solve -> Solve + primal_solution()             (oref_solve, oref_x)
 *   src/osqp-wrapper.h:33-43  update -> UpdateConstraintMatrix, SetBounds    (oref_update_A, oref_update_bounds)
 * and the SQP loop that drives it, src/gomp-solver.h:61-88.
 *
 * Everything marked [MEM] follows SURVEY.md 3.4 (recollection of the OSQP 0.6 sources and
 * the OSQP paper); it is an independent restatement, written from the algorithm, with its
 * own data structures.  The linear-system step is a quasi-definite KKT LDL^T in the style
 * of QDLDL (elimination tree + up-looking numeric factorisation, the classical sparse LDL
 * algorithm), with a minimum-degree ordering standing in for AMD.
 *
 * Deliberate deviations from genuine OSQP, all documented in DESIGN.md:
 *  - adaptive_rho_interval is PINNED (default 25).  OSQP's default 0 derives the interval
 *    from wall-clock time, which makes iteration counts machine dependent.
 *  - the fill-reducing ordering is exact minimum degree, not AMD (affects speed only).
 *  - no polish, no time limit, no printing (reference sets verbose=true; output only).
 */
#include "osqp_ref.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* ---- OSQP constants [MEM: osqp/include/constants.h] ---- */
#define RHO_MIN 1e-06
#define RHO_MAX 1e06
#define RHO_EQ_OVER_RHO_INEQ 1e03
#define RHO_TOL 1e-04
#define MIN_SCALING 1e-04
#define MAX_SCALING 1e+04
#define OSQP_INFTY 1e30
#define OSQP_DIVISION_TOL (1.0 / OSQP_INFTY)

typedef oref_int I;

typedef struct {
    I m, n;   /* rows, cols */
    I *p, *i; /* column pointers (n+1), row indices */
    double *x;
} csc_t;

struct oref_solver {
    I n, m, N;
    oref_settings st;
    csc_t P, A; /* P: upper triangle only.  Both scaled in place. */
    double *q, *l, *u;
    double *D, *E, *Dinv, *Einv, c, cinv;
    double *D_temp, *D_temp_A, *E_temp;
    double *rho_vec, *rho_inv_vec;
    int *constr_type;
    double *x, *z, *y, *x_prev, *z_prev, *xz_tilde, *delta_x, *delta_y;
    double *Ax, *Px, *Aty, *Atdelta_y, *Adelta_x, *Pdelta_x;
    double *sol_x, *sol_y;
    /* KKT (permuted, upper triangular) and maps from data entries into it */
    csc_t K;
    I *perm, *iperm;
    I *PtoK, *AtoK, *rhotoK, *sigmatoK; /* sigmatoK: n entries (diagonal of (1,1) block) */
    /* LDL^T */
    I *etree, *Lnz, *Lp, *Li;
    double *Lx, *Dd, *Ddinv;
    I *iwork;
    unsigned char *bwork;
    double *fwork, *bp, *sol;
    I nnzL;
    /* info */
    int iter, status, rho_updates, nfactor;
    double pri_res, dua_res, obj_val, rho_estimate;
};

/* ------------------------------------------------------------------ small helpers */
static void *xcalloc(size_t n, size_t sz) { return calloc(n ? n : 1, sz); }
static double dmax(double a, double b) { return a > b ? a : b; }
static double dmin(double a, double b) { return a < b ? a : b; }

static double vec_norm_inf(const double *v, I n) {
    double r = 0.0;
    for (I i = 0; i < n; i++) {
        double a = fabs(v[i]);
        if (a > r) r = a;
    }
    return r;
}
static double vec_scaled_norm_inf(const double *s, const double *v, I n) {
    double r = 0.0;
    for (I i = 0; i < n; i++) {
        double a = fabs(s[i] * v[i]);
        if (a > r) r = a;
    }
    return r;
}
static void vec_set(double *v, double a, I n) {
    for (I i = 0; i < n; i++) v[i] = a;
}
/* y (+)= A x */
static void csc_mv(const csc_t *A, const double *x, double *y, int add) {
    if (!add) vec_set(y, 0.0, A->m);
    for (I j = 0; j < A->n; j++) {
        double xj = x[j];
        for (I k = A->p[j]; k < A->p[j + 1]; k++) y[A->i[k]] += A->x[k] * xj;
    }
}
/* y (+)= A' x ; skip_diag for the strictly-lower completion of a triu matrix */
static void csc_tmv(const csc_t *A, const double *x, double *y, int add, int skip_diag) {
    if (!add) vec_set(y, 0.0, A->n);
    for (I j = 0; j < A->n; j++) {
        double acc = y[j];
        for (I k = A->p[j]; k < A->p[j + 1]; k++) {
            I r = A->i[k];
            if (skip_diag && r == j) continue;
            acc += A->x[k] * x[r];
        }
        y[j] = acc;
    }
}
static void sym_triu_mv(const csc_t *P, const double *x, double *y) {
    csc_mv(P, x, y, 0);
    csc_tmv(P, x, y, 1, 1);
}
static void csc_premult_diag(csc_t *A, const double *d) {
    for (I j = 0; j < A->n; j++)
        for (I k = A->p[j]; k < A->p[j + 1]; k++) A->x[k] *= d[A->i[k]];
}
static void csc_postmult_diag(csc_t *A, const double *d) {
    for (I j = 0; j < A->n; j++)
        for (I k = A->p[j]; k < A->p[j + 1]; k++) A->x[k] *= d[j];
}
static void csc_scale(csc_t *A, double c) {
    for (I k = 0; k < A->p[A->n]; k++) A->x[k] *= c;
}
static void csc_inf_norm_cols(const csc_t *A, double *E) {
    for (I j = 0; j < A->n; j++) {
        double r = 0.0;
        for (I k = A->p[j]; k < A->p[j + 1]; k++) r = dmax(r, fabs(A->x[k]));
        E[j] = r;
    }
}
static void csc_inf_norm_rows(const csc_t *A, double *E) {
    vec_set(E, 0.0, A->m);
    for (I j = 0; j < A->n; j++)
        for (I k = A->p[j]; k < A->p[j + 1]; k++) E[A->i[k]] = dmax(E[A->i[k]], fabs(A->x[k]));
}
static void csc_inf_norm_cols_sym_triu(const csc_t *P, double *E) {
    vec_set(E, 0.0, P->n);
    for (I j = 0; j < P->n; j++)
        for (I k = P->p[j]; k < P->p[j + 1]; k++) {
            I i = P->i[k];
            double a = fabs(P->x[k]);
            E[j] = dmax(a, E[j]);
            if (i != j) E[i] = dmax(a, E[i]);
        }
}
static void limit_scaling(double *D, I n) {
    for (I i = 0; i < n; i++) {
        D[i] = D[i] < MIN_SCALING ? 1.0 : D[i];
        D[i] = D[i] > MAX_SCALING ? MAX_SCALING : D[i];
    }
}

/* ------------------------------------------------------------------ Ruiz equilibration
 * [MEM: osqp/src/scaling.c scale_data; OSQP paper Alg. 2]: `scaling` passes of
 *   D_t = 1/sqrt(limit(max(||P_col||inf, ||A_col||inf))),  E_t = 1/sqrt(limit(||A_row||inf)),
 *   P <- D_t P D_t, A <- E_t A D_t, q <- D_t q, then the cost normalisation
 *   c_t = 1/limit(max(mean_j ||P_col j||inf, limit(||q||inf))),  P, q <- c_t P, c_t q
 * inside every pass.  Finally l, u <- E l, E u. */
static void scale_data(oref_solver *s) {
    I n = s->n, m = s->m;
    s->c = 1.0;
    vec_set(s->D, 1.0, n);
    vec_set(s->Dinv, 1.0, n);
    vec_set(s->E, 1.0, m);
    vec_set(s->Einv, 1.0, m);
    for (int it = 0; it < s->st.scaling; it++) {
        csc_inf_norm_cols_sym_triu(&s->P, s->D_temp);
        csc_inf_norm_cols(&s->A, s->D_temp_A);
        for (I j = 0; j < n; j++) s->D_temp[j] = dmax(s->D_temp[j], s->D_temp_A[j]);
        csc_inf_norm_rows(&s->A, s->E_temp);
        limit_scaling(s->D_temp, n);
        limit_scaling(s->E_temp, m);
        for (I j = 0; j < n; j++) s->D_temp[j] = 1.0 / sqrt(s->D_temp[j]);
        for (I i = 0; i < m; i++) s->E_temp[i] = 1.0 / sqrt(s->E_temp[i]);
        csc_premult_diag(&s->P, s->D_temp);
        csc_postmult_diag(&s->P, s->D_temp);
        csc_premult_diag(&s->A, s->E_temp);
        csc_postmult_diag(&s->A, s->D_temp);
        for (I j = 0; j < n; j++) s->q[j] *= s->D_temp[j];
        for (I j = 0; j < n; j++) s->D[j] *= s->D_temp[j];
        for (I i = 0; i < m; i++) s->E[i] *= s->E_temp[i];
        /* cost normalisation */
        csc_inf_norm_cols_sym_triu(&s->P, s->D_temp);
        double c_temp = 0.0;
        for (I j = 0; j < n; j++) c_temp += s->D_temp[j];
        c_temp /= (double)n;
        double nq = vec_norm_inf(s->q, n);
        limit_scaling(&nq, 1);
        c_temp = dmax(c_temp, nq);
        limit_scaling(&c_temp, 1);
        c_temp = 1.0 / c_temp;
        csc_scale(&s->P, c_temp);
        for (I j = 0; j < n; j++) s->q[j] *= c_temp;
        s->c *= c_temp;
    }
    s->cinv = 1.0 / s->c;
    for (I j = 0; j < n; j++) s->Dinv[j] = 1.0 / s->D[j];
    for (I i = 0; i < m; i++) s->Einv[i] = 1.0 / s->E[i];
    for (I i = 0; i < m; i++) {
        s->l[i] *= s->E[i];
        s->u[i] *= s->E[i];
    }
}
/* [MEM: scaling.c unscale_data] */
static void unscale_data(oref_solver *s) {
    csc_scale(&s->P, s->cinv);
    csc_premult_diag(&s->P, s->Dinv);
    csc_postmult_diag(&s->P, s->Dinv);
    for (I j = 0; j < s->n; j++) s->q[j] = s->Dinv[j] * (s->q[j] * s->cinv);
    csc_premult_diag(&s->A, s->Einv);
    csc_postmult_diag(&s->A, s->Dinv);
    for (I i = 0; i < s->m; i++) {
        s->l[i] *= s->Einv[i];
        s->u[i] *= s->Einv[i];
    }
}

/* ------------------------------------------------------------------ rho_vec
 * [MEM: auxil.c set_rho_vec / update_rho_vec] per-constraint step size:
 * loose rows -> RHO_MIN, equality rows -> 1e3*rho, otherwise rho. */
static int classify(double l, double u) {
    if (l < -OSQP_INFTY * MIN_SCALING && u > OSQP_INFTY * MIN_SCALING) return -1;
    if (u - l < RHO_TOL) return 1;
    return 0;
}
static void rho_of_type(const oref_solver *s, int t, double *rho, double *rinv) {
    if (t == -1) {
        *rho = RHO_MIN;
        *rinv = 1.0 / RHO_MIN;
    } else if (t == 1) {
        *rho = RHO_EQ_OVER_RHO_INEQ * s->st.rho;
        *rinv = 1.0 / *rho;
    } else {
        *rho = s->st.rho;
        *rinv = 1.0 / s->st.rho;
    }
}
static void set_rho_vec(oref_solver *s) {
    s->st.rho = dmin(dmax(s->st.rho, RHO_MIN), RHO_MAX);
    for (I i = 0; i < s->m; i++) {
        s->constr_type[i] = classify(s->l[i], s->u[i]);
        rho_of_type(s, s->constr_type[i], &s->rho_vec[i], &s->rho_inv_vec[i]);
    }
}

/* ------------------------------------------------------------------ minimum-degree ordering
 * Exact minimum degree on the elimination graph (clique formation), degree buckets with
 * doubly linked lists.  Stand-in for AMD [MEM: OSQP calls amd_l_order on the KKT pattern]. */
typedef struct {
    I *v;
    I len, cap;
} ilist;
static void il_push(ilist *a, I x) {
    if (a->len == a->cap) {
        a->cap = a->cap ? 2 * a->cap : 8;
        a->v = (I *)realloc(a->v, (size_t)a->cap * sizeof(I));
    }
    a->v[a->len++] = x;
}
static void md_unlink(I v, I *head, I *next, I *prev, const I *deg) {
    if (prev[v] >= 0)
        next[prev[v]] = next[v];
    else
        head[deg[v]] = next[v];
    if (next[v] >= 0) prev[next[v]] = prev[v];
}
static void md_link(I v, I *head, I *next, I *prev, const I *deg) {
    I d = deg[v];
    prev[v] = -1;
    next[v] = head[d];
    if (head[d] >= 0) prev[head[d]] = v;
    head[d] = v;
}
static int min_degree(I N, ilist *adj, I *perm) {
    I *deg = (I *)xcalloc((size_t)N, sizeof(I));
    I *head = (I *)xcalloc((size_t)N + 1, sizeof(I));
    I *next = (I *)xcalloc((size_t)N, sizeof(I));
    I *prev = (I *)xcalloc((size_t)N, sizeof(I));
    I *mark = (I *)xcalloc((size_t)N, sizeof(I));
    for (I d = 0; d <= N; d++) head[d] = -1;
    for (I v = 0; v < N; v++) mark[v] = -1;
    for (I v = 0; v < N; v++) {
        deg[v] = adj[v].len;
        md_link(v, head, next, prev, deg);
    }
    I mind = 0, stamp = 0;
    ilist tmp = {0, 0, 0};
    for (I k = 0; k < N; k++) {
        while (mind <= N && head[mind] < 0) mind++;
        I v = head[mind];
        md_unlink(v, head, next, prev, deg);
        perm[k] = v;
        ilist *nv = &adj[v];
        for (I a = 0; a < nv->len; a++) {
            I u = nv->v[a];
            /* adj[u] <- (adj[u] \ {v}) U (adj[v] \ {u}) */
            stamp++;
            tmp.len = 0;
            mark[u] = stamp;
            mark[v] = stamp;
            for (I b = 0; b < adj[u].len; b++) {
                I w = adj[u].v[b];
                if (w == v) continue;
                mark[w] = stamp;
                il_push(&tmp, w);
            }
            for (I b = 0; b < nv->len; b++) {
                I w = nv->v[b];
                if (mark[w] != stamp) {
                    mark[w] = stamp;
                    il_push(&tmp, w);
                }
            }
            md_unlink(u, head, next, prev, deg);
            adj[u].len = 0;
            for (I b = 0; b < tmp.len; b++) il_push(&adj[u], tmp.v[b]);
            deg[u] = adj[u].len;
            md_link(u, head, next, prev, deg);
            if (deg[u] < mind) mind = deg[u];
        }
        free(nv->v);
        nv->v = 0;
        nv->len = nv->cap = 0;
    }
    free(tmp.v);
    free(deg);
    free(head);
    free(next);
    free(prev);
    free(mark);
    return 0;
}

int oref_kkt_ordering(I n, I m, const I *Pp, const I *Pi, const I *Ap, const I *Ai, I *perm) {
    I N = n + m;
    ilist *adj = (ilist *)xcalloc((size_t)N, sizeof(ilist));
    /* P couples (i,j); use a per-column marker to drop duplicates from a full-symmetric P */
    for (I j = 0; j < n; j++)
        for (I k = Pp[j]; k < Pp[j + 1]; k++) {
            I i = Pi[k];
            if (i < j) { /* upper entry: add both directions once */
                il_push(&adj[i], j);
                il_push(&adj[j], i);
            }
        }
    for (I j = 0; j < n; j++)
        for (I k = Ap[j]; k < Ap[j + 1]; k++) {
            I r = n + Ai[k];
            il_push(&adj[j], r);
            il_push(&adj[r], j);
        }
    int rc = min_degree(N, adj, perm);
    for (I v = 0; v < N; v++) free(adj[v].v);
    free(adj);
    return rc;
}

/* ------------------------------------------------------------------ KKT assembly
 * [MEM: osqp lin_sys/direct/qdldl kkt.c form_KKT]  K = [[P + sigma I, A'], [A, -diag(1/rho)]],
 * upper triangle, symmetrically permuted by `perm`; index maps let update_A / rho updates
 * rewrite values in place. */
static int build_kkt(oref_solver *s) {
    I n = s->n, m = s->m, N = s->N;
    I nnzP = s->P.p[n], nnzA = s->A.p[n];
    /* does P have an explicit diagonal entry in column j? */
    I *diagpos = (I *)xcalloc((size_t)n, sizeof(I));
    I extra = 0;
    for (I j = 0; j < n; j++) {
        diagpos[j] = -1;
        for (I k = s->P.p[j]; k < s->P.p[j + 1]; k++)
            if (s->P.i[k] == j) diagpos[j] = k;
        if (diagpos[j] < 0) extra++;
    }
    I nnzK = nnzP + extra + nnzA + m;
    I *tr = (I *)xcalloc((size_t)nnzK, sizeof(I)), *tc = (I *)xcalloc((size_t)nnzK, sizeof(I));
    I t = 0;
    I *srcP = (I *)xcalloc((size_t)nnzP, sizeof(I));
    I *srcA = (I *)xcalloc((size_t)nnzA, sizeof(I));
    I *srcS = (I *)xcalloc((size_t)n, sizeof(I));
    I *srcR = (I *)xcalloc((size_t)m, sizeof(I));
    for (I j = 0; j < n; j++) {
        for (I k = s->P.p[j]; k < s->P.p[j + 1]; k++) {
            tr[t] = s->P.i[k];
            tc[t] = j;
            srcP[k] = t++;
        }
        if (diagpos[j] >= 0)
            srcS[j] = srcP[diagpos[j]];
        else {
            tr[t] = j;
            tc[t] = j;
            srcS[j] = t++;
        }
    }
    for (I j = 0; j < n; j++)
        for (I k = s->A.p[j]; k < s->A.p[j + 1]; k++) {
            tr[t] = j;
            tc[t] = n + s->A.i[k];
            srcA[k] = t++;
        }
    for (I i = 0; i < m; i++) {
        tr[t] = n + i;
        tc[t] = n + i;
        srcR[i] = t++;
    }
    /* permute + compress to CSC (upper) */
    s->K.m = s->K.n = N;
    s->K.p = (I *)xcalloc((size_t)N + 1, sizeof(I));
    s->K.i = (I *)xcalloc((size_t)nnzK, sizeof(I));
    s->K.x = (double *)xcalloc((size_t)nnzK, sizeof(double));
    I *pos = (I *)xcalloc((size_t)nnzK, sizeof(I));
    for (I e = 0; e < nnzK; e++) {
        I a = s->iperm[tr[e]], b = s->iperm[tc[e]];
        I col = a > b ? a : b;
        s->K.p[col + 1]++;
    }
    for (I j = 0; j < N; j++) s->K.p[j + 1] += s->K.p[j];
    I *fill = (I *)xcalloc((size_t)N, sizeof(I));
    for (I e = 0; e < nnzK; e++) {
        I a = s->iperm[tr[e]], b = s->iperm[tc[e]];
        I col = a > b ? a : b, row = a > b ? b : a;
        I d = s->K.p[col] + fill[col]++;
        s->K.i[d] = row;
        pos[e] = d;
    }
    s->PtoK = (I *)xcalloc((size_t)nnzP, sizeof(I));
    s->AtoK = (I *)xcalloc((size_t)nnzA, sizeof(I));
    s->sigmatoK = (I *)xcalloc((size_t)n, sizeof(I));
    s->rhotoK = (I *)xcalloc((size_t)m, sizeof(I));
    for (I k = 0; k < nnzP; k++) s->PtoK[k] = pos[srcP[k]];
    for (I k = 0; k < nnzA; k++) s->AtoK[k] = pos[srcA[k]];
    for (I j = 0; j < n; j++) s->sigmatoK[j] = pos[srcS[j]];
    for (I i = 0; i < m; i++) s->rhotoK[i] = pos[srcR[i]];
    free(tr); free(tc); free(srcP); free(srcA); free(srcS); free(srcR);
    free(pos); free(fill); free(diagpos);
    return 0;
}
static void kkt_write_values(oref_solver *s) {
    I n = s->n, m = s->m;
    memset(s->K.x, 0, (size_t)s->K.p[s->N] * sizeof(double));
    for (I k = 0; k < s->P.p[n]; k++) s->K.x[s->PtoK[k]] = s->P.x[k];
    for (I j = 0; j < n; j++) s->K.x[s->sigmatoK[j]] += s->st.sigma;
    for (I k = 0; k < s->A.p[n]; k++) s->K.x[s->AtoK[k]] = s->A.x[k];
    for (I i = 0; i < m; i++) s->K.x[s->rhotoK[i]] = -s->rho_inv_vec[i];
}

/* ------------------------------------------------------------------ sparse LDL^T
 * Classical up-looking sparse LDL^T on an upper-triangular CSC matrix (the algorithm QDLDL
 * implements [MEM]): elimination tree + column counts, then row-by-row numeric phase using
 * the etree reach.  Quasi-definite K never needs pivoting. */
static int ldl_symbolic(oref_solver *s) {
    I N = s->N;
    const csc_t *K = &s->K;
    I *work = s->iwork;
    for (I i = 0; i < N; i++) {
        work[i] = -1; /* flag */
        s->Lnz[i] = 0;
        s->etree[i] = -1;
    }
    for (I j = 0; j < N; j++) {
        work[j] = j;
        for (I p = K->p[j]; p < K->p[j + 1]; p++) {
            I i = K->i[p];
            if (i > j) return -1; /* not upper triangular */
            while (work[i] != j) {
                if (s->etree[i] == -1) s->etree[i] = j;
                s->Lnz[i]++;
                work[i] = j;
                i = s->etree[i];
            }
        }
    }
    s->Lp[0] = 0;
    for (I i = 0; i < N; i++) s->Lp[i + 1] = s->Lp[i] + s->Lnz[i];
    s->nnzL = s->Lp[N];
    return 0;
}
static int ldl_numeric(oref_solver *s) {
    I N = s->N;
    const csc_t *K = &s->K;
    I *flag = s->iwork, *pattern = s->iwork + N, *lnz = s->iwork + 2 * N;
    double *Y = s->fwork;
    for (I k = 0; k < N; k++) {
        Y[k] = 0.0;
        I top = N;
        flag[k] = k;
        lnz[k] = 0;
        for (I p = K->p[k]; p < K->p[k + 1]; p++) {
            I i = K->i[p];
            Y[i] += K->x[p];
            I len = 0;
            for (; flag[i] != k; i = s->etree[i]) {
                pattern[len++] = i; /* bottom of the array is scratch; never meets the stack */
                flag[i] = k;
            }
            while (len > 0) pattern[--top] = pattern[--len];
        }
        double dk = Y[k];
        Y[k] = 0.0;
        for (; top < N; top++) {
            I i = pattern[top];
            double yi = Y[i];
            Y[i] = 0.0;
            I p2 = s->Lp[i] + lnz[i];
            for (I p = s->Lp[i]; p < p2; p++) Y[s->Li[p]] -= s->Lx[p] * yi;
            double lki = yi * s->Ddinv[i];
            dk -= lki * yi;
            s->Li[p2] = k;
            s->Lx[p2] = lki;
            lnz[i]++;
        }
        if (dk == 0.0) return -1;
        s->Dd[k] = dk;
        s->Ddinv[k] = 1.0 / dk;
    }
    s->nfactor++;
    return 0;
}
/* x <- K^{-1} b, with permutation [MEM: qdldl_interface.c LDLSolve] */
static void ldl_solve(oref_solver *s, double *sol, const double *b) {
    I N = s->N;
    double *bp = s->bp;
    for (I i = 0; i < N; i++) bp[i] = b[s->perm[i]];
    for (I i = 0; i < N; i++) {
        double v = bp[i];
        for (I p = s->Lp[i]; p < s->Lp[i + 1]; p++) bp[s->Li[p]] -= s->Lx[p] * v;
    }
    for (I i = 0; i < N; i++) bp[i] *= s->Ddinv[i];
    for (I i = N - 1; i >= 0; i--) {
        double v = bp[i];
        for (I p = s->Lp[i]; p < s->Lp[i + 1]; p++) v -= s->Lx[p] * bp[s->Li[p]];
        bp[i] = v;
    }
    for (I i = 0; i < N; i++) sol[s->perm[i]] = bp[i];
}
static int refactor(oref_solver *s) {
    kkt_write_values(s);
    return ldl_numeric(s);
}

/* ------------------------------------------------------------------ setup / cleanup */
void oref_default_settings(oref_settings *st) {
    st->rho = 0.1;
    st->sigma = 1e-6;
    st->alpha = 1.6;
    st->scaling = 10;
    st->adaptive_rho = 1;
    st->adaptive_rho_interval = 25; /* PINNED; genuine OSQP default is 0 (wall-clock) */
    st->adaptive_rho_tolerance = 5.0;
    st->max_iter = 4000;
    st->eps_abs = 1e-3;
    st->eps_rel = 1e-3;
    st->eps_prim_inf = 1e-4;
    st->eps_dual_inf = 1e-4;
    st->scaled_termination = 0;
    st->check_termination = 25;
    st->warm_start = 1;
}

static double *dalloc(I n) { return (double *)xcalloc((size_t)n, sizeof(double)); }

oref_solver *oref_setup(I n, I m, const I *Pp, const I *Pi, const double *Px, const double *q,
                        const I *Ap, const I *Ai, const double *Ax, const double *l,
                        const double *u, const oref_settings *settings, const I *perm) {
    for (I i = 0; i < m; i++)
        if (l[i] > u[i]) return NULL; /* osqp validate_data */
    oref_solver *s = (oref_solver *)xcalloc(1, sizeof(oref_solver));
    s->n = n;
    s->m = m;
    s->N = n + m;
    s->st = *settings;
    /* P: keep the upper triangle only (osqp-cpp takes triangularView<Upper>) */
    I nnzP = 0;
    for (I j = 0; j < n; j++)
        for (I k = Pp[j]; k < Pp[j + 1]; k++)
            if (Pi[k] <= j) nnzP++;
    s->P.m = s->P.n = n;
    s->P.p = (I *)xcalloc((size_t)n + 1, sizeof(I));
    s->P.i = (I *)xcalloc((size_t)nnzP, sizeof(I));
    s->P.x = dalloc(nnzP);
    I t = 0;
    for (I j = 0; j < n; j++) {
        s->P.p[j] = t;
        for (I k = Pp[j]; k < Pp[j + 1]; k++)
            if (Pi[k] <= j) {
                s->P.i[t] = Pi[k];
                s->P.x[t++] = Px[k];
            }
    }
    s->P.p[n] = t;
    I nnzA = Ap[n];
    s->A.m = m;
    s->A.n = n;
    s->A.p = (I *)xcalloc((size_t)n + 1, sizeof(I));
    s->A.i = (I *)xcalloc((size_t)nnzA, sizeof(I));
    s->A.x = dalloc(nnzA);
    memcpy(s->A.p, Ap, (size_t)(n + 1) * sizeof(I));
    memcpy(s->A.i, Ai, (size_t)nnzA * sizeof(I));
    memcpy(s->A.x, Ax, (size_t)nnzA * sizeof(double));
    s->q = dalloc(n);
    if (q) memcpy(s->q, q, (size_t)n * sizeof(double));
    s->l = dalloc(m);
    s->u = dalloc(m);
    for (I i = 0; i < m; i++) { /* osqp-cpp clips bounds to +-OSQP_INFTY [MEM] */
        s->l[i] = dmax(l[i], -OSQP_INFTY);
        s->u[i] = dmin(u[i], OSQP_INFTY);
    }
    s->D = dalloc(n); s->Dinv = dalloc(n); s->E = dalloc(m); s->Einv = dalloc(m);
    s->D_temp = dalloc(n); s->D_temp_A = dalloc(n); s->E_temp = dalloc(m);
    s->rho_vec = dalloc(m); s->rho_inv_vec = dalloc(m);
    s->constr_type = (int *)xcalloc((size_t)m, sizeof(int));
    s->x = dalloc(n); s->z = dalloc(m); s->y = dalloc(m);
    s->x_prev = dalloc(n); s->z_prev = dalloc(m); s->xz_tilde = dalloc(n + m);
    s->delta_x = dalloc(n); s->delta_y = dalloc(m);
    s->Ax = dalloc(m); s->Px = dalloc(n); s->Aty = dalloc(n);
    s->Atdelta_y = dalloc(n); s->Adelta_x = dalloc(m); s->Pdelta_x = dalloc(n);
    s->sol_x = dalloc(n); s->sol_y = dalloc(m);

    if (s->st.scaling) {
        scale_data(s);
    } else {
        s->c = s->cinv = 1.0;
        vec_set(s->D, 1.0, n); vec_set(s->Dinv, 1.0, n);
        vec_set(s->E, 1.0, m); vec_set(s->Einv, 1.0, m);
    }
    set_rho_vec(s);

    I N = s->N;
    s->perm = (I *)xcalloc((size_t)N, sizeof(I));
    s->iperm = (I *)xcalloc((size_t)N, sizeof(I));
    if (perm)
        memcpy(s->perm, perm, (size_t)N * sizeof(I));
    else
        oref_kkt_ordering(n, m, s->P.p, s->P.i, s->A.p, s->A.i, s->perm);
    for (I i = 0; i < N; i++) s->iperm[s->perm[i]] = i;
    build_kkt(s);
    s->etree = (I *)xcalloc((size_t)N, sizeof(I));
    s->Lnz = (I *)xcalloc((size_t)N, sizeof(I));
    s->Lp = (I *)xcalloc((size_t)N + 1, sizeof(I));
    s->iwork = (I *)xcalloc((size_t)4 * N + 4, sizeof(I));
    s->fwork = dalloc(N);
    s->bp = dalloc(N);
    s->sol = dalloc(N);
    s->Dd = dalloc(N);
    s->Ddinv = dalloc(N);
    if (ldl_symbolic(s)) { oref_cleanup(s); return NULL; }
    s->Li = (I *)xcalloc((size_t)s->nnzL, sizeof(I));
    s->Lx = dalloc(s->nnzL);
    if (refactor(s)) { oref_cleanup(s); return NULL; }
    s->status = OREF_UNSOLVED;
    return s;
}

void oref_cleanup(oref_solver *s) {
    if (!s) return;
    free(s->P.p); free(s->P.i); free(s->P.x);
    free(s->A.p); free(s->A.i); free(s->A.x);
    free(s->K.p); free(s->K.i); free(s->K.x);
    free(s->q); free(s->l); free(s->u);
    free(s->D); free(s->Dinv); free(s->E); free(s->Einv);
    free(s->D_temp); free(s->D_temp_A); free(s->E_temp);
    free(s->rho_vec); free(s->rho_inv_vec); free(s->constr_type);
    free(s->x); free(s->z); free(s->y); free(s->x_prev); free(s->z_prev); free(s->xz_tilde);
    free(s->delta_x); free(s->delta_y);
    free(s->Ax); free(s->Px); free(s->Aty); free(s->Atdelta_y); free(s->Adelta_x); free(s->Pdelta_x);
    free(s->sol_x); free(s->sol_y);
    free(s->perm); free(s->iperm); free(s->PtoK); free(s->AtoK); free(s->rhotoK); free(s->sigmatoK);
    free(s->etree); free(s->Lnz); free(s->Lp); free(s->Li); free(s->Lx);
    free(s->Dd); free(s->Ddinv); free(s->iwork); free(s->fwork); free(s->bp); free(s->sol);
    free(s);
}

/* ------------------------------------------------------------------ API updates */
/* [MEM: osqp.c osqp_warm_start_x] x <- Dinv x0 ; z <- A x ; y untouched.
 * Reached from reference src/osqp-wrapper.h:45-49. */
int oref_warm_start_x(oref_solver *s, const double *x) {
    s->st.warm_start = 1;
    for (I j = 0; j < s->n; j++) s->x[j] = s->st.scaling ? s->Dinv[j] * x[j] : x[j];
    csc_mv(&s->A, s->x, s->z, 0);
    return 0;
}
/* [MEM: osqp.c osqp_update_A, full-value form used by osqp-cpp UpdateConstraintMatrix]
 * unscale -> overwrite -> full re-scale -> numeric refactor.  Iterates x,z,y are NOT
 * re-scaled to the new D,E; rho_vec/constr_type are not recomputed here.
 * Reached from reference src/osqp-wrapper.h:36. */
int oref_update_A(oref_solver *s, const double *Ax_new) {
    if (s->st.scaling) unscale_data(s);
    memcpy(s->A.x, Ax_new, (size_t)s->A.p[s->n] * sizeof(double));
    if (s->st.scaling) scale_data(s);
    return refactor(s);
}
/* [MEM: osqp.c osqp_update_bounds + auxil.c update_rho_vec] reached from
 * reference src/osqp-wrapper.h:40. */
int oref_update_bounds(oref_solver *s, const double *l, const double *u) {
    for (I i = 0; i < s->m; i++)
        if (l[i] > u[i]) return 1;
    for (I i = 0; i < s->m; i++) {
        s->l[i] = dmax(l[i], -OSQP_INFTY);
        s->u[i] = dmin(u[i], OSQP_INFTY);
        if (s->st.scaling) {
            s->l[i] *= s->E[i];
            s->u[i] *= s->E[i];
        }
    }
    int changed = 0;
    for (I i = 0; i < s->m; i++) {
        int t = classify(s->l[i], s->u[i]);
        if (t != s->constr_type[i]) {
            s->constr_type[i] = t;
            rho_of_type(s, t, &s->rho_vec[i], &s->rho_inv_vec[i]);
            changed = 1;
        }
    }
    if (changed) return refactor(s);
    return 0;
}

/* ------------------------------------------------------------------ residuals / termination */
/* [MEM: auxil.c compute_pri_res] NB: z_prev is used as scratch and later read by the
 * rho estimate, exactly as in OSQP. */
static double compute_pri_res(oref_solver *s) {
    csc_mv(&s->A, s->x, s->Ax, 0);
    for (I i = 0; i < s->m; i++) s->z_prev[i] = s->Ax[i] - s->z[i];
    if (s->st.scaling && !s->st.scaled_termination)
        return vec_scaled_norm_inf(s->Einv, s->z_prev, s->m);
    return vec_norm_inf(s->z_prev, s->m);
}
/* [MEM: auxil.c compute_dua_res] x_prev <- q + P x + A' y */
static double compute_dua_res(oref_solver *s) {
    sym_triu_mv(&s->P, s->x, s->Px);
    for (I j = 0; j < s->n; j++) s->x_prev[j] = s->q[j] + s->Px[j];
    if (s->m > 0) {
        csc_tmv(&s->A, s->y, s->Aty, 0, 0);
        for (I j = 0; j < s->n; j++) s->x_prev[j] += s->Aty[j];
    }
    if (s->st.scaling && !s->st.scaled_termination)
        return s->cinv * vec_scaled_norm_inf(s->Dinv, s->x_prev, s->n);
    return vec_norm_inf(s->x_prev, s->n);
}
static double compute_obj_val(oref_solver *s) {
    /* 0.5 x'Px + q'x on the scaled problem, times cinv */
    double quad = 0.0;
    for (I j = 0; j < s->n; j++)
        for (I k = s->P.p[j]; k < s->P.p[j + 1]; k++) {
            I i = s->P.i[k];
            if (i == j)
                quad += 0.5 * s->P.x[k] * s->x[i] * s->x[i];
            else
                quad += s->P.x[k] * s->x[i] * s->x[j];
        }
    double lin = 0.0;
    for (I j = 0; j < s->n; j++) lin += s->q[j] * s->x[j];
    double v = quad + lin;
    if (s->st.scaling) v *= s->cinv;
    return v;
}
static void update_info(oref_solver *s, int iter) {
    s->iter = iter;
    s->obj_val = compute_obj_val(s);
    s->pri_res = s->m == 0 ? 0.0 : compute_pri_res(s);
    s->dua_res = compute_dua_res(s);
}
static double compute_pri_tol(oref_solver *s, double eps_abs, double eps_rel) {
    double mx;
    if (s->st.scaling && !s->st.scaled_termination)
        mx = dmax(vec_scaled_norm_inf(s->Einv, s->z, s->m), vec_scaled_norm_inf(s->Einv, s->Ax, s->m));
    else
        mx = dmax(vec_norm_inf(s->z, s->m), vec_norm_inf(s->Ax, s->m));
    return eps_abs + eps_rel * mx;
}
static double compute_dua_tol(oref_solver *s, double eps_abs, double eps_rel) {
    double mx;
    if (s->st.scaling && !s->st.scaled_termination) {
        mx = vec_scaled_norm_inf(s->Dinv, s->q, s->n);
        mx = dmax(mx, vec_scaled_norm_inf(s->Dinv, s->Aty, s->n));
        mx = dmax(mx, vec_scaled_norm_inf(s->Dinv, s->Px, s->n));
        mx *= s->cinv;
    } else {
        mx = vec_norm_inf(s->q, s->n);
        mx = dmax(mx, vec_norm_inf(s->Aty, s->n));
        mx = dmax(mx, vec_norm_inf(s->Px, s->n));
    }
    return eps_abs + eps_rel * mx;
}
/* [MEM: auxil.c is_primal_infeasible] certificate test on delta_y */
static int is_primal_infeasible(oref_solver *s, double eps) {
    I m = s->m;
    for (I i = 0; i < m; i++) {
        if (s->u[i] > OSQP_INFTY * MIN_SCALING) {
            if (s->l[i] < -OSQP_INFTY * MIN_SCALING)
                s->delta_y[i] = 0.0;
            else
                s->delta_y[i] = dmin(s->delta_y[i], 0.0);
        } else if (s->l[i] < -OSQP_INFTY * MIN_SCALING) {
            s->delta_y[i] = dmax(s->delta_y[i], 0.0);
        }
    }
    double nrm;
    if (s->st.scaling && !s->st.scaled_termination) {
        for (I i = 0; i < m; i++) s->Adelta_x[i] = s->E[i] * s->delta_y[i];
        nrm = vec_norm_inf(s->Adelta_x, m);
    } else
        nrm = vec_norm_inf(s->delta_y, m);
    if (nrm > OSQP_DIVISION_TOL) {
        double lhs = 0.0;
        for (I i = 0; i < m; i++)
            lhs += s->u[i] * dmax(s->delta_y[i], 0.0) + s->l[i] * dmin(s->delta_y[i], 0.0);
        if (lhs < -eps * nrm) {
            csc_tmv(&s->A, s->delta_y, s->Atdelta_y, 0, 0);
            if (s->st.scaling && !s->st.scaled_termination)
                for (I j = 0; j < s->n; j++) s->Atdelta_y[j] *= s->Dinv[j];
            return vec_norm_inf(s->Atdelta_y, s->n) < eps * nrm;
        }
    }
    return 0;
}
/* [MEM: auxil.c is_dual_infeasible] certificate test on delta_x */
static int is_dual_infeasible(oref_solver *s, double eps) {
    double nrm, cs;
    if (s->st.scaling && !s->st.scaled_termination) {
        nrm = vec_scaled_norm_inf(s->D, s->delta_x, s->n);
        cs = s->c;
    } else {
        nrm = vec_norm_inf(s->delta_x, s->n);
        cs = 1.0;
    }
    if (nrm > OSQP_DIVISION_TOL) {
        double qdx = 0.0;
        for (I j = 0; j < s->n; j++) qdx += s->q[j] * s->delta_x[j];
        if (qdx < -cs * eps * nrm) {
            sym_triu_mv(&s->P, s->delta_x, s->Pdelta_x);
            if (s->st.scaling && !s->st.scaled_termination)
                for (I j = 0; j < s->n; j++) s->Pdelta_x[j] *= s->Dinv[j];
            if (vec_norm_inf(s->Pdelta_x, s->n) < cs * eps * nrm) {
                csc_mv(&s->A, s->delta_x, s->Adelta_x, 0);
                if (s->st.scaling && !s->st.scaled_termination)
                    for (I i = 0; i < s->m; i++) s->Adelta_x[i] *= s->Einv[i];
                for (I i = 0; i < s->m; i++) {
                    if ((s->u[i] < OSQP_INFTY * MIN_SCALING && s->Adelta_x[i] > eps * nrm) ||
                        (s->l[i] > -OSQP_INFTY * MIN_SCALING && s->Adelta_x[i] < -eps * nrm))
                        return 0;
                }
                return 1;
            }
        }
    }
    return 0;
}
/* [MEM: auxil.c check_termination] */
static int check_termination(oref_solver *s, int approximate) {
    double eps_abs = s->st.eps_abs, eps_rel = s->st.eps_rel;
    double eps_pinf = s->st.eps_prim_inf, eps_dinf = s->st.eps_dual_inf;
    int prim_res_check = 0, dual_res_check = 0, prim_inf_check = 0, dual_inf_check = 0;
    if (s->pri_res > OSQP_INFTY || s->dua_res > OSQP_INFTY) {
        s->status = OREF_NON_CVX;
        s->obj_val = NAN;
        return 1;
    }
    if (approximate) {
        eps_abs *= 10; eps_rel *= 10; eps_pinf *= 10; eps_dinf *= 10;
    }
    if (s->m == 0)
        prim_res_check = 1;
    else {
        double eps_prim = compute_pri_tol(s, eps_abs, eps_rel);
        if (s->pri_res < eps_prim)
            prim_res_check = 1;
        else
            prim_inf_check = is_primal_infeasible(s, eps_pinf);
    }
    double eps_dual = compute_dua_tol(s, eps_abs, eps_rel);
    if (s->dua_res < eps_dual)
        dual_res_check = 1;
    else
        dual_inf_check = is_dual_infeasible(s, eps_dinf);

    if (prim_res_check && dual_res_check) {
        s->status = approximate ? OREF_SOLVED_INACCURATE : OREF_SOLVED;
        return 1;
    } else if (prim_inf_check) {
        s->status = approximate ? OREF_PRIMAL_INFEASIBLE_INACCURATE : OREF_PRIMAL_INFEASIBLE;
        if (s->st.scaling && !s->st.scaled_termination)
            for (I i = 0; i < s->m; i++) s->delta_y[i] *= s->E[i];
        s->obj_val = OSQP_INFTY;
        return 1;
    } else if (dual_inf_check) {
        s->status = approximate ? OREF_DUAL_INFEASIBLE_INACCURATE : OREF_DUAL_INFEASIBLE;
        if (s->st.scaling && !s->st.scaled_termination)
            for (I j = 0; j < s->n; j++) s->delta_x[j] *= s->D[j];
        s->obj_val = -OSQP_INFTY;
        return 1;
    }
    return 0;
}
/* [MEM: auxil.c compute_rho_estimate / adapt_rho + osqp.c osqp_update_rho] */
static double compute_rho_estimate(oref_solver *s) {
    double pri = vec_norm_inf(s->z_prev, s->m);
    double dua = vec_norm_inf(s->x_prev, s->n);
    double pn = dmax(vec_norm_inf(s->z, s->m), vec_norm_inf(s->Ax, s->m));
    pri /= (pn + OSQP_DIVISION_TOL);
    double dn = vec_norm_inf(s->q, s->n);
    dn = dmax(dn, vec_norm_inf(s->Aty, s->n));
    dn = dmax(dn, vec_norm_inf(s->Px, s->n));
    dua /= (dn + OSQP_DIVISION_TOL);
    double est = s->st.rho * sqrt(pri / dua);
    return dmin(dmax(est, RHO_MIN), RHO_MAX);
}
static int update_rho(oref_solver *s, double rho_new) {
    s->st.rho = dmin(dmax(rho_new, RHO_MIN), RHO_MAX);
    for (I i = 0; i < s->m; i++) {
        if (s->constr_type[i] == 0) {
            s->rho_vec[i] = s->st.rho;
            s->rho_inv_vec[i] = 1.0 / s->st.rho;
        } else if (s->constr_type[i] == 1) {
            s->rho_vec[i] = RHO_EQ_OVER_RHO_INEQ * s->st.rho;
            s->rho_inv_vec[i] = 1.0 / s->rho_vec[i];
        }
    }
    return refactor(s);
}
static int adapt_rho(oref_solver *s) {
    double rho_new = compute_rho_estimate(s);
    s->rho_estimate = rho_new;
    if (rho_new > s->st.rho * s->st.adaptive_rho_tolerance ||
        rho_new < s->st.rho / s->st.adaptive_rho_tolerance) {
        int rc = update_rho(s, rho_new);
        s->rho_updates++;
        return rc;
    }
    return 0;
}

static int has_solution(const oref_solver *s) {
    return s->status != OREF_PRIMAL_INFEASIBLE && s->status != OREF_PRIMAL_INFEASIBLE_INACCURATE &&
           s->status != OREF_DUAL_INFEASIBLE && s->status != OREF_DUAL_INFEASIBLE_INACCURATE &&
           s->status != OREF_NON_CVX;
}

/* ------------------------------------------------------------------ the ADMM loop
 * [MEM: osqp.c osqp_solve; OSQP paper Alg. 1].  Reached from reference
 * src/osqp-wrapper.h:52 (solver.Solve()). */
int oref_solve(oref_solver *s) {
    I n = s->n, m = s->m;
    const double alpha = s->st.alpha, sigma = s->st.sigma;
    int can_check = 0, iter;
    s->status = OREF_UNSOLVED;
    if (!s->st.warm_start) {
        vec_set(s->x, 0.0, n); vec_set(s->z, 0.0, m); vec_set(s->y, 0.0, m);
    }
    for (iter = 1; iter <= s->st.max_iter; iter++) {
        double *t;
        t = s->x; s->x = s->x_prev; s->x_prev = t;
        t = s->z; s->z = s->z_prev; s->z_prev = t;
        /* rhs and KKT solve: [x~; nu] */
        for (I j = 0; j < n; j++) s->xz_tilde[j] = sigma * s->x_prev[j] - s->q[j];
        for (I i = 0; i < m; i++) s->xz_tilde[n + i] = s->z_prev[i] - s->rho_inv_vec[i] * s->y[i];
        ldl_solve(s, s->sol, s->xz_tilde);
        for (I j = 0; j < n; j++) s->xz_tilde[j] = s->sol[j];
        for (I i = 0; i < m; i++) s->xz_tilde[n + i] += s->rho_inv_vec[i] * s->sol[n + i];
        /* x, z, y updates */
        for (I j = 0; j < n; j++) {
            s->x[j] = alpha * s->xz_tilde[j] + (1.0 - alpha) * s->x_prev[j];
            s->delta_x[j] = s->x[j] - s->x_prev[j];
        }
        for (I i = 0; i < m; i++) {
            double v = alpha * s->xz_tilde[n + i] + (1.0 - alpha) * s->z_prev[i] + s->rho_inv_vec[i] * s->y[i];
            s->z[i] = dmin(dmax(v, s->l[i]), s->u[i]);
        }
        for (I i = 0; i < m; i++) {
            double d = alpha * s->xz_tilde[n + i] + (1.0 - alpha) * s->z_prev[i] - s->z[i];
            d *= s->rho_vec[i];
            s->delta_y[i] = d;
            s->y[i] += d;
        }
        can_check = s->st.check_termination && (iter % s->st.check_termination == 0);
        if (can_check) {
            update_info(s, iter);
            if (check_termination(s, 0)) break;
        }
        if (s->st.adaptive_rho && s->st.adaptive_rho_interval &&
            (iter % s->st.adaptive_rho_interval == 0)) {
            if (!can_check) update_info(s, iter);
            if (adapt_rho(s)) return s->status = OREF_UNSOLVED;
        }
    }
    if (!can_check) {
        update_info(s, iter - 1);
        check_termination(s, 0);
    }
    if (s->status == OREF_UNSOLVED) {
        if (!check_termination(s, 1)) s->status = OREF_MAX_ITER_REACHED;
    }
    s->rho_estimate = compute_rho_estimate(s);
    /* store_solution */
    if (has_solution(s)) {
        for (I j = 0; j < n; j++) s->sol_x[j] = s->st.scaling ? s->D[j] * s->x[j] : s->x[j];
        for (I i = 0; i < m; i++) s->sol_y[i] = s->st.scaling ? s->cinv * s->E[i] * s->y[i] : s->y[i];
    } else {
        vec_set(s->sol_x, NAN, n);
        vec_set(s->sol_y, NAN, m);
        /* [MEM: auxil.c store_solution] normalise the infeasibility certificate to unit infinity norm */
        if (s->status == OREF_PRIMAL_INFEASIBLE || s->status == OREF_PRIMAL_INFEASIBLE_INACCURATE) {
            double nrm = vec_norm_inf(s->delta_y, m);
            for (I i = 0; i < m; i++) s->delta_y[i] *= 1.0 / nrm;
        }
        if (s->status == OREF_DUAL_INFEASIBLE || s->status == OREF_DUAL_INFEASIBLE_INACCURATE) {
            double nrm = vec_norm_inf(s->delta_x, n);
            for (I j = 0; j < n; j++) s->delta_x[j] *= 1.0 / nrm;
        }
        /* cold start for the next run (iterates cannot restart from NaN) */
        vec_set(s->x, 0.0, n); vec_set(s->z, 0.0, m); vec_set(s->y, 0.0, m);
    }
    return s->status;
}

/* ------------------------------------------------------------------ getters */
const double *oref_x(const oref_solver *s) { return s->sol_x; }
const double *oref_y(const oref_solver *s) { return s->sol_y; }
/* osqp-cpp primal_infeasibility_certificate() / dual_infeasibility_certificate(): work->delta_y / delta_x (meaningful after an
 * infeasible status only) */
const double *oref_delta_y(const oref_solver *s) { return s->delta_y; }
const double *oref_delta_x(const oref_solver *s) { return s->delta_x; }
int oref_status(const oref_solver *s) { return s->status; }
int oref_iter(const oref_solver *s) { return s->iter; }
int oref_rho_updates(const oref_solver *s) { return s->rho_updates; }
double oref_rho(const oref_solver *s) { return s->st.rho; }
double oref_pri_res(const oref_solver *s) { return s->pri_res; }
double oref_dua_res(const oref_solver *s) { return s->dua_res; }
double oref_obj_val(const oref_solver *s) { return s->obj_val; }
oref_int oref_nnz_L(const oref_solver *s) { return s->nnzL; }
int oref_num_factorizations(const oref_solver *s) { return s->nfactor; }
const double *oref_D(const oref_solver *s) { return s->D; }
const double *oref_E(const oref_solver *s) { return s->E; }
double oref_c(const oref_solver *s) { return s->c; }
/* osqp-cpp StatusToExitCode [MEM] */
int oref_exit_code(const oref_solver *s) {
    switch (s->status) {
    case OREF_SOLVED: return OREF_EXIT_OPTIMAL;
    case OREF_SOLVED_INACCURATE: return OREF_EXIT_OPTIMAL_INACCURATE;
    case OREF_PRIMAL_INFEASIBLE: return OREF_EXIT_PRIMAL_INFEASIBLE;
    case OREF_PRIMAL_INFEASIBLE_INACCURATE: return OREF_EXIT_PRIMAL_INFEASIBLE_INACCURATE;
    case OREF_DUAL_INFEASIBLE: return OREF_EXIT_DUAL_INFEASIBLE;
    case OREF_DUAL_INFEASIBLE_INACCURATE: return OREF_EXIT_DUAL_INFEASIBLE_INACCURATE;
    case OREF_MAX_ITER_REACHED: return OREF_EXIT_MAX_ITERATIONS;
    case OREF_NON_CVX: return OREF_EXIT_NON_CONVEX;
    default: return OREF_EXIT_UNKNOWN;
    }
}

/* ------------------------------------------------------------------ batched driver
 * One QP per worker thread (pthreads, dynamic work queue).  This is what the reference would
 * look like run over a batch on all host cores: src/gomp-solver.h:61-71 per problem. */
int oref_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef struct {
    I batch, n, m, nnzA, total; /* total = passes * batch work items, item w solves QP w % batch */
    const I *Pp, *Pi, *Ap, *Ai, *perm;
    const double *Px, *Ax, *l, *u, *x0;
    const oref_settings *settings;
    double *x, *y;
    int *exit_code, *iters, *rho_updates;
    volatile long long next;
    volatile int fail;
} batch_job;

static void *batch_worker(void *arg) {
    batch_job *J = (batch_job *)arg;
    for (;;) {
        I b = (I)__sync_fetch_and_add(&J->next, 1);
        if (b >= J->total) break;
        b %= J->batch;
        oref_solver *s = oref_setup(J->n, J->m, J->Pp, J->Pi, J->Px, NULL, J->Ap, J->Ai,
                                    J->Ax + b * J->nnzA, J->l + b * J->m, J->u + b * J->m,
                                    J->settings, J->perm);
        if (!s) {
            J->fail = 1;
            continue;
        }
        if (J->x0) oref_warm_start_x(s, J->x0 + b * J->n);
        oref_solve(s);
        if (J->x) memcpy(J->x + b * J->n, s->sol_x, (size_t)J->n * sizeof(double));
        if (J->y) memcpy(J->y + b * J->m, s->sol_y, (size_t)J->m * sizeof(double));
        if (J->exit_code) J->exit_code[b] = oref_exit_code(s);
        if (J->iters) J->iters[b] = s->iter;
        if (J->rho_updates) J->rho_updates[b] = s->rho_updates;
        oref_cleanup(s);
    }
    return NULL;
}

/* `passes` passes over the batch as ONE continuous work queue (bench.py's CPU arm: no thread idles at a pass boundary
 * while a long QP of the previous pass finishes; every pass is a full cold setup + warm start + solve of each QP). */
int oref_solve_batch_passes(I batch, I n, I m, const I *Pp, const I *Pi, const double *Px, const I *Ap,
                            const I *Ai, const double *Ax, const double *l, const double *u,
                            const double *x0, const oref_settings *settings, int nthreads, int passes, double *x,
                            double *y, int *exit_code, int *iters, int *rho_updates) {
    I *perm = (I *)xcalloc((size_t)(n + m), sizeof(I));
    oref_kkt_ordering(n, m, Pp, Pi, Ap, Ai, perm);
    batch_job J;
    memset(&J, 0, sizeof J);
    J.batch = batch; J.n = n; J.m = m; J.nnzA = Ap[n];
    J.total = batch * (I)(passes > 0 ? passes : 1);
    J.Pp = Pp; J.Pi = Pi; J.Ap = Ap; J.Ai = Ai; J.perm = perm;
    J.Px = Px; J.Ax = Ax; J.l = l; J.u = u; J.x0 = x0; J.settings = settings;
    J.x = x; J.y = y; J.exit_code = exit_code; J.iters = iters; J.rho_updates = rho_updates;
    if (nthreads <= 0) nthreads = oref_max_threads();
    if ((I)nthreads > J.total) nthreads = (int)J.total;
    if (nthreads <= 1) {
        batch_worker(&J);
    } else {
        pthread_t *th = (pthread_t *)xcalloc((size_t)nthreads, sizeof(pthread_t));
        for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, batch_worker, &J);
        for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
        free(th);
    }
    free(perm);
    return J.fail;
}

int oref_solve_batch(I batch, I n, I m, const I *Pp, const I *Pi, const double *Px, const I *Ap,
                     const I *Ai, const double *Ax, const double *l, const double *u,
                     const double *x0, const oref_settings *settings, int nthreads, double *x,
                     double *y, int *exit_code, int *iters, int *rho_updates) {
    return oref_solve_batch_passes(batch, n, m, Pp, Pi, Px, Ap, Ai, Ax, l, u, x0, settings, nthreads, 1, x, y, exit_code, iters,
                                   rho_updates);
}
