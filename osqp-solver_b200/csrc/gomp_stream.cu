// gomp_stream.cu -- long-horizon back end of libgomp_cuda (GOMP_LINSYS_BLOCKTRI_GLOBAL, sm_100a, fp64).
//
// The register/shared-memory resident kernel (gomp_kernel.cuh) needs the whole QP and its factor on one SM:
// that ends near H = 60 way-points for a 7-DoF arm.  Longer horizons (BASELINE.json configs[3]: H = 400;
// the reference example: H = 802, src/solver-example.cpp:12) run here: the same algorithm, the same
// reduced KKT system S = P + sigma I + A' diag(rho) A and the same two-way block-tridiagonal factorisation
// with explicit pivot-block inverses -- an exact direct solve, so the iteration counts stay those of
// OSQP -- but the QP, its iterates and the factor live in a per-CTA workspace in GLOBAL memory that stays
// L2-resident (H = 400: 1.5 MB per QP; B200 L2: 126 MB), and only what the sequential critical path
// touches is on chip:
//   * the right-hand side / x~ vector of the block solve in shared memory;
//   * the factor STREAMS through shared memory: each of the two substitution warps runs a ring of
//     kRing slots filled by TMA bulk copies (cp.async.bulk + one mbarrier per slot) issued kRing-1
//     stages ahead by the warp's own lane 0, so the L2 latency never sits on the dependent chain.
// One 512-thread CTA per QP (work queue); all threads do the sparse phases (Ruiz, A' gather, fused
// A x~ + projection + dual update, residual norms) as strided loops, the factorisation runs as two
// 224-thread Gauss-Jordan groups (one per elimination chain), the substitutions on warps 0 and 1.
// Intra-problem parallelism is what a single long QP offers: two chains + 512-wide sparse phases.
//
// Replaces the same reference calls as the resident kernel (src/osqp-wrapper.h:28,36,40,46,52); the
// algorithm follows oracle/osqp_ref.c phase by phase (tests/cpp/emulate.cpp is its sequential model).
#include <algorithm>
#include <cstdlib>

#include "gomp_stream.h"

#include <cooperative_groups.h>

namespace gomp_dev {

namespace {

constexpr int kST = 512;    // threads per CTA
constexpr int kGroup = 224; // threads per factor group
constexpr int kMaxW = 4;    // off-diagonal ELL width kept in registers by the substitution warps

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <int K, int NT = kST>
__device__ __forceinline__ void block_max(double (&v)[K], double *sRed, int tid) {
    constexpr int NW = NT / 32;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double a = v[k];
#pragma unroll
        for (int o = 16; o; o >>= 1) a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o));
        if ((tid & 31) == 0) sRed[k * NW + (tid >> 5)] = a;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double a = sRed[k * NW];
#pragma unroll
        for (int j = 1; j < NW; ++j) a = fmax(a, sRed[k * NW + j]);
        v[k] = a;
    }
    __syncthreads();
}
template <int NT = kST>
__device__ __forceinline__ double block_sum(double a, double *sRed, int tid) {
    constexpr int NW = NT / 32;
#pragma unroll
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((tid & 31) == 0) sRed[tid >> 5] = a;
    __syncthreads();
    double r = sRed[0];
#pragma unroll
    for (int j = 1; j < NW; ++j) r += sRed[j];
    __syncthreads();
    return r;
}

// sparse dot products against the pattern arrays in global memory (read-only for the kernel's lifetime)
__device__ __forceinline__ double row_dot_g(const DevSym &S, int i, const double *A, const double *vec) {
    const uint32_t rg = __ldg(S.row_range + i);
    int q = (int)(rg >> 8);
    const int q1 = q + (int)(rg & 0xFFu);
    double acc = 0.0;
    for (; q < q1; ++q) {
        const uint32_t e = __ldg(S.row_csr + q);
        acc = fma(A[e & 0xFFFFu], vec[e >> 16], acc);
    }
    return acc;
}
__device__ __forceinline__ double col_dot_g(const DevSym &S, int p, const double *A, const double *vec, double acc) {
    const uint32_t rg = __ldg(S.col_range + p);
    int k = (int)(rg >> 8);
    const int k1 = k + (int)(rg & 0xFFu);
    for (; k < k1; ++k) acc = fma(A[k], vec[__ldg(S.csc_row + k)], acc);
    return acc;
}
__device__ __forceinline__ double col_absmax_g(const DevSym &S, int p, const double *A) {
    const uint32_t rg = __ldg(S.col_range + p);
    int k = (int)(rg >> 8);
    const int k1 = k + (int)(rg & 0xFFu);
    double nrm = 0.0;
    for (; k < k1; ++k) nrm = fmax(nrm, fabs(A[k]));
    return nrm;
}
__device__ __forceinline__ double pf_absmax_g(const DevSym &S, int p, const double *P) {
    double nrm = 0.0;
    for (int q = __ldg(S.pf_ptr + p), q1 = __ldg(S.pf_ptr + p + 1); q < q1; ++q) nrm = fmax(nrm, fabs(P[__ldg(S.pf_csr + q) & 0xFFFFu]));
    return nrm;
}
__device__ __forceinline__ double pf_dot_g(const DevSym &S, int p, const double *P, const double *vec) {
    double acc = 0.0;
    for (int q = __ldg(S.pf_ptr + p), q1 = __ldg(S.pf_ptr + p + 1); q < q1; ++q) {
        const uint32_t e = __ldg(S.pf_csr + q);
        acc = fma(P[e & 0xFFFFu], vec[e >> 16], acc);
    }
    return acc;
}

// ---- the same products on the ELL views, two rows / columns per call ------------------------------------------
// The entry addresses of an ELL row do not depend on a range word, and two rows are interleaved, so a thread
// has up to 8 independent L2 loads in flight per level instead of one (these loops are L2-latency bound).
// Summation order per row / column is the CSR / CSC order, i.e. the results equal the *_g versions bit for bit.
constexpr uint32_t kNone = 0xFFFFFFFFu;
// Rows are stored longest first (symbolic.h sorts the non-empty rows by entry count), so the longest row of the 256-row block
// the smaller index of a pair falls into bounds both rows: the loops stop there instead of walking the full ELL width (most
// GOMP rows hold one to three entries, the Cartesian ones 2N) -- fewer index words fetched from L2, same entries in the same
// order.  Columns: their lengths are in the range words the loop reads anyway.
__device__ __forceinline__ int ell_row_width(const DevStream &X, int i) { return (int)X.rlen_blk[(i >> 8) < 255 ? (i >> 8) : 255]; }
#define ELL_ROW2(i1, i2, ...)                                                               \
    _Pragma("unroll 4") for (int j__ = 0, w__ = ell_row_width(X, (i1) < (i2) ? (i1) : (i2)); j__ < w__; ++j__) { \
        const uint32_t e1 = __ldg(X.row_ell + (size_t)j__ * mu + (i1));                     \
        const uint32_t e2 = __ldg(X.row_ell + (size_t)j__ * mu + (i2));                     \
        const bool h1 = e1 != kNone, h2 = e2 != kNone;                                      \
        const uint32_t k1 = h1 ? (e1 & 0xFFFFu) : 0u, k2 = h2 ? (e2 & 0xFFFFu) : 0u;         \
        const uint32_t q1 = h1 ? (e1 >> 16) : 0u, q2 = h2 ? (e2 >> 16) : 0u;                 \
        __VA_ARGS__                                                                         \
    }
__device__ __forceinline__ void row_dot_ell2(const DevStream &X, int mu, int i1, int i2, const double *A, const double *vec, double &r1,
                                             double &r2) {
    double a1 = 0.0, a2 = 0.0;
    ELL_ROW2(i1, i2, {
        const double v1 = A[k1], v2 = A[k2], w1 = vec[q1], w2 = vec[q2];
        a1 = h1 ? fma(v1, w1, a1) : a1;
        a2 = h2 ? fma(v2, w2, a2) : a2;
    })
    r1 = a1;
    r2 = a2;
}
__device__ __forceinline__ void row_absmax_ell2(const DevStream &X, int mu, int i1, int i2, const double *A, double &r1, double &r2) {
    double a1 = 0.0, a2 = 0.0;
    ELL_ROW2(i1, i2, {
        const double v1 = A[k1], v2 = A[k2];
        a1 = h1 ? fmax(a1, fabs(v1)) : a1;
        a2 = h2 ? fmax(a2, fabs(v2)) : a2;
        (void)q1; (void)q2;
    })
    r1 = a1;
    r2 = a2;
}
__device__ __forceinline__ void row_scale_ell2(const DevStream &X, int mu, int i1, int i2, bool two, double *A, double s1, double s2) {
    ELL_ROW2(i1, i2, {
        const double v1 = A[k1], v2 = A[k2];
        if (h1) A[k1] = v1 * s1;
        if (h2 && two) A[k2] = v2 * s2;
        (void)q1; (void)q2;
    })
}
// A[k] = (A[k] * row factor) * column factor of the entry's position -- OSQP's row scaling, then its column scaling, in that
// order of rounding -- and the rows' new max norms, in one traversal (the cluster Ruiz kernel: the column factors of the pass
// are complete before any row is scaled)
__device__ __forceinline__ void row_scale2_absmax_ell2(const DevStream &X, int mu, int i1, int i2, bool two, double *A, double s1, double s2,
                                                       const double *dcol, double &r1, double &r2) {
    double a1 = 0.0, a2 = 0.0;
    ELL_ROW2(i1, i2, {
        const double v1 = A[k1], v2 = A[k2], d1 = dcol[q1], d2 = dcol[q2];
        const double t1 = (v1 * s1) * d1, t2 = (v2 * s2) * d2;
        if (h1) {
            A[k1] = t1;
            a1 = fmax(a1, fabs(t1));
        }
        if (h2 && two) {
            A[k2] = t2;
            a2 = fmax(a2, fabs(t2));
        }
    })
    r1 = a1;
    r2 = a2;
}
// columns: values are the contiguous CSC range of the column, their compact rows come from the ELL view
#define ELL_COL2(p1, p2, ...)                                                               \
    const uint32_t rg1 = __ldg(S.col_range + (p1)), rg2 = __ldg(S.col_range + (p2));       \
    const int s1__ = (int)(rg1 >> 8), n1__ = (int)(rg1 & 0xFFu), s2__ = (int)(rg2 >> 8), n2__ = (int)(rg2 & 0xFFu); \
    _Pragma("unroll 4") for (int j__ = 0, w__ = n1__ > n2__ ? n1__ : n2__; j__ < w__; ++j__) { \
        const bool h1 = j__ < n1__, h2 = j__ < n2__;                                        \
        const int k1 = h1 ? s1__ + j__ : 0, k2 = h2 ? s2__ + j__ : 0;                        \
        __VA_ARGS__                                                                         \
    }
__device__ __forceinline__ void col_dot_ell2(const DevSym &S, const DevStream &X, int np, int p1, int p2, const double *A, const double *vec,
                                             double &r1, double &r2) {
    double a1 = r1, a2 = r2;
    ELL_COL2(p1, p2, {
        const int q1 = __ldg(X.col_ell + (size_t)j__ * np + p1), q2 = __ldg(X.col_ell + (size_t)j__ * np + p2);
        const double v1 = A[k1], v2 = A[k2], w1 = vec[h1 ? q1 : 0], w2 = vec[h2 ? q2 : 0];
        a1 = h1 ? fma(v1, w1, a1) : a1;
        a2 = h2 ? fma(v2, w2, a2) : a2;
    })
    r1 = a1;
    r2 = a2;
}
__device__ __forceinline__ void col_absmax2(const DevSym &S, const DevStream &X, int p1, int p2, const double *A, double &r1, double &r2) {
    double a1 = 0.0, a2 = 0.0;
    ELL_COL2(p1, p2, {
        const double v1 = A[k1], v2 = A[k2];
        a1 = h1 ? fmax(a1, fabs(v1)) : a1;
        a2 = h2 ? fmax(a2, fabs(v2)) : a2;
    })
    r1 = a1;
    r2 = a2;
}
__device__ __forceinline__ void col_scale2(const DevSym &S, const DevStream &X, int p1, int p2, bool two, double *A, double s1, double s2) {
    ELL_COL2(p1, p2, {
        const double v1 = A[k1], v2 = A[k2];
        if (h1) A[k1] = v1 * s1;
        if (h2 && two) A[k2] = v2 * s2;
    })
}
// pairs of the strided loops: (i, i2) with i2 = i + kST, clamped to i when it falls off the end
#define ROWS2(i, i2, two) for (int i = tid, i2 = (tid + kST < mu ? tid + kST : tid); i < mu; i += 2 * kST, i2 = (i + kST < mu ? i + kST : i)) if (const bool two = i + kST < mu; true)
#define COLS2(p, p2, two) for (int p = tid, p2 = (tid + kST < np ? tid + kST : tid); p < np; p += 2 * kST, p2 = (p + kST < np ? p + kST : p)) if (const bool two = p + kST < np; true)

// ---- factorisation: one pivot block, by one 224-thread group, entries in registers --------------------------
// mode bit 1: subtract Off_{t-1} Dinv_{t-1} Off_{t-1}' first; bit 2: subtract Off_t' Dinv_{t+1} Off_t.
// `dprev` (shared memory, bs*bs doubles per group) carries the inverse of the previous stage of the chain to the
// next call, so the Schur update never waits on the global copy just written.
// `offbuf` (shared memory, bs*ow doubles per group) stages the off-diagonal block of the Schur update.
// GS = threads of the group (224: one matrix entry per thread for a 14 x 14 block; 128: up to two -- the 256-thread CTAs
// of the small-QP mode run both elimination chains at once with two half-size groups).
// Off-diagonal pattern of a thread's matrix entries (r, c), for patterns shared by every block and at most two entries per
// block row / column: [pass 0 = block above, 1 = block below][entry][0..1] = neighbour index of r (ia) and of c (ib), -1 =
// none, and the positions of the two values inside the staged off-diagonal block (va, vb).  Loaded once per
// factorisation instead of once per stage (the pattern arrays live in L2).
template <int EPL>
struct SchurPattern {
    signed char ia[2][EPL][2], ib[2][EPL][2];
    short va[2][EPL][2], vb[2][EPL][2];
    bool valid;
};
template <int EPL>
__device__ __forceinline__ void load_schur_pattern(const DevSym &S, int bs, int gl, int GS, SchurPattern<EPL> &P) {
    P.valid = S.uniform_off && S.ow <= 2 && S.owT <= 2;
    if (!P.valid) return;
#pragma unroll
    for (int q = 0; q < EPL; ++q) {
        const int e = gl + q * GS, r = e < bs * bs ? e / bs : 0, c = e < bs * bs ? e - r * bs : 0;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            P.ia[0][q][s] = s < S.ow ? S.off_col[r * S.ow + s] : -1;
            P.ib[0][q][s] = s < S.ow ? S.off_col[c * S.ow + s] : -1;
            P.va[0][q][s] = (short)(r * S.ow + s);
            P.vb[0][q][s] = (short)(c * S.ow + s);
            P.ia[1][q][s] = s < S.owT ? S.offT_row[r * S.owT + s] : -1;
            P.ib[1][q][s] = s < S.owT ? S.offT_row[c * S.owT + s] : -1;
            P.va[1][q][s] = s < S.owT ? (short)S.offT_slot[r * S.owT + s] : 0;
            P.vb[1][q][s] = s < S.owT ? (short)S.offT_slot[c * S.owT + s] : 0;
        }
    }
}

template <int BS, int GS>
__device__ __forceinline__ void factor_block_stream(const DevSym &S, double *Dall, const double *Offall, double *pub, double *dprev,
                                                    double *offbuf, bool prev_valid, int t, int mode, int g, int gl,
                                                    const SchurPattern<(BS > 0 ? (BS * BS + GS - 1) / GS : (1024 + GS - 1) / GS)> *sp = nullptr) {
    constexpr int EPL = BS > 0 ? (BS * BS + GS - 1) / GS : (1024 + GS - 1) / GS; // entries per thread (bs <= 32)
    const int bs = BS > 0 ? BS : S.bs, nn = bs * bs;
    double *M = Dall + (size_t)t * nn;
    double val[EPL];
    int ri[EPL], ci[EPL];
#pragma unroll
    for (int q = 0; q < EPL; ++q) {
        const int e = gl + q * GS;
        ri[q] = e < nn ? e / bs : -1;
        ci[q] = e < nn ? e - ri[q] * bs : -1;
        val[q] = e < nn ? M[e] : 0.0;
    }
#pragma unroll 1
    for (int pass = 1; pass <= 2; ++pass) {
        if (!(mode & pass)) continue;
        const bool above = pass == 1;
        const int tb = above ? t - 1 : t;
        const double *Dn = prev_valid ? dprev : Dall + (size_t)(above ? t - 1 : t + 1) * nn;
        const double *gblk = Offall + (size_t)tb * bs * S.ow;
        if (pass == 2 && (mode & 1)) named_bar_sync(4 + g, GS); // middle stage: the first pass is done with offbuf
        for (int e = gl; e < bs * S.ow; e += GS) offbuf[e] = gblk[e];
        named_bar_sync(4 + g, GS);
        const double *blk = offbuf;
        const int tp = S.uniform_off ? 0 : tb; // pattern block (L1-resident when every block shares one pattern)
        const int cnt = above ? S.ow : S.owT;
        if (sp != nullptr && sp->valid) {
            const int ps = above ? 0 : 1;
#pragma unroll
            for (int q = 0; q < EPL; ++q) {
                if (ri[q] < 0) continue;
                double acc = 0.0;
#pragma unroll
                for (int s1 = 0; s1 < 2; ++s1) {
                    const int a = sp->ia[ps][q][s1];
                    if (a < 0) continue;
                    const double va = blk[sp->va[ps][q][s1]];
#pragma unroll
                    for (int s2 = 0; s2 < 2; ++s2) {
                        const int b2 = sp->ib[ps][q][s2];
                        if (b2 >= 0) acc += va * Dn[a * bs + b2] * blk[sp->vb[ps][q][s2]];
                    }
                }
                val[q] -= acc;
            }
        } else
#pragma unroll
        for (int q = 0; q < EPL; ++q) {
            if (ri[q] < 0) continue;
            const int r = ri[q], c = ci[q];
            double acc = 0.0;
            for (int s1 = 0; s1 < cnt; ++s1) {
                int a;
                double va = 0.0;
                if (above) {
                    a = __ldg(S.off_col + (tp * bs + r) * S.ow + s1);
                    if (a >= 0) va = blk[r * S.ow + s1];
                } else {
                    a = __ldg(S.offT_row + (tp * bs + r) * S.owT + s1);
                    if (a >= 0) va = blk[__ldg(S.offT_slot + (tp * bs + r) * S.owT + s1)];
                }
                if (a < 0) continue;
                for (int s2 = 0; s2 < cnt; ++s2) {
                    int b2;
                    double vb = 0.0;
                    if (above) {
                        b2 = __ldg(S.off_col + (tp * bs + c) * S.ow + s2);
                        if (b2 >= 0) vb = blk[c * S.ow + s2];
                    } else {
                        b2 = __ldg(S.offT_row + (tp * bs + c) * S.owT + s2);
                        if (b2 >= 0) vb = blk[__ldg(S.offT_slot + (tp * bs + c) * S.owT + s2)];
                    }
                    if (b2 >= 0) acc += va * Dn[a * bs + b2] * vb;
                }
            }
            val[q] -= acc;
        }
    }
    if (BS > 0 && BS <= 16 && BS % 2 == 0 && EPL <= 2) {
        // Gauss-Jordan sweeps with 2x2 pivot blocks, as in the resident kernel (gomp_kernel.cuh, factor_block_group)
#pragma unroll 1
        for (int k = 0; k < BS; k += 2) {
            double *buf = pub + ((k >> 1) & 1) * 64; // [row k | row k+1 | column k | column k+1], 16 doubles each
#pragma unroll
            for (int q = 0; q < EPL; ++q) {
                const int i = ri[q], j = ci[q];
                if (i < 0) continue;
                if (i == k) buf[j] = val[q];
                if (i == k + 1) buf[16 + j] = val[q];
                if (j == k) buf[32 + i] = val[q];
                if (j == k + 1) buf[48 + i] = val[q];
            }
            named_bar_sync(4 + g, GS);
            const double a = buf[k], b = buf[k + 1], c = buf[16 + k], d = buf[16 + k + 1];
            const double invdet = 1.0 / (a * d - b * c);
#pragma unroll
            for (int q = 0; q < EPL; ++q) {
                const int i = ri[q], j = ci[q];
                if (i < 0) continue;
                const double r1 = buf[j], r2 = buf[16 + j];
                const double c1 = buf[32 + i], c2 = buf[48 + i];
                const bool ik = (i >> 1) == (k >> 1), jk = (j >> 1) == (k >> 1);
                const double u1 = d * r1 - b * r2, u2 = a * r2 - c * r1;
                // all four cases computed, one selected: no divergent paths inside the warp
                const double w1 = c1 * d - c2 * c, w2 = c2 * a - c1 * b;
                const double gen = fma(-(c1 * u1 + c2 * u2), invdet, val[q]);
                const double rowv = (i == k ? u1 : u2) * invdet;
                const double colv = -((j == k ? w1 : w2) * invdet);
                const double piv = (i == k ? (j == k ? d : -b) : (j == k ? -c : a)) * invdet;
                val[q] = ik ? (jk ? piv : rowv) : (jk ? colv : gen);
            }
        }
    } else
#pragma unroll 1
    for (int k = 0; k < bs; ++k) { // Gauss-Jordan without pivoting (SPD block)
        double *rowk = pub + (k & 1) * 64, *colk = rowk + 32;
#pragma unroll
        for (int q = 0; q < EPL; ++q) {
            if (ri[q] == k) rowk[ci[q]] = val[q];
            if (ci[q] == k) colk[ri[q]] = val[q];
        }
        named_bar_sync(4 + g, GS);
        const double pinv = 1.0 / rowk[k];
#pragma unroll
        for (int q = 0; q < EPL; ++q) {
            const int i = ri[q], j = ci[q];
            if (i < 0) continue;
            if (i != k && j != k) val[q] = fma(-(colk[i] * pinv), rowk[j], val[q]);
            else if (i == k && j != k) val[q] *= pinv;
            else if (j == k && i != k) val[q] = -(val[q] * pinv);
            else val[q] = pinv;
        }
    }
    named_bar_sync(4 + g, GS); // every thread is done reading dprev (Schur update) and the last pivot row / column
#pragma unroll
    for (int q = 0; q < EPL; ++q)
        if (ri[q] >= 0) {
            M[gl + q * GS] = val[q];
            dprev[gl + q * GS] = val[q];
        }
    named_bar_sync(4 + g, GS); // the next stage of this chain reads the inverse just stored
}

// ---- block substitution: the step list of one chain warp ------------------------------------------------------
//   warp 0: stages 0..h-1 forward, the middle stage h, stages h-1..0 backward
//   warp 1: stages T-1..h+1 forward,               then stages h+1..T-1 backward
struct Step {
    int t, ob; // stage; off-diagonal block streamed with it (-1: none)
    int kind;  // 0 forward/from above, 1 forward/from below, 2 middle, 3 backward/from below, 4 backward/from above
};
__device__ __forceinline__ Step step_of(int warp, int j, int T, int h) {
    Step s;
    if (warp == 0) {
        if (j < h) { s.t = j; s.ob = j - 1; s.kind = 0; }
        else if (j == h) { s.t = h; s.ob = h - 1; s.kind = 2; }
        else { s.t = 2 * h - j; s.ob = s.t; s.kind = 3; }
    } else {
        const int nb = T - 1 - h;
        if (j < nb) { s.t = T - 1 - j; s.ob = s.t < T - 1 ? s.t : -1; s.kind = 1; }
        else { s.t = h + 1 + (j - nb); s.ob = s.t - 1; s.kind = 4; }
    }
    return s;
}

#define SMARK(ph)                                                                   \
    do {                                                                            \
        if (kTimers && B.prof != nullptr && tid == 0) {                                        \
            const long long now__ = clock64();                                      \
            atomicAdd(&B.prof[ph], (unsigned long long)(now__ - t_phase));          \
            t_phase = now__;                                                        \
        }                                                                           \
    } while (0)

} // namespace

// SM ("small QP"): the QP's A values, w = rho z - y and the CSR / CSC pattern arrays are kept in shared memory as in
// the resident kernel (the sparse phases then run against shared memory instead of L2); iterates, bounds, scaling
// vectors and the factor stay in the global workspace.  With 256 threads such a CTA needs < 113 KB for the N=7, H=50
// shape, so two of them share an SM: one QP's substitution chain overlaps the other's sparse phases.
template <int BS, int NT, bool SM>
__global__ void __launch_bounds__(NT, NT <= 256 ? 2 : 1) gomp_admm_stream_kernel(const __grid_constant__ DevSym S, const __grid_constant__ DevBatch B,
                                                                  const __grid_constant__ DevSettings st,
                                                                  const __grid_constant__ DevStream X) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int kST = NT; // threads of this instantiation (shadows the default used by the 512-thread kernel)
    constexpr int kG = SM ? kStreamSmallGroup : kStreamGroup;         // fast path: stages per bulk copy ...
    constexpr int kR = SM ? kStreamSmallGroupRing : kStreamGroupRing; // ... and groups in flight per chain (= X.grp, X.grp_ring)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bs = BS > 0 ? BS : S.bs;
    const int T = S.T, np = S.n_pad, mu = S.m_used, h = S.T / 2, bb = bs * bs;
    double *const sRHS = reinterpret_cast<double *>(smem_raw + X.oRHS);
    double *const sRing = reinterpret_cast<double *>(smem_raw + X.oRing);
    double *const sPub = reinterpret_cast<double *>(smem_raw + X.oPub);
    double *const sDprev = reinterpret_cast<double *>(smem_raw + X.oDprev); // [2][bs*bs]
    double *const sOffBuf = reinterpret_cast<double *>(smem_raw + X.oOffBuf); // [2][bs*ow]
    double *const sV = reinterpret_cast<double *>(smem_raw + X.oV);
    double *const sMid = reinterpret_cast<double *>(smem_raw + X.oMid); // middle stage: [inverse | Off_{h-1} | Off_h], refreshed per factorisation
    double *const sRed = reinterpret_cast<double *>(smem_raw + X.oRed);
    uint64_t *const sFull = reinterpret_cast<uint64_t *>(smem_raw + X.oBar); // [2][kStreamRing] "slot filled" (bulk copies complete)
    uint64_t *const sEmpty = sFull + 2 * kStreamRing;                        // [2][kStreamRing] "slot consumed" (fast path)
    int *const sWork = reinterpret_cast<int *>(smem_raw + X.oBar + 8 * 4 * kStreamRing);
    // small-QP mode: the three values rho takes over the rows (loose, inequality, equality) and their reciprocals; a row's
    // kind sits in two bits of its owner's `rtype` register, so the update never fetches rho from L2 nor divides by it
    double *const sRhoTab = reinterpret_cast<double *>(smem_raw + X.oBar + 8 * 4 * kStreamRing + 16);

    // A QP may be served by a CLUSTER of CTAs (long horizons, batches that leave SMs idle; launch_stream): rank 0 -- the leader --
    // runs the whole algorithm as a lone CTA would; the other CTAs are helpers that take their share of the two sparse passes
    // of every iteration (right-hand side gather, x / z / y update), reading x~ from and writing the right-hand side to the
    // leader's shared memory through distributed shared memory.  ct / CT: this thread among the cluster's threads.
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int cs = SM ? 1 : (int)cluster.num_blocks(), crank = SM ? 0 : (int)cluster.block_rank();
    const int ct = crank * kST + tid, CT = cs * kST;
    auto csync = [&]() {
        __threadfence(); // (global-memory results of this CTA: ordered by the barrier's release / acquire pair; belt and braces)
        cluster.sync();
    };
    // per-cluster workspace (global, L2-resident)
    double *const ws = X.ws + (size_t)(SM ? (int)blockIdx.x : (int)blockIdx.x / cs) * X.stride;
    double *const gA = SM ? reinterpret_cast<double *>(smem_raw + X.oA) : ws + X.wA;
    double *const gW = SM ? reinterpret_cast<double *>(smem_raw + X.oW) : ws + X.wW;
    double *const gP = ws + X.wP, *const gRho = ws + X.wRho, *const gDinv = ws + X.wDinv,
                 *const gOff = ws + X.wOff, *const gT = ws + X.wT, *const gLo = ws + X.wLo, *const gHi = ws + X.wHi;

    // small-QP mode: pattern arrays in shared memory, staged once per kernel
    uint32_t *const sRowEnt = reinterpret_cast<uint32_t *>(smem_raw + X.oRowEnt);
    uint16_t *const sCscRow = reinterpret_cast<uint16_t *>(smem_raw + X.oCscRow), *const sRowPtr = reinterpret_cast<uint16_t *>(smem_raw + X.oRowRange);
    uint16_t *const sColPk = reinterpret_cast<uint16_t *>(smem_raw + X.oColRange); // column at position p: first entry | length << 12
    // (first entry << 8 | length) of compact row i, from the 16-bit CSR pointers
    auto row_range = [&](int i) { const uint32_t a = sRowPtr[i]; return (a << 8) | ((uint32_t)sRowPtr[i + 1] - a); };
    // the same (first entry << 8 | length) for the column at position p (the host admits this mode only when every column
    // starts below 4096 and holds at most 15 entries)
    auto col_range = [&](int p_) { const uint32_t a = sColPk[p_]; return ((a & 0xFFFu) << 8) | (a >> 12); };
    if (SM) {
        for (int k = tid; k < S.nnzA; k += kST) {
            sRowEnt[k] = S.row_csr[k];
            sCscRow[k] = S.csc_row[k];
        }
        if (tid < 4) {
            sRowEnt[S.nnzA + tid] = 0u; // the 4-wide entry loops may read past the last row
            sCscRow[S.nnzA + tid] = 0;
        }
        for (int k = tid; k < np; k += kST) {
            const uint32_t cr = S.col_range[k];
            sColPk[k] = (uint16_t)((cr >> 8) | ((cr & 0xFFu) << 12));
        }
        for (int k = tid; k <= mu; k += kST) sRowPtr[k] = S.row_ptr[k];
    }
    // ---- the two passes a cluster shares (one CTA: ct = tid, CT = kST -- the plain strided loops)
    enum { OP_RHS = 1, OP_UPDATE = 2, OP_EXIT = 3, OP_ASSEMBLE = 4 };
    volatile int *const sCtl = reinterpret_cast<volatile int *>(sRhoTab); // leader -> helpers: {op, problem, do_info} (region unused outside the small-QP mode)
    // rhs = sigma x - q + A'(rho z - y), columns ct, ct + CT, ...; dst = the leader's right-hand side vector
    auto rhs_share = [&](const double *gx_, const double *gq_, double *dst, bool hasq_) {
        for (int p = ct; p < np; p += 2 * CT) {
            const bool two = p + CT < np;
            const int p2 = two ? p + CT : p;
            double ra = st.sigma * gx_[p], rb = st.sigma * gx_[p2];
            if (hasq_) {
                ra -= gq_[p];
                rb -= gq_[p2];
            }
            col_dot_ell2(S, X, np, p, p2, gA, gW, ra, rb);
            dst[p] = ra;
            if (two) dst[p2] = rb;
        }
    };
    // x <- alpha x~ + (1 - alpha) x;  z~ = A x~, projection, dual step, w = rho z - y; rows ct, ct + CT, ...; src = x~ (leader's)
    auto update_share = [&](double *gx_, double *gz_, double *gy_, double *gdy_, double *gdx_, const double *src, bool do_info_, bool hasq_) {
        for (int p = ct; p < np; p += CT) {
            const double xo = gx_[p];
            const double xn = st.alpha * src[p] + (1.0 - st.alpha) * xo;
            if (do_info_ && hasq_) gdx_[p] = xn - xo;
            gx_[p] = xn;
        }
        for (int i = ct; i < mu; i += 2 * CT) {
            const bool two = i + CT < mu;
            const int i2 = two ? i + CT : i;
            double zta, ztb;
            row_dot_ell2(X, mu, i, i2, gA, src, zta, ztb);
            const double rva = gRho[i], zoa = gz_[i], yoa = gy_[i], loa = gLo[i], hia = gHi[i];
            const double rvb = gRho[i2], zob = gz_[i2], yob = gy_[i2], lob = gLo[i2], hib = gHi[i2];
            {
                const double zr = st.alpha * zta + (1.0 - st.alpha) * zoa;
                const double zn = fmin(fmax(zr + (1.0 / rva) * yoa, loa), hia);
                const double d = rva * (zr - zn), yn = yoa + d;
                gy_[i] = yn;
                gz_[i] = zn;
                gW[i] = fma(rva, zn, -yn); // w = rho z - y for the next gather
                if (do_info_) gdy_[i] = d;
            }
            if (two) {
                const double zr = st.alpha * ztb + (1.0 - st.alpha) * zob;
                const double zn = fmin(fmax(zr + (1.0 / rvb) * yob, lob), hib);
                const double d = rvb * (zr - zn), yn = yob + d;
                gy_[i2] = yn;
                gz_[i2] = zn;
                gW[i2] = fma(rvb, zn, -yn);
                if (do_info_) gdy_[i2] = d;
            }
        }
    };
    // S = P + sigma I + A' diag(rho) A: packed diagonal-block entries and off-diagonal entries ct, ct + CT, ... into the workspace
    auto assemble_share = [&]() {
        const int dpb = bs * (bs + 1) / 2, ndiag = T * dpb;
        for (int e = ct; e < ndiag; e += CT) {
            double acc = 0.0;
            const uint32_t q1 = __ldg(S.sd_ptr + e + 1);
            for (uint32_t q = __ldg(S.sd_ptr + e); q < q1; ++q) {
                const gomp::Contrib cn = S.sd_con[q];
                acc += gRho[cn.row] * gA[cn.k1] * gA[cn.k2];
            }
            const uint16_t pk = __ldg(S.sd_pk + e);
            if (pk != 0xFFFFu) acc += gP[pk];
            const int kind = __ldg(S.sd_kind + e);
            if (kind == 1) acc += st.sigma;
            else if (kind == 2) acc = 1.0;
            const int t = e / dpb, rc = __ldg(S.tri_rc + (e - t * dpb)), rw = rc >> 8, c = rc & 0xFF;
            gDinv[(size_t)t * bb + rw * bs + c] = acc;
            gDinv[(size_t)t * bb + c * bs + rw] = acc;
        }
        for (int e = ct; e < S.noff; e += CT) {
            double acc = 0.0;
            const uint32_t q1 = __ldg(S.so_ptr + e + 1);
            for (uint32_t q = __ldg(S.so_ptr + e); q < q1; ++q) {
                const gomp::Contrib cn = S.so_con[q];
                acc += gRho[cn.row] * gA[cn.k1] * gA[cn.k2];
            }
            const uint16_t pk = __ldg(S.so_pk + e);
            if (pk != 0xFFFFu) acc += gP[pk];
            gOff[e] = acc;
        }
    };
    if constexpr (!SM) {
        if (crank != 0) {
            // ---- helper CTA: serve the leader's passes until it says exit.  Two cluster barriers per pass: "operation published"
            // and "operation complete"; the leader's control words and x~ / right-hand side are read and written in ITS shared memory.
            const volatile int *lctl = cluster.map_shared_rank(const_cast<int *>(reinterpret_cast<volatile int *>(sRhoTab)), 0);
            double *lRHS = cluster.map_shared_rank(sRHS, 0);
            for (;;) {
                csync();
                const int op = lctl[0];
                if (op == OP_EXIT) {
                    csync(); // (the leader's shared memory stays valid until every helper has read the word)
                    break;
                }
                const int hb = lctl[1];
                const bool hinfo = lctl[2] != 0, hq = B.q != nullptr;
                const size_t hmu = (size_t)hb * mu, hnp = (size_t)hb * np;
                if (op == OP_RHS) rhs_share(B.x_state + hnp, hq ? B.qs + hnp : nullptr, lRHS, hq);
                else if (op == OP_ASSEMBLE) assemble_share();
                else update_share(B.x_state + hnp, B.z_state + hmu, B.y_state + hmu, B.dyg + hmu, B.dxg + hnp, lRHS, hinfo, hq);
                csync();
            }
            return;
        }
    }
    if (tid == 0) {
        for (int k = 0; k < 4 * kStreamRing; ++k) mbar_init(sFull + k, 1);
    }
    uint32_t prod_k0 = 0, prod_k1 = 0; // producer warp: groups of completed solves, top / bottom chain
    int pre = 0;                        // producer lane: groups already issued for the coming solve
    uint32_t ring_k = 0; // slots this warp has consumed since the kernel started (slot = ring_k % kRing, parity from ring_k / kRing)
    __syncthreads();

    const int slotD = X.slot_doubles;
    const int halves = bs <= 16 ? 2 : 1; // two lanes per block row when the block has at most 16 rows
    const int rr_ = halves == 2 ? (lane & 15) : lane, half = halves == 2 ? (lane >> 4) : 0;
    const bool rowact = rr_ < bs;
    const int r = rowact ? rr_ : bs - 1;
    const int cchunk = (bs + halves - 1) / halves, c0 = half * cchunk, c1 = min(bs, c0 + cchunk);
    // uniform off-diagonal pattern (every GOMP horizon): this lane's row / column pattern, loaded once
    const bool upat = S.uniform_off && S.ow <= kMaxW && S.owT <= kMaxW;
    int ua[kMaxW], ub[kMaxW], us[kMaxW];
#pragma unroll
    for (int q = 0; q < kMaxW; ++q) {
        ua[q] = (upat && q < S.ow) ? S.off_col[r * S.ow + q] : -1;
        ub[q] = (upat && q < S.owT) ? S.offT_row[r * S.owT + q] : -1;
        us[q] = (upat && q < S.owT) ? S.offT_slot[r * S.owT + q] : 0;
    }

    // sparse products / norms / scalings on two rows or columns, on whichever copy of the pattern this mode uses
    auto row_dot2 = [&](int i1, int i2, const double *vec, double &r1, double &r2) {
        if (SM) {
            r1 = row_dot4(sRowEnt, row_range(i1), gA, vec);
            r2 = row_dot4(sRowEnt, row_range(i2), gA, vec);
        } else {
            row_dot_ell2(X, mu, i1, i2, gA, vec, r1, r2);
        }
    };
    auto row_absmax2 = [&](int i1, int i2, double &r1, double &r2) {
        if (SM) {
            double a[2] = {0.0, 0.0};
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                const uint32_t rg = row_range(w ? i2 : i1);
                for (int q = (int)(rg >> 8), q1 = q + (int)(rg & 0xFFu); q < q1; ++q) a[w] = fmax(a[w], fabs(gA[sRowEnt[q] & 0xFFFFu]));
            }
            r1 = a[0];
            r2 = a[1];
        } else {
            row_absmax_ell2(X, mu, i1, i2, gA, r1, r2);
        }
    };
    auto row_scale2 = [&](int i1, int i2, bool two, double s1, double s2) {
        if (SM) {
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                if (w == 1 && !two) break;
                const uint32_t rg = row_range(w ? i2 : i1);
                const double sc = w ? s2 : s1;
                for (int q = (int)(rg >> 8), q1 = q + (int)(rg & 0xFFu); q < q1; ++q) gA[sRowEnt[q] & 0xFFFFu] *= sc;
            }
        } else {
            row_scale_ell2(X, mu, i1, i2, two, gA, s1, s2);
        }
    };
    // A = (A * row factor) * column factor of the entry's position, and the rows' new max norms (see row_scale2_absmax_ell2)
    auto row_scale2_absmax = [&](int i1, int i2, bool two, double s1, double s2, const double *dcol, double &r1, double &r2) {
        if (SM) {
            double a[2] = {0.0, 0.0};
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                if (w == 1 && !two) break;
                const uint32_t rg = row_range(w ? i2 : i1);
                const double sc = w ? s2 : s1;
                for (int q = (int)(rg >> 8), q1 = q + (int)(rg & 0xFFu); q < q1; ++q) {
                    const uint32_t e = sRowEnt[q];
                    const double t = (gA[e & 0xFFFFu] * sc) * dcol[e >> 16];
                    gA[e & 0xFFFFu] = t;
                    a[w] = fmax(a[w], fabs(t));
                }
            }
            r1 = a[0];
            r2 = a[1];
        } else {
            row_scale2_absmax_ell2(X, mu, i1, i2, two, gA, s1, s2, dcol, r1, r2);
        }
    };
    auto col_dot2 = [&](int p1, int p2, const double *vec, double &r1, double &r2) {
        if (SM) {
            r1 = col_dot4(col_range(p1), gA, sCscRow, vec, r1);
            r2 = col_dot4(col_range(p2), gA, sCscRow, vec, r2);
        } else {
            col_dot_ell2(S, X, np, p1, p2, gA, vec, r1, r2);
        }
    };
    auto col_absmax_2 = [&](int p1, int p2, double &r1, double &r2) {
        if (SM) {
            double a[2] = {0.0, 0.0};
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                const uint32_t rg = col_range(w ? p2 : p1);
                for (int k = (int)(rg >> 8), k1 = k + (int)(rg & 0xFFu); k < k1; ++k) a[w] = fmax(a[w], fabs(gA[k]));
            }
            r1 = a[0];
            r2 = a[1];
        } else {
            col_absmax2(S, X, p1, p2, gA, r1, r2);
        }
    };
    auto col_scale_2 = [&](int p1, int p2, bool two, double s1, double s2) {
        if (SM) {
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                if (w == 1 && !two) break;
                const uint32_t rg = col_range(w ? p2 : p1);
                const double sc = w ? s2 : s1;
                for (int k = (int)(rg >> 8), k1 = k + (int)(rg & 0xFFu); k < k1; ++k) gA[k] *= sc;
            }
        } else {
            col_scale2(S, X, p1, p2, two, gA, s1, s2);
        }
    };
    auto row_dot1 = [&](int i, const double *vec) { return SM ? row_dot4(sRowEnt, row_range(i), gA, vec) : row_dot_g(S, i, gA, vec); };
    auto col_dot1 = [&](int p_, const double *vec) { return SM ? col_dot4(col_range(p_), gA, sCscRow, vec, 0.0) : col_dot_g(S, p_, gA, vec, 0.0); };

    for (;;) {
        if (tid == 0) *sWork = atomicAdd(B.work_counter, 1);
        __syncthreads();
        const int widx = *sWork;
        __syncthreads();
        if (widx >= B.batch) break;
        const int b = B.order != nullptr ? B.order[widx] : widx;
        if (B.active != nullptr && B.active[b] == 0) continue; // masked out: state and outputs stay as they are
        const size_t bmu = (size_t)b * mu, bnp = (size_t)b * np;
        const bool hasq = B.q != nullptr;
        double *const gx = B.x_state + bnp, *const gz = B.z_state + bmu, *const gy = B.y_state + bmu;
        double *const gE = B.Eg + bmu, *const gD = B.Dg + bnp, *const gdy = B.dyg + bmu, *const gdx = B.dxg + bnp;
        double *const gq = hasq ? B.qs + bnp : nullptr;
        long long t_phase = clock64();

        // ------------------------------------------------------------------ load
        // X.pre: the cluster kernel below has already equilibrated this QP (scaled A, P, bounds in its buffers; D, E, q in
        // their per-QP arrays) -- the load is a copy and the Ruiz phase is skipped
        const double *const preq = X.pre != nullptr ? X.pre + (size_t)b * X.pre_stride : nullptr;
        const double *gl = B.l + (size_t)b * S.m, *gu = B.u + (size_t)b * S.m;
        if (preq != nullptr) {
            for (int k = tid; k < S.nnzA; k += kST) gA[k] = __ldcg(preq + X.pA + k);
            for (int k = tid; k < S.nnzPu; k += kST) gP[k] = __ldcg(preq + X.pP + k);
            for (int i = tid; i < mu; i += kST) {
                gLo[i] = __ldcg(preq + X.pLo + i);
                gHi[i] = __ldcg(preq + X.pHi + i);
            }
        } else {
            const double *srcA = B.A_val + (size_t)b * B.A_pitch;
            for (int k = tid; k < S.nnzA; k += kST) gA[k] = srcA[k];
            for (int k = tid; k < S.nnzPu; k += kST) gP[k] = S.Pu_val[k];
            for (int i = tid; i < mu; i += kST) {
                const int orig = S.used_rows[i];
                gLo[i] = fmax(gl[orig], -kInfty);
                gHi[i] = fmin(gu[orig], kInfty);
                gE[i] = 1.0;
            }
            for (int p = tid; p < np; p += kST) {
                const int var = S.pos_var[p];
                gD[p] = 1.0;
                if (hasq) gq[p] = var >= 0 ? B.q[(size_t)b * S.n + var] : 0.0;
            }
        }
        double bad[1] = {0.0};
        for (int e = tid; e < S.n_empty; e += kST) {
            const int rw = S.empty_rows[e];
            if (!(gl[rw] <= 0.0 && 0.0 <= gu[rw])) bad[0] = 1.0;
        }
        block_max<1, NT>(bad, sRed, tid); // (also: the loads above are visible to every thread)
        const bool empty_infeasible = bad[0] != 0.0;
        SMARK(PH_LOAD);

        // ------------------------------------------------------------------ Ruiz equilibration
        // The P side of the scaling (values, their (i, j) positions, the full-symmetric column view) and the column
        // factors of a pass are kept in shared memory while it runs when they fit the regions that are idle before the
        // first iteration -- the right-hand side (P values) and the factor ring (patterns, column factors); otherwise
        // they are read from the workspace / pattern arrays in L2.
        double cscale = 1.0;
        if (preq != nullptr) {
            cscale = __ldcg(preq + X.pC);
        } else {
        const bool pscr = X.p_scratch != 0;
        double *const Pv = pscr ? sRHS : gP;
        double *const dtv = pscr ? sRing : gT;
        const uint32_t *const puij = pscr ? reinterpret_cast<const uint32_t *>(sRing + np) : S.Pu_ij;
        const uint32_t *const pfent = pscr ? puij + S.nnzPu : S.pf_csr;
        const uint16_t *const pfptr = pscr ? reinterpret_cast<const uint16_t *>(pfent + S.npf) : S.pf_ptr;
        // ... and the row factors E, when the ring has room behind the patterns (X.e_scratch = offset in doubles, 0 = none)
        double *const Ev = X.e_scratch > 0 ? sRing + X.e_scratch : gE;
        if (pscr) {
            uint32_t *w_ij = reinterpret_cast<uint32_t *>(sRing + np), *w_ent = w_ij + S.nnzPu;
            uint16_t *w_ptr = reinterpret_cast<uint16_t *>(w_ent + S.npf);
            for (int k = tid; k < S.nnzPu; k += kST) {
                Pv[k] = gP[k];
                w_ij[k] = S.Pu_ij[k];
            }
            for (int k = tid; k < S.npf; k += kST) w_ent[k] = S.pf_csr[k];
            for (int k = tid; k <= np; k += kST) w_ptr[k] = S.pf_ptr[k];
            if (X.e_scratch > 0)
                for (int i = tid; i < mu; i += kST) Ev[i] = 1.0;
            __syncthreads();
        }
        auto pf_absmax = [&](int p_) {
            double nrm = 0.0;
            for (int q = pfptr[p_], q1 = pfptr[p_ + 1]; q < q1; ++q) nrm = fmax(nrm, fabs(Pv[pfent[q] & 0xFFFFu]));
            return nrm;
        };
        // gW holds, per row, the max norm of the row as it stands (every row belongs to the same thread in every phase): a pass
        // traverses A twice -- the column norms, and one row-view traversal that scales an entry by its row factor, then by its
        // column factor, and takes the next pass's row norm from what it writes -- instead of four times
        if (st.scaling > 0) {
            ROWS2(i, i2, two) {
                double na, nb;
                row_absmax2(i, i2, na, nb);
                gW[i] = na;
                if (two) gW[i2] = nb;
            }
        }
#pragma unroll 1
        for (int it = 0; it < st.scaling; ++it) {
            COLS2(p, p2, two) {
                double na, nb;
                col_absmax_2(p, p2, na, nb);
                const double pa = pf_absmax(p), pb = pf_absmax(p2);
                dtv[p] = 1.0 / sqrt(limit_scaling(fmax(pa, na)));
                if (two) dtv[p2] = 1.0 / sqrt(limit_scaling(fmax(pb, nb)));
            }
            __syncthreads();
            ROWS2(i, i2, two) {
                const double ea = 1.0 / sqrt(limit_scaling(gW[i])), eb = 1.0 / sqrt(limit_scaling(gW[i2]));
                double na, nb;
                row_scale2_absmax(i, i2, two, ea, eb, dtv, na, nb);
                gW[i] = na;
                Ev[i] *= ea;
                if (two) {
                    gW[i2] = nb;
                    Ev[i2] *= eb;
                }
            }
#pragma unroll 4
            for (int k = tid; k < S.nnzPu; k += kST) {
                const uint32_t ij = puij[k];
                Pv[k] = (Pv[k] * dtv[ij >> 16]) * dtv[ij & 0xFFFFu];
            }
            __syncthreads();
            double psum = 0.0, qn[1] = {0.0};
            COLS2(p, p2, two) {
                const double da = dtv[p], db = dtv[p2];
                gD[p] *= da;
                if (two) gD[p2] *= db;
                if (hasq) {
                    gq[p] *= da;
                    qn[0] = fmax(qn[0], fabs(gq[p]));
                    if (two) {
                        gq[p2] *= db;
                        qn[0] = fmax(qn[0], fabs(gq[p2]));
                    }
                }
                const double pa = pf_absmax(p), pb = pf_absmax(p2);
                psum += pa;
                if (two) psum += pb;
            }
            psum = block_sum<NT>(psum, sRed, tid);
            block_max<1, NT>(qn, sRed, tid);
            double ct = fmax(psum / (double)S.n, limit_scaling(qn[0]));
            ct = 1.0 / limit_scaling(ct);
            for (int k = tid; k < S.nnzPu; k += kST) Pv[k] *= ct;
            if (hasq)
                for (int p = tid; p < np; p += kST) gq[p] *= ct;
            cscale *= ct;
            __syncthreads();
        }
        if (pscr) { // the scaled objective (and E) go back to the workspace (assembly, residual checks)
            for (int k = tid; k < S.nnzPu; k += kST) gP[k] = Pv[k];
            if (X.e_scratch > 0)
                for (int i = tid; i < mu; i += kST) gE[i] = Ev[i];
        }
        for (int i = tid; i < mu; i += kST) {
            gLo[i] *= gE[i];
            gHi[i] *= gE[i];
        }
        } // (in-kernel Ruiz)
        const double cinv = 1.0 / cscale;
        const bool uns = !st.scaled_termination;
        SMARK(PH_SCALE);

        // ------------------------------------------------------------------ iterates
        const int flags = B.flag_bits;
        double rho;
        int rho_updates;
        if (flags & 1) {
            for (int i = tid; i < mu; i += kST) { gz[i] = 0.0; gy[i] = 0.0; }
            for (int p = tid; p < np; p += kST) gx[p] = 0.0;
            rho = fmin(fmax(st.rho, kRhoMin), kRhoMax);
            rho_updates = 0;
        } else {
            rho = B.rho_state[b];
            rho_updates = B.rho_updates[b];
        }
        if (flags & 2) { // osqp_warm_start_x: x <- Dinv x0, z <- A x
            for (int p = tid; p < np; p += kST) {
                const int var = S.pos_var[p];
                gx[p] = var >= 0 ? (1.0 / gD[p]) * B.x0[(size_t)b * S.n + var] : 0.0;
            }
            __syncthreads();
            ROWS2(i, i2, two) {
                double za, zb;
                row_dot2(i, i2, gx, za, zb);
                gz[i] = za;
                if (two) gz[i2] = zb;
            }
        }
        __syncthreads();

        int status = empty_infeasible ? -3 : -10;
        int iters_done = 0;
        double pri_res = 0.0, dua_res = 0.0;
        bool need_factor = true, w_dirty = true;
        uint32_t rtype = 0; // small-QP mode: kind of row tid + k*kST in bits 2k, 2k+1

#pragma unroll 1
        for (int iter = 1; status == -10 && iter <= st.max_iter; ++iter) {
            if (need_factor) {
                // ---------------- rho per row, assembly of S, two-way block factorisation
                const double rho_eq = kRhoEqOverIneq * rho;
                uint32_t rt = 0;
                for (int i = tid, kq = 0; i < mu; i += kST, kq += 2) {
                    const double lo = gLo[i], hi = gHi[i];
                    double rv = rho;
                    uint32_t kind = 1u;
                    if (lo < -kInfty * kMinScaling && hi > kInfty * kMinScaling) { rv = kRhoMin; kind = 0u; }
                    else if (hi - lo < kRhoTol) { rv = rho_eq; kind = 2u; }
                    gRho[i] = rv;
                    if (SM) rt |= kind << kq;
                }
                if (SM) {
                    rtype = rt;
                    if (tid == 0) {
                        sRhoTab[0] = kRhoMin;
                        sRhoTab[1] = rho;
                        sRhoTab[2] = rho_eq;
                        sRhoTab[4] = 1.0 / kRhoMin;
                        sRhoTab[5] = 1.0 / rho;
                        sRhoTab[6] = 1.0 / rho_eq;
                    }
                }
                __syncthreads();
                if (cs > 1) {
                    if (tid == 0) sCtl[0] = OP_ASSEMBLE;
                    csync(); // (rho per row is in the workspace: the __syncthreads above, then the cluster barrier)
                }
                assemble_share();
                if (cs > 1) csync();
                __syncthreads();
                SMARK(PH_ASSEMBLE);
                if (NT < 2 * kGroup) {
                    // 256-thread CTA: two half-size groups (two matrix entries per thread), one per elimination chain
                    constexpr int kHalf = 128; // four warps per chain (named barriers count whole warps)
                    if (tid < 2 * kHalf) {
                        const int g = tid / kHalf, gl_ = tid - g * kHalf;
                        double *pub = sPub + g * 128, *dprev = sDprev + (size_t)g * bb, *obuf = sOffBuf + (size_t)g * bs * S.ow;
                        SchurPattern<(BS > 0 ? (BS * BS + kHalf - 1) / kHalf : (1024 + kHalf - 1) / kHalf)> spat;
                        load_schur_pattern(S, bs, gl_, kHalf, spat);
                        if (g == 0) {
#pragma unroll 1
                            for (int t = 0; t < h; ++t) factor_block_stream<BS, kHalf>(S, gDinv, gOff, pub, dprev, obuf, true, t, t > 0 ? 1 : 0, 0, gl_, &spat);
                        } else {
#pragma unroll 1
                            for (int t = T - 1; t > h; --t) factor_block_stream<BS, kHalf>(S, gDinv, gOff, pub, dprev, obuf, true, t, t < T - 1 ? 2 : 0, 1, gl_, &spat);
                        }
                    }
                } else if (tid < 2 * kGroup) {
                    const int g = tid / kGroup, gl_ = tid - g * kGroup;
                    double *pub = sPub + g * 128, *dprev = sDprev + (size_t)g * bb, *obuf = sOffBuf + (size_t)g * bs * S.ow;
                    SchurPattern<(BS > 0 ? (BS * BS + kGroup - 1) / kGroup : (1024 + kGroup - 1) / kGroup)> spat;
                    load_schur_pattern(S, bs, gl_, kGroup, spat);
                    if (g == 0) {
#pragma unroll 1
                        for (int t = 0; t < h; ++t) factor_block_stream<BS, kGroup>(S, gDinv, gOff, pub, dprev, obuf, true, t, t > 0 ? 1 : 0, 0, gl_, &spat);
                    } else {
#pragma unroll 1
                        for (int t = T - 1; t > h; --t) factor_block_stream<BS, kGroup>(S, gDinv, gOff, pub, dprev, obuf, true, t, t < T - 1 ? 2 : 0, 1, gl_, &spat);
                    }
                }
                __syncthreads();
                // the middle stage takes both neighbours' inverses from the global copies
                if (tid < kGroup) factor_block_stream<BS, kGroup>(S, gDinv, gOff, sPub, sDprev, sOffBuf, false, h, (h > 0 ? 1 : 0) | (h < T - 1 ? 2 : 0), 0, tid);
                fence_proxy_async_all(); // the factor is read by bulk copies (async proxy) from here on
                __syncthreads();
                {   // the middle stage's operands stay on chip (they are used once per iteration, by one warp, on the critical path)
                    const int blkD = bs * S.ow;
                    for (int k = tid; k < bb; k += kST) sMid[k] = gDinv[(size_t)h * bb + k];
                    for (int k = tid; k < blkD; k += kST) {
                        sMid[bb + k] = h > 0 ? gOff[(size_t)(h - 1) * blkD + k] : 0.0;
                        sMid[bb + blkD + k] = h < T - 1 ? gOff[(size_t)h * blkD + k] : 0.0;
                    }
                }
                __syncthreads();
                need_factor = false;
                w_dirty = true;
                SMARK(PH_FACTOR);
            }
            // ---------------- rhs = sigma x - q + A'(rho z - y)
            if (w_dirty) {
                if constexpr (SM) {
#pragma unroll 4
                    for (int i = tid, kq = 0; i < mu; i += kST, kq += 2) gW[i] = fma(sRhoTab[(rtype >> kq) & 3u], gz[i], -gy[i]);
                } else {
#pragma unroll 4
                    for (int i = tid; i < mu; i += kST) gW[i] = fma(gRho[i], gz[i], -gy[i]);
                }
                __syncthreads();
                w_dirty = false;
            }
            if constexpr (SM) {
                // x lives in L2: the loads of four columns go out before the first gather, one round trip for all of them
#pragma unroll 1
                for (int p0 = tid; p0 < np; p0 += 4 * kST) {
                    double xv[4], qv[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int p = p0 + j * kST;
                        xv[j] = p < np ? gx[p] : 0.0;
                        qv[j] = (hasq && p < np) ? gq[p] : 0.0;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int p = p0 + j * kST;
                        if (p < np) {
                            double ra = st.sigma * xv[j];
                            if (hasq) ra -= qv[j];
                            sRHS[p] = col_dot4(col_range(p), gA, sCscRow, gW, ra);
                        }
                    }
                }
            } else {
                if (cs > 1) {
                    if (tid == 0) {
                        sCtl[0] = OP_RHS;
                        sCtl[1] = b;
                    }
                    csync();
                }
                rhs_share(gx, gq, sRHS, hasq);
                if (cs > 1) csync();
            }
            __syncthreads();
            SMARK(PH_RHS);

            // ---------------- block-tridiagonal solve in place on sRHS (warps 0 and 1)
            const bool last = iter == st.max_iter;
            const bool can_check = st.check_termination && (iter % st.check_termination == 0);
            const bool adapt_now = st.adaptive_rho && st.adaptive_rho_interval && (iter % st.adaptive_rho_interval == 0);
            const bool do_info = can_check || adapt_now || last;
            const bool fastp = BS > 0 && BS <= 16 && upat && S.ow <= 2 && S.owT <= 2;
            if (warp == 2 && fastp) {
                // ---- producer warp of the fast path: keeps both chains' rings filled.  Lane c walks the groups of chain c
                // (issuing a group's bulk copies costs several hundred cycles, so the two chains get a lane each); a slot
                // is refilled as soon as its consumer has released it ("empty" barrier).
                if (lane < 2) {
                    const int blkD = bs * S.ow, grpD = kG * (bb + blkD);
                    const int L0 = h, L1 = T - 1 - h;
                    const int ng0 = 2 * ((L0 + kG - 1) / kG), ng1 = 2 * ((L1 + kG - 1) / kG);
                    fence_proxy_async();
                    // lane c serves chain c.  Groups [pre, ngc) of this solve (the first `pre` were issued at the end of the
                    // previous one), then -- when another iteration is certain to follow with the same factor, i.e. no
                    // residual check / rho update / iteration limit now -- the first groups of the next solve, so that its
                    // sweeps do not start with an L2 round trip.
                    {
                        const int c = lane;
                        const int ngc = c ? ng1 : ng0, L = c ? L1 : L0, ngh = ngc / 2;
                        const uint32_t base = c ? prod_k1 : prod_k0;
                        const int ahead = do_info ? 0 : min(kR, ngc);
                        for (int gq = pre; gq < ngc + ahead; ++gq) {
                            const int gi = gq < ngc ? gq : gq - ngc; // group within its solve
                            const bool first = gi < ngh;
                            const int k = first ? gi : gi - ngh;
                            const int cnt = min(kG, L - k * kG);
                            int t0, dir, tbo;
                            if (c == 0) {
                                if (first) { t0 = k * kG; dir = 1; tbo = -1; }
                                else { t0 = h - 1 - k * kG; dir = -1; tbo = 0; }
                            } else {
                                if (first) { t0 = T - 1 - k * kG; dir = -1; tbo = 0; }
                                else { t0 = h + 1 + k * kG; dir = 1; tbo = -1; }
                            }
                            const uint32_t kk = base + (uint32_t)gq, slot = kk % kR, round = kk / kR;
                            uint64_t *fullb = sFull + c * kStreamRing + slot, *emptyb = sEmpty + c * kStreamRing + slot;
                            {
                                unsigned spins = 0; // first use of a slot: the wait on the previous phase passes at once
                                while (!mbar_try_wait(emptyb, (round & 1u) ^ 1u)) {
                                    if (++spins > (1u << 26)) __trap();
                                }
                            }
                            double *dst = sRing + (size_t)c * X.ring_doubles + (size_t)slot * grpD;
                            const int tlo = dir > 0 ? t0 : t0 - cnt + 1;
                            const int b_lo = max(tlo + tbo, 0), b_hi = min(tlo + cnt - 1 + tbo, T - 2);
                            const uint32_t bytesD = (uint32_t)(cnt * bb) * 8u, bytesO = b_hi >= b_lo ? (uint32_t)((b_hi - b_lo + 1) * blkD) * 8u : 0u;
                            mbar_expect_tx(fullb, bytesD + bytesO);
                            bulk_g2s(dst, gDinv + (size_t)tlo * bb, bytesD, fullb);
                            if (bytesO) bulk_g2s(dst + kG * bb + (b_lo - (tlo + tbo)) * blkD, gOff + (size_t)b_lo * blkD, bytesO, fullb);
                        }
                        pre = ahead;
                    }
                    prod_k0 += (uint32_t)ng0;
                    prod_k1 += (uint32_t)ng1;
                }
            }
            if (warp < 2) {
                const bool two_chains = T - 1 - h > 0; // T <= 2: the bottom chain is empty, warp 1 has nothing to wait for
                double *ring = sRing + (size_t)warp * X.ring_doubles;
                uint64_t *full = sFull + warp * kStreamRing;
                double *vb = sV + warp * 64;
                // one substitution step in plain C++ (generic pattern / block size; also the middle stage of the fast path)
                auto substep = [&](const Step sp, const double *Dm, const double *Oa, const double *Ot, double *vbj) {
                    const bool from_above = (sp.kind == 0 || sp.kind == 2 || sp.kind == 4) && sp.t > 0;
                    const bool from_below = (sp.kind == 1 || sp.kind == 2 || sp.kind == 3) && sp.t < T - 1;
                    const bool fwd = sp.kind <= 2;
                    double a = fwd ? sRHS[sp.t * bs + r] : 0.0;
                    if (from_above) { // rows of stage t against the stage above: Off_{t-1}[r][s] * v_{t-1}[col]
                        const double *src = sRHS + (sp.t - 1) * bs;
                        for (int s = 0; s < S.ow; ++s) {
                            const int cidx = upat ? S.off_col[r * S.ow + s] : __ldg(S.off_col + ((size_t)(sp.t - 1) * bs + r) * S.ow + s);
                            if (cidx >= 0) a = fma(-Oa[r * S.ow + s], src[cidx], a);
                        }
                    }
                    if (from_below) { // columns of Off_t against the stage below: Off_t[row][slot]' * v_{t+1}[row]
                        const double *src = sRHS + (sp.t + 1) * bs;
                        for (int s = 0; s < S.owT; ++s) {
                            const size_t at = (upat ? (size_t)r : (size_t)sp.t * bs + r) * S.owT + s;
                            const int ridx = __ldg(S.offT_row + at);
                            if (ridx >= 0) a = fma(-Ot[__ldg(S.offT_slot + at)], src[ridx], a);
                        }
                    }
                    if (half == 0 && rowact) vbj[r] = a;
                    __syncwarp();
                    double acc = 0.0;
                    for (int c = c0; c < c1; ++c) acc = fma(Dm[r * bs + c], vbj[c], acc);
                    if (halves == 2) acc += __shfl_xor_sync(0xffffffffu, acc, 16);
                    if (half == 0 && rowact) sRHS[sp.t * bs + r] = fwd ? acc : sRHS[sp.t * bs + r] + acc;
                    __syncwarp();
                };
                auto wait_slot = [&](uint32_t slot, uint32_t parity) {
                    unsigned spins = 0;
                    while (!mbar_try_wait(full + slot, parity)) {
                        if (++spins > (1u << 26)) __trap(); // a lost bulk copy must not hang the device
                    }
                };
                // the middle stage needs the forward sweeps of both chains; its result starts both ways back
                auto middle = [&]() {
                    if (warp == 0) {
                        if (two_chains) named_bar_sync(1, 64);
                        Step sp;
                        sp.t = h; sp.ob = h - 1; sp.kind = 2;
                        if (fastp) {
                            // uniform pattern: the lane's neighbour indices are in registers (ua / ub / us), nothing of the
                            // middle stage is fetched from the pattern arrays in L2
                            const double *Oa = sMid + bb, *Ot = sMid + bb + bs * S.ow;
                            double a = sRHS[h * bs + r];
                            if (h > 0) {
                                const double *src = sRHS + (h - 1) * bs;
#pragma unroll
                                for (int s = 0; s < kMaxW; ++s)
                                    if (ua[s] >= 0) a = fma(-Oa[r * S.ow + s], src[ua[s]], a);
                            }
                            if (h < T - 1) {
                                const double *src = sRHS + (h + 1) * bs;
#pragma unroll
                                for (int s = 0; s < kMaxW; ++s)
                                    if (ub[s] >= 0) a = fma(-Ot[us[s]], src[ub[s]], a);
                            }
                            if (half == 0 && rowact) vb[r] = a;
                            __syncwarp();
                            double acc = 0.0;
                            for (int c = c0; c < c1; ++c) acc = fma(sMid[r * bs + c], vb[c], acc);
                            if (halves == 2) acc += __shfl_xor_sync(0xffffffffu, acc, 16);
                            if (half == 0 && rowact) sRHS[h * bs + r] = acc;
                            __syncwarp();
                        } else {
                            substep(sp, sMid, sMid + bb, sMid + bb + bs * S.ow, vb);
                        }
                        if (two_chains) named_bar_sync(2, 64);
                    } else if (two_chains) {
                        named_bar_sync(1, 64);
                        named_bar_sync(2, 64);
                    }
                };
                const bool fast = fastp;
                if (fast) {
                    // ---- fast path: the factor streams in groups of kG stages (one bulk copy for the
                    // pivot-block inverses, one for the off-diagonal blocks); inside a group the hand-scheduled
                    // sweep of the resident kernel runs on the slot as if it were the resident factor
                    const int L = warp == 0 ? h : T - 1 - h; // stages of this chain
                    const int ng = (L + kG - 1) / kG, ngroups = 2 * ng;
                    const int blkD = bs * S.ow, grpD = kG * (bb + blkD);
                    const uint32_t rhs0 = smem_u32(sRHS), vb0 = smem_u32(vb), ring0 = smem_u32(ring);
                    const uint32_t blk_bytes = (uint32_t)blkD * 8u;
                    // per-lane pattern constants: forward = row r of the block above, transposed = column r of the block below
                    const int fg0 = ua[0] >= 0 ? ua[0] : 0, fg1 = ua[1] >= 0 ? ua[1] : 0;
                    const uint32_t fo0 = (uint32_t)(r * S.ow) * 8u, fo1 = fo0 + 8u;
                    const bool fh0 = ua[0] >= 0, fh1 = ua[1] >= 0;
                    const int tg0 = ub[0] >= 0 ? ub[0] : 0, tg1 = ub[1] >= 0 ? ub[1] : 0;
                    const uint32_t to0 = (uint32_t)us[0] * 8u, to1 = (uint32_t)us[1] * 8u;
                    const bool th0 = ub[0] >= 0, th1 = ub[1] >= 0;
                    uint32_t kk = ring_k;
                    // one phase of the chain (way in: fwd, way back: !fwd) = ng groups of up to kG stages.  Everything that
                    // is the same for the groups of a phase (direction, which off-diagonal block a stage uses, the lane's
                    // pattern constants) is settled before the loop; a single warp runs this, so every instruction between
                    // two sweeps is on the critical path of the iteration.
                    auto run_phase = [&](const bool fwd, double prev) {
                        const bool down = (warp == 0) == fwd; // top chain walks down on its way in, the bottom chain on its way back
                        const int dir = down ? 1 : -1, tbo = down ? -1 : 0;
                        int t0 = warp == 0 ? (fwd ? 0 : h - 1) : (fwd ? T - 1 : h + 1);
                        const int g0 = down ? fg0 : tg0, g1 = down ? fg1 : tg1;
                        const uint32_t o0 = down ? fo0 : to0, o1 = down ? fo1 : to1;
                        const bool h0 = down ? fh0 : th0, h1 = down ? fh1 : th1;
                        int left = L;
#pragma unroll 1
                        for (int k = 0; k < ng; ++k, ++kk) {
                            const int cnt = min(kG, left);
                            const uint32_t slot = kk % kR;
                            wait_slot(slot, (kk / kR) & 1u);
                            const int tlo = down ? t0 : t0 - cnt + 1;
                            const uint32_t sl = ring0 + slot * (uint32_t)grpD * 8u;
                            const uint32_t dinv0 = sl - (uint32_t)(tlo * bb) * 8u;                                  // virtual base: stage t at + t*bb*8
                            const uint32_t off0 = sl + (uint32_t)(kG * bb) * 8u - (uint32_t)(tlo + tbo) * blk_bytes; // block tb at + tb*blk_bytes
                            prev = fwd ? chain_sweep_call<BS, true>(dinv0, rhs0, off0, vb0, t0, cnt, dir, tbo, T - 1, blk_bytes, prev, lane, g0, g1, o0, o1, h0, h1)
                                       : chain_sweep_call<BS, false>(dinv0, rhs0, off0, vb0, t0, cnt, dir, tbo, T - 1, blk_bytes, prev, lane, g0, g1, o0, o1, h0, h1);
                            __syncwarp(); // every lane's reads of the slot have retired: hand it back to the producer warp
                            if (lane == 0) mbar_arrive(sEmpty + warp * kStreamRing + slot);
                            t0 += dir * cnt;
                            left -= cnt;
                        }
                        return prev;
                    };
                    run_phase(true, 0.0);
                    middle();
                    run_phase(false, ng > 0 ? sRHS[h * bs + r] : 0.0);
                    ring_k += (uint32_t)ngroups;
                } else {
                    // ---- generic path: one ring slot per stage
                    const int nsteps = warp == 0 ? 2 * h + 1 : 2 * (T - 1 - h);
                    auto issue = [&](int j) { // lane 0: start the bulk copies of step j
                        const Step sp = step_of(warp, j, T, h);
                        const uint32_t slot = (ring_k + (uint32_t)j) % (uint32_t)X.gen_ring;
                        double *dst = ring + (size_t)slot * slotD;
                        const uint32_t bytesD = (uint32_t)bb * 8u, bytesO = (uint32_t)(bs * S.ow) * 8u;
                        const bool has_off = sp.ob >= 0 && sp.ob < T - 1;
                        mbar_expect_tx(full + slot, bytesD + (has_off ? bytesO : 0u));
                        bulk_g2s(dst, gDinv + (size_t)sp.t * bb, bytesD, full + slot);
                        if (has_off) bulk_g2s(dst + bb, gOff + (size_t)sp.ob * bs * S.ow, bytesO, full + slot);
                    };
                    if (lane == 0) {
                        fence_proxy_async();
                        for (int j = 0; j < X.gen_ring - 1 && j < nsteps; ++j) issue(j);
                    }
#pragma unroll 1
                    for (int j = 0; j < nsteps; ++j) {
                        const Step sp = step_of(warp, j, T, h);
                        if (warp == 0 && j == h && two_chains) named_bar_sync(1, 64); // middle stage: needs the bottom chain's forward sweep
                        if (warp == 1 && j == T - 1 - h) {
                            named_bar_sync(1, 64);
                            named_bar_sync(2, 64); // ... and the bottom chain's way back needs x_h
                        }
                        // the slot consumed one step ago is free: refill it kStreamRing-1 steps ahead
                        // (its last generic-proxy reads retired before the previous step's closing __syncwarp)
                        if (lane == 0 && j + X.gen_ring - 1 < nsteps) issue(j + X.gen_ring - 1);
                        const uint32_t kk = ring_k + (uint32_t)j, slot = kk % (uint32_t)X.gen_ring;
                        wait_slot(slot, (kk / (uint32_t)X.gen_ring) & 1u);
                        const double *Dm = ring + (size_t)slot * slotD, *Ob = Dm + bb;
                        // kinds 0, 2, 4 stream Off_{t-1}; kinds 1, 3 stream Off_t; the middle stage reads Off_h from L2
                        substep(sp, Dm, Ob, sp.kind == 2 ? gOff + (size_t)sp.t * bs * S.ow : Ob, vb + (j & 1) * 32);
                        if (warp == 0 && j == h && two_chains) named_bar_sync(2, 64); // x_h published for the bottom chain
                    }
                    ring_k += (uint32_t)nsteps;
                }
            }
            if (kTimers && B.prof != nullptr && tid == 0) atomicAdd(&B.prof[PH_CHAIN], (unsigned long long)(clock64() - t_phase));
            __syncthreads();
            SMARK(PH_SOLVE);

            // ---------------- x, z, y updates
            if constexpr (SM) {
                // the iterates live in L2 (two CTAs per SM leave no registers or shared memory for them): a thread's loads
                // for four elements are issued before any of them is used, so one L2 round trip serves four rows
#pragma unroll 1
                for (int p0 = tid; p0 < np; p0 += 4 * kST) {
                    double xo[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int p = p0 + j * kST;
                        xo[j] = p < np ? gx[p] : 0.0;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int p = p0 + j * kST;
                        if (p < np) {
                            const double xn = st.alpha * sRHS[p] + (1.0 - st.alpha) * xo[j];
                            if (do_info && hasq) gdx[p] = xn - xo[j];
                            gx[p] = xn;
                        }
                    }
                }
#pragma unroll 1
                for (int i0 = tid, kq = 0; i0 < mu; i0 += 4 * kST, kq += 8) {
                    double zo[4], yo[4], lo[4], hi[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int i = i0 + j * kST;
                        const bool in = i < mu;
                        zo[j] = in ? gz[i] : 0.0;
                        yo[j] = in ? gy[i] : 0.0;
                        lo[j] = in ? gLo[i] : 0.0;
                        hi[j] = in ? gHi[i] : 0.0;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int i = i0 + j * kST;
                        if (i < mu) {
                            const uint32_t kind = (rtype >> (kq + 2 * j)) & 3u;
                            const double rvj = sRhoTab[kind], rinv = sRhoTab[4 + kind];
                            const double zt = row_dot4(sRowEnt, row_range(i), gA, sRHS);
                            const double zr = st.alpha * zt + (1.0 - st.alpha) * zo[j];
                            const double zn = fmin(fmax(zr + rinv * yo[j], lo[j]), hi[j]);
                            const double d = rvj * (zr - zn), yn = yo[j] + d;
                            gy[i] = yn;
                            gz[i] = zn;
                            gW[i] = fma(rvj, zn, -yn); // w = rho z - y for the next gather
                            if (do_info) gdy[i] = d;
                        }
                    }
                }
            } else {
                if (cs > 1) {
                    if (tid == 0) {
                        sCtl[0] = OP_UPDATE;
                        sCtl[1] = b;
                        sCtl[2] = do_info ? 1 : 0;
                    }
                    csync();
                }
                update_share(gx, gz, gy, gdy, gdx, sRHS, do_info, hasq);
                if (cs > 1) csync();
            }
            __syncthreads();
            SMARK(PH_UPDATE);

            if (do_info) {
                // ---------------- residuals and the norms OSQP's tolerances need (update_info)
                w_dirty = true; // the infeasibility test below borrows gW
                double pri_u, pri_s, pn_u, pn_s, dua_u, dua_s, dn_u, dn_s;
                {
                    double na[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                    ROWS2(i, i2, two) {
                        double axa, axb;
                        row_dot2(i, i2, gx, axa, axb);
#pragma unroll
                        for (int w = 0; w < 2; ++w) {
                            if (w == 1 && !two) break;
                            const int ii = w ? i2 : i;
                            const double ax = w ? axb : axa;
                            const double zi = gz[ii];
                            const double pr = ax - zi;
                            const double einv = 1.0 / gE[ii];
                            na[0] = fmax(na[0], fabs(einv * pr));
                            na[1] = fmax(na[1], fabs(pr));
                            na[2] = fmax(na[2], fabs(einv * zi));
                            na[3] = fmax(na[3], fabs(einv * ax));
                            na[4] = fmax(na[4], fabs(zi));
                            na[5] = fmax(na[5], fabs(ax));
                        }
                    }
                    block_max<6, NT>(na, sRed, tid);
                    pri_u = na[0]; pri_s = na[1]; pn_u = fmax(na[2], na[3]); pn_s = fmax(na[4], na[5]);
                }
                {
                    double nb[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                    COLS2(p0, p2, two) {
                      double atya = 0.0, atyb = 0.0;
                      col_dot2(p0, p2, gy, atya, atyb);
#pragma unroll
                      for (int w = 0; w < 2; ++w) {
                        if (w == 1 && !two) break;
                        const int p = w ? p2 : p0;
                        const double px = pf_dot_g(S, p, gP, gx);
                        const double aty = w ? atyb : atya;
                        const double qq = hasq ? gq[p] : 0.0;
                        const double du = qq + px + aty;
                        const double dinv = 1.0 / gD[p];
                        nb[0] = fmax(nb[0], fabs(dinv * du));
                        nb[1] = fmax(nb[1], fabs(du));
                        nb[2] = fmax(nb[2], fabs(dinv * px));
                        nb[3] = fmax(nb[3], fabs(dinv * aty));
                        nb[4] = fmax(nb[4], fabs(dinv * qq));
                        nb[5] = fmax(nb[5], fabs(px));
                        nb[6] = fmax(nb[6], fabs(aty));
                        nb[7] = fmax(nb[7], fabs(qq));
                      }
                    }
                    block_max<8, NT>(nb, sRed, tid);
                    dua_u = nb[0]; dua_s = nb[1]; dn_u = fmax(nb[4], fmax(nb[3], nb[2])); dn_s = fmax(nb[7], fmax(nb[6], nb[5]));
                }
                pri_res = uns ? pri_u : pri_s;
                dua_res = uns ? cinv * dua_u : dua_s;
                iters_done = iter;

                bool done = false;
#pragma unroll 1
                for (int stage = 0; stage < 2 && !done; ++stage) {
                    if (stage == 0 && !(can_check || last)) continue;
                    if (stage == 1) {
                        if (adapt_now) { // OSQP adapt_rho: estimate from the scaled residuals
                            const double pr = pri_s / (pn_s + kDivTol);
                            const double dr = dua_s / (dn_s + kDivTol);
                            const double est = fmin(fmax(rho * sqrt(pr / dr), kRhoMin), kRhoMax);
                            if (est > rho * st.adaptive_rho_tolerance || est < rho / st.adaptive_rho_tolerance) {
                                rho = est;
                                rho_updates++;
                                need_factor = true;
                            }
                        }
                        if (!last) break;
                    }
                    const bool approx = stage == 1;
                    double ea = st.eps_abs, er = st.eps_rel, epi = st.eps_prim_inf, edi = st.eps_dual_inf;
                    if (pri_res > kInfty || dua_res > kInfty) {
                        status = -7;
                        done = true;
                        break;
                    }
                    if (approx) {
                        ea *= 10; er *= 10; epi *= 10; edi *= 10;
                    }
                    bool prc = false, drc = false, pic = false, dic = false;
                    const double eps_prim = ea + er * (uns ? pn_u : pn_s);
                    if (mu == 0 || pri_res < eps_prim) {
                        prc = true;
                    } else {
                        // OSQP is_primal_infeasible on delta_y projected onto the polar of the recession cone (kept in gW)
                        double nrm[1] = {0.0}, lhs = 0.0;
                        for (int i = tid; i < mu; i += kST) {
                            double d = gdy[i];
                            const double l_ = gLo[i], u_ = gHi[i];
                            if (u_ > kInfty * kMinScaling) {
                                if (l_ < -kInfty * kMinScaling) d = 0.0;
                                else d = fmin(d, 0.0);
                            } else if (l_ < -kInfty * kMinScaling) {
                                d = fmax(d, 0.0);
                            }
                            gW[i] = d;
                            nrm[0] = fmax(nrm[0], fabs(uns ? gE[i] * d : d));
                            lhs += u_ * fmax(d, 0.0) + l_ * fmin(d, 0.0);
                        }
                        block_max<1, NT>(nrm, sRed, tid);
                        lhs = block_sum<NT>(lhs, sRed, tid);
                        if (nrm[0] > kDivTol && lhs < -epi * nrm[0]) {
                            double v[1] = {0.0};
                            for (int p = tid; p < np; p += kST) {
                                const double acc = col_dot1(p, gW);
                                v[0] = fmax(v[0], fabs(uns ? (1.0 / gD[p]) * acc : acc));
                            }
                            block_max<1, NT>(v, sRed, tid);
                            pic = v[0] < epi * nrm[0];
                        }
                    }
                    const double mx = uns ? cinv * dn_u : dn_s;
                    const double eps_dual = ea + er * mx;
                    if (dua_res < eps_dual) {
                        drc = true;
                    } else if (hasq) {
                        // OSQP is_dual_infeasible on delta_x (q'dx < 0 is impossible when q = 0)
                        double nrm[1] = {0.0}, qdx = 0.0;
                        for (int p = tid; p < np; p += kST) {
                            const double d = gdx[p];
                            nrm[0] = fmax(nrm[0], fabs(uns ? gD[p] * d : d));
                            qdx += gq[p] * d;
                        }
                        block_max<1, NT>(nrm, sRed, tid);
                        qdx = block_sum<NT>(qdx, sRed, tid);
                        const double cs_ = uns ? cscale : 1.0;
                        if (nrm[0] > kDivTol && qdx < -cs_ * edi * nrm[0]) {
                            double v[1] = {0.0};
                            for (int p = tid; p < np; p += kST) {
                                const double acc = pf_dot_g(S, p, gP, gdx);
                                v[0] = fmax(v[0], fabs(uns ? (1.0 / gD[p]) * acc : acc));
                            }
                            block_max<1, NT>(v, sRed, tid);
                            if (v[0] < cs_ * edi * nrm[0]) {
                                double viol[1] = {0.0};
                                for (int i = tid; i < mu; i += kST) {
                                    double acc = row_dot1(i, gdx);
                                    if (uns) acc *= 1.0 / gE[i];
                                    if ((gHi[i] < kInfty * kMinScaling && acc > edi * nrm[0]) ||
                                        (gLo[i] > -kInfty * kMinScaling && acc < -edi * nrm[0]))
                                        viol[0] = 1.0;
                                }
                                block_max<1, NT>(viol, sRed, tid);
                                dic = viol[0] == 0.0;
                            }
                        }
                    }
                    if (prc && drc) {
                        status = approx ? 2 : 1;
                        done = true;
                    } else if (pic) {
                        status = approx ? 3 : -3;
                        done = true;
                    } else if (dic) {
                        status = approx ? 4 : -4;
                        done = true;
                    }
                }
                if (!done && last) status = -2;
                __syncthreads();
                SMARK(PH_INFO);
            }
        }

        // ------------------------------------------------------------------ store
        const bool infeasible = status == -3 || status == 3 || status == -4 || status == 4 || status == -7;
        const double qnan = __longlong_as_double(0x7ff8000000000000LL);
        double *xo = B.x_out + (size_t)b * S.n, *yo = B.y_out + (size_t)b * S.m;
        for (int p = tid; p < np; p += kST) {
            const int var = S.pos_var[p];
            if (var >= 0) xo[var] = infeasible ? qnan : gD[p] * gx[p];
            if (infeasible) gx[p] = 0.0;
        }
        for (int i = tid; i < mu; i += kST) {
            yo[S.used_rows[i]] = infeasible ? qnan : (gE[i] * gy[i]) * cinv;
            if (infeasible) {
                gz[i] = 0.0;
                gy[i] = 0.0;
            }
        }
        for (int e = tid; e < S.n_empty; e += kST) yo[S.empty_rows[e]] = infeasible ? qnan : 0.0;
        if (tid == 0) {
            B.rho_state[b] = rho;
            B.status[b] = status;
            B.iters[b] = iters_done;
            B.rho_updates[b] = rho_updates;
            B.pri_res[b] = pri_res;
            B.dua_res[b] = dua_res;
        }
        __syncthreads();
        SMARK(PH_STORE);
    }
    if (cs > 1) { // release the helpers; they read the word in this CTA's shared memory, so wait until they have
        if (tid == 0) sCtl[0] = OP_EXIT;
        csync();
        csync();
    }
}

// ================================================================================================
//          Ruiz equilibration by a thread-block CLUSTER per QP (long horizons, small batches)
// ================================================================================================
// A long-horizon QP alone on its SM spends a quarter of its solve in the ten Ruiz passes: five strided traversals of A per
// pass against one SM's L2 port (2.9 M of 10.9 M cycles at H = 400).  When the batch leaves SMs idle (batch x cluster size <=
// SM count), this kernel runs first: a cluster of up to 8 CTAs per QP, every thread owning a few rows and columns, the
// phases of a pass separated by the hardware cluster barrier (barrier.cluster, arrive.release / wait.acquire) instead of
// __syncthreads.  It performs the very operations of the in-kernel phase on the same pattern arrays -- max norms are order
// independent and every entry is scaled row factor first, column factor second, as there -- and leaves the scaled A, P, l, u
// and the cost scale in per-QP buffers (X.pre), D, E and the scaled q in their per-QP arrays; the ADMM kernel then loads
// instead of scaling.  Only the cost normalisation's mean of the column norms of P is summed in a different order (per CTA,
// then over the cluster's CTAs in rank order): last-bit differences in c, as between any two summation trees.
__global__ void __launch_bounds__(kStreamThreads, 1) gomp_ruiz_cluster_kernel(const __grid_constant__ DevSym S, const __grid_constant__ DevBatch B,
                                                                               const __grid_constant__ DevSettings st,
                                                                               const __grid_constant__ DevStream X) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ double sRed[2 * (kStreamThreads / 32)];
    const int tid = threadIdx.x;
    const int nb = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int gt = rank * kStreamThreads + tid, GT = nb * kStreamThreads;
    const int nclusters = (int)gridDim.x / nb, cid = (int)blockIdx.x / nb;
    const int np = S.n_pad, mu = S.m_used;
    const bool hasq = B.q != nullptr;
    auto csync = [&]() {
        __threadfence(); // (the barrier's release / acquire pair orders the cluster's global-memory traffic; belt and braces)
        cluster.sync();
    };
    for (int b = cid; b < B.batch; b += nclusters) {
        if (B.active != nullptr && B.active[b] == 0) continue; // (uniform over the cluster)
        double *const pre = X.pre + (size_t)b * X.pre_stride;
        double *const A = pre + X.pA, *const P = pre + X.pP, *const Lo = pre + X.pLo, *const Hi = pre + X.pHi, *const dt = pre + X.pT,
                     *const et = pre + X.pW, *const red = pre + X.pRed;
        double *const E = B.Eg + (size_t)b * mu, *const D = B.Dg + (size_t)b * np;
        double *const q = hasq ? B.qs + (size_t)b * np : nullptr;
        // ---- load (as the ADMM kernel's load phase)
        {
            const double *srcA = B.A_val + (size_t)b * B.A_pitch;
            for (int k = gt; k < S.nnzA; k += GT) __stcg(A + k, srcA[k]);
            for (int k = gt; k < S.nnzPu; k += GT) __stcg(P + k, S.Pu_val[k]);
            const double *gl = B.l + (size_t)b * S.m, *gu = B.u + (size_t)b * S.m;
            for (int i = gt; i < mu; i += GT) {
                const int orig = S.used_rows[i];
                __stcg(Lo + i, fmax(gl[orig], -kInfty));
                __stcg(Hi + i, fmin(gu[orig], kInfty));
                __stcg(E + i, 1.0);
            }
            for (int p = gt; p < np; p += GT) {
                const int var = S.pos_var[p];
                __stcg(D + p, 1.0);
                if (hasq) __stcg(q + p, var >= 0 ? B.q[(size_t)b * S.n + var] : 0.0);
            }
        }
        csync();
        double cscale = 1.0;
        // Row norms of A for the first pass (et holds, per row, the max norm of the row as it stands; every row belongs to the
        // same thread in every phase).  From then on a pass traverses A twice instead of four times: the column norms, and ONE
        // row-view traversal that scales an entry by its row factor and then by its column factor and takes the next pass's row
        // norm from what it writes.
        if (st.scaling > 0) {
            for (int i = gt; i < mu; i += 2 * GT) {
                const bool two = i + GT < mu;
                const int i2 = two ? i + GT : i;
                double na, nb_;
                row_absmax_ell2(X, mu, i, i2, A, na, nb_);
                et[i] = na;
                if (two) et[i2] = nb_;
            }
        }
#pragma unroll 1
        for (int it = 0; it < st.scaling; ++it) {
            // ---- norms of the KKT columns (P and A columns); two at a time on the ELL views, as the in-kernel phase does
            // (these loops are L2-latency bound: independent loads first)
            for (int p = gt; p < np; p += 2 * GT) {
                const bool two = p + GT < np;
                const int p2 = two ? p + GT : p;
                double na, nb_;
                col_absmax2(S, X, p, p2, A, na, nb_);
                const double pa = pf_absmax_g(S, p, P), pb = pf_absmax_g(S, p2, P);
                dt[p] = 1.0 / sqrt(limit_scaling(fmax(pa, na)));
                if (two) dt[p2] = 1.0 / sqrt(limit_scaling(fmax(pb, nb_)));
            }
            csync();
            // ---- A by the row factors, then by the column factors; E; P by the column factors on both sides
            for (int i = gt; i < mu; i += 2 * GT) {
                const bool two = i + GT < mu;
                const int i2 = two ? i + GT : i;
                const double ea = 1.0 / sqrt(limit_scaling(et[i])), eb = 1.0 / sqrt(limit_scaling(et[i2]));
                double na, nb_;
                row_scale2_absmax_ell2(X, mu, i, i2, two, A, ea, eb, dt, na, nb_);
                et[i] = na;
                E[i] *= ea;
                if (two) {
                    et[i2] = nb_;
                    E[i2] *= eb;
                }
            }
#pragma unroll 4
            for (int k = gt; k < S.nnzPu; k += GT) {
                const uint32_t ij = __ldg(S.Pu_ij + k);
                P[k] = (P[k] * dt[ij >> 16]) * dt[ij & 0xFFFFu];
            }
            csync();
            // ---- D and q by the column factors; cost normalisation from the scaled P
            double psum = 0.0, qn = 0.0;
            for (int p = gt; p < np; p += 2 * GT) {
                const bool two = p + GT < np;
                const int p2 = two ? p + GT : p;
                const double da = dt[p], db = dt[p2];
                D[p] *= da;
                if (two) D[p2] *= db;
                if (hasq) {
                    q[p] *= da;
                    qn = fmax(qn, fabs(q[p]));
                    if (two) {
                        q[p2] *= db;
                        qn = fmax(qn, fabs(q[p2]));
                    }
                }
                psum += pf_absmax_g(S, p, P);
                if (two) psum += pf_absmax_g(S, p2, P);
            }
            psum = block_sum<kStreamThreads>(psum, sRed, tid);
            {
                double qv[1] = {qn};
                block_max<1, kStreamThreads>(qv, sRed, tid);
                qn = qv[0];
            }
            if (tid == 0) {
                red[rank] = psum;
                red[16 + rank] = qn;
            }
            csync();
            double ps = 0.0, qm = 0.0;
            for (int r = 0; r < nb; ++r) {
                ps += red[r];
                qm = fmax(qm, red[16 + r]);
            }
            double ct = fmax(ps / (double)S.n, limit_scaling(qm));
            ct = 1.0 / limit_scaling(ct);
#pragma unroll 4
            for (int k = gt; k < S.nnzPu; k += GT) P[k] *= ct;
            if (hasq)
                for (int p = gt; p < np; p += GT) q[p] *= ct;
            cscale *= ct;
            csync();
        }
        for (int i = gt; i < mu; i += GT) {
            const double e = E[i];
            Lo[i] *= e;
            Hi[i] *= e;
        }
        if (gt == 0) __stcg(pre + X.pC, cscale);
        csync(); // (the next QP of this cluster reuses nothing, but keeps the CTAs of the cluster in step)
    }
}

cudaError_t launch_prescale(const DevSym &S, const DevBatch &B, const DevSettings &st, const DevStream &X, int clusters, int cluster_size,
                            cudaStream_t stream) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * cluster_size), 1, 1);
    cfg.blockDim = dim3(kStreamThreads, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster_size;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, gomp_ruiz_cluster_kernel, S, B, st, X);
}

cudaError_t prepare_stream(int bs, bool small, size_t smem_bytes, int *ctas_per_sm) {
    cudaError_t e;
#define PREP(FN, NTHR)                                                                                    \
    e = cudaFuncSetAttribute(FN, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);           \
    if (e != cudaSuccess) return e;                                                                       \
    cudaFuncSetAttribute(FN, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); \
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, FN, NTHR, smem_bytes);
    if (small) {
        if (bs == 14) { PREP((gomp_admm_stream_kernel<14, kStreamSmallThreads, true>), kStreamSmallThreads) }
        PREP((gomp_admm_stream_kernel<0, kStreamSmallThreads, true>), kStreamSmallThreads)
    }
    if (bs == 14) { PREP((gomp_admm_stream_kernel<14, kStreamThreads, false>), kStreamThreads) }
    if (bs == 12) { PREP((gomp_admm_stream_kernel<12, kStreamThreads, false>), kStreamThreads) } // N = 6: the reference example's arm
    PREP((gomp_admm_stream_kernel<0, kStreamThreads, false>), kStreamThreads)
#undef PREP
}
// how many clusters of `cluster_size` CTAs of the 512-thread ADMM kernel the device holds at once (GPC boundaries: fewer than
// SMs / cluster_size)
cudaError_t stream_max_clusters(int bs, int cluster_size, size_t smem_bytes, int *n) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)cluster_size, 1, 1);
    cfg.blockDim = dim3(kStreamThreads, 1, 1);
    cfg.dynamicSmemBytes = smem_bytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster_size;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (bs == 14) return cudaOccupancyMaxActiveClusters(n, gomp_admm_stream_kernel<14, kStreamThreads, false>, &cfg);
    if (bs == 12) return cudaOccupancyMaxActiveClusters(n, gomp_admm_stream_kernel<12, kStreamThreads, false>, &cfg);
    return cudaOccupancyMaxActiveClusters(n, gomp_admm_stream_kernel<0, kStreamThreads, false>, &cfg);
}

cudaError_t launch_stream(const DevSym &S, const DevBatch &B, const DevSettings &st, const DevStream &X, int grid, int cluster_size,
                          size_t smem_bytes, cudaStream_t stream) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3((unsigned)(X.small ? kStreamSmallThreads : kStreamThreads), 1, 1);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)(X.small ? 1 : std::max(1, cluster_size));
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (X.small) {
        if (S.bs == 14) return cudaLaunchKernelEx(&cfg, gomp_admm_stream_kernel<14, kStreamSmallThreads, true>, S, B, st, X);
        return cudaLaunchKernelEx(&cfg, gomp_admm_stream_kernel<0, kStreamSmallThreads, true>, S, B, st, X);
    }
    if (S.bs == 14) return cudaLaunchKernelEx(&cfg, gomp_admm_stream_kernel<14, kStreamThreads, false>, S, B, st, X);
    if (S.bs == 12) return cudaLaunchKernelEx(&cfg, gomp_admm_stream_kernel<12, kStreamThreads, false>, S, B, st, X);
    return cudaLaunchKernelEx(&cfg, gomp_admm_stream_kernel<0, kStreamThreads, false>, S, B, st, X);
}

} // namespace gomp_dev
