// gomp_kernel.cuh -- the persistent ADMM kernel of libgomp_cuda (sm_100a, fp64) and its launchers.
//
// One CTA pulls QPs from a device work queue and runs the whole OSQP algorithm for each of
// them on chip (see gomp_cuda.cu for the overview).  Every kernel variant (rows per thread,
// columns per thread, compile-time block size) is instantiated in its own translation unit
// (gomp_kernel_inst.cu with -DGOMP_VARIANT=...), so variants compile in parallel.
//
// Algorithm per QP, following OSQP 0.6 as restated in oracle/osqp_ref.c:
//   load -> Ruiz equilibration (10 passes) -> rho classification -> [warm start] ->
//   { factor S = P + sigma I + A' diag(rho) A ; ADMM iterations ; every check_termination
//     iterations residuals/termination/infeasibility ; adaptive rho -> refactor } -> store.
// Reference call sites replaced: src/osqp-wrapper.h:28 (Init), :46 (SetPrimalWarmStart),
// :52 (Solve), :36,:40 (UpdateConstraintMatrix, SetBounds).
#ifndef GOMP_KERNEL_CUH
#define GOMP_KERNEL_CUH

#include <cuda_runtime.h>

#include <cstdint>

#include "symbolic.h"

namespace gomp_dev {

// ---- OSQP constants ---------------------------------------------------------
constexpr double kRhoMin = 1e-6, kRhoMax = 1e6, kRhoEqOverIneq = 1e3, kRhoTol = 1e-4;
constexpr double kMinScaling = 1e-4, kMaxScaling = 1e4, kInfty = 1e30, kDivTol = 1.0 / 1e30;

constexpr int kThreads = 256;
constexpr int kRedK = 14; // widest multi-value block reduction

// ---- device-side views ------------------------------------------------------------------------
struct DevSym {
    int n, m, nnzA, nnzPu, m_used, bs, T, n_pad, roww, colw, ow, owT, dpb, pfw, n_empty, noff;
    int uniform_off; // every off-diagonal block has the same pattern -> indices cached in registers
    const int32_t *pos_var, *used_rows, *empty_rows;
    const uint32_t *row_ent;
    const uint8_t *row_len;
    const uint32_t *col_start;
    const uint8_t *col_len;
    const uint16_t *col_row;
    const double *Pu_val;
    const uint32_t *Pu_ij, *pf_ent;
    const uint8_t *pf_len;
    const int8_t *off_col, *offT_row;
    const uint16_t *offT_slot;
    const uint32_t *sd_ptr;
    const gomp::Contrib *sd_con;
    const uint16_t *sd_pk;
    const uint8_t *sd_kind; // 0 plain entry, 1 diagonal, 2 diagonal of a padding variable
    const uint32_t *so_ptr;
    const gomp::Contrib *so_con;
    const uint16_t *so_pk;
    const uint16_t *tri_rc; // packed lower index -> (row << 8 | col)
    // shared-memory carve-up (in doubles)
    int oA, oP, oW, oRHS, oDinv, oOff, oInv, oRed, oEnd;
};

struct DevBatch {
    int batch;
    int bulk_ok; // A rows are 16-byte aligned and a multiple of 16 bytes: stage with cp.async.bulk
    const double *A_val, *l, *u, *q, *x0;
    long long A_pitch;
    double *x_state, *z_state, *y_state, *rho_state;
    int *flags; // bit0: cold start, bit1: primal warm start pending
    double *Eg, *Dg, *qs, *dyg, *dxg;
    double *x_out, *y_out;
    int *status, *iters, *rho_updates;
    double *pri_res, *dua_res, *obj_val;
    int *work_counter;
};

struct DevSettings {
    double rho, sigma, alpha, adaptive_rho_tolerance, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf;
    int scaling, adaptive_rho, adaptive_rho_interval, max_iter, scaled_termination, check_termination;
};

// ---- PTX helpers: mbarrier + TMA bulk copy ------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- block reductions (results broadcast to every thread; deterministic order) --------------------
template <int K, int NT>
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
template <int NT>
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

__device__ __forceinline__ double limit_scaling(double v) {
    v = v < kMinScaling ? 1.0 : v;
    return v > kMaxScaling ? kMaxScaling : v;
}

// ---- warp-level pieces of the block-tridiagonal factor / solve ---------------------------------------
// u = Dinv_t * v, one entry of v and u per lane (lanes >= bs carry zeros).
template <int BS>
__device__ __forceinline__ double gemv_dinv(const double *__restrict__ Dt, double v, int lane, int bs) {
    const int rl = lane < bs ? lane : bs - 1;
    const int rbase = rl * (rl + 1) / 2;
    double a0 = 0.0, a1 = 0.0;
    if (BS > 0) {
#pragma unroll
        for (int c = 0; c < BS; ++c) {
            const double vc = __shfl_sync(0xffffffffu, v, c);
            const int idx = rl >= c ? rbase + c : c * (c + 1) / 2 + rl;
            const double m = Dt[idx];
            if (c & 1) a1 = fma(m, vc, a1);
            else a0 = fma(m, vc, a0);
        }
    } else {
        for (int c = 0; c < bs; ++c) {
            const double vc = __shfl_sync(0xffffffffu, v, c);
            const int idx = rl >= c ? rbase + c : c * (c + 1) / 2 + rl;
            a0 = fma(Dt[idx], vc, a0);
        }
    }
    const double r = a0 + a1;
    return lane < bs ? r : 0.0;
}

struct OffIdx { // per-lane pattern of one off-diagonal block (<= 4 entries per row / column cached)
    int col[4];
    int trow[4];
    int tslot[4];
};
__device__ __forceinline__ void load_off_idx(const DevSym &S, int t, int lane, int bs, OffIdx &ix) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        ix.col[s] = (lane < bs && s < S.ow) ? (int)S.off_col[((size_t)t * bs + lane) * S.ow + s] : -1;
        const bool okT = lane < bs && s < S.owT;
        ix.trow[s] = okT ? (int)S.offT_row[((size_t)t * bs + lane) * S.owT + s] : -1;
        ix.tslot[s] = okT ? (int)S.offT_slot[((size_t)t * bs + lane) * S.owT + s] : 0;
    }
}
// sum_s Off_t[lane][s] * v[col]   (rows of block t+1 from a vector over block t)
__device__ __forceinline__ double off_mul(const DevSym &S, const double *__restrict__ sOff, int t, double v, int lane, int bs,
                                          const OffIdx &ix, bool cached) {
    double acc = 0.0;
    if (cached) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int a = ix.col[s];
            const double up = __shfl_sync(0xffffffffu, v, a >= 0 ? a : 0);
            if (a >= 0) acc = fma(sOff[((size_t)t * bs + lane) * S.ow + s], up, acc);
        }
    } else {
        for (int s = 0; s < S.ow; ++s) {
            const int a = lane < bs ? (int)S.off_col[((size_t)t * bs + lane) * S.ow + s] : -1;
            const double up = __shfl_sync(0xffffffffu, v, a >= 0 ? a : 0);
            if (a >= 0) acc = fma(sOff[((size_t)t * bs + lane) * S.ow + s], up, acc);
        }
    }
    return acc;
}
// Off_t' * v: (columns of block t from a vector over block t+1)
__device__ __forceinline__ double offT_mul(const DevSym &S, const double *__restrict__ sOff, int t, double v, int lane, int bs,
                                           const OffIdx &ix, bool cached) {
    double acc = 0.0;
    const double *blk = sOff + (size_t)t * bs * S.ow;
    if (cached) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int r = ix.trow[s];
            const double up = __shfl_sync(0xffffffffu, v, r >= 0 ? r : 0);
            if (r >= 0) acc = fma(blk[ix.tslot[s]], up, acc);
        }
    } else {
        for (int s = 0; s < S.owT; ++s) {
            const int r = lane < bs ? (int)S.offT_row[((size_t)t * bs + lane) * S.owT + s] : -1;
            const int slot = lane < bs ? (int)S.offT_slot[((size_t)t * bs + lane) * S.owT + s] : 0;
            const double up = __shfl_sync(0xffffffffu, v, r >= 0 ? r : 0);
            if (r >= 0) acc = fma(blk[slot], up, acc);
        }
    }
    return acc;
}

__device__ __forceinline__ double dinv_at(const double *Dt, int r, int c) {
    return r >= c ? Dt[r * (r + 1) / 2 + c] : Dt[c * (c + 1) / 2 + r];
}
// D_t -= Off_{t-1} Dinv_{t-1} Off_{t-1}'  (above = true)   or   D_t -= Off_t' Dinv_{t+1} Off_t  (above = false); one warp
static __device__ __noinline__ void schur_update(const DevSym &S, double *sDinv, const double *sOff, int t, int lane, int bs, bool above) {
    const int dpb = S.dpb;
    if (above) {
        const int ow = S.ow;
        const double *Dp = sDinv + (size_t)(t - 1) * dpb;
        for (int e = lane; e < dpb; e += 32) {
            const int rc = S.tri_rc[e], r = rc >> 8, c = rc & 0xFF;
            double acc = 0.0;
            for (int s = 0; s < ow; ++s) {
                const int a = S.off_col[((size_t)(t - 1) * bs + r) * ow + s];
                if (a < 0) continue;
                const double va = sOff[((size_t)(t - 1) * bs + r) * ow + s];
                for (int s2 = 0; s2 < ow; ++s2) {
                    const int b = S.off_col[((size_t)(t - 1) * bs + c) * ow + s2];
                    if (b < 0) continue;
                    acc += va * dinv_at(Dp, a, b) * sOff[((size_t)(t - 1) * bs + c) * ow + s2];
                }
            }
            sDinv[(size_t)t * dpb + e] -= acc;
        }
    } else {
        const int owT = S.owT;
        const double *Dn = sDinv + (size_t)(t + 1) * dpb;
        const double *blk = sOff + (size_t)t * bs * S.ow;
        for (int e = lane; e < dpb; e += 32) {
            const int rc = S.tri_rc[e], r = rc >> 8, c = rc & 0xFF;
            double acc = 0.0;
            for (int s = 0; s < owT; ++s) {
                const int a = S.offT_row[((size_t)t * bs + r) * owT + s];
                if (a < 0) continue;
                const double va = blk[S.offT_slot[((size_t)t * bs + r) * owT + s]];
                for (int s2 = 0; s2 < owT; ++s2) {
                    const int b = S.offT_row[((size_t)t * bs + c) * owT + s2];
                    if (b < 0) continue;
                    acc += va * dinv_at(Dn, a, b) * blk[S.offT_slot[((size_t)t * bs + c) * owT + s2]];
                }
            }
            sDinv[(size_t)t * dpb + e] -= acc;
        }
    }
    __syncwarp();
}
// In-place inverse of the packed SPD block M (one warp, Gauss-Jordan without pivoting on a full
// bs x bs scratch copy F; all 32 lanes sweep the bs*bs entries of each elimination step).
static __device__ __noinline__ void invert_block(const DevSym &S, double *M, double *F, int lane, int bs) {
    const int dpb = S.dpb;
    for (int e = lane; e < dpb; e += 32) {
        const int rc = S.tri_rc[e], r = rc >> 8, c = rc & 0xFF;
        const double v = M[e];
        F[r * bs + c] = v;
        F[c * bs + r] = v;
    }
    __syncwarp();
    const int nn = bs * bs;
    for (int k = 0; k < bs; ++k) {
        const double pinv = 1.0 / F[k * bs + k];
        for (int e = lane; e < nn; e += 32) {
            const int i = e / bs, j = e - i * bs;
            if (i != k && j != k) F[e] = fma(-(F[i * bs + k] * pinv), F[k * bs + j], F[e]);
        }
        __syncwarp();
        for (int e = lane; e < bs; e += 32) {
            if (e != k) {
                F[e * bs + k] = -(F[e * bs + k] * pinv);
                F[k * bs + e] *= pinv;
            } else {
                F[k * bs + k] = pinv;
            }
        }
        __syncwarp();
    }
    for (int e = lane; e < dpb; e += 32) {
        const int rc = S.tri_rc[e];
        M[e] = F[(rc >> 8) * bs + (rc & 0xFF)];
    }
    __syncwarp();
}

// ================================================================================================
//                                        the ADMM kernel
// ================================================================================================
#define ROWS_LOOP(j, i) _Pragma("unroll") for (int j = 0, i = tid; j < RPT; ++j, i += NT) if (i < mu)
#define COLS_LOOP(c, p) _Pragma("unroll") for (int c = 0, p = tid; c < CPT; ++c, p += NT) if (p < np)

template <int NT, int RPT, int CPT, int BS>
__global__ void __launch_bounds__(NT, 2) gomp_admm_kernel(const DevSym S, const DevBatch B, const DevSettings st) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bs = BS > 0 ? BS : S.bs;
    const int T = S.T, np = S.n_pad, mu = S.m_used, dpb = S.dpb, h = S.T / 2;
    double *const sA = smem + S.oA;
    double *const sP = smem + S.oP;
    double *const sW = smem + S.oW;
    double *const sRHS = smem + S.oRHS;
    double *const sDinv = smem + S.oDinv;
    double *const sOff = smem + S.oOff;
    double *const sInv = smem + S.oInv;
    double *const sRed = smem + S.oRed;
    uint64_t *const sBar = reinterpret_cast<uint64_t *>(smem + S.oEnd);
    int *const sWork = reinterpret_cast<int *>(smem + S.oEnd + 1);

    const bool small_off = S.ow <= 4 && S.owT <= 4; // pattern of a block row fits the per-lane cache
    const bool cached_off = S.uniform_off && small_off;
    const bool hasq = B.q != nullptr;
    OffIdx oix;
    load_off_idx(S, 0, lane, bs, oix); // valid for every block when the pattern is uniform (T >= 2)

    if (tid == 0) mbar_init(sBar, 1);
    uint32_t bar_phase = 0;
    __syncthreads();

    for (;;) {
        if (tid == 0) *sWork = atomicAdd(B.work_counter, 1);
        __syncthreads();
        const int b = *sWork;
        if (b >= B.batch) break;
        const size_t bmu = (size_t)b * mu, bnp = (size_t)b * np;

        // ------------------------------------------------------------------ load
        const double *gA = B.A_val + (size_t)b * B.A_pitch;
        if (B.bulk_ok) {
            if (tid == 0) {
                mbar_expect_tx(sBar, (uint32_t)S.nnzA * 8u);
                bulk_g2s(sA, gA, (uint32_t)S.nnzA * 8u, sBar);
            }
        } else {
            for (int k = tid; k < S.nnzA; k += NT) sA[k] = gA[k];
        }
        for (int k = tid; k < S.nnzPu; k += NT) sP[k] = S.Pu_val[k];
        const double *gl = B.l + (size_t)b * S.m, *gu = B.u + (size_t)b * S.m;
        double z[RPT], y[RPT], lo[RPT], hi[RPT];
        double x[CPT];
        ROWS_LOOP(j, i) {
            const int orig = S.used_rows[i];
            lo[j] = fmax(gl[orig], -kInfty); // osqp-cpp clips bounds to +-OSQP_INFTY
            hi[j] = fmin(gu[orig], kInfty);
        }
        int bad = 0;
        for (int e = tid; e < S.n_empty; e += NT) {
            const int r = S.empty_rows[e];
            if (!(gl[r] <= 0.0 && 0.0 <= gu[r])) bad = 1;
        }
        if (B.bulk_ok) {
            unsigned spins = 0;
            while (!mbar_try_wait(sBar, bar_phase)) {
                if (++spins > (1u << 26)) __trap(); // a lost bulk copy must not hang the device
            }
            bar_phase ^= 1;
        }
        const int empty_infeasible = __syncthreads_or(bad);

        // ------------------------------------------------------------------ Ruiz equilibration
        double cscale = 1.0;
        {
            double Ei[RPT], Dp[CPT], qv[CPT];
            ROWS_LOOP(j, i) Ei[j] = 1.0;
            COLS_LOOP(c, p) {
                const int var = S.pos_var[p];
                qv[c] = (hasq && var >= 0) ? B.q[(size_t)b * S.n + var] : 0.0;
                Dp[c] = 1.0;
            }
#pragma unroll 1
            for (int it = 0; it < st.scaling; ++it) {
                double dt[CPT], et[RPT];
                COLS_LOOP(c, p) {
                    double nrm = 0.0;
                    const int pl = S.pf_len[p];
                    for (int j = 0; j < pl; ++j) nrm = fmax(nrm, fabs(sP[S.pf_ent[(size_t)j * np + p] & 0xFFFFu]));
                    const int cl = S.col_len[p], cs = S.col_start[p];
                    for (int j = 0; j < cl; ++j) nrm = fmax(nrm, fabs(sA[cs + j]));
                    dt[c] = 1.0 / sqrt(limit_scaling(nrm));
                    sRHS[p] = dt[c];
                }
                ROWS_LOOP(j, i) {
                    double nrm = 0.0;
                    const int rl = S.row_len[i];
                    for (int q = 0; q < rl; ++q) nrm = fmax(nrm, fabs(sA[S.row_ent[(size_t)q * mu + i] & 0xFFFFu]));
                    et[j] = 1.0 / sqrt(limit_scaling(nrm));
                }
                __syncthreads();
                ROWS_LOOP(j, i) {
                    const int rl = S.row_len[i];
                    for (int q = 0; q < rl; ++q) sA[S.row_ent[(size_t)q * mu + i] & 0xFFFFu] *= et[j];
                    Ei[j] *= et[j];
                }
                for (int k = tid; k < S.nnzPu; k += NT) {
                    const uint32_t ij = S.Pu_ij[k];
                    sP[k] = (sP[k] * sRHS[ij >> 16]) * sRHS[ij & 0xFFFFu];
                }
                __syncthreads();
                double psum = 0.0, qn[1] = {0.0};
                COLS_LOOP(c, p) {
                    const int cl = S.col_len[p], cs = S.col_start[p];
                    for (int j = 0; j < cl; ++j) sA[cs + j] *= dt[c];
                    qv[c] *= dt[c];
                    Dp[c] *= dt[c];
                    double nrm = 0.0;
                    const int pl = S.pf_len[p];
                    for (int j = 0; j < pl; ++j) nrm = fmax(nrm, fabs(sP[S.pf_ent[(size_t)j * np + p] & 0xFFFFu]));
                    psum += nrm;
                    qn[0] = fmax(qn[0], fabs(qv[c]));
                }
                psum = block_sum<NT>(psum, sRed, tid);
                block_max<1, NT>(qn, sRed, tid);
                double ct = fmax(psum / (double)S.n, limit_scaling(qn[0]));
                ct = 1.0 / limit_scaling(ct);
                for (int k = tid; k < S.nnzPu; k += NT) sP[k] *= ct;
                COLS_LOOP(c, p) qv[c] *= ct;
                cscale *= ct;
                __syncthreads();
            }
            ROWS_LOOP(j, i) {
                lo[j] *= Ei[j];
                hi[j] *= Ei[j];
                B.Eg[bmu + i] = Ei[j];
            }
            COLS_LOOP(c, p) {
                B.Dg[bnp + p] = Dp[c];
                if (hasq) B.qs[bnp + p] = qv[c];
            }
        }
        const double cinv = 1.0 / cscale;
        const bool uns = !st.scaled_termination;

        // ------------------------------------------------------------------ rho classification
        unsigned ctype = 0; // 2 bits per owned row: 0 inequality, 1 equality, 2 loose
        ROWS_LOOP(j, i) {
            unsigned t = 0;
            if (lo[j] < -kInfty * kMinScaling && hi[j] > kInfty * kMinScaling) t = 2;
            else if (hi[j] - lo[j] < kRhoTol) t = 1;
            ctype |= t << (2 * j);
        }

        // ------------------------------------------------------------------ iterates
        const int flags = B.flags[b];
        double rho;
        int rho_updates;
        if (flags & 1) {
            ROWS_LOOP(j, i) { z[j] = 0.0; y[j] = 0.0; }
            COLS_LOOP(c, p) x[c] = 0.0;
            rho = fmin(fmax(st.rho, kRhoMin), kRhoMax);
            rho_updates = 0;
        } else {
            ROWS_LOOP(j, i) {
                z[j] = B.z_state[bmu + i];
                y[j] = B.y_state[bmu + i];
            }
            COLS_LOOP(c, p) x[c] = B.x_state[bnp + p];
            rho = B.rho_state[b];
            rho_updates = B.rho_updates[b];
        }
        if (flags & 2) { // osqp_warm_start_x: x <- Dinv x0, z <- A x
            COLS_LOOP(c, p) {
                const int var = S.pos_var[p];
                x[c] = var >= 0 ? (1.0 / B.Dg[bnp + p]) * B.x0[(size_t)b * S.n + var] : 0.0;
                sRHS[p] = x[c];
            }
            __syncthreads();
            ROWS_LOOP(j, i) {
                double acc = 0.0;
                const int rl = S.row_len[i];
                for (int q = 0; q < rl; ++q) {
                    const uint32_t e = S.row_ent[(size_t)q * mu + i];
                    acc = fma(sA[e & 0xFFFFu], sRHS[e >> 16], acc);
                }
                z[j] = acc;
            }
            __syncthreads();
        }

        double rho_eq = kRhoEqOverIneq * rho, rinv = 1.0 / rho, rinv_eq = 1.0 / rho_eq;
        const double rinv_min = 1.0 / kRhoMin;
#define RHO_OF(t) ((t) == 0 ? rho : ((t) == 1 ? rho_eq : kRhoMin))
#define RINV_OF(t) ((t) == 0 ? rinv : ((t) == 1 ? rinv_eq : rinv_min))

        // ------------------------------------------------------------------ ADMM loop
        int status = empty_infeasible ? -3 : -10; // OSQP status_val
        int iters_done = 0;
        double pri_res = 0.0, dua_res = 0.0;
        bool need_factor = true, w_dirty = true;

#pragma unroll 1
        for (int iter = 1; status == -10 && iter <= st.max_iter; ++iter) {
            if (need_factor) {
                // ---------------- assemble S and factor it (two-way block elimination, explicit pivot inverses)
                ROWS_LOOP(j, i) {
                    const unsigned t = (ctype >> (2 * j)) & 3u;
                    sW[i] = RHO_OF(t);
                }
                __syncthreads();
                const int ndiag = T * dpb;
                for (int e = tid; e < ndiag; e += NT) {
                    double acc = 0.0;
                    const uint32_t q1 = S.sd_ptr[e + 1];
                    for (uint32_t q = S.sd_ptr[e]; q < q1; ++q) {
                        const gomp::Contrib cn = S.sd_con[q];
                        acc += sW[cn.row] * sA[cn.k1] * sA[cn.k2];
                    }
                    const uint16_t pk = S.sd_pk[e];
                    if (pk != 0xFFFFu) acc += sP[pk];
                    const int kind = S.sd_kind[e];
                    if (kind == 1) acc += st.sigma;
                    else if (kind == 2) acc = 1.0;
                    sDinv[e] = acc;
                }
                for (int e = tid; e < S.noff; e += NT) {
                    double acc = 0.0;
                    const uint32_t q1 = S.so_ptr[e + 1];
                    for (uint32_t q = S.so_ptr[e]; q < q1; ++q) {
                        const gomp::Contrib cn = S.so_con[q];
                        acc += sW[cn.row] * sA[cn.k1] * sA[cn.k2];
                    }
                    const uint16_t pk = S.so_pk[e];
                    if (pk != 0xFFFFu) acc += sP[pk];
                    sOff[e] = acc;
                }
                __syncthreads();
                if (warp == 0) {
                    for (int t = 0; t < h; ++t) {
                        if (t > 0) schur_update(S, sDinv, sOff, t, lane, bs, true);
                        invert_block(S, sDinv + (size_t)t * dpb, sInv, lane, bs);
                    }
                } else if (warp == 1) {
                    for (int t = T - 1; t > h; --t) {
                        if (t < T - 1) schur_update(S, sDinv, sOff, t, lane, bs, false);
                        invert_block(S, sDinv + (size_t)t * dpb, sInv + bs * bs, lane, bs);
                    }
                }
                __syncthreads();
                if (warp == 0) {
                    if (h > 0) schur_update(S, sDinv, sOff, h, lane, bs, true);
                    if (h < T - 1) schur_update(S, sDinv, sOff, h, lane, bs, false);
                    invert_block(S, sDinv + (size_t)h * dpb, sInv, lane, bs);
                }
                __syncthreads();
                need_factor = false;
                w_dirty = true;
            }
            if (w_dirty) {
                ROWS_LOOP(j, i) {
                    const unsigned t = (ctype >> (2 * j)) & 3u;
                    sW[i] = fma(RHO_OF(t), z[j], -y[j]);
                }
                w_dirty = false;
                __syncthreads();
            }
            // ---------------- rhs = sigma x - q + A'(rho z - y)
            COLS_LOOP(c, p) {
                double acc = st.sigma * x[c];
                if (hasq) acc -= B.qs[bnp + p];
                const int cl = S.col_len[p], cs = S.col_start[p];
                for (int j = 0; j < cl; ++j) acc = fma(sA[cs + j], sW[S.col_row[(size_t)j * np + p]], acc);
                sRHS[p] = acc;
            }
            __syncthreads();
            // ---------------- block-tridiagonal solve, in place on sRHS
            if (warp == 0) { // top chain: blocks 0 .. h-1, then the middle block, then back up
                double uprev = 0.0;
                for (int t = 0; t < h; ++t) {
                    double yv = lane < bs ? sRHS[t * bs + lane] : 0.0;
                    if (t > 0) {
                        if (small_off && !cached_off) load_off_idx(S, t - 1, lane, bs, oix);
                        yv -= off_mul(S, sOff, t - 1, uprev, lane, bs, oix, small_off);
                    }
                    uprev = gemv_dinv<BS>(sDinv + (size_t)t * dpb, yv, lane, bs);
                    if (lane < bs) sRHS[t * bs + lane] = uprev;
                }
                named_bar_sync(1, 64);
                double yv = lane < bs ? sRHS[h * bs + lane] : 0.0;
                if (h > 0) {
                    if (small_off && !cached_off) load_off_idx(S, h - 1, lane, bs, oix);
                    yv -= off_mul(S, sOff, h - 1, uprev, lane, bs, oix, small_off);
                }
                if (h < T - 1) {
                    if (small_off && !cached_off) load_off_idx(S, h, lane, bs, oix);
                    const double unext = lane < bs ? sRHS[(h + 1) * bs + lane] : 0.0;
                    yv -= offT_mul(S, sOff, h, unext, lane, bs, oix, small_off);
                }
                double xn = gemv_dinv<BS>(sDinv + (size_t)h * dpb, yv, lane, bs);
                if (lane < bs) sRHS[h * bs + lane] = xn;
                named_bar_sync(2, 64);
                for (int t = h - 1; t >= 0; --t) { // x_t = u_t - Dinv_t Off_t' x_{t+1}
                    if (small_off && !cached_off) load_off_idx(S, t, lane, bs, oix);
                    const double v = offT_mul(S, sOff, t, xn, lane, bs, oix, small_off);
                    const double corr = gemv_dinv<BS>(sDinv + (size_t)t * dpb, v, lane, bs);
                    xn = (lane < bs ? sRHS[t * bs + lane] : 0.0) - corr;
                    if (lane < bs) sRHS[t * bs + lane] = xn;
                }
            } else if (warp == 1) { // bottom chain: blocks T-1 .. h+1, then back down
                double unext = 0.0;
                for (int t = T - 1; t > h; --t) {
                    double yv = lane < bs ? sRHS[t * bs + lane] : 0.0;
                    if (t < T - 1) {
                        if (small_off && !cached_off) load_off_idx(S, t, lane, bs, oix);
                        yv -= offT_mul(S, sOff, t, unext, lane, bs, oix, small_off);
                    }
                    unext = gemv_dinv<BS>(sDinv + (size_t)t * dpb, yv, lane, bs);
                    if (lane < bs) sRHS[t * bs + lane] = unext;
                }
                named_bar_sync(1, 64);
                named_bar_sync(2, 64);
                double xp = lane < bs ? sRHS[h * bs + lane] : 0.0;
                for (int t = h + 1; t < T; ++t) { // x_t = u_t - Dinv_t Off_{t-1} x_{t-1}
                    if (small_off && !cached_off) load_off_idx(S, t - 1, lane, bs, oix);
                    const double v = off_mul(S, sOff, t - 1, xp, lane, bs, oix, small_off);
                    const double corr = gemv_dinv<BS>(sDinv + (size_t)t * dpb, v, lane, bs);
                    xp = (lane < bs ? sRHS[t * bs + lane] : 0.0) - corr;
                    if (lane < bs) sRHS[t * bs + lane] = xp;
                }
            }
            __syncthreads();

            const bool last = iter == st.max_iter;
            const bool can_check = st.check_termination && (iter % st.check_termination == 0);
            const bool adapt_now = st.adaptive_rho && st.adaptive_rho_interval && (iter % st.adaptive_rho_interval == 0);
            const bool do_info = can_check || adapt_now || last;
            // ---------------- x, z, y updates (z~ = A x~ fused with the projection and the dual step)
            COLS_LOOP(c, p) {
                const double xn = st.alpha * sRHS[p] + (1.0 - st.alpha) * x[c];
                if (do_info && hasq) B.dxg[bnp + p] = xn - x[c];
                x[c] = xn;
            }
            ROWS_LOOP(j, i) {
                double zt = 0.0;
                const int rl = S.row_len[i];
                for (int q = 0; q < rl; ++q) {
                    const uint32_t e = S.row_ent[(size_t)q * mu + i];
                    zt = fma(sA[e & 0xFFFFu], sRHS[e >> 16], zt);
                }
                const unsigned t = (ctype >> (2 * j)) & 3u;
                const double zr = st.alpha * zt + (1.0 - st.alpha) * z[j];
                const double zn = fmin(fmax(zr + RINV_OF(t) * y[j], lo[j]), hi[j]);
                const double d = RHO_OF(t) * (zr - zn);
                y[j] += d;
                z[j] = zn;
                sW[i] = fma(RHO_OF(t), zn, -y[j]); // w = rho z - y for the next gather
                if (do_info) B.dyg[bmu + i] = d;
            }
            __syncthreads(); // sRHS is overwritten by the next gather, which reads sW

            if (do_info) {
                // ---------------- residuals and the norms OSQP's tolerances need (update_info)
                double nv[kRedK];
                COLS_LOOP(c, p) sRHS[p] = x[c];
                ROWS_LOOP(j, i) sW[i] = y[j];
                w_dirty = true;
                __syncthreads();
#pragma unroll
                for (int k = 0; k < kRedK; ++k) nv[k] = 0.0;
                ROWS_LOOP(j, i) {
                    double ax = 0.0;
                    const int rl = S.row_len[i];
                    for (int q = 0; q < rl; ++q) {
                        const uint32_t e = S.row_ent[(size_t)q * mu + i];
                        ax = fma(sA[e & 0xFFFFu], sRHS[e >> 16], ax);
                    }
                    const double pr = ax - z[j];
                    const double einv = 1.0 / B.Eg[bmu + i];
                    nv[0] = fmax(nv[0], fabs(einv * pr));
                    nv[1] = fmax(nv[1], fabs(pr));
                    nv[2] = fmax(nv[2], fabs(einv * z[j]));
                    nv[3] = fmax(nv[3], fabs(einv * ax));
                    nv[4] = fmax(nv[4], fabs(z[j]));
                    nv[5] = fmax(nv[5], fabs(ax));
                }
                COLS_LOOP(c, p) {
                    double px = 0.0, aty = 0.0;
                    const int pl = S.pf_len[p];
                    for (int j = 0; j < pl; ++j) {
                        const uint32_t e = S.pf_ent[(size_t)j * np + p];
                        px = fma(sP[e & 0xFFFFu], sRHS[e >> 16], px);
                    }
                    const int cl = S.col_len[p], cs = S.col_start[p];
                    for (int j = 0; j < cl; ++j) aty = fma(sA[cs + j], sW[S.col_row[(size_t)j * np + p]], aty);
                    const double qq = hasq ? B.qs[bnp + p] : 0.0;
                    const double du = qq + px + aty;
                    const double dinv = 1.0 / B.Dg[bnp + p];
                    nv[6] = fmax(nv[6], fabs(dinv * du));
                    nv[7] = fmax(nv[7], fabs(du));
                    nv[8] = fmax(nv[8], fabs(dinv * px));
                    nv[9] = fmax(nv[9], fabs(dinv * aty));
                    nv[10] = fmax(nv[10], fabs(dinv * qq));
                    nv[11] = fmax(nv[11], fabs(px));
                    nv[12] = fmax(nv[12], fabs(aty));
                    nv[13] = fmax(nv[13], fabs(qq));
                }
                block_max<kRedK, NT>(nv, sRed, tid);
                pri_res = uns ? nv[0] : nv[1];
                dua_res = uns ? cinv * nv[6] : nv[7];
                iters_done = iter;

                // ---------------- termination: stage 0 exact tolerances, stage 1 (only at max_iter) 10x tolerances
                bool done = false;
#pragma unroll 1
                for (int stage = 0; stage < 2 && !done; ++stage) {
                    if (stage == 0 && !(can_check || last)) continue;
                    if (stage == 1) {
                        if (adapt_now) { // OSQP adapt_rho: estimate from the scaled residuals
                            const double pr = nv[1] / (fmax(nv[4], nv[5]) + kDivTol);
                            const double dr = nv[7] / (fmax(nv[13], fmax(nv[12], nv[11])) + kDivTol);
                            const double est = fmin(fmax(rho * sqrt(pr / dr), kRhoMin), kRhoMax);
                            if (est > rho * st.adaptive_rho_tolerance || est < rho / st.adaptive_rho_tolerance) {
                                rho = est;
                                rho_eq = kRhoEqOverIneq * rho;
                                rinv = 1.0 / rho;
                                rinv_eq = 1.0 / rho_eq;
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
                    const double eps_prim = ea + er * (uns ? fmax(nv[2], nv[3]) : fmax(nv[4], nv[5]));
                    if (mu == 0 || pri_res < eps_prim) {
                        prc = true;
                    } else {
                        // OSQP is_primal_infeasible on delta_y
                        double dyr[RPT];
                        double nrm[1] = {0.0}, lhs = 0.0;
                        ROWS_LOOP(j, i) {
                            double d = B.dyg[bmu + i];
                            if (hi[j] > kInfty * kMinScaling) {
                                if (lo[j] < -kInfty * kMinScaling) d = 0.0;
                                else d = fmin(d, 0.0);
                            } else if (lo[j] < -kInfty * kMinScaling) {
                                d = fmax(d, 0.0);
                            }
                            dyr[j] = d;
                            nrm[0] = fmax(nrm[0], fabs(uns ? B.Eg[bmu + i] * d : d));
                            lhs += hi[j] * fmax(d, 0.0) + lo[j] * fmin(d, 0.0);
                        }
                        block_max<1, NT>(nrm, sRed, tid);
                        lhs = block_sum<NT>(lhs, sRed, tid);
                        if (nrm[0] > kDivTol && lhs < -epi * nrm[0]) {
                            ROWS_LOOP(j, i) sW[i] = dyr[j];
                            __syncthreads();
                            double v[1] = {0.0};
                            COLS_LOOP(c, p) {
                                double acc = 0.0;
                                const int cl = S.col_len[p], cs = S.col_start[p];
                                for (int j = 0; j < cl; ++j) acc = fma(sA[cs + j], sW[S.col_row[(size_t)j * np + p]], acc);
                                v[0] = fmax(v[0], fabs(uns ? (1.0 / B.Dg[bnp + p]) * acc : acc));
                            }
                            block_max<1, NT>(v, sRed, tid);
                            pic = v[0] < epi * nrm[0];
                        }
                    }
                    const double mx = uns ? cinv * fmax(nv[10], fmax(nv[9], nv[8])) : fmax(nv[13], fmax(nv[12], nv[11]));
                    const double eps_dual = ea + er * mx;
                    if (dua_res < eps_dual) {
                        drc = true;
                    } else if (hasq) {
                        // OSQP is_dual_infeasible on delta_x (q'dx < 0 is impossible when q = 0)
                        double nrm[1] = {0.0}, qdx = 0.0;
                        COLS_LOOP(c, p) {
                            const double d = B.dxg[bnp + p];
                            nrm[0] = fmax(nrm[0], fabs(uns ? B.Dg[bnp + p] * d : d));
                            qdx += B.qs[bnp + p] * d;
                        }
                        block_max<1, NT>(nrm, sRed, tid);
                        qdx = block_sum<NT>(qdx, sRed, tid);
                        const double cs_ = uns ? cscale : 1.0;
                        if (nrm[0] > kDivTol && qdx < -cs_ * edi * nrm[0]) {
                            COLS_LOOP(c, p) sRHS[p] = B.dxg[bnp + p];
                            __syncthreads();
                            double v[1] = {0.0};
                            COLS_LOOP(c, p) {
                                double acc = 0.0;
                                const int pl = S.pf_len[p];
                                for (int j = 0; j < pl; ++j) {
                                    const uint32_t e = S.pf_ent[(size_t)j * np + p];
                                    acc = fma(sP[e & 0xFFFFu], sRHS[e >> 16], acc);
                                }
                                v[0] = fmax(v[0], fabs(uns ? (1.0 / B.Dg[bnp + p]) * acc : acc));
                            }
                            block_max<1, NT>(v, sRed, tid);
                            if (v[0] < cs_ * edi * nrm[0]) {
                                int viol = 0;
                                ROWS_LOOP(j, i) {
                                    double acc = 0.0;
                                    const int rl = S.row_len[i];
                                    for (int q = 0; q < rl; ++q) {
                                        const uint32_t e = S.row_ent[(size_t)q * mu + i];
                                        acc = fma(sA[e & 0xFFFFu], sRHS[e >> 16], acc);
                                    }
                                    if (uns) acc *= 1.0 / B.Eg[bmu + i];
                                    if ((hi[j] < kInfty * kMinScaling && acc > edi * nrm[0]) ||
                                        (lo[j] > -kInfty * kMinScaling && acc < -edi * nrm[0]))
                                        viol = 1;
                                }
                                dic = !__syncthreads_or(viol);
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
            }
        }
#undef RHO_OF
#undef RINV_OF

        // ------------------------------------------------------------------ store
        const bool infeasible = status == -3 || status == 3 || status == -4 || status == 4 || status == -7;
        const double qnan = __longlong_as_double(0x7ff8000000000000LL);
        double *xo = B.x_out + (size_t)b * S.n, *yo = B.y_out + (size_t)b * S.m;
        COLS_LOOP(c, p) {
            const int var = S.pos_var[p];
            if (var >= 0) xo[var] = infeasible ? qnan : B.Dg[bnp + p] * x[c];
            B.x_state[bnp + p] = infeasible ? 0.0 : x[c];
        }
        ROWS_LOOP(j, i) {
            const double e = B.Eg[bmu + i];
            yo[S.used_rows[i]] = infeasible ? qnan : (e * y[j]) * cinv;
            B.z_state[bmu + i] = infeasible ? 0.0 : z[j];
            B.y_state[bmu + i] = infeasible ? 0.0 : y[j];
        }
        for (int e = tid; e < S.n_empty; e += NT) yo[S.empty_rows[e]] = infeasible ? qnan : 0.0;
        if (tid == 0) {
            B.rho_state[b] = rho;
            B.flags[b] = 0;
            B.status[b] = status;
            B.iters[b] = iters_done;
            B.rho_updates[b] = rho_updates;
            B.pri_res[b] = pri_res;
            B.dua_res[b] = dua_res;
        }
        fence_proxy_async(); // generic-proxy accesses to sA are ordered before the next QP's bulk copy
        __syncthreads();
    }
}

#undef ROWS_LOOP
#undef COLS_LOOP

// ---- launch interface (one explicit instantiation per translation unit) ------------------------------
template <int RPT, int CPT, int BS>
cudaError_t prepare_variant(size_t smem_bytes, int *ctas_per_sm) {
    auto fn = gomp_admm_kernel<kThreads, RPT, CPT, BS>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, fn, kThreads, smem_bytes);
}
template <int RPT, int CPT, int BS>
cudaError_t launch_variant(const DevSym &S, const DevBatch &B, const DevSettings &st, int grid, size_t smem_bytes,
                           cudaStream_t stream) {
    gomp_admm_kernel<kThreads, RPT, CPT, BS><<<grid, kThreads, smem_bytes, stream>>>(S, B, st);
    return cudaGetLastError();
}

} // namespace gomp_dev
#endif // GOMP_KERNEL_CUH
