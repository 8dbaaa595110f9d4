// gomp_kernel.cuh -- the persistent ADMM kernel of libgomp_cuda (sm_100a, fp64) and its launchers.
//
// One 512-thread CTA per SM pulls QPs from a device work queue and runs the whole OSQP algorithm
// for each of them on chip.  Roles inside the CTA:
//   * warps 0 and 1 are SOLVER warps.  They own the block-tridiagonal linear algebra: the two-way
//     ("twisted") block factorisation of S = P + sigma I + A' diag(rho) A -- warp 0 eliminates
//     stages 0..h-1 downwards, warp 1 stages T-1..h+1 upwards, they meet at the middle stage h --
//     with explicit inverses of the pivot blocks (register-resident Gauss-Jordan, one matrix
//     column per lane), and the forward / backward block substitutions of every ADMM iteration
//     (one 2N x 2N GEMV per stage, operands broadcast through shared memory).
//   * warps 2..15 are OWNER warps (448 threads).  Each thread owns a fixed set of constraint rows
//     and variables; the ADMM iterates nobody else touches (z, y, l, u per row; x per variable)
//     live in its REGISTERS for the whole solve.  Owners do the fused SpMV + projection + dual
//     update, the A' gather, Ruiz equilibration, residuals and norms.
// Shared memory holds what is exchanged: the QP's A values (staged with one TMA bulk copy,
// cp.async.bulk + mbarrier), the scaled objective P, w = rho z - y, the right-hand side / x~, the
// factor (full rows of every inverted pivot block + ELL off-diagonal blocks) and -- loaded once per
// kernel, shared by every QP the CTA processes -- the sparsity-pattern index arrays.
//
// Algorithm per QP, following OSQP 0.6 as restated in oracle/osqp_ref.c:
//   load -> Ruiz equilibration -> rho classification -> [warm start] ->
//   { factor ; ADMM iterations ; every check_termination iterations residuals / termination /
//     infeasibility ; adaptive rho -> refactor } -> store.
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

constexpr int kThreads = 512;              // 16 warps
constexpr int kSolverThreads = 64;         // warps 0, 1
constexpr int kOwners = kThreads - kSolverThreads; // 448 row / variable owners
constexpr int kRedK = 14;                  // widest multi-value block reduction
// On-device phase timers (GOMP_PROFILE_PHASES at run time) are compiled in only with -DGOMP_PHASE_TIMERS=1
// (make PROFILE=1): they keep a 64-bit time stamp alive through the whole solve, which the 128-register budget feels.
#ifndef GOMP_PHASE_TIMERS
#define GOMP_PHASE_TIMERS 0
#endif
constexpr bool kTimers = GOMP_PHASE_TIMERS != 0;

// ---- device-side views ------------------------------------------------------------------------
struct DevSym {
    int n, m, nnzA, nnzPu, m_used, bs, T, n_pad, ow, owT, n_empty, noff, npf;
    int uniform_off; // every off-diagonal block has the same sparsity pattern
    int bounds_in_smem; // scaled l, u of the QP fit in shared memory (else: per-QP global scratch)
    int p_in_smem; // the P-side pattern arrays (pf_csr, pf_ptr, Pu_ij) fit in shared memory
    const int32_t *pos_var, *used_rows, *empty_rows;
    // pattern arrays (global copies; staged into shared memory once per kernel)
    const uint32_t *row_csr, *row_range, *col_range, *pf_csr, *Pu_ij;
    const uint16_t *row_ptr, *csc_row, *pf_ptr, *offT_slot;
    const int8_t *off_col, *offT_row;
    const double *Pu_val;
    // assembly lists of S (global)
    const uint32_t *sd_ptr;
    const gomp::Contrib *sd_con;
    const uint16_t *sd_pk;
    const uint8_t *sd_kind; // 0 plain entry, 1 diagonal, 2 diagonal of a padding variable
    const uint32_t *so_ptr;
    const gomp::Contrib *so_con;
    const uint16_t *so_pk;
    const uint16_t *tri_rc; // packed lower index -> (row << 8 | col)
    // shared-memory carve-up (byte offsets)
    int oA, oP, oW, oRHS, oLo, oHi, oDinv, oOff, oV, oScr, oRed, oRowRange, oRowEnt, oColRange, oPfEnt, oPuIJ, oRowPtr, oCscRow, oPfPtr, oOffTslot,
        oOffCol, oOffTrow, oBar;
};

struct DevBatch {
    int batch;
    int bulk_ok; // A rows are 16-byte aligned and a multiple of 16 bytes: stage with cp.async.bulk
    const double *A_val, *l, *u, *q, *x0;
    long long A_pitch;
    double *x_state, *z_state, *y_state, *rho_state;
    int flag_bits; // bit0: cold start, bit1: primal warm start pending (whole batch)
    double *Eg, *Dg, *qs, *dyg, *dxg, *Lg, *Ug;
    double *x_out, *y_out;
    int *status, *iters, *rho_updates;
    double *pri_res, *dua_res, *obj_val;
    int *work_counter;
    const int *active; // optional per-QP mask (NULL = all): QPs with 0 are skipped, their state and outputs stay
    const int *order;  // optional work-queue order (NULL = by index): longest expected solves first
    unsigned long long *prof; // optional: per-phase cycle counters (thread 0 of every CTA adds), NULL = off
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
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- reductions over the owner warps (named barrier 3; results broadcast to every owner; fixed order) ----
__device__ __forceinline__ void owners_sync() { named_bar_sync(3, kOwners); }
template <int K>
__device__ __forceinline__ void owners_max(double (&v)[K], double *sRed, int oid) {
    constexpr int NW = kOwners / 32;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double a = v[k];
#pragma unroll
        for (int o = 16; o; o >>= 1) a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o));
        if ((oid & 31) == 0) sRed[k * NW + (oid >> 5)] = a;
    }
    owners_sync();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double a = sRed[k * NW];
#pragma unroll
        for (int j = 1; j < NW; ++j) a = fmax(a, sRed[k * NW + j]);
        v[k] = a;
    }
    owners_sync();
}
__device__ __forceinline__ double owners_sum(double a, double *sRed, int oid) {
    constexpr int NW = kOwners / 32;
#pragma unroll
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((oid & 31) == 0) sRed[oid >> 5] = a;
    owners_sync();
    double r = sRed[0];
#pragma unroll
    for (int j = 1; j < NW; ++j) r += sRed[j];
    owners_sync();
    return r;
}

__device__ __forceinline__ double limit_scaling(double v) {
    v = v < kMinScaling ? 1.0 : v;
    return v > kMaxScaling ? kMaxScaling : v;
}

// Everything the solver warps need, all in shared memory.  Only the base pointer lives in a register;
// the region offsets come straight from the kernel-parameter constant bank.
struct SolverCtx {
    uint32_t sm; // shared-window address of the dynamic shared memory base (keeps every access an LDS/STS)
    int oDinv, oOff, oRHS, oV, oScr, oOffCol, oOffTrow, oOffTslot;
    int bs, T, ow, owT;
    template <typename U>
    __device__ __forceinline__ U *at(int off) const { return reinterpret_cast<U *>(__cvta_shared_to_generic((size_t)(sm + (uint32_t)off))); }
    __device__ __forceinline__ double *dinv() const { return at<double>(oDinv); }  // [T][bs][bs] inverted pivot blocks (full rows)
    __device__ __forceinline__ double *off() const { return at<double>(oOff); }    // [(T-1)][bs][ow] ELL off-diagonal values
    __device__ __forceinline__ double *rhs() const { return at<double>(oRHS); }    // [n_pad]
    __device__ __forceinline__ double *vbuf() const { return at<double>(oV); }     // [32] per-warp broadcast buffer
    __device__ __forceinline__ const int8_t *offCol() const { return at<const int8_t>(oOffCol); }
    __device__ __forceinline__ const int8_t *offTrow() const { return at<const int8_t>(oOffTrow); }
    __device__ __forceinline__ const uint16_t *offTslot() const { return at<const uint16_t>(oOffTslot); }
};

// u = M_row . v   (row of the lane's pivot-block inverse times the broadcast vector)
template <int BS>
__device__ __forceinline__ double row_dot(const double *__restrict__ Mrow, const double *__restrict__ v, int bs) {
    if (BS > 0) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
        for (int c = 0; c < BS; c += 2) {
            const double2 m = *reinterpret_cast<const double2 *>(Mrow + c);
            const double2 w = *reinterpret_cast<const double2 *>(v + c);
            if ((c >> 1) & 1) {
                a2 = fma(m.x, w.x, a2);
                a3 = fma(m.y, w.y, a3);
            } else {
                a0 = fma(m.x, w.x, a0);
                a1 = fma(m.y, w.y, a1);
            }
        }
        return (a0 + a1) + (a2 + a3);
    } else {
        double a0 = 0.0;
        for (int c = 0; c < bs; ++c) a0 = fma(Mrow[c], v[c], a0);
        return a0;
    }
}
// sum_s Off_t[lane][s] * src[col]: rows of stage t+1 from a vector over stage t (lane < bs)
__device__ __forceinline__ double off_mul(const SolverCtx &C, int t, const double *__restrict__ src, int lane) {
    double acc = 0.0;
    const int base = (t * C.bs + lane) * C.ow;
    for (int s = 0; s < C.ow; ++s) {
        const int a = C.offCol()[base + s];
        if (a >= 0) acc = fma(C.off()[base + s], src[a], acc);
    }
    return acc;
}
// (Off_t' src)[lane]: columns of stage t from a vector over stage t+1 (lane < bs)
__device__ __forceinline__ double offT_mul(const SolverCtx &C, int t, const double *__restrict__ src, int lane) {
    double acc = 0.0;
    const int base = (t * C.bs + lane) * C.owT;
    const double *blk = C.off() + (size_t)t * C.bs * C.ow;
    for (int s = 0; s < C.owT; ++s) {
        const int r = C.offTrow()[base + s];
        if (r >= 0) acc = fma(blk[C.offTslot()[base + s]], src[r], acc);
    }
    return acc;
}
// one substitution step: returns (Dinv_t * yv)[lane]  (yv distributed one entry per lane)
template <int BS>
__device__ __forceinline__ double apply_dinv(const SolverCtx &C, int t, double yv, int lane) {
    const int bs = BS > 0 ? BS : C.bs;
    if (lane < bs) C.vbuf()[lane] = yv;
    __syncwarp();
    const int rl = lane < bs ? lane : bs - 1;
    const double u = row_dot<BS>(C.dinv() + ((size_t)t * bs + rl) * bs, C.vbuf(), bs);
    __syncwarp();
    return u;
}

// ---- substitution step of the fast path (compile-time block size, <= 2 off-diagonal entries per block
// row / column).  The previous stage's result stays in a register of its lane and is gathered with two
// shuffles; everything independent of it -- the lane's row of the inverted pivot block, its right-hand
// side entry, the off-diagonal values -- is loaded first.  The block vector is broadcast through a
// double-buffered shared-memory line, so one __syncwarp per step.
struct OffPattern { // per-lane indices of one off-diagonal block row (forward) and column (transposed)
    int oi0, oi1;   // columns of Off_t[lane][.] (-1: none)
    int ti0, ti1;   // rows r with Off_t[r][lane] != 0 (-1: none)
    int ts0, ts1;   // their value slots inside block t
};
__device__ __forceinline__ void load_off_pattern(const SolverCtx &C, int tb, int rl, OffPattern &P) {
    const int bf = (tb * C.bs + rl) * C.ow, bt = (tb * C.bs + rl) * C.owT;
    P.oi0 = C.offCol()[bf];
    P.oi1 = C.ow > 1 ? C.offCol()[bf + 1] : -1;
    P.ti0 = C.offTrow()[bt];
    P.ts0 = C.offTslot()[bt];
    P.ti1 = C.owT > 1 ? C.offTrow()[bt + 1] : -1;
    P.ts1 = C.owT > 1 ? C.offTslot()[bt + 1] : 0;
}
// fwd:  returns (Dinv_t (b - off . uprev))[lane]        bwd:  returns (b - Dinv_t (off . uprev))[lane]
// tb < 0: no off-diagonal part.  transposed: use Off_tb' (columns of stage tb from stage tb+1).
template <int BS>
__device__ __forceinline__ double chain_step(const SolverCtx &C, int t, int tb, bool transposed, bool fwd, double uprev, int lane,
                                             const OffPattern &P) {
    constexpr int NB = BS > 0 ? BS : 2;
    const int rl = lane < NB ? lane : NB - 1;
    // ---- independent of the previous stage
    const double *row = C.dinv() + ((size_t)t * NB + rl) * NB;
    double m[NB];
#pragma unroll
    for (int c = 0; c < NB; c += 2) {
        const double2 d = *reinterpret_cast<const double2 *>(row + c);
        m[c] = d.x;
        m[c + 1] = d.y;
    }
    const double b = C.rhs()[t * NB + rl];
    double ov0 = 0.0, ov1 = 0.0;
    int s0 = 0, s1 = 0;
    if (tb >= 0) {
        if (!transposed) {
            const double *ob = C.off() + (size_t)(tb * NB + rl) * C.ow;
            s0 = P.oi0;
            s1 = P.oi1;
            if (s0 >= 0) ov0 = ob[0];
            if (s1 >= 0) ov1 = ob[1];
        } else {
            const double *blk = C.off() + (size_t)tb * NB * C.ow;
            s0 = P.ti0;
            s1 = P.ti1;
            if (s0 >= 0) ov0 = blk[P.ts0];
            if (s1 >= 0) ov1 = blk[P.ts1];
        }
    }
    // ---- dependent part
    const double up0 = __shfl_sync(0xffffffffu, uprev, s0 >= 0 ? s0 : 0);
    const double up1 = __shfl_sync(0xffffffffu, uprev, s1 >= 0 ? s1 : 0);
    const double off = fma(ov0, up0, ov1 * up1);
    const double v = fwd ? b - off : off;
    double *vb = C.vbuf() + (t & 1) * 16;
    if (lane < NB) vb[lane] = v;
    __syncwarp();
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
    for (int c = 0; c < NB; c += 2) {
        const double2 w = *reinterpret_cast<const double2 *>(vb + c);
        if ((c >> 1) & 1) {
            a2 = fma(m[c], w.x, a2);
            a3 = fma(m[c + 1], w.y, a3);
        } else {
            a0 = fma(m[c], w.x, a0);
            a1 = fma(m[c + 1], w.y, a1);
        }
    }
    const double acc = (a0 + a1) + (a2 + a3);
    const double res = fwd ? acc : b - acc;
    if (lane < NB) C.rhs()[t * NB + lane] = res;
    return res;
}

// ================================================================================================
//                                        the ADMM kernel
// ================================================================================================
// phase timers: thread 0 accumulates clock64() deltas per phase when B.prof is set
enum { PH_LOAD = 0, PH_SCALE, PH_FACTOR, PH_RHS, PH_SOLVE, PH_UPDATE, PH_INFO, PH_STORE, PH_ASSEMBLE, PH_CHAIN, PH_COUNT };
#define PHASE_MARK(ph)                                                              \
    do {                                                                            \
        if (kTimers && B.prof != nullptr && tid == kSolverThreads) {                           \
            const long long now__ = clock64();                                      \
            atomicAdd(&B.prof[ph], (unsigned long long)(now__ - t_phase));          \
            t_phase = now__;                                                        \
        }                                                                           \
    } while (0)

#define ROWS_LOOP(j, i) _Pragma("unroll") for (int j = 0, i = oid; j < RPT; ++j, i += kOwners) if (i < mu)
#define COLS_LOOP(c, p) _Pragma("unroll") for (int c = 0, p = oid; c < CPT; ++c, p += kOwners) if (p < np)
// for every entry e = (pos << 16 | k) of compact row i
#define ROW_ENTRIES(i, e) \
    for (int q__ = (int)(sRowRange[i] >> 8), q1__ = q__ + (int)(sRowRange[i] & 0xFFu); q__ < q1__; ++q__) \
        if (const uint32_t e = sRowEnt[q__]; true)
// for every CSC value index k of the variable at position p
#define COL_ENTRIES(p, k) for (int k = (int)(sColRange[p] >> 8), k1__ = k + (int)(sColRange[p] & 0xFFu); k < k1__; ++k)
// for every entry e = (other_pos << 16 | kPu) of the full-symmetric P at position p
#define PF_ENTRIES(p, e) \
    for (int q__ = sPfPtr[p], q1__ = sPfPtr[(p) + 1]; q__ < q1__; ++q__) \
        if (const uint32_t e = sPfEnt[q__]; true)

// ---- raw shared-memory accessors (32-bit shared-window addresses): the substitution sweeps below are
// latency-critical single-warp code, so their loads/stores are written out explicitly
__device__ __forceinline__ double lds_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double2 lds_f64x2(uint32_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

// One substitution sweep over `count` stages starting at stage t0 and moving by dir (+1 / -1), in place on rhs.
//   fwd:  rhs_t <- Dinv_t (rhs_t - O . prev)      bwd:  rhs_t <- rhs_t - Dinv_t (O . prev)
// where prev is the previous stage's result (kept in a register per lane, gathered with two shuffles) and O
// is the off-diagonal block tb = t + tbo (its transpose when the per-lane offsets describe a column).
// Per-lane pattern constants (uniform over the stages): g0/g1 source lanes, o0/o1 byte offsets of the two
// values inside block tb, has0/has1.  Returns the last stage's result.
template <int BS, bool fwd>
__device__ __forceinline__ double chain_sweep_impl(uint32_t dinv0, uint32_t rhs0, uint32_t off0, uint32_t vb0, int t0, int count, int dir,
                                                   int tbo, int nblk, uint32_t blk_bytes, double prev, int lane, int g0, int g1,
                                                   uint32_t o0, uint32_t o1, bool has0, bool has1) {
    // Two lanes per block row: lane = r + 16*half; half 0 multiplies columns [0, CH), half 1 columns [CH, NB);
    // CH is even so both halves start 16-byte aligned.  Halves the loads and FMAs each lane issues per step.
    constexpr int NB = BS > 0 ? BS : 2;
    constexpr int CH = ((NB / 2) + 1) & ~1;              // 8 for NB = 14
    constexpr int NC = CH > NB - CH ? CH : NB - CH;       // columns a lane handles at most
    const int r = lane & 15, half = lane >> 4;
    const int rl = r < NB ? r : NB - 1;
    const bool act = half == 0 && r < NB;                 // the lane that owns row r's scalar work
    const int c0 = half ? CH : 0, nc = half ? NB - CH : CH;
    // running shared-window addresses (one add per stage instead of re-deriving them from the stage index)
    uint32_t arow = dinv0 + (uint32_t)((t0 * NB + rl) * NB + c0) * 8u; // this lane's half row of the stage's inverse
    uint32_t ab = rhs0 + (uint32_t)(t0 * NB + rl) * 8u;                 // its right-hand side entry
    int tb = t0 + tbo;                                                   // off-diagonal block of the stage
    uint32_t ao = off0 + (uint32_t)tb * blk_bytes;                       // (only dereferenced when 0 <= tb < nblk)
    uint32_t par = (uint32_t)(t0 & 1) * 128u;                            // double-buffered broadcast line: vb0 + 0 / + 128
    const uint32_t darow = (uint32_t)(dir * NB * NB * 8), dab = (uint32_t)(dir * NB * 8), dao = (uint32_t)dir * blk_bytes;
#pragma unroll 1
    for (int it = 0; it < count; ++it) {
        // ---- independent of the previous stage
        double m[NC];
#pragma unroll
        for (int c = 0; c < NC; c += 2) {
            if (c < nc) {
                const double2 d = lds_f64x2(arow + (uint32_t)c * 8u);
                m[c] = d.x;
                m[c + 1] = d.y;
            } else {
                m[c] = 0.0;
                m[c + 1] = 0.0;
            }
        }
        const double b = lds_f64(ab);
        const bool blk = (unsigned)tb < (unsigned)nblk;
        const double ov0 = (blk && has0) ? lds_f64(ao + o0) : 0.0;
        const double ov1 = (blk && has1) ? lds_f64(ao + o1) : 0.0;
        // ---- dependent chain (prev[r] is held by both lanes of row r; sources are lanes < 16)
        const double up0 = __shfl_sync(0xffffffffu, prev, g0);
        const double up1 = __shfl_sync(0xffffffffu, prev, g1);
        const double off = fma(ov0, up0, ov1 * up1);
        const double v = fwd ? b - off : off;
        const uint32_t avb = vb0 + par;
        if (act) sts_f64(avb + (uint32_t)r * 8u, v);
        __syncwarp();
        double2 w[NC / 2];
#pragma unroll
        for (int c = 0; c < NC; c += 2) w[c / 2] = c < nc ? lds_f64x2(avb + (uint32_t)(c0 + c) * 8u) : make_double2(0.0, 0.0);
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int c = 0; c < NC; c += 2) {
            a0 = fma(m[c], w[c / 2].x, a0);
            a1 = fma(m[c + 1], w[c / 2].y, a1);
        }
        double acc = a0 + a1;
        acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        prev = fwd ? acc : b - acc;
        if (act) sts_f64(ab, prev);
        arow += darow;
        ab += dab;
        ao += dao;
        tb += dir;
        par ^= 128u;
    }
    return prev;
}
// Out-of-line instances (one per sweep direction of the recurrence): scheduled and register-allocated on their own,
// whatever surrounds the call -- the inlined copies lost 20-30 % to the code around them.
template <int BS, bool fwd>
static __device__ __noinline__ double chain_sweep_call(uint32_t dinv0, uint32_t rhs0, uint32_t off0, uint32_t vb0, int t0, int count, int dir,
                                                       int tbo, int nblk, uint32_t blk_bytes, double prev, int lane, int g0, int g1,
                                                       uint32_t o0, uint32_t o1, bool has0, bool has1) {
    return chain_sweep_impl<BS, fwd>(dinv0, rhs0, off0, vb0, t0, count, dir, tbo, nblk, blk_bytes, prev, lane, g0, g1, o0, o1, has0, has1);
}
// (used by the streamed kernel, where the out-of-line instances are 10 % faster than inlined copies; the resident
// kernel keeps the inlined chain_sweep below, which is the faster one inside solver_role)
template <int BS>
__device__ __forceinline__ double chain_sweep_ool(uint32_t dinv0, uint32_t rhs0, uint32_t off0, uint32_t vb0, int t0, int count, int dir,
                                                  int tbo, int nblk, uint32_t blk_bytes, bool fwd, double prev, int lane, int g0, int g1,
                                                  uint32_t o0, uint32_t o1, bool has0, bool has1) {
    return fwd ? chain_sweep_call<BS, true>(dinv0, rhs0, off0, vb0, t0, count, dir, tbo, nblk, blk_bytes, prev, lane, g0, g1, o0, o1, has0, has1)
               : chain_sweep_call<BS, false>(dinv0, rhs0, off0, vb0, t0, count, dir, tbo, nblk, blk_bytes, prev, lane, g0, g1, o0, o1, has0, has1);
}

// Inlined form of the same sweep (stage addresses derived from the stage index), as used by solver_role.
template <int BS>
__device__ __forceinline__ double chain_sweep(uint32_t dinv0, uint32_t rhs0, uint32_t off0, uint32_t vb0, int t0, int count, int dir,
                                              int tbo, int nblk, uint32_t blk_bytes, bool fwd, double prev, int lane, int g0, int g1,
                                              uint32_t o0, uint32_t o1, bool has0, bool has1) {
    // Two lanes per block row: lane = r + 16*half; half 0 multiplies columns [0, CH), half 1 columns [CH, NB);
    // CH is even so both halves start 16-byte aligned.  Halves the loads and FMAs each lane issues per step.
    constexpr int NB = BS > 0 ? BS : 2;
    constexpr int CH = ((NB / 2) + 1) & ~1;              // 8 for NB = 14
    constexpr int NC = CH > NB - CH ? CH : NB - CH;       // columns a lane handles at most
    const int r = lane & 15, half = lane >> 4;
    const int rl = r < NB ? r : NB - 1;
    const bool act = half == 0 && r < NB;                 // the lane that owns row r's scalar work
    const int c0 = half ? CH : 0, nc = half ? NB - CH : CH;
    int t = t0;
#pragma unroll 1
    for (int it = 0; it < count; ++it, t += dir) {
        // ---- independent of the previous stage
        const uint32_t arow = dinv0 + (uint32_t)((t * NB + rl) * NB + c0) * 8u;
        double m[NC];
#pragma unroll
        for (int c = 0; c < NC; c += 2) {
            if (c < nc) {
                const double2 d = lds_f64x2(arow + (uint32_t)c * 8u);
                m[c] = d.x;
                m[c + 1] = d.y;
            } else {
                m[c] = 0.0;
                m[c + 1] = 0.0;
            }
        }
        const uint32_t ab = rhs0 + (uint32_t)(t * NB + rl) * 8u;
        const double b = lds_f64(ab);
        const int tb = t + tbo;
        const bool blk = tb >= 0 && tb < nblk;
        const uint32_t ao = off0 + (uint32_t)(blk ? tb : 0) * blk_bytes;
        const double ov0 = (blk && has0) ? lds_f64(ao + o0) : 0.0;
        const double ov1 = (blk && has1) ? lds_f64(ao + o1) : 0.0;
        // ---- dependent chain (prev[r] is held by both lanes of row r; sources are lanes < 16)
        const double up0 = __shfl_sync(0xffffffffu, prev, g0);
        const double up1 = __shfl_sync(0xffffffffu, prev, g1);
        const double off = fma(ov0, up0, ov1 * up1);
        const double v = fwd ? b - off : off;
        const uint32_t avb = vb0 + (uint32_t)(t & 1) * 128u;
        if (act) sts_f64(avb + (uint32_t)r * 8u, v);
        __syncwarp();
        double2 w[NC / 2];
#pragma unroll
        for (int c = 0; c < NC; c += 2) w[c / 2] = c < nc ? lds_f64x2(avb + (uint32_t)(c0 + c) * 8u) : make_double2(0.0, 0.0);
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int c = 0; c < NC; c += 2) {
            a0 = fma(m[c], w[c / 2].x, a0);
            a1 = fma(m[c + 1], w[c / 2].y, a1);
        }
        double acc = a0 + a1;
        acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        prev = fwd ? acc : b - acc;
        if (act) sts_f64(ab, prev);
    }
    return prev;
}

// ---- cooperative block factorisation -----------------------------------------------------------------
// The two elimination chains (stages 0..h-1 downwards, T-1..h+1 upwards) are sequential over stages, and
// inverting one pivot block is sequential over its bs pivots; a single warp spends ~10 cycles per
// dependent instruction on that.  So each chain is run by a GROUP of 7 warps (224 threads) in which every
// thread owns matrix entries in registers: per pivot step the owners of row k and column k publish them
// to shared memory, one named barrier, every thread updates its own entries.  Critical path per pivot:
// publish -> barrier -> 3 shared loads -> reciprocal -> 2 FMAs.
constexpr int kFactorGroup = 224; // threads per chain group (warps 0-6: top chain, warps 7-13: bottom chain)

template <int BS>
__device__ __forceinline__ void factor_block_group(const DevSym &S, unsigned char *smem_raw, int t, int mode, int g, int gl) {
    const int bs = BS > 0 ? BS : S.bs;
    const int nn = bs * bs;
    constexpr int EPL = BS > 0 ? (BS * BS + kFactorGroup - 1) / kFactorGroup : 5; // entries per thread (bs <= 32)
    double *const Dall = reinterpret_cast<double *>(smem_raw + S.oDinv);
    const double *const Offall = reinterpret_cast<const double *>(smem_raw + S.oOff);
    const int8_t *const offCol = reinterpret_cast<const int8_t *>(smem_raw + S.oOffCol);
    const int8_t *const offTrow = reinterpret_cast<const int8_t *>(smem_raw + S.oOffTrow);
    const uint16_t *const offTslot = reinterpret_cast<const uint16_t *>(smem_raw + S.oOffTslot);
    double *const pub = reinterpret_cast<double *>(smem_raw + S.oScr) + g * 128; // [parity][row k (32) | column k (32)]
    double *M = Dall + (size_t)t * nn;
    double val[EPL];
    int ri[EPL], ci[EPL];
#pragma unroll
    for (int q = 0; q < EPL; ++q) {
        const int e = gl + q * kFactorGroup;
        ri[q] = e < nn ? e / bs : -1;
        ci[q] = e < nn ? e - ri[q] * bs : -1;
        val[q] = e < nn ? M[e] : 0.0;
    }
    // neighbour updates: minus Off_{t-1} Dinv_{t-1} Off_{t-1}' (mode bit 1), minus Off_t' Dinv_{t+1} Off_t (bit 2)
#pragma unroll 1
    for (int pass = 1; pass <= 2; ++pass) {
        if (!(mode & pass)) continue;
        const bool above = pass == 1;
        const int tb = above ? t - 1 : t;
        const double *Dn = Dall + (size_t)(above ? t - 1 : t + 1) * nn;
        const double *blk = Offall + (size_t)tb * bs * S.ow;
        const int cnt = above ? S.ow : S.owT;
#pragma unroll
        for (int q = 0; q < EPL; ++q) {
            if (ri[q] < 0) continue;
            const int r = ri[q], c = ci[q];
            double acc = 0.0;
            for (int s1 = 0; s1 < cnt; ++s1) {
                int a;
                double va = 0.0;
                if (above) {
                    a = offCol[(tb * bs + r) * S.ow + s1];
                    if (a >= 0) va = blk[r * S.ow + s1];
                } else {
                    a = offTrow[(tb * bs + r) * S.owT + s1];
                    if (a >= 0) va = blk[offTslot[(tb * bs + r) * S.owT + s1]];
                }
                if (a < 0) continue;
                for (int s2 = 0; s2 < cnt; ++s2) {
                    int b2;
                    double vb = 0.0;
                    if (above) {
                        b2 = offCol[(tb * bs + c) * S.ow + s2];
                        if (b2 >= 0) vb = blk[c * S.ow + s2];
                    } else {
                        b2 = offTrow[(tb * bs + c) * S.owT + s2];
                        if (b2 >= 0) vb = blk[offTslot[(tb * bs + c) * S.owT + s2]];
                    }
                    if (b2 >= 0) acc += va * Dn[a * bs + b2] * vb;
                }
            }
            val[q] -= acc;
        }
    }
    if (BS > 0 && BS <= 16 && BS % 2 == 0 && EPL == 1) {
        // Gauss-Jordan sweeps with 2x2 pivot blocks (SPD block, no pivoting): half the barriers and half the
        // dependent reciprocal chains of the scalar version below.  Sweep of the index pair K = {k, k+1}:
        //   M_KK <- P^-1,  M_Kj <- P^-1 M_Kj,  M_iK <- -M_iK P^-1,  M_ij <- M_ij - M_iK P^-1 M_Kj
        // with P^-1 = adj(P) / det(P); everything but the final scaling by 1/det is independent of the division.
        const int i = ri[0], j = ci[0];
        double v = val[0];
#pragma unroll 1
        for (int k = 0; k < BS; k += 2) {
            double *buf = pub + ((k >> 1) & 1) * 64; // [row k | row k+1 | column k | column k+1], 16 doubles each
            if (i >= 0) {
                if (i == k) buf[j] = v;
                if (i == k + 1) buf[16 + j] = v;
                if (j == k) buf[32 + i] = v;
                if (j == k + 1) buf[48 + i] = v;
            }
            named_bar_sync(4 + g, kFactorGroup);
            if (i < 0) continue;
            const double a = buf[k], b = buf[k + 1], c = buf[16 + k], d = buf[16 + k + 1]; // P = [a b; c d]
            const double r1 = buf[j], r2 = buf[16 + j];                                     // M_Kj
            const double c1 = buf[32 + i], c2 = buf[48 + i];                                // M_iK
            const double invdet = 1.0 / (a * d - b * c);
            const bool ik = (i >> 1) == (k >> 1), jk = (j >> 1) == (k >> 1);
            const double u1 = d * r1 - b * r2, u2 = a * r2 - c * r1; // adj(P) M_Kj
            const double w1 = c1 * d - c2 * c, w2 = c2 * a - c1 * b; // M_iK adj(P)
            // all four cases computed, one selected: no divergent paths inside the warp
            const double gen = fma(-(c1 * u1 + c2 * u2), invdet, v);
            const double rowv = (i == k ? u1 : u2) * invdet;
            const double colv = -((j == k ? w1 : w2) * invdet);
            const double piv = (i == k ? (j == k ? d : -b) : (j == k ? -c : a)) * invdet;
            v = ik ? (jk ? piv : rowv) : (jk ? colv : gen);
        }
        val[0] = v;
    } else
    // Gauss-Jordan without pivoting (SPD block), entries in registers
#pragma unroll 1
    for (int k = 0; k < bs; ++k) {
        double *rowk = pub + (k & 1) * 64, *colk = rowk + 32;
#pragma unroll
        for (int q = 0; q < EPL; ++q) {
            if (ri[q] == k) rowk[ci[q]] = val[q];
            if (ci[q] == k) colk[ri[q]] = val[q];
        }
        named_bar_sync(4 + g, kFactorGroup);
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
#pragma unroll
    for (int q = 0; q < EPL; ++q)
        if (ri[q] >= 0) M[gl + q * kFactorGroup] = val[q];
    named_bar_sync(4 + g, kFactorGroup); // the next stage of this chain reads the inverse just stored
}

// phase 0: both chains (threads 0..447);  phase 1: the middle stage (group 0) after a CTA-wide barrier
// One shared out-of-line copy: the factorisation runs once or twice per QP, from both roles; keeping it out of the
// callers keeps their hot loops (substitution sweeps, gather / update) compact and their layout independent of it.
template <int BS>
static __device__ __noinline__ void factor_cooperative(const DevSym &S, unsigned char *smem_raw, int tid, int phase) {
    if (tid >= 2 * kFactorGroup) return;
    const int g = tid / kFactorGroup, gl = tid - g * kFactorGroup;
    const int T = S.T, h = S.T / 2;
    if (phase == 0) {
        if (g == 0) {
#pragma unroll 1
            for (int t = 0; t < h; ++t) factor_block_group<BS>(S, smem_raw, t, t > 0 ? 1 : 0, 0, gl);
        } else {
#pragma unroll 1
            for (int t = T - 1; t > h; --t) factor_block_group<BS>(S, smem_raw, t, t < T - 1 ? 2 : 0, 1, gl);
        }
    } else if (g == 0) {
        factor_block_group<BS>(S, smem_raw, h, (h > 0 ? 1 : 0) | (h < T - 1 ? 2 : 0), 0, gl);
    }
}

// ---- sparse dot products of the owners ---------------------------------------------------------------------
// row i of A times a dense vector in shared memory; `range` = (first entry << 8 | length), entries are
// (position << 16 | value index).  The first four entries are loaded unconditionally (the entry array is
// padded, surplus products are masked), so the loads of several rows of one thread can be issued together;
// longer rows continue in row_rest4.
__device__ __forceinline__ double row_first4(const uint32_t *__restrict__ ent, uint32_t range, const double *__restrict__ A,
                                             const double *__restrict__ vec) {
    const int q = (int)(range >> 8), len = (int)(range & 0xFFu);
    const uint32_t e0 = ent[q], e1 = ent[q + 1], e2 = ent[q + 2], e3 = ent[q + 3];
    const double a0 = A[e0 & 0xFFFFu], a1 = A[e1 & 0xFFFFu], a2 = A[e2 & 0xFFFFu], a3 = A[e3 & 0xFFFFu];
    const double v0 = vec[e0 >> 16], v1 = vec[e1 >> 16], v2 = vec[e2 >> 16], v3 = vec[e3 >> 16];
    double acc = len > 0 ? a0 * v0 : 0.0;
    acc = len > 1 ? fma(a1, v1, acc) : acc;
    acc = len > 2 ? fma(a2, v2, acc) : acc;
    acc = len > 3 ? fma(a3, v3, acc) : acc;
    return acc;
}
__device__ __forceinline__ double row_rest4(const uint32_t *__restrict__ ent, uint32_t range, const double *__restrict__ A,
                                            const double *__restrict__ vec, double acc) {
    int q = (int)(range >> 8) + 4;
    const int q1 = (int)(range >> 8) + (int)(range & 0xFFu);
#pragma unroll 1 // (the compiler would unroll this by four at every inlined site: code size, not speed)
    for (; q < q1; q += 4) {
        const uint32_t e0 = ent[q], e1 = ent[q + 1], e2 = ent[q + 2], e3 = ent[q + 3];
        const double a0 = A[e0 & 0xFFFFu], a1 = A[e1 & 0xFFFFu], a2 = A[e2 & 0xFFFFu], a3 = A[e3 & 0xFFFFu];
        const double v0 = vec[e0 >> 16], v1 = vec[e1 >> 16], v2 = vec[e2 >> 16], v3 = vec[e3 >> 16];
        acc = fma(a0, v0, acc);
        if (q + 1 < q1) acc = fma(a1, v1, acc);
        if (q + 2 < q1) acc = fma(a2, v2, acc);
        if (q + 3 < q1) acc = fma(a3, v3, acc);
    }
    return acc;
}
__device__ __forceinline__ double row_dot4(const uint32_t *__restrict__ ent, uint32_t range, const double *__restrict__ A,
                                           const double *__restrict__ vec) {
    double acc = row_first4(ent, range, A, vec);
    if ((range & 0xFFu) > 4u) acc = row_rest4(ent, range, A, vec, acc);
    return acc;
}
// column of A (contiguous value range) times a row-indexed vector: sum_k A[k] * vec[rowOf[k]]
__device__ __forceinline__ double col_dot4(uint32_t range, const double *__restrict__ A, const uint16_t *__restrict__ rowOf,
                                           const double *__restrict__ vec, double acc) {
    int k = (int)(range >> 8);
    const int k1 = k + (int)(range & 0xFFu);
#pragma unroll 1
    for (; k < k1; k += 4) {
        const int r0 = rowOf[k], r1 = rowOf[k + 1], r2 = rowOf[k + 2], r3 = rowOf[k + 3];
        const double a0 = A[k], a1 = A[k + 1], a2 = A[k + 2], a3 = A[k + 3];
        const double v0 = vec[r0];
        acc = fma(a0, v0, acc);
        if (k + 1 < k1) acc = fma(a1, vec[r1], acc);
        if (k + 2 < k1) acc = fma(a2, vec[r2], acc);
        if (k + 3 < k1) acc = fma(a3, vec[r3], acc);
    }
    return acc;
}

// Assembly of S = P + sigma I + A' diag(rho) A into the full-row pivot blocks and the ELL
// off-diagonal blocks; destination driven (deterministic, no atomics), every thread of the CTA helps.
// sW holds rho per compact row.
static __device__ __noinline__ void assemble_S(const DevSym &S, double sigma, const double *sW, const double *sA, const double *sP,
                                               double *sDinv, double *sOff, int tid, int bs) {
    const int dpb = bs * (bs + 1) / 2, ndiag = S.T * dpb, bb = bs * bs;
    // The contribution lists live in global memory (L2): two entries per thread are walked together so that two
    // chains of dependent loads (pointer -> contribution -> operands) are in flight instead of one.  The order of
    // the additions per entry is unchanged.
    for (int e = tid; e < ndiag; e += 2 * kThreads) {
        const int e2 = e + kThreads;
        const bool two = e2 < ndiag;
        const int e2c = two ? e2 : e;
        uint32_t qa = S.sd_ptr[e], qb = S.sd_ptr[e2c];
        const uint32_t qa1 = S.sd_ptr[e + 1], qb1 = two ? S.sd_ptr[e2c + 1] : qb;
        const uint16_t pka = S.sd_pk[e], pkb = S.sd_pk[e2c];
        const int kinda = S.sd_kind[e], kindb = S.sd_kind[e2c];
        double acca = 0.0, accb = 0.0;
        while (qa < qa1 || qb < qb1) {
            const bool ha = qa < qa1, hb = qb < qb1;
            const gomp::Contrib ca = S.sd_con[ha ? qa : qa1 - 1], cb = S.sd_con[hb ? qb : (ha ? qa : qb1 - 1)];
            const double pa = sW[ca.row] * sA[ca.k1] * sA[ca.k2], pb = sW[cb.row] * sA[cb.k1] * sA[cb.k2];
            if (ha) acca += pa;
            if (hb) accb += pb;
            ++qa;
            ++qb;
        }
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            if (w == 1 && !two) break;
            const int ee = w ? e2 : e;
            double acc = w ? accb : acca;
            const uint16_t pk = w ? pkb : pka;
            const int kind = w ? kindb : kinda;
            if (pk != 0xFFFFu) acc += sP[pk];
            if (kind == 1) acc += sigma;
            else if (kind == 2) acc = 1.0;
            const int t = ee / dpb, rc = S.tri_rc[ee - t * dpb], r = rc >> 8, c = rc & 0xFF;
            sDinv[(size_t)t * bb + r * bs + c] = acc;
            sDinv[(size_t)t * bb + c * bs + r] = acc;
        }
    }
    for (int e = tid; e < S.noff; e += kThreads) {
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
}

// control word the owners publish for the solver warps before every CTA-wide rendezvous
enum { CTRL_SOLVE = 0, CTRL_FACTOR = 1, CTRL_EXIT = 2 };

// shared-memory regions: addressed as base + constant-bank offset, no pointer registers kept alive
#define sA (reinterpret_cast<double *>(smem_raw + S.oA))
#define sP (reinterpret_cast<double *>(smem_raw + S.oP))
#define sW (reinterpret_cast<double *>(smem_raw + S.oW))
#define sRHS (reinterpret_cast<double *>(smem_raw + S.oRHS))
#define sDinv (reinterpret_cast<double *>(smem_raw + S.oDinv))
#define sOff (reinterpret_cast<double *>(smem_raw + S.oOff))
#define sRed (reinterpret_cast<double *>(smem_raw + S.oRed))
#define sRowEnt (reinterpret_cast<uint32_t *>(smem_raw + S.oRowEnt))
#define sRowRange (reinterpret_cast<uint32_t *>(smem_raw + S.oRowRange))
#define sColRange (reinterpret_cast<uint32_t *>(smem_raw + S.oColRange))
#define sCscRow (reinterpret_cast<uint16_t *>(smem_raw + S.oCscRow))
#define sOffTslot (reinterpret_cast<uint16_t *>(smem_raw + S.oOffTslot))
#define sOffCol (reinterpret_cast<int8_t *>(smem_raw + S.oOffCol))
#define sOffTrow (reinterpret_cast<int8_t *>(smem_raw + S.oOffTrow))
#define sBar (reinterpret_cast<uint64_t *>(smem_raw + S.oBar))
#define sWork (reinterpret_cast<int *>(smem_raw + S.oBar + 8))
#define sCtrl (reinterpret_cast<volatile int *>(smem_raw + S.oBar + 12))

// ================================================================================================
//                       solver role (warps 0 and 1): one QP's linear algebra
// ================================================================================================
template <int BS>
static __device__ __noinline__ void solver_role(const DevSym &S, double sigma, unsigned long long *prof) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bs = BS > 0 ? BS : S.bs;
    const int T = S.T, h = S.T / 2, bb = bs * bs;
    SolverCtx C;
    C.sm = smem_u32(smem_raw);
    C.oDinv = S.oDinv; C.oOff = S.oOff; C.oRHS = S.oRHS;
    C.oV = S.oV + warp * 256; C.oScr = S.oScr + warp * bb * 8;
    C.oOffCol = S.oOffCol; C.oOffTrow = S.oOffTrow; C.oOffTslot = S.oOffTslot;
    C.bs = bs; C.T = T; C.ow = S.ow; C.owT = S.owT;
    for (;;) {
        __syncthreads(); // (E) control word + inputs published
        const int ctrl = *sCtrl;
        if (ctrl == CTRL_EXIT) break;
        if (ctrl == CTRL_FACTOR) {
            assemble_S(S, sigma, sW, sA, sP, sDinv, sOff, tid, bs);
            __syncthreads(); // (B) S assembled
            factor_cooperative<BS>(S, smem_raw, tid, 0);
            __syncthreads(); // (C) both halves eliminated
            factor_cooperative<BS>(S, smem_raw, tid, 1);
            __syncthreads(); // (D) factor complete
            continue;
        }
        // block-tridiagonal solve, in place on sRHS
        const long long t_chain0 = clock64();
        if (BS > 0 && BS <= 16 && S.uniform_off && S.ow <= 2 && S.owT <= 2) {
            // ---- fast path (uniform off-diagonal pattern): register-carried stage results, explicit LDS/STS
            const int rl = (lane & 15) < bs ? (lane & 15) : bs - 1; // two lanes per block row (lane = row + 16*half)
            OffPattern pat;
            load_off_pattern(C, 0, rl, pat);
            const uint32_t dinv0 = C.sm + (uint32_t)C.oDinv, rhs0 = C.sm + (uint32_t)C.oRHS, off0 = C.sm + (uint32_t)C.oOff,
                           vb0 = C.sm + (uint32_t)C.oV;
            const uint32_t blk_bytes = (uint32_t)(bs * C.ow) * 8u;
            // forward pattern: row rl of Off_tb (values at (rl*ow + s));  transposed: column rl (values at the stored slots)
            const int fg0 = pat.oi0 >= 0 ? pat.oi0 : 0, fg1 = pat.oi1 >= 0 ? pat.oi1 : 0;
            const uint32_t fo0 = (uint32_t)(rl * C.ow) * 8u, fo1 = fo0 + 8u;
            const bool fh0 = pat.oi0 >= 0, fh1 = pat.oi1 >= 0;
            const int tg0 = pat.ti0 >= 0 ? pat.ti0 : 0, tg1 = pat.ti1 >= 0 ? pat.ti1 : 0;
            const uint32_t to0 = (uint32_t)pat.ts0 * 8u, to1 = (uint32_t)pat.ts1 * 8u;
            const bool th0 = pat.ti0 >= 0, th1 = pat.ti1 >= 0;
            if (warp == 0) { // stages 0 .. h-1 forward (Off_{t-1} u_{t-1}), the middle stage, then back up (Off_t' x_{t+1})
                const double u = chain_sweep<BS>(dinv0, rhs0, off0, vb0, 0, h, +1, -1, T - 1, blk_bytes, true, 0.0, lane, fg0, fg1, fo0, fo1, fh0, fh1);
                (void)u;
                __syncwarp();
                if (kTimers && prof != nullptr && tid == 0) atomicAdd(&prof[10], (unsigned long long)(clock64() - t_chain0));
                named_bar_sync(1, 64);
                if (kTimers && prof != nullptr && tid == 0) atomicAdd(&prof[11], (unsigned long long)(clock64() - t_chain0));
                double yv = lane < bs ? sRHS[h * bs + lane] : 0.0;
                if (lane < bs) {
                    if (h > 0) yv -= off_mul(C, h - 1, sRHS + (h - 1) * bs, lane);
                    if (h < T - 1) yv -= offT_mul(C, h, sRHS + (h + 1) * bs, lane);
                }
                const double xm = apply_dinv<BS>(C, h, yv, lane);
                if (lane < bs) sRHS[h * bs + lane] = xm;
                named_bar_sync(2, 64);
                if (kTimers && prof != nullptr && tid == 0) atomicAdd(&prof[12], (unsigned long long)(clock64() - t_chain0));
                const double xm2 = sRHS[h * bs + rl]; // both lanes of a row hold the middle stage's value
                chain_sweep<BS>(dinv0, rhs0, off0, vb0, h - 1, h, -1, 0, T - 1, blk_bytes, false, xm2, lane, tg0, tg1, to0, to1, th0, th1);
            } else { // stages T-1 .. h+1 forward (Off_t' u_{t+1}), then back down (Off_{t-1} x_{t-1})
                chain_sweep<BS>(dinv0, rhs0, off0, vb0, T - 1, T - 1 - h, -1, 0, T - 1, blk_bytes, true, 0.0, lane, tg0, tg1, to0, to1, th0, th1);
                __syncwarp();
                named_bar_sync(1, 64);
                named_bar_sync(2, 64);
                const double xm = sRHS[h * bs + rl];
                chain_sweep<BS>(dinv0, rhs0, off0, vb0, h + 1, T - 1 - h, +1, -1, T - 1, blk_bytes, false, xm, lane, fg0, fg1, fo0, fo1, fh0, fh1);
            }
        } else
        if (warp == 0) { // stages 0 .. h-1 forward, the middle stage, then back up
            for (int t = 0; t < h; ++t) {
                double yv = lane < bs ? sRHS[t * bs + lane] : 0.0;
                if (t > 0 && lane < bs) yv -= off_mul(C, t - 1, sRHS + (t - 1) * bs, lane);
                const double u = apply_dinv<BS>(C, t, yv, lane);
                if (lane < bs) sRHS[t * bs + lane] = u;
                __syncwarp();
            }
            named_bar_sync(1, 64);
            double yv = lane < bs ? sRHS[h * bs + lane] : 0.0;
            if (lane < bs) {
                if (h > 0) yv -= off_mul(C, h - 1, sRHS + (h - 1) * bs, lane);
                if (h < T - 1) yv -= offT_mul(C, h, sRHS + (h + 1) * bs, lane);
            }
            const double xm = apply_dinv<BS>(C, h, yv, lane);
            if (lane < bs) sRHS[h * bs + lane] = xm;
            named_bar_sync(2, 64);
            for (int t = h - 1; t >= 0; --t) { // x_t = u_t - Dinv_t Off_t' x_{t+1}
                const double v = lane < bs ? offT_mul(C, t, sRHS + (t + 1) * bs, lane) : 0.0;
                const double corr = apply_dinv<BS>(C, t, v, lane);
                if (lane < bs) sRHS[t * bs + lane] -= corr;
                __syncwarp();
            }
        } else { // stages T-1 .. h+1 forward (upwards), then back down
            for (int t = T - 1; t > h; --t) {
                double yv = lane < bs ? sRHS[t * bs + lane] : 0.0;
                if (t < T - 1 && lane < bs) yv -= offT_mul(C, t, sRHS + (t + 1) * bs, lane);
                const double u = apply_dinv<BS>(C, t, yv, lane);
                if (lane < bs) sRHS[t * bs + lane] = u;
                __syncwarp();
            }
            named_bar_sync(1, 64);
            named_bar_sync(2, 64);
            for (int t = h + 1; t < T; ++t) { // x_t = u_t - Dinv_t Off_{t-1} x_{t-1}
                const double v = lane < bs ? off_mul(C, t - 1, sRHS + (t - 1) * bs, lane) : 0.0;
                const double corr = apply_dinv<BS>(C, t, v, lane);
                if (lane < bs) sRHS[t * bs + lane] -= corr;
                __syncwarp();
            }
        }
        if (kTimers && prof != nullptr && tid == 0) atomicAdd(&prof[PH_CHAIN], (unsigned long long)(clock64() - t_chain0));
        __syncthreads(); // (F) x~ ready
    }
}

template <int RPT, int CPT, int BS>
__global__ void __launch_bounds__(kThreads, 1) gomp_admm_kernel(const __grid_constant__ DevSym S, const __grid_constant__ DevBatch B,
                                                                const __grid_constant__ DevSettings st) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool owner = tid >= kSolverThreads;
    const int oid = tid - kSolverThreads;
    const int bs = BS > 0 ? BS : S.bs;
    const int T = S.T, np = S.n_pad, mu = S.m_used, h = S.T / 2, bb = bs * bs;

    // ---- once per kernel: sparsity pattern -> shared memory (identical for every QP of the batch)
    for (int k = tid; k < S.nnzA; k += kThreads) {
        sRowEnt[k] = S.row_csr[k];
        sCscRow[k] = S.csc_row[k];
    }
    for (int k = tid; k < np; k += kThreads) sColRange[k] = S.col_range[k];
    for (int k = tid; k < mu; k += kThreads) sRowRange[k] = S.row_range[k];
    if (tid < 4) sRowEnt[S.nnzA + tid] = 0u; // the 4-wide entry loops may read past the last row
    // P-side patterns are only touched by the scaling and the residual phases: shared memory if there is room
    const uint32_t *sPfEnt = S.pf_csr, *sPuIJ = S.Pu_ij;
    const uint16_t *sPfPtr = S.pf_ptr;
    if (S.p_in_smem) {
        uint32_t *e = reinterpret_cast<uint32_t *>(smem_raw + S.oPfEnt), *ij = reinterpret_cast<uint32_t *>(smem_raw + S.oPuIJ);
        uint16_t *pp = reinterpret_cast<uint16_t *>(smem_raw + S.oPfPtr);
        for (int k = tid; k < S.npf; k += kThreads) e[k] = S.pf_csr[k];
        for (int k = tid; k <= np; k += kThreads) pp[k] = S.pf_ptr[k];
        for (int k = tid; k < S.nnzPu; k += kThreads) ij[k] = S.Pu_ij[k];
        sPfEnt = e;
        sPuIJ = ij;
        sPfPtr = pp;
    }
    for (int k = tid; k < (T - 1) * bs * S.ow; k += kThreads) sOffCol[k] = S.off_col[k];
    for (int k = tid; k < (T - 1) * bs * S.owT; k += kThreads) {
        sOffTrow[k] = S.offT_row[k];
        sOffTslot[k] = S.offT_slot[k];
    }
    if (tid == kSolverThreads) mbar_init(sBar, 1);
    uint32_t bar_phase = 0;
    __syncthreads();

    for (;;) {
        if (tid == kSolverThreads) *sWork = atomicAdd(B.work_counter, 1);
        __syncthreads();
        const int widx = *sWork;
        if (widx >= B.batch) break;
        const int b = B.order != nullptr ? B.order[widx] : widx;
        if (B.active != nullptr && B.active[b] == 0) { // masked out (converged SQP problem): nothing to do
            __syncthreads();                          // everyone has read *sWork before it is overwritten
            continue;
        }

        if (!owner) {
            // SOLVER warps: wait for the owners' control word, then factor or substitute (separate function:
            // its registers are allocated independently of the owner role's long-lived state)
            solver_role<BS>(S, st.sigma, B.prof);
            fence_proxy_async();
            continue;
        }

        // =========================================================================================
        // OWNER warps: everything else.
        // =========================================================================================
        const size_t bmu = (size_t)b * mu, bnp = (size_t)b * np;
        const bool hasq = B.q != nullptr;
        long long t_phase = clock64();

        // ------------------------------------------------------------------ load
        const double *gA = B.A_val + (size_t)b * B.A_pitch;
        if (B.bulk_ok) {
            if (oid == 0) {
                mbar_expect_tx(sBar, (uint32_t)S.nnzA * 8u);
                bulk_g2s(sA, gA, (uint32_t)S.nnzA * 8u, sBar);
            }
        } else {
            for (int k = oid; k < S.nnzA; k += kOwners) sA[k] = gA[k];
        }
        for (int k = oid; k < S.nnzPu; k += kOwners) sP[k] = S.Pu_val[k];
        const double *gl = B.l + (size_t)b * S.m, *gu = B.u + (size_t)b * S.m;
        // scaled bounds: shared memory when they fit, else this QP's slice of a global scratch (L2-resident)
        double *const sLo = S.bounds_in_smem ? reinterpret_cast<double *>(smem_raw + S.oLo) : B.Lg + bmu;
        double *const sHi = S.bounds_in_smem ? reinterpret_cast<double *>(smem_raw + S.oHi) : B.Ug + bmu;
        double z[RPT], y[RPT]; // bounds live in shared memory (sLo, sHi): 2 loads per row and iteration, 4*RPT fewer registers
        double x[CPT];
        ROWS_LOOP(j, i) {
            const int orig = S.used_rows[i];
            sLo[i] = fmax(gl[orig], -kInfty); // osqp-cpp clips bounds to +-OSQP_INFTY
            sHi[i] = fmin(gu[orig], kInfty);
        }
        double bad[1] = {0.0};
        for (int e = oid; e < S.n_empty; e += kOwners) {
            const int r = S.empty_rows[e];
            if (!(gl[r] <= 0.0 && 0.0 <= gu[r])) bad[0] = 1.0;
        }
        if (B.bulk_ok) {
            unsigned spins = 0;
            while (!mbar_try_wait(sBar, bar_phase)) {
                if (++spins > (1u << 26)) __trap(); // a lost bulk copy must not hang the device
            }
            bar_phase ^= 1;
        }
        owners_max<1>(bad, sRed, oid);
        const bool empty_infeasible = bad[0] != 0.0;
        PHASE_MARK(PH_LOAD);

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
            // rn: max norms of the owned rows as they stand.  A pass traverses A twice -- the column norms, and one row-view
            // traversal that scales an entry by its row factor, then by its column factor (OSQP's order of rounding; the column
            // factors of the pass are published in sRHS for P anyway) and takes the next pass's row norm from what it writes --
            // instead of four times
            double rn[RPT];
#pragma unroll
            for (int j = 0; j < RPT; ++j) rn[j] = 0.0;
            if (st.scaling > 0) {
                ROWS_LOOP(j, i) {
                    double nrm = 0.0;
                    ROW_ENTRIES(i, e) nrm = fmax(nrm, fabs(sA[e & 0xFFFFu]));
                    rn[j] = nrm;
                }
            }
#pragma unroll 1
            for (int it = 0; it < st.scaling; ++it) {
                double dt[CPT], et[RPT];
                COLS_LOOP(c, p) {
                    double nrm = 0.0;
                    PF_ENTRIES(p, e) nrm = fmax(nrm, fabs(sP[e & 0xFFFFu]));
                    COL_ENTRIES(p, k) nrm = fmax(nrm, fabs(sA[k]));
                    dt[c] = 1.0 / sqrt(limit_scaling(nrm));
                    sRHS[p] = dt[c];
                }
                ROWS_LOOP(j, i) et[j] = 1.0 / sqrt(limit_scaling(rn[j]));
                owners_sync();
                ROWS_LOOP(j, i) {
                    double nrm = 0.0;
                    ROW_ENTRIES(i, e) {
                        const double t = (sA[e & 0xFFFFu] * et[j]) * sRHS[e >> 16];
                        sA[e & 0xFFFFu] = t;
                        nrm = fmax(nrm, fabs(t));
                    }
                    rn[j] = nrm;
                    Ei[j] *= et[j];
                }
                for (int k = oid; k < S.nnzPu; k += kOwners) {
                    const uint32_t ij = sPuIJ[k];
                    sP[k] = (sP[k] * sRHS[ij >> 16]) * sRHS[ij & 0xFFFFu];
                }
                owners_sync();
                double psum = 0.0, qn[1] = {0.0};
                COLS_LOOP(c, p) {
                    qv[c] *= dt[c];
                    Dp[c] *= dt[c];
                    double nrm = 0.0;
                    PF_ENTRIES(p, e) nrm = fmax(nrm, fabs(sP[e & 0xFFFFu]));
                    psum += nrm;
                    qn[0] = fmax(qn[0], fabs(qv[c]));
                }
                psum = owners_sum(psum, sRed, oid);
                owners_max<1>(qn, sRed, oid);
                double ct = fmax(psum / (double)S.n, limit_scaling(qn[0]));
                ct = 1.0 / limit_scaling(ct);
                for (int k = oid; k < S.nnzPu; k += kOwners) sP[k] *= ct;
                COLS_LOOP(c, p) qv[c] *= ct;
                cscale *= ct;
                owners_sync();
            }
            ROWS_LOOP(j, i) {
                sLo[i] *= Ei[j];
                sHi[i] *= Ei[j];
                B.Eg[bmu + i] = Ei[j];
            }
            COLS_LOOP(c, p) {
                B.Dg[bnp + p] = Dp[c];
                if (hasq) B.qs[bnp + p] = qv[c];
            }
        }
        const double cinv = 1.0 / cscale;
        const bool uns = !st.scaled_termination;
        PHASE_MARK(PH_SCALE);

        // ------------------------------------------------------------------ rho classification
        unsigned ctype = 0; // 2 bits per owned row: 0 inequality, 1 equality, 2 loose
        ROWS_LOOP(j, i) {
            unsigned t = 0;
            if (sLo[i] < -kInfty * kMinScaling && sHi[i] > kInfty * kMinScaling) t = 2;
            else if (sHi[i] - sLo[i] < kRhoTol) t = 1;
            ctype |= t << (2 * j);
        }

        // ------------------------------------------------------------------ iterates
        const int flags = B.flag_bits;
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
            owners_sync();
            ROWS_LOOP(j, i) {
                z[j] = row_dot4(sRowEnt, sRowRange[i], sA, sRHS);
            }
            owners_sync();
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
                // ---------------- assemble S (all threads), then the solver warps factor it
                ROWS_LOOP(j, i) {
                    const unsigned t = (ctype >> (2 * j)) & 3u;
                    sW[i] = RHO_OF(t);
                }
                if (oid == 0) *sCtrl = CTRL_FACTOR;
                __syncthreads(); // (E)
                assemble_S(S, st.sigma, sW, sA, sP, sDinv, sOff, tid, bs);
                __syncthreads(); // (B)
                PHASE_MARK(PH_ASSEMBLE);
                factor_cooperative<BS>(S, smem_raw, tid, 0);
                __syncthreads(); // (C)
                factor_cooperative<BS>(S, smem_raw, tid, 1);
                __syncthreads(); // (D)
                need_factor = false;
                w_dirty = true;
                PHASE_MARK(PH_FACTOR);
            }
            if (w_dirty) {
                ROWS_LOOP(j, i) {
                    const unsigned t = (ctype >> (2 * j)) & 3u;
                    sW[i] = fma(RHO_OF(t), z[j], -y[j]);
                }
                w_dirty = false;
                owners_sync();
            }
            // ---------------- rhs = sigma x - q + A'(rho z - y)
            COLS_LOOP(c, p) {
                double acc = st.sigma * x[c];
                if (hasq) acc -= B.qs[bnp + p];
                sRHS[p] = col_dot4(sColRange[p], sA, sCscRow, sW, acc);
            }
            if (oid == 0) *sCtrl = CTRL_SOLVE;
            PHASE_MARK(PH_RHS);
            __syncthreads(); // (E) rhs published; the solver warps substitute
            __syncthreads(); // (F) x~ in sRHS
            PHASE_MARK(PH_SOLVE);

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
            // z~ = A x~ for all owned rows first (independent loads issued together), then the updates
            uint32_t rng[RPT];
            double ztv[RPT];
#pragma unroll
            for (int j = 0; j < RPT; ++j) rng[j] = oid + j * kOwners < mu ? sRowRange[oid + j * kOwners] : 0u;
#pragma unroll
            for (int j = 0; j < RPT; ++j) ztv[j] = row_first4(sRowEnt, rng[j], sA, sRHS);
#pragma unroll
            for (int j = 0; j < RPT; ++j)
                if ((rng[j] & 0xFFu) > 4u) ztv[j] = row_rest4(sRowEnt, rng[j], sA, sRHS, ztv[j]);
            ROWS_LOOP(j, i) {
                const double zt = ztv[j];
                const unsigned t = (ctype >> (2 * j)) & 3u;
                const double zr = st.alpha * zt + (1.0 - st.alpha) * z[j];
                const double zn = fmin(fmax(zr + RINV_OF(t) * y[j], sLo[i]), sHi[i]);
                const double d = RHO_OF(t) * (zr - zn);
                y[j] += d;
                z[j] = zn;
                sW[i] = fma(RHO_OF(t), zn, -y[j]); // w = rho z - y for the next gather
                if (do_info) B.dyg[bmu + i] = d;
            }
            owners_sync(); // sRHS is overwritten by the next gather, which reads sW
            PHASE_MARK(PH_UPDATE);

            if (do_info) {
                // ---------------- residuals and the norms OSQP's tolerances need (update_info)
                long long t_info = clock64();
#define INFO_MARK(slot)                                                              \
    do {                                                                             \
        if (kTimers && B.prof != nullptr && tid == kSolverThreads) {                            \
            const long long now__ = clock64();                                       \
            atomicAdd(&B.prof[slot], (unsigned long long)(now__ - t_info));          \
            t_info = now__;                                                          \
        }                                                                            \
    } while (0)
                COLS_LOOP(c, p) sRHS[p] = x[c];
                ROWS_LOOP(j, i) sW[i] = y[j];
                w_dirty = true;
                owners_sync();
                double pri_u, pri_s, pn_u, pn_s, dua_u, dua_s, dn_u, dn_s; // residuals / normalisers, unscaled and scaled
                {
                    double na[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                    ROWS_LOOP(j, i) {
                        const double ax = row_dot4(sRowEnt, sRowRange[i], sA, sRHS);
                        const double pr = ax - z[j];
                        const double einv = 1.0 / B.Eg[bmu + i];
                        na[0] = fmax(na[0], fabs(einv * pr));
                        na[1] = fmax(na[1], fabs(pr));
                        na[2] = fmax(na[2], fabs(einv * z[j]));
                        na[3] = fmax(na[3], fabs(einv * ax));
                        na[4] = fmax(na[4], fabs(z[j]));
                        na[5] = fmax(na[5], fabs(ax));
                    }
                    owners_max<6>(na, sRed, oid);
                    pri_u = na[0]; pri_s = na[1]; pn_u = fmax(na[2], na[3]); pn_s = fmax(na[4], na[5]);
                    INFO_MARK(13);
                }
                {
                    double nb[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                    COLS_LOOP(c, p) {
                        double px = 0.0;
                        PF_ENTRIES(p, e) px = fma(sP[e & 0xFFFFu], sRHS[e >> 16], px);
                        const double aty = col_dot4(sColRange[p], sA, sCscRow, sW, 0.0);
                        const double qq = hasq ? B.qs[bnp + p] : 0.0;
                        const double du = qq + px + aty;
                        const double dinv = 1.0 / B.Dg[bnp + p];
                        nb[0] = fmax(nb[0], fabs(dinv * du));
                        nb[1] = fmax(nb[1], fabs(du));
                        nb[2] = fmax(nb[2], fabs(dinv * px));
                        nb[3] = fmax(nb[3], fabs(dinv * aty));
                        nb[4] = fmax(nb[4], fabs(dinv * qq));
                        nb[5] = fmax(nb[5], fabs(px));
                        nb[6] = fmax(nb[6], fabs(aty));
                        nb[7] = fmax(nb[7], fabs(qq));
                    }
                    owners_max<8>(nb, sRed, oid);
                    dua_u = nb[0]; dua_s = nb[1]; dn_u = fmax(nb[4], fmax(nb[3], nb[2])); dn_s = fmax(nb[7], fmax(nb[6], nb[5]));
                    INFO_MARK(14);
                }
                pri_res = uns ? pri_u : pri_s;
                dua_res = uns ? cinv * dua_u : dua_s;
                iters_done = iter;

                // ---------------- termination: stage 0 exact tolerances, stage 1 (only at max_iter) 10x tolerances
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
                    const double eps_prim = ea + er * (uns ? pn_u : pn_s);
                    if (mu == 0 || pri_res < eps_prim) {
                        prc = true;
                    } else {
                        // OSQP is_primal_infeasible on delta_y (projected onto the polar of the recession cone, kept in sW)
                        double nrm[1] = {0.0}, lhs = 0.0;
                        ROWS_LOOP(j, i) {
                            double d = B.dyg[bmu + i];
                            const double l_ = sLo[i], u_ = sHi[i];
                            if (u_ > kInfty * kMinScaling) {
                                if (l_ < -kInfty * kMinScaling) d = 0.0;
                                else d = fmin(d, 0.0);
                            } else if (l_ < -kInfty * kMinScaling) {
                                d = fmax(d, 0.0);
                            }
                            sW[i] = d;
                            nrm[0] = fmax(nrm[0], fabs(uns ? B.Eg[bmu + i] * d : d));
                            lhs += u_ * fmax(d, 0.0) + l_ * fmin(d, 0.0);
                        }
                        owners_max<1>(nrm, sRed, oid); // (also orders the sW writes before the gather below)
                        lhs = owners_sum(lhs, sRed, oid);
                        if (nrm[0] > kDivTol && lhs < -epi * nrm[0]) {
                            double v[1] = {0.0};
                            COLS_LOOP(c, p) {
                                const double acc = col_dot4(sColRange[p], sA, sCscRow, sW, 0.0);
                                v[0] = fmax(v[0], fabs(uns ? (1.0 / B.Dg[bnp + p]) * acc : acc));
                            }
                            owners_max<1>(v, sRed, oid);
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
                        COLS_LOOP(c, p) {
                            const double d = B.dxg[bnp + p];
                            nrm[0] = fmax(nrm[0], fabs(uns ? B.Dg[bnp + p] * d : d));
                            qdx += B.qs[bnp + p] * d;
                        }
                        owners_max<1>(nrm, sRed, oid);
                        qdx = owners_sum(qdx, sRed, oid);
                        const double cs_ = uns ? cscale : 1.0;
                        if (nrm[0] > kDivTol && qdx < -cs_ * edi * nrm[0]) {
                            COLS_LOOP(c, p) sRHS[p] = B.dxg[bnp + p];
                            owners_sync();
                            double v[1] = {0.0};
                            COLS_LOOP(c, p) {
                                double acc = 0.0;
                                PF_ENTRIES(p, e) acc = fma(sP[e & 0xFFFFu], sRHS[e >> 16], acc);
                                v[0] = fmax(v[0], fabs(uns ? (1.0 / B.Dg[bnp + p]) * acc : acc));
                            }
                            owners_max<1>(v, sRed, oid);
                            if (v[0] < cs_ * edi * nrm[0]) {
                                double viol[1] = {0.0};
                                ROWS_LOOP(j, i) {
                                    double acc = row_dot4(sRowEnt, sRowRange[i], sA, sRHS);
                                    if (uns) acc *= 1.0 / B.Eg[bmu + i];
                                    if ((sHi[i] < kInfty * kMinScaling && acc > edi * nrm[0]) ||
                                        (sLo[i] > -kInfty * kMinScaling && acc < -edi * nrm[0]))
                                        viol[0] = 1.0;
                                }
                                owners_max<1>(viol, sRed, oid);
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
                INFO_MARK(15);
#undef INFO_MARK
                PHASE_MARK(PH_INFO);
            }
        }
#undef RHO_OF
#undef RINV_OF
        if (oid == 0) *sCtrl = CTRL_EXIT;
        __syncthreads(); // (E) releases the solver warps from this QP

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
        for (int e = oid; e < S.n_empty; e += kOwners) yo[S.empty_rows[e]] = infeasible ? qnan : 0.0;
        if (oid == 0) {
            B.rho_state[b] = rho;
            B.status[b] = status;
            B.iters[b] = iters_done;
            B.rho_updates[b] = rho_updates;
            B.pri_res[b] = pri_res;
            B.dua_res[b] = dua_res;
        }
        fence_proxy_async(); // generic-proxy accesses to sA are ordered before the next QP's bulk copy
        PHASE_MARK(PH_STORE);
    }
}

#undef sA
#undef sP
#undef sW
#undef sRHS
#undef sDinv
#undef sOff
#undef sRed
#undef sRowRange
#undef sRowEnt
#undef sColRange
#undef sCscRow
#undef sOffTslot
#undef sOffCol
#undef sOffTrow
#undef sBar
#undef sWork
#undef sCtrl
#undef ROWS_LOOP
#undef COLS_LOOP
#undef ROW_ENTRIES
#undef COL_ENTRIES
#undef PF_ENTRIES

// ---- launch interface (one explicit instantiation per translation unit) ------------------------------
template <int RPT, int CPT, int BS>
cudaError_t prepare_variant(size_t smem_bytes, int *ctas_per_sm) {
    auto fn = gomp_admm_kernel<RPT, CPT, BS>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, fn, kThreads, smem_bytes);
}
template <int RPT, int CPT, int BS>
cudaError_t launch_variant(const DevSym &S, const DevBatch &B, const DevSettings &st, int grid, size_t smem_bytes,
                           cudaStream_t stream) {
    gomp_admm_kernel<RPT, CPT, BS><<<grid, kThreads, smem_bytes, stream>>>(S, B, st);
    return cudaGetLastError();
}

} // namespace gomp_dev
#endif // GOMP_KERNEL_CUH
