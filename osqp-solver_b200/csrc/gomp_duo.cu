// gomp_duo.cu -- the two-slot ADMM kernel of libgomp_cuda (GOMP_LINSYS_BLOCKTRI_TMEM; sm_100a, fp64).
//
// Same algorithm as gomp_kernel.cuh (OSQP 0.6 restated: Ruiz, rho classification, reduced-KKT ADMM with a
// two-way block-tridiagonal factorisation and explicit pivot-block inverses, OSQP's termination / infeasibility
// tests, adaptive rho), re-mapped so that TWO QPs are resident per SM and every phase issues far fewer
// instructions per QP:
//   * one 512-thread CTA per SM = two independent SLOTS.  A slot is 2 substitution warps + 6 owner warps and runs
//     one QP at a time from the batch's work queue; the slots share nothing but the sparsity-pattern arrays in
//     shared memory and synchronise only through their own named barriers.
//   * the factor (25 + 1 + 24 inverted 14 x 14 pivot blocks per QP, 78 KB) lives in TENSOR MEMORY: every lane of a
//     substitution warp keeps its half rows of all its stages in its own TMEM lane (tcgen05.st after the
//     factorisation, one tcgen05.ld.32x32b.x16 per stage, issued a stage ahead).  The tensor cores themselves are
//     not used (nothing here is a dense contraction beyond 14 x 14 in fp64); TMEM is the 256 KB of lane-private
//     storage that lets two factors stay on chip next to two QPs' worth of shared memory.
//   * the block factorisation runs inside the two substitution warps (in-warp symmetric Gauss-Jordan sweeps, pivot
//     column exchanged through a 128-byte shared line) while the owner warps assemble the next stage's block of
//     S = P + sigma I + A' diag(rho) A one round ahead (named-barrier producer / consumer pair).
//   * sparse phases walk "ELL by slot" views (symbolic.h: build_owner_ell): uniform trip counts for the whole CTA,
//     one conflict-free load per entry, no per-row ranges or divergent tails.
// Reference call sites replaced: src/osqp-wrapper.h:28 (Init), :46 (SetPrimalWarmStart), :52 (Solve), :36,:40
// (UpdateConstraintMatrix, SetBounds) -- as for the other back ends.
#include "gomp_duo.h"

#include <type_traits>

namespace gomp_dev {

// ---- tensor memory as lane-private storage ---------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t ta, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(ta)
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t ta, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(ta),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                 "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ double mk_f64(uint32_t lo, uint32_t hi) { return __hiloint2double((int)hi, (int)lo); }
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void sts_f64x2(uint32_t a, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}

// named barriers of slot s: base 1 + 5 s
enum { DB_SLOT = 0, DB_OWNERS = 1, DB_PAIR = 2, DB_RING_LOADED = 3, DB_RING_WRITTEN = 4 };
__device__ __forceinline__ int duo_bar(int slot, int which) { return 1 + 5 * slot + which; }

enum { DUO_SOLVE = 0, DUO_FACTOR = 1, DUO_EXIT = 2 };

// ---- reductions over the owner warps of one slot (fixed order) ----------------------------------------------
template <int K>
__device__ __forceinline__ void duo_max(double (&v)[K], double *sRed, int oid, int bar) {
    constexpr int NW = kDuoOwners / 32;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double a = v[k];
#pragma unroll
        for (int o = 16; o; o >>= 1) a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o));
        if ((oid & 31) == 0) sRed[k * NW + (oid >> 5)] = a;
    }
    named_bar_sync(bar, kDuoOwners);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double a = sRed[k * NW];
#pragma unroll
        for (int j = 1; j < NW; ++j) a = fmax(a, sRed[k * NW + j]);
        v[k] = a;
    }
    named_bar_sync(bar, kDuoOwners);
}
__device__ __forceinline__ double duo_sum(double a, double *sRed, int oid, int bar) {
    constexpr int NW = kDuoOwners / 32;
#pragma unroll
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((oid & 31) == 0) sRed[oid >> 5] = a;
    named_bar_sync(bar, kDuoOwners);
    double r = sRed[0];
#pragma unroll
    for (int j = 1; j < NW; ++j) r += sRed[j];
    named_bar_sync(bar, kDuoOwners);
    return r;
}

// ================================================================================================
//                substitution warps: block factorisation and the sweeps of every iteration
// ================================================================================================
// Lane layout of a substitution warp (as chain_sweep in gomp_kernel.cuh): lane = r + 16 * half; half 0 holds columns
// 0..7 of block row r, half 1 columns 8..NB-1 (padded with zeros to 8).  Per stage a lane owns 16 TMEM columns.
struct ChainPat {              // per-lane constants of the uniform off-diagonal pattern (factorisation)
    unsigned long long cidx;   // for the lane's 8 columns j: nibble (2c + s) = s-th source index of row / column j (15: none)
    uint32_t cslot[4];         // ... byte (2c + s) = its value slot inside a block
};

// One sweep over `count` stages.  tq: TMEM address of the warp's lane quarter; col: first stage's column, dcol = +-16;
// ab / dab: shared address of the lane's right-hand-side entry of the first stage and its step; tb / dtb: off-diagonal
// block of the first stage and its step (blocks outside [0, nblk) contribute nothing).
//   fwd:  rhs_t <- Dinv_t (rhs_t - O . prev)      bwd:  rhs_t <- rhs_t - Dinv_t (O . prev)
// The stage's half row (16 TMEM columns) is requested as soon as the previous stage's products have consumed the
// registers: the load completes under the dependent tail of that stage and the head of this one (single buffer: the
// kernel runs at 96 registers per thread, and every register this loop saves keeps its four broadcast loads in flight
// together).
template <int NB, bool FWD>
static __device__ __forceinline__ double duo_sweep(uint32_t tq, int col, int dcol, uint32_t ab, int dab, uint32_t off0, uint32_t blk_bytes, int tb,
                                                int dtb, int nblk, uint32_t vb0, int count, double prev, int lane, int g0, int g1, uint32_t o0,
                                                uint32_t o1, bool has0, bool has1) {
    const int r = lane & 15, half = lane >> 4;
    const bool act = half == 0 && r < NB;
    const uint32_t avs = vb0 + (uint32_t)r * 8u, avl = vb0 + (half ? 64u : 0u);
    uint32_t c[16];
    auto clampc = [](int cc) { return (uint32_t)(cc < 0 ? 0 : (cc > kDuoTmemCols - 16 ? kDuoTmemCols - 16 : cc)); };
    tmem_ld16(tq + clampc(col), c);
    uint32_t par = 0;
#pragma unroll 1
    for (int it = 0; it < count; ++it) {
        // ---- independent of the previous stage
        const bool blk = (unsigned)tb < (unsigned)nblk;
        const uint32_t ao = off0 + (uint32_t)tb * blk_bytes;
        const double b = lds_f64(ab);
        const double ov0 = (blk && has0) ? lds_f64(ao + o0) : 0.0;
        const double ov1 = (blk && has1) ? lds_f64(ao + o1) : 0.0;
        // ---- dependent chain
        const double up0 = __shfl_sync(0xffffffffu, prev, g0);
        const double up1 = __shfl_sync(0xffffffffu, prev, g1);
        const double off = fma(ov0, up0, ov1 * up1);
        const double v = FWD ? b - off : off;
        if (act) sts_f64(avs + par, v);
        __syncwarp();
        const uint32_t aw = avl + par;
        const double2 w0 = lds_f64x2(aw), w1 = lds_f64x2(aw + 16), w2 = lds_f64x2(aw + 32), w3 = lds_f64x2(aw + 48);
        tmem_wait_ld();
        double a0 = mk_f64(c[0], c[1]) * w0.x, a1 = mk_f64(c[2], c[3]) * w0.y;
        a0 = fma(mk_f64(c[4], c[5]), w1.x, a0);
        a1 = fma(mk_f64(c[6], c[7]), w1.y, a1);
        a0 = fma(mk_f64(c[8], c[9]), w2.x, a0);
        a1 = fma(mk_f64(c[10], c[11]), w2.y, a1);
        a0 = fma(mk_f64(c[12], c[13]), w3.x, a0);
        a1 = fma(mk_f64(c[14], c[15]), w3.y, a1);
        col += dcol;
        tmem_ld16(tq + clampc(col), c); // next stage's half row
        double acc = a0 + a1;
        acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        prev = FWD ? acc : b - acc;
        if (act) sts_f64(ab, prev);
        ab += (uint32_t)dab;
        tb += dtb;
        par ^= 128u;
    }
    tmem_wait_ld();
    return prev;
}

// sum_{s1,s2} O[r][s1] D[a_s1][b_s2] O[j][s2] for the lane's 8 columns j (D: previous inverse in shared memory)
template <int NB>
__device__ __forceinline__ void duo_schur(double (&acc)[8], uint32_t dprev, uint32_t oblk, int ra0, int ra1, uint32_t ro0, uint32_t ro1,
                                          bool rh0, bool rh1, const ChainPat &P) {
    const double va0 = rh0 ? lds_f64(oblk + ro0) : 0.0, va1 = rh1 ? lds_f64(oblk + ro1) : 0.0;
    const uint32_t row0 = dprev + (uint32_t)(ra0 * NB) * 8u, row1 = dprev + (uint32_t)(ra1 * NB) * 8u;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        double a = 0.0;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int e = 2 * c + s;
            const uint32_t b = (uint32_t)(P.cidx >> (4 * e)) & 15u;
            const uint32_t slot = (P.cslot[e >> 2] >> (8 * (e & 3))) & 0xFFu;
            const bool valid = b != 15u;
            const double vb = valid ? lds_f64(oblk + slot * 8u) : 0.0;
            const uint32_t bb = valid ? b : 0u;
            const double d0 = lds_f64(row0 + bb * 8u), d1 = lds_f64(row1 + bb * 8u);
            a = fma(fma(va0, d0, va1 * d1), vb, a);
        }
        acc[c] = a;
    }
}

// in-warp inversion of an SPD NB x NB block held as half rows (symmetric sweep operator: the matrix stays symmetric,
// so the pivot row is the pivot column; after all sweeps it holds -inverse).  colb: this warp's [2][16] exchange line.
template <int NB>
__device__ __forceinline__ void duo_invert(double (&m)[8], uint32_t colb, int r, int rl, int half) {
    const uint32_t c0b = half ? 64u : 0u;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        const int kh = k >= 8 ? 1 : 0, kc = k - 8 * kh;
        const uint32_t cb = colb + (uint32_t)(k & 1) * 128u;
        if (half == kh && r < NB) sts_f64(cb + (uint32_t)r * 8u, m[kc]);
        __syncwarp();
        const double p = lds_f64(cb + (uint32_t)k * 8u);
        const double ark = lds_f64(cb + (uint32_t)rl * 8u);
        const double2 q0 = lds_f64x2(cb + c0b), q1 = lds_f64x2(cb + c0b + 16), q2 = lds_f64x2(cb + c0b + 32), q3 = lds_f64x2(cb + c0b + 48);
        const double akj[8] = {q0.x, q0.y, q1.x, q1.y, q2.x, q2.y, q3.x, q3.y};
        const double pinv = 1.0 / p;
        const double f = ark * pinv;
        const bool rk = rl == k;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const double gen = fma(-f, akj[c], m[c]);
            const double rowv = m[c] * pinv;
            double v = rk ? rowv : gen;
            if (c == kc) v = (half == kh) ? (rk ? -pinv : f) : v;
            m[c] = v;
        }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) m[c] = -m[c];
}

// The block factorisation of one QP, as seen by one substitution warp (cw 0: down chain, 1: up chain).  Out of line:
// its per-lane constants (pattern nibbles, packed-triangle offsets) and temporaries are allocated on their own and are
// not live across the substitution sweeps of the caller.  Round R: the owners have staged stage R of the down chain and
// stage T-1-R of the up chain in the ring (round h: the middle stage); the warp subtracts the Schur term of its previous
// stage, inverts the block in registers and stores its half rows to its own TMEM lane.
template <int NB>
static __device__ __noinline__ void duo_factor(int T, int ow, int owT, const int8_t *__restrict__ off_col, const int8_t *__restrict__ offT_row,
                                               const uint16_t *__restrict__ offT_slot, uint32_t off0, uint32_t scr, uint32_t ring, uint32_t tq,
                                               int slot, int cw, int lane) {
    const int h = T / 2;
    const int r = lane & 15, half = lane >> 4, rl = r < NB ? r : NB - 1;
    const int c0 = half ? 8 : 0;
    const uint32_t dprev_own = scr + (uint32_t)(cw * NB * NB) * 8u, dprev_other = scr + (uint32_t)((cw ^ 1) * NB * NB) * 8u;
    const uint32_t colb = scr + (uint32_t)(2 * NB * NB) * 8u + (uint32_t)cw * 256u;
    const int bPair = duo_bar(slot, DB_PAIR), bLoaded = duo_bar(slot, DB_RING_LOADED), bWritten = duo_bar(slot, DB_RING_WRITTEN);
    const uint32_t blk_bytes = (uint32_t)(NB * ow) * 8u;
    constexpr int dpb = NB * (NB + 1) / 2;
    const int nstages = cw == 0 ? h : T - 1 - h;
    // pattern of the lane's own row (forward view for the down chain, transposed for the up chain) and of its 8 columns
    ChainPat P;
    int ra0, ra1;
    uint32_t ro0, ro1;
    bool rh0, rh1;
    {
        int a0, a1;
        if (cw == 0) {
            a0 = off_col[rl * ow];
            a1 = ow > 1 ? off_col[rl * ow + 1] : -1;
            ro0 = (uint32_t)(rl * ow) * 8u;
            ro1 = ro0 + 8u;
        } else {
            a0 = offT_row[rl * owT];
            a1 = owT > 1 ? offT_row[rl * owT + 1] : -1;
            ro0 = (uint32_t)offT_slot[rl * owT] * 8u;
            ro1 = owT > 1 ? (uint32_t)offT_slot[rl * owT + 1] * 8u : 0u;
        }
        ra0 = a0 >= 0 ? a0 : 0; ra1 = a1 >= 0 ? a1 : 0;
        rh0 = a0 >= 0; rh1 = a1 >= 0;
        P.cidx = 0ull;
        P.cslot[0] = P.cslot[1] = P.cslot[2] = P.cslot[3] = 0u;
        for (int c = 0; c < 8; ++c)
            for (int s = 0; s < 2; ++s) {
                const int j = c0 + c, e = 2 * c + s;
                int b = -1, sl = 0;
                if (j < NB) {
                    if (cw == 0) { // down chain: row j of Off (forward view)
                        if (s < ow) { b = off_col[j * ow + s]; sl = j * ow + s; }
                    } else {       // up chain: column j of Off (transposed view)
                        if (s < owT) { b = offT_row[j * owT + s]; sl = offT_slot[j * owT + s]; }
                    }
                }
                P.cidx |= (unsigned long long)(b >= 0 ? b : 15) << (4 * e);
                P.cslot[e >> 2] |= (uint32_t)(b >= 0 ? sl : 0) << (8 * (e & 3));
            }
    }
    uint32_t tri[8]; // packed lower-triangle offsets of the lane's 8 entries of a block row
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int j = c0 + c < NB ? c0 + c : NB - 1;
        tri[c] = (uint32_t)(rl >= j ? rl * (rl + 1) / 2 + j : j * (j + 1) / 2 + rl) * 8u;
    }
#pragma unroll 1
    for (int R = 0; R <= h; ++R) {
        named_bar_sync(bWritten, kDuoSlotThreads); // the owners have staged round R's blocks
        const bool mine = R < nstages;
        const bool mid = R == h;
        const int t = mid ? h : (cw == 0 ? R : T - 1 - R);
        double m[8];
        if (mine || (mid && cw == 0)) {
            const uint32_t src = ring + (uint32_t)(mid ? 0 : cw * (dpb + 1)) * 8u;
#pragma unroll
            for (int c = 0; c < 8; ++c) m[c] = c0 + c < NB ? lds_f64(src + tri[c]) : 0.0;
        } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) m[c] = 0.0;
        }
        double sch[8];
        const bool has_prev = R > 0 && (mine || mid);
        if (has_prev) {
            const int tb = cw == 0 ? t - 1 : t;
            duo_schur<NB>(sch, dprev_own, off0 + (uint32_t)tb * blk_bytes, ra0, ra1, ro0, ro1, rh0, rh1, P);
#pragma unroll
            for (int c = 0; c < 8; ++c) m[c] -= sch[c];
        }
        if (R < h) {
            // (the loads above are consumed before the arrive below, whatever the barrier's ordering of pending reads)
            if (((m[0] + m[1]) + (m[2] + m[3])) + ((m[4] + m[5]) + (m[6] + m[7])) == -1.2345e300) __trap();
            named_bar_arrive(bLoaded, kDuoSlotThreads); // the ring may be overwritten with round R + 1
        }
        if (mid) {
            // the up chain hands its Schur term over through its previous-inverse buffer, the down chain finishes
            if (cw == 1) {
                __syncwarp();
                if (r < NB) {
#pragma unroll
                    for (int c = 0; c < 8; c += 2)
                        if (c0 + c < NB) sts_f64x2(dprev_own + (uint32_t)(r * NB + c0 + c) * 8u, sch[c], sch[c + 1]);
                }
            }
            named_bar_sync(bPair, kDuoChainThreads);
            if (cw == 0) {
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (c0 + c < NB) m[c] -= lds_f64(dprev_other + (uint32_t)(rl * NB + c0 + c) * 8u);
            }
        }
        if (mine || (mid && cw == 0)) {
            duo_invert<NB>(m, colb, r, rl, half);
            uint32_t regs[16];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                regs[2 * c] = (uint32_t)__double2loint(m[c]);
                regs[2 * c + 1] = (uint32_t)__double2hiint(m[c]);
            }
            tmem_st16(tq + (uint32_t)(R * 16), regs);
            if (!mid) {
                __syncwarp(); // everybody is done reading the previous inverse
                if (r < NB) {
#pragma unroll
                    for (int c = 0; c < 8; c += 2)
                        if (c0 + c < NB) sts_f64x2(dprev_own + (uint32_t)(r * NB + c0 + c) * 8u, m[c], m[c + 1]);
                }
                __syncwarp();
            }
        }
    }
    tmem_wait_st();
}

template <int NB>
static __device__ __forceinline__ void duo_chain_role(const DevSym &S, const DevDuo &X, unsigned long long *prof, uint32_t tmem, int slot, int cw, int lane) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = S.T, h = S.T / 2;
    const uint32_t sb = smem_u32(smem_raw) + (uint32_t)(X.slot_base + slot * X.slot_bytes);
    const uint32_t rhs0 = sb + (uint32_t)X.oRHS, off0 = sb + (uint32_t)X.oOff, vb0 = sb + (uint32_t)X.oV + (uint32_t)cw * 256u;
    const uint32_t tq = tmem + ((uint32_t)((threadIdx.x >> 5) & 3) << 21); // lane quarter of this warp: (32 * quarter) << 16
    const int r = lane & 15, half = lane >> 4, rl = r < NB ? r : NB - 1;
    const bool act = half == 0 && r < NB;
    const int bSlot = duo_bar(slot, DB_SLOT), bPair = duo_bar(slot, DB_PAIR);
    const uint32_t blk_bytes = (uint32_t)(NB * S.ow) * 8u;
    const int nstages = cw == 0 ? h : T - 1 - h; // stages of this chain

    // ---- per-lane constants of the sweeps (uniform off-diagonal pattern: block 0 stands for all)
    int fg0, fg1, tg0, tg1;
    uint32_t fo0, fo1, to0, to1;
    bool fh0, fh1, th0, th1;
    {
        const int o0 = S.off_col[rl * S.ow], o1 = S.ow > 1 ? S.off_col[rl * S.ow + 1] : -1;
        fg0 = o0 >= 0 ? o0 : 0; fg1 = o1 >= 0 ? o1 : 0;
        fo0 = (uint32_t)(rl * S.ow) * 8u; fo1 = fo0 + 8u;
        fh0 = o0 >= 0; fh1 = o1 >= 0;
        const int t0 = S.offT_row[rl * S.owT], t1 = S.owT > 1 ? S.offT_row[rl * S.owT + 1] : -1;
        tg0 = t0 >= 0 ? t0 : 0; tg1 = t1 >= 0 ? t1 : 0;
        to0 = (uint32_t)S.offT_slot[rl * S.owT] * 8u;
        to1 = S.owT > 1 ? (uint32_t)S.offT_slot[rl * S.owT + 1] * 8u : 0u;
        th0 = t0 >= 0; th1 = t1 >= 0;
    }

    for (;;) {
        named_bar_sync(bSlot, kDuoSlotThreads); // (E) control word + inputs published
        const int ctrl = *reinterpret_cast<volatile int *>(smem_raw + X.slot_base + slot * X.slot_bytes + X.oCtl);
        if (ctrl == DUO_EXIT) break;
        if (ctrl == DUO_FACTOR) {
            duo_factor<NB>(T, S.ow, S.owT, S.off_col, S.offT_row, S.offT_slot, off0, sb + (uint32_t)X.oScr, sb + (uint32_t)X.oRing, tq, slot, cw, lane);
            named_bar_sync(bSlot, kDuoSlotThreads); // (D) factor complete
            continue;
        }
        // ---------------------------------------------------------------- block-tridiagonal solve, in place on the rhs
        const long long t_chain0 = kTimers ? clock64() : 0;
        if (cw == 0) { // stages 0 .. h-1 forward (Off_{t-1} u_{t-1}), the middle stage, then back up (Off_t' x_{t+1})
            const double u = duo_sweep<NB, true>(tq, 0, 16, rhs0 + (uint32_t)rl * 8u, NB * 8, off0, blk_bytes, -1, +1, T - 1, vb0, h, 0.0, lane,
                                                 fg0, fg1, fo0, fo1, fh0, fh1);
            uint32_t cm[16];
            tmem_ld16(tq + (uint32_t)(h * 16), cm);
            if (kTimers && prof != nullptr && lane == 0) atomicAdd(&prof[10], (unsigned long long)(clock64() - t_chain0));
            named_bar_sync(bPair, kDuoChainThreads); // the up chain has reached stage h + 1
            if (kTimers && prof != nullptr && lane == 0) atomicAdd(&prof[11], (unsigned long long)(clock64() - t_chain0));
            // v = rhs_h - Off_{h-1} u_{h-1} - Off_h' u_{h+1}
            const uint32_t abm = rhs0 + (uint32_t)(h * NB + rl) * 8u;
            const double b = lds_f64(abm);
            const uint32_t aoA = off0 + (uint32_t)(h - 1) * blk_bytes, aoB = off0 + (uint32_t)h * blk_bytes;
            const double fv0 = fh0 ? lds_f64(aoA + fo0) : 0.0, fv1 = fh1 ? lds_f64(aoA + fo1) : 0.0;
            const bool hasB = h < T - 1;
            const double tv0 = (hasB && th0) ? lds_f64(aoB + to0) : 0.0, tv1 = (hasB && th1) ? lds_f64(aoB + to1) : 0.0;
            const uint32_t nxt = rhs0 + (uint32_t)((hasB ? h + 1 : h) * NB) * 8u;
            const double un0 = lds_f64(nxt + (uint32_t)tg0 * 8u), un1 = lds_f64(nxt + (uint32_t)tg1 * 8u);
            const double up0 = __shfl_sync(0xffffffffu, u, fg0), up1 = __shfl_sync(0xffffffffu, u, fg1);
            const double v = (b - fma(fv0, up0, fv1 * up1)) - fma(tv0, un0, tv1 * un1);
            if (act) sts_f64(vb0 + (uint32_t)r * 8u, v);
            tmem_wait_ld();
            __syncwarp();
            const uint32_t aw = vb0 + (half ? 64u : 0u);
            const double2 w0 = lds_f64x2(aw), w1 = lds_f64x2(aw + 16), w2 = lds_f64x2(aw + 32), w3 = lds_f64x2(aw + 48);
            double a0 = mk_f64(cm[0], cm[1]) * w0.x, a1 = mk_f64(cm[2], cm[3]) * w0.y;
            a0 = fma(mk_f64(cm[4], cm[5]), w1.x, a0);
            a1 = fma(mk_f64(cm[6], cm[7]), w1.y, a1);
            a0 = fma(mk_f64(cm[8], cm[9]), w2.x, a0);
            a1 = fma(mk_f64(cm[10], cm[11]), w2.y, a1);
            a0 = fma(mk_f64(cm[12], cm[13]), w3.x, a0);
            a1 = fma(mk_f64(cm[14], cm[15]), w3.y, a1);
            double xm = a0 + a1;
            xm += __shfl_xor_sync(0xffffffffu, xm, 16);
            if (act) sts_f64(abm, xm);
            __syncwarp();
            named_bar_sync(bPair, kDuoChainThreads); // x_h published
            if (kTimers && prof != nullptr && lane == 0) atomicAdd(&prof[12], (unsigned long long)(clock64() - t_chain0));
            duo_sweep<NB, false>(tq, (h - 1) * 16, -16, rhs0 + (uint32_t)((h - 1) * NB + rl) * 8u, -NB * 8, off0, blk_bytes, h - 1, -1, T - 1, vb0,
                                 h, xm, lane, tg0, tg1, to0, to1, th0, th1);
        } else { // stages T-1 .. h+1 forward (Off_t' u_{t+1}), then back down (Off_{t-1} x_{t-1})
            duo_sweep<NB, true>(tq, 0, 16, rhs0 + (uint32_t)((T - 1) * NB + rl) * 8u, -NB * 8, off0, blk_bytes, T - 1, -1, T - 1, vb0, nstages,
                                0.0, lane, tg0, tg1, to0, to1, th0, th1);
            __syncwarp();
            named_bar_sync(bPair, kDuoChainThreads);
            named_bar_sync(bPair, kDuoChainThreads);
            const double xm = lds_f64(rhs0 + (uint32_t)(h * NB + rl) * 8u);
            duo_sweep<NB, false>(tq, (nstages - 1) * 16, -16, rhs0 + (uint32_t)((h + 1) * NB + rl) * 8u, NB * 8, off0, blk_bytes, h, +1, T - 1, vb0,
                                 nstages, xm, lane, fg0, fg1, fo0, fo1, fh0, fh1);
        }
        if (kTimers && prof != nullptr && cw == 0 && lane == 0) atomicAdd(&prof[PH_CHAIN], (unsigned long long)(clock64() - t_chain0));
        named_bar_sync(bSlot, kDuoSlotThreads); // (F) x~ ready
    }
}

// ================================================================================================
//                                   owner warps: everything else
// ================================================================================================
// Register economy rules this role (96 registers per thread with 18 warps per CTA): every shared-memory access is an
// explicit ld/st.shared on a 32-bit address formed from ONE base register plus kernel-parameter offsets; ELL entries carry
// premultiplied byte offsets; rho and 1/rho per row type come from a 3-entry shared table; scalars that are only needed at
// the residual checks (cost scale, rho) live in shared memory.
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// Walk the ELL lines of the N row (ROWS) or column slots of this thread "line-major", four slots at a time: line q of
// the four slots is loaded first (independent loads, issued together), then handed to f(slot0, entries).  A slot shorter
// than its group yields the surplus entry `pad` (zero value, zero operand).  The per-thread dependent chain of a traversal
// is max-length-of-group steps instead of the sum of all lengths.
template <bool ROWS, int N, int G0 = 0, class F>
__device__ __forceinline__ void duo_ell_lines(const DevDuo &X, uint32_t mine, uint32_t pad, F &&f) {
    if constexpr (G0 < N) {
        // Groups of 2, 3, then 4 slots: the slots hold rows / columns longest first, so the lengths differ most among the first
        // ones (GOMP, N = 7: rows 7 7 3 2 2 1 1 1 1, columns 7 7 4 4) -- a group pads every slot to its longest, and groups of
        // four would walk 37 row entries and 28 column entries per thread where 27 and 22 carry values.
        constexpr int GW = G0 == 0 ? 2 : (G0 == 2 ? 3 : 4);
        constexpr int G = (N - G0) < GW ? (N - G0) : GW;
        int maxlen = 0;
#pragma unroll
        for (int g = 0; g < G; ++g) maxlen = max(maxlen, ROWS ? X.rlen[G0 + g] : X.clen[G0 + g]);
#pragma unroll 1
        for (int q = 0; q < maxlen; ++q) {
            uint32_t e[G];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int len = ROWS ? X.rlen[G0 + g] : X.clen[G0 + g];
                const int base = ROWS ? X.rbase[G0 + g] : X.cbase[G0 + g];
                e[g] = q < len ? lds_u32(mine + (uint32_t)base + (uint32_t)q * (kDuoOwners * 4u)) : pad;
            }
            f(std::integral_constant<int, G0>{}, e);
        }
        duo_ell_lines<ROWS, N, G0 + G>(X, mine, pad, f);
    }
}
// acc[slot] += sum over the slot's entries of value * vec[operand]   (values at aVal, operand vector at aVec)
#define DUO_DOT(ROWSEL, N, accarr, aVal, aVec, mine, pad)                                           \
    duo_ell_lines<ROWSEL, N>(X, mine, pad, [&](auto g0c, const auto &e) {                            \
        constexpr int g0 = decltype(g0c)::value, G = (int)(sizeof(e) / sizeof(e[0]));                \
        double a__[G], v__[G];                                                                       \
        _Pragma("unroll") for (int g = 0; g < G; ++g) {                                              \
            a__[g] = lds_f64((aVal) + LO16(e[g]));                                                   \
            v__[g] = lds_f64((aVec) + HI16(e[g]));                                                   \
        }                                                                                            \
        _Pragma("unroll") for (int g = 0; g < G; ++g) accarr[g0 + g] = fma(a__[g], v__[g], accarr[g0 + g]); \
    })
// nrm[slot] = max(nrm[slot], |value|) over the slot's entries
#define DUO_ABSMAX(ROWSEL, N, nrmarr, aVal, mine, pad)                                              \
    duo_ell_lines<ROWSEL, N>(X, mine, pad, [&](auto g0c, const auto &e) {                            \
        constexpr int g0 = decltype(g0c)::value, G = (int)(sizeof(e) / sizeof(e[0]));                \
        double a__[G];                                                                               \
        _Pragma("unroll") for (int g = 0; g < G; ++g) a__[g] = lds_f64((aVal) + LO16(e[g]));         \
        _Pragma("unroll") for (int g = 0; g < G; ++g) nrmarr[g0 + g] = fmax(nrmarr[g0 + g], fabs(a__[g])); \
    })
// value *= fac[slot] for the slot's entries
#define DUO_SCALE(ROWSEL, N, facarr, aVal, mine, pad)                                               \
    duo_ell_lines<ROWSEL, N>(X, mine, pad, [&](auto g0c, const auto &e) {                            \
        constexpr int g0 = decltype(g0c)::value, G = (int)(sizeof(e) / sizeof(e[0]));                \
        double a__[G];                                                                               \
        _Pragma("unroll") for (int g = 0; g < G; ++g) a__[g] = lds_f64((aVal) + LO16(e[g]));         \
        _Pragma("unroll") for (int g = 0; g < G; ++g) sts_f64((aVal) + LO16(e[g]), a__[g] * facarr[g0 + g]); \
    })

// value = (value * fac[slot]) * vec[operand] for the slot's entries (OSQP's row scaling, then its column scaling, in that
// order of rounding); nrm[slot] = max(nrm[slot], |new value|)
#define DUO_SCALE2_ABSMAX(ROWSEL, N, facarr, nrmarr, aVal, aVec, mine, pad)                          \
    duo_ell_lines<ROWSEL, N>(X, mine, pad, [&](auto g0c, const auto &e) {                            \
        constexpr int g0 = decltype(g0c)::value, G = (int)(sizeof(e) / sizeof(e[0]));                \
        double a__[G], v__[G];                                                                       \
        _Pragma("unroll") for (int g = 0; g < G; ++g) {                                              \
            a__[g] = lds_f64((aVal) + LO16(e[g]));                                                   \
            v__[g] = lds_f64((aVec) + HI16(e[g]));                                                   \
        }                                                                                            \
        _Pragma("unroll") for (int g = 0; g < G; ++g) {                                              \
            const double t__ = (a__[g] * facarr[g0 + g]) * v__[g];                                   \
            sts_f64((aVal) + LO16(e[g]), t__);                                                       \
            nrmarr[g0 + g] = fmax(nrmarr[g0 + g], fabs(t__));                                        \
        }                                                                                            \
    })

#define ROWS_LOOP(j, i) _Pragma("unroll") for (int j = 0, i = oid; j < RPT; ++j, i += kDuoOwners) if (i < mu)
#define COLS_LOOP(c, p8) _Pragma("unroll") for (int c = 0; c < CPT; ++c) if (const int p8 = pos8[c]; p8 >= 0)
// entries e = (8 position << 16 | 8 value index) of the row this thread owns in slot j
#define ROW_ENTRIES(j, e)                                                                                          \
    for (uint32_t q__ = 0, n__ = (uint32_t)X.rlen[j], l__ = aRowMine + (uint32_t)X.rbase[j]; q__ < n__; ++q__, l__ += kDuoOwners * 4u) \
        if (const uint32_t e = lds_u32(l__); true)
// entries e = (8 compact row << 16 | 8 value index) of the column this thread owns in slot c
#define COL_ENTRIES(c, e)                                                                                          \
    for (uint32_t q__ = 0, n__ = (uint32_t)X.clen[c], l__ = aColMine + (uint32_t)X.cbase[c]; q__ < n__; ++q__, l__ += kDuoOwners * 4u) \
        if (const uint32_t e = lds_u32(l__); true)
// entries e = (8 other position << 16 | 8 kPu) of the full-symmetric P at the position owned in slot c (global memory)
#define PF_ENTRIES(c, e)                                                                                           \
    for (int q__ = 0, n__ = X.plen[c], l__ = X.pbase[c] + oid; q__ < n__; ++q__, l__ += kDuoOwners)                 \
        if (const uint32_t e = X.pf_ell[l__]; true)
#define LO16(e) ((e) & 0xFFFFu)
#define HI16(e) ((e) >> 16)

#define DUO_MARK(ph)                                                                       \
    do {                                                                                   \
        if (kTimers && B.prof != nullptr && oid == 0) {                                    \
            const long long now__ = clock64();                                             \
            atomicAdd(&B.prof[ph], (unsigned long long)(now__ - t_phase));                 \
            t_phase = now__;                                                               \
        }                                                                                  \
    } while (0)

// st.shared without the compiler-level memory clobber of sts_f64: global loads of later rows may be scheduled ahead of it
// (shared-memory accesses are all volatile asm and keep their order among themselves)
static __device__ __forceinline__ void sts_f64_nc(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v)); }

template <int RPT, int CPT, int NB>
static __device__ __forceinline__ void duo_owner_role(const DevSym &S, const DevBatch &B, const DevSettings &st, const DevDuo &X, int slot, int oid) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t sm0 = smem_u32(smem_raw);
    const uint32_t sl = sm0 + (uint32_t)(X.slot_base + slot * X.slot_bytes);
#define aA (sl + (uint32_t)X.oA)
#define aP (sl + (uint32_t)X.oP)
#define aW (sl + (uint32_t)X.oW)
#define aR (sl + (uint32_t)X.oRHS)
#define aLo (sl + (uint32_t)X.oLo)
#define aHi (sl + (uint32_t)X.oHi)
#define aOff (sl + (uint32_t)X.oOff)
#define aRing (sl + (uint32_t)X.oRing)
#define aCtl (sl + (uint32_t)X.oCtl)
#define aTab (sl + (uint32_t)X.oCtl + 32u)
    double *const sRed = reinterpret_cast<double *>(smem_raw + X.slot_base + slot * X.slot_bytes + X.oRed);
    uint64_t *const sBar = reinterpret_cast<uint64_t *>(smem_raw + X.slot_base + slot * X.slot_bytes + X.oCtl + 8);
    const uint32_t aRowMine = sm0 + (uint32_t)X.oRowEll + (uint32_t)oid * 4u, aColMine = sm0 + (uint32_t)X.oColEll + (uint32_t)oid * 4u;
    const int T = S.T, np = S.n_pad, mu = S.m_used, h = S.T / 2;
    constexpr int dpb = NB * (NB + 1) / 2;
    const int bSlot = duo_bar(slot, DB_SLOT), bOwn = duo_bar(slot, DB_OWNERS), bLoaded = duo_bar(slot, DB_RING_LOADED),
              bWritten = duo_bar(slot, DB_RING_WRITTEN);
#define OWNERS_SYNC() named_bar_sync(bOwn, kDuoOwners)

    const uint32_t rpad = ((uint32_t)(np * 8) << 16) | (uint32_t)(S.nnzA * 8), cpad = ((uint32_t)(mu * 8) << 16) | (uint32_t)(S.nnzA * 8);
    int pos8[CPT]; // 8 * position owned in column slot c, -1: none
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        const uint16_t p = c < X.cpt ? X.col_pos[c * kDuoOwners + oid] : (uint16_t)0xFFFFu; // (the table holds X.cpt <= CPT slots)
        pos8[c] = p == 0xFFFFu ? -1 : 8 * (int)p;
    }
    // zero elements behind the arrays (targets of surplus ELL entries)
    if (oid == 0) {
        sts_f64(aA + (uint32_t)S.nnzA * 8u, 0.0);
        sts_f64(aP + (uint32_t)S.nnzPu * 8u, 0.0);
        sts_f64(aW + (uint32_t)mu * 8u, 0.0);
        sts_f64(aR + (uint32_t)np * 8u, 0.0);
        mbar_init(sBar, 1);
    }
    uint32_t bar_phase = 0;
    OWNERS_SYNC();

    bool first = true;
    for (;;) {
        int widx;
        if (first) {
            widx = (int)blockIdx.x + slot * (int)gridDim.x; // the first QP of every slot is assigned statically: an even spread over the SMs
            first = false;
        } else {
            if (oid == 0) sts_u32(aCtl + 4u, (uint32_t)(atomicAdd(B.work_counter, 1) + kDuoSlots * (int)gridDim.x));
            OWNERS_SYNC();
            widx = (int)lds_u32(aCtl + 4u);
            OWNERS_SYNC(); // everybody has read the slot before it can be written again
        }
        if (widx >= B.batch) break;
        const int b = B.order != nullptr ? B.order[widx] : widx;
        if (B.active != nullptr && B.active[b] == 0) continue; // masked out (converged SQP problem)

        const bool hasq = B.q != nullptr;
        long long t_phase = kTimers ? clock64() : 0;

        // ------------------------------------------------------------------ load
        if (B.bulk_ok) {
            if (oid == 0) {
                mbar_expect_tx(sBar, (uint32_t)S.nnzA * 8u);
                bulk_g2s(smem_raw + X.slot_base + slot * X.slot_bytes + X.oA, B.A_val + (size_t)b * B.A_pitch, (uint32_t)S.nnzA * 8u, sBar);
            }
        } else {
            const double *gA = B.A_val + (size_t)b * B.A_pitch;
            for (int k = oid; k < S.nnzA; k += kDuoOwners) sts_f64(aA + (uint32_t)k * 8u, gA[k]);
        }
        for (int k0 = oid; k0 < S.nnzPu; k0 += 4 * kDuoOwners) { // (four loads in flight per step: the stores are compiler barriers)
            double v_[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) v_[g] = k0 + g * kDuoOwners < S.nnzPu ? S.Pu_val[k0 + g * kDuoOwners] : 0.0;
#pragma unroll
            for (int g = 0; g < 4; ++g)
                if (k0 + g * kDuoOwners < S.nnzPu) sts_f64(aP + (uint32_t)(k0 + g * kDuoOwners) * 8u, v_[g]);
        }
        {
            const double *gl = B.l + (size_t)b * S.m, *gu = B.u + (size_t)b * S.m;
            {   // (all of a thread's bounds requested before the first shared-memory store: the stores are compiler barriers,
                //  and row by row every row cost a dependent L2 + DRAM round trip of its own)
                int orig_[RPT];
                double lo_[RPT], hi_[RPT];
#pragma unroll
                for (int j = 0; j < RPT; ++j) orig_[j] = oid + j * kDuoOwners < mu ? S.used_rows[oid + j * kDuoOwners] : -1;
#pragma unroll
                for (int j = 0; j < RPT; ++j) {
                    lo_[j] = orig_[j] >= 0 ? gl[orig_[j]] : 0.0;
                    hi_[j] = orig_[j] >= 0 ? gu[orig_[j]] : 0.0;
                }
                ROWS_LOOP(j, i) {
                    sts_f64(aLo + (uint32_t)i * 8u, fmax(lo_[j], -kInfty)); // osqp-cpp clips bounds to +-OSQP_INFTY
                    sts_f64(aHi + (uint32_t)i * 8u, fmin(hi_[j], kInfty));
                }
            }
            double bad[1] = {0.0};
            // (eight rows per step: the index loads, then the sixteen bound loads of a step are in flight together -- the bounds
            //  come from DRAM, one dependent round trip per row made this loop two thirds of the load phase)
            for (int e0 = oid; e0 < S.n_empty; e0 += 8 * kDuoOwners) {
                int r_[8];
                double lo_[8], hi_[8];
#pragma unroll
                for (int g = 0; g < 8; ++g) r_[g] = e0 + g * kDuoOwners < S.n_empty ? S.empty_rows[e0 + g * kDuoOwners] : -1;
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    lo_[g] = r_[g] >= 0 ? gl[r_[g]] : 0.0;
                    hi_[g] = r_[g] >= 0 ? gu[r_[g]] : 0.0;
                }
#pragma unroll
                for (int g = 0; g < 8; ++g)
                    if (!(lo_[g] <= 0.0 && 0.0 <= hi_[g])) bad[0] = 1.0;
            }
            if (B.bulk_ok) {
                unsigned spins = 0;
                while (!mbar_try_wait(sBar, bar_phase)) {
                    if (++spins > (1u << 26)) __trap(); // a lost bulk copy must not hang the device
                }
                bar_phase ^= 1;
            }
            duo_max<1>(bad, sRed, oid, bOwn);
            if (oid == 0) sts_u32(aCtl + 96u, bad[0] != 0.0 ? 1u : 0u);
        }
        DUO_MARK(PH_LOAD);

        // ------------------------------------------------------------------ Ruiz equilibration
        {
            double cscale = 1.0;
            double Dp[CPT], qv[CPT];
            // E accumulates in the (still unused) w region, one entry per owned row; rn: row norms of A for the coming pass --
            // a pass scales A in ONE row-view traversal ((a * e_row) * d_col, the column factors read from the right-hand-side
            // region where they are published for P anyway) and takes the next pass's row norms from the values it writes:
            // two traversals of A per pass (column norms, scale) instead of four
            double rn[RPT];
            ROWS_LOOP(j, i) sts_f64(aW + (uint32_t)i * 8u, 1.0);
#pragma unroll
            for (int j = 0; j < RPT; ++j) rn[j] = 0.0;
#pragma unroll
            for (int c = 0; c < CPT; ++c) { qv[c] = 0.0; Dp[c] = 1.0; }
            if (hasq) {
                COLS_LOOP(c, p8) {
                    const int var = S.pos_var[p8 >> 3];
                    qv[c] = var >= 0 ? B.q[(size_t)b * S.n + var] : 0.0;
                }
            }
            // P-side patterns of this thread (static, global memory): read once per QP, held in registers over the passes
            constexpr int PFW = 4, PPT = 4;
            const uint32_t ppad = ((uint32_t)(np * 8) << 16) | (uint32_t)(S.nnzPu * 8);
            uint32_t pfe[CPT][PFW], puij[PPT];
#pragma unroll
            for (int c = 0; c < CPT; ++c)
#pragma unroll
                for (int q = 0; q < PFW; ++q) pfe[c][q] = q < X.plen[c] ? X.pf_ell[X.pbase[c] + q * kDuoOwners + oid] : ppad;
#pragma unroll
            for (int k = 0; k < PPT; ++k) puij[k] = oid + k * kDuoOwners < S.nnzPu ? S.Pu_ij[oid + k * kDuoOwners] : 0u;
            if (st.scaling > 0) DUO_ABSMAX(true, RPT, rn, aA, aRowMine, rpad);
#pragma unroll 1
            for (int it = 0; it < st.scaling; ++it) {
                double dt[CPT], et[RPT];
#pragma unroll
                for (int c = 0; c < CPT; ++c) dt[c] = 0.0;
#pragma unroll
                for (int c = 0; c < CPT; ++c)
#pragma unroll
                    for (int q = 0; q < PFW; ++q) dt[c] = fmax(dt[c], fabs(lds_f64(aP + LO16(pfe[c][q]))));
                DUO_ABSMAX(false, CPT, dt, aA, aColMine, cpad);
#pragma unroll
                for (int c = 0; c < CPT; ++c) dt[c] = 1.0 / sqrt(limit_scaling(dt[c]));
#pragma unroll
                for (int j = 0; j < RPT; ++j) { et[j] = 1.0 / sqrt(limit_scaling(rn[j])); rn[j] = 0.0; }
                COLS_LOOP(c, p8) sts_f64(aR + (uint32_t)p8, dt[c]);
                OWNERS_SYNC();
                DUO_SCALE2_ABSMAX(true, RPT, et, rn, aA, aR, aRowMine, rpad);
                ROWS_LOOP(j, i) sts_f64(aW + (uint32_t)i * 8u, lds_f64(aW + (uint32_t)i * 8u) * et[j]);
#pragma unroll
                for (int k = 0; k < PPT; ++k) {
                    const int kk = oid + k * kDuoOwners;
                    if (kk < S.nnzPu)
                        sts_f64(aP + (uint32_t)kk * 8u,
                                (lds_f64(aP + (uint32_t)kk * 8u) * lds_f64(aR + (puij[k] >> 16) * 8u)) * lds_f64(aR + (puij[k] & 0xFFFFu) * 8u));
                }
                OWNERS_SYNC();
                double psum = 0.0, qn[1] = {0.0};
                COLS_LOOP(c, p8) {
                    qv[c] *= dt[c];
                    Dp[c] *= dt[c];
                    double nrm = 0.0;
#pragma unroll
                    for (int q = 0; q < PFW; ++q) nrm = fmax(nrm, fabs(lds_f64(aP + LO16(pfe[c][q]))));
                    psum += nrm;
                    qn[0] = fmax(qn[0], fabs(qv[c]));
                }
                psum = duo_sum(psum, sRed, oid, bOwn);
                if (hasq) duo_max<1>(qn, sRed, oid, bOwn); // (q = 0: the norm is 0 on every thread without asking the others)
                double ct = fmax(psum / (double)S.n, limit_scaling(qn[0]));
                ct = 1.0 / limit_scaling(ct);
                for (int k = oid; k < S.nnzPu; k += kDuoOwners) sts_f64(aP + (uint32_t)k * 8u, lds_f64(aP + (uint32_t)k * 8u) * ct);
#pragma unroll
                for (int c = 0; c < CPT; ++c) qv[c] *= ct;
                cscale *= ct;
                OWNERS_SYNC();
            }
            double Ei[RPT];
#pragma unroll
            for (int j = 0; j < RPT; ++j) Ei[j] = 1.0;
            ROWS_LOOP(j, i) Ei[j] = lds_f64(aW + (uint32_t)i * 8u);
            const size_t bmu = (size_t)b * mu, bnp = (size_t)b * np;
            ROWS_LOOP(j, i) {
                sts_f64(aLo + (uint32_t)i * 8u, lds_f64(aLo + (uint32_t)i * 8u) * Ei[j]);
                sts_f64(aHi + (uint32_t)i * 8u, lds_f64(aHi + (uint32_t)i * 8u) * Ei[j]);
                B.Eg[bmu + i] = Ei[j];
            }
            COLS_LOOP(c, p8) {
                B.Dg[bnp + (p8 >> 3)] = Dp[c];
                if (hasq) B.qs[bnp + (p8 >> 3)] = qv[c];
            }
            if (oid == 0) {
                sts_f64(aCtl + 16u, cscale);
                sts_f64(aCtl + 24u, 1.0 / cscale);
            }
        }
        const bool uns = !st.scaled_termination;
        DUO_MARK(PH_SCALE);

        // ------------------------------------------------------------------ rho classification
        unsigned ctype = 0; // 2 bits per owned row: 0 inequality, 1 equality, 2 loose
        ROWS_LOOP(j, i) {
            const double lo = lds_f64(aLo + (uint32_t)i * 8u), hi = lds_f64(aHi + (uint32_t)i * 8u);
            unsigned t = 0;
            if (lo < -kInfty * kMinScaling && hi > kInfty * kMinScaling) t = 2;
            else if (hi - lo < kRhoTol) t = 1;
            ctype |= t << (2 * j);
        }

        // ------------------------------------------------------------------ iterates
        double z[RPT], y[RPT];
        double x[CPT];
        const int flags = B.flag_bits;
        int rho_updates;
        {
            double rho;
            const size_t bmu = (size_t)b * mu, bnp = (size_t)b * np;
#pragma unroll
            for (int j = 0; j < RPT; ++j) { z[j] = 0.0; y[j] = 0.0; }
#pragma unroll
            for (int c = 0; c < CPT; ++c) x[c] = 0.0;
            if (flags & 1) {
                rho = fmin(fmax(st.rho, kRhoMin), kRhoMax);
                rho_updates = 0;
            } else {
                ROWS_LOOP(j, i) {
                    z[j] = B.z_state[bmu + i];
                    y[j] = B.y_state[bmu + i];
                }
                COLS_LOOP(c, p8) x[c] = B.x_state[bnp + (p8 >> 3)];
                rho = B.rho_state[b];
                rho_updates = B.rho_updates[b];
            }
            if (oid == 0) { // rho / 1/rho per row type (inequality, equality, loose)
                sts_f64(aTab, rho);
                sts_f64(aTab + 8u, kRhoEqOverIneq * rho);
                sts_f64(aTab + 16u, kRhoMin);
                sts_f64(aTab + 32u, 1.0 / rho);
                sts_f64(aTab + 40u, 1.0 / (kRhoEqOverIneq * rho));
                sts_f64(aTab + 48u, 1.0 / kRhoMin);
            }
            if (flags & 2) { // osqp_warm_start_x: x <- Dinv x0, z <- A x
                {   // (loads first, shared-memory stores after: see the bounds above)
                    int var_[CPT];
                    double d_[CPT], x0_[CPT];
#pragma unroll
                    for (int c = 0; c < CPT; ++c) var_[c] = pos8[c] >= 0 ? S.pos_var[pos8[c] >> 3] : -1;
#pragma unroll
                    for (int c = 0; c < CPT; ++c) {
                        d_[c] = pos8[c] >= 0 ? B.Dg[bnp + (pos8[c] >> 3)] : 1.0;
                        x0_[c] = var_[c] >= 0 ? B.x0[(size_t)b * S.n + var_[c]] : 0.0;
                    }
                    COLS_LOOP(c, p8) {
                        x[c] = var_[c] >= 0 ? (1.0 / d_[c]) * x0_[c] : 0.0;
                        sts_f64(aR + (uint32_t)p8, x[c]);
                    }
                }
                OWNERS_SYNC();
#pragma unroll
                for (int j = 0; j < RPT; ++j) z[j] = 0.0;
                DUO_DOT(true, RPT, z, aA, aR, aRowMine, rpad);
            }
            OWNERS_SYNC();
        }
#define RHO_OF(t) lds_f64(aTab + (t) * 8u)
#define RINV_OF(t) lds_f64(aTab + 32u + (t) * 8u)

        // ------------------------------------------------------------------ ADMM loop
        int status = lds_u32(aCtl + 96u) ? -3 : -10; // OSQP status_val (an empty row excluding 0 makes the QP infeasible outright)
        int iters_done = 0;
        bool need_factor = true, w_dirty = true;

#pragma unroll 1
        for (int iter = 1; status == -10 && iter <= st.max_iter; ++iter) {
            if (need_factor) {
                // ---------------- assemble S block by block, one round ahead of the substitution warps that factor it
                ROWS_LOOP(j, i) {
                    const unsigned t = (ctype >> (2 * j)) & 3u;
                    sts_f64(aW + (uint32_t)i * 8u, RHO_OF(t));
                }
                if (oid == 0) sts_u32(aCtl, DUO_FACTOR);
                named_bar_sync(bSlot, kDuoSlotThreads); // (E)
                const unsigned long long apad = (unsigned long long)(mu * 8) | ((unsigned long long)(S.nnzA * 8) << 16) | ((unsigned long long)(S.nnzA * 8) << 32);
                // sum of rho[row] A[k1] A[k2] over a list of packed contribution words (stride kDuoOwners), four loads at a time
                auto contrib_sum = [&](const unsigned long long *cp, int maxc) {
                    double acc = 0.0;
#pragma unroll 1
                    for (int k0 = 0; k0 < maxc; k0 += 4) {
                        unsigned long long cw_[4];
#pragma unroll
                        for (int g = 0; g < 4; ++g) cw_[g] = k0 + g < maxc ? cp[(size_t)(k0 + g) * kDuoOwners] : apad;
                        double r_[4], a1_[4], a2_[4];
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            r_[g] = lds_f64(aW + (uint32_t)(cw_[g] & 0xFFFFu));
                            a1_[g] = lds_f64(aA + (uint32_t)((cw_[g] >> 16) & 0xFFFFu));
                            a2_[g] = lds_f64(aA + (uint32_t)((cw_[g] >> 32) & 0xFFFFu));
                        }
#pragma unroll
                        for (int g = 0; g < 4; ++g) acc += r_[g] * a1_[g] * a2_[g];
                    }
                    return acc;
                };
#pragma unroll 1
                for (int w = 0; w < X.as_opt; ++w) {
                    const int e = oid + w * kDuoOwners;
                    const uint32_t meta = X.as_ometa[w * kDuoOwners + oid];
                    double acc = contrib_sum(X.as_ocon + (size_t)w * X.as_maxco * kDuoOwners + oid, X.as_maxco);
                    acc += lds_f64(aP + (meta & 0xFFFFu));
                    if (e < S.noff) sts_f64(aOff + (uint32_t)e * 8u, acc);
                }
#pragma unroll 1
                for (int R = 0; R <= h; ++R) {
                    // blocks of round R: stage R of the down chain, stage T-1-R of the up chain; round h: the middle stage
                    constexpr int EPT = (2 * dpb + kDuoOwners - 1) / kDuoOwners; // packed entries per owner and round
                    double val[EPT];
                    int dst[EPT];
#pragma unroll
                    for (int w = 0; w < EPT; ++w) {
                        const int e = oid + w * kDuoOwners;
                        val[w] = 0.0;
                        dst[w] = -1;
                        const uint32_t meta = X.as_meta[(R * EPT + w) * kDuoOwners + oid];
                        double acc = contrib_sum(X.as_con + (size_t)(R * EPT + w) * X.as_maxc * kDuoOwners + oid, X.as_maxc);
                        acc += lds_f64(aP + (meta & 0xFFFFu));
                        const uint32_t kind = (meta >> 16) & 3u;
                        if (kind == 1u) acc += st.sigma;
                        else if (kind == 2u) acc = 1.0;
                        val[w] = acc;
                        if ((meta >> 18) & 1u) dst[w] = e >= dpb ? e + 1 : e; // (second block of the round: behind the first, padded by one)
                    }
                    if (R > 0) named_bar_sync(bLoaded, kDuoSlotThreads); // round R-1 has been read out of the ring
#pragma unroll
                    for (int w = 0; w < EPT; ++w)
                        if (dst[w] >= 0) sts_f64(aRing + (uint32_t)dst[w] * 8u, val[w]);
                    named_bar_arrive(bWritten, kDuoSlotThreads);
                }
                named_bar_sync(bSlot, kDuoSlotThreads);                    // (D)
                if (oid == 0) sts_f64(aR + (uint32_t)np * 8u, 0.0); // (the ring / scratch may alias the right-hand side region)
                need_factor = false;
                w_dirty = true;
                DUO_MARK(PH_FACTOR);
            }
            if (w_dirty) {
                ROWS_LOOP(j, i) {
                    const unsigned t = (ctype >> (2 * j)) & 3u;
                    sts_f64(aW + (uint32_t)i * 8u, fma(RHO_OF(t), z[j], -y[j]));
                }
                w_dirty = false;
                OWNERS_SYNC();
            }
            // ---------------- rhs = sigma x - q + A'(rho z - y)
            {
                double acc[CPT];
#pragma unroll
                for (int c = 0; c < CPT; ++c) acc[c] = st.sigma * x[c];
                if (hasq) {
                    COLS_LOOP(c, p8) acc[c] -= B.qs[(size_t)b * np + (p8 >> 3)];
                }
                DUO_DOT(false, CPT, acc, aA, aW, aColMine, cpad);
                COLS_LOOP(c, p8) sts_f64(aR + (uint32_t)p8, acc[c]);
            }
            if (oid == 0) sts_u32(aCtl, DUO_SOLVE);
            DUO_MARK(PH_RHS);
            named_bar_sync(bSlot, kDuoSlotThreads); // (E) rhs published; the substitution warps solve
            named_bar_sync(bSlot, kDuoSlotThreads); // (F) x~ in sRHS
            DUO_MARK(PH_SOLVE);

            const bool last = iter == st.max_iter;
            const bool can_check = st.check_termination && (iter % st.check_termination == 0);
            const bool adapt_now = st.adaptive_rho && st.adaptive_rho_interval && (iter % st.adaptive_rho_interval == 0);
            const bool do_info = can_check || adapt_now || last;
            // ---------------- x, z, y updates (z~ = A x~ fused with the projection and the dual step)
            COLS_LOOP(c, p8) {
                const double xn = st.alpha * lds_f64(aR + (uint32_t)p8) + (1.0 - st.alpha) * x[c];
                if (do_info && hasq) B.dxg[(size_t)b * np + (p8 >> 3)] = xn - x[c];
                x[c] = xn;
            }
            long long t_sub = kTimers ? clock64() : 0;
            double ztv[RPT];
#pragma unroll
            for (int j = 0; j < RPT; ++j) ztv[j] = 0.0;
            DUO_DOT(true, RPT, ztv, aA, aR, aRowMine, rpad);
            if (kTimers && B.prof != nullptr && oid == 0) { const long long n_ = clock64(); atomicAdd(&B.prof[13], (unsigned long long)(n_ - t_sub)); t_sub = n_; }
            // two rows at a time: bounds and rho of both first (independent loads), then the two dependent chains side by side
#pragma unroll
            for (int j0 = 0; j0 < RPT; j0 += 2) {
                double lo_[2], hi_[2], rho_[2], rinv_[2];
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    if (j0 + g < RPT) {
                        const uint32_t i8 = (uint32_t)(oid + (j0 + g) * kDuoOwners) * 8u;
                        const unsigned t = (ctype >> (2 * (j0 + g))) & 3u;
                        lo_[g] = lds_f64(aLo + i8);
                        hi_[g] = lds_f64(aHi + i8);
                        rho_[g] = RHO_OF(t);
                        rinv_[g] = RINV_OF(t);
                    }
                }
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    if (j0 + g < RPT) {
                        const int j = j0 + g, i = oid + j * kDuoOwners;
                        const double zr = st.alpha * ztv[j] + (1.0 - st.alpha) * z[j];
                        const double zn = fmin(fmax(zr + rinv_[g] * y[j], lo_[g]), hi_[g]);
                        const double d = rho_[g] * (zr - zn);
                        y[j] += d;
                        z[j] = zn;
                        if (i < mu) {
                            sts_f64(aW + (uint32_t)i * 8u, fma(rho_[g], zn, -y[j])); // w = rho z - y for the next gather
                            if (do_info) B.dyg[(size_t)b * mu + i] = d;
                        }
                    }
                }
            }
            if (kTimers && B.prof != nullptr && oid == 0) { const long long n_ = clock64(); atomicAdd(&B.prof[14], (unsigned long long)(n_ - t_sub)); t_sub = n_; }
            OWNERS_SYNC(); // sRHS is overwritten by the next gather, which reads sW
            if (kTimers && B.prof != nullptr && oid == 0) { const long long n_ = clock64(); atomicAdd(&B.prof[15], (unsigned long long)(n_ - t_sub)); t_sub = n_; }
            DUO_MARK(PH_UPDATE);

            if (do_info) {
                // ---------------- residuals and the norms OSQP's tolerances need (update_info)
                const size_t bmu = (size_t)b * mu, bnp = (size_t)b * np;
                const double cscale = lds_f64(aCtl + 16u), cinv = lds_f64(aCtl + 24u);
                COLS_LOOP(c, p8) sts_f64(aR + (uint32_t)p8, x[c]);
                ROWS_LOOP(j, i) sts_f64(aW + (uint32_t)i * 8u, y[j]);
                w_dirty = true;
                OWNERS_SYNC();
                double pri_u, pri_s, pn_u, pn_s, dua_u, dua_s, dn_u, dn_s; // residuals / normalisers, unscaled and scaled
                {
                    double na[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                    double axv[RPT];
#pragma unroll
                    for (int j = 0; j < RPT; ++j) axv[j] = 0.0;
                    DUO_DOT(true, RPT, axv, aA, aR, aRowMine, rpad);
                    ROWS_LOOP(j, i) {
                        const double ax = axv[j];
                        const double pr = ax - z[j];
                        const double einv = 1.0 / B.Eg[bmu + i];
                        na[0] = fmax(na[0], fabs(einv * pr));
                        na[1] = fmax(na[1], fabs(pr));
                        na[2] = fmax(na[2], fabs(einv * z[j]));
                        na[3] = fmax(na[3], fabs(einv * ax));
                        na[4] = fmax(na[4], fabs(z[j]));
                        na[5] = fmax(na[5], fabs(ax));
                    }
                    duo_max<6>(na, sRed, oid, bOwn);
                    pri_u = na[0]; pri_s = na[1]; pn_u = fmax(na[2], na[3]); pn_s = fmax(na[4], na[5]);
                }
                {
                    double nb[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                    double atyv[CPT];
#pragma unroll
                    for (int c = 0; c < CPT; ++c) atyv[c] = 0.0;
                    DUO_DOT(false, CPT, atyv, aA, aW, aColMine, cpad);
                    COLS_LOOP(c, p8) {
                        double px = 0.0;
                        PF_ENTRIES(c, e) px = fma(lds_f64(aP + LO16(e)), lds_f64(aR + HI16(e)), px);
                        const double aty = atyv[c];
                        const double qq = hasq ? B.qs[bnp + (p8 >> 3)] : 0.0;
                        const double du = qq + px + aty;
                        const double dinv = 1.0 / B.Dg[bnp + (p8 >> 3)];
                        nb[0] = fmax(nb[0], fabs(dinv * du));
                        nb[1] = fmax(nb[1], fabs(du));
                        nb[2] = fmax(nb[2], fabs(dinv * px));
                        nb[3] = fmax(nb[3], fabs(dinv * aty));
                        nb[4] = fmax(nb[4], fabs(dinv * qq));
                        nb[5] = fmax(nb[5], fabs(px));
                        nb[6] = fmax(nb[6], fabs(aty));
                        nb[7] = fmax(nb[7], fabs(qq));
                    }
                    duo_max<8>(nb, sRed, oid, bOwn);
                    dua_u = nb[0]; dua_s = nb[1]; dn_u = fmax(nb[4], fmax(nb[3], nb[2])); dn_s = fmax(nb[7], fmax(nb[6], nb[5]));
                }
                const double pri_res = uns ? pri_u : pri_s;
                const double dua_res = uns ? cinv * dua_u : dua_s;
                iters_done = iter;
                if (oid == 0) {
                    B.pri_res[b] = pri_res;
                    B.dua_res[b] = dua_res;
                }

                // ---------------- termination: stage 0 exact tolerances, stage 1 (only at max_iter) 10x tolerances
                bool done = false;
#pragma unroll 1
                for (int stage = 0; stage < 2 && !done; ++stage) {
                    if (stage == 0 && !(can_check || last)) continue;
                    if (stage == 1) {
                        if (adapt_now) { // OSQP adapt_rho: estimate from the scaled residuals
                            const double rho = RHO_OF(0u);
                            const double pr = pri_s / (pn_s + kDivTol);
                            const double dr = dua_s / (dn_s + kDivTol);
                            const double est = fmin(fmax(rho * sqrt(pr / dr), kRhoMin), kRhoMax);
                            if (est > rho * st.adaptive_rho_tolerance || est < rho / st.adaptive_rho_tolerance) {
                                OWNERS_SYNC(); // everybody has read the old rho
                                if (oid == 0) {
                                    sts_f64(aTab, est);
                                    sts_f64(aTab + 8u, kRhoEqOverIneq * est);
                                    sts_f64(aTab + 32u, 1.0 / est);
                                    sts_f64(aTab + 40u, 1.0 / (kRhoEqOverIneq * est));
                                }
                                OWNERS_SYNC();
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
                            const double l_ = lds_f64(aLo + (uint32_t)i * 8u), u_ = lds_f64(aHi + (uint32_t)i * 8u);
                            if (u_ > kInfty * kMinScaling) {
                                if (l_ < -kInfty * kMinScaling) d = 0.0;
                                else d = fmin(d, 0.0);
                            } else if (l_ < -kInfty * kMinScaling) {
                                d = fmax(d, 0.0);
                            }
                            sts_f64_nc(aW + (uint32_t)i * 8u, d); // (delta_y and E of the other rows are requested ahead of it)
                            nrm[0] = fmax(nrm[0], fabs(uns ? B.Eg[bmu + i] * d : d));
                            lhs += u_ * fmax(d, 0.0) + l_ * fmin(d, 0.0);
                        }
                        duo_max<1>(nrm, sRed, oid, bOwn); // (also orders the sW writes before the gather below)
                        lhs = duo_sum(lhs, sRed, oid, bOwn);
                        if (nrm[0] > kDivTol && lhs < -epi * nrm[0]) {
                            double v[1] = {0.0};
                            double accv[CPT];
#pragma unroll
                            for (int c = 0; c < CPT; ++c) accv[c] = 0.0;
                            DUO_DOT(false, CPT, accv, aA, aW, aColMine, cpad);
                            COLS_LOOP(c, p8) v[0] = fmax(v[0], fabs(uns ? (1.0 / B.Dg[bnp + (p8 >> 3)]) * accv[c] : accv[c]));
                            duo_max<1>(v, sRed, oid, bOwn);
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
                        COLS_LOOP(c, p8) {
                            const double d = B.dxg[bnp + (p8 >> 3)];
                            nrm[0] = fmax(nrm[0], fabs(uns ? B.Dg[bnp + (p8 >> 3)] * d : d));
                            qdx += B.qs[bnp + (p8 >> 3)] * d;
                        }
                        duo_max<1>(nrm, sRed, oid, bOwn);
                        qdx = duo_sum(qdx, sRed, oid, bOwn);
                        const double cs_ = uns ? cscale : 1.0;
                        if (nrm[0] > kDivTol && qdx < -cs_ * edi * nrm[0]) {
                            COLS_LOOP(c, p8) sts_f64(aR + (uint32_t)p8, B.dxg[bnp + (p8 >> 3)]);
                            OWNERS_SYNC();
                            double v[1] = {0.0};
                            COLS_LOOP(c, p8) {
                                double acc = 0.0;
                                PF_ENTRIES(c, e) acc = fma(lds_f64(aP + LO16(e)), lds_f64(aR + HI16(e)), acc);
                                v[0] = fmax(v[0], fabs(uns ? (1.0 / B.Dg[bnp + (p8 >> 3)]) * acc : acc));
                            }
                            duo_max<1>(v, sRed, oid, bOwn);
                            if (v[0] < cs_ * edi * nrm[0]) {
                                double viol[1] = {0.0};
                                double adx[RPT];
#pragma unroll
                                for (int j = 0; j < RPT; ++j) adx[j] = 0.0;
                                DUO_DOT(true, RPT, adx, aA, aR, aRowMine, rpad);
                                ROWS_LOOP(j, i) {
                                    double acc = adx[j];
                                    if (uns) acc *= 1.0 / B.Eg[bmu + i];
                                    if ((lds_f64(aHi + (uint32_t)i * 8u) < kInfty * kMinScaling && acc > edi * nrm[0]) ||
                                        (lds_f64(aLo + (uint32_t)i * 8u) > -kInfty * kMinScaling && acc < -edi * nrm[0]))
                                        viol[0] = 1.0;
                                }
                                duo_max<1>(viol, sRed, oid, bOwn);
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
                DUO_MARK(PH_INFO);
            }
        }
#undef RHO_OF
#undef RINV_OF

        // ------------------------------------------------------------------ store
        {
            const size_t bmu = (size_t)b * mu, bnp = (size_t)b * np;
            const bool infeasible = status == -3 || status == 3 || status == -4 || status == 4 || status == -7;
            const double qnan = __longlong_as_double(0x7ff8000000000000LL);
            const double cinv = lds_f64(aCtl + 24u);
            double *xo = B.x_out + (size_t)b * S.n, *yo = B.y_out + (size_t)b * S.m;
            COLS_LOOP(c, p8) {
                const int var = S.pos_var[p8 >> 3];
                if (var >= 0) xo[var] = infeasible ? qnan : B.Dg[bnp + (p8 >> 3)] * x[c];
                B.x_state[bnp + (p8 >> 3)] = infeasible ? 0.0 : x[c];
            }
            ROWS_LOOP(j, i) {
                const double e = B.Eg[bmu + i];
                yo[S.used_rows[i]] = infeasible ? qnan : (e * y[j]) * cinv;
                B.z_state[bmu + i] = infeasible ? 0.0 : z[j];
                B.y_state[bmu + i] = infeasible ? 0.0 : y[j];
            }
            for (int e0 = oid; e0 < S.n_empty; e0 += 8 * kDuoOwners) {
                int r_[8];
#pragma unroll
                for (int g = 0; g < 8; ++g) r_[g] = e0 + g * kDuoOwners < S.n_empty ? S.empty_rows[e0 + g * kDuoOwners] : -1;
#pragma unroll
                for (int g = 0; g < 8; ++g)
                    if (r_[g] >= 0) yo[r_[g]] = infeasible ? qnan : 0.0;
            }
            if (oid == 0) {
                B.rho_state[b] = lds_f64(aTab);
                B.status[b] = status;
                B.iters[b] = iters_done;
                B.rho_updates[b] = rho_updates;
                if (iters_done == 0) {
                    B.pri_res[b] = 0.0;
                    B.dua_res[b] = 0.0;
                }
            }
        }
        fence_proxy_async(); // generic-proxy accesses to sA are ordered before the next QP's bulk copy
        OWNERS_SYNC();       // every owner is done with this QP's shared data before the next one is loaded
        DUO_MARK(PH_STORE);
    }
    if (oid == 0) sts_u32(aCtl, DUO_EXIT);
    named_bar_sync(bSlot, kDuoSlotThreads); // (E) releases the substitution warps
#undef OWNERS_SYNC
#undef aA
#undef aP
#undef aW
#undef aR
#undef aLo
#undef aHi
#undef aOff
#undef aRing
#undef aCtl
#undef aTab
}

template <int RPT, int CPT, int NB>
__global__ void __launch_bounds__(kDuoThreads, 1) gomp_duo_kernel(const __grid_constant__ DevSym S, const __grid_constant__ DevBatch B,
                                                                  const __grid_constant__ DevSettings st, const __grid_constant__ DevDuo X) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // ---- once per kernel: the ELL pattern views -> shared memory; zero the substitution warps' exchange lines
    {
        uint32_t *re = reinterpret_cast<uint32_t *>(smem_raw + X.oRowEll), *ce = reinterpret_cast<uint32_t *>(smem_raw + X.oColEll);
        for (int k = tid; k < X.n_row_ell; k += kDuoThreads) re[k] = X.row_ell[k];
        for (int k = tid; k < X.n_col_ell; k += kDuoThreads) ce[k] = X.col_ell[k];
        for (int s = 0; s < kDuoSlots; ++s) {
            double *v = reinterpret_cast<double *>(smem_raw + X.slot_base + s * X.slot_bytes + X.oV);
            double *scr = reinterpret_cast<double *>(smem_raw + X.slot_base + s * X.slot_bytes + X.oScr);
            for (int k = tid; k < 64; k += kDuoThreads) v[k] = 0.0;
            for (int k = tid; k < 2 * NB * NB + 64; k += kDuoThreads) scr[k] = 0.0;
        }
    }
    if (warp == 0) tmem_alloc(&s_tmem, kDuoTmemCols);
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    const uint32_t tmem = s_tmem;
    // The substitution warps are the four HIGHEST warps of the CTA (the issue arbiter favours high warp ids: their dependent
    // chain is the critical path of every iteration), one per TMEM lane quarter.
    constexpr int kOwnerThreads = kDuoSlots * kDuoOwners;
    if (tid >= kOwnerThreads) {
        const int cwp = warp - kOwnerThreads / 32; // 0..3
        duo_chain_role<NB>(S, X, B.prof, tmem, cwp >> 1, cwp & 1, lane);
    } else {
        duo_owner_role<RPT, CPT, NB>(S, B, st, X, tid / kDuoOwners, tid % kDuoOwners);
    }
    tmem_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, kDuoTmemCols);
}

#undef ROWS_LOOP
#undef COLS_LOOP
#undef ROW_ENTRIES
#undef COL_ENTRIES
#undef PF_ENTRIES
#undef DUO_MARK
#undef LO16
#undef HI16
#undef DUO_DOT
#undef DUO_ABSMAX
#undef DUO_SCALE

// ---- launch interface ------------------------------------------------------------------------------------
#if GOMP_DUO_OWNERS >= 224
#define DUO_VARIANTS(F) F(7, 4, 14) F(8, 4, 14) F(6, 3, 12)
#else
#define DUO_VARIANTS(F) F(9, 4, 14) F(10, 4, 14) F(7, 4, 12)
#endif

bool duo_variant_exists(int rpt, int cpt, int bs) {
#define CHK(R, C, N) if (rpt <= R && cpt <= C && bs == N) return true;
    DUO_VARIANTS(CHK)
#undef CHK
    return false;
}
cudaError_t prepare_duo(int rpt, int cpt, int bs, size_t smem_bytes, int *ctas_per_sm) {
#define PREP(R, C, N)                                                                                       \
    if (rpt <= R && cpt <= C && bs == N) {                                                                  \
        auto fn = gomp_duo_kernel<R, C, N>;                                                                 \
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes); \
        if (e != cudaSuccess) return e;                                                                     \
        cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); \
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, fn, kDuoThreads, smem_bytes);     \
    }
    DUO_VARIANTS(PREP)
#undef PREP
    return cudaErrorInvalidValue;
}
cudaError_t launch_duo(int rpt, int cpt, int bs, const DevSym &S, const DevBatch &B, const DevSettings &st, const DevDuo &X, int grid,
                       size_t smem_bytes, cudaStream_t stream) {
#define LAUNCH(R, C, N)                                                                                     \
    if (rpt <= R && cpt <= C && bs == N) {                                                                  \
        gomp_duo_kernel<R, C, N><<<grid, kDuoThreads, smem_bytes, stream>>>(S, B, st, X);                   \
        return cudaGetLastError();                                                                          \
    }
    DUO_VARIANTS(LAUNCH)
#undef LAUNCH
    return cudaErrorInvalidValue;
}

} // namespace gomp_dev
