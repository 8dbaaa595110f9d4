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
#include "gomp_stream.h"

namespace gomp_dev {

namespace {

constexpr int kST = 512;    // threads per CTA
constexpr int kGroup = 224; // threads per factor group
constexpr int kMaxW = 4;    // off-diagonal ELL width kept in registers by the substitution warps

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <int K>
__device__ __forceinline__ void block_max(double (&v)[K], double *sRed, int tid) {
    constexpr int NW = kST / 32;
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
__device__ __forceinline__ double block_sum(double a, double *sRed, int tid) {
    constexpr int NW = kST / 32;
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

// ---- factorisation: one pivot block, by one 224-thread group, entries in registers --------------------------
// mode bit 1: subtract Off_{t-1} Dinv_{t-1} Off_{t-1}' first; bit 2: subtract Off_t' Dinv_{t+1} Off_t.
__device__ __forceinline__ void factor_block_stream(const DevSym &S, double *Dall, const double *Offall, double *pub, int t, int mode,
                                                    int g, int gl) {
    constexpr int EPL = 5; // ceil(32 * 32 / 224)
    const int bs = S.bs, nn = bs * bs;
    double *M = Dall + (size_t)t * nn;
    double val[EPL];
    int ri[EPL], ci[EPL];
#pragma unroll
    for (int q = 0; q < EPL; ++q) {
        const int e = gl + q * kGroup;
        ri[q] = e < nn ? e / bs : -1;
        ci[q] = e < nn ? e - ri[q] * bs : -1;
        val[q] = e < nn ? M[e] : 0.0;
    }
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
                    a = __ldg(S.off_col + (tb * bs + r) * S.ow + s1);
                    if (a >= 0) va = blk[r * S.ow + s1];
                } else {
                    a = __ldg(S.offT_row + (tb * bs + r) * S.owT + s1);
                    if (a >= 0) va = blk[__ldg(S.offT_slot + (tb * bs + r) * S.owT + s1)];
                }
                if (a < 0) continue;
                for (int s2 = 0; s2 < cnt; ++s2) {
                    int b2;
                    double vb = 0.0;
                    if (above) {
                        b2 = __ldg(S.off_col + (tb * bs + c) * S.ow + s2);
                        if (b2 >= 0) vb = blk[c * S.ow + s2];
                    } else {
                        b2 = __ldg(S.offT_row + (tb * bs + c) * S.owT + s2);
                        if (b2 >= 0) vb = blk[__ldg(S.offT_slot + (tb * bs + c) * S.owT + s2)];
                    }
                    if (b2 >= 0) acc += va * Dn[a * bs + b2] * vb;
                }
            }
            val[q] -= acc;
        }
    }
#pragma unroll 1
    for (int k = 0; k < bs; ++k) { // Gauss-Jordan without pivoting (SPD block)
        double *rowk = pub + (k & 1) * 64, *colk = rowk + 32;
#pragma unroll
        for (int q = 0; q < EPL; ++q) {
            if (ri[q] == k) rowk[ci[q]] = val[q];
            if (ci[q] == k) colk[ri[q]] = val[q];
        }
        named_bar_sync(4 + g, kGroup);
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
        if (ri[q] >= 0) M[gl + q * kGroup] = val[q];
    named_bar_sync(4 + g, kGroup); // the next stage of this chain reads the inverse just stored
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

} // namespace

__global__ void __launch_bounds__(kST, 1) gomp_admm_stream_kernel(const __grid_constant__ DevSym S, const __grid_constant__ DevBatch B,
                                                                  const __grid_constant__ DevSettings st,
                                                                  const __grid_constant__ DevStream X) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bs = S.bs, T = S.T, np = S.n_pad, mu = S.m_used, h = S.T / 2, bb = bs * bs;
    double *const sRHS = reinterpret_cast<double *>(smem_raw + X.oRHS);
    double *const sRing = reinterpret_cast<double *>(smem_raw + X.oRing);
    double *const sPub = reinterpret_cast<double *>(smem_raw + X.oPub);
    double *const sV = reinterpret_cast<double *>(smem_raw + X.oV);
    double *const sRed = reinterpret_cast<double *>(smem_raw + X.oRed);
    uint64_t *const sFull = reinterpret_cast<uint64_t *>(smem_raw + X.oBar); // [2][kRing]
    int *const sWork = reinterpret_cast<int *>(smem_raw + X.oBar + 8 * 2 * kStreamRing);

    // per-CTA workspace (global, L2-resident)
    double *const ws = X.ws + (size_t)blockIdx.x * X.stride;
    double *const gA = ws + X.wA, *const gP = ws + X.wP, *const gW = ws + X.wW, *const gRho = ws + X.wRho, *const gDinv = ws + X.wDinv,
                 *const gOff = ws + X.wOff, *const gT = ws + X.wT, *const gLo = ws + X.wLo, *const gHi = ws + X.wHi;

    if (tid == 0) {
        for (int k = 0; k < 2 * kStreamRing; ++k) mbar_init(sFull + k, 1);
    }
    uint32_t ring_k = 0; // slots this warp has consumed since the kernel started (slot = ring_k % kRing, parity from ring_k / kRing)
    __syncthreads();

    const int slotD = X.slot_doubles;
    const int halves = bs <= 16 ? 2 : 1;
    const int rr_ = halves == 2 ? (lane & 15) : lane, half = halves == 2 ? (lane >> 4) : 0;
    const bool rowact = rr_ < bs;
    const int r = rowact ? rr_ : bs - 1;
    const int cchunk = (bs + halves - 1) / halves, c0 = half * cchunk, c1 = min(bs, c0 + cchunk);

    for (;;) {
        if (tid == 0) *sWork = atomicAdd(B.work_counter, 1);
        __syncthreads();
        const int b = *sWork;
        __syncthreads();
        if (b >= B.batch) break;
        const size_t bmu = (size_t)b * mu, bnp = (size_t)b * np;
        const bool hasq = B.q != nullptr;
        double *const gx = B.x_state + bnp, *const gz = B.z_state + bmu, *const gy = B.y_state + bmu;
        double *const gE = B.Eg + bmu, *const gD = B.Dg + bnp, *const gdy = B.dyg + bmu, *const gdx = B.dxg + bnp;
        double *const gq = hasq ? B.qs + bnp : nullptr;

        // ------------------------------------------------------------------ load
        {
            const double *srcA = B.A_val + (size_t)b * B.A_pitch;
            for (int k = tid; k < S.nnzA; k += kST) gA[k] = srcA[k];
            for (int k = tid; k < S.nnzPu; k += kST) gP[k] = S.Pu_val[k];
        }
        const double *gl = B.l + (size_t)b * S.m, *gu = B.u + (size_t)b * S.m;
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
        double bad[1] = {0.0};
        for (int e = tid; e < S.n_empty; e += kST) {
            const int rw = S.empty_rows[e];
            if (!(gl[rw] <= 0.0 && 0.0 <= gu[rw])) bad[0] = 1.0;
        }
        block_max<1>(bad, sRed, tid); // (also: the loads above are visible to every thread)
        const bool empty_infeasible = bad[0] != 0.0;

        // ------------------------------------------------------------------ Ruiz equilibration
        double cscale = 1.0;
#pragma unroll 1
        for (int it = 0; it < st.scaling; ++it) {
            for (int p = tid; p < np; p += kST) {
                const double nrm = fmax(pf_absmax_g(S, p, gP), col_absmax_g(S, p, gA));
                gT[p] = 1.0 / sqrt(limit_scaling(nrm));
            }
            for (int i = tid; i < mu; i += kST) {
                const uint32_t rg = __ldg(S.row_range + i);
                double nrm = 0.0;
                for (int q = (int)(rg >> 8), q1 = q + (int)(rg & 0xFFu); q < q1; ++q) nrm = fmax(nrm, fabs(gA[__ldg(S.row_csr + q) & 0xFFFFu]));
                gW[i] = 1.0 / sqrt(limit_scaling(nrm));
            }
            __syncthreads();
            for (int i = tid; i < mu; i += kST) {
                const uint32_t rg = __ldg(S.row_range + i);
                const double et = gW[i];
                for (int q = (int)(rg >> 8), q1 = q + (int)(rg & 0xFFu); q < q1; ++q) gA[__ldg(S.row_csr + q) & 0xFFFFu] *= et;
                gE[i] *= et;
            }
            for (int k = tid; k < S.nnzPu; k += kST) {
                const uint32_t ij = __ldg(S.Pu_ij + k);
                gP[k] = (gP[k] * gT[ij >> 16]) * gT[ij & 0xFFFFu];
            }
            __syncthreads();
            double psum = 0.0, qn[1] = {0.0};
            for (int p = tid; p < np; p += kST) {
                const double dt = gT[p];
                const uint32_t rg = __ldg(S.col_range + p);
                for (int k = (int)(rg >> 8), k1 = k + (int)(rg & 0xFFu); k < k1; ++k) gA[k] *= dt;
                gD[p] *= dt;
                if (hasq) {
                    gq[p] *= dt;
                    qn[0] = fmax(qn[0], fabs(gq[p]));
                }
                psum += pf_absmax_g(S, p, gP);
            }
            psum = block_sum(psum, sRed, tid);
            block_max<1>(qn, sRed, tid);
            double ct = fmax(psum / (double)S.n, limit_scaling(qn[0]));
            ct = 1.0 / limit_scaling(ct);
            for (int k = tid; k < S.nnzPu; k += kST) gP[k] *= ct;
            if (hasq)
                for (int p = tid; p < np; p += kST) gq[p] *= ct;
            cscale *= ct;
            __syncthreads();
        }
        for (int i = tid; i < mu; i += kST) {
            gLo[i] *= gE[i];
            gHi[i] *= gE[i];
        }
        const double cinv = 1.0 / cscale;
        const bool uns = !st.scaled_termination;

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
            for (int i = tid; i < mu; i += kST) gz[i] = row_dot_g(S, i, gA, gx);
        }
        __syncthreads();

        int status = empty_infeasible ? -3 : -10;
        int iters_done = 0;
        double pri_res = 0.0, dua_res = 0.0;
        bool need_factor = true;

#pragma unroll 1
        for (int iter = 1; status == -10 && iter <= st.max_iter; ++iter) {
            if (need_factor) {
                // ---------------- rho per row, assembly of S, two-way block factorisation
                const double rho_eq = kRhoEqOverIneq * rho;
                for (int i = tid; i < mu; i += kST) {
                    const double lo = gLo[i], hi = gHi[i];
                    double rv = rho;
                    if (lo < -kInfty * kMinScaling && hi > kInfty * kMinScaling) rv = kRhoMin;
                    else if (hi - lo < kRhoTol) rv = rho_eq;
                    gRho[i] = rv;
                }
                __syncthreads();
                {
                    const int dpb = bs * (bs + 1) / 2, ndiag = T * dpb;
                    for (int e = tid; e < ndiag; e += kST) {
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
                    for (int e = tid; e < S.noff; e += kST) {
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
                }
                __syncthreads();
                if (tid < 2 * kGroup) {
                    const int g = tid / kGroup, gl_ = tid - g * kGroup;
                    double *pub = sPub + g * 128;
                    if (g == 0) {
#pragma unroll 1
                        for (int t = 0; t < h; ++t) factor_block_stream(S, gDinv, gOff, pub, t, t > 0 ? 1 : 0, 0, gl_);
                    } else {
#pragma unroll 1
                        for (int t = T - 1; t > h; --t) factor_block_stream(S, gDinv, gOff, pub, t, t < T - 1 ? 2 : 0, 1, gl_);
                    }
                }
                __syncthreads();
                if (tid < kGroup) factor_block_stream(S, gDinv, gOff, sPub, h, (h > 0 ? 1 : 0) | (h < T - 1 ? 2 : 0), 0, tid);
                fence_proxy_async_all(); // the factor is read by bulk copies (async proxy) from here on
                __syncthreads();
                need_factor = false;
            }
            // ---------------- rhs = sigma x - q + A'(rho z - y)
            for (int i = tid; i < mu; i += kST) gW[i] = fma(gRho[i], gz[i], -gy[i]);
            __syncthreads();
            for (int p = tid; p < np; p += kST) {
                double acc = st.sigma * gx[p];
                if (hasq) acc -= gq[p];
                sRHS[p] = col_dot_g(S, p, gA, gW, acc);
            }
            __syncthreads();

            // ---------------- block-tridiagonal solve in place on sRHS (warps 0 and 1)
            if (warp < 2) {
                const int nsteps = warp == 0 ? 2 * h + 1 : 2 * (T - 1 - h);
                const bool two_chains = T - 1 - h > 0; // T <= 2: the bottom chain is empty, warp 1 has nothing to wait for
                double *ring = sRing + (size_t)warp * kStreamRing * slotD;
                uint64_t *full = sFull + warp * kStreamRing;
                double *vb = sV + warp * 64;
                auto issue = [&](int j) { // lane 0: start the bulk copies of step j
                    const Step sp = step_of(warp, j, T, h);
                    const uint32_t slot = (ring_k + (uint32_t)j) % kStreamRing;
                    double *dst = ring + (size_t)slot * slotD;
                    const uint32_t bytesD = (uint32_t)bb * 8u, bytesO = (uint32_t)(bs * S.ow) * 8u;
                    const bool has_off = sp.ob >= 0 && sp.ob < T - 1;
                    mbar_expect_tx(full + slot, bytesD + (has_off ? bytesO : 0u));
                    bulk_g2s(dst, gDinv + (size_t)sp.t * bb, bytesD, full + slot);
                    if (has_off) bulk_g2s(dst + bb, gOff + (size_t)sp.ob * bs * S.ow, bytesO, full + slot);
                };
                if (lane == 0) {
                    fence_proxy_async();
                    for (int j = 0; j < kStreamRing - 1 && j < nsteps; ++j) issue(j);
                }
#pragma unroll 1
                for (int j = 0; j < nsteps; ++j) {
                    const Step sp = step_of(warp, j, T, h);
                    if (warp == 0 && j == h && two_chains) named_bar_sync(1, 64); // middle stage: needs the bottom chain's forward sweep
                    if (warp == 1 && j == T - 1 - h) {
                        named_bar_sync(1, 64);
                        named_bar_sync(2, 64);                          // ... and the bottom chain's way back needs x_h
                    }
                    // the slot consumed one step ago is free: refill it kRing-1 steps ahead
                    if (lane == 0 && j + kStreamRing - 1 < nsteps) {
                        fence_proxy_async();
                        issue(j + kStreamRing - 1);
                    }
                    // off-diagonal pattern of this step (does not depend on the chain)
                    const bool from_above = sp.kind == 0 || sp.kind == 2 || sp.kind == 4;
                    const bool from_below = sp.kind == 1 || sp.kind == 2 || sp.kind == 3;
                    const bool fwd = sp.kind <= 2;
                    int ia[kMaxW], ib[kMaxW], sb[kMaxW];
#pragma unroll
                    for (int s = 0; s < kMaxW; ++s) {
                        ia[s] = -1;
                        ib[s] = -1;
                        sb[s] = 0;
                        if (from_above && sp.t > 0 && s < S.ow) ia[s] = __ldg(S.off_col + ((size_t)(sp.t - 1) * bs + r) * S.ow + s);
                        if (from_below && sp.t < T - 1 && s < S.owT) {
                            ib[s] = __ldg(S.offT_row + ((size_t)sp.t * bs + r) * S.owT + s);
                            sb[s] = __ldg(S.offT_slot + ((size_t)sp.t * bs + r) * S.owT + s);
                        }
                    }
                    const uint32_t kk = ring_k + (uint32_t)j, slot = kk % kStreamRing, parity = (kk / kStreamRing) & 1u;
                    {
                        unsigned spins = 0;
                        while (!mbar_try_wait(full + slot, parity)) {
                            if (++spins > (1u << 26)) __trap();
                        }
                    }
                    const double *Dm = ring + (size_t)slot * slotD, *Ob = Dm + bb;
                    double a = fwd ? sRHS[sp.t * bs + r] : 0.0;
                    if (from_above && sp.t > 0) {
                        // rows of stage t against the stage above: Off_{t-1}[r][s] * v_{t-1}[col]
                        const double *Oa = Ob; // streamed block is Off_{t-1} for kinds 0, 2, 4
                        const double *src = sRHS + (sp.t - 1) * bs;
#pragma unroll
                        for (int s = 0; s < kMaxW; ++s)
                            if (ia[s] >= 0) a = fma(-Oa[r * S.ow + s], src[ia[s]], a);
                        for (int s = kMaxW; s < S.ow; ++s) {
                            const int cidx = __ldg(S.off_col + ((size_t)(sp.t - 1) * bs + r) * S.ow + s);
                            if (cidx >= 0) a = fma(-Oa[r * S.ow + s], src[cidx], a);
                        }
                    }
                    if (from_below && sp.t < T - 1) {
                        // columns of Off_t against the stage below: Off_t[row][slot]' * v_{t+1}[row]
                        const double *Ot = sp.kind == 2 ? gOff + (size_t)sp.t * bs * S.ow : Ob; // the middle stage streams Off_{h-1}
                        const double *src = sRHS + (sp.t + 1) * bs;
#pragma unroll
                        for (int s = 0; s < kMaxW; ++s)
                            if (ib[s] >= 0) a = fma(-Ot[sb[s]], src[ib[s]], a);
                        for (int s = kMaxW; s < S.owT; ++s) {
                            const int ridx = __ldg(S.offT_row + ((size_t)sp.t * bs + r) * S.owT + s);
                            if (ridx >= 0) a = fma(-Ot[__ldg(S.offT_slot + ((size_t)sp.t * bs + r) * S.owT + s)], src[ridx], a);
                        }
                    }
                    double *vbj = vb + (j & 1) * 32;
                    if (half == 0 && rowact) vbj[r] = a;
                    __syncwarp();
                    double acc = 0.0;
                    for (int c = c0; c < c1; ++c) acc = fma(Dm[r * bs + c], vbj[c], acc);
                    if (halves == 2) acc += __shfl_xor_sync(0xffffffffu, acc, 16);
                    if (half == 0 && rowact) sRHS[sp.t * bs + r] = fwd ? acc : sRHS[sp.t * bs + r] + acc;
                    __syncwarp();
                    if (warp == 0 && j == h && two_chains) named_bar_sync(2, 64); // x_h published for the bottom chain
                }
                ring_k += (uint32_t)nsteps;
            }
            __syncthreads();

            const bool last = iter == st.max_iter;
            const bool can_check = st.check_termination && (iter % st.check_termination == 0);
            const bool adapt_now = st.adaptive_rho && st.adaptive_rho_interval && (iter % st.adaptive_rho_interval == 0);
            const bool do_info = can_check || adapt_now || last;
            // ---------------- x, z, y updates
            for (int p = tid; p < np; p += kST) {
                const double xo = gx[p];
                const double xn = st.alpha * sRHS[p] + (1.0 - st.alpha) * xo;
                if (do_info && hasq) gdx[p] = xn - xo;
                gx[p] = xn;
            }
            for (int i = tid; i < mu; i += kST) {
                const double zt = row_dot_g(S, i, gA, sRHS);
                const double rv = gRho[i], zo = gz[i], yo_ = gy[i];
                const double zr = st.alpha * zt + (1.0 - st.alpha) * zo;
                const double zn = fmin(fmax(zr + (1.0 / rv) * yo_, gLo[i]), gHi[i]);
                const double d = rv * (zr - zn);
                gy[i] = yo_ + d;
                gz[i] = zn;
                if (do_info) gdy[i] = d;
            }
            __syncthreads();

            if (do_info) {
                // ---------------- residuals and the norms OSQP's tolerances need (update_info)
                double pri_u, pri_s, pn_u, pn_s, dua_u, dua_s, dn_u, dn_s;
                {
                    double na[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                    for (int i = tid; i < mu; i += kST) {
                        const double ax = row_dot_g(S, i, gA, gx);
                        const double zi = gz[i];
                        const double pr = ax - zi;
                        const double einv = 1.0 / gE[i];
                        na[0] = fmax(na[0], fabs(einv * pr));
                        na[1] = fmax(na[1], fabs(pr));
                        na[2] = fmax(na[2], fabs(einv * zi));
                        na[3] = fmax(na[3], fabs(einv * ax));
                        na[4] = fmax(na[4], fabs(zi));
                        na[5] = fmax(na[5], fabs(ax));
                    }
                    block_max<6>(na, sRed, tid);
                    pri_u = na[0]; pri_s = na[1]; pn_u = fmax(na[2], na[3]); pn_s = fmax(na[4], na[5]);
                }
                {
                    double nb[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                    for (int p = tid; p < np; p += kST) {
                        const double px = pf_dot_g(S, p, gP, gx);
                        const double aty = col_dot_g(S, p, gA, gy, 0.0);
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
                    block_max<8>(nb, sRed, tid);
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
                        block_max<1>(nrm, sRed, tid);
                        lhs = block_sum(lhs, sRed, tid);
                        if (nrm[0] > kDivTol && lhs < -epi * nrm[0]) {
                            double v[1] = {0.0};
                            for (int p = tid; p < np; p += kST) {
                                const double acc = col_dot_g(S, p, gA, gW, 0.0);
                                v[0] = fmax(v[0], fabs(uns ? (1.0 / gD[p]) * acc : acc));
                            }
                            block_max<1>(v, sRed, tid);
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
                        block_max<1>(nrm, sRed, tid);
                        qdx = block_sum(qdx, sRed, tid);
                        const double cs_ = uns ? cscale : 1.0;
                        if (nrm[0] > kDivTol && qdx < -cs_ * edi * nrm[0]) {
                            double v[1] = {0.0};
                            for (int p = tid; p < np; p += kST) {
                                const double acc = pf_dot_g(S, p, gP, gdx);
                                v[0] = fmax(v[0], fabs(uns ? (1.0 / gD[p]) * acc : acc));
                            }
                            block_max<1>(v, sRed, tid);
                            if (v[0] < cs_ * edi * nrm[0]) {
                                double viol[1] = {0.0};
                                for (int i = tid; i < mu; i += kST) {
                                    double acc = row_dot_g(S, i, gA, gdx);
                                    if (uns) acc *= 1.0 / gE[i];
                                    if ((gHi[i] < kInfty * kMinScaling && acc > edi * nrm[0]) ||
                                        (gLo[i] > -kInfty * kMinScaling && acc < -edi * nrm[0]))
                                        viol[0] = 1.0;
                                }
                                block_max<1>(viol, sRed, tid);
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
    }
}

cudaError_t prepare_stream(size_t smem_bytes, int *ctas_per_sm) {
    cudaError_t e = cudaFuncSetAttribute(gomp_admm_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, gomp_admm_stream_kernel, kST, smem_bytes);
}
cudaError_t launch_stream(const DevSym &S, const DevBatch &B, const DevSettings &st, const DevStream &X, int grid, size_t smem_bytes,
                          cudaStream_t stream) {
    gomp_admm_stream_kernel<<<grid, kST, smem_bytes, stream>>>(S, B, st, X);
    return cudaGetLastError();
}

} // namespace gomp_dev
