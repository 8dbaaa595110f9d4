// tmem_probe.cu -- can tensor memory (TMEM, 256 KB per SM) hold the block factor of the substitution sweeps?
// 1. correctness of tcgen05.alloc / st / ld (32x32b shapes) as plain lane-private storage;
// 2. latency and throughput of tcgen05.ld;
// 3. cycles per stage of four forms of the 14x14 substitution step:
//      V0  the kernel's own chain_sweep<14> (matrix rows from shared memory, two lanes per row)
//      V1  two lanes per row, matrix half-rows + off-diagonal values from TMEM (x16 per stage)
//      V2  one lane per row, two chains per warp, rows from TMEM (x32 per stage)
//      V3  one lane per (q_j, v_j) row pair, four chains per warp, off-diagonal product lane-local (x64 per stage)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_probe tools/tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../osqp-solver_b200/csrc/gomp_kernel.cuh"
#include "../osqp-solver_b200/csrc/gomp_duo.cu"
using namespace gomp_dev;

__device__ __forceinline__ void tmem_ld32(uint32_t ta, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(ta)
        : "memory");
}
__device__ __forceinline__ void tmem_ld64(uint32_t ta, uint32_t (&r)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
        "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
          "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]),
          "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]),
          "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(ta)
        : "memory");
}
__device__ __forceinline__ double mk(uint32_t lo, uint32_t hi) { return __hiloint2double((int)hi, (int)lo); }

constexpr int NB = 14, T = 50, H = 25;

// ---- V1: two lanes per row; lane = r + 16*half; half 0: columns 0..7, half 1: columns 8..13 + the two off-diagonal values
template <bool FWD>
__device__ __forceinline__ double step_v1(const uint32_t (&c)[16], uint32_t ab, uint32_t avb, double prev, int r, int half, bool act, int g0, int g1) {
    const double b = lds_f64(ab);
    const double up0 = __shfl_sync(0xffffffffu, prev, g0), up1 = __shfl_sync(0xffffffffu, prev, g1);
    const double ov0 = mk(c[12], c[13]), ov1 = mk(c[14], c[15]);
    const double off = fma(ov0, up0, ov1 * up1);
    const double v = FWD ? b - off : off;
    if (act) sts_f64(avb + (uint32_t)r * 8u, v);
    __syncwarp();
    const uint32_t aw = avb + (half ? 64u : 0u);
    const double2 w0 = lds_f64x2(aw), w1 = lds_f64x2(aw + 16), w2 = lds_f64x2(aw + 32), w3 = lds_f64x2(aw + 48);
    double a0 = mk(c[0], c[1]) * w0.x, a1 = mk(c[2], c[3]) * w0.y;
    a0 = fma(mk(c[4], c[5]), w1.x, a0);
    a1 = fma(mk(c[6], c[7]), w1.y, a1);
    a0 = fma(mk(c[8], c[9]), w2.x, a0);
    a1 = fma(mk(c[10], c[11]), w2.y, a1);
    if (!half) {
        a0 = fma(mk(c[12], c[13]), w3.x, a0);
        a1 = fma(mk(c[14], c[15]), w3.y, a1);
    }
    double acc = a0 + a1;
    acc += __shfl_xor_sync(0xffffffffu, acc, 16);
    const double res = FWD ? acc : b - acc;
    if (act) sts_f64(ab, res);
    return res;
}
template <bool FWD>
__device__ __noinline__ double sweep_v1(uint32_t tb, uint32_t rhs0, uint32_t vb0, int count, double prev, int lane) {
    const int r = lane & 15, half = lane >> 4, rl = r < NB ? r : NB - 1;
    const bool act = half == 1 && r < NB;
    const int g0 = rl, g1 = (rl + 7) % NB;
    uint32_t ca[16], cb[16];
    tmem_ld16(tb, ca);
    uint32_t ab = rhs0 + rl * 8u;
#pragma unroll 1
    for (int it = 0; it < count; it += 2) {
        tmem_wait_ld();
        tmem_ld16(tb + (it + 1) * 16, cb);
        prev = step_v1<FWD>(ca, ab, vb0, prev, r, half, act, g0, g1);
        ab += NB * 8;
        tmem_wait_ld();
        tmem_ld16(tb + (it + 2) * 16, ca);
        if (it + 1 < count) prev = step_v1<FWD>(cb, ab, vb0 + 128, prev, r, half, act, g0, g1);
        ab += NB * 8;
    }
    tmem_wait_ld();
    return prev;
}

// ---- V2: one lane per row, two chains per warp (lanes 0..13 and 16..29); 14 row entries + 2 off-diagonal values per stage
template <bool FWD>
__device__ __forceinline__ double step_v2(const uint32_t (&c)[32], uint32_t ab, uint32_t avb, double prev, int r, bool act, int g0, int g1) {
    const double b = lds_f64(ab);
    const double up0 = __shfl_sync(0xffffffffu, prev, g0), up1 = __shfl_sync(0xffffffffu, prev, g1);
    const double off = fma(mk(c[28], c[29]), up0, mk(c[30], c[31]) * up1);
    const double v = FWD ? b - off : off;
    if (act) sts_f64(avb + (uint32_t)r * 8u, v);
    __syncwarp();
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const double2 w = lds_f64x2(avb + k * 16);
        if (k & 1) {
            a2 = fma(mk(c[4 * k], c[4 * k + 1]), w.x, a2);
            a3 = fma(mk(c[4 * k + 2], c[4 * k + 3]), w.y, a3);
        } else {
            a0 = fma(mk(c[4 * k], c[4 * k + 1]), w.x, a0);
            a1 = fma(mk(c[4 * k + 2], c[4 * k + 3]), w.y, a1);
        }
    }
    const double acc = (a0 + a1) + (a2 + a3);
    const double res = FWD ? acc : b - acc;
    if (act) sts_f64(ab, res);
    return res;
}
template <bool FWD>
__device__ __noinline__ double sweep_v2(uint32_t tb, uint32_t rhs0, uint32_t vb0, int count, double prev, int lane) {
    const int r = lane & 15, ch = lane >> 4, rl = r < NB ? r : NB - 1;
    const bool act = r < NB;
    const int g0 = ch * 16 + rl, g1 = ch * 16 + (rl + 7) % NB;
    uint32_t ca[32], cb[32];
    tmem_ld32(tb, ca);
    uint32_t ab = rhs0 + (ch * H * NB + rl) * 8u;
    const uint32_t vb = vb0 + ch * 256;
#pragma unroll 1
    for (int it = 0; it < count; it += 2) {
        tmem_wait_ld();
        tmem_ld32(tb + ((it + 1) % 15) * 32, cb);
        prev = step_v2<FWD>(ca, ab, vb, prev, r, act, g0, g1);
        ab += NB * 8;
        tmem_wait_ld();
        tmem_ld32(tb + ((it + 2) % 15) * 32, ca);
        if (it + 1 < count) prev = step_v2<FWD>(cb, ab, vb + 128, prev, r, act, g0, g1);
        ab += NB * 8;
    }
    tmem_wait_ld();
    return prev;
}

// ---- V3: lane j (< 7) of chain ch = lane >> 3 owns rows j and j + 7; off-diagonal product is lane-local
template <bool FWD>
__device__ __forceinline__ void step_v3(const uint32_t (&c)[64], uint32_t ab, uint32_t avb, double &pq, double &pv, int j, bool act) {
    const double bq = lds_f64(ab), bv = lds_f64(ab + 56);
    const double oq = fma(mk(c[56], c[57]), pq, mk(c[58], c[59]) * pv), ov = mk(c[60], c[61]) * pv;
    const double vq = FWD ? bq - oq : oq, vv = FWD ? bv - ov : ov;
    if (act) {
        sts_f64(avb + (uint32_t)j * 8u, vq);
        sts_f64(avb + (uint32_t)(j + 7) * 8u, vv);
    }
    __syncwarp();
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const double2 w = lds_f64x2(avb + k * 16);
        a0 = fma(mk(c[4 * k], c[4 * k + 1]), w.x, a0);
        a1 = fma(mk(c[4 * k + 2], c[4 * k + 3]), w.y, a1);
        a2 = fma(mk(c[28 + 4 * k], c[28 + 4 * k + 1]), w.x, a2);
        a3 = fma(mk(c[28 + 4 * k + 2], c[28 + 4 * k + 3]), w.y, a3);
    }
    const double rq = a0 + a1, rv = a2 + a3;
    pq = FWD ? rq : bq - rq;
    pv = FWD ? rv : bv - rv;
    if (act) {
        sts_f64(ab, pq);
        sts_f64(ab + 56, pv);
    }
}
template <bool FWD>
__device__ __noinline__ double sweep_v3(uint32_t tb, uint32_t rhs0, uint32_t vb0, int count, double prev, int lane) {
    const int j = lane & 7, ch = lane >> 3, jl = j < 7 ? j : 6;
    const bool act = j < 7;
    uint32_t ca[64], cb[64];
    tmem_ld64(tb, ca);
    uint32_t ab = rhs0 + ((ch & 1) * H * NB + jl) * 8u + (ch >> 1) * (T * NB * 8u);
    const uint32_t vb = vb0 + ch * 256;
    double pq = prev, pv = prev;
#pragma unroll 1
    for (int it = 0; it < count; it += 2) {
        tmem_wait_ld();
        tmem_ld64(tb + ((it + 1) & 7) * 64, cb);
        step_v3<FWD>(ca, ab, vb, pq, pv, jl, act);
        ab += NB * 8;
        tmem_wait_ld();
        tmem_ld64(tb + ((it + 2) & 7) * 64, ca);
        if (it + 1 < count) step_v3<FWD>(cb, ab, vb + 128, pq, pv, jl, act);
        ab += NB * 8;
    }
    tmem_wait_ld();
    return pq + pv;
}

__global__ void __launch_bounds__(128, 1) probe(long long *cyc, double *out, int *errs, int reps, int nwarps) {
    extern __shared__ __align__(16) unsigned char sm[];
    __shared__ uint32_t tmem_base_s;
    double *D = (double *)sm;      // T*NB*NB  (V0)
    double *rhs = D + T * NB * NB; // 2*T*NB
    double *O = rhs + 2 * T * NB;  // T*NB*2
    double *vb = O + T * NB * 2;   // 4 warps x 4 chains x 32 doubles
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < T * NB * NB; i += blockDim.x) D[i] = ((i % (NB + 1)) == 0 ? 0.5 : 0.001) * (1 + (i % 7) * 1e-3);
    for (int i = tid; i < T * NB * 2; i += blockDim.x) O[i] = 0.01;
    for (int i = tid; i < 2 * T * NB; i += blockDim.x) rhs[i] = 1.0 + i * 1e-3;
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    const uint32_t tbase = tmem_base_s;
    const uint32_t tl = tbase + ((uint32_t)(warp * 32) << 16); // this warp's lane quarter
    int k = 0;
    // ---- 1. write / read back the whole quarter
    {
        uint32_t v[16];
        for (int c = 0; c < 512; c += 16) {
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = (uint32_t)(tid * 1000 + c + q);
            tmem_st16(tl + c, v);
        }
        tmem_wait_st();
        int bad = 0;
        for (int c = 0; c < 512; c += 16) {
            tmem_ld16(tl + c, v);
            tmem_wait_ld();
#pragma unroll
            for (int q = 0; q < 16; ++q) bad += v[q] != (uint32_t)(tid * 1000 + c + q);
        }
        uint32_t w[64];
        tmem_ld64(tl + 100, w);
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 64; ++q) bad += w[q] != (uint32_t)(tid * 1000 + 100 + q);
        atomicAdd(errs, bad);
        if (tid == 0) errs[1] = (int)tbase;
    }
    __syncthreads();
    // ---- 2. latency (dependent address) and throughput of tcgen05.ld.x16 / x32 / x64, one warp and all warps
    if (warp < nwarps) {
        uint32_t v[16];
        uint32_t a = tl;
        long long t0 = clock64();
        for (int i = 0; i < 1024; ++i) {
            tmem_ld16(a, v);
            tmem_wait_ld();
            a = tl + (v[0] & 0x10u); // data dependent column (0 or 16)
        }
        long long t1 = clock64();
        if (lane == 0) cyc[warp * 16 + k] = t1 - t0;
        out[tid] = (double)a;
        k++;
        uint32_t w[64];
        t0 = clock64();
        for (int i = 0; i < 1024; ++i) {
            tmem_ld64(tl + (i & 7) * 64, w);
            tmem_wait_ld();
            a += w[0] + w[63];
        }
        t1 = clock64();
        if (lane == 0) cyc[warp * 16 + k] = t1 - t0;
        out[tid] += (double)a;
        k++;
    } else k += 2;
    __syncthreads();
    // ---- 3. fill TMEM with a well-scaled "factor": diagonal 0.5, rest 1e-3, off-diagonal values 1e-2
    {
        uint32_t v[16];
        for (int c = 0; c < 512; c += 16) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const double d = 0.001 + 1e-5 * q;
                v[2 * q] = (uint32_t)__double2loint(d);
                v[2 * q + 1] = (uint32_t)__double2hiint(d);
            }
            tmem_st16(tl + c, v);
        }
        tmem_wait_st();
    }
    __syncthreads();
    const uint32_t rhs0 = smem_u32(rhs), vb0 = smem_u32(vb) + warp * 1024;
    double acc = 0;
    if (warp < nwarps) {
        // V0
        const int rl = (lane & 15) < NB ? (lane & 15) : NB - 1;
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            acc += chain_sweep<NB>(smem_u32(D), rhs0, smem_u32(O), vb0, 0, H, +1, -1, T - 1, NB * 2 * 8, true, 0.0, lane, rl, (rl + 7) % NB, rl * 16, rl * 16 + 8, true, true);
            acc += chain_sweep<NB>(smem_u32(D), rhs0, smem_u32(O), vb0, H - 1, H, -1, 0, T - 1, NB * 2 * 8, false, acc, lane, rl, (rl + 7) % NB, rl * 16, rl * 16 + 8, true, true);
        }
        long long t1 = clock64();
        if (lane == 0) cyc[warp * 16 + k] = t1 - t0;
        k++;
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            acc += sweep_v1<true>(tl, rhs0, vb0, H, 0.0, lane);
            acc += sweep_v1<false>(tl, rhs0, vb0, H, acc * 1e-3, lane);
        }
        t1 = clock64();
        if (lane == 0) cyc[warp * 16 + k] = t1 - t0;
        k++;
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            acc += duo_sweep<NB, true>(tl, 0, 16, rhs0 + rl * 8u, NB * 8, smem_u32(O), NB * 2 * 8, -1, +1, T - 1, vb0, H, 0.0, lane, rl, (rl + 7) % NB,
                                       rl * 16, rl * 16 + 8, true, true);
            acc += duo_sweep<NB, false>(tl, (H - 1) * 16, -16, rhs0 + ((H - 1) * NB + rl) * 8u, -NB * 8, smem_u32(O), NB * 2 * 8, H - 1, -1, T - 1, vb0, H,
                                        acc * 1e-3, lane, rl, (rl + 7) % NB, rl * 16, rl * 16 + 8, true, true);
        }
        t1 = clock64();
        if (lane == 0) cyc[warp * 16 + k] = t1 - t0;
        k++;
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            acc += sweep_v2<true>(tl, rhs0, vb0, H, 0.0, lane);
            acc += sweep_v2<false>(tl, rhs0, vb0, H, acc * 1e-3, lane);
        }
        t1 = clock64();
        if (lane == 0) cyc[warp * 16 + k] = t1 - t0;
        k++;
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            acc += sweep_v3<true>(tl, rhs0, vb0, H, 0.0, lane);
            acc += sweep_v3<false>(tl, rhs0, vb0, H, acc * 1e-3, lane);
        }
        t1 = clock64();
        if (lane == 0) cyc[warp * 16 + k] = t1 - t0;
        k++;
    }
    out[tid] += acc;
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

int main() {
    long long *cyc;
    double *out;
    int *errs;
    cudaMalloc(&cyc, 64 * 8);
    cudaMalloc(&out, 8 * 128);
    cudaMalloc(&errs, 8);
    size_t smem = (T * NB * NB + 2 * T * NB + T * NB * 2 + 4 * 128 + 64) * 8;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int reps = 100;
    for (int nw : {1, 2, 4}) {
        cudaMemset(errs, 0, 8);
        cudaMemset(cyc, 0, 64 * 8);
        probe<<<1, 128, smem>>>(cyc, out, errs, reps, nw);
        probe<<<1, 128, smem>>>(cyc, out, errs, reps, nw);
        cudaError_t e = cudaDeviceSynchronize();
        long long c[64];
        int he[2];
        cudaMemcpy(c, cyc, sizeof c, cudaMemcpyDeviceToHost);
        cudaMemcpy(he, errs, 8, cudaMemcpyDeviceToHost);
        printf("== %d warps active: %s, readback mismatches %d, tmem base 0x%x\n", nw, cudaGetErrorString(e), he[0], he[1]);
        for (int w = 0; w < nw; ++w)
            printf("  warp %d: ld.x16 dependent %.1f cyc, ld.x64+wait %.1f cyc | per stage: V0 smem 2 lanes/row %.1f, V1 tmem 2 lanes/row %.1f, "
                   "V2 tmem 1 lane/row (2 chains) %.1f, V3 tmem row pair/lane (4 chains) %.1f, duo_sweep %.1f\n",
                   w, c[w * 16 + 0] / 1024.0, c[w * 16 + 1] / 1024.0, c[w * 16 + 2] / (reps * 50.0), c[w * 16 + 3] / (reps * 50.0),
                   c[w * 16 + 5] / (reps * 50.0), c[w * 16 + 6] / (reps * 50.0), c[w * 16 + 4] / (reps * 50.0));
    }
    return 0;
}
