// sweep_probe.cu -- times gomp_dev::chain_sweep<14> (the kernel's own substitution sweep) in isolation:
// 1 warp, and 2 warps (top + bottom chain) as in the kernel.
#include <cstdio>
#include "../osqp-solver_b200/csrc/gomp_kernel.cuh"
using namespace gomp_dev;
constexpr int NB = 14, T = 50;
__global__ void probe(long long *cyc, double *out, int reps, int nwarps, int park) {
    extern __shared__ __align__(16) unsigned char sm[];
    double *D = (double *)sm;                 // T*NB*NB
    double *rhs = D + T * NB * NB;            // T*NB
    double *O = rhs + T * NB;                 // (T-1)*NB*2
    double *vb = O + T * NB * 2;              // 64 per warp
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < T * NB * NB; i += blockDim.x) D[i] = ((i % (NB + 1)) == 0 ? 0.5 : 0.001) * (1 + (i % 7) * 1e-3);
    for (int i = tid; i < T * NB * 2; i += blockDim.x) O[i] = 0.01;
    for (int i = tid; i < T * NB; i += blockDim.x) rhs[i] = 1.0 + i * 1e-3;
    __syncthreads();
    if (warp >= nwarps) {
        if (park) __syncthreads();
        return;
    }
    const uint32_t dinv0 = smem_u32(D), rhs0 = smem_u32(rhs), off0 = smem_u32(O), vb0 = smem_u32(vb) + warp * 256;
    const int rl = lane < NB ? lane : NB - 1;
    const int h = T / 2;
    long long t0 = clock64();
    double acc = 0;
    for (int r = 0; r < reps; ++r) {
        if (warp == 0) {
            acc += chain_sweep<NB>(dinv0, rhs0, off0, vb0, 0, h, +1, -1, T - 1, NB * 2 * 8, true, 0.0, lane, rl, (rl + 7) % NB, rl * 16, rl * 16 + 8, true, true);
            acc += chain_sweep<NB>(dinv0, rhs0, off0, vb0, h - 1, h, -1, 0, T - 1, NB * 2 * 8, false, acc, lane, rl, (rl + 7) % NB, rl * 16, rl * 16 + 8, true, true);
        } else {
            acc += chain_sweep<NB>(dinv0, rhs0, off0, vb0, T - 1, T - 1 - h, -1, 0, T - 1, NB * 2 * 8, true, 0.0, lane, rl, (rl + 7) % NB, rl * 16, rl * 16 + 8, true, true);
            acc += chain_sweep<NB>(dinv0, rhs0, off0, vb0, h + 1, T - 1 - h, +1, -1, T - 1, NB * 2 * 8, false, acc, lane, rl, (rl + 7) % NB, rl * 16, rl * 16 + 8, true, true);
        }
    }
    long long t1 = clock64();
    if (park) __syncthreads();
    if (lane == 0) cyc[warp] = t1 - t0;
    out[tid] = acc;
}
int main() {
    long long *cyc; double *out;
    cudaMalloc(&cyc, 64); cudaMalloc(&out, 8 * 512);
    size_t smem = (T * NB * NB + T * NB + T * NB * 2 + 128) * 8;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int reps = 100;
    for (int nw = 1; nw <= 2; ++nw)
        for (int threads : {64, 512}) {
            for (int park = 0; park < 2; ++park) {
            probe<<<1, threads, smem>>>(cyc, out, reps, nw, park);
            probe<<<1, threads, smem>>>(cyc, out, reps, nw, park);
            cudaDeviceSynchronize();
            long long c[2]; cudaMemcpy(c, cyc, 16, cudaMemcpyDeviceToHost);
            printf("warps active %d, block %3d threads, others parked at barrier %d: warp0 %.1f cycles/step %s\n", nw, threads, park, (double)c[0] / (reps * 50), cudaGetErrorString(cudaGetLastError()));
            }
        }
}
