// chain_probe.cu -- cycles per block-substitution step (14x14 block, one warp) for several formulations.
#include <cstdio>
#include <cuda_runtime.h>
#define NB 14
#define T 24
__device__ __forceinline__ void sync() { __syncwarp(); }

// rows of Dinv in smem D[t][r][c]; off values O[t][r][2]; indices oi0/oi1 per lane; rhs in smem
template <int V>
__global__ void probe(long long *cyc, double *out, int reps) {
    __shared__ __align__(16) double D[T * NB * NB];
    __shared__ __align__(16) double O[T * NB * 2];
    __shared__ __align__(16) double rhs[T * NB];
    __shared__ __align__(16) double vb[64];
    const int lane = threadIdx.x;
    for (int i = lane; i < T * NB * NB; i += 32) D[i] = ((i % (NB + 1)) == 0 ? 0.5 : 0.001) * (1 + (i % 7) * 1e-3);
    for (int i = lane; i < T * NB * 2; i += 32) O[i] = 0.01;
    for (int i = lane; i < T * NB; i += 32) rhs[i] = 1.0 + i * 1e-3;
    __syncwarp();
    const int rl = lane < NB ? lane : NB - 1;
    const int oi0 = rl, oi1 = (rl + 7) % NB;
    long long t0 = clock64();
    double u = 0.0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int t = 0; t < T; ++t) {
            if (V == 0) { // smem gather, 2 syncs (version B)
                const double *row = D + (t * NB + rl) * NB;
                double b = rhs[t * NB + rl];
                double off = 0.0;
                if (t > 0) off = fma(O[(t * NB + rl) * 2], rhs[(t - 1) * NB + oi0], O[(t * NB + rl) * 2 + 1] * rhs[(t - 1) * NB + oi1]);
                double v = b - off;
                if (lane < NB) vb[lane] = v;
                sync();
                double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
                for (int c = 0; c < NB; c += 2) {
                    double2 m = *(const double2 *)(row + c), w = *(const double2 *)(vb + c);
                    if ((c >> 1) & 1) { a2 = fma(m.x, w.x, a2); a3 = fma(m.y, w.y, a3); } else { a0 = fma(m.x, w.x, a0); a1 = fma(m.y, w.y, a1); }
                }
                u = (a0 + a1) + (a2 + a3);
                if (lane < NB) rhs[t * NB + lane] = u;
                sync();
            } else if (V == 1) { // shfl gather, double-buffered vb, 1 sync (version A)
                const double *row = D + (t * NB + rl) * NB;
                double m[NB];
#pragma unroll
                for (int c = 0; c < NB; c += 2) { double2 d = *(const double2 *)(row + c); m[c] = d.x; m[c + 1] = d.y; }
                double b = rhs[t * NB + rl];
                double ov0 = O[(t * NB + rl) * 2], ov1 = O[(t * NB + rl) * 2 + 1];
                double up0 = __shfl_sync(~0u, u, oi0), up1 = __shfl_sync(~0u, u, oi1);
                double v = b - fma(ov0, up0, ov1 * up1);
                double *vv = vb + (t & 1) * 16;
                if (lane < NB) vv[lane] = v;
                sync();
                double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
                for (int c = 0; c < NB; c += 2) {
                    double2 w = *(const double2 *)(vv + c);
                    if ((c >> 1) & 1) { a2 = fma(m[c], w.x, a2); a3 = fma(m[c + 1], w.y, a3); } else { a0 = fma(m[c], w.x, a0); a1 = fma(m[c + 1], w.y, a1); }
                }
                u = (a0 + a1) + (a2 + a3);
                if (lane < NB) rhs[t * NB + lane] = u;
            } else if (V == 2) { // pure GEMV chain via smem broadcast (no off-diagonal): floor of the broadcast scheme
                const double *row = D + (t * NB + rl) * NB;
                double *vv = vb + (t & 1) * 16;
                if (lane < NB) vv[lane] = u + 1.0;
                sync();
                double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
                for (int c = 0; c < NB; c += 2) {
                    double2 m = *(const double2 *)(row + c), w = *(const double2 *)(vv + c);
                    if ((c >> 1) & 1) { a2 = fma(m.x, w.x, a2); a3 = fma(m.y, w.y, a3); } else { a0 = fma(m.x, w.x, a0); a1 = fma(m.y, w.y, a1); }
                }
                u = (a0 + a1) + (a2 + a3);
            } else if (V == 3) { // pure GEMV chain via 14 shuffles
                const double *row = D + (t * NB + rl) * NB;
                double v = u + 1.0, a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
                for (int c = 0; c < NB; c += 2) {
                    double2 m = *(const double2 *)(row + c);
                    double w0 = __shfl_sync(~0u, v, c), w1 = __shfl_sync(~0u, v, c + 1);
                    if ((c >> 1) & 1) { a2 = fma(m.x, w0, a2); a3 = fma(m.y, w1, a3); } else { a0 = fma(m.x, w0, a0); a1 = fma(m.y, w1, a1); }
                }
                u = (a0 + a1) + (a2 + a3);
            } else if (V == 4) { // 2 lanes per row: lane = r + 16*half, 7 columns each, shfl_xor combine; smem broadcast
                const int r = lane & 15, half = lane >> 4;
                const int rr = r < NB ? r : NB - 1;
                const double *row = D + (t * NB + rr) * NB + half * 7;
                double *vv = vb + (t & 1) * 16;
                if (lane < NB) vv[lane] = u + 1.0;
                sync();
                double a0 = 0, a1 = 0;
#pragma unroll
                for (int c = 0; c < 7; ++c) {
                    double m = row[c], w = vv[half * 7 + c];
                    if (c & 1) a1 = fma(m, w, a1); else a0 = fma(m, w, a0);
                }
                double s = a0 + a1;
                s += __shfl_xor_sync(~0u, s, 16);
                u = s;
            } else if (V == 5) { // version B but rows loaded BEFORE the dependent gather (explicit order)
                const double *row = D + (t * NB + rl) * NB;
                double m[NB];
#pragma unroll
                for (int c = 0; c < NB; c += 2) { double2 d = *(const double2 *)(row + c); m[c] = d.x; m[c + 1] = d.y; }
                double b = rhs[t * NB + rl];
                double ov0 = O[(t * NB + rl) * 2], ov1 = O[(t * NB + rl) * 2 + 1];
                double off = 0.0;
                if (t > 0) off = fma(ov0, rhs[(t - 1) * NB + oi0], ov1 * rhs[(t - 1) * NB + oi1]);
                double v = b - off;
                double *vv = vb + (t & 1) * 16;
                if (lane < NB) vv[lane] = v;
                sync();
                double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
                for (int c = 0; c < NB; c += 2) {
                    double2 w = *(const double2 *)(vv + c);
                    if ((c >> 1) & 1) { a2 = fma(m[c], w.x, a2); a3 = fma(m[c + 1], w.y, a3); } else { a0 = fma(m[c], w.x, a0); a1 = fma(m[c + 1], w.y, a1); }
                }
                u = (a0 + a1) + (a2 + a3);
                if (lane < NB) rhs[t * NB + lane] = u;
                sync();
            }
        }
    }
    long long t1 = clock64();
    if (lane == 0) cyc[V] = t1 - t0;
    out[lane + 32 * V] = u + rhs[lane];
}

int main() {
    long long *cyc; double *out;
    cudaMalloc(&cyc, 64); cudaMalloc(&out, 8 * 32 * 8);
    const int reps = 200;
    probe<0><<<1, 32>>>(cyc, out, reps); probe<1><<<1, 32>>>(cyc, out, reps); probe<2><<<1, 32>>>(cyc, out, reps);
    probe<3><<<1, 32>>>(cyc, out, reps); probe<4><<<1, 32>>>(cyc, out, reps); probe<5><<<1, 32>>>(cyc, out, reps);
    cudaDeviceSynchronize();
    probe<0><<<1, 32>>>(cyc, out, reps); probe<1><<<1, 32>>>(cyc, out, reps); probe<2><<<1, 32>>>(cyc, out, reps);
    probe<3><<<1, 32>>>(cyc, out, reps); probe<4><<<1, 32>>>(cyc, out, reps); probe<5><<<1, 32>>>(cyc, out, reps);
    cudaDeviceSynchronize();
    long long c[8]; cudaMemcpy(c, cyc, 64, cudaMemcpyDeviceToHost);
    const char *nm[] = {"B: smem gather, 2 syncs", "A: shfl gather, 1 sync", "GEMV only, smem bcast", "GEMV only, 14 shfl", "GEMV only, 2 lanes/row", "B': rows first, smem gather"};
    for (int i = 0; i < 6; ++i) printf("%-32s %7.1f cycles/step\n", nm[i], (double)c[i] / (reps * T));
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
