// latency_probe.cu -- dependent-issue latencies on the target GPU (cycles): DFMA, DADD, DMUL, DDIV, FFMA,
// LDS (pointer chase), LDS.128 broadcast + DFMA, SHFL (64-bit), STS->syncwarp->LDS round trip, bar.sync.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latency_probe tools/latency_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

#define N 4096
__global__ void probe(double *out, long long *cyc, int *chase_init) {
    __shared__ int chase[1024];
    __shared__ __align__(16) double buf[1024];
    const int lane = threadIdx.x;
    for (int i = lane; i < 1024; i += 32) { chase[i] = chase_init[i]; buf[i] = 1.0 + i * 1e-9; }
    __syncwarp();
    double a = 1.0 + lane * 1e-9, b = 1.0000001, c = 1e-9;
    long long t0, t1;
    int k = 0;
    // DFMA
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = fma(a, b, c);
    t1 = clock64(); if (lane == 0) cyc[k] = t1 - t0; k++;
    // DADD
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = a + c;
    t1 = clock64(); if (lane == 0) cyc[k] = t1 - t0; k++;
    // DMUL
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = a * b;
    t1 = clock64(); if (lane == 0) cyc[k] = t1 - t0; k++;
    // DDIV
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N / 16; ++i) a = 1.0 / a + 1.5;
    t1 = clock64(); if (lane == 0) cyc[k] = (t1 - t0) * 16; k++;
    // FFMA
    float fa = 1.0f + lane, fb = 1.0001f, fc = 1e-3f;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) fa = fmaf(fa, fb, fc);
    t1 = clock64(); if (lane == 0) cyc[k] = t1 - t0; k++;
    // LDS pointer chase
    int p = lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) p = chase[p];
    t1 = clock64(); if (lane == 0) cyc[k] = t1 - t0; k++;
    // SHFL 64-bit dependent
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = __shfl_sync(0xffffffffu, a, (lane + 1) & 31);
    t1 = clock64(); if (lane == 0) cyc[k] = t1 - t0; k++;
    // STS -> syncwarp -> LDS (other lane) round trip
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) { buf[lane] = a; __syncwarp(); a = buf[(lane + 1) & 31] + 0.0; __syncwarp(); }
    t1 = clock64(); if (lane == 0) cyc[k] = t1 - t0; k++;
    // 14x14 GEMV step as in the solver: STS v, syncwarp, 7 LDS.128 row + 7 LDS.128 bcast, 14 DFMA (4 acc), syncwarp
    const double *row = buf + 64 + (lane % 14) * 14;
    t0 = clock64();
    for (int i = 0; i < N / 8; ++i) {
        if (lane < 14) buf[lane] = a;
        __syncwarp();
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
        for (int cc = 0; cc < 14; cc += 2) {
            const double2 m = *reinterpret_cast<const double2 *>(row + cc);
            const double2 w = *reinterpret_cast<const double2 *>(buf + cc);
            if ((cc >> 1) & 1) { a2 = fma(m.x, w.x, a2); a3 = fma(m.y, w.y, a3); } else { a0 = fma(m.x, w.x, a0); a1 = fma(m.y, w.y, a1); }
        }
        a = ((a0 + a1) + (a2 + a3)) * 1e-3;
        __syncwarp();
    }
    t1 = clock64(); if (lane == 0) cyc[k] = (t1 - t0) * 8; k++;
    // bar.sync of one warp-pair sized named barrier is not measurable here; plain __syncwarp cost
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) __syncwarp();
    t1 = clock64(); if (lane == 0) cyc[k] = t1 - t0; k++;
    out[lane] = a + fa + p;
}

int main() {
    double *out; long long *cyc; int *ci;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 16 * 8); cudaMalloc(&ci, 1024 * 4);
    int h[1024];
    for (int i = 0; i < 1024; ++i) h[i] = (i * 37 + 11) & 1023;
    cudaMemcpy(ci, h, sizeof h, cudaMemcpyHostToDevice);
    probe<<<1, 32>>>(out, cyc, ci);
    probe<<<1, 32>>>(out, cyc, ci);
    cudaDeviceSynchronize();
    long long c[16];
    cudaMemcpy(c, cyc, sizeof c, cudaMemcpyDeviceToHost);
    const char *names[] = {"DFMA", "DADD", "DMUL", "DDIV(+add)", "FFMA", "LDS chase", "SHFL.f64", "STS-sync-LDS-sync", "GEMV14 step", "syncwarp"};
    for (int i = 0; i < 10; ++i) printf("%-20s %8.1f cycles/op\n", names[i], (double)c[i] / N);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
