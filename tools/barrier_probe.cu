// barrier_probe.cu -- cost of one named barrier (bar.sync id, n) per iteration for n = 64..448 threads,
// and of a publish->barrier->load round trip, on the target GPU.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(long long *cyc, double *out, int nthreads, int reps) {
    __shared__ double buf[512];
    const int tid = threadIdx.x;
    buf[tid] = tid;
    __syncthreads();
    if (tid >= nthreads) return;
    double v = tid;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        buf[tid] = v;
        asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
        v = buf[(tid + 33) % nthreads] + 1.0;
    }
    long long t1 = clock64();
    if (tid == 0) cyc[0] = t1 - t0;
    out[tid] = v;
}
int main() {
    long long *cyc; double *out;
    cudaMalloc(&cyc, 8); cudaMalloc(&out, 512 * 8);
    for (int n : {32, 64, 128, 224, 256, 448, 512}) {
        probe<<<1, 512>>>(cyc, out, n, 1000);
        probe<<<1, 512>>>(cyc, out, n, 1000);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("bar.sync over %3d threads: %.1f cycles per STS+barrier+LDS round (%s)\n", n, c / 1000.0, cudaGetErrorString(cudaGetLastError()));
    }
}
