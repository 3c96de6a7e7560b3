// Dependent-issue latency of FP64 / shuffle instructions on B200 (one warp, clock64 around a dependent chain).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* t, double x, double y) {
    const int N = 4096;
    double a = x + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = fma(a, y, x);
    long long t1 = clock64();
    double b = x + threadIdx.x;
#pragma unroll 16
    for (int i = 0; i < N; ++i) b = b + y;
    long long t2 = clock64();
    double c = x + threadIdx.x;
#pragma unroll 16
    for (int i = 0; i < N; ++i) c = __shfl_xor_sync(0xffffffffu, c, 1) + 0.0 * y;   // shuffle (2x SHFL) + DADD
    long long t3 = clock64();
    float f = (float)x + threadIdx.x;
#pragma unroll 16
    for (int i = 0; i < N; ++i) f = fmaf(f, (float)y, (float)x);
    long long t4 = clock64();
    double d = x + threadIdx.x;
#pragma unroll 16
    for (int i = 0; i < N; ++i) d = __shfl_xor_sync(0xffffffffu, d, 1);
    long long t5 = clock64();
    unsigned long long u = (unsigned long long)threadIdx.x + (unsigned long long)x;
#pragma unroll 16
    for (int i = 0; i < N; ++i) { unsigned long long o = __shfl_xor_sync(0xffffffffu, u, 1); u = o > u ? o + 1 : u + 1; }
    long long t6 = clock64();
    out[threadIdx.x] = a + b + c + f + d + (double)u;
    if (threadIdx.x == 0) { t[0] = (t1 - t0); t[1] = (t2 - t1); t[2] = (t3 - t2); t[3] = (t4 - t3); t[4] = (t5 - t4); t[5] = (t6 - t5); t[6] = N; }
}
int main() {
    double* o; long long* t; cudaMalloc(&o, 32 * 8); cudaMalloc(&t, 64);
    k<<<1, 32>>>(o, t, 1.0, 1.0000001);
    cudaDeviceSynchronize();
    long long h[7]; cudaMemcpy(h, t, 56, cudaMemcpyDeviceToHost);
    const char* n[] = {"DFMA dependent", "DADD dependent", "SHFL(f64)+DADD round", "FFMA dependent", "SHFL(f64) only", "SHFL(u64)+cmp/sel round"};
    for (int i = 0; i < 6; ++i) printf("%-26s %.1f cycles\n", n[i], (double)h[i] / h[6]);
    return 0;
}
