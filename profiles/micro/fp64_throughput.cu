// FP64 issue throughput on B200 with the register LDR kernel's shape: 512 threads (4 warps per SM sub-partition),
// each warp issuing 32 independent DFMAs (a 4 x 8 register tile) per round.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_throughput fp64_throughput.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512, 1) k(double* out, long long* t, double x, int rounds, int nwarps) {
    double a[4][8], v[8], w[4];
    for (int u = 0; u < 4; ++u) { w[u] = x * (u + 1); for (int i = 0; i < 8; ++i) a[u][i] = x + threadIdx.x + u * 8 + i; }
    for (int i = 0; i < 8; ++i) v[i] = x * 0.001 * (i + 1);
    __syncthreads();
    long long t0 = clock64();
    if ((threadIdx.x >> 5) < nwarps) {
        for (int r = 0; r < rounds; ++r) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[u][i] = fma(-v[i], w[u], a[u][i]);
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int u = 0; u < 4; ++u) for (int i = 0; i < 8; ++i) s += a[u][i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) t[0] = t1 - t0;
}
int main() {
    double* o; long long* t; cudaMalloc(&o, 512 * 8); cudaMalloc(&t, 64);
    const int rounds = 1000;
    for (int nw : {1, 4, 8, 16}) {
        k<<<1, 512>>>(o, t, 1.0, rounds, nw);
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, t, 8, cudaMemcpyDeviceToHost);
        printf("%2d warps x 32 independent DFMA per round: %.1f cycles/round (warp 0)\n", nw, (double)h / rounds);
    }
    return 0;
}
