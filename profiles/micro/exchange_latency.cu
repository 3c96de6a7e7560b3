// Microbenchmark: cost of one all-to-all candidate exchange round in a thread-block cluster of 4 on B200.
//   mode 0: DSMEM bulk copy (cp.async.bulk.shared::cluster) + mbarrier complete_tx wait   (what sla_ldr_reg.cu uses)
//   mode 1: per-element remote stores (st.shared::cluster via map_shared_rank) + cluster.sync()
//   mode 2: cluster.sync() only
//   mode 3: __syncthreads() only
//   mode 4: per-element remote stores + remote mbarrier.arrive.release.cluster / try_wait.acquire.cluster
//   mode 5: st.async (remote store that completes bytes on the destination CTA's mbarrier): no staging, no fence
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exchange_latency exchange_latency.cu ; run: ./exchange_latency
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
constexpr int CS = 4, VLEN = 264, ROUNDS = 2000;

__device__ __forceinline__ unsigned su32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa(const void* p, unsigned r) { unsigned o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(su32(p)), "r"(r)); return o; }

__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(512, 1) k(int mode, int bytes, long long* out) {
    __shared__ __align__(16) double recv[2][CS][VLEN];
    __shared__ __align__(16) double stage[2][VLEN];
    __shared__ unsigned long long bar[2];
    __shared__ unsigned long long bar2[2];   // arrival-count barriers of mode 4 (CS remote arrivals per phase)
    cg::cluster_group cl = cg::this_cluster();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, rank = cl.block_rank();
    for (int i = tid; i < 2 * VLEN; i += 512) ((double*)stage)[i] = i;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bar[1])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(&bar2[0])), "r"(CS));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(&bar2[1])), "r"(CS));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    cl.sync();
    long long t0 = clock64();
    for (int k = 0; k < ROUNDS; ++k) {
        const int par = k & 1;
        if (mode == 0) {
            if (warp == 1) {   // "winner warp": stage + publish
                for (int r = lane; r < bytes / 8; r += 32) stage[par][r] = k + r;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane < CS) {
                    unsigned dst = mapa(&recv[par][rank][0], lane), b = mapa(&bar[par], lane);
                    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "r"(su32(&stage[par][0])), "r"(bytes), "r"(b) : "memory");
                }
            }
            if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(&bar[par])), "r"(CS * bytes) : "memory");
            unsigned done = 0;
            while (!done) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(su32(&bar[par])), "r"((k >> 1) & 1) : "memory");
        } else if (mode == 1) {
            if (warp == 1) {
                for (int r = lane; r < bytes / 8; r += 32) {
                    double v = k + r;
                    for (int q = 0; q < CS; ++q) cl.map_shared_rank(&recv[par][rank][0], q)[r] = v;
                }
            }
            cl.sync();
        } else if (mode == 2) {
            cl.sync();
        } else if (mode == 3) {
            __syncthreads();
        } else if (mode == 4) {
            if (warp == 1) {
                for (int r = lane; r < bytes / 8; r += 32) {
                    double v = k + r;
                    for (int q = 0; q < CS; ++q) cl.map_shared_rank(&recv[par][rank][0], q)[r] = v;
                }
                __syncwarp();
                if (lane < CS) {
                    unsigned b = mapa(&bar2[par], lane);
                    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(b) : "memory");
                }
            }
            unsigned done = 0;
            while (!done) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(su32(&bar2[par])), "r"((k >> 1) & 1) : "memory");
        } else {
            if (warp == 1) {
                for (int r = lane; r < bytes / 8; r += 32) {
                    double v = k + r;
                    for (int q = 0; q < CS; ++q) {
                        unsigned dst = mapa(&recv[par][rank][r], q), b = mapa(&bar[par], q);
                        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(dst), "l"(__double_as_longlong(v)), "r"(b) : "memory");
                    }
                }
            }
            if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(&bar[par])), "r"(CS * bytes) : "memory");
            unsigned done = 0;
            while (!done) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(su32(&bar[par])), "r"((k >> 1) & 1) : "memory");
        }
    }
    long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / ROUNDS;
    if (recv[0][0][5] == -1.0) out[1] = 1;
}

int main() {
    long long* d; cudaMalloc(&d, 16);
    const char* names[] = {"bulk copy + mbarrier", "remote stores + cluster.sync", "cluster.sync only", "__syncthreads only",
                           "remote stores + remote arrive", "st.async + complete_tx"};
    for (int mode = 0; mode < 6; ++mode)
        for (int bytes : {2048, 256}) {
            k<<<CS, 512>>>(mode, bytes, d);
            cudaError_t e = cudaDeviceSynchronize();
            long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            printf("%-32s %5d B per peer: %6lld cycles/round (%s)\n", names[mode], bytes, h, cudaGetErrorString(e));
            if (mode == 2 || mode == 3) break;
        }
    return 0;
}
