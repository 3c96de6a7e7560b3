// Instantiations + runtime dispatch of the batched DMMA GEMM (see sla_gemm.cuh).
#include <cstdlib>

#include "sla_gemm.cuh"
#include "sla_kernels.h"

namespace sla {

template <typename T, int OPA, int OPB> static cudaError_t gemm_pick_tile(const GemmArgs<T>& p, cudaStream_t s) {
    if constexpr (is_cplx<T>::value) {
        static const int force_tile_c = getenv("SLA_GEMM_TILE") ? atoi(getenv("SLA_GEMM_TILE")) : 0;   // experiments
        if (force_tile_c == 12864) return launch_gemm_cfg<T, OPA, OPB, 128, 64, 32, 32, 8, 3>(p, s);
        return launch_gemm_cfg<T, OPA, OPB, 64, 64, 32, 32, 8, 3>(p, s);
    } else {
        // Measured on B200 (profiles/gemm_tiles_r1.txt): 64x64 CTA tiles with 4 warps (3-4 CTAs resident per SM)
        // beat 128x128 with 8 warps (31.0 vs 27.7 TFLOP/s at N=256, batch 1280); 80x80 tiles avoid the padding
        // waste when the size is a multiple of 80 but not of 64 (N=400).
        static const int force_tile = getenv("SLA_GEMM_TILE") ? atoi(getenv("SLA_GEMM_TILE")) : 0;   // experiments
        if (force_tile == 64) return launch_gemm_cfg<T, OPA, OPB, 64, 64, 32, 32, 16, 3>(p, s);
        if (force_tile == 6432) return launch_gemm_cfg<T, OPA, OPB, 64, 64, 32, 32, 32, 2>(p, s);
        if (force_tile == 6424) return launch_gemm_cfg<T, OPA, OPB, 64, 64, 32, 32, 16, 2>(p, s);
        if (force_tile == 6444) return launch_gemm_cfg<T, OPA, OPB, 64, 64, 32, 32, 16, 4>(p, s);
        if (force_tile == 6484) return launch_gemm_cfg<T, OPA, OPB, 64, 64, 32, 32, 8, 4>(p, s);
        if (force_tile == 80) return launch_gemm_cfg<T, OPA, OPB, 80, 80, 40, 40, 16, 3>(p, s);
        if (force_tile == 12864) return launch_gemm_cfg<T, OPA, OPB, 128, 64, 64, 32, 16, 3>(p, s);
        if (force_tile == 128) return launch_gemm_cfg<T, OPA, OPB, 128, 128, 64, 32, 16, 3>(p, s);
        auto padded = [&](int t) { return (long long)((p.M + t - 1) / t) * ((p.N + t - 1) / t) * t * t; };
        if (padded(80) * 10 < padded(64) * 9) return launch_gemm_cfg<T, OPA, OPB, 80, 80, 40, 40, 16, 3>(p, s);
        // full 64x64x16 tiles run the precomputed-address loader: two stages are enough to cover the L2 latency there
        // (31.4 vs 31.2 TFLOP/s at batch 128, 33.6 vs 33.4 at batch 1280, C2 step -0.1 ms: profiles/gemm_fast_r2.txt)
        if (p.M % 64 == 0 && p.N % 64 == 0 && p.K % 16 == 0) return launch_gemm_cfg<T, OPA, OPB, 64, 64, 32, 32, 16, 2>(p, s);
        return launch_gemm_cfg<T, OPA, OPB, 64, 64, 32, 32, 16, 3>(p, s);
    }
}

template <typename T> cudaError_t gemm_batched(int opa, int opb, const GemmArgs<T>& p, cudaStream_t s) {
    if constexpr (!is_cplx<T>::value) {
        // one A for the whole batch (the group products expK * X): persistent A-stationary TMA kernel where the shape fits
        if (opa == OP_N && opb == OP_N && p.strideA == 0) {
            const cudaError_t e = gemm_shared_a_batched(p, s);
            if (e != cudaErrorNotSupported) return e;
        }
    }
    if (opa == OP_N && opb == OP_N) return gemm_pick_tile<T, OP_N, OP_N>(p, s);
    if (opa == OP_C && opb == OP_N) return gemm_pick_tile<T, OP_C, OP_N>(p, s);
    if (opa == OP_N && opb == OP_C) return gemm_pick_tile<T, OP_N, OP_C>(p, s);
    return gemm_pick_tile<T, OP_C, OP_C>(p, s);
}

template cudaError_t gemm_batched<double>(int, int, const GemmArgs<double>&, cudaStream_t);
template cudaError_t gemm_batched<cplx>(int, int, const GemmArgs<cplx>&, cudaStream_t);

}  // namespace sla
