// Instantiations + runtime dispatch of the batched DMMA GEMM (see sla_gemm.cuh).
#include "sla_gemm.cuh"
#include "sla_kernels.h"

namespace sla {

template <typename T, int OPA, int OPB> static cudaError_t gemm_pick_tile(const GemmArgs<T>& p, cudaStream_t s) {
    if constexpr (is_cplx<T>::value) {
        if (p.M > 64 && p.N > 32) return launch_gemm_cfg<T, OPA, OPB, 128, 64, 32, 32, 8, 3>(p, s);
        return launch_gemm_cfg<T, OPA, OPB, 64, 64, 32, 32, 8, 3>(p, s);
    } else {
        // 128x128 tiles waste (ceil(M/128)*128/M)^2; prefer 64x64 when that waste is large
        long long t128 = (long long)((p.M + 127) / 128) * ((p.N + 127) / 128) * 128 * 128;
        long long t64 = (long long)((p.M + 63) / 64) * ((p.N + 63) / 64) * 64 * 64;
        if (p.M > 64 && p.N > 64 && t128 * 10 <= t64 * 12)
            return launch_gemm_cfg<T, OPA, OPB, 128, 128, 64, 32, 16, 3>(p, s);
        return launch_gemm_cfg<T, OPA, OPB, 64, 64, 32, 32, 16, 3>(p, s);
    }
}

template <typename T> cudaError_t gemm_batched(int opa, int opb, const GemmArgs<T>& p, cudaStream_t s) {
    if (opa == OP_N && opb == OP_N) return gemm_pick_tile<T, OP_N, OP_N>(p, s);
    if (opa == OP_C && opb == OP_N) return gemm_pick_tile<T, OP_C, OP_N>(p, s);
    if (opa == OP_N && opb == OP_C) return gemm_pick_tile<T, OP_N, OP_C>(p, s);
    return gemm_pick_tile<T, OP_C, OP_C>(p, s);
}

template cudaError_t gemm_batched<double>(int, int, const GemmArgs<double>&, cudaStream_t);
template cudaError_t gemm_batched<cplx>(int, int, const GemmArgs<cplx>&, cudaStream_t);

}  // namespace sla
