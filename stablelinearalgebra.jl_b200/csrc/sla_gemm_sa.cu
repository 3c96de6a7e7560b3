// Batched GEMM with ONE A operand shared by the whole batch:  C[b] = diag(r_b) * (A * B[b]) * diag(c_b)   (real FP64).
//
// This is the shape of 80 % of a chain's GEMM flops: the group products X <- diag(e_l) * (expK * X) multiply the SAME
// expK (strideA = 0) into every chain's running product (test/runtests.jl:96-99, `mul!(A, B, B_bar)` in a loop).  The
// generic kernel (sla_gemm.cuh) re-streams its 64 x K panel of A for every 64 x 64 tile of every batch entry through a
// cp.async ring; here the kernel is PERSISTENT and A-STATIONARY:
//   * one CTA per SM; a CTA keeps a 64 x K row panel of A in shared memory for its whole life and walks over the
//     (batch entry, 128-column block) work units of that panel, so only B is streamed (half the operand traffic) and the
//     tail of a 3.46-wave launch of 2048 small CTAs disappears (1024 units over 148 CTAs: 6.92 rounds of 7);
//   * operands arrive by tensor-map TMA (cp.async.bulk.tensor.2d, 128-byte swizzle, SASS: UTMALDG) issued by ONE producer
//     thread into a 5-stage ring; completion is tracked by mbarriers (full / empty per stage), there is no CTA-wide
//     barrier in the main loop and no per-thread address arithmetic for the loads;
//   * 8 consumer warps (2 x 4, each a 32 x 32 register tile of 8 x 8 DMMA.8x8x4 accumulators) read their fragments
//     straight from the swizzled tiles -- the swizzle makes both fragment patterns conflict-free without padding;
//   * the epilogue (diagonal scalings, 16-byte stores) of unit u overlaps the producer's prefetch of unit u + 1.
#include <cuda.h>

#include <mutex>

#include "sla_gemm.cuh"
#include "sla_kernels.h"

namespace sla {

namespace {

constexpr int SA_BM = 64, SA_BN = 128, SA_BK = 16, SA_STAGES = 5;
constexpr int SA_CONSUMERS = 8;                       // warps
constexpr int SA_THREADS = (SA_CONSUMERS + 1) * 32;   // + the producer warp
constexpr int SA_KMAX = 256;
constexpr int SA_STAGE_BYTES = SA_BN * SA_BK * 8;     // 16 KB: box {16 k, 128 n}
constexpr int SA_ABOX_BYTES = 16 * SA_KMAX * 8;       // 32 KB: box {16 m, 256 k}; four of them hold the 64 x K panel
constexpr size_t SA_SMEM = 1024 /* alignment slack */ + 4 * SA_ABOX_BYTES + SA_STAGES * SA_STAGE_BYTES + 256;

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
// one box of a 2-D tensor map -> shared memory; completes (by bytes) on `bar`
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}

// 128-byte swizzle (CU_TENSOR_MAP_SWIZZLE_128B): rows of 128 bytes, the 16-byte chunk index is XORed with (row & 7).
// Element (row, e) of a tile of doubles with 16 doubles per row:
__device__ __forceinline__ int swz(int row, int e) { return row * 16 + ((((e >> 1) ^ row) & 7) << 1) + (e & 1); }

__global__ void __launch_bounds__(SA_THREADS, 1)
gemm_sa_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, GemmArgs<double> p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    double* As = reinterpret_cast<double*>(base);                                  // 4 boxes [k][16 m], swizzled
    double* Bs = reinterpret_cast<double*>(base + 4 * SA_ABOX_BYTES);              // stages [128 n][16 k], swizzled
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(base + 4 * SA_ABOX_BYTES + SA_STAGES * SA_STAGE_BYTES);
    unsigned long long* full = bars;                  // [SA_STAGES]
    unsigned long long* empty = bars + SA_STAGES;     // [SA_STAGES]
    unsigned long long* abar = bars + 2 * SA_STAGES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int mpanels = p.M / SA_BM;
    const int panel = blockIdx.x % mpanels;                      // this CTA's row panel of A, for its whole life
    const int cta_in_panel = blockIdx.x / mpanels, ctas_per_panel = gridDim.x / mpanels;
    const int nblocks = p.N / SA_BN;
    const int units = p.batch * nblocks;                         // (batch entry, column block) pairs of this panel
    const int KT = p.K / SA_BK;
    const int m0 = panel * SA_BM;

    if (tid == 0) {
        for (int s = 0; s < SA_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], SA_CONSUMERS); }
        mbar_init(abar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == SA_CONSUMERS) {
        // ------------------------------------------------------------------ producer: one thread drives the TMA unit
        if (lane == 0) {
            mbar_expect_tx(abar, 4 * 16 * p.K * 8);
            for (int q = 0; q < 4; ++q) tma_load_2d(As + q * (SA_ABOX_BYTES / 8), &mapA, m0 + 16 * q, 0, abar);
            int it = 0;
            for (int u = cta_in_panel; u < units; u += ctas_per_panel) {
                const int b = u / nblocks, n0 = (u % nblocks) * SA_BN;
                const int col0 = b * p.N + n0;                   // column of the (batch * N) x K view of B
                for (int kt = 0; kt < KT; ++kt, ++it) {
                    const int s = it % SA_STAGES;
                    mbar_wait(&empty[s], ((it / SA_STAGES) & 1) ^ 1);
                    mbar_expect_tx(&full[s], SA_STAGE_BYTES);
                    tma_load_2d(Bs + s * (SA_STAGE_BYTES / 8), &mapB, kt * SA_BK, col0, &full[s]);
                }
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers
    const int wm = warp & 1, wn = warp >> 1;           // 2 x 4 warps of 32 x 32
    const int fi = lane >> 2, fk = lane & 3;
    mbar_wait(abar, 0);
    int it = 0;
    for (int u = cta_in_panel; u < units; u += ctas_per_panel) {
        const int b = u / nblocks, n0 = (u % nblocks) * SA_BN;
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
        // scaling multipliers of this lane's rows / columns, loaded before the main loop
        const double* rs = nullptr;
        if (p.rs) rs = (p.rs_div > 0) ? p.rs + (long long)(b / p.rs_div) * p.stride_rs_hi + (long long)(b % p.rs_div) * p.stride_rs
                                      : p.rs + (long long)b * p.stride_rs;
        const double* cs = p.cs ? p.cs + (long long)b * p.stride_cs : nullptr;
        double rsv[4][2], csv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + wm * 32 + j * 8 + 2 * fk;
            rsv[j][0] = rs ? rs[m] : 1.0;
            rsv[j][1] = rs ? rs[m + 1] : 1.0;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) csv[i] = cs ? cs[n0 + wn * 32 + i * 8 + fi] : 1.0;

        for (int kt = 0; kt < KT; ++kt, ++it) {
            const int s = it % SA_STAGES;
            mbar_wait(&full[s], (it / SA_STAGES) & 1);
            const double* sb = Bs + s * (SA_STAGE_BYTES / 8);
            const int kg = kt * SA_BK;
            double nf[2][4], mf[2][4];
            auto load_frags = [&](int buf, int kk) {
#pragma unroll
                for (int i = 0; i < 4; ++i) nf[buf][i] = sb[swz(wn * 32 + i * 8 + fi, kk + fk)];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ml = wm * 32 + j * 8 + fi;           // row of the panel: box ml / 16, element ml % 16
                    mf[buf][j] = As[(ml >> 4) * (SA_ABOX_BYTES / 8) + swz(kg + kk + fk, ml & 15)];
                }
            };
            load_frags(0, 0);
#pragma unroll
            for (int kk = 0; kk < SA_BK; kk += 4) {
                const int cur = (kk >> 2) & 1;
                if (kk + 4 < SA_BK) load_frags(cur ^ 1, kk + 4);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], nf[cur][i], mf[cur][j]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);                 // this warp is done with the stage
        }
        // epilogue: lane owns C[m, m+1][n], m = m0 + wm*32 + j*8 + 2 fk, n = n0 + wn*32 + i*8 + fi
        double* C = p.C + (long long)b * p.strideC;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int n = n0 + wn * 32 + i * 8 + fi;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int m = m0 + wm * 32 + j * 8 + 2 * fk;
                double2* dst = reinterpret_cast<double2*>(C + (long long)n * p.ldc + m);
                double v0 = acc[i][j][0] * rsv[j][0] * csv[i], v1 = acc[i][j][1] * rsv[j][1] * csv[i];
                if (p.accumulate) { const double2 o = *dst; v0 += o.x; v1 += o.y; }
                *dst = make_double2(v0, v1);
            }
        }
    }
}

// ---- tensor maps: encoded on the host through the driver entry point, cached (the chain reuses its buffers every step)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

struct MapKey { const void* ptr; unsigned long long inner, outer, stride; unsigned b0, b1; int dev; };
struct MapEntry { MapKey key; CUtensorMap map; bool used; };
bool same(const MapKey& a, const MapKey& b) {
    return a.ptr == b.ptr && a.inner == b.inner && a.outer == b.outer && a.stride == b.stride && a.b0 == b.b0 && a.b1 == b.b1 && a.dev == b.dev;
}
// 2-D tensor of doubles: `inner` contiguous elements per row, `outer` rows `stride` elements apart; box {b0, b1}
bool get_map(CUtensorMap* out, const void* ptr, unsigned long long inner, unsigned long long outer, unsigned long long stride,
             unsigned b0, unsigned b1) {
    static std::mutex mu;
    static MapEntry cache[32];
    static int next = 0;
    const MapKey key{ptr, inner, outer, stride, b0, b1, current_device()};
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < 32; ++i)
        if (cache[i].used && same(cache[i].key, key)) { *out = cache[i].map; return true; }
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t gdim[2] = {inner, outer};
    const cuuint64_t gstride[1] = {stride * 8};
    const cuuint32_t box[2] = {b0, b1};
    const cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    if (fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    cache[next] = MapEntry{key, m, true};
    next = (next + 1) % 32;
    *out = m;
    return true;
}

}  // namespace

// cudaErrorNotSupported: the shape is not served here (the caller falls back to the generic kernel)
cudaError_t gemm_shared_a_batched(const GemmArgs<double>& p, cudaStream_t stream) {
    // Measured (profiles/gemm_sa_r2.txt): correct, but NOT faster than the generic cp.async kernel at 4 CTAs per SM
    // (28.7 = 28.7 TFLOP/s at batch 128, 24.3 < 26.5 at batch 64 where 128 units over 37 CTAs per panel leave a tail,
    // 30.4 < 30.9 at batch 1280): with one CTA per SM the eight consumer warps run their epilogues in lock step and the
    // DMMA pipe idles meanwhile.  Opt-in (SLA_GEMM_SA=1) until the consumers are split into two ping-pong groups.
    static const bool on = getenv("SLA_GEMM_SA") && atoi(getenv("SLA_GEMM_SA")) != 0;
    if (!on) return cudaErrorNotSupported;
    if (p.strideA != 0 || p.batch < 8) return cudaErrorNotSupported;
    if (p.M % SA_BM || p.N % SA_BN || p.K % SA_BK || p.K > SA_KMAX || p.M > 1024) return cudaErrorNotSupported;
    if (p.lda % 2 || p.ldb != p.K || p.strideB != (long long)p.K * p.N || p.ldc % 2) return cudaErrorNotSupported;   // B[b] dense: one 2-D view
    if ((((uintptr_t)p.A) | ((uintptr_t)p.B) | ((uintptr_t)p.C)) & 15) return cudaErrorNotSupported;
    if (p.strideC % 2) return cudaErrorNotSupported;
    static int sms[kMaxDevices] = {};
    const int dev = current_device();
    if (dev < 0 || dev >= kMaxDevices) return cudaErrorNotSupported;
    if (sms[dev] == 0) {
        if (cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return cudaErrorNotSupported;
        if (cudaFuncSetAttribute(gemm_sa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SA_SMEM) != cudaSuccess) { sms[dev] = 0; return cudaErrorNotSupported; }
    }
    CUtensorMap mapA, mapB;
    // A(m, k) at k * lda + m: inner = m; box {16 m, K k}.   B[b](k, n) at (b * N + n) * K + k: inner = k; box {16 k, 128 n}
    if (!get_map(&mapA, p.A, (unsigned long long)p.M, (unsigned long long)p.K, (unsigned long long)p.lda, 16, (unsigned)p.K)) return cudaErrorNotSupported;
    if (!get_map(&mapB, p.B, (unsigned long long)p.K, (unsigned long long)p.batch * p.N, (unsigned long long)p.K, SA_BK, SA_BN)) return cudaErrorNotSupported;
    const int mpanels = p.M / SA_BM;
    const int units = p.batch * (p.N / SA_BN);
    int per_panel = sms[dev] / mpanels;
    if (per_panel > units) per_panel = units;
    if (per_panel < 1) return cudaErrorNotSupported;
    gemm_sa_kernel<<<mpanels * per_panel, SA_THREADS, SA_SMEM, stream>>>(mapA, mapB, p);
    ++g_launch_count;
    return cudaGetLastError();
}

}  // namespace sla
