// Strided-batched column-major GEMM on the FP64 tensor pipe (DMMA.8x8x4) with fused diagonal
// scaling:   C[b] (+)= rowscale(r[b]) * ( op(A[b]) * op(B[b]) ) * colscale(c[b])
//
// This replaces every `mul!(C, A, B)` of the reference (BLAS ?gemm via LinearAlgebra.mul!,
// e.g. src/overloaded_functions.jl:133,141; src/exported_functions.jl:64,146,158,169,170) together
// with the `Diagonal` scalings that surround it (overloaded_functions.jl:134-135,
// exported_functions.jl:41-42,147,155).
//
// Design (B200): operands are staged global->shared with cp.async (LDGSTS, .cg so the streams do
// not pollute L1) through a 3-stage ring; each warp owns a WM x WN register tile built from
// 8x8 DMMA accumulators.  The MMA is used "transposed" (MMA-M <-> matrix column n, MMA-N <-> matrix
// row m) so that the two accumulator values a lane owns are adjacent rows of one column of the
// column-major C and are written with one 16-byte store.  Shared tiles are padded so that the
// fragment reads (lane t -> [t/4][t%4]) are bank-conflict free for both operand orientations.
// Complex GEMM runs 4 real DMMAs per tile pair on the interleaved (re,im) data.
#pragma once
#include <atomic>
#include <cstdlib>

#include "sla_common.cuh"

namespace sla {

// number of kernels this library has launched in this process (bench.py reports it as gpu_launches)
extern std::atomic<unsigned long long> g_launch_count;

constexpr int kMaxDevices = 64;
inline int current_device() { int d = -1; return cudaGetDevice(&d) == cudaSuccess ? d : -1; }

enum Op : int { OP_N = 0, OP_C = 1 };  // OP_C = conjugate transpose (plain transpose for real)

template <typename T> struct GemmArgs {
    const T* A; const T* B; T* C;
    long long strideA, strideB, strideC;  // in elements; 0 = operand shared by the whole batch
    int lda, ldb, ldc;
    int M, N, K;
    // row / column scaling vectors are plain MULTIPLIERS (the D, D^-1, min(D,1), 1/max(D,1) factors are
    // prepared by scale_prep_batched); rs_mode / cs_mode are ignored by the kernel
    const double* rs; long long stride_rs; int rs_mode;   // row scaling (length M), may be null
    int rs_div; long long stride_rs_hi;                    // rs_div>0: rs offset = (b/rs_div)*stride_rs_hi + (b%rs_div)*stride_rs
    const double* cs; long long stride_cs; int cs_mode;   // column scaling (length N), may be null
    int accumulate;                                        // C += ... instead of C = ...
    int batch;
};

template <typename T> struct TilePad { static constexpr int value = 4; };
template <> struct TilePad<cplx> { static constexpr int value = 2; };

// One operand tile in shared memory.  IDXC=true : stored [k][idx] (idx contiguous, ld = BIDX+pad)
//                                     IDXC=false: stored [idx][k] (k contiguous,   ld = BK+pad)
template <typename T, bool IDXC, int BIDX, int BK> struct OperandTile {
    static constexpr int PAD = TilePad<T>::value;
    static constexpr int LD = (IDXC ? BIDX : BK) + PAD;
    static constexpr int ELEMS = (IDXC ? BK : BIDX) * LD;
    static constexpr int EPC = 16 / (int)sizeof(T);  // elements per 16-byte chunk

    // global element (idx, k) lives at g[idx*sidx + k*sk] with exactly one of sidx/sk == 1
    template <int NT, bool VEC>
    __device__ static __forceinline__ void load(T* s, const T* g, int ld, int idx0, int k0, int idx_max,
                                                int k_max, int tid) {
        if (IDXC) {
            constexpr int CPR = BIDX / EPC;  // chunks per k-row
            constexpr int TOTAL = CPR * BK;
            for (int c = tid; c < TOTAL; c += NT) {
                int kk = c / CPR, ii = (c % CPR) * EPC;
                T* dst = s + kk * LD + ii;
                int gi = idx0 + ii, gk = k0 + kk;
                const T* src = g + (long long)gk * ld + gi;
                int nvalid = (gk < k_max) ? max(0, min(EPC, idx_max - gi)) : 0;
                copy_chunk<VEC>(dst, src, nvalid, g);
            }
        } else {
            constexpr int CPR = BK / EPC;  // chunks per idx-row
            constexpr int TOTAL = CPR * BIDX;
            for (int c = tid; c < TOTAL; c += NT) {
                int ii = c / CPR, kk = (c % CPR) * EPC;
                T* dst = s + ii * LD + kk;
                int gi = idx0 + ii, gk = k0 + kk;
                const T* src = g + (long long)gi * ld + gk;
                int nvalid = (gi < idx_max) ? max(0, min(EPC, k_max - gk)) : 0;
                copy_chunk<VEC>(dst, src, nvalid, g);
            }
        }
    }
    template <bool VEC>
    __device__ static __forceinline__ void copy_chunk(T* dst, const T* src, int nvalid, const T* safe) {
        if constexpr (VEC || EPC == 1) {
            cp_async16(dst, nvalid > 0 ? (const void*)src : (const void*)safe, nvalid * (int)sizeof(T));
        } else {  // real data whose rows are only 8-byte aligned (odd ld): two 8-byte copies
            cp_async8(dst, nvalid > 0 ? (const void*)src : (const void*)safe, nvalid > 0 ? 8 : 0);
            cp_async8((char*)dst + 8, nvalid > 1 ? (const void*)((const char*)src + 8) : (const void*)safe,
                      nvalid > 1 ? 8 : 0);
        }
    }
    // fragment element [idx][k]
    __device__ static __forceinline__ T frag(const T* s, int idx, int k) {
        return IDXC ? s[k * LD + idx] : s[idx * LD + k];
    }
};

// Loader of the FAST instantiation (every tile full, 16-byte aligned rows): the per-thread chunk -> address map is
// computed ONCE, before the k loop.  A thread's PER chunks of an operand tile are c = tid + i NT; when NT is a multiple
// of the chunks per row they share one column offset and sit a fixed number of rows apart, so a k-tile costs
// PER (64-bit add + LDGSTS with an immediate shared-memory offset) instead of the ~30 instructions of index / bounds /
// zero-fill arithmetic per chunk of the general loader (measured round 2: 275 of the 490 instructions a warp issues
// per k-tile were that arithmetic, profiles/gemm_fast_r2.txt).
template <typename T, bool IDXC, int BIDX, int BK, int NT> struct FastLoader {
    using Tile = OperandTile<T, IDXC, BIDX, BK>;
    static constexpr int EPC = Tile::EPC, LD = Tile::LD;
    static constexpr int CPR = (IDXC ? BIDX : BK) / EPC;      // 16-byte chunks per contiguous row of the tile
    static constexpr int ROWS = IDXC ? BK : BIDX;
    static constexpr int TOTAL = CPR * ROWS;
    static constexpr bool OK = (TOTAL % NT == 0) && ((IDXC ? BIDX : BK) % EPC == 0);
    static constexpr int PER = OK ? TOTAL / NT : 1;
    static constexpr bool AFFINE = (NT % CPR == 0);
    static constexpr int RSTEP = AFFINE ? NT / CPR : 0;       // rows between consecutive chunks of a thread

    const char* g;                 // this thread's chunk 0 of the current k-tile
    long long gstep, kstep;        // bytes between its consecutive chunks (AFFINE) / between k-tiles
    unsigned s0;                   // shared-memory byte offset of chunk 0 inside the operand tile
    int goff[AFFINE ? 1 : PER];    // !AFFINE: byte offsets of the chunks from g / inside the tile
    unsigned soff[AFFINE ? 1 : PER];

    // tile origin: global element (idx0, k = 0) of an operand with leading dimension ld
    __device__ __forceinline__ void init(const T* base, int ld, int idx0, int tid) {
        const long long row_b = (long long)ld * (long long)sizeof(T);
        const T* origin = IDXC ? base + idx0 : base + (long long)idx0 * ld;
        kstep = IDXC ? row_b * BK : (long long)BK * (long long)sizeof(T);
        if constexpr (AFFINE) {
            const int r = tid / CPR, cc = tid % CPR;
            g = reinterpret_cast<const char*>(origin) + r * row_b + (long long)cc * 16;
            gstep = row_b * RSTEP;
            s0 = (unsigned)((r * LD + cc * EPC) * (int)sizeof(T));
        } else {
            g = reinterpret_cast<const char*>(origin);
            gstep = 0; s0 = 0;
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int c = tid + i * NT, r = c / CPR, cc = c % CPR;
                goff[i] = (int)(r * row_b) + cc * 16;
                soff[i] = (unsigned)((r * LD + cc * EPC) * (int)sizeof(T));
            }
        }
    }
    // issue the copies of the current k-tile into the tile at shared address `stile`, then step to the next k-tile
    __device__ __forceinline__ void issue(unsigned stile) {
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            if constexpr (AFFINE) {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(stile + s0 + (unsigned)(i * RSTEP * LD * (int)sizeof(T))),
                             "l"(g + i * gstep));
            } else {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(stile + soff[i]), "l"(g + goff[i]));
            }
        }
        g += kstep;
    }
};

// acc(8x8 tile pair) += nfrag (x) mfrag
__device__ __forceinline__ void mma_tile(double (&acc)[2], double nf, double mf) { dmma884(acc[0], acc[1], nf, mf); }
__device__ __forceinline__ void mma_tile(cplx (&acc)[2], cplx nf, cplx mf) {
    dmma884(acc[0].re, acc[1].re, nf.re, mf.re);
    dmma884(acc[0].re, acc[1].re, -nf.im, mf.im);
    dmma884(acc[0].im, acc[1].im, nf.re, mf.im);
    dmma884(acc[0].im, acc[1].im, nf.im, mf.re);
}

template <typename T, int OPA, int OPB, int BM, int BN, int WM, int WN, int BK, int STAGES, bool VEC, bool FAST = false>
// 64x64 real tiles: cap the registers at 128 so that 4 CTAs (not 3) are resident per SM -- fewer, fuller waves for the
// 1024-2048 CTA launches of a 64-128 chain batch
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32, (sizeof(T) == 8 && BM * BN <= 64 * 64) ? 4 : 1) gemm_kernel(GemmArgs<T> p) {
    constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
    constexpr int NT = WARPS_M * WARPS_N * 32;
    constexpr int MF = WM / 8, NF = WN / 8;
    // op(A)[m][k]: OP_N -> A(m,k) col-major => m contiguous (IDXC); OP_C -> stored (k,m) => k contiguous
    using TileA = OperandTile<T, OPA == OP_N, BM, BK>;
    // op(B)[k][n]: OP_N -> B(k,n) col-major => k contiguous; OP_C -> stored (n,k) => n contiguous (IDXC)
    using TileB = OperandTile<T, OPB == OP_C, BN, BK>;
    constexpr int STAGE_ELEMS = TileA::ELEMS + TileB::ELEMS;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WARPS_M, wn = warp / WARPS_M;
    const int tiles_m = (p.M + BM - 1) / BM;
    const int tiles = tiles_m * ((p.N + BN - 1) / BN);
    const int tile = blockIdx.x % tiles;
    const int m0 = (tile % tiles_m) * BM, n0 = (tile / tiles_m) * BN;
    const int b = blockIdx.x / tiles;

    const T* A = p.A + (long long)b * p.strideA;
    const T* B = p.B + (long long)b * p.strideB;
    T* C = p.C + (long long)b * p.strideC;

    T acc[NF][MF][2];
#pragma unroll
    for (int i = 0; i < NF; ++i)
#pragma unroll
        for (int j = 0; j < MF; ++j) { acc[i][j][0] = mk<T>(0.0); acc[i][j][1] = mk<T>(0.0); }

    // scaling multipliers of this lane's rows / columns: loaded up front so that their global-memory latency
    // hides behind the main loop instead of stalling the epilogue
    const double* rs = nullptr;
    if (p.rs) rs = (p.rs_div > 0) ? p.rs + (long long)(b / p.rs_div) * p.stride_rs_hi + (long long)(b % p.rs_div) * p.stride_rs
                                  : p.rs + (long long)b * p.stride_rs;
    const double* cs = p.cs ? p.cs + (long long)b * p.stride_cs : nullptr;
    double rsv[MF][2], csv_[NF];
    {
        const int fk0 = lane & 3, fi0 = lane >> 2;
#pragma unroll
        for (int j = 0; j < MF; ++j) {
            const int m = m0 + wm * WM + j * 8 + 2 * fk0;
            rsv[j][0] = (rs && m < p.M) ? rs[m] : 1.0;
            rsv[j][1] = (rs && m + 1 < p.M) ? rs[m + 1] : 1.0;
        }
#pragma unroll
        for (int i = 0; i < NF; ++i) {
            const int n = n0 + wn * WN + i * 8 + fi0;
            csv_[i] = (cs && n < p.N) ? cs[n] : 1.0;
        }
    }

    const int KT = (p.K + BK - 1) / BK;
    using FastA = FastLoader<T, OPA == OP_N, BM, BK, NT>;
    using FastB = FastLoader<T, OPB == OP_C, BN, BK, NT>;
    FastA fla;
    FastB flb;
    const unsigned smem_base = smem_u32(smem);
    if constexpr (FAST) {
        fla.init(A, p.lda, m0, tid);
        flb.init(B, p.ldb, n0, tid);
    }
    auto load_stage = [&](int stage, int kt) {
#ifdef SLA_GEMM_ABLATE_LOADS   // timing experiment only (wrong results): no global->shared traffic after the prologue
        if (kt >= STAGES - 1) return;
#endif
        if constexpr (FAST) {   // k-tiles are requested in order: the loaders carry the running pointers
            const unsigned st = smem_base + (unsigned)(stage * STAGE_ELEMS * (int)sizeof(T));
            fla.issue(st);
            flb.issue(st + (unsigned)(TileA::ELEMS * (int)sizeof(T)));
        } else {
            T* sa = smem + stage * STAGE_ELEMS;
            T* sb = sa + TileA::ELEMS;
            TileA::template load<NT, VEC>(sa, A, p.lda, m0, kt * BK, p.M, p.K, tid);
            TileB::template load<NT, VEC>(sb, B, p.ldb, n0, kt * BK, p.N, p.K, tid);
        }
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_stage(s, s);
        cp_async_commit();
    }

    const int fi = lane >> 2, fk = lane & 3;
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nk = kt + STAGES - 1;
            if (nk < KT) load_stage(nk % STAGES, nk);
            cp_async_commit();
        }
        const T* sa = smem + (kt % STAGES) * STAGE_ELEMS;
        const T* sb = sa + TileA::ELEMS;
        // register double buffering of the fragments: the LDS of step kk+4 are issued before the DMMAs of kk
        T nf[2][NF], mf[2][MF];
        auto load_frags = [&](int buf, int kk) {
#pragma unroll
            for (int i = 0; i < NF; ++i) {
                T v = TileB::frag(sb, wn * WN + i * 8 + fi, kk + fk);
                nf[buf][i] = (OPB == OP_C) ? conj_(v) : v;
            }
#pragma unroll
            for (int j = 0; j < MF; ++j) {
                T v = TileA::frag(sa, wm * WM + j * 8 + fi, kk + fk);
                mf[buf][j] = (OPA == OP_C) ? conj_(v) : v;
            }
        };
        load_frags(0, 0);
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            const int cur = (kk >> 2) & 1;
            if (kk + 4 < BK) load_frags(cur ^ 1, kk + 4);
#pragma unroll
            for (int i = 0; i < NF; ++i)
#pragma unroll
                for (int j = 0; j < MF; ++j) mma_tile(acc[i][j], nf[cur][i], mf[cur][j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: lane owns C[m, m+1][n] with m = m0 + wm*WM + j*8 + 2*(lane&3), n = n0 + wn*WN + i*8 + lane/4
    const bool vecC = ((p.ldc & 1) == 0 || sizeof(T) == 16) && ((((uintptr_t)C) & 15) == 0);
#pragma unroll
    for (int i = 0; i < NF; ++i) {
        const int n = n0 + wn * WN + i * 8 + fi;
        if (n >= p.N) continue;
        const double csv = csv_[i];
#pragma unroll
        for (int j = 0; j < MF; ++j) {
            const int m = m0 + wm * WM + j * 8 + 2 * fk;
            if (m >= p.M) continue;
            T* dst = C + (long long)n * p.ldc + m;
            T v0 = acc[i][j][0], v1 = acc[i][j][1];
            const bool has1 = (m + 1 < p.M);
            if (rs) {
                v0 = v0 * rsv[j][0];
                v1 = v1 * rsv[j][1];
            }
            if (cs) {
                v0 = v0 * csv;
                v1 = v1 * csv;
            }
            bool done = false;
            if constexpr (sizeof(T) == 8) {
                if (vecC && has1) {
                    double2* d2 = reinterpret_cast<double2*>(dst);
                    if (p.accumulate) {
                        double2 o = *d2;
                        v0 = v0 + o.x;
                        v1 = v1 + o.y;
                    }
                    *d2 = make_double2(v0, v1);
                    done = true;
                }
            }
            if (!done) {
                if (p.accumulate) {
                    v0 = v0 + dst[0];
                    if (has1) v1 = v1 + dst[1];
                }
                dst[0] = v0;
                if (has1) dst[1] = v1;
            }
        }
    }
}

template <typename T, int OPA, int OPB, int BM, int BN, int WM, int WN, int BK, int STAGES, bool VEC, bool FAST = false>
static cudaError_t launch_gemm_vec(const GemmArgs<T>& p, cudaStream_t stream) {
    using TileA = OperandTile<T, OPA == OP_N, BM, BK>;
    using TileB = OperandTile<T, OPB == OP_C, BN, BK>;
    constexpr int SMEM = STAGES * (TileA::ELEMS + TileB::ELEMS) * (int)sizeof(T);
    constexpr int NT = (BM / WM) * (BN / WN) * 32;
    auto kern = gemm_kernel<T, OPA, OPB, BM, BN, WM, WN, BK, STAGES, VEC, FAST>;
    // the attribute is per device (context): cached per instantiation AND device ordinal
    static bool configured[kMaxDevices] = {};
    const int dev = current_device();
    if (dev < 0 || dev >= kMaxDevices || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < kMaxDevices) configured[dev] = true;
    }
    int tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
    if (tiles <= 0 || p.batch <= 0) return cudaSuccess;
    kern<<<(unsigned)((long long)tiles * p.batch), NT, SMEM, stream>>>(p);
    ++g_launch_count;
    return cudaGetLastError();
}

// 16-byte cp.async needs even leading dimensions/strides and 16-byte aligned bases (always true for complex)
template <typename T, int OPA, int OPB, int BM, int BN, int WM, int WN, int BK, int STAGES>
static cudaError_t launch_gemm_cfg(const GemmArgs<T>& p, cudaStream_t stream) {
    bool vec = true;
    if (sizeof(T) == 8)
        vec = ((p.lda & 1) == 0) && ((p.ldb & 1) == 0) && ((p.strideA & 1) == 0) && ((p.strideB & 1) == 0) &&
              ((((uintptr_t)p.A) & 15) == 0) && ((((uintptr_t)p.B) & 15) == 0);
    constexpr int NT = (BM / WM) * (BN / WN) * 32;
    if constexpr (FastLoader<T, OPA == OP_N, BM, BK, NT>::OK && FastLoader<T, OPB == OP_C, BN, BK, NT>::OK) {
        // every tile full and every row 16-byte aligned: the loader with precomputed addresses
        static const bool fast_off = getenv("SLA_GEMM_FAST") && atoi(getenv("SLA_GEMM_FAST")) == 0;   // A/B experiments
        if (vec && !fast_off && p.M % BM == 0 && p.N % BN == 0 && p.K % BK == 0 && p.K >= BK)
            return launch_gemm_vec<T, OPA, OPB, BM, BN, WM, WN, BK, STAGES, true, true>(p, stream);
    }
    if (vec) return launch_gemm_vec<T, OPA, OPB, BM, BN, WM, WN, BK, STAGES, true>(p, stream);
    return launch_gemm_vec<T, OPA, OPB, BM, BN, WM, WN, BK, STAGES, false>(p, stream);
}

}  // namespace sla
