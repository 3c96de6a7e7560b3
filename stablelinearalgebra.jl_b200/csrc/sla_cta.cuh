// CTA-level building blocks for the "one CTA owns one matrix" LAPACK-class kernels (pivoted QR,
// Q formation, LU, triangular solves).  The matrix lives in global memory (L2-resident at the sizes
// of interest), panels live in shared memory, and the BLAS-3 parts run on DMMA.8x8x4 straight from
// the shared panels.
#pragma once
#include "sla_common.cuh"

namespace sla {

constexpr int CTA_THREADS = 512;
constexpr int CTA_WARPS = CTA_THREADS / 32;

template <typename T> __host__ __device__ constexpr int panel_ld(int n) {
    // leading dimension of an [i][idx] shared panel: conflict-free DMMA fragment reads need
    // ld = 4 (mod 16) for 8-byte and ld = 2 (mod 8) for 16-byte elements
    return sizeof(T) == 8 ? ((n + 15) / 16) * 16 + 4 : ((n + 7) / 8) * 8 + 2;
}

// block-wide sum; every thread gets the result.  `red` holds >= CTA_WARPS entries.  Two barriers.
template <typename V> __device__ __forceinline__ V block_sum(V v, V* red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    V s = red[0];
#pragma unroll
    for (int w = 1; w < CTA_WARPS; ++w) s = s + red[w];
    __syncthreads();
    return s;
}

// block-wide argmax (first maximal index wins); two barriers
__device__ __forceinline__ int block_argmax(double v, int i, double* redv, int* redi) {
    warp_argmax(v, i);
    if ((threadIdx.x & 31) == 0) { redv[threadIdx.x >> 5] = v; redi[threadIdx.x >> 5] = i; }
    __syncthreads();
    double bv = redv[0];
    int bi = redi[0];
#pragma unroll
    for (int w = 1; w < CTA_WARPS; ++w) {
        double ov = redv[w];
        int oi = redi[w];
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    __syncthreads();
    return bi;
}

__device__ __forceinline__ void mma_acc(double (&acc)[2], double nf, double mf) { dmma884(acc[0], acc[1], nf, mf); }
__device__ __forceinline__ void mma_acc(cplx (&acc)[2], cplx nf, cplx mf) {
    dmma884(acc[0].re, acc[1].re, nf.re, mf.re);
    dmma884(acc[0].re, acc[1].re, -nf.im, mf.im);
    dmma884(acc[0].im, acc[1].im, nf.re, mf.im);
    dmma884(acc[0].im, acc[1].im, nf.im, mf.re);
}

// C[r][c] -= sum_{i<kb} P[i][r] * (CONJU ? conj(U[i][c]) : U[i][c])   for r in [r_lo,r_hi), c in [c_lo,c_hi)
// C global column-major (ldc); P, U shared panels stored [i*ld + idx].  Rows i in [kb, kb4) of both
// panels must be zero (kb4 = kb rounded up to 4).  All threads of the CTA must call; no barrier inside.
template <typename T, bool CONJU>
__device__ void cta_rank_update(T* __restrict__ C, int ldc, int r_lo, int r_hi, int c_lo, int c_hi,
                                const T* __restrict__ Ps, const T* __restrict__ Us, int ld, int kb,
                                int n_clamp) {
    if (r_lo >= r_hi || c_lo >= c_hi) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fi = lane >> 2, fk = lane & 3;
    const int kb4 = (kb + 3) & ~3;
    const int r_base = r_lo & ~1;
    constexpr int NI = (sizeof(T) == 8) ? 4 : 2;  // 8-column fragments per warp tile (register budget)
    constexpr int TC = NI * 8;
    const int tiles_r = (r_hi - r_base + 31) / 32, tiles_c = (c_hi - c_lo + TC - 1) / TC;
    const bool vec = (sizeof(T) == 8) && ((ldc & 1) == 0) && ((((uintptr_t)C) & 15) == 0);
    for (int t = warp; t < tiles_r * tiles_c; t += CTA_WARPS) {
        const int r0 = r_base + (t % tiles_r) * 32, c0 = c_lo + (t / tiles_r) * TC;
        T acc[NI][4][2];
        // load C tile
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int c = c0 + i * 8 + fi;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = r0 + j * 8 + 2 * fk;
                T v0 = mk<T>(0.0), v1 = mk<T>(0.0);
                if (c < c_hi) {
                    const T* src = C + (long long)c * ldc + r;
                    bool done = false;
                    if constexpr (sizeof(T) == 8) {
                        if (vec && r >= r_lo && r + 1 < r_hi) {
                            double2 t2 = *reinterpret_cast<const double2*>(src);
                            v0 = t2.x; v1 = t2.y; done = true;
                        }
                    }
                    if (!done) {
                        if (r >= r_lo && r < r_hi) v0 = src[0];
                        if (r + 1 >= r_lo && r + 1 < r_hi) v1 = src[1];
                    }
                }
                acc[i][j][0] = v0; acc[i][j][1] = v1;
            }
        }
        for (int kk = 0; kk < kb4; kk += 4) {
            T nf[NI], mf[4];
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                T u = Us[(kk + fk) * ld + min(c0 + i * 8 + fi, n_clamp)];
                nf[i] = -(CONJU ? conj_(u) : u);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) mf[j] = Ps[(kk + fk) * ld + min(r0 + j * 8 + fi, n_clamp)];
#pragma unroll
            for (int i = 0; i < NI; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) mma_acc(acc[i][j], nf[i], mf[j]);
        }
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int c = c0 + i * 8 + fi;
            if (c >= c_hi) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = r0 + j * 8 + 2 * fk;
                T* dst = C + (long long)c * ldc + r;
                bool done = false;
                if constexpr (sizeof(T) == 8) {
                    if (vec && r >= r_lo && r + 1 < r_hi) {
                        *reinterpret_cast<double2*>(dst) = make_double2(acc[i][j][0], acc[i][j][1]);
                        done = true;
                    }
                }
                if (!done) {
                    if (r >= r_lo && r < r_hi) dst[0] = acc[i][j][0];
                    if (r + 1 >= r_lo && r + 1 < r_hi) dst[1] = acc[i][j][1];
                }
            }
        }
    }
}

// W[i][c] = sum_{r in [r_lo,r_hi)} conj(P[i][r]) * C[r][c]   for i < kb4 (kb rounded up to 8), c in [c_lo,c_hi)
// P shared panel [i*ld + r] (rows i >= kb must be zero), C global column-major, W shared [i*ld + c].
// Entries P[i][r] for r outside [r_lo, r_hi) must be zero OR r-range 4-aligned: the k-loop starts at
// r_lo rounded down to 4 and runs to r_hi rounded up to 4 with C reads masked to [r_lo,r_hi).
template <typename T>
__device__ void cta_panel_dot(T* __restrict__ Ws, const T* __restrict__ Ps, int ld, int kb,
                              const T* __restrict__ C, int ldc, int r_lo, int r_hi, int c_lo, int c_hi,
                              int n_clamp) {
    if (c_lo >= c_hi) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fi = lane >> 2, fk = lane & 3;
    const int nif = (kb + 7) / 8;  // 8-row fragments of W actually needed (<= 4)
    const int tiles_c = (c_hi - c_lo + 7) / 8;
    const int k_lo = r_lo & ~3;
    for (int t = warp; t < tiles_c; t += CTA_WARPS) {
        const int c = c_lo + t * 8 + fi;
        const bool cvalid = c < c_hi;
        const T* col = C + (long long)min(c, c_hi - 1) * ldc;
        T acc[4][2];
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[j][0] = mk<T>(0.0); acc[j][1] = mk<T>(0.0); }
        // eight row steps per trip: all eight global loads of C are in flight before the first DMMA (one L2 round trip
        // per 32 rows instead of one per 4: this loop is latency-bound, profiles/ldr_stream_phases_r2.txt)
        T cn[8];                                       // the NEXT trip's eight loads are issued before this trip's DMMAs
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int r = k_lo + 4 * u + fk;
            cn[u] = (cvalid && r >= r_lo && r < r_hi) ? col[r] : mk<T>(0.0);
        }
        for (int kk0 = k_lo; kk0 < r_hi; kk0 += 32) {
            T cf[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) cf[u] = cn[u];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = kk0 + 32 + 4 * u + fk;
                cn[u] = (cvalid && r >= r_lo && r < r_hi) ? col[r] : mk<T>(0.0);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = kk0 + 4 * u + fk;
                if (kk0 + 4 * u < r_hi) {
                    const int rc = min(r, n_clamp);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (j < nif) {
                            T pf = conj_(Ps[(j * 8 + fi) * ld + rc]);
                            mma_acc(acc[j], cf[u], pf);
                        }
                    }
                }
            }
        }
        // lane holds D[ci = fi][ii = 2*fk + {0,1}]  ->  W[j*8 + 2*fk + {0,1}][c]
        if (cvalid) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j < nif) {
                    Ws[(j * 8 + 2 * fk) * ld + c] = acc[j][0];
                    Ws[(j * 8 + 2 * fk + 1) * ld + c] = acc[j][1];
                }
            }
        }
    }
}

}  // namespace sla
