// Batched partial-pivot LU with fused log|det| / sign, and the LU-based inverse.
//
// Replaces det_lu! (src/lu.jl:123-140: getrf! + diagonal/pivot scan) and inv_lu! (src/lu.jl:149-158:
// det_lu! + getri!) -- LAPACK ?getrf / ?getri behind src/lu.jl:63-93 -- with one kernel launch per
// batch: one CTA owns one matrix.  Right-looking blocked LU: the NB-column panel is factorised in
// shared memory (pivot search = block-wide argmax with "first maximal index" tie-break like i?amax),
// the panel's row interchanges are applied to the rest of the matrix as one gathered permutation,
// U12 comes from a register-resident forward substitution and the rank-NB trailing update runs on
// DMMA from the shared panels.  The inverse is formed as X = U^-1 (L^-1 P) by two blocked triangular
// sweeps built from the same pieces (mathematically the same inverse ?getri returns).
#include <cooperative_groups.h>
#include <cstdio>

#include "sla_cta.cuh"
#include "sla_kernels.h"

namespace cg = cooperative_groups;

namespace sla {

namespace {

// a NaN/Inf pivot means an earlier factorisation of the same composite call was singular: report it like a zero pivot
__device__ __forceinline__ bool lu_finite(double x) { return fabs(x) <= 1.7976931348623157e308; }
__device__ __forceinline__ bool lu_finite(cplx x) { return lu_finite(x.re) && lu_finite(x.im); }

constexpr int LU_NB = 32;

template <typename T> __host__ __device__ inline size_t lu_smem_bytes(int n, int NB) {
    size_t ld = panel_ld<T>(n);
    size_t b = 2 * NB * ld * sizeof(T);       // Ps, Us
    b += 32 * sizeof(double) + 32 * sizeof(T);  // redv, redt
    b += (size_t)n * sizeof(int) * 2;         // ipiv, rowpos
    b += (4 * NB + 64) * sizeof(int);         // touched rows, sources, redi, misc
    return b + 64;
}

#ifdef SLA_LU_PROFILE
#define LUP_DECL long long lup_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long lupt_ = clock64();
#define LUP(i) do { __syncthreads(); long long t_ = clock64(); lup_[i] += t_ - lupt_; lupt_ = t_; } while (0)
#define LUP_DUMP(tag) do { if (blockIdx.x == 0 && threadIdx.x == 0) printf("%s: panelLU %lld  store %lld  permsim %lld  permapply %lld  U12 %lld  update %lld  other %lld\n", tag, lup_[0], lup_[1], lup_[2], lup_[3], lup_[4], lup_[5], lup_[6]); } while (0)
#else
#define LUP_DECL
#define LUP(i) do { } while (0)
#define LUP_DUMP(tag) do { } while (0)
#endif

// PIPE: the CTA runs as rank 0 of a cluster whose other CTAs form L^-1 P (the forward sweep of the inverse) panel by
// panel WHILE this CTA factorises: one cluster barrier per panel, right after the panel (final with respect to its own
// pivots) and its pivots are published.  The later interchanges that touch this panel's rows in global memory only
// happen after the NEXT barrier, i.e. after the peers have consumed the panel (lu_forward_pipe).
template <typename T, int NB, bool PIPE = false>
__device__ void lu_factor(T* __restrict__ A, int lda, int n, T* Ps, T* Us, int ld, int* ipiv, int* rowpos,
                          int* tl, int* tsrc, double* redv, int* redi, int* misc, int* info_out) {
    const int tid = threadIdx.x;
    for (int r = tid; r < n; r += CTA_THREADS) rowpos[r] = -1;
    LUP_DECL
    for (int p0 = 0; p0 < n; p0 += NB) {
        const int jb = min(NB, n - p0), kend = p0 + jb;
        LUP(6);
        if (n <= CTA_THREADS) {
            // ---- register-row panel LU: thread r holds row r of the panel (NB values) for the whole panel.
            // Per column: warp argmax -> block maximum (1 barrier), the pivot row and row k change owners through
            // shared memory (1 barrier), every thread below scales its entry and updates its row from registers.
            const int r = tid;
            const bool inrow = (r >= p0 && r < n);
            T arow[NB];
#pragma unroll
            for (int i = 0; i < NB; ++i) arow[i] = (i < jb && inrow) ? A[(long long)(p0 + i) * lda + r] : mk<T>(0.0);
            T* rowP = Us;            // the pivot row of the current column
            T* rowK = Us + NB;       // the row it displaces
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                if (j < jb) {
                    const int k = p0 + j;
                    int piv;
                    if constexpr (!is_cplx<T>::value) {
                        // exact 64-bit maximum of |a| by two 32-bit warp reductions (the bit pattern of a non-negative
                        // double is monotone), first maximal row by a ballot; 0 = inactive row or NaN
                        const double v = (r >= k && r < n) ? abs1_(arow[j]) : -1.0;
                        const unsigned long long key = (v >= 0.0) ? (unsigned long long)__double_as_longlong(v) + 1ull : 0ull;
                        const unsigned khi = (unsigned)(key >> 32);
                        const unsigned mhi = __reduce_max_sync(0xffffffffu, khi);
                        const unsigned mlo = __reduce_max_sync(0xffffffffu, khi == mhi ? (unsigned)key : 0u);
                        const unsigned long long wk = ((unsigned long long)mhi << 32) | mlo;
                        const unsigned who = __ballot_sync(0xffffffffu, key == wk);
                        if ((tid & 31) == 0) {
                            reinterpret_cast<unsigned long long*>(redv)[tid >> 5] = wk;
                            redi[tid >> 5] = (wk != 0ull) ? (tid & ~31) + __ffs(who) - 1 : 0x7fffffff;
                        }
                        __syncthreads();
                        const int l16 = tid & 15;
                        const unsigned long long ok = reinterpret_cast<const unsigned long long*>(redv)[l16];
                        const unsigned ohi = (unsigned)(ok >> 32);
                        const unsigned ghi = __reduce_max_sync(0xffffffffu, ohi);
                        const unsigned glo = __reduce_max_sync(0xffffffffu, ohi == ghi ? (unsigned)ok : 0u);
                        const unsigned long long gk = ((unsigned long long)ghi << 32) | glo;
                        const unsigned gwho = __ballot_sync(0xffffffffu, ok == gk) & 0xffffu;   // warps in row order
                        piv = (gk != 0ull) ? redi[__ffs(gwho) - 1] : 0x7fffffff;
                    } else {
                        double v = (r >= k && r < n) ? abs1_(arow[j]) : -1.0;
                        int idx = r;
                        if (!(v >= 0.0)) { v = -1.0; idx = 0x7fffffff; }   // inactive rows and NaNs never win
                        warp_argmax(v, idx);
                        if ((tid & 31) == 0) { redv[tid >> 5] = v; redi[tid >> 5] = idx; }
                        __syncthreads();
                        double bv = redv[0];
                        piv = redi[0];
#pragma unroll
                        for (int w = 1; w < CTA_WARPS; ++w) {
                            const double ov = redv[w];
                            const int oi = redi[w];
                            if (ov > bv || (ov == bv && oi < piv)) { bv = ov; piv = oi; }
                        }
                    }
                    if (piv >= n) piv = k;   // NaN guard
                    if (r == piv) {
#pragma unroll
                        for (int i = 0; i < NB; ++i) rowP[i] = arow[i];
                    }
                    if (r == k && piv != k) {
#pragma unroll
                        for (int i = 0; i < NB; ++i) rowK[i] = arow[i];
                    }
                    if (tid == 0) ipiv[k] = piv;
                    __syncthreads();
                    if (piv != k) {
                        if (r == k) {
#pragma unroll
                            for (int i = 0; i < NB; ++i) arow[i] = rowP[i];
                        } else if (r == piv) {
#pragma unroll
                            for (int i = 0; i < NB; ++i) arow[i] = rowK[i];
                        }
                    }
                    const T pv = rowP[j];
                    if (is_zero(pv) || !lu_finite(pv)) {
                        if (tid == 0 && info_out && *info_out == 0) *info_out = k + 1;
                    }
                    if (!is_zero(pv) && r > k && r < n) {
                        T l;
                        if constexpr (!is_cplx<T>::value) {   // reciprocal scaling like ?getf2 (division below sfmin)
                            if (fabs(pv) >= 2.2250738585072014e-308) l = arow[j] * __drcp_rn(pv);
                            else l = arow[j] / pv;
                        } else {
                            l = arow[j] / pv;
                        }
                        arow[j] = l;
#pragma unroll
                        for (int i = j + 1; i < NB; ++i) fms(arow[i], l, rowP[i]);
                    }
                }
            }
            // the factored panel -> shared (the block steps below read it from there)
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                for (int rr = tid; rr < n; rr += CTA_THREADS) Ps[(size_t)i * ld + rr] = (rr >= p0) ? arow[i] : mk<T>(0.0);
            }
            __syncthreads();
        } else {
            // ---- load panel rows [p0, n)
            for (int i = 0; i < NB; ++i) {
                for (int r = tid; r < n; r += CTA_THREADS)
                    Ps[(size_t)i * ld + r] = (i < jb && r >= p0) ? A[(long long)(p0 + i) * lda + r] : mk<T>(0.0);
            }
            __syncthreads();
            // ---- unblocked LU of the panel in shared memory
            for (int j = 0; j < jb; ++j) {
                const int k = p0 + j;
                T* colj = Ps + (size_t)j * ld;
                double bv = -1.0; int bi = 0x7fffffff;
                for (int r = k + tid; r < n; r += CTA_THREADS) {
                    double v = abs1_(colj[r]);
                    if (v > bv) { bv = v; bi = r; }
                }
                int piv = block_argmax(bv, bi, redv, redi);
                if (piv >= n) piv = k;  // NaN guard
                if (tid < jb && piv != k) {
                    T* ci = Ps + (size_t)tid * ld;
                    T t = ci[k]; ci[k] = ci[piv]; ci[piv] = t;
                }
                if (tid == 0) ipiv[k] = piv;
                __syncthreads();
                const T pv = colj[k];
                if (is_zero(pv) || !lu_finite(pv)) {
                    if (tid == 0 && info_out && *info_out == 0) *info_out = k + 1;
                }
                if (!is_zero(pv)) {
                    for (int r = k + 1 + tid; r < n; r += CTA_THREADS) {
                        T l = colj[r] / pv;
                        colj[r] = l;
                        for (int i = j + 1; i < jb; ++i) fms(Ps[(size_t)i * ld + r], l, Ps[(size_t)i * ld + k]);
                    }
                }
                __syncthreads();
            }
        }
        LUP(0);
        // ---- panel back to global
        for (int i = 0; i < jb; ++i)
            for (int r = p0 + tid; r < n; r += CTA_THREADS) A[(long long)(p0 + i) * lda + r] = Ps[(size_t)i * ld + r];
        if constexpr (PIPE) {
            __threadfence();
            cg::this_cluster().sync();   // panel p0 and ipiv[p0, kend) are published
        }
        LUP(1);
        // ---- net row permutation of this panel (thread 0 simulates the jb interchanges)
        if (tid == 0) {
            int nt = jb;
            for (int q = 0; q < jb; ++q) { tl[q] = p0 + q; tsrc[q] = p0 + q; rowpos[p0 + q] = q; }
            for (int j = 0; j < jb; ++j) {
                int pj = ipiv[p0 + j];
                if (pj == p0 + j) continue;
                int pos = rowpos[pj];
                if (pos < 0) { pos = nt++; tl[pos] = pj; tsrc[pos] = pj; rowpos[pj] = pos; }
                int t = tsrc[j]; tsrc[j] = tsrc[pos]; tsrc[pos] = t;
            }
            for (int q = 0; q < nt; ++q) rowpos[tl[q]] = -1;
            misc[0] = nt;
        }
        __syncthreads();
        LUP(2);
        {
            const int nt = misc[0];
            const int ncols_out = n - jb;             // columns outside the panel: [0,p0) U [kend,n)
            const int ch = (NB * ld) / (2 * NB);      // columns per chunk staged through Us ([q][cl], q < 2*NB)
            for (int cb = 0; cb < ncols_out; cb += ch) {
                const int cn = min(ch, ncols_out - cb);
                // a warp per touched row, lanes over the columns (no integer division in the index math)
                for (int q = tid >> 5; q < nt; q += CTA_WARPS) {
                    const int rs_ = tsrc[q];
                    if (rs_ == tl[q]) continue;   // this row stays where it is
                    for (int cl = tid & 31; cl < cn; cl += 32) {
                        int c = cb + cl; c = (c < p0) ? c : c + jb;
                        Us[(size_t)q * ch + cl] = A[(long long)c * lda + rs_];
                    }
                }
                __syncthreads();
                for (int q = tid >> 5; q < nt; q += CTA_WARPS) {
                    const int rd_ = tl[q];
                    if (tsrc[q] == rd_) continue;
                    for (int cl = tid & 31; cl < cn; cl += 32) {
                        int c = cb + cl; c = (c < p0) ? c : c + jb;
                        A[(long long)c * lda + rd_] = Us[(size_t)q * ch + cl];
                    }
                }
                __syncthreads();
            }
        }
        LUP(3);
        if (kend < n) {
            // ---- U12 = L11^-1 A12  (thread per column, registers)
            for (int c = kend + tid; c < n; c += CTA_THREADS) {
                T a[NB];
                T* col = A + (long long)c * lda + p0;
#pragma unroll
                for (int i = 0; i < NB; ++i) a[i] = (i < jb) ? col[i] : mk<T>(0.0);
#pragma unroll
                for (int i = 1; i < NB; ++i) {
                    if (i < jb) {
#pragma unroll
                        for (int q = 0; q < i; ++q) fms(a[i], Ps[(size_t)q * ld + p0 + i], a[q]);
                    }
                }
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    if (i < jb) col[i] = a[i];
                    Us[(size_t)i * ld + c] = a[i];
                }
            }
            __syncthreads();
            LUP(4);
            cta_rank_update<T, false>(A, lda, kend, n, kend, n, Ps, Us, ld, jb, n - 1);
            __syncthreads();
            LUP(5);
        }
    }
    LUP_DUMP("lu_factor");
}

// logdet = sum log|u_ii|, sgn = prod sign(u_ii) * (-1)^{#interchanges}   (src/lu.jl:128-139)
template <typename T>
__device__ void lu_logdet(const T* __restrict__ A, int lda, int n, const int* ipiv, double* redv, T* redt,
                          double& logdet, T& sgn) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double ld_acc = 0.0;
    T sg = mk<T>(1.0);
    for (int i = tid; i < n; i += CTA_THREADS) {
        T u = A[(long long)i * lda + i];
        ld_acc += log(abs_(u));
        sg = sg * sign_(u);
        if (ipiv[i] != i) sg = -sg;
    }
    ld_acc = warp_sum(ld_acc);
    for (int o = 16; o > 0; o >>= 1) {
        T other;
        if constexpr (sizeof(T) == 8) other = __shfl_xor_sync(0xffffffffu, sg, o);
        else { other.re = __shfl_xor_sync(0xffffffffu, sg.re, o); other.im = __shfl_xor_sync(0xffffffffu, sg.im, o); }
        sg = sg * other;
    }
    if (lane == 0) { redv[warp] = ld_acc; redt[warp] = sg; }
    __syncthreads();
    logdet = 0.0;
    sgn = mk<T>(1.0);
    for (int w = 0; w < CTA_WARPS; ++w) { logdet += redv[w]; sgn = sgn * redt[w]; }
    __syncthreads();
}

// One panel of the forward sweep Y = L^-1 (.) on the columns [c_lo, c_hi) of X: substitution with the unit-lower L11 in
// registers (thread per column), then the rank-NB update of the rows below on the tensor cores.
template <typename T, int NB>
__device__ void lu_forward_step(const T* __restrict__ A, int lda, int n, T* __restrict__ X, int ldx, T* Ps, T* Us, int ld,
                                int p0, int c_lo, int c_hi) {
    const int tid = threadIdx.x;
    const int jb = min(NB, n - p0), kend = p0 + jb;
    for (int i = 0; i < NB; ++i)
        for (int r = tid; r < n; r += CTA_THREADS)
            Ps[(size_t)i * ld + r] = (i < jb && r > p0 + i) ? A[(long long)(p0 + i) * lda + r] : mk<T>(0.0);
    __syncthreads();
    for (int c = c_lo + tid; c < c_hi; c += CTA_THREADS) {
        T a[NB];
        T* col = X + (long long)c * ldx + p0;
#pragma unroll
        for (int i = 0; i < NB; ++i) a[i] = (i < jb) ? col[i] : mk<T>(0.0);
#pragma unroll
        for (int i = 1; i < NB; ++i) {
            if (i < jb) {
#pragma unroll
                for (int q = 0; q < i; ++q) fms(a[i], Ps[(size_t)q * ld + p0 + i], a[q]);
            }
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            if (i < jb) col[i] = a[i];
            Us[(size_t)i * ld + c] = a[i];
        }
    }
    __syncthreads();
    cta_rank_update<T, false>(X, ldx, kend, n, c_lo, c_hi, Ps, Us, ld, jb, n - 1);
    __syncthreads();
}

// Rank >= 1 of a PIPE cluster: Y = L^-1 P I (or L^-1 P RHS) on the columns [c_lo, c_hi) of X, one panel per cluster
// barrier, concurrently with the factorisation in rank 0 (lu_factor<PIPE>).  P is applied incrementally -- the
// interchanges of panel p to the rows of X when panel p arrives -- which is equivalent to permuting first: the rows of L
// below a panel and the rows of X are then in the same (time-p) order, and later interchanges permute both alike.
template <typename T, int NB>
__device__ void lu_forward_pipe(const T* __restrict__ A, int lda, int n, T* __restrict__ X, int ldx, T* Ps, T* Us, int ld,
                                const int* __restrict__ ipiv0, int* piv_s, const T* __restrict__ rhs, int ldrhs,
                                int c_lo, int c_hi) {
    const int tid = threadIdx.x;
    for (int c = c_lo + (tid >> 5); c < c_hi; c += CTA_WARPS)
        for (int r = (tid & 31); r < n; r += 32)
            X[(long long)c * ldx + r] = rhs ? rhs[(long long)c * ldrhs + r] : mk<T>(r == c ? 1.0 : 0.0);
    __syncthreads();
    for (int p0 = 0; p0 < n; p0 += NB) {
        const int jb = min(NB, n - p0);
        cg::this_cluster().sync();       // rank 0 has published panel p0 and its pivots
        if (tid < jb) piv_s[tid] = ipiv0[p0 + tid];   // distributed shared memory of rank 0
        __syncthreads();
        // the panel's interchanges, in order, on this CTA's columns: a thread per column
        for (int c = c_lo + tid; c < c_hi; c += CTA_THREADS) {
            T* col = X + (long long)c * ldx;
            for (int j = 0; j < jb; ++j) {
                const int k = p0 + j, pv = piv_s[j];
                if (pv != k) { const T t = col[k]; col[k] = col[pv]; col[pv] = t; }
            }
        }
        __syncthreads();
        lu_forward_step<T, NB>(A, lda, n, X, ldx, Ps, Us, ld, p0, c_lo, c_hi);
    }
}

// X = A^-1 * RHS from the LU factors (RHS = identity when rhs == nullptr, i.e. the inverse), columns [c_lo, c_hi) of X
// only: the columns of X are independent, so the CTAs of a cluster split them (see lu_kernel).
template <typename T, int NB>
__device__ void lu_inverse(const T* __restrict__ A, int lda, int n, T* __restrict__ X, int ldx, T* Ps, T* Us,
                           int ld, const int* ipiv, int* rowpos, const T* __restrict__ rhs, int ldrhs, T* rdiag,
                           int c_lo, int c_hi, bool skip_forward = false) {
    const int tid = threadIdx.x;
    if (!skip_forward) {
        // X = P (row-permuted identity): row r of P*I is e_{src(r)}
        for (int r = tid; r < n; r += CTA_THREADS) rowpos[r] = r;
        __syncthreads();
        if (tid == 0)
            for (int k = 0; k < n; ++k) {
                int pk = ipiv[k];
                if (pk != k) { int t = rowpos[k]; rowpos[k] = rowpos[pk]; rowpos[pk] = t; }
            }
        __syncthreads();
        for (int c = c_lo + (tid >> 5); c < c_hi; c += CTA_WARPS)
            for (int r = (tid & 31); r < n; r += 32)
                X[(long long)c * ldx + r] = rhs ? rhs[(long long)c * ldrhs + rowpos[r]] : mk<T>(rowpos[r] == c ? 1.0 : 0.0);
        __syncthreads();
    }
    // ---- forward sweep: Y = L^-1 P   (skipped when the cluster peers formed it during the factorisation)
    if (!skip_forward)
        for (int p0 = 0; p0 < n; p0 += NB) lu_forward_step<T, NB>(A, lda, n, X, ldx, Ps, Us, ld, p0, c_lo, c_hi);
    // ---- backward sweep: X = U^-1 Y
    const int npan = (n + NB - 1) / NB;
    for (int pi = npan - 1; pi >= 0; --pi) {
        const int p0 = pi * NB, jb = min(NB, n - p0);
        for (int i = 0; i < NB; ++i)
            for (int r = tid; r < n; r += CTA_THREADS)
                Ps[(size_t)i * ld + r] = (i < jb && r <= p0 + i) ? A[(long long)(p0 + i) * lda + r] : mk<T>(0.0);
        __syncthreads();
        // reciprocals of the panel's diagonal, once per panel (the substitution below multiplies)
        if (tid < jb) rdiag[tid] = mk<T>(1.0) / Ps[(size_t)tid * ld + p0 + tid];
        __syncthreads();
        for (int c = c_lo + tid; c < c_hi; c += CTA_THREADS) {
            T a[NB];
            T* col = X + (long long)c * ldx + p0;
#pragma unroll
            for (int i = 0; i < NB; ++i) a[i] = (i < jb) ? col[i] : mk<T>(0.0);
#pragma unroll
            for (int i = NB - 1; i >= 0; --i) {
                if (i < jb) {
#pragma unroll
                    for (int q = i + 1; q < NB; ++q)
                        if (q < jb) fms(a[i], Ps[(size_t)q * ld + p0 + i], a[q]);
                    a[i] = a[i] * rdiag[i];
                }
            }
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                if (i < jb) col[i] = a[i];
                Us[(size_t)i * ld + c] = a[i];
            }
        }
        __syncthreads();
        cta_rank_update<T, false>(X, ldx, 0, p0, c_lo, c_hi, Ps, Us, ld, jb, n - 1);
        __syncthreads();
    }
}

// One matrix per CTA, or per CLUSTER of CS CTAs: the factorisation (pivot search = a serial chain) runs in CTA 0; the
// two triangular sweeps that form X = U^-1 L^-1 P -- 3/4 of the flops of an inverse -- are independent per column of X, so
// the cluster splits the columns.  Small batches (64 matrices on 148 SMs) otherwise leave more than half of the GPU idle.
template <typename T, int NB>
__global__ void __launch_bounds__(CTA_THREADS, 1) lu_kernel(LuArgs<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::cluster_group cl = cg::this_cluster();
    const int CS = (int)cl.num_blocks(), rank = (int)cl.block_rank();
    const int n = p.n, b = blockIdx.x / CS, tid = threadIdx.x;
    const int ld = panel_ld<T>(n);
    T* Ps = reinterpret_cast<T*>(smem_raw);
    T* Us = Ps + (size_t)NB * ld;
    T* redt = Us + (size_t)NB * ld;
    double* redv = reinterpret_cast<double*>(redt + 32);
    int* ipiv = reinterpret_cast<int*>(redv + 32);
    int* rowpos = ipiv + n;
    int* tl = rowpos + n;
    int* tsrc = tl + 2 * NB;
    int* redi = tsrc + 2 * NB;
    int* misc = redi + 32;

    T* A = p.A + (long long)b * p.strideA;
    // info is cleared once per C-API entry point (sla_capi.cu: clear_info), not per kernel: composite calls run two or
    // three LU factorisations and the FIRST failure has to survive (the reference throws at every getrf!, lu.jl:75)
    int* info = p.info ? p.info + b : nullptr;
    double logdet = 0.0; T sgn = mk<T>(1.0);
    // PIPE (a cluster forming an inverse / a solve): rank 0 factorises, the other ranks form Y = L^-1 P panel by panel at
    // the same time (lu_forward_pipe), their columns of X split evenly among them; afterwards only the backward sweep is
    // left, split over ALL ranks.  SLA_LU_PIPE=0 (p.pipe == 0): the factorisation first, then both sweeps split.
    const bool pipe = (CS > 1) && p.invert && p.pipe;
    T* Xm = p.X ? p.X + (long long)b * p.strideX : nullptr;
    T* rhs_m = p.rhs ? p.rhs + (long long)b * p.strideRhs : nullptr;
    if (rank == 0) {
        if (pipe) lu_factor<T, NB, true>(A, p.lda, n, Ps, Us, ld, ipiv, rowpos, tl, tsrc, redv, redi, misc, info);
        else lu_factor<T, NB, false>(A, p.lda, n, Ps, Us, ld, ipiv, rowpos, tl, tsrc, redv, redi, misc, info);
        lu_logdet<T>(A, p.lda, n, ipiv, redv, redt, logdet, sgn);
        if (p.ipiv_out)
            for (int i = tid; i < n; i += CTA_THREADS) p.ipiv_out[(long long)b * n + i] = ipiv[i] + 1;
    } else if (pipe) {
        const int fper = (((n + CS - 2) / (CS - 1)) + 31) & ~31;
        const int f_lo = min(n, (rank - 1) * fper), f_hi = min(n, f_lo + fper);
        lu_forward_pipe<T, NB>(A, p.lda, n, Xm, p.ldx, Ps, Us, ld, cl.map_shared_rank(ipiv, 0), tl, rhs_m, p.ldrhs, f_lo, f_hi);
    }
    if (CS > 1) {
        __threadfence();   // the factors in global memory are read by the peers after the barrier
        cl.sync();
        if (rank != 0) {
            const int* rip = cl.map_shared_rank(ipiv, 0);
            for (int i = tid; i < n; i += CTA_THREADS) ipiv[i] = rip[i];
        }
        cl.sync();         // CTA 0 may reuse / leave its shared memory only after the peers have copied the pivots
    }
    if (p.invert) {
        // column slice of this CTA (multiples of the 32-column warp tile of cta_rank_update)
        const int per = (((n + CS - 1) / CS) + 31) & ~31;
        const int c_lo = min(n, rank * per), c_hi = min(n, c_lo + per);
        T* X = Xm;
        T* rhs = rhs_m;
        lu_inverse<T, NB>(A, p.lda, n, X, p.ldx, Ps, Us, ld, ipiv, rowpos, rhs, p.ldrhs, redt, c_lo, c_hi, pipe);
        if (rhs) {   // ldiv_lu!: the solution replaces the right-hand side (src/lu.jl:167-176)
            __syncthreads();
            for (int c = c_lo + (tid >> 5); c < c_hi; c += CTA_WARPS)
                for (int r = (tid & 31); r < n; r += 32) rhs[(long long)c * p.ldrhs + r] = X[(long long)c * p.ldx + r];
        } else if (p.copy_back) {
            if (CS > 1) cl.sync();   // every CTA has finished reading the factors that A^-1 now overwrites
            for (int c = c_lo + (tid >> 5); c < c_hi; c += CTA_WARPS)
                for (int r = (tid & 31); r < n; r += 32) A[(long long)c * p.lda + r] = X[(long long)c * p.ldx + r];
        }
        logdet = -logdet;      // src/lu.jl:157
        sgn = conj_(sgn);
    }
    if (tid == 0 && rank == 0) {
        if (p.logdet) p.logdet[b] = logdet;
        if (p.sign) p.sign[b] = sgn;
    }
}

}  // namespace

// cluster size of the inverse: as many CTAs per matrix as keep the whole batch in ONE wave (SLA_LU_CLUSTER overrides)
static int lu_cluster_size(int n, int batch, int invert) {
    static const int forced = getenv("SLA_LU_CLUSTER") ? atoi(getenv("SLA_LU_CLUSTER")) : 0;
    if (!invert) return 1;
    if (forced == 1 || forced == 2 || forced == 4) return forced;
    static int sms[kMaxDevices] = {};
    const int dev = current_device();
    int nsm = 148;
    if (dev >= 0 && dev < kMaxDevices) {
        if (sms[dev] == 0 && cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms[dev] = 148;
        nsm = sms[dev];
    }
    if (n >= 512 && (long long)batch * 4 <= nsm) return 4;
    if (n >= 128 && (long long)batch * 2 <= nsm) return 2;
    return 1;
}

template <typename T> cudaError_t lu_batched(const LuArgs<T>& p_in, cudaStream_t stream) {
    if (p_in.batch <= 0 || p_in.n <= 0) return cudaSuccess;
    if (lu_small_applies<T>(p_in)) return lu_small_inverse_batched<T>(p_in, stream);   // n <= 64, inverse: sla_lu_small.cu
    static const bool pipe_off = getenv("SLA_LU_PIPE") && atoi(getenv("SLA_LU_PIPE")) == 0;   // A/B experiments
    LuArgs<T> p = p_in;
    p.pipe = pipe_off ? 0 : 1;
    int nb = (sizeof(T) == 8) ? LU_NB : LU_NB / 2;  // register-resident substitution columns: a[NB]
    const size_t limit = 227 * 1024;
    while (nb >= 8 && lu_smem_bytes<T>(p.n, nb) > limit) nb >>= 1;
    if (nb < 8) return cudaErrorInvalidValue;
    size_t smem = lu_smem_bytes<T>(p.n, nb);
    const int cs = lu_cluster_size(p.n, p.batch, p.invert);
    auto launch = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)p.batch * cs, 1, 1);
        cfg.blockDim = dim3(CTA_THREADS, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)cs;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = cs > 1 ? 1 : 0;
        e = cudaLaunchKernelEx(&cfg, kern, p);
        ++g_launch_count;
        return e != cudaSuccess ? e : cudaGetLastError();
    };
    if (nb == 32) return launch(lu_kernel<T, 32>);
    if (nb == 16) return launch(lu_kernel<T, 16>);
    return launch(lu_kernel<T, 8>);
}

template <typename T> bool lu_supported(int n) {
    return n > 0 && lu_smem_bytes<T>(n, 8) <= (size_t)227 * 1024;
}
template bool lu_supported<double>(int);
template bool lu_supported<cplx>(int);

template cudaError_t lu_batched<double>(const LuArgs<double>&, cudaStream_t);
template cudaError_t lu_batched<cplx>(const LuArgs<cplx>&, cudaStream_t);

}  // namespace sla
