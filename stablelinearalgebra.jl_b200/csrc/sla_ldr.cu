// Batched LDR factorisation kernel:  A P = Q R'  ->  L = Q, d = |diag R'|, R = D^-1 R' P^T.
//
// Replaces `ldr!(F, ws)` (src/LDR.jl:160-188) = geqp3! (src/qr.jl:70-92, LAPACK ?geqp3) + triu/abs/
// ldiv!/mul_P! (LDR.jl:169-182, developer_functions.jl:20-26) + orgqr! (src/qr.jl:95-109, LAPACK
// ?orgqr/?ungqr) by ONE kernel launch for the whole batch: one CTA owns one matrix.
//
// Algorithm (own formulation, same mathematics as LAPACK's blocked QRCP):
//   * greedy column pivoting on down-dated partial column norms (first maximal index wins), with the
//     LAPACK cancellation safeguard -- but instead of cutting the panel short when a down-date is
//     unreliable, the affected column's norm is recomputed exactly on the spot from the panel;
//   * panel of NB Householder reflectors kept in shared memory (V) together with F = tau * A^H v
//     (the "laqps" auxiliary matrix), so that per column only ONE pass over the trailing matrix is
//     needed (the GEMV that produces the new column of F); the pivot row is kept up to date;
//   * rank-NB trailing update A22 -= V2 * F2^H on the FP64 tensor pipe (DMMA) directly from the shared
//     panels;
//   * T factors of the compact-WY form are built for free from the Gram dot products V^H v that the
//     per-column GEMV produces anyway, and reused to form Q = H_1...H_n blockwise (backward
//     accumulation, two DMMA products per panel).
// Householder conventions are LAPACK's (?larfg): beta = -sign(Re alpha)*||(alpha,x)||, H = I - tau v v^H.
#include <cstdio>

#include "sla_cta.cuh"
#include "sla_kernels.h"

// -DSLA_LDRS_PROFILE: per-phase cycle counters of CTA 0 (thread 0), printed at the end of the kernel
#ifdef SLA_LDRS_PROFILE
#define SP_DECL long long sp_[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long spt_ = clock64();
#define SP(i) do { long long t_ = clock64(); sp_[i] += t_ - spt_; spt_ = t_; } while (0)
#define SP_DUMP() do { if (blockIdx.x == 0 && threadIdx.x == 0) printf("ldr_kernel n=%d: init %lld | P1 fetch+apply %lld | P2 reflector %lld | scale+store %lld | GEMV %lld | gram %lld | P4 finish/downdate %lld | flagged %lld | argmax tail %lld | T + rank update %lld | R extract %lld | Q formation %lld\n", n, sp_[0], sp_[1], sp_[2], sp_[3], sp_[4], sp_[5], sp_[6], sp_[7], sp_[8], sp_[9], sp_[10], sp_[11]); } while (0)
#else
#define SP_DECL
#define SP(i) do { } while (0)
#define SP_DUMP() do { } while (0)
#endif

namespace sla {

namespace {

constexpr double TOL3Z = 1.4901161193847656e-08;  // sqrt(eps), LAPACK ?laqps

template <typename T> struct LdrSmem {
    T* Vs; T* Fs; T* Gp; T* Ts; T* tau; T* auxv;
    double* vn1; double* vn2; double* redv;
    int* jpvt; int* redi; int* flag; int* nflag;
    int ld;
};

template <typename T> __host__ __device__ inline size_t ldr_smem_bytes(int n, int NB) {
    size_t ld = panel_ld<T>(n);
    size_t b = 0;
    b += 2 * NB * ld * sizeof(T);              // Vs, Fs
    b += 2 * (size_t)NB * NB * sizeof(T);      // Gp, Ts
    b += (size_t)n * sizeof(T);                // tau
    b += (size_t)NB * sizeof(T);               // auxv
    b += 2 * (size_t)n * sizeof(double);       // vn1, vn2
    b += 32 * sizeof(double);                  // redv
    b += (size_t)n * sizeof(int);              // jpvt
    b += (size_t)n * sizeof(int);              // flag list
    b += 64 * sizeof(int);                     // redi + nflag
    return b + 64;
}

template <typename T> __device__ inline LdrSmem<T> carve(unsigned char* raw, int n, int NB) {
    LdrSmem<T> s;
    s.ld = panel_ld<T>(n);
    T* t = reinterpret_cast<T*>(raw);
    s.Vs = t; t += (size_t)NB * s.ld;
    s.Fs = t; t += (size_t)NB * s.ld;
    s.Gp = t; t += NB * NB;
    s.Ts = t; t += NB * NB;
    s.tau = t; t += n;
    s.auxv = t; t += NB;
    double* dd = reinterpret_cast<double*>(t);
    s.vn1 = dd; dd += n;
    s.vn2 = dd; dd += n;
    s.redv = dd; dd += 32;
    int* ii = reinterpret_cast<int*>(dd);
    s.jpvt = ii; ii += n;
    s.flag = ii; ii += n;
    s.redi = ii; ii += 32;
    s.nflag = ii;
    return s;
}

// F(c) = tau * sum_{r>=k} conj(A[r][c]) * v[r]  for every trailing column c in [c_lo, n): the one pass
// over the trailing matrix that column-pivoted QR needs per column (the BLAS-2 half of ?geqp3).
// Bandwidth-bound on the L2->SM path, so the mapping keeps many independent 16-byte loads in flight:
// LPC lanes share a column (fewer lanes as the remaining height shrinks) and every lane works on UC
// column groups at once; partial sums are combined with warp shuffles.
template <typename T, int UC>
__device__ __forceinline__ void trailing_gemv(const T* __restrict__ A, int lda, int c_lo, int n,
                                              const T* __restrict__ v, int k, T tau, T* __restrict__ Fout,
                                              bool vec) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ncols = n - c_lo;
    if (ncols <= 0) return;
    const int mrem = n - (k & ~1);
    const int per_lane = (sizeof(T) == 8 && vec) ? 2 : 1;
    const int lpc = (mrem > 64 * per_lane) ? 32 : ((mrem > 32 * per_lane) ? 16 : 8);
    const int cpw = 32 / lpc;
    const int sub = lane / lpc, sl = lane % lpc;
    const int cols_per_pass = cpw * UC;
    for (int g = warp; g * cols_per_pass < ncols; g += CTA_WARPS) {
        const int cbase = c_lo + g * cols_per_pass + sub;
        T acc[UC];
#pragma unroll
        for (int u = 0; u < UC; ++u) acc[u] = mk<T>(0.0);
        bool done = false;
        if constexpr (sizeof(T) == 8) {
            if (vec) {
                double acc2[UC];
#pragma unroll
                for (int u = 0; u < UC; ++u) acc2[u] = 0.0;
                // chunks of 4 row steps: all UC*4 16-byte loads are issued before the first FMA (one L2 round
                // trip per chunk instead of one per row step)
                const bool cok[4] = {cbase < n, cbase + cpw < n, cbase + 2 * cpw < n, cbase + 3 * cpw < n};
                for (int r0 = (k & ~1) + 2 * sl; r0 < n; r0 += 8 * lpc) {
                    double2 a[UC][4];
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int r = r0 + it * 2 * lpc;
#pragma unroll
                        for (int u = 0; u < UC; ++u)
                            a[u][it] = (r < n && cok[u & 3])
                                           ? __ldcg(reinterpret_cast<const double2*>(A + (long long)(cbase + u * cpw) * lda + r))
                                           : make_double2(0.0, 0.0);
                    }
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int r = r0 + it * 2 * lpc;
                        const double2 w = (r < n) ? *reinterpret_cast<const double2*>(v + r) : make_double2(0.0, 0.0);
#pragma unroll
                        for (int u = 0; u < UC; ++u) {
                            acc[u] = fma(a[u][it].x, w.x, acc[u]);
                            acc2[u] = fma(a[u][it].y, w.y, acc2[u]);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < UC; ++u) acc[u] += acc2[u];
                done = true;
            }
        }
        if (!done) {
#pragma unroll 2
            for (int r = k + sl; r < n; r += lpc) {
                const T w = v[r];
#pragma unroll
                for (int u = 0; u < UC; ++u) {
                    const int c = cbase + u * cpw;
                    if (c < n) fmacc(acc[u], conj_(A[(long long)c * lda + r]), w);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UC; ++u) {
            T a = acc[u];
            for (int o = lpc >> 1; o > 0; o >>= 1) {
                if constexpr (sizeof(T) == 8) a += __shfl_xor_sync(0xffffffffu, a, o);
                else { a.re += __shfl_xor_sync(0xffffffffu, a.re, o); a.im += __shfl_xor_sync(0xffffffffu, a.im, o); }
            }
            const int c = cbase + u * cpw;
            if (sl == 0 && c < n) Fout[c] = tau * a;
        }
    }
}

}  // namespace

template <typename T, int NB>
__global__ void __launch_bounds__(CTA_THREADS, 1) ldr_kernel(LdrArgs<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n;
    LdrSmem<T> s = carve<T>(smem_raw, n, NB);
    const int ld = s.ld;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    T* A = p.L + (long long)b * p.strideL;
    const int lda = p.ldl;
    T* Tw = p.Twork + (long long)b * p.strideT;
    const bool vec = (sizeof(T) == 8) && ((lda & 1) == 0) && ((n & 1) == 0) && ((((uintptr_t)A) & 15) == 0);

    SP_DECL
    // ---------------------------------------------------------------- init: column norms, jpvt
    const T* Ain = p.Ain ? p.Ain + (long long)b * p.strideAin : nullptr;
    for (int c = warp; c < n; c += CTA_WARPS) {
        const T* col = Ain ? Ain + (long long)c * p.ldain : A + (long long)c * lda;
        T* dcol = A + (long long)c * lda;
        double acc = 0.0;
        bool nz = false;
        for (int r = lane; r < n; r += 32) {
            T x = col[r];
            if (Ain) dcol[r] = x;
            acc += abs2_(x);
            nz |= !is_zero(x);
        }
        acc = warp_sum(acc);
        nz = __any_sync(0xffffffffu, nz);
        if (lane == 0 && ldr_norm2_out_of_range(acc, nz)) ldr_raise_range(p.info, b);
        if (lane == 0) { double nr = sqrt(acc); s.vn1[c] = nr; s.vn2[c] = nr; s.jpvt[c] = c; }
    }
    __syncthreads();
    int pvt;
    {
        double bv = -1.0; int bi = 0x7fffffff;
        for (int c = tid; c < n; c += CTA_THREADS) {
            double v = s.vn1[c];
            if (v > bv) { bv = v; bi = c; }
        }
        pvt = block_argmax(bv, bi, s.redv, s.redi);
        if (pvt >= n) pvt = 0;  // all-NaN guard
    }

    SP(0);
    // ---------------------------------------------------------------- pivoted QR, panel by panel
    for (int p0 = 0; p0 < n; p0 += NB) {
        const int jb = min(NB, n - p0);
        for (int i = tid; i < NB * NB; i += CTA_THREADS) s.Gp[i] = mk<T>(0.0);
        for (int j = 0; j < jb; ++j) {
            const int k = p0 + j;
            T* vj = s.Vs + (size_t)j * ld;
            SP(9);
            // ---- P1: bring the pivot column to position k, apply the panel's reflectors to it
            double ss = 0.0;
            for (int r = tid; r < n; r += CTA_THREADS) {
                T x = A[(long long)pvt * lda + r];
                if (pvt != k) {
                    T xk = A[(long long)k * lda + r];
                    A[(long long)pvt * lda + r] = xk;
                }
                if (r >= k) {
                    for (int i = 0; i < j; ++i) fms(x, s.Vs[(size_t)i * ld + r], conj_(s.Fs[(size_t)i * ld + pvt]));
                    vj[r] = x;
                    if (r > k) ss += abs2_(x);
                } else {
                    if (pvt != k) A[(long long)k * lda + r] = x;
                    vj[r] = mk<T>(0.0);
                }
            }
            ss = warp_sum(ss);
            if (lane == 0) s.redv[warp] = ss;
            __syncthreads();
            SP(1);
            // ---- P2: Householder reflector (LAPACK ?larfg)
            double xnorm2 = 0.0;
#pragma unroll
            for (int w = 0; w < CTA_WARPS; ++w) xnorm2 += s.redv[w];
            const T alpha = vj[k];
            T tau, scal;
            double beta;
            if (xnorm2 == 0.0 && imag_(alpha) == 0.0) {
                tau = mk<T>(0.0); scal = mk<T>(0.0); beta = real_(alpha);
            } else {
                beta = -copysign(sqrt(abs2_(alpha) + xnorm2), real_(alpha));
                tau = mk<T>((beta - real_(alpha)) / beta, -imag_(alpha) / beta);
                scal = mk<T>(1.0) / (alpha - mk<T>(beta));
            }
            __syncthreads();  // everyone has read vj[k] and redv
            SP(2);
            for (int r = tid; r < n; r += CTA_THREADS) {
                if (r > k) {
                    T v = vj[r] * scal;
                    vj[r] = v;
                    A[(long long)k * lda + r] = v;
                } else if (r == k) {
                    vj[r] = mk<T>(1.0);
                    A[(long long)k * lda + r] = mk<T>(beta);
                }
            }
            if (pvt != k) {
                if (tid < j) s.Fs[(size_t)tid * ld + pvt] = s.Fs[(size_t)tid * ld + k];
                if (tid == CTA_THREADS - 1) {
                    int t = s.jpvt[pvt]; s.jpvt[pvt] = s.jpvt[k]; s.jpvt[k] = t;
                    s.vn1[pvt] = s.vn1[k]; s.vn2[pvt] = s.vn2[k];
                }
            }
            if (tid == CTA_THREADS - 2) s.tau[k] = tau;
            __syncthreads();
            SP(3);
            // ---- P3: one pass over the trailing matrix: F(c,j) = tau * A(k:n,c)^H v ; Gram dots
            trailing_gemv<T, 4>(A, lda, k + 1, n, vj, k, tau, s.Fs + (size_t)j * ld, vec);
            SP(4);
            for (int i = warp; i < j; i += CTA_WARPS) {
                const T* vi = s.Vs + (size_t)i * ld;
                T acc = mk<T>(0.0);
                for (int r = k + lane; r < n; r += 32) fmacc(acc, conj_(vi[r]), vj[r]);
                acc = warp_sum(acc);
                if (lane == 0) { s.Gp[i * NB + j] = acc; s.auxv[i] = -(tau * acc); }
            }
            if (tid == 0) *s.nflag = 0;
            __syncthreads();
            SP(5);
            // ---- P4: finish F(:,j), update pivot row k, down-date the partial norms
            double bv = -1.0; int bi = 0x7fffffff;
            for (int c = k + 1 + tid; c < n; c += CTA_THREADS) {
                T f = s.Fs[(size_t)j * ld + c];
                for (int i = 0; i < j; ++i) fmacc(f, s.Fs[(size_t)i * ld + c], s.auxv[i]);
                s.Fs[(size_t)j * ld + c] = f;
                T a = A[(long long)c * lda + k];
                for (int i = 0; i < j; ++i) fms(a, s.Vs[(size_t)i * ld + k], conj_(s.Fs[(size_t)i * ld + c]));
                a = a - conj_(f);  // i = j term, V(k,j) = 1
                A[(long long)c * lda + k] = a;
                double v1 = s.vn1[c];
                if (v1 != 0.0) {
                    double temp = abs_(a) / v1;
                    temp = fmax(0.0, (1.0 + temp) * (1.0 - temp));
                    double q = v1 / s.vn2[c];
                    double temp2 = temp * q * q;
                    if (temp2 <= TOL3Z) {
                        int slot = atomicAdd(s.nflag, 1);
                        s.flag[slot] = c;
                    } else {
                        v1 *= sqrt(temp);
                        s.vn1[c] = v1;
                    }
                }
                if (v1 > bv) { bv = v1; bi = c; }
            }
            warp_argmax(bv, bi);
            if (lane == 0) { s.redv[warp] = bv; s.redi[warp] = bi; }
            __syncthreads();
            SP(6);
            const int nflag = *s.nflag;
            if (nflag > 0) {
                // exact recomputation of the flagged columns' norms below row k (panel applied on the fly)
                for (int f = warp; f < nflag; f += CTA_WARPS) {
                    const int c = s.flag[f];
                    const T* col = A + (long long)c * lda;
                    double acc = 0.0;
                    for (int r = k + 1 + lane; r < n; r += 32) {
                        T x = col[r];
                        for (int i = 0; i <= j; ++i) fms(x, s.Vs[(size_t)i * ld + r], conj_(s.Fs[(size_t)i * ld + c]));
                        acc += abs2_(x);
                    }
                    acc = warp_sum(acc);
                    if (lane == 0) { double nr = sqrt(acc); s.vn1[c] = nr; s.vn2[c] = nr; }
                }
                __syncthreads();
                bv = -1.0; bi = 0x7fffffff;
                for (int c = k + 1 + tid; c < n; c += CTA_THREADS) {
                    double v = s.vn1[c];
                    if (v > bv) { bv = v; bi = c; }
                }
                pvt = block_argmax(bv, bi, s.redv, s.redi);
                SP(7);
            } else {
                double fv = s.redv[0]; int fidx = s.redi[0];
#pragma unroll
                for (int w = 1; w < CTA_WARPS; ++w) {
                    double ov = s.redv[w]; int oi = s.redi[w];
                    if (ov > fv || (ov == fv && oi < fidx)) { fv = ov; fidx = oi; }
                }
                pvt = fidx;
                __syncthreads();
            }
            if (pvt >= n) pvt = min(k + 1, n - 1);
            SP(8);
        }
        // ---- panel done: T factor (one warp) || rank-jb trailing update (everyone)
        const int kend = p0 + jb;
        {
            const int jb4 = (jb + 3) & ~3;
            for (int i = jb * ld + tid; i < jb4 * ld; i += CTA_THREADS) { s.Vs[i] = mk<T>(0.0); s.Fs[i] = mk<T>(0.0); }
            __syncthreads();
        }
        if (warp == CTA_WARPS - 1 && lane < jb) {
            // row `lane` of T:  T(l,l) = tau_l ; T(l,i) = -tau_i * sum_{q=l}^{i-1} T(l,q) G(q,i)
            const int l = lane;
            T* Trow = s.Ts + l * NB;
            Trow[l] = s.tau[p0 + l];
            for (int i = l + 1; i < jb; ++i) {
                T acc = mk<T>(0.0);
                for (int q = l; q < i; ++q) fmacc(acc, Trow[q], s.Gp[q * NB + i]);
                Trow[i] = -(s.tau[p0 + i] * acc);
            }
            T* Tg = Tw + (size_t)(p0 / NB) * NB * NB;
            for (int i = 0; i < jb; ++i) Tg[l * NB + i] = (i >= l) ? Trow[i] : mk<T>(0.0);
        }
        cta_rank_update<T, true>(A, lda, kend, n, kend, n, s.Vs, s.Fs, ld, jb, n - 1);
        __syncthreads();
    }

    SP(9);
    // ---------------------------------------------------------------- d = |diag R'|, R = D^-1 R' P^T
    {
        double* d = p.d + (long long)b * p.stride_d;
        T* R = p.R + (long long)b * p.strideR;
        for (int r = tid; r < n; r += CTA_THREADS) {
            const double dr = abs_(A[(long long)r * lda + r]);
            d[r] = dr;
            for (int c = 0; c < n; ++c) {
                T v = (c >= r) ? A[(long long)c * lda + r] / dr : mk<T>(0.0);
                R[(long long)s.jpvt[c] * p.ldr + r] = v;
            }
        }
        if (p.jpvt_out)
            for (int c = tid; c < n; c += CTA_THREADS) p.jpvt_out[(long long)b * n + c] = s.jpvt[c] + 1;
    }
    __syncthreads();
    SP(10);

    // ---------------------------------------------------------------- Q = H_1 ... H_n (backward, blocked)
    const int npan = (n + NB - 1) / NB;
    for (int pi = npan - 1; pi >= 0; --pi) {
        const int p0 = pi * NB, jb = min(NB, n - p0), kend = p0 + jb;
        // a. panel reflectors -> shared (unit diagonal, zeros above), T -> shared
        for (int i = 0; i < NB; ++i) {
            const int c = p0 + i;
            for (int r = tid; r < n; r += CTA_THREADS) {
                T v = mk<T>(0.0);
                if (i < jb) {
                    if (r > c) v = A[(long long)c * lda + r];
                    else if (r == c) v = mk<T>(1.0);
                }
                s.Vs[(size_t)i * ld + r] = v;
            }
        }
        {
            const T* Tg = Tw + (size_t)pi * NB * NB;
            for (int i = tid; i < NB * NB; i += CTA_THREADS) {
                int l = i / NB, q = i % NB;
                s.Ts[i] = (l < jb && q < jb) ? Tg[i] : mk<T>(0.0);
            }
        }
        __syncthreads();
        // b. initial content of the region this panel touches: [I 0; 0 Qsub] and zeros above
        for (int i = 0; i < jb; ++i) {
            const int c = p0 + i;
            for (int r = tid; r < n; r += CTA_THREADS) A[(long long)c * lda + r] = mk<T>(r == c ? 1.0 : 0.0);
        }
        for (int c = kend + warp; c < n; c += CTA_WARPS)
            for (int r = p0 + lane; r < kend; r += 32) A[(long long)c * lda + r] = mk<T>(0.0);
        __syncthreads();
        // c. W = V^H C
        cta_panel_dot<T>(s.Fs, s.Vs, ld, jb, A, lda, p0, n, p0, n, n - 1);
        __syncthreads();
        // d. W <- T W   (upper-triangular T, in place top-down)
        for (int c = p0 + tid; c < n; c += CTA_THREADS) {
            for (int l = 0; l < jb; ++l) {
                T acc = mk<T>(0.0);
                for (int q = l; q < jb; ++q) fmacc(acc, s.Ts[l * NB + q], s.Fs[(size_t)q * ld + c]);
                s.Fs[(size_t)l * ld + c] = acc;
            }
            for (int l = jb; l < ((jb + 3) & ~3); ++l) s.Fs[(size_t)l * ld + c] = mk<T>(0.0);
        }
        __syncthreads();
        // e. C -= V W
        cta_rank_update<T, false>(A, lda, p0, n, p0, n, s.Vs, s.Fs, ld, jb, n - 1);
        __syncthreads();
    }
    SP(11);
    SP_DUMP();
}

template <typename T> static int pick_nb(int n) {
    const size_t limit = 227 * 1024;
    if (ldr_smem_bytes<T>(n, 32) <= limit) return 32;
    if (ldr_smem_bytes<T>(n, 16) <= limit) return 16;
    if (ldr_smem_bytes<T>(n, 8) <= limit) return 8;
    return 0;
}

template <typename T> size_t ldr_twork_elems(int n) {
    int nb = pick_nb<T>(n);
    if (nb == 0) return 0;
    return (size_t)((n + nb - 1) / nb) * nb * nb;
}

template <typename T, int NB> static cudaError_t launch_ldr_nb(const LdrArgs<T>& p, cudaStream_t stream) {
    auto kern = ldr_kernel<T, NB>;
    size_t smem = ldr_smem_bytes<T>(p.n, NB);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<p.batch, CTA_THREADS, smem, stream>>>(p);
    ++g_launch_count;
    return cudaGetLastError();
}

template <typename T> cudaError_t ldr_batched(const LdrArgs<T>& p, cudaStream_t stream) {
    if (p.batch <= 0 || p.n <= 0) return cudaSuccess;
    switch (pick_nb<T>(p.n)) {
        case 32: return launch_ldr_nb<T, 32>(p, stream);
        case 16: return launch_ldr_nb<T, 16>(p, stream);
        case 8: return launch_ldr_nb<T, 8>(p, stream);
        default: return cudaErrorInvalidValue;
    }
}

template cudaError_t ldr_batched<double>(const LdrArgs<double>&, cudaStream_t);
template size_t ldr_twork_elems<double>(int);
template cudaError_t ldr_batched<cplx>(const LdrArgs<cplx>&, cudaStream_t);
template size_t ldr_twork_elems<cplx>(int);

}  // namespace sla
