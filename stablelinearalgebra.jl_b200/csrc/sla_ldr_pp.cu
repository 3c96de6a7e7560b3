// Batched LDR factorisation with PANEL-RESTRICTED column pivoting (the communication-avoiding /
// tournament-pivoted variant of the pivoted QR, docs/src/references.bib:60-69) -- opt-in: SLA_LDR_PIVOT=panel.
//
// `ldr!(F, ws)` (src/LDR.jl:160-188) asks ?geqp3 for a GREEDY pivot per column, and the greedy choice needs the partial
// norms of ALL remaining columns after every reflector: one pass over the whole trailing matrix (a GEMV) and one global
// argmax per column -- the latency chain that bounds every strict kernel of this library (sections 4.2-4.4 of DESIGN.md).
// Here the pivots are chosen PER PANEL:
//   1. at a panel boundary the exact partial norms of all remaining columns are recomputed (one pass per NB columns
//      instead of one per column) and the NB largest columns become the panel;
//   2. inside the panel -- NB columns in shared memory -- the pivoting is greedy on exact norms, as ?laqp2 does it;
//   3. the trailing matrix then gets ONE compact-WY block update on the tensor cores (DMMA): W = V^H A, W <- T^H W,
//      A -= V W.
// The result is a valid factorisation A P = Q R' with |r_kk| non-increasing inside every panel and from panel to panel up
// to the panel's own spread; on the matrices a DQMC chain produces (columns graded over 50-100 decades) the Green's
// function, log|det| and sign agree with the ?geqp3 route at its own noise floor (1e-12, tests: test_chain_*_panel), but
// the pivot SEQUENCE -- and with it d and R entry by entry -- differs from LAPACK's whenever the greedy choice falls
// outside the current panel.  That is why it is a mode and not the default.
//
// One CTA per matrix; the matrix lives in global memory (L2), the panel in shared memory.  No per-column pass over the
// trailing matrix, no per-column cluster exchange: per column only CTA-local barriers and 2 NB-column-wide operations.
#include "sla_cta.cuh"
#include "sla_kernels.h"

namespace sla {

namespace {

template <typename T> __host__ __device__ inline size_t pp_smem_bytes(int n, int NB) {
    const size_t ld = panel_ld<T>(n);
    size_t b = 2 * NB * ld * sizeof(T);                 // panel (V) and W / staging
    b += 2 * (size_t)NB * NB * sizeof(T);               // Gram, T
    b += (size_t)n * sizeof(T);                         // tau
    b += (size_t)n * sizeof(unsigned long long);        // selection keys
    b += (size_t)n * sizeof(int);                       // jpvt
    b += (size_t)NB * (3 * sizeof(double) + 6 * sizeof(int));   // beta, pnf, pnt | sel, jsel, order, act, disp, vac
    b += 32 * sizeof(unsigned long long) + 16 * sizeof(int);
    return b + 128;
}

__device__ __forceinline__ double pp_nan_inf(double v) { return (v == v) ? v : INFINITY; }
// (squared norm, index) -> 64-bit key: larger norm wins, then the smaller index
__device__ __forceinline__ unsigned long long pp_key(double v, int c) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(pp_nan_inf(v));
    return (b & ~0x7FFull) | (unsigned long long)(2047 - c);
}
__device__ __forceinline__ int pp_key_index(unsigned long long key) { return 2047 - (int)(key & 0x7FFull); }
__device__ __forceinline__ unsigned long long pp_warp_max(unsigned long long k) {
    const unsigned hi = (unsigned)(k >> 32);
    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? (unsigned)k : 0u);
    return ((unsigned long long)mhi << 32) | mlo;
}

template <typename T, int NB>
__global__ void __launch_bounds__(CTA_THREADS, 1) ldr_pp_kernel(LdrArgs<T> p, int form_q) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n;
    const int ld = panel_ld<T>(n);
    T* Ps = reinterpret_cast<T*>(smem_raw);                      // the panel: [NB][ld]
    T* Ws = Ps + (size_t)NB * ld;                                // W = V^H A2 / staging: [NB][ld]
    T* Gs = Ws + (size_t)NB * ld;                                // Gram matrix V^H V: [NB][NB]
    T* Ts = Gs + NB * NB;                                        // T of the panel: [NB][NB]
    T* tau = Ts + NB * NB;                                       // [n]
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(tau + n);   // [n]
    unsigned long long* red = keys + n;                          // [32]
    double* beta = reinterpret_cast<double*>(red + 32);          // [NB]
    double* pnf = beta + NB;                                     // [NB] squared norm of a panel column, rows >= k
    double* pnt = pnf + NB;                                      // [NB] ... rows > k
    int* jpvt = reinterpret_cast<int*>(pnt + NB);                // [n]
    int* sel = jpvt + n;                                         // [NB] column positions picked for the panel
    int* jsel = sel + NB;                                        // [NB] their original column indices
    int* order = jsel + NB;                                      // [NB] slot chosen for panel position j
    int* act = order + NB;                                       // [NB] slot still unpivoted
    int* disp = act + NB;                                        // [NB] displaced / vacated pairs
    int* vac = disp + NB;
    int* misc = vac + NB;                                        // [16]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    T* A = p.L + (long long)b * p.strideL;
    const int lda = p.ldl;
    T* Tw = p.Twork + (long long)b * p.strideT;
    const T* Ain = p.Ain ? p.Ain + (long long)b * p.strideAin : nullptr;

    for (int c = tid; c < n; c += CTA_THREADS) jpvt[c] = c;

    for (int p0 = 0; p0 < n; p0 += NB) {
        const int jb = min(NB, n - p0), kend = p0 + jb;
        // ---- (a) exact squared partial norms (rows >= p0) of every remaining column -> selection keys
        for (int c = p0 + warp; c < n; c += CTA_WARPS) {
            const bool first = (p0 == 0) && Ain;
            const T* col = first ? Ain + (long long)c * p.ldain : A + (long long)c * lda;
            T* dcol = A + (long long)c * lda;
            double acc = 0.0;
            bool nz = false;
            for (int r = p0 + lane; r < n; r += 32) {
                const T x = col[r];
                if (first) dcol[r] = x;
                acc += abs2_(x);
                nz |= !is_zero(x);
            }
            acc = warp_sum(acc);
            if (p0 == 0) {
                nz = __any_sync(0xffffffffu, nz);
                if (lane == 0 && ldr_norm2_out_of_range(acc, nz)) ldr_raise_range(p.info, b);
            }
            if (lane == 0) keys[c] = pp_key(acc, c);
        }
        __syncthreads();
        // ---- (b) the jb largest columns become the panel (slot i = i-th largest)
        for (int i = 0; i < jb; ++i) {
            unsigned long long best = 0ull;
            for (int c = p0 + tid; c < n; c += CTA_THREADS) { const unsigned long long k = keys[c]; best = k > best ? k : best; }
            best = pp_warp_max(best);
            if (lane == 0) red[warp] = best;
            __syncthreads();
            unsigned long long win = red[0];
#pragma unroll
            for (int w = 1; w < CTA_WARPS; ++w) win = red[w] > win ? red[w] : win;
            int c = pp_key_index(win);
            if (win == 0ull || c < p0 || c >= n) c = p0 + i;          // degenerate input (NaNs everywhere): natural order
            if (tid == 0) { sel[i] = c; keys[c] = 0ull; }
            __syncthreads();
        }
        // ---- (c) gather the selected columns (all n rows) into the panel slots; make room at positions p0 .. kend - 1
        for (int i = 0; i < jb; ++i) {
            const T* src = A + (long long)sel[i] * lda;
            for (int r = tid; r < n; r += CTA_THREADS) Ps[(size_t)i * ld + r] = src[r];
        }
        if (tid == 0) {
            for (int q = 0; q < jb; ++q) act[q] = 0;                  // here: panel position p0 + q keeps a selected column
            for (int i = 0; i < jb; ++i) { jsel[i] = jpvt[sel[i]]; if (sel[i] < kend) act[sel[i] - p0] = 1; }
            int nd = 0, nv = 0;
            for (int q = 0; q < jb; ++q) if (!act[q]) disp[nd++] = p0 + q;
            for (int i = 0; i < jb; ++i) if (sel[i] >= kend) vac[nv++] = sel[i];
            for (int t = 0; t < nd; ++t) jpvt[vac[t]] = jpvt[disp[t]];
            misc[0] = nd;
            for (int q = 0; q < jb; ++q) act[q] = 1;                  // from here on: slot q not yet pivoted
        }
        __syncthreads();
        {
            const int nd = misc[0];
            for (int t = 0; t < nd; ++t) {                            // displaced column -> the slot a selected column left
                const T* src = A + (long long)disp[t] * lda;
                T* dst = A + (long long)vac[t] * lda;
                for (int r = tid; r < n; r += CTA_THREADS) dst[r] = src[r];
            }
        }
        for (int s = warp; s < jb; s += CTA_WARPS) {                   // exact norms of the panel columns: rows >= p0, rows > p0
            const T* a = Ps + (size_t)s * ld;
            double nf = 0.0, nt = 0.0;
            for (int r = p0 + lane; r < n; r += 32) { const double v = abs2_(a[r]); nf += v; if (r > p0) nt += v; }
            nf = warp_sum(nf); nt = warp_sum(nt);
            if (lane == 0) { pnf[s] = nf; pnt[s] = nt; }
        }
        __syncthreads();
        // ---- (d) panel factorisation in shared memory, greedy pivoting among the panel's columns (?laqp2 on exact norms)
        for (int j = 0; j < jb; ++j) {
            const int k = p0 + j;
            if (warp == 0) {
                unsigned long long key = (lane < jb && act[lane]) ? (pp_key(pnf[lane], lane) | 1ull << 63) : 0ull;
                key = pp_warp_max(key);
                if (lane == 0) { const int s = pp_key_index(key); order[j] = s; act[s] = 0; }
            }
            __syncthreads();
            const int sj = order[j];
            T* x = Ps + (size_t)sj * ld;
            const T alpha = x[k];
            const double xnorm2 = pnt[sj];
            T tk, scal;
            double bt;
            if (xnorm2 == 0.0 && imag_(alpha) == 0.0) {               // LAPACK ?larfg
                tk = mk<T>(0.0); scal = mk<T>(0.0); bt = real_(alpha);
            } else {
                bt = -copysign(sqrt(abs2_(alpha) + xnorm2), real_(alpha));
                tk = mk<T>((bt - real_(alpha)) / bt, -imag_(alpha) / bt);
                scal = mk<T>(1.0) / (alpha - mk<T>(bt));
            }
            __syncthreads();   // everybody has read alpha before it is overwritten
            for (int r = k + tid; r < n; r += CTA_THREADS) x[r] = (r == k) ? mk<T>(1.0) : x[r] * scal;
            if (tid == 0) { tau[k] = tk; beta[j] = bt; }
            __syncthreads();
            // H^H = I - conj(tau) v v^H on the unpivoted slots: a warp per slot, dot + update + next norms in one go
            const T ctau = conj_(tk);
            for (int s = warp; s < jb; s += CTA_WARPS) {
                if (!act[s]) continue;
                T* a = Ps + (size_t)s * ld;
                T dot = mk<T>(0.0);
                for (int r = k + lane; r < n; r += 32) fmacc(dot, conj_(x[r]), a[r]);
                dot = warp_sum(dot);
                const T w = ctau * dot;
                double nf = 0.0, nt = 0.0;
                for (int r = k + lane; r < n; r += 32) {
                    T v = a[r];
                    fms(v, x[r], w);
                    a[r] = v;
                    const double v2 = abs2_(v);
                    if (r > k) nf += v2;
                    if (r > k + 1) nt += v2;
                }
                nf = warp_sum(nf); nt = warp_sum(nt);
                if (lane == 0) { pnf[s] = nf; pnt[s] = nt; }
            }
            __syncthreads();
        }
        // ---- (e) position j of the panel <-> slot order[j]: columns back to global memory (rows above the panel block
        //      hold earlier rows of R', the block's upper triangle its own rows, beta on the diagonal, v below), V to Ws
        if (tid < jb) jpvt[p0 + tid] = jsel[order[tid]];
        for (int j = 0; j < jb; ++j) {
            const int k = p0 + j;
            const T* src = Ps + (size_t)order[j] * ld;
            T* dst = A + (long long)k * lda;
            T* vj = Ws + (size_t)j * ld;
            const T bj = mk<T>(beta[j]);
            for (int r = tid; r < n; r += CTA_THREADS) {
                const T v = src[r];
                dst[r] = (r == k) ? bj : v;
                vj[r] = (r < k) ? mk<T>(0.0) : v;                     // unit diagonal is already in place
            }
        }
        for (int i = jb * ld + tid; i < NB * ld; i += CTA_THREADS) Ws[i] = mk<T>(0.0);
        __syncthreads();
        { T* t = Ps; Ps = Ws; Ws = t; }                               // Ps = V (position order), Ws free
        for (int i = jb * ld + tid; i < NB * ld; i += CTA_THREADS) Ws[i] = mk<T>(0.0);
        // ---- (g) Gram matrix of the reflectors and the compact-WY T (?larft):  T(l,l) = tau_l, T(l,i) = -tau_i sum_q T(l,q) G(q,i)
        for (int pr = warp; pr < jb * jb; pr += CTA_WARPS) {
            const int i = pr / jb, j = pr % jb;
            if (i >= j) continue;
            const T* vi = Ps + (size_t)i * ld;
            const T* vj = Ps + (size_t)j * ld;
            T acc = mk<T>(0.0);
            for (int r = p0 + j + lane; r < n; r += 32) fmacc(acc, conj_(vi[r]), vj[r]);
            acc = warp_sum(acc);
            if (lane == 0) Gs[i * NB + j] = acc;
        }
        __syncthreads();
        if (warp == CTA_WARPS - 1) {
            for (int i = lane; i < NB * NB; i += 32) Ts[i] = mk<T>(0.0);
            __syncwarp();
            if (lane < jb) {
                const int l = lane;
                T* Trow = Ts + l * NB;
                Trow[l] = tau[p0 + l];
                for (int i = l + 1; i < jb; ++i) {
                    T acc = mk<T>(0.0);
                    for (int q = l; q < i; ++q) fmacc(acc, Trow[q], Gs[q * NB + i]);
                    Trow[i] = -(tau[p0 + i] * acc);
                }
                if (form_q) {
                    T* Tg = Tw + (size_t)(p0 / NB) * NB * NB;
                    for (int i = 0; i < jb; ++i) Tg[l * NB + i] = (i >= l) ? Trow[i] : mk<T>(0.0);
                }
            }
        }
        __syncthreads();
        // ---- (h) trailing matrix:  A2 <- (I - V T^H V^H) A2  on the tensor cores
        if (kend < n) {
            cta_panel_dot<T>(Ws, Ps, ld, jb, A, lda, p0, n, kend, n, n - 1);           // W = V^H A2
            __syncthreads();
            for (int c = kend + tid; c < n; c += CTA_THREADS) {                         // W <- T^H W  (T^H lower triangular: bottom-up)
                for (int l = jb - 1; l >= 0; --l) {
                    T acc = mk<T>(0.0);
                    for (int q = 0; q <= l; ++q) fmacc(acc, conj_(Ts[q * NB + l]), Ws[(size_t)q * ld + c]);
                    Ws[(size_t)l * ld + c] = acc;
                }
                for (int l = jb; l < ((jb + 7) & ~7) && l < NB; ++l) Ws[(size_t)l * ld + c] = mk<T>(0.0);
            }
            __syncthreads();
            cta_rank_update<T, false>(A, lda, p0, n, kend, n, Ps, Ws, ld, jb, n - 1);   // A2 -= V W
            __syncthreads();
        }
    }

    // ---------------------------------------------------------------- d = |diag R'|, R = D^-1 R' P^T, taus
    {
        double* d = p.d + (long long)b * p.stride_d;
        T* R = p.R + (long long)b * p.strideR;
        for (int r = tid; r < n; r += CTA_THREADS) {
            const double dr = abs_(A[(long long)r * lda + r]);
            d[r] = dr;
            for (int c = 0; c < n; ++c) {
                T v = (c >= r) ? A[(long long)c * lda + r] / dr : mk<T>(0.0);
                R[(long long)jpvt[c] * p.ldr + r] = v;
            }
        }
        if (p.jpvt_out)
            for (int c = tid; c < n; c += CTA_THREADS) p.jpvt_out[(long long)b * n + c] = jpvt[c] + 1;
        if (!form_q) {   // Q is formed by orgq_dmma_batched (real, n <= 256): it wants the taus as a plain vector
            if constexpr (!is_cplx<T>::value)
                for (int r = tid; r < n; r += CTA_THREADS) Tw[r] = tau[r];
        }
    }
    __syncthreads();
    if (!form_q) return;

    // ---------------------------------------------------------------- Q = H_1 ... H_n (backward, blocked; as sla_ldr.cu)
    const int npan = (n + NB - 1) / NB;
    for (int pi = npan - 1; pi >= 0; --pi) {
        const int p0 = pi * NB, jb = min(NB, n - p0), kend = p0 + jb;
        for (int i = 0; i < NB; ++i) {
            const int c = p0 + i;
            for (int r = tid; r < n; r += CTA_THREADS) {
                T v = mk<T>(0.0);
                if (i < jb) {
                    if (r > c) v = A[(long long)c * lda + r];
                    else if (r == c) v = mk<T>(1.0);
                }
                Ps[(size_t)i * ld + r] = v;
            }
        }
        {
            const T* Tg = Tw + (size_t)pi * NB * NB;
            for (int i = tid; i < NB * NB; i += CTA_THREADS) {
                int l = i / NB, q = i % NB;
                Ts[i] = (l < jb && q < jb) ? Tg[i] : mk<T>(0.0);
            }
        }
        __syncthreads();
        for (int i = 0; i < jb; ++i) {
            const int c = p0 + i;
            for (int r = tid; r < n; r += CTA_THREADS) A[(long long)c * lda + r] = mk<T>(r == c ? 1.0 : 0.0);
        }
        for (int c = kend + warp; c < n; c += CTA_WARPS)
            for (int r = p0 + lane; r < kend; r += 32) A[(long long)c * lda + r] = mk<T>(0.0);
        __syncthreads();
        cta_panel_dot<T>(Ws, Ps, ld, jb, A, lda, p0, n, p0, n, n - 1);
        __syncthreads();
        for (int c = p0 + tid; c < n; c += CTA_THREADS) {
            for (int l = 0; l < jb; ++l) {
                T acc = mk<T>(0.0);
                for (int q = l; q < jb; ++q) fmacc(acc, Ts[l * NB + q], Ws[(size_t)q * ld + c]);
                Ws[(size_t)l * ld + c] = acc;
            }
            for (int l = jb; l < ((jb + 3) & ~3); ++l) Ws[(size_t)l * ld + c] = mk<T>(0.0);
        }
        __syncthreads();
        cta_rank_update<T, false>(A, lda, p0, n, p0, n, Ps, Ws, ld, jb, n - 1);
        __syncthreads();
    }
}

template <typename T> int pp_pick_nb(int n) {
    const size_t limit = 227 * 1024;
    if (pp_smem_bytes<T>(n, 32) <= limit) return 32;
    if (pp_smem_bytes<T>(n, 16) <= limit) return 16;
    if (pp_smem_bytes<T>(n, 8) <= limit) return 8;
    return 0;
}

template <typename T, int NB> cudaError_t launch_pp(const LdrArgs<T>& p, int form_q, cudaStream_t stream) {
    auto kern = ldr_pp_kernel<T, NB>;
    const size_t smem = pp_smem_bytes<T>(p.n, NB);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<p.batch, CTA_THREADS, smem, stream>>>(p, form_q);
    ++g_launch_count;
    return cudaGetLastError();
}

}  // namespace

// the T factors of the in-kernel Q formation share the workspace sizing of the streaming kernel (ldr_twork_elems)
template <typename T> bool ldr_pp_supported(int n) {
    const int nb = pp_pick_nb<T>(n);
    return nb > 0 && n <= 2040 && (size_t)((n + nb - 1) / nb) * nb * nb <= ldr_twork_elems<T>(n);
}

template <typename T> cudaError_t ldr_pp_batched(const LdrArgs<T>& p, cudaStream_t stream) {
    if (p.batch <= 0 || p.n <= 0) return cudaSuccess;
    if (!ldr_pp_supported<T>(p.n) || p.Twork == nullptr) return cudaErrorNotSupported;
    // real, n <= 256: Q on the tensor cores by the separate launch (sla_orgq.cu); otherwise inside the kernel
    bool dmma_q = false;
    if constexpr (!is_cplx<T>::value) dmma_q = (p.n <= 256) && (p.strideT >= p.n);
    const int form_q = dmma_q ? 0 : 1;
    cudaError_t e;
    switch (pp_pick_nb<T>(p.n)) {
        case 32: e = launch_pp<T, 32>(p, form_q, stream); break;
        case 16: e = launch_pp<T, 16>(p, form_q, stream); break;
        case 8: e = launch_pp<T, 8>(p, form_q, stream); break;
        default: return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    if constexpr (!is_cplx<T>::value) {
        if (dmma_q) return orgq_dmma_batched(p, stream);
    }
    return cudaSuccess;
}

template bool ldr_pp_supported<double>(int);
template bool ldr_pp_supported<cplx>(int);
template cudaError_t ldr_pp_batched<double>(const LdrArgs<double>&, cudaStream_t);
template cudaError_t ldr_pp_batched<cplx>(const LdrArgs<cplx>&, cudaStream_t);

}  // namespace sla
