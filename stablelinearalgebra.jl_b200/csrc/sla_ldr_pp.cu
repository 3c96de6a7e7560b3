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
#include <cstdio>
#include <cstdlib>

#include "sla_cta.cuh"
#include "sla_kernels.h"

// -DSLA_PP_PROFILE: per-phase cycle counters of CTA 0 (thread 0), printed at the end of the kernel
#ifdef SLA_PP_PROFILE
#define PP_DECL long long pp_[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long ppt_ = clock64();
#define PP(i) do { long long t_ = clock64(); pp_[i] += t_ - ppt_; ppt_ = t_; } while (0)
#define PP_DUMP() do { if (blockIdx.x == 0 && threadIdx.x == 0) printf("ldr_pp_kernel n=%d NB=%d: norms %lld | select %lld | gather+move %lld | panel pivot+reflector %lld | panel apply %lld | write back %lld | gram+T %lld | W=V^H A %lld | T^H W %lld | A -= V W %lld | R extract %lld | Q %lld\n", n, NB, pp_[0], pp_[1], pp_[2], pp_[3], pp_[4], pp_[5], pp_[6], pp_[7], pp_[8], pp_[9], pp_[10], pp_[11]); } while (0)
#else
#define PP_DECL
#define PP(i) do { } while (0)
#define PP_DUMP() do { } while (0)
#endif

namespace sla {

namespace {

template <typename T> __host__ __device__ inline size_t pp_smem_bytes(int n, int NB) {
    const size_t ld = panel_ld<T>(n);
    size_t b = 2 * NB * ld * sizeof(T);                 // panel (V) and W / staging
    b += 2 * (size_t)NB * NB * sizeof(T);               // Gram, T
    b += (size_t)n * sizeof(T);                         // tau
    b += (size_t)n * sizeof(unsigned long long);        // selection keys
    size_t np2 = 64; while (np2 < (size_t)n) np2 <<= 1;
    b += np2 * sizeof(unsigned long long);              // bitonic sort buffer
    b += (size_t)n * sizeof(double);                    // squared partial norms accumulated by the trailing update
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

// cta_rank_update (sla_cta.cuh) + the squared column norms of the UPDATED values in rows >= nrow_lo, accumulated into
// vnacc[c] (shared memory, zeroed by the caller): the next panel's pivot norms come for free, from exactly the numbers
// that were stored.
template <typename T>
__device__ void pp_rank_update_norms(T* __restrict__ C, int ldc, int r_lo, int r_hi, int c_lo, int c_hi, const T* __restrict__ Ps,
                                     const T* __restrict__ Us, int ld, int kb, int n_clamp, int nrow_lo, double* vnacc) {
    if (r_lo >= r_hi || c_lo >= c_hi) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fi = lane >> 2, fk = lane & 3;
    const int kb4 = (kb + 3) & ~3;
    const int r_base = r_lo & ~1;
    constexpr int NI = (sizeof(T) == 8) ? 4 : 2;
    constexpr int TC = NI * 8;
    const int tiles_r = (r_hi - r_base + 31) / 32, tiles_c = (c_hi - c_lo + TC - 1) / TC;
    const bool vec = (sizeof(T) == 8) && ((ldc & 1) == 0) && ((((uintptr_t)C) & 15) == 0);
    for (int t = warp; t < tiles_r * tiles_c; t += CTA_WARPS) {
        const int r0 = r_base + (t % tiles_r) * 32, c0 = c_lo + (t / tiles_r) * TC;
        T acc[NI][4][2];
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
            for (int i = 0; i < NI; ++i) nf[i] = -Us[(kk + fk) * ld + min(c0 + i * 8 + fi, n_clamp)];
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
            double ss = 0.0;
            if (c < c_hi) {
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
                    if (r >= nrow_lo && r < r_hi) ss += abs2_(acc[i][j][0]);
                    if (r + 1 >= nrow_lo && r + 1 < r_hi) ss += abs2_(acc[i][j][1]);
                }
            }
            ss += __shfl_xor_sync(0xffffffffu, ss, 1);
            ss += __shfl_xor_sync(0xffffffffu, ss, 2);
            if (fk == 0 && c < c_hi) atomicAdd(&vnacc[c], ss);
        }
    }
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
    int np2 = 64; while (np2 < n) np2 <<= 1;
    unsigned long long* skeys = keys + n;                        // [np2] bitonic sort buffer
    unsigned long long* red = skeys + np2;                       // [32]
    double* beta = reinterpret_cast<double*>(red + 32);          // [NB]
    double* pnf = beta + NB;                                     // [NB] squared norm of a panel column, rows >= k
    double* pnt = pnf + NB;                                      // [NB] ... rows > k
    double* vnacc = pnt + NB;                                    // [n] squared norms (rows >= kend) from the trailing update
    int* jpvt = reinterpret_cast<int*>(vnacc + n);               // [n]
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
    PP_DECL

    for (int p0 = 0; p0 < n; p0 += NB) {
        const int jb = min(NB, n - p0), kend = p0 + jb;
        PP(9);
        // ---- (a) exact squared partial norms (rows >= p0) of every remaining column -> selection keys.  Panel 0 makes a
        //      pass over the input (and copies it into L); later panels take the sums the trailing update accumulated
        //      from the very values it stored (pp_rank_update_norms)
        if (p0 == 0) {
            for (int c = warp; c < n; c += CTA_WARPS) {
                const T* col = Ain ? Ain + (long long)c * p.ldain : A + (long long)c * lda;
                T* dcol = A + (long long)c * lda;
                double acc = 0.0;
                bool nz = false;
                for (int r0 = lane; r0 < n; r0 += 256) {          // eight independent loads in flight
                    T x[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) { const int r = r0 + 32 * u; x[u] = (r < n) ? col[r] : mk<T>(0.0); }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int r = r0 + 32 * u;
                        if (r < n) { if (Ain) dcol[r] = x[u]; acc += abs2_(x[u]); nz |= !is_zero(x[u]); }
                    }
                }
                acc = warp_sum(acc);
                nz = __any_sync(0xffffffffu, nz);
                if (lane == 0 && ldr_norm2_out_of_range(acc, nz)) ldr_raise_range(p.info, b);
                if (lane == 0) keys[c] = pp_key(acc, c);
            }
        } else {
            for (int c = p0 + tid; c < n; c += CTA_THREADS) keys[c] = pp_key(vnacc[c], c);
        }
        __syncthreads();
        PP(0);
        // ---- (b) the jb largest columns become the panel (slot i = i-th largest): bitonic sort of the keys, descending
        {
            int m = 64; while (m < n - p0) m <<= 1;
            for (int i = tid; i < m; i += CTA_THREADS) skeys[i] = (p0 + i < n) ? keys[p0 + i] : 0ull;
            __syncthreads();
            for (int k2 = 2; k2 <= m; k2 <<= 1) {
                for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
                    for (int i = tid; i < m; i += CTA_THREADS) {
                        const int ixj = i ^ j2;
                        if (ixj > i) {
                            const unsigned long long a = skeys[i], c = skeys[ixj];
                            const bool desc = ((i & k2) == 0);
                            if (desc ? (a < c) : (a > c)) { skeys[i] = c; skeys[ixj] = a; }
                        }
                    }
                    __syncthreads();
                }
            }
            if (tid < jb) {
                const unsigned long long win = skeys[tid];
                int c = pp_key_index(win);
                if (win == 0ull || c < p0 || c >= n) c = p0 + tid;    // degenerate input (NaNs everywhere): natural order
                sel[tid] = c;
            }
            __syncthreads();
        }
        PP(1);
        // ---- (c) gather the selected columns (all n rows) into the panel slots; make room at positions p0 .. kend - 1
        for (int e = tid; e < jb * n; e += CTA_THREADS) {
            const int i = e / n, r = e - i * n;
            Ps[(size_t)i * ld + r] = A[(long long)sel[i] * lda + r];
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
            for (int e = tid; e < nd * n; e += CTA_THREADS) {         // displaced column -> the slot a selected column left
                const int t = e / n, r = e - t * n;
                A[(long long)vac[t] * lda + r] = A[(long long)disp[t] * lda + r];
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
        PP(2);
        // ---- (d) panel factorisation in shared memory, greedy pivoting among the panel's columns (?laqp2 on exact norms).
        //      Every warp picks the pivot redundantly from the same norms (no barrier for the choice), the unpivoted
        //      slots are a register bit mask, the pivot's row-k element stays in place (the implicit 1 of v is handled
        //      where v is used): two CTA barriers per column.
        unsigned actmask = (jb >= 32) ? 0xffffffffu : ((1u << jb) - 1u);
        for (int j = 0; j < jb; ++j) {
            const int k = p0 + j;
            unsigned long long key = (lane < jb && ((actmask >> lane) & 1u)) ? (pp_key(pnf[lane], lane) | 1ull << 63) : 0ull;
            key = pp_warp_max(key);
            const int sj = pp_key_index(key);
            actmask &= ~(1u << sj);
            T* x = Ps + (size_t)sj * ld;
            const T alpha = x[k];
            const double xnorm2 = pnt[sj];
            // LAPACK ?larfg, by one lane per warp (the FP64 square root and the two divisions are long software sequences:
            // 512 redundant copies of them would occupy the FP64 pipe for longer than their latency)
            double sc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};                 // tau.re, tau.im, scal.re, scal.im, beta
            if (lane == 0) {
                T tk0, scal0;
                double bt0;
                if (xnorm2 == 0.0 && imag_(alpha) == 0.0) {
                    tk0 = mk<T>(0.0); scal0 = mk<T>(0.0); bt0 = real_(alpha);
                } else {
                    bt0 = -copysign(sqrt(abs2_(alpha) + xnorm2), real_(alpha));
                    const double ibt = 1.0 / bt0;
                    tk0 = mk<T>((bt0 - real_(alpha)) * ibt, -imag_(alpha) * ibt);
                    scal0 = mk<T>(1.0) / (alpha - mk<T>(bt0));
                }
                sc[0] = real_(tk0); sc[1] = imag_(tk0); sc[2] = real_(scal0); sc[3] = imag_(scal0); sc[4] = bt0;
            }
#pragma unroll
            for (int u = 0; u < 5; ++u) sc[u] = __shfl_sync(0xffffffffu, sc[u], 0);
            const T tk = mk<T>(sc[0], sc[1]), scal = mk<T>(sc[2], sc[3]);
            const double bt = sc[4];
            for (int r = k + 1 + tid; r < n; r += CTA_THREADS) x[r] = x[r] * scal;
            if (tid == 0) { tau[k] = tk; beta[j] = bt; order[j] = sj; }
            __syncthreads();
            PP(3);
            // H^H = I - conj(tau) v v^H on the unpivoted slots: a warp per slot, dot + update + next norms in one go
            const T ctau = conj_(tk);
            {
                // a warp serves slots warp and warp + 16 (NB <= 32) TOGETHER: two independent dependency chains per pass
                const int s0 = warp, s1 = warp + CTA_WARPS;
                const bool on0 = s0 < jb && ((actmask >> s0) & 1u), on1 = s1 < jb && ((actmask >> s1) & 1u);
                if (on0 || on1) {
                    T* a0 = Ps + (size_t)(on0 ? s0 : s1) * ld;
                    T* a1 = Ps + (size_t)(on1 ? s1 : s0) * ld;         // both valid; the second is skipped when !(on0 && on1)
                    const bool two = on0 && on1;
                    T dot0 = mk<T>(0.0), dot1 = mk<T>(0.0);
                    for (int r = k + lane; r < n; r += 32) {
                        const T xv = (r == k) ? mk<T>(1.0) : conj_(x[r]);
                        fmacc(dot0, xv, a0[r]);
                        if (two) fmacc(dot1, xv, a1[r]);
                    }
                    dot0 = warp_sum(dot0);
                    if (two) dot1 = warp_sum(dot1);
                    const T w0 = ctau * dot0, w1 = ctau * dot1;
                    double nf0 = 0.0, nt0 = 0.0, nf1 = 0.0, nt1 = 0.0;
                    for (int r = k + lane; r < n; r += 32) {
                        const T xv = (r == k) ? mk<T>(1.0) : x[r];
                        T v = a0[r];
                        fms(v, xv, w0);
                        a0[r] = v;
                        const double v2 = abs2_(v);
                        if (r > k) nf0 += v2;
                        if (r > k + 1) nt0 += v2;
                        if (two) {
                            T u = a1[r];
                            fms(u, xv, w1);
                            a1[r] = u;
                            const double u2 = abs2_(u);
                            if (r > k) nf1 += u2;
                            if (r > k + 1) nt1 += u2;
                        }
                    }
                    nf0 = warp_sum(nf0); nt0 = warp_sum(nt0);
                    if (two) { nf1 = warp_sum(nf1); nt1 = warp_sum(nt1); }
                    if (lane == 0) {
                        const int sa = on0 ? s0 : s1;
                        pnf[sa] = nf0; pnt[sa] = nt0;
                        if (two) { pnf[s1] = nf1; pnt[s1] = nt1; }
                    }
                }
            }
            __syncthreads();
            PP(4);
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
                vj[r] = (r < k) ? mk<T>(0.0) : ((r == k) ? mk<T>(1.0) : v);
            }
        }
        for (int i = jb * ld + tid; i < NB * ld; i += CTA_THREADS) Ws[i] = mk<T>(0.0);
        __syncthreads();
        { T* t = Ps; Ps = Ws; Ws = t; }                               // Ps = V (position order), Ws free
        for (int i = jb * ld + tid; i < NB * ld; i += CTA_THREADS) Ws[i] = mk<T>(0.0);
        PP(5);
        // ---- (g) Gram matrix of the reflectors and the compact-WY T (?larft):  T(l,l) = tau_l, T(l,i) = -tau_i sum_q T(l,q) G(q,i)
        // G = V^H V on the tensor cores, both operands from the shared panel: one 8 x 8 tile of G per warp and trip
        {
            constexpr int TG = NB / 8;
            const int fi = lane >> 2, fk = lane & 3;
            const int r_lo = p0 & ~3;
            for (int t = warp; t < TG * TG; t += CTA_WARPS) {
                const int ti = t / TG, tj = t % TG;
                if (ti > tj) continue;                                  // only the upper triangle is used
                const T* vi = Ps + (size_t)(ti * 8 + fi) * ld;
                const T* vj = Ps + (size_t)(tj * 8 + fi) * ld;
                T acc[2] = {mk<T>(0.0), mk<T>(0.0)};
                for (int r = r_lo; r < n; r += 4) {
                    const int rr = min(r + fk, ld - 1);
                    const T af = (r + fk < n) ? conj_(vi[rr]) : mk<T>(0.0);
                    const T bf = (r + fk < n) ? vj[rr] : mk<T>(0.0);
                    mma_acc(acc, af, bf);                               // D[i = fi][j = 2 fk + e] += conj(V_i[r]) V_j[r]
                }
                Gs[(ti * 8 + fi) * NB + tj * 8 + 2 * fk] = acc[0];
                Gs[(ti * 8 + fi) * NB + tj * 8 + 2 * fk + 1] = acc[1];
            }
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
        PP(6);
        // ---- (h) trailing matrix:  A2 <- (I - V T^H V^H) A2  on the tensor cores
        if (kend < n) {
            cta_panel_dot<T>(Ws, Ps, ld, jb, A, lda, p0, n, kend, n, n - 1);           // W = V^H A2
            __syncthreads();
            PP(7);
            for (int c = kend + tid; c < n; c += CTA_THREADS) {                         // W <- T^H W  (T^H lower triangular: bottom-up)
                for (int l = jb - 1; l >= 0; --l) {
                    T ac[4] = {mk<T>(0.0), mk<T>(0.0), mk<T>(0.0), mk<T>(0.0)};   // four chains: the sum is latency-bound
                    int q = 0;
                    for (; q + 3 <= l; q += 4) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) fmacc(ac[u], conj_(Ts[(q + u) * NB + l]), Ws[(size_t)(q + u) * ld + c]);
                    }
                    for (; q <= l; ++q) fmacc(ac[0], conj_(Ts[q * NB + l]), Ws[(size_t)q * ld + c]);
                    Ws[(size_t)l * ld + c] = (ac[0] + ac[1]) + (ac[2] + ac[3]);
                }
                for (int l = jb; l < ((jb + 7) & ~7) && l < NB; ++l) Ws[(size_t)l * ld + c] = mk<T>(0.0);
            }
            __syncthreads();
            PP(8);
            for (int c = kend + tid; c < n; c += CTA_THREADS) vnacc[c] = 0.0;
            __syncthreads();
            pp_rank_update_norms<T>(A, lda, p0, n, kend, n, Ps, Ws, ld, jb, n - 1, kend, vnacc);   // A2 -= V W, norms of rows >= kend
            __syncthreads();
        }
    }
    PP(9);

    // ---------------------------------------------------------------- d = |diag R'|, R = D^-1 R' P^T, taus
    {
        double* d = p.d + (long long)b * p.stride_d;
        T* R = p.R + (long long)b * p.strideR;
        double* dall = vnacc;                                          // the norms are dead: d of every position
        for (int r = tid; r < n; r += CTA_THREADS) { const double dr = abs_(A[(long long)r * lda + r]); dall[r] = dr; d[r] = dr; }
        __syncthreads();
        for (int c = warp; c < n; c += CTA_WARPS) {                    // position c holds original column jpvt[c]
            const T* src = A + (long long)c * lda;
            T* dst = R + (long long)jpvt[c] * p.ldr;
            for (int r = lane; r < n; r += 32) dst[r] = (r <= c) ? src[r] / dall[r] : mk<T>(0.0);
        }
        if (p.jpvt_out)
            for (int c = tid; c < n; c += CTA_THREADS) p.jpvt_out[(long long)b * n + c] = jpvt[c] + 1;
        if (!form_q) {   // Q is formed by orgq_dmma_batched (real, n <= 256): it wants the taus as a plain vector
            if constexpr (!is_cplx<T>::value)
                for (int r = tid; r < n; r += CTA_THREADS) Tw[r] = tau[r];
        }
    }
    __syncthreads();
    PP(10);
    if (!form_q) { PP_DUMP(); return; }

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
    PP(11);
    PP_DUMP();
}

// Panel width: as wide as the shared memory allows, at most 32; real matrices of n <= 256 take 16 (measured faster: C2 step
// 31.8 ms against 33.1 with NB = 32 and 35.2 with NB = 8; N = 400: 16 and 32 equal).  The width does NOT decide the accuracy:
// on the C1-size sweep the G(0) error of this mode is 5e-12 .. 3e-9 for NB = 8, 16 and 32 alike (strict path: 6e-13 .. 3e-11;
// profiles/ldr_panel_r2.txt).  SLA_PP_NB=8|16|32 overrides (experiments).
template <typename T> int pp_pick_nb(int n) {
    const size_t limit = 227 * 1024;
    static const int forced = getenv("SLA_PP_NB") ? atoi(getenv("SLA_PP_NB")) : 0;
    int cap = (!is_cplx<T>::value && n <= 256) ? 16 : 32;
    if (forced == 8 || forced == 16 || forced == 32) cap = forced;
    if (cap >= 32 && pp_smem_bytes<T>(n, 32) <= limit) return 32;
    if (cap >= 16 && pp_smem_bytes<T>(n, 16) <= limit) return 16;
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
