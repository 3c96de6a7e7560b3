// LDR factorisation of SMALL matrices (n <= 64): one CTA per matrix, the matrix in registers, ONE barrier per column.
//
// Replaces `ldr!(F, ws)` (src/LDR.jl:160-188: geqp3! src/qr.jl:70-92, triu/abs/ldiv!/mul_P! LDR.jl:169-182, orgqr!
// src/qr.jl:95-109) for the sizes where the cluster kernels' per-column machinery (DSMEM candidate exchange, mbarriers)
// costs more than the arithmetic: BASELINE config C1 (N=64, one chain) spends 10 of its ~100 launches and 70 % of its
// time in `ldr!` of a single 64 x 64 matrix.
//
// Shape: thread (c, q) = (tid / 4, tid % 4) holds rows q, q+4, q+8, ... of column c (NMAX/4 values), so a dot product
// over a column is an in-thread FMA chain plus TWO shuffle rounds inside a group of 4 lanes.  Per column k, with one
// __syncthreads:
//   1. every warp picks the pivot redundantly from the published keys (exact squared partial norms; three `redux`
//      warp reductions on the key's bit pattern; first maximal index wins);
//   2. every thread reads the pivot column from its mirror in shared memory together with the reflector scalars
//      (beta, tau, 1/(alpha - beta): LAPACK ?larfg) that the column's own group published speculatively, forms
//      v = (0, 1, x * scal) and applies H^H = I - conj(tau) v v^H to its column (?laqp2: ?larf with conj(tau));
//   3. every live column recomputes its partial norms EXACTLY from the registers (no down-dating, hence no
//      cancellation safeguard), computes the reflector it would generate at step k+1 and publishes key, scalars and
//      the column itself (double-buffered by the parity of k; a retired column's mirror is never written again and
//      doubles as the reflector store for the Q formation).
// No physical column swaps: a column retired at step k keeps R'(0..k, k) in its registers and goes to its ORIGINAL
// column of R = D^-1 R' P^T.  Q = H_0 ... H_{n-1} I needs no barrier at all: every thread applies the reflectors to its
// slice of a column of the identity, reading them from the mirror.
#include <stdlib.h>

#include "sla_kernels.h"

namespace sla {

namespace {

constexpr unsigned FULL = 0xffffffffu;

template <typename T> struct SmallPad { static constexpr int value = 2; };        // 16 bytes between lane slices
template <> struct SmallPad<cplx> { static constexpr int value = 1; };

template <typename T, int NMAX> struct SmallCfg {
    static constexpr int RPL = NMAX / 4;                       // rows per lane: row = 4 i + q
    static constexpr int THREADS = NMAX * 4;
    static constexpr int QLD = RPL + SmallPad<T>::value;       // elements between the slices of two lanes of a group
    static constexpr int CLD = 4 * QLD;                        // elements between two columns of the mirror
    static constexpr int KPL = NMAX / 32;                      // pivot keys per lane
    static constexpr size_t SMEM = (size_t)NMAX * CLD * sizeof(T)        // mirror
                                   + (size_t)6 * NMAX * sizeof(T)        // sc_scal[2], sc_tau[2], tau_s, scal_s
                                   + (size_t)5 * NMAX * sizeof(double)   // keys[2], sc_beta[2], d_s
                                   + (size_t)NMAX * sizeof(int);         // jpvt_s
};

__device__ __forceinline__ double shx(double v, int m) { return __shfl_xor_sync(FULL, v, m); }
__device__ __forceinline__ cplx shx(cplx v, int m) { return cplx{shx(v.re, m), shx(v.im, m)}; }
// value of lane `src` of the caller's group of 4
__device__ __forceinline__ double shq(double v, int src) { return __shfl_sync(FULL, v, src, 4); }
__device__ __forceinline__ cplx shq(cplx v, int src) { return cplx{shq(v.re, src), shq(v.im, src)}; }
template <typename T> __device__ __forceinline__ T group_sum(T v) {
    v = v + shx(v, 1);
    return v + shx(v, 2);
}

template <typename T, int RPL> __device__ __forceinline__ void load_slice(T (&x)[RPL], const T* __restrict__ src) {
    if constexpr (is_cplx<T>::value) {
#pragma unroll
        for (int i = 0; i < RPL; ++i) x[i] = src[i];
    } else {
#pragma unroll
        for (int i = 0; i < RPL; i += 2) {
            const double2 t = *reinterpret_cast<const double2*>(src + i);
            x[i] = t.x; x[i + 1] = t.y;
        }
    }
}
template <typename T, int RPL> __device__ __forceinline__ void store_slice(T* __restrict__ dst, const T (&x)[RPL]) {
    if constexpr (is_cplx<T>::value) {
#pragma unroll
        for (int i = 0; i < RPL; ++i) dst[i] = x[i];
    } else {
#pragma unroll
        for (int i = 0; i < RPL; i += 2) *reinterpret_cast<double2*>(dst + i) = make_double2(x[i], x[i + 1]);
    }
}

// v = (0 above row k, 1 in row k, x * scal below) for the caller's rows 4 i + q
template <typename T, int RPL>
__device__ __forceinline__ void make_reflector(T (&v)[RPL], int q, int k, T scal) {
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
        const int r = 4 * i + q;
        v[i] = (r > k) ? v[i] * scal : mk<T>(r == k ? 1.0 : 0.0);
    }
}
template <typename T, int RPL> __device__ __forceinline__ T slice_dot(const T (&v)[RPL], const T (&a)[RPL]) {
    T s0 = mk<T>(0.0), s1 = mk<T>(0.0);
#pragma unroll
    for (int i = 0; i < RPL; i += 2) {
        fmacc(s0, conj_(v[i]), a[i]);
        fmacc(s1, conj_(v[i + 1]), a[i + 1]);
    }
    return s0 + s1;
}

template <typename T, int NMAX>
__global__ void __launch_bounds__(NMAX * 4) ldr_small_kernel(LdrArgs<T> p) {
    using C = SmallCfg<T, NMAX>;
    constexpr int RPL = C::RPL, QLD = C::QLD, CLD = C::CLD;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* mirror = reinterpret_cast<T*>(smem_raw);
    T* sc_scal = mirror + (size_t)NMAX * CLD;   // [2][NMAX]  1 / (alpha - beta) of the reflector a column would generate
    T* sc_tau = sc_scal + 2 * NMAX;             // [2][NMAX]
    T* tau_s = sc_tau + 2 * NMAX;               // [NMAX] by step
    T* scal_s = tau_s + NMAX;                   // [NMAX] by step
    double* keys = reinterpret_cast<double*>(scal_s + NMAX);   // [2][NMAX]  squared partial norm, -1 = retired
    double* sc_beta = keys + 2 * NMAX;          // [2][NMAX]
    double* d_s = sc_beta + 2 * NMAX;           // [NMAX] |beta| by step
    int* jpvt_s = reinterpret_cast<int*>(d_s + NMAX);          // [NMAX] column retired at step k

    const int n = p.n, b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31;
    const int c = tid >> 2, q = tid & 3;
    const bool col_ok = c < n;
    T* const mycol = mirror + (size_t)c * CLD + q * QLD;

    // ---------------------------------------------------------------- load
    T a[RPL];
    {
        const T* src = p.Ain ? p.Ain + (long long)b * p.strideAin + (long long)c * p.ldain
                             : p.L + (long long)b * p.strideL + (long long)c * p.ldl;
        double ss = 0.0;
        int nz = 0;
#pragma unroll
        for (int i = 0; i < RPL; ++i) {
            const int r = 4 * i + q;
            a[i] = (col_ok && r < n) ? src[r] : mk<T>(0.0);
            ss += abs2_(a[i]);
            nz |= !is_zero(a[i]);
        }
        ss = group_sum(ss);
        nz |= __shfl_xor_sync(FULL, nz, 1);
        nz |= __shfl_xor_sync(FULL, nz, 2);
        if (q == 0 && col_ok && ldr_norm2_out_of_range(ss, nz != 0)) ldr_raise_range(p.info, b);
    }
    bool live = col_ok;
    int mypos = -1;   // step at which this column was retired

    // exact partial norms below row kn, the reflector this column would generate at step kn, and the column itself
    auto publish = [&](int kn) {
        double p0 = 0.0, p1 = 0.0, e = 0.0;
        T alpha = mk<T>(0.0);
#pragma unroll
        for (int i = 0; i < RPL; ++i) {
            const int r = 4 * i + q;
            const double m = abs2_(a[i]);
            if (r > kn) { if (i & 1) p1 += m; else p0 += m; }
            if (r == kn) { e = m; alpha = a[i]; }
        }
        const double xn2 = group_sum(p0 + p1);           // bit-identical in the four lanes
        const int src = kn & 3;
        e = shq(e, src);
        alpha = shq(alpha, src);
        T tau, scal;
        double beta;
        if (xn2 == 0.0 && imag_(alpha) == 0.0) {         // ?larfg: H = I
            tau = mk<T>(0.0); scal = mk<T>(0.0); beta = real_(alpha);
        } else {
            beta = -copysign(sqrt(e + xn2), real_(alpha));
            tau = mk<T>((beta - real_(alpha)) / beta, -imag_(alpha) / beta);
            scal = mk<T>(1.0) / (alpha - mk<T>(beta));
        }
        if (q == 0) {
            const int o = (kn & 1) * NMAX + c;
            keys[o] = live ? xn2 + e : -1.0;
            sc_scal[o] = scal; sc_tau[o] = tau; sc_beta[o] = beta;
        }
        if (live) store_slice<T, RPL>(mycol, a);
    };

    publish(0);
    __syncthreads();

    // ---------------------------------------------------------------- pivoted QR
#pragma unroll 1
    for (int k = 0; k < n; ++k) {
        const int par = (k & 1) * NMAX;
        // 1. pivot: largest key, first index among equals
        double bv = -2.0;
        int bi = 0;
#pragma unroll
        for (int j = 0; j < C::KPL; ++j) {
            const double v = keys[par + lane + 32 * j];
            if (v > bv) { bv = v; bi = lane + 32 * j; }
        }
        const int hi = __double2hiint(bv);
        const unsigned lo = (unsigned)__double2loint(bv);
        const int m1 = __reduce_max_sync(FULL, hi);
        const unsigned m2 = __reduce_max_sync(FULL, hi == m1 ? lo : 0u);
        int pvt = __reduce_min_sync(FULL, (hi == m1 && lo == m2) ? bi : 0x7fffffff);
        pvt = min(pvt, NMAX - 1);
        // 2. its reflector, applied to this thread's column
        const T scal = sc_scal[par + pvt], tau = sc_tau[par + pvt];
        const double beta = sc_beta[par + pvt];
        T v[RPL];
        load_slice<T, RPL>(v, mirror + (size_t)pvt * CLD + q * QLD);
        make_reflector<T, RPL>(v, q, k, scal);
        const T w = conj_(tau) * group_sum(slice_dot<T, RPL>(v, a));
        if (c == pvt) {
            live = false;
            mypos = k;
#pragma unroll
            for (int i = 0; i < RPL; ++i)
                if (4 * i + q == k) a[i] = mk<T>(beta);
            if (q == 0) { tau_s[k] = tau; scal_s[k] = scal; d_s[k] = fabs(beta); jpvt_s[k] = pvt; }
        } else if (live) {
#pragma unroll
            for (int i = 0; i < RPL; ++i) fms(a[i], v[i], w);
        }
        // 3. keys, candidate reflectors and columns for the next step
        if (k + 1 < n) publish(k + 1);
        __syncthreads();
    }

    // ---------------------------------------------------------------- d = |diag R'|, R = D^-1 R' P^T
    {
        T* R = p.R + (long long)b * p.strideR + (long long)c * p.ldr;
        if (col_ok) {
#pragma unroll
            for (int i = 0; i < RPL; ++i) {
                const int r = 4 * i + q;
                if (r < n) R[r] = (r <= mypos) ? a[i] / d_s[r] : mk<T>(0.0);
            }
        }
        if (tid < n) {
            p.d[(long long)b * p.stride_d + tid] = d_s[tid];
            if (p.jpvt_out) p.jpvt_out[(long long)b * n + tid] = jpvt_s[tid] + 1;
        }
    }

    // ---------------------------------------------------------------- Q = H_0 ... H_{n-1} I
    {
#pragma unroll
        for (int i = 0; i < RPL; ++i) a[i] = mk<T>((4 * i + q == c) ? 1.0 : 0.0);
        // H_k e_c = e_c for k > c: the warp starts at its largest column
        const int kstart = min(n - 1, 8 * (tid >> 5) + 7);
        T x[RPL];
        load_slice<T, RPL>(x, mirror + (size_t)jpvt_s[kstart] * CLD + q * QLD);
#pragma unroll 1
        for (int k = kstart; k >= 0; --k) {
            T v[RPL];
#pragma unroll
            for (int i = 0; i < RPL; ++i) v[i] = x[i];
            if (k > 0) load_slice<T, RPL>(x, mirror + (size_t)jpvt_s[k - 1] * CLD + q * QLD);   // next reflector in flight
            make_reflector<T, RPL>(v, q, k, scal_s[k]);
            const T w = tau_s[k] * group_sum(slice_dot<T, RPL>(v, a));
#pragma unroll
            for (int i = 0; i < RPL; ++i) fms(a[i], v[i], w);
        }
        T* Q = p.L + (long long)b * p.strideL + (long long)c * p.ldl;
        if (col_ok) {
#pragma unroll
            for (int i = 0; i < RPL; ++i) {
                const int r = 4 * i + q;
                if (r < n) Q[r] = a[i];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Version 2 (SLA_LDR_SMALL=2): the same algorithm with the ROW BLOCK of the current column as a compile-time constant.
//
// Version 1 above spends ~550 instructions per column and warp, most of them compares and selects: a thread's 16 rows
// live in registers, so "rows above / on / below the diagonal" is decided per element and per column.  Here the column
// loop is unrolled over KB = k / 4 (row = 4 i + q: the elements i < KB of every thread lie above row k, the elements
// i > KB below it, and only element KB depends on the lane: q <, ==, > k % 4), so
//   * dead rows cost nothing (the vector work of a column shrinks with k: half on average),
//   * live rows are plain FMAs without predicates; one boundary element per thread carries the case distinction,
//   * only the rows >= 4 KB of a column are mirrored / read back.
// The reflector scalars are no longer computed speculatively by every column: once the pivot is known every thread
// forms beta, tau and 1 / (alpha - beta) of THAT column from its published norm and its diagonal element, and the
// square root and the reciprocals overlap with the raw dot product  s = sum_{r > k} conj(x_r) a_r, which does not
// need them:  v^H a = a_k + conj(scal) s.  Loop body per column: publish norms and column -> barrier -> pivot -> apply.
// The Q formation runs the same blocked form backwards (no barrier).
// ---------------------------------------------------------------------------------------------
// diagnostics (-DSLA_SMALL_PROFILE, profiles/ldr_small_phases.py): cycles per phase of the column loop, thread 0 of CTA 0
struct SmallProf {
#ifdef SLA_SMALL_PROFILE
    long long t, acc[8];
    __device__ __forceinline__ void start() { t = clock64(); for (int i = 0; i < 8; ++i) acc[i] = 0; }
    __device__ __forceinline__ void stamp(int i) { const long long now = clock64(); acc[i] += now - t; t = now; }
#else
    __device__ __forceinline__ void start() {}
    __device__ __forceinline__ void stamp(int) {}
#endif
};

template <typename T, int RPL, int I0> __device__ __forceinline__ void load_rows(T (&x)[RPL], const T* __restrict__ src) {
    if constexpr (is_cplx<T>::value) {
#pragma unroll
        for (int i = I0; i < RPL; ++i) x[i] = src[i];
    } else {
#pragma unroll
        for (int i = (I0 & ~1); i < RPL; i += 2) {
            const double2 t = *reinterpret_cast<const double2*>(src + i);
            x[i] = t.x; x[i + 1] = t.y;
        }
    }
}
template <typename T, int RPL, int I0> __device__ __forceinline__ void store_rows(T* __restrict__ dst, const T (&x)[RPL]) {
    if constexpr (is_cplx<T>::value) {
#pragma unroll
        for (int i = I0; i < RPL; ++i) dst[i] = x[i];
    } else {
#pragma unroll
        for (int i = (I0 & ~1); i < RPL; i += 2) *reinterpret_cast<double2*>(dst + i) = make_double2(x[i], x[i + 1]);
    }
}

// One reflector H = I - t v v^H, v = (1 in row k = 4 KB + kq, x * scal below), applied to the caller's slice `a` of a
// column (t = conj(tau) during the factorisation, tau during the Q formation); x = the caller's rows of the reflector column.
template <typename T, int RPL, int KB>
__device__ __forceinline__ void apply_reflector(T (&a)[RPL], const T (&x)[RPL], int q, int kq, T t, T scal, bool update) {
    T s0 = mk<T>(0.0), s1 = mk<T>(0.0);
#pragma unroll
    for (int i = KB + 1; i < RPL; ++i) {
        if ((i - KB) & 1) fmacc(s0, conj_(x[i]), a[i]); else fmacc(s1, conj_(x[i]), a[i]);
    }
    T s = s0 + s1;
    if (q > kq) fmacc(s, conj_(x[KB]), a[KB]);
    s = group_sum(s);
    const T ak = shq(a[KB], kq);
    const T w = t * (ak + conj_(scal) * s);
    if (update) {                                       // (the shuffles above are executed by every lane of the warp)
        const T ws = scal * w;
#pragma unroll
        for (int i = KB + 1; i < RPL; ++i) fms(a[i], x[i], ws);
        if (q > kq) fms(a[KB], x[KB], ws);
        else if (q == kq) a[KB] = a[KB] - w;
    }
}

template <typename T, int NMAX, int KB> struct Small2Block {
    using C = SmallCfg<T, NMAX>;
    static constexpr int RPL = C::RPL, QLD = C::QLD, CLD = C::CLD;

    // columns k = 4 KB .. 4 KB + 3 of the pivoted QR
    static __device__ __forceinline__ void factor(T (&a)[RPL], bool& live, int& mypos, int n, int c, int q, int lane,
                                                  T* mirror, T* mycol, T* tau_s, T* scal_s, double* keys, double* xn_s,
                                                  double* d_s, int* jpvt_s, SmallProf& pf) {
#pragma unroll 1
        for (int kq = 0; kq < 4; ++kq) {
            const int k = 4 * KB + kq;
            if (k >= n) break;
            const int par = (k & 1) * NMAX;
            // 1. exact partial norms of this column below / from row k, and the live rows of the column itself
            {
                const double mb = abs2_(a[KB]);
                double p0 = 0.0, p1 = 0.0;
#pragma unroll
                for (int i = KB + 1; i < RPL; ++i) {
                    if ((i - KB) & 1) p0 += abs2_(a[i]); else p1 += abs2_(a[i]);
                }
                double pl = p0 + p1;
                if (q > kq) pl += mb;
                const double xn2 = group_sum(pl);            // bit-identical in the four lanes
                const double e = shq(mb, kq);
                if (q == 0) { keys[par + c] = live ? xn2 + e : -1.0; xn_s[par + c] = xn2; }
                if (live) store_rows<T, RPL, KB>(mycol, a);
            }
            pf.stamp(0);
            __syncthreads();
            pf.stamp(1);
            // 2. pivot: largest key, first index among equals (every warp redundantly)
            double bv = -2.0;
            int bi = 0;
#pragma unroll
            for (int j = 0; j < C::KPL; ++j) {
                const double v = keys[par + lane + 32 * j];
                if (v > bv) { bv = v; bi = lane + 32 * j; }
            }
            const int hi = __double2hiint(bv);
            const unsigned lo = (unsigned)__double2loint(bv);
            const int m1 = __reduce_max_sync(FULL, hi);
            const unsigned m2 = __reduce_max_sync(FULL, hi == m1 ? lo : 0u);
            int pvt = __reduce_min_sync(FULL, (hi == m1 && lo == m2) ? bi : 0x7fffffff);
            pvt = min(pvt, NMAX - 1);
#ifdef SLA_SMALL_PROFILE
            if (pvt == 12345) pf.t += 1;   // (the stamp waits for the pivot)
#endif
            pf.stamp(2);
            // 3. its reflector (LAPACK ?larfg), every thread for itself; x = the caller's rows of the pivot column
            const double xn2p = xn_s[par + pvt];
            const T alpha = mirror[(size_t)pvt * CLD + kq * QLD + KB];
            T x[RPL];
            load_rows<T, RPL, KB>(x, mirror + (size_t)pvt * CLD + q * QLD);
            T tau, scal;
            double beta;
            if (xn2p == 0.0 && imag_(alpha) == 0.0) {        // H = I
                tau = mk<T>(0.0); scal = mk<T>(0.0); beta = real_(alpha);
            } else {
                beta = -copysign(sqrt(abs2_(alpha) + xn2p), real_(alpha));
                tau = mk<T>((beta - real_(alpha)) / beta, -imag_(alpha) / beta);
                scal = mk<T>(1.0) / (alpha - mk<T>(beta));
            }
#ifdef SLA_SMALL_PROFILE
            if (real_(scal) == 12345.678) pf.t += 1;   // (the stamp waits for the scalars)
#endif
            pf.stamp(3);
            apply_reflector<T, RPL, KB>(a, x, q, kq, conj_(tau), scal, live && c != pvt);   // ?laqp2: ?larf with conj(tau)
#ifdef SLA_SMALL_PROFILE
            if (real_(a[KB]) == 12345.678) pf.t += 1;
#endif
            pf.stamp(4);
            if (c == pvt) {
                live = false;
                mypos = k;
                if (q == kq) a[KB] = mk<T>(beta);
                if (q == 0) { tau_s[k] = tau; scal_s[k] = scal; d_s[k] = fabs(beta); jpvt_s[k] = pvt; }
            }
        }
    }

    // H_k for k = 4 KB + 3 .. 4 KB applied to the caller's slice of a column of Q (k <= kstart only)
    static __device__ __forceinline__ void formq(T (&a)[RPL], int n, int kstart, int q, const T* mirror, const T* tau_s,
                                                 const T* scal_s, const int* jpvt_s) {
#pragma unroll 1
        for (int kq = 3; kq >= 0; --kq) {
            const int k = 4 * KB + kq;
            if (k >= n || k > kstart) continue;
            T x[RPL];
            load_rows<T, RPL, KB>(x, mirror + (size_t)jpvt_s[k] * CLD + q * QLD);
            apply_reflector<T, RPL, KB>(a, x, q, kq, tau_s[k], scal_s[k], true);
        }
    }
};

template <typename T, int NMAX, int KB> struct Small2Run {
    using B = Small2Block<T, NMAX, KB>;
    template <typename... A> static __device__ __forceinline__ void factor(int n, A&&... args) {
        if (4 * KB < n) {
            B::factor(static_cast<A&&>(args)...);
            if constexpr (KB + 1 < B::RPL) Small2Run<T, NMAX, KB + 1>::factor(n, static_cast<A&&>(args)...);
        }
    }
    template <typename... A> static __device__ __forceinline__ void formq(int n, A&&... args) {
        if (4 * KB < n) {
            if constexpr (KB + 1 < B::RPL) Small2Run<T, NMAX, KB + 1>::formq(n, static_cast<A&&>(args)...);
            B::formq(static_cast<A&&>(args)...);
        }
    }
};

template <typename T, int NMAX>
__global__ void __launch_bounds__(NMAX * 4) ldr_small2_kernel(LdrArgs<T> p) {
    using C = SmallCfg<T, NMAX>;
    constexpr int RPL = C::RPL, QLD = C::QLD, CLD = C::CLD;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* mirror = reinterpret_cast<T*>(smem_raw);
    T* tau_s = mirror + (size_t)NMAX * CLD;     // [NMAX] by step
    T* scal_s = tau_s + NMAX;                   // [NMAX] by step
    double* keys = reinterpret_cast<double*>(scal_s + NMAX);   // [2][NMAX]  squared norm of rows >= k, -1 = retired
    double* xn_s = keys + 2 * NMAX;             // [2][NMAX]  squared norm of rows > k
    double* d_s = xn_s + 2 * NMAX;              // [NMAX] |beta| by step
    int* jpvt_s = reinterpret_cast<int*>(d_s + NMAX);          // [NMAX] column retired at step k

    const int n = p.n, b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31;
    const int c = tid >> 2, q = tid & 3;
    const bool col_ok = c < n;
    T* const mycol = mirror + (size_t)c * CLD + q * QLD;

    T a[RPL];
    {
        const T* src = p.Ain ? p.Ain + (long long)b * p.strideAin + (long long)c * p.ldain
                             : p.L + (long long)b * p.strideL + (long long)c * p.ldl;
        double ss = 0.0;
        int nz = 0;
#pragma unroll
        for (int i = 0; i < RPL; ++i) {
            const int r = 4 * i + q;
            a[i] = (col_ok && r < n) ? src[r] : mk<T>(0.0);
            ss += abs2_(a[i]);
            nz |= !is_zero(a[i]);
        }
        ss = group_sum(ss);
        nz |= __shfl_xor_sync(FULL, nz, 1);
        nz |= __shfl_xor_sync(FULL, nz, 2);
        if (q == 0 && col_ok && ldr_norm2_out_of_range(ss, nz != 0)) ldr_raise_range(p.info, b);
    }
    bool live = col_ok;
    int mypos = -1;   // step at which this column was retired

    SmallProf pf;
    pf.start();
    Small2Run<T, NMAX, 0>::factor(n, a, live, mypos, n, c, q, lane, mirror, mycol, tau_s, scal_s, keys, xn_s, d_s, jpvt_s, pf);
    __syncthreads();   // the scalars of the last step
    pf.stamp(5);

    // ---------------------------------------------------------------- d = |diag R'|, R = D^-1 R' P^T
    {
        T* R = p.R + (long long)b * p.strideR + (long long)c * p.ldr;
        if (col_ok) {
#pragma unroll
            for (int i = 0; i < RPL; ++i) {
                const int r = 4 * i + q;
                if (r < n) R[r] = (r <= mypos) ? a[i] / d_s[r] : mk<T>(0.0);
            }
        }
        if (tid < n) {
            p.d[(long long)b * p.stride_d + tid] = d_s[tid];
            if (p.jpvt_out) p.jpvt_out[(long long)b * n + tid] = jpvt_s[tid] + 1;
        }
    }

    pf.stamp(6);
    // ---------------------------------------------------------------- Q = H_0 ... H_{n-1} I
    {
#pragma unroll
        for (int i = 0; i < RPL; ++i) a[i] = mk<T>((4 * i + q == c) ? 1.0 : 0.0);
        // H_k e_c = e_c for k > c: the warp starts at its largest column
        const int kstart = min(n - 1, 8 * (tid >> 5) + 7);
        Small2Run<T, NMAX, 0>::formq(n, a, n, kstart, q, mirror, tau_s, scal_s, jpvt_s);
        T* Q = p.L + (long long)b * p.strideL + (long long)c * p.ldl;
        if (col_ok) {
#pragma unroll
            for (int i = 0; i < RPL; ++i) {
                const int r = 4 * i + q;
                if (r < n) Q[r] = a[i];
            }
        }
    }
#ifdef SLA_SMALL_PROFILE
    pf.stamp(7);
    if (p.prof && b == 0 && tid == 0)
        for (int i = 0; i < 8; ++i) p.prof[i] = pf.acc[i];
#endif
}

// ---------------------------------------------------------------------------------------------
// Version 3 (the default): version 2's blocked column loop with the per-column dependency chain shortened further.  The
// phase counters of version 2 (profiles/ldr_small_r2.txt; a dependent FP64 operation costs ~19 cycles here) put 1460
// cycles on a column: norms 400, barrier 40, pivot search 270, sqrt + reciprocals of ?larfg 500, v^H a and update 260.
//   * every column forms the reflector it WOULD generate (?larfg: one square root, two reciprocals) while the pivot
//     search and the raw dot product with the pivot column run -- three independent instruction streams of the same
//     thread -- and a second barrier hands the winner's scalars over: 270 + 500 cycles become ~460 + a barrier;
//   * the partial norms for the next column are accumulated element by element right behind the update FMAs (four
//     independent chains) instead of in a pass of their own;
//   * pivot keys are compared as integers (non-negative doubles order like their bit patterns);
//   * nothing between the two barriers branches (?larfg from MUFU seeds and Newton steps, see small_larfg), so the three
//     streams really are interleaved by the instruction scheduler.
// ---------------------------------------------------------------------------------------------
// ?larfg without a branch, so that the compiler can interleave it with the pivot search and the dot product (IEEE sqrt and
// division come with slow-path branches that cut the basic block: measured, the three streams then run one after the
// other, 1100 cycles instead of ~450).  MUFU seeds (>= 20 bits) + Newton steps in FMAs: 1 / sqrt(t) by one cubically
// convergent step (error ~2^-58), sqrt(t) = t y with one correction step, 1 / beta = -+y, 1 / (alpha - beta) by two
// quadratic steps.  Results agree with the IEEE sequence to an ulp or two; the operands are within the range the norm
// guard admits (squares representable), where the seeds are valid.
__device__ __forceinline__ double rsqrt_seed(double x) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
__device__ __forceinline__ double rcp_seed(double x) { double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
template <typename T> __device__ __forceinline__ void small_larfg(T alpha, double xn2, T& tau, T& scal, double& beta) {
    const double ar = real_(alpha), ai = imag_(alpha);
    const double t = abs2_(alpha) + xn2;
    const double y0 = rsqrt_seed(t);
    const double h = fma(-t * y0, y0, 1.0);
    const double y = fma(y0 * h, fma(h, 0.375, 0.5), y0);      // 1 / sqrt(t)
    double g = t * y;
    g = fma(fma(-g, g, t), 0.5 * y, g);                        // sqrt(t)
    const double b = -copysign(g, ar);
    const double ib = -copysign(y, ar);                        // 1 / beta
    const double dr = ar - b;                                  // alpha - beta = (dr, ai), |dr| = |ar| + g >= |ai|
    double m = dr;
    if constexpr (is_cplx<T>::value) m = fma(dr, dr, ai * ai);
    double z = rcp_seed(m);
    z = fma(z, fma(-m, z, 1.0), z);
    z = fma(z, fma(-m, z, 1.0), z);                            // 1 / m
    const bool ident = (xn2 == 0.0 && ai == 0.0);              // H = I
    if constexpr (is_cplx<T>::value) {
        tau = ident ? mk<T>(0.0) : mk<T>((b - ar) * ib, -ai * ib);
        scal = ident ? mk<T>(0.0) : mk<T>(dr * z, -ai * z);
    } else {
        tau = ident ? mk<T>(0.0) : mk<T>((b - ar) * ib);
        scal = ident ? mk<T>(0.0) : mk<T>(z);
    }
    beta = ident ? ar : b;
}

template <typename T, int NMAX, int KB> struct Small3Block {
    using C = SmallCfg<T, NMAX>;
    static constexpr int RPL = C::RPL, QLD = C::QLD, CLD = C::CLD;

    // Norms of the caller's column for the NEXT step (row block KB if `same`, else KB + 1; lane kqn of the group holds the
    // diagonal element), its key and its live rows; alpha_o / xn2_o = what ?larfg needs if this column becomes the pivot.
    static __device__ __forceinline__ void publish(const T (&a)[RPL], bool live, int c, int q, int kqn, bool same, int parn,
                                                   T* mycol, double* keys, T& alpha_o, double& xn2_o) {
        double r0 = 0.0, r1 = 0.0, r2 = 0.0, r3 = 0.0;
#pragma unroll
        for (int i = KB + 2; i < RPL; ++i) {
            const int j = (i - KB) & 3;
            if (j == 0) r0 += abs2_(a[i]); else if (j == 1) r1 += abs2_(a[i]); else if (j == 2) r2 += abs2_(a[i]); else r3 += abs2_(a[i]);
        }
        double m1 = 0.0;
        T a1 = mk<T>(0.0);
        if constexpr (KB + 1 < RPL) { a1 = a[KB + 1]; m1 = abs2_(a1); }
        const double mb = abs2_(a[KB]);
        const double rest = (r0 + r1) + (r2 + r3);
        double pl, eloc;
        T aloc;
        if (same) { pl = rest + m1; if (q > kqn) pl += mb; eloc = mb; aloc = a[KB]; }
        else { pl = rest; if (q > kqn) pl += m1; eloc = m1; aloc = a1; }
        xn2_o = group_sum(pl);                           // bit-identical in the four lanes
        const double e = shq(eloc, kqn);
        alpha_o = shq(aloc, kqn);
        if (q == 0) keys[parn + c] = live ? xn2_o + e : -1.0;
        if (live) store_rows<T, RPL, KB>(mycol, a);
    }

    // columns k = 4 KB .. 4 KB + 3 of the pivoted QR
    static __device__ __forceinline__ void factor(T (&a)[RPL], bool& live, int& mypos, T& alpha_o, double& xn2_o, int n, int c,
                                                  int q, int lane, T* mirror, T* mycol, T* tau_s, T* scal_s, T* sc_tau,
                                                  T* sc_scal, double* keys, double* sc_beta, double* d_s, int* jpvt_s,
                                                  SmallProf& pf) {
#pragma unroll 1
        for (int kq = 0; kq < 4; ++kq) {
            const int k = 4 * KB + kq;
            if (k >= n) break;
            const int par = (k & 1) * NMAX;
            __syncthreads();                             // #1: keys and columns for step k
            pf.stamp(1);
            // a. the reflector this column would generate
            T tau_o, scal_o;
            double beta_o;
            small_larfg<T>(alpha_o, xn2_o, tau_o, scal_o, beta_o);   // (stored below, after the loads of the other two streams)
            // b. pivot: largest key, first index among equals (every warp redundantly)
            long long bv = -0x7fffffffffffffffLL;
            int bi = 0;
#pragma unroll
            for (int j = 0; j < C::KPL; ++j) {
                const long long v = __double_as_longlong(keys[par + lane + 32 * j]);
                if (v > bv) { bv = v; bi = lane + 32 * j; }
            }
            const int hi = (int)(bv >> 32);
            const unsigned lo = (unsigned)(bv & 0xffffffffLL);
            const int m1 = __reduce_max_sync(FULL, hi);
            const unsigned m2 = __reduce_max_sync(FULL, hi == m1 ? lo : 0u);
            int pvt = __reduce_min_sync(FULL, (hi == m1 && lo == m2) ? bi : 0x7fffffff);
            pvt = min(pvt, NMAX - 1);
            // c. raw dot product with the pivot column: v^H a = a_k + conj(scal) sum_{r > k} conj(x_r) a_r
            T x[RPL];
            load_rows<T, RPL, KB>(x, mirror + (size_t)pvt * CLD + q * QLD);
            T s0 = mk<T>(0.0), s1 = mk<T>(0.0);
#pragma unroll
            for (int i = KB + 1; i < RPL; ++i) {
                if ((i - KB) & 1) fmacc(s0, conj_(x[i]), a[i]); else fmacc(s1, conj_(x[i]), a[i]);
            }
            T s = s0 + s1;
            if (q > kq) fmacc(s, conj_(x[KB]), a[KB]);
            s = group_sum(s);
            const T ak = shq(a[KB], kq);
            // (a never-taken store that consumes s: keeps the dot product on this side of the barrier)
            if (__double_as_longlong(real_(s)) == 0x7ff8dead0000beefLL) keys[par + c] = 0.0;
            // Stored by all four lanes of the group (same values: a store by lane 0 alone makes the compiler wrap the whole
            // of ?larfg into a divergent branch), and stored LAST: shared-memory loads are not moved above a shared-memory
            // store, so an early store would serialise the streams.
            sc_tau[c] = tau_o; sc_scal[c] = scal_o; sc_beta[c] = beta_o;
#ifdef SLA_SMALL_PROFILE
            if (real_(s) == 12345.678 || sc_beta[c] == 12345.678) pf.t += 1;
#endif
            pf.stamp(2);
            __syncthreads();                             // #2: every column's candidate reflector is published
            pf.stamp(3);
            const T tau = sc_tau[pvt], scal = sc_scal[pvt];
            const double beta = sc_beta[pvt];
            const T w = conj_(tau) * (ak + conj_(scal) * s);   // ?laqp2: ?larf with conj(tau)
            if (c == pvt) {
                live = false;
                mypos = k;
                if (q == kq) a[KB] = mk<T>(beta);
                if (q == 0) { tau_s[k] = tau; scal_s[k] = scal; d_s[k] = fabs(beta); jpvt_s[k] = pvt; }
            } else if (live) {
                const T ws = scal * w;
#pragma unroll
                for (int i = KB + 1; i < RPL; ++i) fms(a[i], x[i], ws);
                if (q > kq) fms(a[KB], x[KB], ws);
                else if (q == kq) a[KB] = a[KB] - w;
            }
#ifdef SLA_SMALL_PROFILE
            if (real_(a[KB]) == 12345.678) pf.t += 1;
#endif
            pf.stamp(4);
            if (k + 1 < n) publish(a, live, c, q, kq < 3 ? kq + 1 : 0, kq < 3, ((k + 1) & 1) * NMAX, mycol, keys, alpha_o, xn2_o);
#ifdef SLA_SMALL_PROFILE
            if (xn2_o == 12345.678) pf.t += 1;
#endif
            pf.stamp(0);
        }
    }
};

template <typename T, int NMAX, int KB> struct Small3Run {
    using B = Small3Block<T, NMAX, KB>;
    template <typename... A> static __device__ __forceinline__ void factor(int n, A&&... args) {
        if (4 * KB < n) {
            B::factor(static_cast<A&&>(args)...);
            if constexpr (KB + 1 < B::RPL) Small3Run<T, NMAX, KB + 1>::factor(n, static_cast<A&&>(args)...);
        }
    }
};

template <typename T, int NMAX>
__global__ void __launch_bounds__(NMAX * 4) ldr_small3_kernel(LdrArgs<T> p) {
    using C = SmallCfg<T, NMAX>;
    constexpr int RPL = C::RPL, QLD = C::QLD, CLD = C::CLD;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* mirror = reinterpret_cast<T*>(smem_raw);
    T* tau_s = mirror + (size_t)NMAX * CLD;     // [NMAX] by step
    T* scal_s = tau_s + NMAX;                   // [NMAX] by step
    T* sc_tau = scal_s + NMAX;                  // [NMAX] by column: the reflector the column would generate at this step
    T* sc_scal = sc_tau + NMAX;                 // [NMAX]
    double* keys = reinterpret_cast<double*>(sc_scal + NMAX);   // [2][NMAX]  squared norm of rows >= k, -1 = retired
    double* sc_beta = keys + 2 * NMAX;          // [NMAX]
    double* d_s = sc_beta + NMAX;               // [NMAX] |beta| by step
    int* jpvt_s = reinterpret_cast<int*>(d_s + NMAX);           // [NMAX] column retired at step k

    const int n = p.n, b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31;
    const int c = tid >> 2, q = tid & 3;
    const bool col_ok = c < n;
    T* const mycol = mirror + (size_t)c * CLD + q * QLD;

    T a[RPL];
    {
        const T* src = p.Ain ? p.Ain + (long long)b * p.strideAin + (long long)c * p.ldain
                             : p.L + (long long)b * p.strideL + (long long)c * p.ldl;
        double ss = 0.0;
        int nz = 0;
#pragma unroll
        for (int i = 0; i < RPL; ++i) {
            const int r = 4 * i + q;
            a[i] = (col_ok && r < n) ? src[r] : mk<T>(0.0);
            ss += abs2_(a[i]);
            nz |= !is_zero(a[i]);
        }
        ss = group_sum(ss);
        nz |= __shfl_xor_sync(FULL, nz, 1);
        nz |= __shfl_xor_sync(FULL, nz, 2);
        if (q == 0 && col_ok && ldr_norm2_out_of_range(ss, nz != 0)) ldr_raise_range(p.info, b);
    }
    bool live = col_ok;
    int mypos = -1;   // step at which this column was retired
    T alpha_o;
    double xn2_o;
    SmallProf pf;
    pf.start();
    Small3Block<T, NMAX, 0>::publish(a, live, c, q, 0, true, 0, mycol, keys, alpha_o, xn2_o);
    Small3Run<T, NMAX, 0>::factor(n, a, live, mypos, alpha_o, xn2_o, n, c, q, lane, mirror, mycol, tau_s, scal_s, sc_tau, sc_scal,
                                  keys, sc_beta, d_s, jpvt_s, pf);
    __syncthreads();   // the scalars of the last step
    pf.stamp(5);

    // ---------------------------------------------------------------- d = |diag R'|, R = D^-1 R' P^T
    {
        T* R = p.R + (long long)b * p.strideR + (long long)c * p.ldr;
        if (col_ok) {
#pragma unroll
            for (int i = 0; i < RPL; ++i) {
                const int r = 4 * i + q;
                if (r < n) R[r] = (r <= mypos) ? a[i] / d_s[r] : mk<T>(0.0);
            }
        }
        if (tid < n) {
            p.d[(long long)b * p.stride_d + tid] = d_s[tid];
            if (p.jpvt_out) p.jpvt_out[(long long)b * n + tid] = jpvt_s[tid] + 1;
        }
    }

    pf.stamp(6);
    // ---------------------------------------------------------------- Q = H_0 ... H_{n-1} I
    {
#pragma unroll
        for (int i = 0; i < RPL; ++i) a[i] = mk<T>((4 * i + q == c) ? 1.0 : 0.0);
        // H_k e_c = e_c for k > c: the warp starts at its largest column
        const int kstart = min(n - 1, 8 * (tid >> 5) + 7);
        Small2Run<T, NMAX, 0>::formq(n, a, n, kstart, q, mirror, tau_s, scal_s, jpvt_s);
        T* Q = p.L + (long long)b * p.strideL + (long long)c * p.ldl;
        if (col_ok) {
#pragma unroll
            for (int i = 0; i < RPL; ++i) {
                const int r = 4 * i + q;
                if (r < n) Q[r] = a[i];
            }
        }
    }
#ifdef SLA_SMALL_PROFILE
    pf.stamp(7);
    if (p.prof && b == 0 && tid == 0)
        for (int i = 0; i < 8; ++i) p.prof[i] = pf.acc[i];
#endif
}

template <typename T, int NMAX> cudaError_t launch_small(const LdrArgs<T>& p, cudaStream_t stream) {
    using C = SmallCfg<T, NMAX>;
    auto kern = ldr_small_kernel<T, NMAX>;
    if (C::SMEM > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return e;
    }
    kern<<<p.batch, C::THREADS, C::SMEM, stream>>>(p);
    ++g_launch_count;
    return cudaGetLastError();
}
template <typename T, int NMAX, int VERSION> cudaError_t launch_small2(const LdrArgs<T>& p, cudaStream_t stream) {
    using C = SmallCfg<T, NMAX>;
    auto kern = VERSION == 3 ? ldr_small3_kernel<T, NMAX> : ldr_small2_kernel<T, NMAX>;
    if (C::SMEM > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return e;
    }
    kern<<<p.batch, C::THREADS, C::SMEM, stream>>>(p);
    ++g_launch_count;
    return cudaGetLastError();
}

}  // namespace

template <typename T> bool ldr_small_supported(int n) { return n >= 1 && n <= 64; }

template <typename T> cudaError_t ldr_small_batched(const LdrArgs<T>& p, cudaStream_t stream) {
    if (p.batch <= 0 || p.n <= 0) return cudaSuccess;
    if (!ldr_small_supported<T>(p.n)) return cudaErrorInvalidValue;
    // version 3 unless SLA_LDR_SMALL=1 / 2 asks for an earlier one; measured on one B200 (profiles/ldr_small_r2.txt), ldr! of one
    // 64 x 64 matrix / of 2048 of them / ComplexF64 / the C1 step:  v1 100.4 us, 1363 us, 166 us, 1.502 ms;  v2 67.4, 556, 121,
    // 1.189;  v3 61.3, 530, 96.1, 1.106
    static const int version = [] { const char* e = getenv("SLA_LDR_SMALL"); const int v = e ? atoi(e) : 3; return (v == 1 || v == 2) ? v : 3; }();
    if (version == 3) {
        if (p.n <= 32) return launch_small2<T, 32, 3>(p, stream);
        return launch_small2<T, 64, 3>(p, stream);
    }
    if (version != 1) {
        if (p.n <= 32) return launch_small2<T, 32, 2>(p, stream);
        return launch_small2<T, 64, 2>(p, stream);
    }
    if (p.n <= 32) return launch_small<T, 32>(p, stream);
    return launch_small<T, 64>(p, stream);
}

template bool ldr_small_supported<double>(int);
template bool ldr_small_supported<cplx>(int);
template cudaError_t ldr_small_batched<double>(const LdrArgs<double>&, cudaStream_t);
template cudaError_t ldr_small_batched<cplx>(const LdrArgs<cplx>&, cudaStream_t);

}  // namespace sla
