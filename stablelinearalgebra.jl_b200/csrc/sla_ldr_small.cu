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

}  // namespace

template <typename T> bool ldr_small_supported(int n) { return n >= 1 && n <= 64; }

template <typename T> cudaError_t ldr_small_batched(const LdrArgs<T>& p, cudaStream_t stream) {
    if (p.batch <= 0 || p.n <= 0) return cudaSuccess;
    if (!ldr_small_supported<T>(p.n)) return cudaErrorInvalidValue;
    if (p.n <= 32) return launch_small<T, 32>(p, stream);
    return launch_small<T, 64>(p, stream);
}

template bool ldr_small_supported<double>(int);
template bool ldr_small_supported<cplx>(int);
template cudaError_t ldr_small_batched<double>(const LdrArgs<double>&, cudaStream_t);
template cudaError_t ldr_small_batched<cplx>(const LdrArgs<cplx>&, cudaStream_t);

}  // namespace sla
