// inv_lu! for SMALL matrices (n <= 64): one CTA per matrix, the matrix in registers, ONE barrier per column.
//
// Replaces `inv_lu!(A, ws)` (src/lu.jl:149-158: det_lu! src/lu.jl:123-140 = getrf! + diagonal/pivot scan, then getri!;
// LAPACK behind src/lu.jl:63-93) where the general kernel (sla_lu.cu: panels through shared memory, interchanges through
// global memory, two blocked triangular sweeps, 512 threads and two barriers per column) is all latency: measured in the
// step of BASELINE config C1 (N = 64, one chain; profiles/launches_r2v_C1_step_summary.txt) the two LU inverses of
// `inv_IpA!` took ~0.2 ms each, as much as the ten factorisations of the chain together.
//
// Algorithm: Gauss-Jordan elimination with partial pivoting, in place, WITHOUT moving rows ("implicit" pivoting: step k
// picks the largest |a_rk| among the rows that have not been a pivot yet -- the same rows, hence the same pivots u_kk and
// the same multipliers below the pivot as ?getrf -- and remembers p_k).  Column k of the working array is replaced by the
// column the inverse gains in that step, so after n steps  slot c holds column p_c of E = Pi A^-1  (Pi[p_k, k] = 1), i.e.
//     A^-1[k, p_c] = slot_c[p_k].
// Mathematically the inverse ?getri returns; rounding differs in the order of the eliminations above the diagonal.
// log|det| = sum log|u_kk|, sign = prod sign(u_kk) * parity of k -> p_k (inversion count), negated / conjugated for the
// inverse like src/lu.jl:157.  Ties in |a_rk| go to the smallest row index (LAPACK: smallest current position).
//
// Shape: thread (c, q) = (tid / 4, tid % 4) holds rows q, q+4, ... of slot c (NMAX / 4 values).  Per step k:
//   A. the warp that owns slot k publishes column k as it stands, finds the pivot over it with all 32 lanes (integer compares
//      on |a| bit patterns, three `redux` rounds) and publishes p_k and 1 / u_kk -- nothing else: seven of the eight warps wait
//      for this one;
//   -- barrier --
//   B. every other slot:  a_pc = its element in row p_k (one shuffle inside the group),  t = a_pc / u_kk,
//      a_r -= a_rk t  for r != p_k,  a_pk = t;  slot k turns into the column the inverse gains: -a_rk / u_kk, 1 / u_kk in row p_k.
#include <stdlib.h>

#include "sla_kernels.h"

namespace sla {

namespace {

constexpr unsigned LS_FULL = 0xffffffffu;

template <typename T> struct LuSmallPad { static constexpr int value = 2; };      // 16 bytes between lane slices
template <> struct LuSmallPad<cplx> { static constexpr int value = 1; };

__device__ __forceinline__ double ls_shq(double v, int src) { return __shfl_sync(LS_FULL, v, src, 4); }
__device__ __forceinline__ cplx ls_shq(cplx v, int src) { return cplx{ls_shq(v.re, src), ls_shq(v.im, src)}; }
__device__ __forceinline__ double ls_shx(double v, int m) { return __shfl_xor_sync(LS_FULL, v, m); }
__device__ __forceinline__ cplx ls_shx(cplx v, int m) { return cplx{ls_shx(v.re, m), ls_shx(v.im, m)}; }
__device__ __forceinline__ long long ls_shx(long long v, int m) { return __shfl_xor_sync(LS_FULL, v, m); }

// a[idx] for a warp-uniform idx without predicated moves over the whole array: log2(RPL) uniform branches
template <typename T, int RPL, int LO, int HI> __device__ __forceinline__ T ls_get(const T (&a)[RPL], int idx) {
    if constexpr (HI - LO == 1) {
        return a[LO];
    } else {
        constexpr int MID = (LO + HI) / 2;
        return (idx < MID) ? ls_get<T, RPL, LO, MID>(a, idx) : ls_get<T, RPL, MID, HI>(a, idx);
    }
}
template <typename T, int RPL, int LO, int HI> __device__ __forceinline__ void ls_set(T (&a)[RPL], int idx, T v) {
    if constexpr (HI - LO == 1) {
        a[LO] = v;
    } else {
        constexpr int MID = (LO + HI) / 2;
        if (idx < MID) ls_set<T, RPL, LO, MID>(a, idx, v); else ls_set<T, RPL, MID, HI>(a, idx, v);
    }
}

template <typename T, int RPL> __device__ __forceinline__ void ls_load(T (&x)[RPL], const T* __restrict__ src) {
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
template <typename T, int RPL> __device__ __forceinline__ void ls_store(T* __restrict__ dst, const T (&x)[RPL]) {
    if constexpr (is_cplx<T>::value) {
#pragma unroll
        for (int i = 0; i < RPL; ++i) dst[i] = x[i];
    } else {
#pragma unroll
        for (int i = 0; i < RPL; i += 2) *reinterpret_cast<double2*>(dst + i) = make_double2(x[i], x[i + 1]);
    }
}
// 1 / u_kk: MUFU seed + two Newton steps (an ulp or two from the IEEE quotient, without its slow-path call on the chain)
// inside the range where the seed is valid; IEEE division for zero, tiny, huge and non-finite pivots
__device__ __forceinline__ double ls_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = fma(y, fma(-x, y, 1.0), y);
    y = fma(y, fma(-x, y, 1.0), y);
    const double ax = fabs(x);
    if (!(ax > 1e-290 && ax < 1e290)) y = 1.0 / x;
    return y;
}
__device__ __forceinline__ cplx ls_rcp(cplx x) { return cplx{1.0, 0.0} / x; }

// |a| as ?getf2 / i?amax compare it, as an integer key: the bit pattern of a non-negative double is monotone; NaN -> -1
__device__ __forceinline__ long long ls_key(double x) {
    const double v = fabs(x);
    return (v == v) ? __double_as_longlong(v) : -1LL;
}
__device__ __forceinline__ long long ls_key(cplx x) {
    const double v = fabs(x.re) + fabs(x.im);
    return (v == v) ? __double_as_longlong(v) : -1LL;
}
__device__ __forceinline__ bool ls_finite(double x) { return isfinite(x); }
__device__ __forceinline__ bool ls_finite(cplx x) { return isfinite(x.re) && isfinite(x.im); }

template <typename T, int NMAX> struct LuSmallCfg {
    static constexpr int RPL = NMAX / 4;
    static constexpr int THREADS = NMAX * 4;
    static constexpr int QLD = RPL + LuSmallPad<T>::value;     // elements between the multiplier slices of two lanes
};

template <typename T, int NMAX>
__global__ void __launch_bounds__(NMAX * 4) lu_small_inv_kernel(LuArgs<T> p) {
    using C = LuSmallCfg<T, NMAX>;
    constexpr int RPL = C::RPL, QLD = C::QLD;
    __shared__ __align__(16) T col_s[2][4 * QLD];    // column k at step k, [lane of the group][row / 4]
    __shared__ T rcp_s[2];                           // 1 / u_kk
    __shared__ int piv_s[2];                         // p_k
    __shared__ T pv_s[NMAX];                         // u_kk by step
    __shared__ int prow_s[NMAX];                     // p_k by step
    __shared__ int kofr_s[NMAX];                     // step at which row r was the pivot row
    __shared__ double red_s[NMAX / 8];
    __shared__ T redt_s[NMAX / 8];

    const int n = p.n, b = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int c = tid >> 2, q = tid & 3;
    const bool slot_ok = c < n;
    T* A = p.A + (long long)b * p.strideA;

    T a[RPL];
    {
        const T* src = A + (long long)c * p.lda;
#pragma unroll
        for (int i = 0; i < RPL; ++i) {
            const int r = 4 * i + q;
            a[i] = (slot_ok && r < n) ? src[r] : mk<T>(0.0);
        }
    }
    unsigned used2 = 0;                              // bit h: row lane + 32 h has been a pivot row (rows >= n count as used)
#pragma unroll
    for (int h = 0; h < NMAX / 32; ++h)
        if ((tid & 31) + 32 * h >= n) used2 |= 1u << h;
    int bad = 0;                                     // first step with a zero / non-finite pivot, 1-based (group of slot k)

#pragma unroll 1
    for (int k = 0; k < n; ++k) {
        const int par = k & 1;
        // ---- A. the warp of slot k: column k, pivot, reciprocal
        if ((k >> 3) == warp) {
            if (c == k) ls_store<T, RPL>(&col_s[par][q * QLD], a);   // column k as it stands: everybody's multipliers come from it
            // The search runs over the published column with all 32 lanes (rows lane and lane + 32), not inside the four lanes
            // that hold it: the seven other warps wait for this one (ncu: 62 % of the kernel's stall samples were their
            // barrier waits when the group searched its 16 registers per lane serially).
            __syncwarp();
            const int lane = tid & 31;
            long long bk = -3;
            int br = 0;
#pragma unroll
            for (int h = 0; h < NMAX / 32; ++h) {
                const int r = lane + 32 * h;
                const long long key = ((used2 >> h) & 1u) ? -2LL : ls_key(col_s[par][(r & 3) * QLD + (r >> 2)]);
                if (key > bk) { bk = key; br = r; }
            }
            const int hi = (int)(bk >> 32);
            const unsigned lo = (unsigned)(bk & 0xffffffffLL);
            const int m1 = __reduce_max_sync(LS_FULL, hi);
            const unsigned m2 = __reduce_max_sync(LS_FULL, hi == m1 ? lo : 0u);
            int pr = __reduce_min_sync(LS_FULL, (hi == m1 && lo == m2) ? br : 0x7fffffff);   // first maximal row
            pr = min(pr, NMAX - 1);                  // (an unused row always exists: keys of unused rows are >= -1)
            const T pv = col_s[par][(pr & 3) * QLD + (pr >> 2)];
            const T rc = ls_rcp(pv);
            if (c == k) {
                const bool singular = is_zero(pv) || !ls_finite(pv);
                if (singular && bad == 0) bad = k + 1;
                if (q == 0) { rcp_s[par] = rc; piv_s[par] = pr; pv_s[k] = pv; prow_s[k] = pr; kofr_s[pr] = k; }
            }
        }
        __syncthreads();
        // ---- B. every slot
        const int pr = piv_s[par];
        const T rc = rcp_s[par];
        const int ip = pr >> 2, qp = pr & 3;
        const T apc = ls_shq(ls_get<T, RPL, 0, RPL>(a, ip), qp);
        if (slot_ok) {
            if (c != k) {
                const T t = apc * rc;                // a_pc / u_kk: the new element in row p_k
                T col[RPL];
                ls_load<T, RPL>(col, &col_s[par][q * QLD]);
#pragma unroll
                for (int i = 0; i < RPL; ++i) fms(a[i], col[i], t);
                if (q == qp) ls_set<T, RPL, 0, RPL>(a, ip, t);
            } else {                                 // the column the inverse gains: -a_rk / u_kk, 1 / u_kk in row p_k
                const T nrc = -rc;
#pragma unroll
                for (int i = 0; i < RPL; ++i) a[i] = a[i] * nrc;
                if (q == qp) ls_set<T, RPL, 0, RPL>(a, ip, rc);
            }
        }
        if ((pr & 31) == (tid & 31)) used2 |= 1u << (pr >> 5);
    }
    __syncthreads();

    // ---- A^-1[k_r, p_c] = slot_c[r]
    if (slot_ok) {
        const int pc = prow_s[c];
        T* X = p.X ? p.X + (long long)b * p.strideX + (long long)pc * p.ldx : nullptr;
        T* Ab = p.copy_back ? A + (long long)pc * p.lda : nullptr;
#pragma unroll
        for (int i = 0; i < RPL; ++i) {
            const int r = 4 * i + q;
            if (r < n) {
                const int kr = kofr_s[r];
                if (X) X[kr] = a[i];
                if (Ab) Ab[kr] = a[i];
            }
        }
    }
    // ---- info, log|det|, sign  (of the INVERSE: src/lu.jl:157)
    if (bad != 0 && q == 0 && p.info) atomicCAS(p.info + b, 0, bad);
    if (p.logdet || p.sign) {
        const bool mine = tid < n;
        const T u = mine ? pv_s[tid] : mk<T>(1.0);
        double lg = mine ? log(abs_(u)) : 0.0;
        T sg = mine ? sign_(u) : mk<T>(1.0);
        int inv = 0;                                 // inversions (j > k, p_j < p_k) of this thread's k
        if (mine) {
            const int pk = prow_s[tid];
            for (int j = tid + 1; j < n; ++j) inv += (prow_s[j] < pk);
        }
        const int odd = __syncthreads_count(inv & 1);
        lg = warp_sum(lg);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sg = sg * ls_shx(sg, o);
        if ((tid & 31) == 0 && warp < NMAX / 8) { red_s[warp] = lg; redt_s[warp] = sg; }
        __syncthreads();
        if (tid == 0) {
            double L = 0.0;
            T S = mk<T>((odd & 1) ? -1.0 : 1.0);
            for (int w = 0; w < (n + 31) / 32; ++w) { L += red_s[w]; S = S * redt_s[w]; }
            if (p.logdet) p.logdet[b] = -L;
            if (p.sign) p.sign[b] = conj_(S);
        }
    }
}

template <typename T, int NMAX> cudaError_t launch_lu_small(const LuArgs<T>& p, cudaStream_t stream) {
    lu_small_inv_kernel<T, NMAX><<<p.batch, NMAX * 4, 0, stream>>>(p);
    ++g_launch_count;
    return cudaGetLastError();
}

}  // namespace

// the inverse form without a right-hand side and without exported pivots (what inv_lu! and the inverses inside inv_IpA! /
// inv_IpUV! / inv_UpV! / inv_invUpV! ask for); SLA_LU_SMALL=0 switches the kernel off
template <typename T> bool lu_small_applies(const LuArgs<T>& p) {
    static const bool on = !(getenv("SLA_LU_SMALL") && atoi(getenv("SLA_LU_SMALL")) == 0);
    return on && p.invert && p.rhs == nullptr && p.ipiv_out == nullptr && p.n >= 1 && p.n <= 64 && (p.X != nullptr || p.copy_back);
}

template <typename T> cudaError_t lu_small_inverse_batched(const LuArgs<T>& p, cudaStream_t stream) {
    if (p.batch <= 0 || p.n <= 0) return cudaSuccess;
    if (p.n <= 32) return launch_lu_small<T, 32>(p, stream);
    return launch_lu_small<T, 64>(p, stream);
}

template bool lu_small_applies<double>(const LuArgs<double>&);
template bool lu_small_applies<cplx>(const LuArgs<cplx>&);
template cudaError_t lu_small_inverse_batched<double>(const LuArgs<double>&, cudaStream_t);
template cudaError_t lu_small_inverse_batched<cplx>(const LuArgs<cplx>&, cudaStream_t);

}  // namespace sla
