// Shared device helpers for the sm_100a kernels of libsla_b200.
//
// FP64 tensor-core math on sm_100a is the warp-level DMMA.8x8x4 (every mma.sync f64 shape lowers
// to it; tcgen05/UMMA has no f64 kind), so the building block here is `dmma884` and everything
// else (cp.async staging, warp-shuffle reductions, the tiny complex type) is written around it.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace sla {

// ------------------------------------------------------------------------------------------
// scalar types: double and an interleaved (re,im) complex that is layout-compatible with
// Julia's ComplexF64 / numpy complex128.
// ------------------------------------------------------------------------------------------
struct __align__(16) cplx {
    double re, im;
};

template <typename T> struct is_cplx { static constexpr bool value = false; };
template <> struct is_cplx<cplx> { static constexpr bool value = true; };

__host__ __device__ __forceinline__ double make_T(double re, double, double*) { return re; }
__host__ __device__ __forceinline__ cplx make_T(double re, double im, cplx*) { return cplx{re, im}; }
template <typename T> __host__ __device__ __forceinline__ T mk(double re, double im = 0.0) {
    return make_T(re, im, (T*)nullptr);
}

__host__ __device__ __forceinline__ double conj_(double x) { return x; }
__host__ __device__ __forceinline__ cplx conj_(cplx x) { return cplx{x.re, -x.im}; }
__host__ __device__ __forceinline__ double real_(double x) { return x; }
__host__ __device__ __forceinline__ double real_(cplx x) { return x.re; }
__host__ __device__ __forceinline__ double imag_(double) { return 0.0; }
__host__ __device__ __forceinline__ double imag_(cplx x) { return x.im; }
__host__ __device__ __forceinline__ double abs2_(double x) { return x * x; }
__host__ __device__ __forceinline__ double abs2_(cplx x) { return x.re * x.re + x.im * x.im; }
__device__ __forceinline__ double abs_(double x) { return fabs(x); }
__device__ __forceinline__ double abs_(cplx x) { return hypot(x.re, x.im); }
// |re|+|im| -- what izamax (LAPACK pivot search for complex) compares; |x| for real
__device__ __forceinline__ double abs1_(double x) { return fabs(x); }
__device__ __forceinline__ double abs1_(cplx x) { return fabs(x.re) + fabs(x.im); }

__host__ __device__ __forceinline__ cplx operator+(cplx a, cplx b) { return cplx{a.re + b.re, a.im + b.im}; }
__host__ __device__ __forceinline__ cplx operator-(cplx a, cplx b) { return cplx{a.re - b.re, a.im - b.im}; }
__host__ __device__ __forceinline__ cplx operator-(cplx a) { return cplx{-a.re, -a.im}; }
__host__ __device__ __forceinline__ cplx operator*(cplx a, cplx b) {
    return cplx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
__host__ __device__ __forceinline__ cplx operator*(cplx a, double b) { return cplx{a.re * b, a.im * b}; }
__host__ __device__ __forceinline__ cplx operator*(double a, cplx b) { return cplx{a * b.re, a * b.im}; }
__host__ __device__ __forceinline__ cplx operator/(cplx a, double b) { return cplx{a.re / b, a.im / b}; }
__device__ __forceinline__ cplx operator/(cplx a, cplx b) {
    // Smith's algorithm (what Julia/LAPACK zladiv-class division amounts to for our ranges)
    if (fabs(b.re) >= fabs(b.im)) {
        double r = b.im / b.re, den = b.re + b.im * r;
        return cplx{(a.re + a.im * r) / den, (a.im - a.re * r) / den};
    }
    double r = b.re / b.im, den = b.re * r + b.im;
    return cplx{(a.re * r + a.im) / den, (a.im * r - a.re) / den};
}
__host__ __device__ __forceinline__ cplx& operator+=(cplx& a, cplx b) { a.re += b.re; a.im += b.im; return a; }
__host__ __device__ __forceinline__ cplx& operator-=(cplx& a, cplx b) { a.re -= b.re; a.im -= b.im; return a; }
__host__ __device__ __forceinline__ bool is_zero(double x) { return x == 0.0; }
__host__ __device__ __forceinline__ bool is_zero(cplx x) { return x.re == 0.0 && x.im == 0.0; }

// fused multiply-subtract acc -= a*b (complex: 4 FMAs)
__device__ __forceinline__ void fms(double& acc, double a, double b) { acc = fma(-a, b, acc); }
__device__ __forceinline__ void fms(cplx& acc, cplx a, cplx b) {
    acc.re = fma(-a.re, b.re, acc.re); acc.re = fma(a.im, b.im, acc.re);
    acc.im = fma(-a.re, b.im, acc.im); acc.im = fma(-a.im, b.re, acc.im);
}
__device__ __forceinline__ void fmacc(double& acc, double a, double b) { acc = fma(a, b, acc); }
__device__ __forceinline__ void fmacc(cplx& acc, cplx a, cplx b) {
    acc.re = fma(a.re, b.re, acc.re); acc.re = fma(-a.im, b.im, acc.re);
    acc.im = fma(a.re, b.im, acc.im); acc.im = fma(a.im, b.re, acc.im);
}

// sign(z) as Julia defines it: z/|z| (0 -> 0)
__device__ __forceinline__ double sign_(double x) { return (x > 0.0) ? 1.0 : ((x < 0.0) ? -1.0 : 0.0); }
__device__ __forceinline__ cplx sign_(cplx x) {
    double a = hypot(x.re, x.im);
    return (a == 0.0) ? cplx{0.0, 0.0} : cplx{x.re / a, x.im / a};
}

// ------------------------------------------------------------------------------------------
// diagonal-scaling modes shared by the GEMM epilogue and the element-wise kernels.  They are the
// D, D^-1, D_min = min(D,1) and D_max^-1 = 1/max(D,1) factors of exported_functions.jl:15-21.
// ------------------------------------------------------------------------------------------
enum ScaleMode : int {
    SCALE_NONE = 0,
    SCALE_MUL = 1,       // x * d
    SCALE_DIV = 2,       // x / d
    SCALE_MUL_MIN1 = 3,  // x * min(d,1)
    SCALE_DIV_MAX1 = 4,  // x / max(d,1)
    SCALE_MUL_MAX1 = 5,  // x * max(d,1)
    SCALE_DIV_MIN1 = 6,  // x / min(d,1)
};

template <typename T> __device__ __forceinline__ T apply_scale(T x, double d, int mode) {
    switch (mode) {
        case SCALE_MUL: return x * d;
        case SCALE_DIV: return x / d;
        case SCALE_MUL_MIN1: return x * fmin(d, 1.0);
        case SCALE_DIV_MAX1: return x / fmax(d, 1.0);
        case SCALE_MUL_MAX1: return x * fmax(d, 1.0);
        case SCALE_DIV_MIN1: return x / fmin(d, 1.0);
        default: return x;
    }
}

// ------------------------------------------------------------------------------------------
// DMMA.8x8x4:  D(8x8) += A(8x4) * B(4x8).  Lane t holds A[t/4][t%4], B[t%4][t/4],
// D[t/4][2*(t%4)+{0,1}].
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
        : "+d"(d0), "+d"(d1)
        : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------
// cp.async (LDGSTS) helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
// 16-byte copy with zero-fill of the bytes beyond src_bytes (0, 8 or 16)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
                 "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
                 "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

// ------------------------------------------------------------------------------------------
// warp reductions
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ cplx warp_sum(cplx v) {
    v.re = warp_sum(v.re);
    v.im = warp_sum(v.im);
    return v;
}
// argmax with "first maximal index wins" (idamax semantics); NaNs lose
__device__ __forceinline__ void warp_argmax(double& v, int& i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, v, o);
        int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
}

}  // namespace sla
