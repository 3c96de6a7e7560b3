// Element-wise and small reduction kernels: the Diagonal scalings, adjoints, identity fills and
// scalar bookkeeping that sit between the GEMM / LDR / LU kernels.  All are HBM/L2-streaming kernels:
// thread-per-row so that every warp access is a contiguous run of one column (column-major data).
#include "sla_kernels.h"

namespace sla {

std::atomic<unsigned long long> g_launch_count{0};

namespace {

template <typename T>
__global__ void scale_kernel(T* dst, long long sdst, int lddst, const T* src, long long ssrc, int ldsrc, int op,
                             const double* rs, long long srs, int rs_mode, const double* cs, long long scs,
                             int cs_mode, int n) {
    const int b = blockIdx.z;
    T* D = dst + (long long)b * sdst;
    const T* S = src + (long long)b * ssrc;
    const double* r = rs ? rs + (long long)b * srs : nullptr;
    const double* c = cs ? cs + (long long)b * scs : nullptr;
    if (op == OP_N) {
        const int row = blockIdx.x * blockDim.x + threadIdx.x;
        if (row >= n) return;
        const double rv = r ? r[row] : 1.0;
        for (int col = blockIdx.y; col < n; col += gridDim.y) {
            T v = S[(long long)col * ldsrc + row];
            if (r) v = apply_scale(v, rv, rs_mode);
            if (c) v = apply_scale(v, c[col], cs_mode);
            D[(long long)col * lddst + row] = v;
        }
    } else {
        // adjoint through a shared 32x32 tile (coalesced on both sides)
        __shared__ T tile[32][33];
        const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // blockDim.x = 256 -> ty in 0..7
        const int tiles = (n + 31) / 32;
        for (int t = blockIdx.x + blockIdx.y * gridDim.x; t < tiles * tiles; t += gridDim.x * gridDim.y) {
            const int tr = (t % tiles) * 32, tc = (t / tiles) * 32;  // dst tile origin (row, col)
            // dst[row][col] = conj(src[col][row]) : read src rows tc.., cols tr..
            for (int j = ty; j < 32; j += 8) {
                int srow = tc + tx, scol = tr + j;
                if (srow < n && scol < n) tile[j][tx] = S[(long long)scol * ldsrc + srow];
            }
            __syncthreads();
            for (int j = ty; j < 32; j += 8) {
                int drow = tr + tx, dcol = tc + j;
                if (drow < n && dcol < n) {
                    T v = conj_(tile[tx][j]);
                    if (r) v = apply_scale(v, r[drow], rs_mode);
                    if (c) v = apply_scale(v, c[dcol], cs_mode);
                    D[(long long)dcol * lddst + drow] = v;
                }
            }
            __syncthreads();
        }
    }
}

template <typename T> __global__ void identity_kernel(T* A, long long sA, int lda, int n) {
    T* D = A + (long long)blockIdx.z * sA;
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    for (int col = blockIdx.y; col < n; col += gridDim.y) D[(long long)col * lda + row] = mk<T>(row == col ? 1.0 : 0.0);
}

__global__ void fill_kernel(double* d, double v, long long count) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) d[i] = v;
}

template <typename T>
__global__ void ipa_combine_kernel(T* M, T* Rinv, const T* L, const double* d, int n) {
    const int b = blockIdx.z;
    const long long off = (long long)b * n * n;
    const double* db = d + (long long)b * n;
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    for (int col = blockIdx.y; col < n; col += gridDim.y) {
        const double dv = db[col];
        const long long e = off + (long long)col * n + row;
        T lm = L[e] * fmin(dv, 1.0);          // exported_functions.jl:37-42
        T ri = Rinv[e] / fmax(dv, 1.0);       // :53-54
        Rinv[e] = ri;
        M[e] = lm + ri;                       // :57
    }
}

__global__ void logsum_kernel(double* out, const double* d, long long sd, int mode, int n) {
    const double* db = d + (long long)blockIdx.x * sd;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double v = db[i];
        if (mode == SCALE_MUL_MAX1) v = fmax(v, 1.0);
        else if (mode == SCALE_MUL_MIN1) v = fmin(v, 1.0);
        acc += log(v);
    }
    __shared__ double red[32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (blockDim.x + 31) / 32; ++w) s += red[w];
        out[blockIdx.x] = s;
    }
}

template <typename T> __global__ void combine_kernel(CombineArgs<T> a) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.batch) return;
    double l = 0.0;
    for (int k = 0; k < a.nl; ++k) l += a.lcoef[k] * a.lterm[k][b];
    T s = mk<T>(1.0);
    for (int k = 0; k < a.ns; ++k) {
        T v = a.sterm[k][b];
        s = s * (a.sconj[k] ? conj_(v) : v);
    }
    if (a.logdet) a.logdet[b] = l;
    if (a.sign) a.sign[b] = s;
}

// out[i] = the MULTIPLIER that realises ScaleMode `mode` with d[i]
__global__ void scale_prep_kernel(double* out, const double* d, int mode, long long count) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double v = d[i];
    double r = v;
    switch (mode) {
        case SCALE_DIV: r = 1.0 / v; break;
        case SCALE_MUL_MIN1: r = fmin(v, 1.0); break;
        case SCALE_DIV_MAX1: r = 1.0 / fmax(v, 1.0); break;
        case SCALE_MUL_MAX1: r = fmax(v, 1.0); break;
        case SCALE_DIV_MIN1: r = 1.0 / fmin(v, 1.0); break;
        default: break;
    }
    out[i] = r;
}

__global__ void hs_exp_kernel(double* ev, const signed char* f, double ep, double em, long long count) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) ev[i] = f[i] > 0 ? ep : em;
}

// out[0..3] = [sum logdet, sum Re sign, sum Im sign, batch]: the local part of the cross-walker reduction, in a FIXED
// summation order (one block, strided partial sums, tree over warps) so that the result does not depend on timing
template <typename T> __global__ void walker_sums_kernel(double* out, const double* logdet, const T* sign, int batch) {
    __shared__ double red[3][32];
    double a = 0.0, b = 0.0, c = 0.0;
    for (int i = threadIdx.x; i < batch; i += blockDim.x) {
        a += logdet[i];
        b += real_(sign[i]);
        c += imag_(sign[i]);
    }
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; red[2][threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int nw = blockDim.x >> 5;
        a = threadIdx.x < nw ? red[0][threadIdx.x] : 0.0;
        b = threadIdx.x < nw ? red[1][threadIdx.x] : 0.0;
        c = threadIdx.x < nw ? red[2][threadIdx.x] : 0.0;
        a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
        if (threadIdx.x == 0) { out[0] = a; out[1] = b; out[2] = c; out[3] = (double)batch; }
    }
}

inline dim3 grid_rows_cols(int n, int batch, int threads) {
    return dim3((n + threads - 1) / threads, min(n, 64), batch);
}

}  // namespace

template <typename T>
cudaError_t scale_batched(T* dst, long long sdst, int lddst, const T* src, long long ssrc, int ldsrc, int op,
                          const double* rs, long long srs, int rs_mode, const double* cs, long long scs,
                          int cs_mode, int n, int batch, cudaStream_t s) {
    if (n <= 0 || batch <= 0) return cudaSuccess;
    for (int b0 = 0; b0 < batch; b0 += 65535) {
        int nb = min(65535, batch - b0);
        dim3 grid = (op == OP_N) ? grid_rows_cols(n, nb, 256) : dim3(min((n + 31) / 32, 64), min((n + 31) / 32, 64), nb);
        scale_kernel<T><<<grid, 256, 0, s>>>(dst + (long long)b0 * sdst, sdst, lddst, src + (long long)b0 * ssrc, ssrc,
                                              ldsrc, op, rs ? rs + (long long)b0 * srs : nullptr, srs, rs_mode,
                                              cs ? cs + (long long)b0 * scs : nullptr, scs, cs_mode, n);
        ++g_launch_count;
    }
    return cudaGetLastError();
}

template <typename T>
cudaError_t set_identity_batched(T* A, long long sA, int lda, int n, int batch, cudaStream_t s) {
    if (n <= 0 || batch <= 0) return cudaSuccess;
    for (int b0 = 0; b0 < batch; b0 += 65535) {
        int nb = min(65535, batch - b0);
        identity_kernel<T><<<grid_rows_cols(n, nb, 256), 256, 0, s>>>(A + (long long)b0 * sA, sA, lda, n);
        ++g_launch_count;
    }
    return cudaGetLastError();
}

cudaError_t fill_batched(double* d, double value, long long count, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    fill_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(d, value, count);
    ++g_launch_count;
    return cudaGetLastError();
}

template <typename T>
cudaError_t ipa_combine_batched(T* M, T* Rinv, const T* L, const double* d, int n, int batch, cudaStream_t s) {
    if (n <= 0 || batch <= 0) return cudaSuccess;
    for (int b0 = 0; b0 < batch; b0 += 65535) {
        int nb = min(65535, batch - b0);
        long long off = (long long)b0 * n * n;
        ipa_combine_kernel<T><<<grid_rows_cols(n, nb, 256), 256, 0, s>>>(M + off, Rinv + off, L + off,
                                                                          d + (long long)b0 * n, n);
        ++g_launch_count;
    }
    return cudaGetLastError();
}

cudaError_t logsum_batched(double* out, const double* d, long long sd, int mode, int n, int batch, cudaStream_t s) {
    if (batch <= 0) return cudaSuccess;
    logsum_kernel<<<batch, 256, 0, s>>>(out, d, sd, mode, n);
    ++g_launch_count;
    return cudaGetLastError();
}

template <typename T> cudaError_t combine_batched(const CombineArgs<T>& a, cudaStream_t s) {
    if (a.batch <= 0) return cudaSuccess;
    combine_kernel<T><<<(a.batch + 127) / 128, 128, 0, s>>>(a);
    ++g_launch_count;
    return cudaGetLastError();
}

cudaError_t scale_prep_batched(double* out, const double* d, int mode, long long count, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    scale_prep_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(out, d, mode, count);
    ++g_launch_count;
    return cudaGetLastError();
}

cudaError_t hs_exp_batched(double* ev, const signed char* fields, double lam, long long count, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    hs_exp_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(ev, fields, exp(lam), exp(-lam), count);
    ++g_launch_count;
    return cudaGetLastError();
}

// out[b] = max_ij |A[b](i,j) - B[b](i,j)|: the wrap error a DQMC sweep monitors at every re-stabilisation (sla_sweep).
// One CTA per matrix; NaNs propagate (fmax would drop them) so that a broken propagation cannot hide.
template <typename T> __global__ void maxabsdiff_kernel(double* out, const T* A, const T* B, long long nn, long long stride) {
    __shared__ double red[8];
    __shared__ int bad[8];
    const T* a = A + (long long)blockIdx.x * stride;
    const T* b = B + (long long)blockIdx.x * stride;
    double m = 0.0;
    int nanf = 0;
    for (long long i = threadIdx.x; i < nn; i += blockDim.x) {
        const double v = abs_(a[i] - b[i]);
        nanf |= !(v == v);
        m = fmax(m, v);
    }
    for (int o = 16; o > 0; o >>= 1) { m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o)); nanf |= __shfl_xor_sync(0xffffffffu, nanf, o); }
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = m; bad[threadIdx.x >> 5] = nanf; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { m = fmax(m, red[w]); nanf |= bad[w]; }
        out[blockIdx.x] = nanf ? nan("") : m;
    }
}
template <typename T> cudaError_t maxabsdiff_batched(double* out, const T* A, const T* B, long long nn, long long stride, int batch, cudaStream_t s) {
    if (batch <= 0) return cudaSuccess;
    maxabsdiff_kernel<T><<<batch, 256, 0, s>>>(out, A, B, nn, stride);
    ++g_launch_count;
    return cudaGetLastError();
}
template cudaError_t maxabsdiff_batched<double>(double*, const double*, const double*, long long, long long, int, cudaStream_t);
template cudaError_t maxabsdiff_batched<cplx>(double*, const cplx*, const cplx*, long long, long long, int, cudaStream_t);

// ---- Float32 storage types (sla_capi_f32.cu): promotion / demotion of operand arrays, identity in the storage type
__global__ void cvt_f32_f64_kernel(double* dst, const float* src, long long count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) dst[i] = (double)src[i];
}
__global__ void cvt_f64_f32_kernel(float* dst, const double* src, long long count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) dst[i] = (float)src[i];   // round to nearest even; out-of-range values become +-Inf like a Float32 store
}
__global__ void identity_f32_kernel(float* A, long long nn, int n, int reals, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long e = i / reals;            // element index over the batch
    const int part = (int)(i % reals);
    const long long m = e % nn;
    A[i] = (part == 0 && (m / n) == (m % n)) ? 1.0f : 0.0f;
}
cudaError_t convert_f32_f64(double* dst, const float* src, long long count, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    cvt_f32_f64_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(dst, src, count);
    ++g_launch_count;
    return cudaGetLastError();
}
cudaError_t convert_f64_f32(float* dst, const double* src, long long count, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    cvt_f64_f32_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(dst, src, count);
    ++g_launch_count;
    return cudaGetLastError();
}
cudaError_t set_identity_f32(float* A, long long nn, int n, int batch, bool complex_, cudaStream_t s) {
    const int reals = complex_ ? 2 : 1;
    const long long total = nn * batch * reals;
    if (total <= 0) return cudaSuccess;
    identity_f32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(A, nn, n, reals, total);
    ++g_launch_count;
    return cudaGetLastError();
}

template <typename T> cudaError_t walker_sums_batched(double* out4, const double* logdet, const T* sign, int batch, cudaStream_t s) {
    walker_sums_kernel<T><<<1, 256, 0, s>>>(out4, logdet, sign, batch);
    ++g_launch_count;
    return cudaGetLastError();
}
template cudaError_t walker_sums_batched<double>(double*, const double*, const double*, int, cudaStream_t);
template cudaError_t walker_sums_batched<cplx>(double*, const double*, const cplx*, int, cudaStream_t);

#define INST(T)                                                                                                   \
    template cudaError_t scale_batched<T>(T*, long long, int, const T*, long long, int, int, const double*,      \
                                          long long, int, const double*, long long, int, int, int, cudaStream_t); \
    template cudaError_t set_identity_batched<T>(T*, long long, int, int, int, cudaStream_t);                     \
    template cudaError_t ipa_combine_batched<T>(T*, T*, const T*, const double*, int, int, cudaStream_t);         \
    template cudaError_t combine_batched<T>(const CombineArgs<T>&, cudaStream_t);
INST(double)
INST(cplx)

}  // namespace sla
