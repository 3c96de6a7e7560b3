// Q = H_0 H_1 ... H_{n-1} I on the FP64 tensor cores (real, n <= 256): the Q-formation launch that follows a
// factorisation-only launch of the register LDR kernel (sla_ldr_reg.cu, MODE 1: reflectors below the diagonal
// of L, taus in the workspace).  Replaces `orgqr!` (src/qr.jl:95-109, called from src/LDR.jl:186).
//
// Q formation has no pivoting, so the reflectors are applied eight at a time in compact-WY form,
//     Q <- Q - V (T (V^T Q)),
// and every product is a DMMA.8x8x4 chain whose operands are ALREADY where the tensor core wants them:
//   * a warp owns 8 columns of Q (all 256 rows) as accumulator fragments: lane (fi = lane/4, fk = lane%4) holds
//     rows 8s + 2fk + {0,1} of column fi of its group, s = 0..31 -- 64 doubles per lane, never moved;
//   * Z^T = Q^T V  : A operand = the lane's own Q value (the k-slot <-> row assignment is free, so it is chosen
//     to be the row the lane holds), B operand = V from shared memory (one 16-byte load per two steps);
//   * Y^T = Z^T T^T: two DMMAs, A = the Z^T fragment as it came out, B = T from shared memory;
//   * Q^T -= Y^T V^T: accumulates straight into the Q fragments, A = the Y^T fragment as it came out.
// No warp shuffle anywhere; the transposed reductions of the vector-unit version (one per reflector, then one per
// four) are gone.  T (8x8, ?larft recurrence) comes from the Gram matrix V^T V, itself eight DMMAs per warp.
// A cluster of 4 CTAs forms the Q of one matrix (column c -> CTA c % 4); the only cluster-wide step is the barrier
// before Q overwrites the reflectors in L.
//
// STATUS (round 1): correct (orthogonality 2e-15, all sizes), but NOT faster than the vector-unit MODE 2 launch of
// sla_ldr_reg.cu: 0.325 ms vs 0.275 ms for 64 matrices of N = 256 -- every warp re-reads the whole chunk of V from shared
// memory in both sweeps (8x redundant traffic, DMMA pipe 18 % busy), and the per-chunk barriers leave the light column
// groups waiting for the heavy one.  It is therefore OFF by default (SLA_Q_DMMA=1 selects it); a next version has to
// give a warp more columns per V fragment (two warps per row half with a shared-memory reduction of Z).
#include <cooperative_groups.h>

#include "sla_cta.cuh"
#include "sla_kernels.h"

namespace cg = cooperative_groups;

namespace sla {

namespace {

constexpr int OQ_THREADS = 256;   // 8 warps: one group of 8 columns each
constexpr int OQ_LDV = 264;       // row stride of a reflector in the ring: = 8 (mod 16) doubles -> the 16-byte B loads of
                                  // the Z sweep are bank-conflict free
constexpr int OQ_RING = 2 * 8 * OQ_LDV;
// (the update sweep's loads are 4-way bank-conflicted with this stride; an XOR row swizzle that removes the conflicts
// cost more in index math than it saved: 0.335 vs 0.325 ms)
__device__ __forceinline__ int oq_swz(int) { return 0; }
constexpr size_t OQ_SMEM = (size_t)(OQ_RING + 8 * 64 + 8 * 64 + 256) * sizeof(double) + 64;

__global__ void __launch_bounds__(OQ_THREADS, 1) orgq_dmma_kernel(LdrArgs<double> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* const ring = reinterpret_cast<double*>(smem_raw);   // [2][8][OQ_LDV]
    double* const Gp = ring + OQ_RING;                          // [8 warps][64] partial Gram matrices
    double* const Tw = Gp + 8 * 64;                             // [8 warps][64]  T (row-major), one private copy per warp
    double* const taus = Tw + 8 * 64;                           // [256]

    const int n = p.n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fi = lane >> 2, fk = lane & 3;
    const int rank = (int)cg::this_cluster().block_rank();
    const int b = blockIdx.x / 4;
    double* Lg = p.L + (long long)b * p.strideL;
    const int ldl = p.ldl;
    const double* tau_g = p.Twork + (long long)b * p.strideT;
    // column groups: group g holds columns [32 g, 32 g + 32); the work grows with g (column c is touched by the
    // reflectors k <= c), warps w and w + 4 share an SM sub-partition -> pair a heavy group with a light one.
    // (Spreading a warp's columns over the whole range balances the warps but was slower: 0.385 vs 0.325 ms.)
    const int g = (warp < 4) ? 7 - warp : warp - 4;
    const int c = 32 * g + 4 * fi + rank;   // this lane's column of Q

    for (int r = tid; r < 256; r += OQ_THREADS) taus[r] = (r < n) ? tau_g[r] : 0.0;
    for (int r = tid; r < OQ_RING; r += OQ_THREADS) ring[r] = 0.0;   // rows >= n stay zero for ever

    double Qt[32][2];
#pragma unroll
    for (int s = 0; s < 32; ++s) {
        Qt[s][0] = (8 * s + 2 * fk == c && c < n) ? 1.0 : 0.0;
        Qt[s][1] = (8 * s + 2 * fk + 1 == c && c < n) ? 1.0 : 0.0;
    }
    __syncthreads();

    const int nchunk = (n + 7) / 8;
    const bool vecq = (((size_t)ldl * sizeof(double)) % 16 == 0) && ((((uintptr_t)Lg) & 15) == 0) && (n % 2 == 0);
    auto load_chunk = [&](int ch, int buf) {   // 32 threads per reflector
        const int i = tid >> 5, k = ch * 8 + i;
        double* dst = ring + (size_t)buf * 8 * OQ_LDV + (size_t)i * OQ_LDV;
        if (k < n) {
            const double* src = Lg + (long long)k * ldl;
            if (vecq) {
                for (int r = (tid & 31) * 2; r < n; r += 64) cp_async16(dst + (r ^ oq_swz(i)), src + r, 16);
            } else {
                for (int r = (tid & 31); r < n; r += 32) dst[r ^ oq_swz(i)] = src[r];
            }
        }
        cp_async_commit();
    };
    load_chunk(nchunk - 1, (nchunk - 1) & 1);

    for (int ch = nchunk - 1; ch >= 0; --ch) {
        cp_async_wait<0>();
        __syncthreads();   // chunk ch has landed; everybody is done with chunk ch + 1 (the other buffer)
        if (ch > 0) load_chunk(ch - 1, (ch - 1) & 1);
        const double* vb = ring + (size_t)(ch & 1) * 8 * OQ_LDV;
        const int k0 = ch * 8, kcnt = min(8, n - k0), s0 = ch;   // rows < k0 = 8 s0 are zero in every reflector of the chunk
        // reflector j of the chunk, row k0 + r8 (only block s0 needs the mask): 0 above the diagonal, 1 on it;
        // reflectors beyond n do not exist
        auto masked = [&](double raw, int j, int r8) { return (r8 < j) ? 0.0 : (r8 == j ? 1.0 : raw); };
        const bool jx_fi = fi < kcnt;   // reflector fi exists

        // ---- Gram matrix G = V^T V: every warp sums a share of the row steps on the tensor core
        {
            double g0 = 0.0, g1 = 0.0;
            for (int t = 2 * s0 + warp; t < 64; t += 8) {
                double x = vb[(size_t)fi * OQ_LDV + ((4 * t + fk) ^ oq_swz(fi))];
                if (t < 2 * s0 + 2) x = masked(x, fi, 4 * t + fk - k0);
                if (!jx_fi) x = 0.0;
                dmma884(g0, g1, x, x);   // G[j = fi][j' = 2 fk + e]
            }
            Gp[warp * 64 + fi * 8 + 2 * fk] = g0;
            Gp[warp * 64 + fi * 8 + 2 * fk + 1] = g1;
        }
        __syncthreads();
        if (32 * g + 31 < k0) continue;   // none of this warp's columns is touched by the chunk (warp-uniform)

        // ---- T (forward, column-wise ?larft) -- every active warp builds its own copy (no second block barrier)
        double* const Tm = Tw + warp * 64;
        {
            double gs0 = 0.0, gs1 = 0.0;   // entries lane and lane + 32 of G
#pragma unroll
            for (int w = 0; w < 8; ++w) { gs0 += Gp[w * 64 + lane]; gs1 += Gp[w * 64 + 32 + lane]; }
            Tm[lane] = gs0;                // G staged in the T buffer, turned into T row by row below
            Tm[32 + lane] = gs1;
            __syncwarp();
            double Trow[8];
            if (lane < 8) {                // lane i builds row i of T:  T(i,j) = -tau_j sum_{l=i}^{j-1} T(i,l) G(l,j)
                const int i = lane;
                const double ti = taus[k0 + i];
#pragma unroll
                for (int j = 0; j < 8; ++j) Trow[j] = (j == i) ? ti : 0.0;
#pragma unroll
                for (int j = 1; j < 8; ++j) {
                    if (j > i) {
                        double acc = 0.0;
#pragma unroll
                        for (int l = 0; l < 8; ++l)
                            if (l >= i && l < j) acc = fma(Trow[l], Tm[l * 8 + j], acc);
                        Trow[j] = -taus[k0 + j] * acc;
                    }
                }
            }
            __syncwarp();
            if (lane < 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) Tm[lane * 8 + j] = Trow[j];
            }
            __syncwarp();
        }

        // ---- Z^T = Q^T V over the rows >= k0
        // (four independent accumulator pairs: a single DMMA accumulation chain would be latency-bound)
        double za[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) { za[i][0] = 0.0; za[i][1] = 0.0; }
#pragma unroll
        for (int s = 0; s < 32; ++s) {
            if (s >= s0) {
                double2 vv = *reinterpret_cast<const double2*>(vb + (size_t)fi * OQ_LDV + ((8 * s + 2 * fk) ^ oq_swz(fi)));
                if (s == s0) { vv.x = masked(vv.x, fi, 2 * fk); vv.y = masked(vv.y, fi, 2 * fk + 1); }
                if (!jx_fi) { vv.x = 0.0; vv.y = 0.0; }
                dmma884(za[(2 * s) & 3][0], za[(2 * s) & 3][1], Qt[s][0], vv.x);   // Z^T[col = fi][j = 2 fk + e]
                dmma884(za[(2 * s + 1) & 3][0], za[(2 * s + 1) & 3][1], Qt[s][1], vv.y);
            }
        }
        const double z0 = (za[0][0] + za[1][0]) + (za[2][0] + za[3][0]);
        const double z1 = (za[0][1] + za[1][1]) + (za[2][1] + za[3][1]);
        // ---- Y^T = Z^T T^T   (k slot (fk, kk) <-> i = 2 fk + kk: the fragment as it came out)
        double y0 = 0.0, y1 = 0.0;
        {
            const double2 t2 = *reinterpret_cast<const double2*>(Tm + fi * 8 + 2 * fk);   // T[j = fi][i = 2 fk + {0,1}]
            dmma884(y0, y1, z0, t2.x);
            dmma884(y0, y1, z1, t2.y);
        }
        // ---- Q^T -= Y^T V^T   (k slot (fk, kk) <-> j = 2 fk + kk), straight into the Q fragments
        {
            const double ny0 = -y0, ny1 = -y1;
            const double* va = vb + (size_t)(2 * fk) * OQ_LDV + fi;
            const int swz = oq_swz(2 * fk);   // the same for reflectors 2 fk and 2 fk + 1
            const bool ja = 2 * fk < kcnt, jb = 2 * fk + 1 < kcnt;
#pragma unroll
            for (int s = 0; s < 32; ++s) {
                if (s >= s0) {
                    double b0 = va[(8 * s) ^ swz], b1 = va[OQ_LDV + ((8 * s) ^ swz)];
                    if (s == s0) { b0 = masked(b0, 2 * fk, fi); b1 = masked(b1, 2 * fk + 1, fi); }
                    if (!ja) b0 = 0.0;
                    if (!jb) b1 = 0.0;
                    dmma884(Qt[s][0], Qt[s][1], ny0, b0);
                    dmma884(Qt[s][0], Qt[s][1], ny1, b1);
                }
            }
        }
    }
    cg::this_cluster().sync();   // every CTA of the matrix is done reading the reflectors from L

    if (c < n) {
        double* dst = Lg + (long long)c * ldl;
        const bool v2 = (((size_t)ldl * sizeof(double)) % 16 == 0) && ((((uintptr_t)Lg) & 15) == 0);
#pragma unroll
        for (int s = 0; s < 32; ++s) {
            const int r = 8 * s + 2 * fk;
            if (v2 && r + 1 < n) *reinterpret_cast<double2*>(dst + r) = make_double2(Qt[s][0], Qt[s][1]);
            else {
                if (r < n) dst[r] = Qt[s][0];
                if (r + 1 < n) dst[r + 1] = Qt[s][1];
            }
        }
    }
}

}  // namespace

// Q formation for real matrices with n <= 256 after a MODE 1 launch of ldr_reg_kernel (taus in p.Twork)
cudaError_t orgq_dmma_batched(const LdrArgs<double>& p, cudaStream_t stream) {
    if (p.batch <= 0 || p.n <= 0) return cudaSuccess;
    if (p.n > 256 || p.Twork == nullptr || p.strideT < p.n) return cudaErrorNotSupported;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(orgq_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OQ_SMEM);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)p.batch * 4, 1, 1);
    cfg.blockDim = dim3(OQ_THREADS, 1, 1);
    cfg.dynamicSmemBytes = OQ_SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 4;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, orgq_dmma_kernel, p);
    ++g_launch_count;
    return e;
}

}  // namespace sla
