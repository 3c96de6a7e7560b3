// Q = H_0 H_1 ... H_{n-1} I on the FP64 tensor cores (real, n <= 256): the Q-formation launch that follows a
// factorisation-only launch of the register LDR kernel (sla_ldr_reg.cu, MODE 1: reflectors below the diagonal
// of L, taus in the workspace).  Replaces `orgqr!` (src/qr.jl:95-109, called from src/LDR.jl:186).
//
// Q formation has no pivoting, so the reflectors are applied eight at a time in compact-WY form,
//     Q <- Q - V (T (V^T Q)),
// and every product is a DMMA.8x8x4 chain whose operands are ALREADY where the tensor core wants them:
//   * Q lives in accumulator fragments: lane (fi = lane/4, fk = lane%4) of warp w holds rows 8s + 2fk + {0,1}, s = w + 8i,
//     of column 32g + 4fi + (CTA rank) for the 8 column groups g -- 64 doubles per lane, never moved;
//   * Z^T = Q^T V  : A operand = the lane's own Q value (the k-slot <-> row assignment is free, so it is chosen
//     to be the row the lane holds), B operand = V from shared memory (one 16-byte load per two steps);
//   * Y^T = Z^T T^T: two DMMAs, A = the Z^T fragment as it came out, B = T from shared memory;
//   * Q^T -= Y^T V^T: accumulates straight into the Q fragments, A = the Y^T fragment as it came out.
// No warp shuffle anywhere; the transposed reductions of the vector-unit version (one per reflector, then one per
// four) are gone.  T (8x8, ?larft recurrence) comes from the Gram matrix V^T V, itself eight DMMAs per warp.
// A cluster of 4 CTAs forms the Q of one matrix (column c -> CTA c % 4); the only cluster-wide step is the barrier
// before Q overwrites the reflectors in L.
//
// Round-1 history: a column-distributed version (a warp owns 8 columns, all rows) was correct but slower than the
// vector-unit launch (0.325 vs 0.275 ms for 64 matrices of N = 256): every warp re-read the whole chunk of V in both
// sweeps and the per-chunk barriers left the light column groups waiting for the heavy one.  The ROW-distributed
// version below loads each V fragment once per row block for all column groups and sheds dead rows evenly: 0.215 ms.
#include <cooperative_groups.h>

#include "sla_cta.cuh"
#include "sla_kernels.h"

namespace cg = cooperative_groups;

namespace sla {

namespace {

constexpr int OQ_THREADS = 256;   // 8 warps
constexpr int OQ_LDV = 264;       // row stride of a reflector in the ring: = 8 (mod 16) doubles -> the 16-byte B loads of
                                  // the Z sweep are bank-conflict free
constexpr int OQ_NBUF = 3;       // chunks of reflectors in the ring
constexpr int OQ_RING = OQ_NBUF * 8 * OQ_LDV;
// Row swizzle of the ring: reflectors 2,3,6,7 of a chunk store row r at r ^ 8, so that the Z sweep (lanes = reflector
// fi, 16 bytes each) AND the update sweep (lanes = reflector 2 fk + kk, 8 consecutive rows) both need the minimal
// number of wavefronts (the update sweep is 4-way conflicted without it).
__device__ __forceinline__ int oq_swz(int j) { return ((j >> 1) & 1) << 3; }
// ring | partial Z^T [8 warps][NG groups][32 lanes][2] | T of every chunk [32][64] | scratch [8 warps][64] | Y^T [NG groups][32][2] | taus
template <int NG> constexpr size_t oq_smem() { return (size_t)(OQ_RING + 8 * NG * 64 + 32 * 64 + 8 * 64 + NG * 64 + 256) * sizeof(double) + 64; }

// ROW-DISTRIBUTED: warp w owns the row blocks s = w, w+8, w+16, w+24 (8 rows each) of ALL 64 columns of its CTA
// (8 column groups).  A V fragment is then loaded once per row block and used for every column group (a warp that
// owns columns re-reads the whole chunk of V: that version was bound by shared-memory traffic and by the barrier
// imbalance between heavy and light column groups), the dead rows (< k0) are shed evenly by all warps, and the
// price is one reduction of the partial Z^T through shared memory per chunk (warp g finishes column group g).
//
// NG = column groups per CTA (8 columns each): NG = 8 -> a cluster of 4 CTAs per matrix, one CTA per SM (Q = 128 registers
// per lane); NG = 4 -> a cluster of 8 CTAs with half the columns each, TWO CTAs resident per SM, so that one CTA's
// per-chunk barriers and shared-memory round trips overlap with the other's tensor-core work.
template <int NG>
__global__ void __launch_bounds__(OQ_THREADS, NG == 8 ? 1 : 2) orgq_dmma_kernel(LdrArgs<double> p) {
    constexpr int CS = 32 / NG;   // CTAs per matrix; column c = CS (8 g + fi) + rank
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* const ring = reinterpret_cast<double*>(smem_raw);   // [OQ_NBUF][8][OQ_LDV]
    double* const Zp = ring + OQ_RING;                          // [8 warps][8 groups][32 lanes][2]
    double* const Tall = Zp + 8 * NG * 64;                      // [32 chunks][64]  T of every chunk (row-major)
    double* const Tw = Tall + 32 * 64;                          // [8 warps][64]  scratch of the T prologue
    double* const Ys = Tw + 8 * 64;                             // [8 groups][32 lanes][2]
    double* const taus = Ys + NG * 64;                          // [256]

    const int n = p.n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fi = lane >> 2, fk = lane & 3;
    const int rank = (int)cg::this_cluster().block_rank();
    const int b = blockIdx.x / CS;
    double* Lg = p.L + (long long)b * p.strideL;
    const int ldl = p.ldl;
    const double* tau_g = p.Twork + (long long)b * p.strideT;

    for (int r = tid; r < 256; r += OQ_THREADS) taus[r] = (r < n) ? tau_g[r] : 0.0;
    for (int r = tid; r < OQ_RING; r += OQ_THREADS) ring[r] = 0.0;   // rows >= n stay zero for ever

    // lane (fi, fk) holds Q[row = 8 (warp + 8 i) + 2 fk + e][col = 32 g + 4 fi + rank]
    double Qt[NG][4][2];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const int c = CS * (8 * g + fi) + rank;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = 8 * (warp + 8 * i) + 2 * fk;
            Qt[g][i][0] = (r == c && c < n) ? 1.0 : 0.0;
            Qt[g][i][1] = (r + 1 == c && c < n) ? 1.0 : 0.0;
        }
    }
    // The three cluster-wide barriers of this kernel are split into arrive / wait so that their latency hides behind work
    // that does not depend on them (round 2: 10 % of the stall samples sat in cluster.sync()).
    auto cl_arrive = [] { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); };
    auto cl_wait = [] { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); };
    cl_arrive();   // #1: taus / ring are set (also the CTA-wide barrier for them) and this CTA runs

    const int nchunk = (n + 7) / 8;
    // ---- prologue: T of EVERY chunk (they depend on V only), straight from global memory: G = V^T V on the tensor
    // core, then the ?larft recurrence on 8 lanes -- off the per-chunk critical path.  The chunks are shared out over the
    // CTAs of the cluster (chunk cq -> CTA cq % CS, warp cq / CS) and each T is written into every CTA's shared memory
    // (DSMEM stores): a warp builds at most one or two T instead of four.
    // (n <= 256: at most 32 chunks over CS x 8 warps -- one chunk per warp, or none)
    const int cq = rank + CS * warp;
    const bool has_t = cq < nchunk;
    double g0a = 0.0, g1a = 0.0, g0b = 0.0, g1b = 0.0;
    if (has_t) {
        const int kj = cq * 8 + fi;                    // lane's reflector (column kj of L)
        const double* col = Lg + (long long)kj * ldl;
        for (int t = 2 * cq; t < 64; t += 8) {         // eight loads in flight, two independent accumulation chains
            double x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = 4 * (t + u) + fk;
                x[u] = (kj < n && r < n && r > kj) ? col[r] : ((r == kj && kj < n) ? 1.0 : 0.0);
            }
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
                dmma884(g0a, g1a, x[u], x[u]);         // G[j = fi][j' = 2 fk + e]
                dmma884(g0b, g1b, x[u + 1], x[u + 1]);
            }
        }
    }
    cl_wait();     // #1: every CTA of the cluster runs (its shared memory may be written); taus visible
    if (has_t) {
        const int k0 = cq * 8;
        double* const Gm = Tw + warp * 64;
        Gm[fi * 8 + 2 * fk] = g0a + g0b;
        Gm[fi * 8 + 2 * fk + 1] = g1a + g1b;
        __syncwarp();
        double Trow[8];
        if (lane < 8) {                                // lane i builds row i of T:  T(i,j) = -tau_j sum_{l=i}^{j-1} T(i,l) G(l,j)
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
                        if (l >= i && l < j) acc = fma(Trow[l], Gm[l * 8 + j], acc);
                    Trow[j] = -taus[k0 + j] * acc;
                }
            }
            for (int pr = 0; pr < CS; ++pr) {
                double* const Tpeer = cg::this_cluster().map_shared_rank(Tall, pr);
#pragma unroll
                for (int j = 0; j < 8; ++j) Tpeer[cq * 64 + i * 8 + j] = Trow[j];
            }
        }
        __syncwarp();
    }
    cl_arrive();   // #2: this CTA's T blocks are written everywhere
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
    // Three-buffer ring, two barriers per chunk: chunk ch - 2 is requested after the first barrier of chunk ch (everybody
    // has left the update of chunk ch + 1, the previous user of that buffer) and awaited before the second barrier of
    // chunk ch - 1, so a warp runs from its update of chunk ch straight into its partial Z of chunk ch - 1.
    load_chunk(nchunk - 1, (nchunk - 1) % OQ_NBUF);
    if (nchunk >= 2) load_chunk(nchunk - 2, (nchunk - 2) % OQ_NBUF); else cp_async_commit();
    cp_async_wait<1>();
    cl_wait();         // #2: every T has arrived
    __syncthreads();   // chunk nchunk - 1 has landed for everybody

    for (int ch = nchunk - 1; ch >= 0; --ch) {
        if (ch == 0) cl_arrive();   // #3: all of this CTA's reads of the reflectors in L have completed
        const double* vb = ring + (size_t)(ch % OQ_NBUF) * 8 * OQ_LDV;
        const int k0 = ch * 8, kcnt = min(8, n - k0), s0 = ch;   // rows < k0 = 8 s0 are zero in every reflector of the chunk
        const int g0 = k0 / (8 * CS);                            // column groups < g0 lie left of the chunk: untouched
        // reflector j of the chunk, row k0 + r8 (only block s0 needs the mask): 0 above the diagonal, 1 on it
        auto masked = [&](double raw, int j, int r8) { return (r8 < j) ? 0.0 : (r8 == j ? 1.0 : raw); };
        const bool jx_fi = fi < kcnt;                            // reflector fi exists
        const bool ja = 2 * fk < kcnt, jb = 2 * fk + 1 < kcnt;
        const int swz_z = oq_swz(fi), swz_u = oq_swz(2 * fk);

        // ---- partial Z^T over this warp's live row blocks
        {
            double z[NG][2];
#pragma unroll
            for (int g = 0; g < NG; ++g) { z[g][0] = 0.0; z[g][1] = 0.0; }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int sb = warp + 8 * i;
                if (sb >= s0) {
                    double2 vv = *reinterpret_cast<const double2*>(vb + (size_t)fi * OQ_LDV + ((8 * sb + 2 * fk) ^ swz_z));
                    if (sb == s0) { vv.x = masked(vv.x, fi, 2 * fk); vv.y = masked(vv.y, fi, 2 * fk + 1); }
                    if (!jx_fi) { vv.x = 0.0; vv.y = 0.0; }
#pragma unroll
                    for (int g = 0; g < NG; ++g)
                        if (g >= g0) {
                            dmma884(z[g][0], z[g][1], Qt[g][i][0], vv.x);   // Z^T[col = fi][j = 2 fk + e]
                            dmma884(z[g][0], z[g][1], Qt[g][i][1], vv.y);
                        }
                }
            }
#pragma unroll
            for (int g = 0; g < NG; ++g)
                if (g >= g0) *reinterpret_cast<double2*>(Zp + ((size_t)(warp * NG + g) * 32 + lane) * 2) = make_double2(z[g][0], z[g][1]);
        }
        __syncthreads();   // partial Z^T written; everybody has left the update of chunk ch + 1
        if (ch >= 2) load_chunk(ch - 2, (ch - 2) % OQ_NBUF); else cp_async_commit();

        // ---- warp g finishes column group g: Z^T = sum of the partials, T, Y^T = Z^T T^T
        if (warp >= g0 && warp < NG) {
            const double* const Tm = Tall + ch * 64;   // from the prologue
            double z0 = 0.0, z1 = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                const double2 zp = *reinterpret_cast<const double2*>(Zp + ((size_t)(w * NG + warp) * 32 + lane) * 2);
                z0 += zp.x; z1 += zp.y;
            }
            // Y^T = Z^T T^T   (k slot (fk, kk) <-> i = 2 fk + kk: the fragment as it came out)
            double y0 = 0.0, y1 = 0.0;
            const double2 t2 = *reinterpret_cast<const double2*>(Tm + fi * 8 + 2 * fk);   // T[j = fi][i = 2 fk + {0,1}]
            dmma884(y0, y1, z0, t2.x);
            dmma884(y0, y1, z1, t2.y);
            *reinterpret_cast<double2*>(Ys + ((size_t)warp * 32 + lane) * 2) = make_double2(-y0, -y1);
        }
        cp_async_wait<1>();   // this thread's part of chunk ch - 1 (requested one chunk ago) has landed
        __syncthreads();      // Y^T written; chunk ch - 1 visible to everybody

        // ---- Q^T -= Y^T V^T on this warp's live row blocks, every live column group (k slot <-> j = 2 fk + kk)
        {
            double ny[NG][2];
#pragma unroll
            for (int g = 0; g < NG; ++g)
                if (g >= g0) {
                    const double2 y2 = *reinterpret_cast<const double2*>(Ys + ((size_t)g * 32 + lane) * 2);
                    ny[g][0] = y2.x; ny[g][1] = y2.y;
                } else { ny[g][0] = 0.0; ny[g][1] = 0.0; }
            const double* const va = vb + (size_t)(2 * fk) * OQ_LDV + fi;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int sb = warp + 8 * i;
                if (sb >= s0) {
                    double b0 = va[(8 * sb) ^ swz_u], b1 = va[OQ_LDV + ((8 * sb) ^ swz_u)];
                    if (sb == s0) { b0 = masked(b0, 2 * fk, fi); b1 = masked(b1, 2 * fk + 1, fi); }
                    if (!ja) b0 = 0.0;
                    if (!jb) b1 = 0.0;
#pragma unroll
                    for (int g = 0; g < NG; ++g)
                        if (g >= g0) {
                            dmma884(Qt[g][i][0], Qt[g][i][1], ny[g][0], b0);
                            dmma884(Qt[g][i][0], Qt[g][i][1], ny[g][1], b1);
                        }
                }
            }
        }
    }
    cl_wait();   // #3: every CTA of the matrix is done reading the reflectors from L

    {
        const bool v2 = (((size_t)ldl * sizeof(double)) % 16 == 0) && ((((uintptr_t)Lg) & 15) == 0);
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const int c = CS * (8 * g + fi) + rank;
            if (c >= n) continue;
            double* dst = Lg + (long long)c * ldl;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = 8 * (warp + 8 * i) + 2 * fk;
                if (v2 && r + 1 < n) *reinterpret_cast<double2*>(dst + r) = make_double2(Qt[g][i][0], Qt[g][i][1]);
                else {
                    if (r < n) dst[r] = Qt[g][i][0];
                    if (r + 1 < n) dst[r + 1] = Qt[g][i][1];
                }
            }
        }
    }
}

}  // namespace

template <int NG> static cudaError_t launch_orgq(const LdrArgs<double>& p, cudaStream_t stream) {
    constexpr int CS = 32 / NG;
    static bool configured[kMaxDevices] = {};   // the attribute is per device
    const int dev = current_device();
    if (dev < 0 || dev >= kMaxDevices || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(orgq_dmma_kernel<NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)oq_smem<NG>());
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < kMaxDevices) configured[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)p.batch * CS, 1, 1);
    cfg.blockDim = dim3(OQ_THREADS, 1, 1);
    cfg.dynamicSmemBytes = oq_smem<NG>();
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, orgq_dmma_kernel<NG>, p);
    ++g_launch_count;
    return e;
}

// Q formation for real matrices with n <= 256 after a MODE 1 launch of ldr_reg_kernel (taus in p.Twork)
cudaError_t orgq_dmma_batched(const LdrArgs<double>& p, cudaStream_t stream) {
    if (p.batch <= 0 || p.n <= 0) return cudaSuccess;
    if (p.n > 256 || p.Twork == nullptr || p.strideT < p.n) return cudaErrorNotSupported;
    // default: clusters of 8 CTAs with 32 columns each, two CTAs per SM -- 0.178 vs 0.190 ms alone for 64 x N=256, and
    // 32.24 vs 32.61 ms per C2 step (the half-size CTAs interleave with the side-stream GEMM CTAs; gpurun_out/r2_orgq_ng.log).
    // SLA_ORGQ_NG=8: clusters of 4 CTAs with 64 columns each, one CTA per SM (round 1)
    static const int ng = getenv("SLA_ORGQ_NG") ? atoi(getenv("SLA_ORGQ_NG")) : 4;
    if (ng == 8) return launch_orgq<8>(p, stream);
    return launch_orgq<4>(p, stream);
}

}  // namespace sla
