// Batched LDR factorisation for LARGE matrices and SMALL batches (C4: 32 chains of N = 1024): the L2-streaming
// algorithm of sla_ldr.cu (?laqps panel: V and F = tau A^H v in shared memory, ONE pass over the trailing matrix per
// column, rank-NB DMMA trailing update, compact-WY Q formation) with a CLUSTER of CS CTAs per matrix instead of one CTA.
//
// Replaces `ldr!(F, ws)` (src/LDR.jl:160-188 = geqp3! src/qr.jl:70-92 + orgqr! src/qr.jl:95-109) where one CTA per
// matrix leaves most of the GPU idle: the per-column pass streams the whole trailing matrix through ONE SM (8 MB at
// N = 1024, ~110 us per column), so 32 matrices keep 32 of 148 SMs busy.  Here the trailing matrix is split by COLUMN
// POSITION, block-cyclically (blocks of 32 positions) over the CTAs of a cluster:
//   * every CTA keeps the whole panel state (V, F, T, tau, the position -> column map) in its own shared memory and
//     redoes the cheap O(n NB) steps redundantly and deterministically, so that all copies stay bit-identical;
//   * the expensive steps -- the GEMV over the trailing matrix, the pivot-row update / norm down-date, the rank-NB
//     trailing update and all three products of the Q formation -- run on the CTA's own positions only;
//   * no physical column swaps: positions are mapped to matrix columns through `colidx` (swapping two ints in every
//     CTA replaces two column copies through global memory and the read/write hazard that would need a barrier);
//   * per column two cluster barriers: after the GEMV (its results are broadcast into every CTA's F through
//     distributed shared memory) and after the pivot search (each CTA's best column + the norms of the column that
//     changes position travel in a 32-byte message per CTA); one more per panel after the trailing update.
// Q is formed in place in L from a copy of the reflectors (LdrArgs::Qwork, n*n per matrix), because without swaps
// reflector k lives in matrix column colidx[k], which the backward accumulation would overwrite.
#include <cooperative_groups.h>

#include "sla_cta.cuh"
#include "sla_kernels.h"

namespace cg = cooperative_groups;

namespace sla {

namespace {

constexpr double TOL3Z_MC = 1.4901161193847656e-08;  // sqrt(eps), LAPACK ?laqps
constexpr int MC_BW = 32;                             // ownership block: 32 consecutive column positions
constexpr int MC_MAXCS = 8;

struct alignas(16) McMsg {   // one CTA's contribution to the pivot search
    double bv;               // its largest partial column norm (-1: none)
    int bi;                  // ... and that column's position
    int has_norm;            // this CTA owns the position that is filled next: v1 / v2 are its norms
    double v1, v2;
};

template <typename T> struct McSmem {
    T* Vs; T* Fs; T* Gp; T* Ts; T* tau; T* auxv;
    double* vn1; double* vn2; double* redv;
    int* colidx; int* flag; int* redi; int* nflag;
    McMsg* msgs;             // [2][MC_MAXCS]
    int ld;
};

template <typename T> __host__ __device__ inline size_t mc_smem_bytes(int n, int NB) {
    size_t ld = panel_ld<T>(n);
    size_t b = 0;
    b += 2 * NB * ld * sizeof(T);              // Vs, Fs
    b += 2 * (size_t)NB * NB * sizeof(T);      // Gp, Ts
    b += (size_t)n * sizeof(T);                // tau
    b += (size_t)NB * sizeof(T);               // auxv
    b += 2 * (size_t)n * sizeof(double);       // vn1, vn2
    b += 32 * sizeof(double);                  // redv
    b += 2 * MC_MAXCS * sizeof(McMsg) + 16;    // msgs (+ alignment)
    b += 2 * (size_t)n * sizeof(int);          // colidx, flag list
    b += 64 * sizeof(int);                     // redi + nflag
    return b + 64;
}

template <typename T> __device__ inline McSmem<T> mc_carve(unsigned char* raw, int n, int NB) {
    McSmem<T> s;
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
    dd = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(dd) + 15) & ~(uintptr_t)15);   // odd n: McMsg is 16-byte aligned
    s.msgs = reinterpret_cast<McMsg*>(dd); dd += 2 * MC_MAXCS * (sizeof(McMsg) / sizeof(double));
    int* ii = reinterpret_cast<int*>(dd);
    s.colidx = ii; ii += n;
    s.flag = ii; ii += n;
    s.redi = ii; ii += 32;
    s.nflag = ii;
    return s;
}

// Block-cyclic ownership of column positions: CTA `rank` owns the blocks rank, rank + CS, ... of MC_BW positions.  Its
// positions are enumerated by a contiguous "own index" o: block o / BW of its own, offset o % BW.
struct Own {
    int CS, rank, n;
    __device__ __forceinline__ int pos(int o) const { return ((o / MC_BW) * CS + rank) * MC_BW + (o % MC_BW); }
    __device__ __forceinline__ bool mine(int c) const { return (c / MC_BW) % CS == rank; }
    __device__ __forceinline__ int count() const {   // own indices (including positions >= n in the last block)
        const int nblk = (n + MC_BW - 1) / MC_BW;
        return nblk > rank ? ((nblk - rank + CS - 1) / CS) * MC_BW : 0;
    }
    // smallest own index whose position is >= c
    __device__ __forceinline__ int first_ge(int c) const {
        const int blk = c / MC_BW;
        const int ob = blk + ((rank - blk % CS) + CS) % CS;   // first own block >= blk
        int o = (ob / CS) * MC_BW;
        if (ob == blk) o += c % MC_BW;
        return o;
    }
};

// F(c) = tau * sum_{r >= k} conj(A[col(c)][r]) v[r] for the CTA's own positions c > k; the result goes into row `frow`
// of F in EVERY CTA of the cluster (fs_all[q] = that CTA's F row).  Same load pattern as trailing_gemv (sla_ldr.cu).
template <typename T, int UC>
__device__ __forceinline__ void mc_gemv(const T* __restrict__ A, int lda, const Own own, const int* __restrict__ colidx,
                                        const T* __restrict__ v, int k, T tau, T* const* fs_all, size_t frow_off, bool vec) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = own.n;
    const int o_lo = own.first_ge(k + 1), o_hi = own.count();
    const int ncols = o_hi - o_lo;
    if (ncols <= 0) return;
    const int mrem = n - (k & ~1);
    const int per_lane = (sizeof(T) == 8 && vec) ? 2 : 1;
    const int lpc = (mrem > 64 * per_lane) ? 32 : ((mrem > 32 * per_lane) ? 16 : 8);
    const int cpw = 32 / lpc;
    const int sub = lane / lpc, sl = lane % lpc;
    const int cols_per_pass = cpw * UC;
    for (int g = warp; g * cols_per_pass < ncols; g += CTA_WARPS) {
        const int obase = o_lo + g * cols_per_pass + sub;
        int cpos[UC];
        const T* cptr[UC];
#pragma unroll
        for (int u = 0; u < UC; ++u) {
            const int o = obase + u * cpw;
            const int c = (o < o_hi) ? own.pos(o) : n;
            cpos[u] = (c < n) ? c : -1;
            cptr[u] = A + (long long)colidx[c < n ? c : 0] * lda;
        }
        T acc[UC];
#pragma unroll
        for (int u = 0; u < UC; ++u) acc[u] = mk<T>(0.0);
        bool done = false;
        if constexpr (sizeof(T) == 8) {
            if (vec) {
                double acc2[UC];
#pragma unroll
                for (int u = 0; u < UC; ++u) acc2[u] = 0.0;
                for (int r0 = (k & ~1) + 2 * sl; r0 < n; r0 += 8 * lpc) {
                    double2 a[UC][4];
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int r = r0 + it * 2 * lpc;
#pragma unroll
                        for (int u = 0; u < UC; ++u)
                            a[u][it] = (r < n && cpos[u] >= 0) ? __ldcg(reinterpret_cast<const double2*>(cptr[u] + r))
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
                for (int u = 0; u < UC; ++u)
                    if (cpos[u] >= 0) fmacc(acc[u], conj_(cptr[u][r]), w);
            }
        }
#pragma unroll
        for (int u = 0; u < UC; ++u) {
            T a = acc[u];
            for (int o = lpc >> 1; o > 0; o >>= 1) {
                if constexpr (sizeof(T) == 8) a += __shfl_xor_sync(0xffffffffu, a, o);
                else { a.re += __shfl_xor_sync(0xffffffffu, a.re, o); a.im += __shfl_xor_sync(0xffffffffu, a.im, o); }
            }
            if (sl == 0 && cpos[u] >= 0) {
                const T f = tau * a;
                for (int q = 0; q < own.CS; ++q) fs_all[q][frow_off + cpos[u]] = f;
            }
        }
    }
}

// C[r][col(c)] -= sum_{i<kb} P[i][r] * (CONJU ? conj(U[i][c]) : U[i][c]) for r in [r_lo, r_hi) and the CTA's own
// positions c >= c_lo (c < n); cmap == nullptr: position = column.  Port of cta_rank_update (sla_cta.cuh) to the
// block-cyclic ownership: a warp tile is 32 rows x TC positions and never straddles an ownership block.
template <typename T, bool CONJU>
__device__ void mc_rank_update(T* __restrict__ C, int ldc, int r_lo, int r_hi, int c_lo, const Own own,
                               const int* __restrict__ cmap, const T* __restrict__ Ps, const T* __restrict__ Us, int ld,
                               int kb, int n_clamp) {
    const int n = own.n;
    if (r_lo >= r_hi || c_lo >= n) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fi = lane >> 2, fk = lane & 3;
    const int kb4 = (kb + 3) & ~3;
    const int r_base = r_lo & ~1;
    constexpr int NI = (sizeof(T) == 8) ? 4 : 2;
    constexpr int TC = NI * 8;
    static_assert(MC_BW % TC == 0, "a tile stays inside one ownership block");
    const int o_start = own.first_ge(c_lo) / TC * TC, o_hi = own.count();
    if (o_start >= o_hi) return;
    const int tiles_r = (r_hi - r_base + 31) / 32, tiles_c = (o_hi - o_start + TC - 1) / TC;
    const bool vec = (sizeof(T) == 8) && ((ldc & 1) == 0) && ((((uintptr_t)C) & 15) == 0);
    for (int t = warp; t < tiles_r * tiles_c; t += CTA_WARPS) {
        const int r0 = r_base + (t % tiles_r) * 32, o0 = o_start + (t / tiles_r) * TC;
        T acc[NI][4][2];
        int cp[NI];
        T* cptr[NI];
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int c = own.pos(o0 + i * 8 + fi);
            cp[i] = (c >= c_lo && c < n) ? c : -1;
            cptr[i] = C + (long long)(cp[i] >= 0 ? (cmap ? cmap[c] : c) : 0) * ldc;
        }
#pragma unroll
        for (int i = 0; i < NI; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = r0 + j * 8 + 2 * fk;
                T v0 = mk<T>(0.0), v1 = mk<T>(0.0);
                if (cp[i] >= 0) {
                    const T* src = cptr[i] + r;
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
            for (int i = 0; i < NI; ++i) {
                T u = Us[(kk + fk) * ld + (cp[i] >= 0 ? cp[i] : n_clamp)];
                if (cp[i] < 0) u = mk<T>(0.0);
                nf[i] = -(CONJU ? conj_(u) : u);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) mf[j] = Ps[(kk + fk) * ld + min(r0 + j * 8 + fi, n_clamp)];
#pragma unroll
            for (int i = 0; i < NI; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) mma_acc(acc[i][j], nf[i], mf[j]);
        }
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            if (cp[i] < 0) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = r0 + j * 8 + 2 * fk;
                T* dst = cptr[i] + r;
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
            }
        }
    }
}

// W[i][c] = sum_{r in [r_lo, r_hi)} conj(P[i][r]) * C[r][c] for the CTA's own COLUMNS c >= c_lo (position = column):
// port of cta_panel_dot (sla_cta.cuh); a warp tile is 8 own columns.
template <typename T>
__device__ void mc_panel_dot(T* __restrict__ Ws, const T* __restrict__ Ps, int ld, int kb, const T* __restrict__ C, int ldc,
                             int r_lo, int r_hi, int c_lo, const Own own, int n_clamp) {
    const int n = own.n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fi = lane >> 2, fk = lane & 3;
    const int nif = (kb + 7) / 8;
    const int o_start = own.first_ge(c_lo) & ~7, o_hi = own.count();
    if (o_start >= o_hi) return;
    const int tiles_c = (o_hi - o_start + 7) / 8;
    const int k_lo = r_lo & ~3;
    for (int t = warp; t < tiles_c; t += CTA_WARPS) {
        const int c = own.pos(o_start + t * 8 + fi);
        const bool cvalid = c >= c_lo && c < n;
        const T* col = C + (long long)(cvalid ? c : 0) * ldc;
        T acc[4][2];
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[j][0] = mk<T>(0.0); acc[j][1] = mk<T>(0.0); }
        // eight row steps per trip: all eight global loads of C are in flight before the first DMMA (one L2 round trip
        // per 32 rows instead of one per 4: this loop is latency-bound, profiles/ldr_stream_phases_r2.txt)
        T cn[8];                                       // the NEXT trip's eight loads are issued before this trip's DMMAs
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int r = k_lo + 4 * u + fk;
            cn[u] = (cvalid && r >= r_lo && r < r_hi) ? col[r] : mk<T>(0.0);
        }
        for (int kk0 = k_lo; kk0 < r_hi; kk0 += 32) {
            T cf[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) cf[u] = cn[u];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = kk0 + 32 + 4 * u + fk;
                cn[u] = (cvalid && r >= r_lo && r < r_hi) ? col[r] : mk<T>(0.0);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = kk0 + 4 * u + fk;
                if (kk0 + 4 * u < r_hi) {
                    const int rc = min(r, n_clamp);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (j < nif) {
                            T pf = conj_(Ps[(j * 8 + fi) * ld + rc]);
                            mma_acc(acc[j], cf[u], pf);
                        }
                    }
                }
            }
        }
        if (cvalid) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j < nif) {
                    Ws[(j * 8 + 2 * fk) * ld + c] = acc[j][0];
                    Ws[(j * 8 + 2 * fk + 1) * ld + c] = acc[j][1];
                }
            }
        }
    }
}

template <typename T, int NB>
__global__ void __launch_bounds__(CTA_THREADS, 1) ldr_mc_kernel(LdrArgs<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::cluster_group cl = cg::this_cluster();
    const int n = p.n;
    McSmem<T> s = mc_carve<T>(smem_raw, n, NB);
    const int ld = s.ld;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int CS = (int)cl.num_blocks(), rank = (int)cl.block_rank();
    const int b = blockIdx.x / CS;
    const Own own{CS, rank, n};
    T* A = p.L + (long long)b * p.strideL;
    const int lda = p.ldl;
    T* Tw = p.Twork + (long long)b * p.strideT;
    T* Qw = p.Qwork + (long long)b * p.strideQ;
    const bool vec = (sizeof(T) == 8) && ((lda & 1) == 0) && ((n & 1) == 0) && ((((uintptr_t)A) & 15) == 0);
    // this CTA's view of every CTA's F panel and message board (distributed shared memory)
    T* fs_all[MC_MAXCS];
    McMsg* msg_all[MC_MAXCS];
#pragma unroll
    for (int q = 0; q < MC_MAXCS; ++q) {
        fs_all[q] = (q < CS) ? cl.map_shared_rank(s.Fs, q) : nullptr;
        msg_all[q] = (q < CS) ? cl.map_shared_rank(s.msgs, q) : nullptr;
    }

    // pivot search across the cluster: every CTA posts its best own position (and, if it owns position `next`, that
    // position's norms) on every CTA's board, one cluster barrier, everybody reduces the same CS messages
    double nx1 = 0.0, nx2 = 0.0;   // norms of the column that sat at position `next` (valid after exchange)
    auto exchange = [&](double bv, int bi, int next, int parity) -> int {
        if (tid < CS) {
            McMsg m;
            m.bv = bv; m.bi = bi;
            m.has_norm = (next < n && own.mine(next)) ? 1 : 0;
            m.v1 = m.has_norm ? s.vn1[next] : 0.0;
            m.v2 = m.has_norm ? s.vn2[next] : 0.0;
            msg_all[tid][parity * MC_MAXCS + rank] = m;
        }
        cl.sync();
        double fv = -1.0; int fidx = 0x7fffffff;
        for (int q = 0; q < CS; ++q) {
            const McMsg m = s.msgs[parity * MC_MAXCS + q];
            if (m.bv > fv || (m.bv == fv && m.bi < fidx)) { fv = m.bv; fidx = m.bi; }
            if (m.has_norm) { nx1 = m.v1; nx2 = m.v2; }
        }
        return fidx;
    };

    // ---------------------------------------------------------------- init: own column norms, colidx = identity
    const T* Ain = p.Ain ? p.Ain + (long long)b * p.strideAin : nullptr;
    const int n_own = own.count();
    for (int o = warp; o < n_own; o += CTA_WARPS) {
        const int c = own.pos(o);
        if (c >= n) continue;
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
        if (lane == 0) { double nr = sqrt(acc); s.vn1[c] = nr; s.vn2[c] = nr; }
    }
    for (int c = tid; c < n; c += CTA_THREADS) s.colidx[c] = c;
    __syncthreads();
    int pvt;
    {
        double bv = -1.0; int bi = 0x7fffffff;
        for (int o = tid; o < n_own; o += CTA_THREADS) {
            const int c = own.pos(o);
            if (c < n) { double v = s.vn1[c]; if (v > bv) { bv = v; bi = c; } }
        }
        const int lb = block_argmax(bv, bi, s.redv, s.redi);
        const double lv = (lb < n) ? s.vn1[lb] : -1.0;
        __threadfence();   // the copy of Ain into L is read by the other CTAs after the barrier inside exchange()
        pvt = exchange(lb < n ? lv : -1.0, lb, 0, 1);
        if (pvt >= n) pvt = 0;  // all-NaN guard
    }

    // ---------------------------------------------------------------- pivoted QR, panel by panel
    for (int p0 = 0; p0 < n; p0 += NB) {
        const int jb = min(NB, n - p0);
        for (int i = tid; i < NB * NB; i += CTA_THREADS) s.Gp[i] = mk<T>(0.0);
        for (int j = 0; j < jb; ++j) {
            const int k = p0 + j;
            T* vj = s.Vs + (size_t)j * ld;
            // ---- P1 (every CTA): fetch the pivot column, apply the panel's reflectors to it
            const T* pcol = A + (long long)s.colidx[pvt] * lda;
            double ss = 0.0;
            for (int r = tid; r < n; r += CTA_THREADS) {
                if (r >= k) {
                    T x = pcol[r];
                    for (int i = 0; i < j; ++i) fms(x, s.Vs[(size_t)i * ld + r], conj_(s.Fs[(size_t)i * ld + pvt]));
                    vj[r] = x;
                    if (r > k) ss += abs2_(x);
                } else {
                    vj[r] = mk<T>(0.0);
                }
            }
            ss = warp_sum(ss);
            if (lane == 0) s.redv[warp] = ss;
            __syncthreads();
            // ---- P2: Householder reflector (LAPACK ?larfg), identical in every CTA
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
            __syncthreads();  // everyone has read vj[k] and redv, and colidx[pvt] / Fs[:, pvt]
            for (int r = tid; r < n; r += CTA_THREADS) {
                if (r > k) vj[r] = vj[r] * scal;
                else if (r == k) vj[r] = mk<T>(1.0);
            }
            // position bookkeeping (replicated): the column at position k moves to position pvt
            if (pvt != k) {
                if (tid < j) s.Fs[(size_t)tid * ld + pvt] = s.Fs[(size_t)tid * ld + k];
                if (tid == CTA_THREADS - 1) {
                    const int t = s.colidx[pvt]; s.colidx[pvt] = s.colidx[k]; s.colidx[k] = t;
                    s.vn1[pvt] = nx1; s.vn2[pvt] = nx2;   // used by the owner of pvt only
                }
            }
            if (tid == CTA_THREADS - 2) s.tau[k] = tau;
            __syncthreads();
            // ---- P3: the pass over the trailing matrix, own positions only; results into every CTA's F.  Gram dots.
            mc_gemv<T, 4>(A, lda, own, s.colidx, vj, k, tau, fs_all, (size_t)j * ld, vec);
            for (int i = warp; i < j; i += CTA_WARPS) {
                const T* vi = s.Vs + (size_t)i * ld;
                T acc = mk<T>(0.0);
                for (int r = k + lane; r < n; r += 32) fmacc(acc, conj_(vi[r]), vj[r]);
                acc = warp_sum(acc);
                if (lane == 0) { s.Gp[i * NB + j] = acc; s.auxv[i] = -(tau * acc); }
            }
            if (tid == 0) *s.nflag = 0;
            cl.sync();   // B2: every F(c, j) has landed in every CTA; every CTA is done reading the pivot column
            // the reflector (rows > k) and beta go to the pivot's matrix column, by the owner of position k
            if (own.mine(k)) {
                T* kcol = A + (long long)s.colidx[k] * lda;
                for (int r = k + tid; r < n; r += CTA_THREADS) kcol[r] = (r == k) ? mk<T>(beta) : vj[r];
            }
            // ---- P4: finish F(:, j) everywhere; own positions: pivot row k, down-date of the partial norms
            double bv = -1.0; int bi = 0x7fffffff;
            for (int c = k + 1 + tid; c < n; c += CTA_THREADS) {
                T f = s.Fs[(size_t)j * ld + c];
                for (int i = 0; i < j; ++i) fmacc(f, s.Fs[(size_t)i * ld + c], s.auxv[i]);
                s.Fs[(size_t)j * ld + c] = f;
                if (!own.mine(c)) continue;
                T* ak = A + (long long)s.colidx[c] * lda + k;
                T a = *ak;
                for (int i = 0; i < j; ++i) fms(a, s.Vs[(size_t)i * ld + k], conj_(s.Fs[(size_t)i * ld + c]));
                a = a - conj_(f);  // i = j term, V(k,j) = 1
                *ak = a;
                double v1 = s.vn1[c];
                if (v1 != 0.0) {
                    double temp = abs_(a) / v1;
                    temp = fmax(0.0, (1.0 + temp) * (1.0 - temp));
                    double q = v1 / s.vn2[c];
                    double temp2 = temp * q * q;
                    if (temp2 <= TOL3Z_MC) {
                        int slot = atomicAdd(s.nflag, 1);
                        s.flag[slot] = c;
                    } else {
                        v1 *= sqrt(temp);
                        s.vn1[c] = v1;
                    }
                }
                if (v1 > bv) { bv = v1; bi = c; }
            }
            __syncthreads();
            const int nflag = *s.nflag;
            if (nflag > 0) {
                // exact recomputation of the flagged (own) columns' norms below row k, panel applied on the fly
                for (int f = warp; f < nflag; f += CTA_WARPS) {
                    const int c = s.flag[f];
                    const T* col = A + (long long)s.colidx[c] * lda;
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
                const int o_lo = own.first_ge(k + 1);
                for (int o = o_lo + tid; o < n_own; o += CTA_THREADS) {
                    const int c = own.pos(o);
                    if (c < n) { double v = s.vn1[c]; if (v > bv) { bv = v; bi = c; } }
                }
            }
            const int lb = block_argmax(bv, bi, s.redv, s.redi);
            const double lv = (lb < n) ? s.vn1[lb] : -1.0;
            __threadfence();   // pivot-row / reflector stores are read by other CTAs after the barrier
            pvt = exchange(lb < n ? lv : -1.0, lb, k + 1, k & 1);
            if (pvt >= n) pvt = min(k + 1, n - 1);
        }
        // ---- panel done: T factor (one warp of CTA 0) || rank-jb trailing update of the own positions
        const int kend = p0 + jb;
        {
            const int jb4 = (jb + 3) & ~3;
            for (int i = jb * ld + tid; i < jb4 * ld; i += CTA_THREADS) { s.Vs[i] = mk<T>(0.0); s.Fs[i] = mk<T>(0.0); }
            __syncthreads();
        }
        if (warp == CTA_WARPS - 1 && lane < jb) {
            const int l = lane;
            T* Trow = s.Ts + l * NB;
            Trow[l] = s.tau[p0 + l];
            for (int i = l + 1; i < jb; ++i) {
                T acc = mk<T>(0.0);
                for (int q = l; q < i; ++q) fmacc(acc, Trow[q], s.Gp[q * NB + i]);
                Trow[i] = -(s.tau[p0 + i] * acc);
            }
            if (rank == 0) {
                T* Tg = Tw + (size_t)(p0 / NB) * NB * NB;
                for (int i = 0; i < jb; ++i) Tg[l * NB + i] = (i >= l) ? Trow[i] : mk<T>(0.0);
            }
        }
        mc_rank_update<T, true>(A, lda, kend, n, kend, own, s.colidx, s.Vs, s.Fs, ld, jb, n - 1);
        __threadfence();
        cl.sync();   // B4: the next pivot column may belong to another CTA
    }

    // ---------------------------------------------------------------- d = |diag R'|, R = D^-1 R' P^T
    double* dall = s.vn2;   // the norms are dead: d of every position, in every CTA
    for (int r = tid; r < n; r += CTA_THREADS) dall[r] = abs_(A[(long long)s.colidx[r] * lda + r]);
    __syncthreads();
    {
        double* d = p.d + (long long)b * p.stride_d;
        T* R = p.R + (long long)b * p.strideR;
        if (rank == 0)
            for (int r = tid; r < n; r += CTA_THREADS) d[r] = dall[r];
        // position c holds matrix column colidx[c]: R[:, colidx[c]] = R'(0..c, c) / d, zero below -- own positions
        for (int o = warp; o < n_own; o += CTA_WARPS) {
            const int c = own.pos(o);
            if (c >= n) continue;
            const T* src = A + (long long)s.colidx[c] * lda;
            T* dst = R + (long long)s.colidx[c] * p.ldr;
            for (int r = lane; r < n; r += 32) dst[r] = (r <= c) ? src[r] / dall[r] : mk<T>(0.0);
        }
        if (p.jpvt_out && rank == 0)
            for (int c = tid; c < n; c += CTA_THREADS) p.jpvt_out[(long long)b * n + c] = s.colidx[c] + 1;
    }
    // the reflectors leave L (matrix column colidx[k], rows > k) for the scratch copy: own COLUMNS
    for (int o = warp; o < n_own; o += CTA_WARPS) {
        const int c = own.pos(o);
        if (c >= n) continue;
        const T* src = A + (long long)c * lda;
        T* dst = Qw + (long long)c * n;
        for (int r = lane; r < n; r += 32) dst[r] = src[r];
    }
    __threadfence();
    cl.sync();   // B5: all of Qw is written, nobody reads the factorisation in L any more

    // ---------------------------------------------------------------- Q = H_1 ... H_n (backward, blocked), own COLUMNS
    const int npan = (n + NB - 1) / NB;
    for (int pi = npan - 1; pi >= 0; --pi) {
        const int p0 = pi * NB, jb = min(NB, n - p0), kend = p0 + jb;
        // a. panel reflectors -> shared (unit diagonal, zeros above), T -> shared
        for (int i = 0; i < NB; ++i) {
            const int c = p0 + i;
            const T* vcol = Qw + (long long)s.colidx[c < n ? c : 0] * n;
            for (int r = tid; r < n; r += CTA_THREADS) {
                T v = mk<T>(0.0);
                if (i < jb) {
                    if (r > c) v = vcol[r];
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
        // b. initial content of the region this panel touches: [I 0; 0 Qsub] and zeros above -- own columns
        for (int i = 0; i < jb; ++i) {
            const int c = p0 + i;
            if (!own.mine(c)) continue;
            for (int r = tid; r < n; r += CTA_THREADS) A[(long long)c * lda + r] = mk<T>(r == c ? 1.0 : 0.0);
        }
        {
            const int o_lo = own.first_ge(kend);
            for (int o = o_lo + warp; o < n_own; o += CTA_WARPS) {
                const int c = own.pos(o);
                if (c >= n) continue;
                for (int r = p0 + lane; r < kend; r += 32) A[(long long)c * lda + r] = mk<T>(0.0);
            }
        }
        __syncthreads();
        // c. W = V^H C
        mc_panel_dot<T>(s.Fs, s.Vs, ld, jb, A, lda, p0, n, p0, own, n - 1);
        __syncthreads();
        // d. W <- T W   (upper-triangular T, in place top-down)
        {
            const int o_lo = own.first_ge(p0);
            for (int o = o_lo + tid; o < n_own; o += CTA_THREADS) {
                const int c = own.pos(o);
                if (c >= n) continue;
                for (int l = 0; l < jb; ++l) {
                    T acc = mk<T>(0.0);
                    for (int q = l; q < jb; ++q) fmacc(acc, s.Ts[l * NB + q], s.Fs[(size_t)q * ld + c]);
                    s.Fs[(size_t)l * ld + c] = acc;
                }
                for (int l = jb; l < ((jb + 3) & ~3); ++l) s.Fs[(size_t)l * ld + c] = mk<T>(0.0);
            }
        }
        __syncthreads();
        // e. C -= V W
        mc_rank_update<T, false>(A, lda, p0, n, p0, own, nullptr, s.Vs, s.Fs, ld, jb, n - 1);
        __syncthreads();
    }
    cl.sync();   // no CTA leaves while a peer may still address its shared memory
}

template <typename T> int mc_pick_nb(int n) {
    const size_t limit = 227 * 1024;
    if (mc_smem_bytes<T>(n, 32) <= limit) return 32;
    if (mc_smem_bytes<T>(n, 16) <= limit) return 16;
    if (mc_smem_bytes<T>(n, 8) <= limit) return 8;
    return 0;
}

template <typename T, int NB> cudaError_t launch_mc(const LdrArgs<T>& p, int cs, cudaStream_t stream) {
    auto kern = ldr_mc_kernel<T, NB>;
    const size_t smem = mc_smem_bytes<T>(p.n, NB);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)p.batch * cs, 1, 1);
    cfg.blockDim = dim3(CTA_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, p);
    ++g_launch_count;
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace

// the T factors must use the SAME panel width as the streaming kernel's workspace sizing (ldr_twork_elems): the shared
// memory of this kernel is a few hundred bytes larger, so its NB can only be equal or smaller
template <typename T> bool ldr_mc_supported(int n) {
    const int nb = mc_pick_nb<T>(n);
    return nb > 0 && (size_t)((n + nb - 1) / nb) * nb * nb <= ldr_twork_elems<T>(n);
}

template <typename T> cudaError_t ldr_mc_batched(const LdrArgs<T>& p, int cs, cudaStream_t stream) {
    if (p.batch <= 0 || p.n <= 0) return cudaSuccess;
    if (cs < 2 || cs > MC_MAXCS || (cs & (cs - 1)) || p.Qwork == nullptr || p.strideQ < (long long)p.n * p.n) return cudaErrorInvalidValue;
    switch (mc_pick_nb<T>(p.n)) {
        case 32: return launch_mc<T, 32>(p, cs, stream);
        case 16: return launch_mc<T, 16>(p, cs, stream);
        case 8: return launch_mc<T, 8>(p, cs, stream);
        default: return cudaErrorInvalidValue;
    }
}

template bool ldr_mc_supported<double>(int);
template bool ldr_mc_supported<cplx>(int);
template cudaError_t ldr_mc_batched<double>(const LdrArgs<double>&, int, cudaStream_t);
template cudaError_t ldr_mc_batched<cplx>(const LdrArgs<cplx>&, int, cudaStream_t);

}  // namespace sla
