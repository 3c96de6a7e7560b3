// Batched LDR factorisation, on-chip version: one thread-block CLUSTER owns one matrix.
//
// Replaces `ldr!(F, ws)` (src/LDR.jl:160-188: geqp3! + triu/abs/ldiv!/mul_P! + orgqr!) like sla_ldr.cu,
// but keeps the whole matrix in the shared memory of a cluster of CS CTAs (column-cyclic: CTA q owns
// the columns c = q (mod CS)), so that the per-column pass over the trailing matrix -- the BLAS-2 half
// of column-pivoted QR, which bounds the factorisation -- runs at shared-memory speed instead of L2
// round trips, and so that ONE cluster barrier per column is the only global synchronisation.
//
// Algorithm (own formulation; LAPACK ?laqp2 / ?larfg / ?ung2r mathematics, greedy pivoting on
// down-dated partial norms with LAPACK's cancellation safeguard):
//   * speculative pivots: after its pass every CTA takes its best local column, builds the Householder
//     reflector that column WOULD generate and writes it (plus norm, index, tau, beta) into every peer's
//     shared memory (DSMEM).  After the barrier all CTAs pick the same winner from the CS candidates;
//     nobody ever waits for an "owner" to finish serial work.
//   * no physical column swaps: a column's position in A*P is recorded when it wins, and since
//     R = D^-1 R' P^T wants R' back in the ORIGINAL column order each CTA writes its own columns of R
//     directly -- the permutation never materialises.
//   * lazy reflector application: the pass for column k applies the pending H_{k-1} to every remaining
//     local column AND accumulates v_k^H a for the new one (one read + one write per element).
//   * a column whose norm down-date is unreliable gets its norm recomputed exactly, immediately, by the
//     warp that owns it (LAPACK would cut the panel short instead).
//   * reflectors leave the SM through the TMA (cp.async.bulk shared->global), so the store latency never
//     sits in front of the cluster barrier's release.
//   * Q = H_0 ... H_{n-1} is then formed in place of the local columns: every CTA applies the reflectors
//     (streamed back through a cp.async ring) to its own columns of the identity, a warp keeping a
//     group of columns in REGISTERS for a whole chunk of reflectors.  No cluster traffic in this phase.
#include <cooperative_groups.h>

#include "sla_cta.cuh"
#include "sla_kernels.h"

namespace cg = cooperative_groups;

namespace sla {

namespace {

constexpr double TOL3Z_C = 1.4901161193847656e-08;  // sqrt(eps), LAPACK ?laqps / ?laqp2
constexpr int QKC = 8;                               // reflectors per chunk in the Q phase

template <typename T> struct CandMeta {   // what a CTA publishes about its candidate
    double val;      // down-dated norm (-1: no active column left)
    double beta;
    T tau;
    int idx;         // global column index
    int pad;
};

template <typename T> __host__ __device__ inline int cl_ldc(int n) { return sizeof(T) == 8 ? ((n + 1) & ~1) : n; }
template <typename T> __host__ __device__ inline int cl_vlen(int n) { return cl_ldc<T>(n) + 8; }

template <typename T> struct ClLayout {
    size_t offA, offCand, offVloc, offW, offTau, offMeta, offVn1, offVn2, offD, offRedv, offPos, offRedi, offMisc, total;
    __host__ __device__ ClLayout(int n, int cs) {
        const size_t nloc = (n + cs - 1) / cs, ldc = cl_ldc<T>(n), vlen = cl_vlen<T>(n);
        size_t o = 0;
        offA = o; o += nloc * ldc * sizeof(T);
        // candidate reflectors [2][cs][vlen]; the Q-phase ring [2][QKC][ldc] aliases this region
        size_t cand = 2 * (size_t)cs * vlen * sizeof(T), ring = 2 * (size_t)QKC * ldc * sizeof(T);
        offCand = o; o += (cand > ring ? cand : ring);
        offVloc = o; o += 2 * vlen * sizeof(T);
        offW = o; o += 2 * nloc * sizeof(T);
        offTau = o; o += (size_t)n * sizeof(T);
        offMeta = o; o += 2 * 8 * sizeof(CandMeta<T>);
        o = (o + 15) & ~(size_t)15;
        offVn1 = o; o += nloc * sizeof(double);
        offVn2 = o; o += nloc * sizeof(double);
        offD = o; o += (size_t)n * sizeof(double);
        offRedv = o; o += 64 * sizeof(double);
        offPos = o; o += nloc * sizeof(int);
        offRedi = o; o += 64 * sizeof(int);
        offMisc = o; o += 16 * sizeof(double);
        total = o + 64;
    }
};

template <int CS> __device__ __forceinline__ void cluster_barrier() {
    if constexpr (CS == 1) __syncthreads();
    else cg::this_cluster().sync();
}
template <int CS, typename P> __device__ __forceinline__ P* peer(P* local, int rank) {
    if constexpr (CS == 1) return local;
    else return cg::this_cluster().map_shared_rank(local, rank);
}

__device__ __forceinline__ double shfl_xor_T(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
__device__ __forceinline__ cplx shfl_xor_T(cplx v, int o) {
    return cplx{__shfl_xor_sync(0xffffffffu, v.re, o), __shfl_xor_sync(0xffffffffu, v.im, o)};
}
// NaN norms must still be selectable as pivots (otherwise the column would never retire)
__device__ __forceinline__ double nan_as_inf(double v) { return (v == v) ? v : INFINITY; }

template <typename T, int G> __device__ __forceinline__ void warp_sum_group(T (&a)[G]) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int u = 0; u < G; ++u) a[u] = a[u] + shfl_xor_T(a[u], o);
    }
}
template <typename T, int G> __device__ __forceinline__ T pick(const T (&a)[G], int u) {
    T r = a[0];
#pragma unroll
    for (int i = 1; i < G; ++i) if (u == i) r = a[i];
    return r;
}

// TMA bulk store shared -> global (bytes multiple of 16, both 16-byte aligned); one thread issues
__device__ __forceinline__ void bulk_store_s2g(void* gdst, const void* ssrc, unsigned bytes) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

}  // namespace

// CPW: local columns a warp processes together in the QR pass.  RPLQ: rows per lane of the Q-phase
// register tile (n <= 32*RPLQ);  CPWQ columns per warp there.
template <typename T, int CS, int CPW, int RPLQ, int CPWQ>
__global__ void __launch_bounds__(CTA_THREADS, 1) ldr_cluster_kernel(LdrArgs<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n;
    const int nloc = (n + CS - 1) / CS;
    const int ldc = cl_ldc<T>(n), vlen = cl_vlen<T>(n);
    const ClLayout<T> lay(n, CS);
    T* const A = reinterpret_cast<T*>(smem_raw + lay.offA);
    T* const vcand = reinterpret_cast<T*>(smem_raw + lay.offCand);     // [2][CS][vlen]
    T* const vloc = reinterpret_cast<T*>(smem_raw + lay.offVloc);      // [2][vlen]  private copies of winners
    T* const wbuf = reinterpret_cast<T*>(smem_raw + lay.offW);         // [2][nloc]
    T* const taus = reinterpret_cast<T*>(smem_raw + lay.offTau);
    CandMeta<T>* const meta = reinterpret_cast<CandMeta<T>*>(smem_raw + lay.offMeta);   // [2][8]
    double* const vn1 = reinterpret_cast<double*>(smem_raw + lay.offVn1);
    double* const vn2 = reinterpret_cast<double*>(smem_raw + lay.offVn2);
    double* const dvec = reinterpret_cast<double*>(smem_raw + lay.offD);
    double* const redv = reinterpret_cast<double*>(smem_raw + lay.offRedv);
    int* const pos = reinterpret_cast<int*>(smem_raw + lay.offPos);
    int* const redi = reinterpret_cast<int*>(smem_raw + lay.offRedi);
    double* const misc = reinterpret_cast<double*>(smem_raw + lay.offMisc);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int rank = 0;
    if constexpr (CS > 1) rank = (int)cg::this_cluster().block_rank();
    const int b = blockIdx.x / CS;
    constexpr bool CP = is_cplx<T>::value;
    constexpr int EPV = CP ? 1 : 2;   // elements per 16-byte vector

    T* Lg = p.L + (long long)b * p.strideL;
    const int ldl = p.ldl;
    const T* Ain = p.Ain ? p.Ain + (long long)b * p.strideAin : Lg;
    const int ldin = p.Ain ? p.ldain : ldl;
    const bool bulk_ok = ((((uintptr_t)Lg) & 15) == 0) && (((size_t)ldl * sizeof(T)) % 16 == 0) && (CP || (n % 2 == 0));

    // ------------------------------------------------------------------ load local columns, norms
    for (int lc = warp; lc < nloc; lc += CTA_WARPS) {
        const int c = lc * CS + rank;
        T* col = A + (size_t)lc * ldc;
        double acc = 0.0;
        if (c < n) {
            const T* src = Ain + (long long)c * ldin;
            for (int r = lane; r < ldc; r += 32) {
                T x = (r < n) ? src[r] : mk<T>(0.0);
                col[r] = x;
                acc += abs2_(x);
            }
        } else {
            for (int r = lane; r < ldc; r += 32) col[r] = mk<T>(0.0);
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            double nr = sqrt(acc);
            vn1[lc] = nr; vn2[lc] = nr;
            pos[lc] = (c < n) ? -1 : n;   // padding columns count as already chosen
            wbuf[lc] = mk<T>(0.0); wbuf[nloc + lc] = mk<T>(0.0);
        }
    }
    for (int r = tid; r < 2 * vlen; r += CTA_THREADS) vloc[r] = mk<T>(0.0);
    for (int r = tid; r < 2 * CS * vlen; r += CTA_THREADS) vcand[r] = mk<T>(0.0);
    __syncthreads();
    cluster_barrier<CS>();   // peers' candidate buffers are zeroed before anyone writes into them

    // CTA-wide best active column -> (cand, cand_v); all threads get the result (2 barriers)
    auto cta_candidate = [&](double bv, int bi, int& cand, double& cand_v) {
        if (lane == 0) { redv[warp] = bv; redi[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            double v = (lane < CTA_WARPS) ? redv[lane] : -2.0;
            int i = (lane < CTA_WARPS) ? redi[lane] : 0x7fffffff;
            warp_argmax(v, i);
            if (lane == 0) { redv[32] = v; redi[32] = i; }
        }
        __syncthreads();
        cand_v = redv[32]; cand = redi[32];
    };

    // Build the reflector the local candidate column would generate for position `kpos` (rows kpos..n-1),
    // with the pending reflector (vpend, wpend) applied on the fly, and publish it to every CTA.
    auto publish_candidate = [&](int kpos, int cand, double cand_v, const T* vpend, const T* wpend, bool pending) {
        const int par = kpos & 1;
        if (cand >= n) {   // no active column left in this CTA
            if (tid < CS) peer<CS>(meta, tid)[par * 8 + rank].val = -1.0;
            return;
        }
        const int lcs = cand / CS;
        const T* col = A + (size_t)lcs * ldc;
        const T wp = pending ? wpend[lcs] : mk<T>(0.0);
        // rows [rz, kpos) are written as zeros: the pass for column kpos starts reading v at the (aligned)
        // first row of the PENDING reflector, kpos-1, and the slot still holds an older vector there
        const int rz = (kpos >= 2 ? kpos - 2 : 0) & ~(EPV - 1);
        double ss = 0.0;
        T xmine[(32 * RPLQ + CTA_THREADS - 1) / CTA_THREADS];
#pragma unroll
        for (int t = 0; t < (32 * RPLQ + CTA_THREADS - 1) / CTA_THREADS; ++t) {
            const int r = rz + tid + t * CTA_THREADS;
            T x = mk<T>(0.0);
            if (r < n && r >= kpos) {
                x = col[r];
                if (pending) fms(x, vpend[r], wp);
                if (r > kpos) ss += abs2_(x);
                else redv[40] = real_(x), redv[41] = imag_(x);   // alpha
            }
            xmine[t] = x;
        }
        const double xnorm2 = block_sum(ss, redv);
        const T alpha = mk<T>(redv[40], redv[41]);
        T tau, scal; double beta;
        if (xnorm2 == 0.0 && imag_(alpha) == 0.0) {
            tau = mk<T>(0.0); scal = mk<T>(0.0); beta = real_(alpha);
        } else {
            beta = -copysign(sqrt(abs2_(alpha) + xnorm2), real_(alpha));
            tau = mk<T>((beta - real_(alpha)) / beta, -imag_(alpha) / beta);
            scal = mk<T>(1.0) / (alpha - mk<T>(beta));
        }
        T* slot = vcand + ((size_t)par * CS + rank) * vlen;
#pragma unroll
        for (int t = 0; t < (32 * RPLQ + CTA_THREADS - 1) / CTA_THREADS; ++t) {
            const int r = rz + tid + t * CTA_THREADS;
            if (r < n) {
                T v = mk<T>(0.0);
                if (r > kpos) v = xmine[t] * scal;
                else if (r == kpos) v = mk<T>(1.0);
#pragma unroll
                for (int q = 0; q < CS; ++q) peer<CS>(slot, q)[r] = v;
            }
        }
        if (tid < CS) {
            CandMeta<T> m;
            m.val = cand_v; m.beta = beta; m.tau = tau; m.idx = cand; m.pad = 0;
            peer<CS>(meta, tid)[par * 8 + rank] = m;
        }
    };

    // ------------------------------------------------------------------ prologue: candidates for position 0
    int cand; double cand_v;
    {
        double bv = -1.0; int bi = 0x7fffffff;
        for (int lc = tid; lc < nloc; lc += CTA_THREADS)
            if (pos[lc] < 0) { double v = nan_as_inf(vn1[lc]); int c = lc * CS + rank; if (v > bv || (v == bv && c < bi)) { bv = v; bi = c; } }
        warp_argmax(bv, bi);
        cta_candidate(bv, bi, cand, cand_v);
        publish_candidate(0, cand, cand_v, vloc, wbuf, false);
    }

    // ------------------------------------------------------------------ pivoted Householder QR
    for (int k = 0; k < n; ++k) {
        const int par = k & 1;
        if (tid == 0) bulk_wait_read();   // the TMA has finished reading the candidate slot it was given
        cluster_barrier<CS>();            // the ONLY cluster-wide synchronisation per column
        // ---- winner of the CS candidates (identical decision in every CTA)
        int win = 0;
        {
            double gv = meta[par * 8].val; int gi = meta[par * 8].idx;
#pragma unroll
            for (int q = 1; q < CS; ++q) {
                const double ov = meta[par * 8 + q].val; const int oi = meta[par * 8 + q].idx;
                if (ov > gv || (ov == gv && ov >= 0.0 && oi < gi)) { gv = ov; gi = oi; win = q; }
            }
        }
        const CandMeta<T> wm = meta[par * 8 + win];
        const int pc = wm.idx, plc = pc / CS;
        const T tauk = wm.tau;
        const T* vcur = vcand + ((size_t)par * CS + win) * vlen;
        T* vsave = vloc + (size_t)par * vlen;              // private copy: v_k is "pending" during the next pass
        const T* vprev = vloc + (size_t)(par ^ 1) * vlen;
        T* wcur = wbuf + (size_t)par * nloc;
        const T* wprev = wbuf + (size_t)(par ^ 1) * nloc;
        if (tid == 0) { taus[k] = tauk; dvec[k] = fabs(wm.beta); }
        if (rank == win && tid == 32) {
            // finalise the pivot column: row k-1 gets the pending reflector's contribution, row k = beta
            T* col = A + (size_t)plc * ldc;
            if (k > 0) col[k - 1] = col[k - 1] - wprev[plc];
            col[k] = mk<T>(wm.beta);
            pos[plc] = k;
        }
        // reflector k -> global (rows > k of column k of L), through the TMA, by CTA k mod CS
        if (rank == (k % CS) && k + 1 < n) {
            const int r0 = (k + 1) & ~(EPV - 1);
            if (bulk_ok) {
                if (tid == 0) bulk_store_s2g(Lg + (long long)k * ldl + r0, vcur + r0, (unsigned)((n - r0) * sizeof(T)));
            } else {
                for (int r = k + 1 + tid; r < n; r += CTA_THREADS) Lg[(long long)k * ldl + r] = vcur[r];
            }
        }
        if (k == n - 1) break;
        __syncthreads();   // pos[] update visible before the pass
        // ---- fused pass: apply pending H_{k-1}, accumulate v_k^H a, down-date norms, local candidate
        const T tauc = conj_(tauk);
        const int rb = (k > 0) ? ((k - 1) & ~(EPV - 1)) : 0;
        double bv = -1.0; int bi = 0x7fffffff;
        for (int g0 = warp; g0 < nloc; g0 += CTA_WARPS * CPW) {
            // this warp's columns of the group: lc = g0 + u*CTA_WARPS
            T wp[CPW], acc[CPW];
            bool act[CPW];
            bool any = false;
#pragma unroll
            for (int u = 0; u < CPW; ++u) {
                const int lc = g0 + u * CTA_WARPS;
                act[u] = (lc < nloc) && (pos[lc] < 0);
                wp[u] = act[u] ? wprev[lc] : mk<T>(0.0);
                acc[u] = mk<T>(0.0);
                any |= act[u];
            }
            if (g0 == warp) {   // one warp-wide copy of v_k into the private buffer (for the next pass)
                for (int r = rb + lane; r < n; r += 32) if (warp == (r >> 5) % CTA_WARPS) vsave[r] = vcur[r];
            }
            if (!any) continue;
            if constexpr (!CP) {
                for (int r = rb + 2 * lane; r < n; r += 64) {
                    const double2 vp = *reinterpret_cast<const double2*>(vprev + r);
                    const double2 vc = *reinterpret_cast<const double2*>(vcur + r);
#pragma unroll
                    for (int u = 0; u < CPW; ++u) {
                        if (act[u]) {
                            double2* ptr = reinterpret_cast<double2*>(A + (size_t)(g0 + u * CTA_WARPS) * ldc + r);
                            double2 a = *ptr;
                            a.x = fma(-vp.x, wp[u], a.x);
                            a.y = fma(-vp.y, wp[u], a.y);
                            *ptr = a;
                            acc[u] = fma(a.x, vc.x, acc[u]);
                            acc[u] = fma(a.y, vc.y, acc[u]);
                        }
                    }
                }
            } else {
                for (int r = rb + lane; r < n; r += 32) {
                    const T vp = vprev[r], vc = vcur[r];
#pragma unroll
                    for (int u = 0; u < CPW; ++u) {
                        if (act[u]) {
                            T* ptr = A + (size_t)(g0 + u * CTA_WARPS) * ldc + r;
                            T a = *ptr;
                            fms(a, vp, wp[u]);
                            *ptr = a;
                            fmacc(acc[u], conj_(vc), a);
                        }
                    }
                }
            }
            warp_sum_group<T, CPW>(acc);
            __syncwarp();
            // lane u (< CPW) finishes column u: w, pivot-row entry, norm down-date
            bool flagged = false;
            double v1 = -1.0;
            int myc = 0x7fffffff;
            T wnew = mk<T>(0.0);
            const int mylc = g0 + lane * CTA_WARPS;
            const bool mine = (lane < CPW) && (mylc < nloc) && (pos[mylc] < 0);
            if (mine) {
                wnew = tauc * pick<T, CPW>(acc, lane);
                wcur[mylc] = wnew;
                myc = mylc * CS + rank;
                v1 = vn1[mylc];
                if (v1 != 0.0) {
                    const T ak = A[(size_t)mylc * ldc + k] - wnew;   // a'(k,c): v_k(k) = 1
                    double temp = abs_(ak) / v1;
                    temp = fmax(0.0, (1.0 + temp) * (1.0 - temp));
                    const double q = v1 / vn2[mylc];
                    if (temp * q * q <= TOL3Z_C) flagged = true;
                    else { v1 *= sqrt(temp); vn1[mylc] = v1; }
                }
            }
            unsigned fl = __ballot_sync(0xffffffffu, flagged);
            while (fl) {   // exact norm of a flagged column below row k (new reflector applied on the fly)
                const int u = __ffs(fl) - 1;
                fl &= fl - 1;
                const int lc = g0 + u * CTA_WARPS;
                const T wsel = mk<T>(__shfl_sync(0xffffffffu, real_(wnew), u), __shfl_sync(0xffffffffu, imag_(wnew), u));
                const T* col = A + (size_t)lc * ldc;
                double ss = 0.0;
                for (int r = k + 1 + lane; r < n; r += 32) {
                    T x = col[r];
                    fms(x, vcur[r], wsel);
                    ss += abs2_(x);
                }
                ss = warp_sum(ss);
                if (lane == u) { v1 = sqrt(ss); vn1[lc] = v1; vn2[lc] = v1; }
            }
            if (mine) {
                v1 = nan_as_inf(v1);
                if (v1 > bv || (v1 == bv && myc < bi)) { bv = v1; bi = myc; }
            }
        }
        warp_argmax(bv, bi);
        cta_candidate(bv, bi, cand, cand_v);
        // ---- speculative reflector for position k+1 from the local candidate (pending = H_k)
        publish_candidate(k + 1, cand, cand_v, vcur, wcur, true);
    }
    if (tid == 0) bulk_wait_all();   // all reflectors have landed in global memory
    __syncthreads();

    // ------------------------------------------------------------------ d, R (original column order)
    {
        T* Rg = p.R + (long long)b * p.strideR;
        for (int lc = warp; lc < nloc; lc += CTA_WARPS) {
            const int c = lc * CS + rank;
            if (c >= n) continue;
            const int pk = pos[lc];
            const T* col = A + (size_t)lc * ldc;
            T* dst = Rg + (long long)c * p.ldr;
            for (int r = lane; r < n; r += 32) dst[r] = (r <= pk) ? col[r] / dvec[r] : mk<T>(0.0);
        }
        if (rank == 0) {
            double* d = p.d + (long long)b * p.stride_d;
            for (int r = tid; r < n; r += CTA_THREADS) d[r] = dvec[r];
        }
    }
    __syncthreads();

    // ------------------------------------------------------------------ Q = H_0 ... H_{n-1} applied to I
    for (int lc = warp; lc < nloc; lc += CTA_WARPS) {
        const int c = lc * CS + rank;
        T* col = A + (size_t)lc * ldc;
        for (int r = lane; r < ldc; r += 32) col[r] = mk<T>((r == c && c < n) ? 1.0 : 0.0);
    }
    __threadfence();
    cluster_barrier<CS>();   // every CTA's reflectors are in global memory and visible

    {
        T* const ring = vcand;   // [2][QKC][ldc], aliases the candidate buffers (no longer needed)
        const int nchunk = (n + QKC - 1) / QKC;
        const bool vecq = (((size_t)ldl * sizeof(T)) % 16 == 0) && ((((uintptr_t)Lg) & 15) == 0) && (CP || (ldc == n));
        auto load_chunk = [&](int ch, int buf) {
            // raw copy of columns [ch*QKC, ch*QKC+QKC) of Lg; entries on/above the diagonal are masked at use
            T* dstb = ring + (size_t)buf * QKC * ldc;
            const int per_col = ldc / EPV;
            for (int e = tid; e < QKC * per_col; e += CTA_THREADS) {
                const int i = e / per_col, r = (e % per_col) * EPV;
                const int kk = ch * QKC + i;
                T* dst = dstb + (size_t)i * ldc + r;
                if (kk < n && vecq) {
                    cp_async16(dst, Lg + (long long)kk * ldl + r, 16);
                } else if (kk < n) {
                    for (int t = 0; t < EPV; ++t) dst[t] = (r + t < n) ? Lg[(long long)kk * ldl + r + t] : mk<T>(0.0);
                }
            }
            cp_async_commit();
        };
        load_chunk(nchunk - 1, (nchunk - 1) & 1);
        for (int ch = nchunk - 1; ch >= 0; --ch) {
            if (ch > 0) load_chunk(ch - 1, (ch - 1) & 1);
            if (ch > 0) cp_async_wait<1>(); else cp_async_wait<0>();
            __syncthreads();
            const T* vb = ring + (size_t)(ch & 1) * QKC * ldc;
            const int k0 = ch * QKC;
            const int kcnt = min(QKC, n - k0);
            const int t0 = k0 >> 5;   // row tiles below this are untouched by the chunk (v is zero there)
            // warp's columns: lc = warp + 16*(g*CPWQ + u): spread over the index range for load balance
            for (int g = 0; (g * CPWQ) * CTA_WARPS + warp < nloc; ++g) {
                int lcs[CPWQ]; bool on[CPWQ]; bool any = false;
#pragma unroll
                for (int u = 0; u < CPWQ; ++u) {
                    lcs[u] = warp + CTA_WARPS * (g * CPWQ + u);
                    // H_k touches column c only if c >= k: skip columns left of the whole chunk
                    on[u] = (lcs[u] < nloc) && (lcs[u] * CS + rank < n) && (lcs[u] * CS + rank >= k0);
                    any |= on[u];
                }
                if (!any) continue;
                T q[CPWQ][RPLQ];
#pragma unroll
                for (int u = 0; u < CPWQ; ++u)
#pragma unroll
                    for (int t = 0; t < RPLQ; ++t) {
                        const int r = lane + 32 * t;
                        q[u][t] = (on[u] && t >= t0 && r < n) ? A[(size_t)lcs[u] * ldc + r] : mk<T>(0.0);
                    }
                for (int i = kcnt - 1; i >= 0; --i) {
                    const int k = k0 + i;
                    const T* vk = vb + (size_t)i * ldc;
                    const T tk = taus[k];
                    T vv[RPLQ];
#pragma unroll
                    for (int t = 0; t < RPLQ; ++t) {
                        const int r = lane + 32 * t;
                        vv[t] = (t >= t0 && r < n && r > k) ? vk[r] : mk<T>(r == k ? 1.0 : 0.0);
                    }
                    T acc[CPWQ];
#pragma unroll
                    for (int u = 0; u < CPWQ; ++u) {
                        acc[u] = mk<T>(0.0);
#pragma unroll
                        for (int t = 0; t < RPLQ; ++t) fmacc(acc[u], conj_(vv[t]), q[u][t]);
                    }
                    warp_sum_group<T, CPWQ>(acc);
#pragma unroll
                    for (int u = 0; u < CPWQ; ++u) {
                        // columns with c < k are untouched by H_k (their dot product is exactly zero anyway)
                        const T wq = tk * acc[u];
#pragma unroll
                        for (int t = 0; t < RPLQ; ++t) fms(q[u][t], vv[t], wq);
                    }
                }
#pragma unroll
                for (int u = 0; u < CPWQ; ++u)
#pragma unroll
                    for (int t = 0; t < RPLQ; ++t) {
                        const int r = lane + 32 * t;
                        if (on[u] && t >= t0 && r < n) A[(size_t)lcs[u] * ldc + r] = q[u][t];
                    }
            }
            __syncthreads();
        }
    }
    cluster_barrier<CS>();   // every CTA is done reading reflectors from Lg before anyone overwrites them
    for (int lc = warp; lc < nloc; lc += CTA_WARPS) {
        const int c = lc * CS + rank;
        if (c >= n) continue;
        const T* col = A + (size_t)lc * ldc;
        T* dst = Lg + (long long)c * ldl;
        for (int r = lane; r < n; r += 32) dst[r] = col[r];
    }
}

// smallest cluster size whose CTAs can hold the matrix (0: does not fit / n too large for the register tile)
template <typename T> int ldr_cluster_size(int n) {
    const size_t limit = 227 * 1024;
    if (n > 512) return 0;
    for (int cs : {1, 2, 4, 8})
        if (ClLayout<T>(n, cs).total <= limit) return cs;
    return 0;
}

template <typename T, int CS, int RPLQ> static cudaError_t launch_cluster(const LdrArgs<T>& p, cudaStream_t stream) {
    constexpr int CPW = is_cplx<T>::value ? 2 : 4;
    constexpr int CPWQ0 = (is_cplx<T>::value ? 16 : 32) / RPLQ;   // register tile of <= 32 (16 complex) values per lane
    constexpr int CPWQ = (RPLQ >= 16 || CPWQ0 < 1) ? 1 : (CPWQ0 > 4 ? 4 : CPWQ0);
    auto kern = ldr_cluster_kernel<T, CS, CPW, RPLQ, CPWQ>;
    const size_t smem = ClLayout<T>(p.n, CS).total;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)p.batch * CS, 1, 1);
    cfg.blockDim = dim3(CTA_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, p);
    ++g_launch_count;
    return e;
}

template <typename T, int CS> static cudaError_t launch_cluster_rpl(const LdrArgs<T>& p, cudaStream_t stream) {
    if (p.n <= 64) return launch_cluster<T, CS, 2>(p, stream);
    if (p.n <= 128) return launch_cluster<T, CS, 4>(p, stream);
    if (p.n <= 256) return launch_cluster<T, CS, 8>(p, stream);
    return launch_cluster<T, CS, 16>(p, stream);
}

template <typename T> cudaError_t ldr_cluster_batched(const LdrArgs<T>& p, cudaStream_t stream) {
    if (p.batch <= 0 || p.n <= 0) return cudaSuccess;
    switch (ldr_cluster_size<T>(p.n)) {
        case 1: return launch_cluster_rpl<T, 1>(p, stream);
        case 2: return launch_cluster_rpl<T, 2>(p, stream);
        case 4: return launch_cluster_rpl<T, 4>(p, stream);
        case 8: return launch_cluster_rpl<T, 8>(p, stream);
        default: return cudaErrorNotSupported;
    }
}

template int ldr_cluster_size<double>(int);
template int ldr_cluster_size<cplx>(int);
template cudaError_t ldr_cluster_batched<double>(const LdrArgs<double>&, cudaStream_t);
template cudaError_t ldr_cluster_batched<cplx>(const LdrArgs<cplx>&, cudaStream_t);

}  // namespace sla
