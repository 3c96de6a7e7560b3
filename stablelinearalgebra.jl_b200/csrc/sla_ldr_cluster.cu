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
// first row the pass for column k touches (row k-1 of the pending reflector), as a multiple of 2
__device__ __forceinline__ int rb_k(int k) { return k > 0 ? ((k - 1) & ~1) : 0; }
// NaN norms must still be selectable as pivots (otherwise the column would never retire)
__device__ __forceinline__ double nan_as_inf(double v) { return (v == v) ? v : INFINITY; }

template <typename T, int G> __device__ __forceinline__ void warp_sum_group(T (&a)[G]) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int u = 0; u < G; ++u) a[u] = a[u] + shfl_xor_T(a[u], o);
    }
}
// Sum G (= 2 or 4) per-lane values over the warp with G-fold fewer shuffles: afterwards every lane of
// lane-group u (32/G consecutive lanes) holds the total of value u.
template <typename T, int G> __device__ __forceinline__ T warp_sum_transposed(const T (&a)[G], int lane) {
    static_assert(G == 2 || G == 4, "G must be 2 or 4");
    T kk;
    if constexpr (G == 4) {
        const bool hi16 = lane & 16;
        T s0 = hi16 ? a[0] : a[2], s1 = hi16 ? a[1] : a[3];
        T k0 = hi16 ? a[2] : a[0], k1 = hi16 ? a[3] : a[1];
        k0 = k0 + shfl_xor_T(s0, 16);
        k1 = k1 + shfl_xor_T(s1, 16);
        const bool hi8 = lane & 8;
        T s = hi8 ? k0 : k1;
        kk = hi8 ? k1 : k0;
        kk = kk + shfl_xor_T(s, 8);
    } else {
        const bool hi16 = lane & 16;
        T s = hi16 ? a[0] : a[1];
        kk = hi16 ? a[1] : a[0];
        kk = kk + shfl_xor_T(s, 16);
        kk = kk + shfl_xor_T(kk, 8);
    }
    kk = kk + shfl_xor_T(kk, 4);
    kk = kk + shfl_xor_T(kk, 2);
    kk = kk + shfl_xor_T(kk, 1);
    return kk;
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
        bool nz = false;
        if (c < n) {
            const T* src = Ain + (long long)c * ldin;
            for (int r = lane; r < ldc; r += 32) {
                T x = (r < n) ? src[r] : mk<T>(0.0);
                col[r] = x;
                acc += abs2_(x);
                nz |= !is_zero(x);
            }
        } else {
            for (int r = lane; r < ldc; r += 32) col[r] = mk<T>(0.0);
        }
        acc = warp_sum(acc);
        nz = __any_sync(0xffffffffu, nz);
        if (lane == 0) {
            if (c < n && ldr_norm2_out_of_range(acc, nz)) ldr_raise_range(p.info, b);
            vn1[lc] = acc; vn2[lc] = acc;   // SQUARED partial norms (no sqrt on the per-column critical path)
            pos[lc] = (c < n) ? -1 : n;   // padding columns count as already chosen
            wbuf[lc] = mk<T>(0.0); wbuf[nloc + lc] = mk<T>(0.0);
        }
    }
    for (int r = tid; r < 2 * vlen; r += CTA_THREADS) vloc[r] = mk<T>(0.0);
    for (int r = tid; r < 2 * CS * vlen; r += CTA_THREADS) vcand[r] = mk<T>(0.0);
    __syncthreads();
    cluster_barrier<CS>();   // peers' candidate buffers are zeroed before anyone writes into them

    // After the pass: per-warp candidates go to shared memory, ONE block barrier, then warp 0 alone picks
    // the CTA's candidate, builds the reflector that column would generate for position `kpos` (rows
    // kpos..n-1; the pending reflector (vpend, wpend) is applied on the fly, the stored column is left
    // untouched) and publishes it to every CTA.  The other warps go straight on to the cluster barrier.
    auto select_and_publish = [&](double bv, int bi, int kpos, const T* vpend, const T* wpend, bool pending) {
        if (lane == 0) { redv[warp] = bv; redi[warp] = bi; }
        __syncthreads();
        if (warp != 0) return;
        double cand_v = (lane < CTA_WARPS) ? redv[lane] : -2.0;
        int cand = (lane < CTA_WARPS) ? redi[lane] : 0x7fffffff;
        warp_argmax(cand_v, cand);
        const int par = kpos & 1;
        if (cand >= n) {   // no active column left in this CTA
            if (lane < CS) peer<CS>(meta, lane)[par * 8 + rank].val = -1.0;
            return;
        }
        const int lcs = cand / CS;
        const T* col = A + (size_t)lcs * ldc;
        const T wp = pending ? wpend[lcs] : mk<T>(0.0);
        // rows [rz, kpos) are written as zeros: the pass for column kpos starts reading v at the (aligned)
        // first row of the PENDING reflector, kpos-1, and the slot still holds an older vector there
        const int rz = (kpos >= 2 ? kpos - 2 : 0) & ~(EPV - 1);
        const int t_lo = rz >> 5;
        double ss = 0.0;
        T alpha = mk<T>(0.0);
        T x[RPLQ];
#pragma unroll
        for (int t = 0; t < RPLQ; ++t) {
            const int r = lane + 32 * t;
            x[t] = mk<T>(0.0);
            if (t >= t_lo && r < n && r >= kpos) {
                T xv = col[r];
                if (pending) fms(xv, vpend[r], wp);
                if (r > kpos) ss += abs2_(xv); else alpha = xv;
                x[t] = xv;
            }
        }
        const double xnorm2 = warp_sum(ss);
        alpha = warp_sum(alpha);   // exactly one lane holds a non-zero alpha
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
        for (int t = 0; t < RPLQ; ++t) {
            const int r = lane + 32 * t;
            if (t >= t_lo && r < n && r >= rz) {
                T v = mk<T>(0.0);
                if (r > kpos) v = x[t] * scal;
                else if (r == kpos) v = mk<T>(1.0);
#pragma unroll
                for (int q = 0; q < CS; ++q) peer<CS>(slot, q)[r] = v;
            }
        }
        if (lane < CS) {
            CandMeta<T> m;
            m.val = cand_v; m.beta = beta; m.tau = tau; m.idx = cand; m.pad = 0;
            peer<CS>(meta, lane)[par * 8 + rank] = m;
        }
    };

    // ------------------------------------------------------------------ prologue: candidates for position 0
    {
        double bv = -1.0; int bi = 0x7fffffff;
        for (int lc = tid; lc < nloc; lc += CTA_THREADS)
            if (pos[lc] < 0) { double v = nan_as_inf(vn1[lc]); int c = lc * CS + rank; if (v > bv || (v == bv && c < bi)) { bv = v; bi = c; } }
        warp_argmax(bv, bi);
        select_and_publish(bv, bi, 0, vloc, wbuf, false);
    }

    // ------------------------------------------------------------------ pivoted Householder QR
    // per-phase cycle counters of cluster 0 (diagnostics; compiled in only with -DSLA_LDR_PROFILE,
    // see profiles/ldr_phase_cycles.py)
#ifdef SLA_LDR_PROFILE
    long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bool prof = (p.prof != nullptr) && (blockIdx.x < CS);
    long long tq = prof ? clock64() : 0;
#define PROF_MARK(i) do { if (prof) { long long t_ = clock64(); pt[i] += t_ - tq; tq = t_; } } while (0)
#else
#define PROF_MARK(i) do { } while (0)
#endif
    for (int k = 0; k < n; ++k) {
        const int par = k & 1;
        if (tid == 0) bulk_wait_read();   // the TMA has finished reading the candidate slot it was given
        PROF_MARK(0);                      // tail of the previous iteration (publish for warp 0, nothing for others)
        cluster_barrier<CS>();            // the ONLY cluster-wide synchronisation per column
        PROF_MARK(1);                      // cluster barrier (arrive + wait)
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
        PROF_MARK(2);                      // winner pick, finalisation, TMA issue
        // private copy of v_k: it is the "pending" reflector during the next pass
        for (int r = (rb_k(k) & ~31) + tid; r < n; r += CTA_THREADS) vsave[r] = vcur[r];
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
                // pos[] of the column that just won is being written concurrently: exclude it explicitly
                act[u] = (lc < nloc) && (pos[lc] < 0) && !(rank == win && lc == plc);
                wp[u] = act[u] ? wprev[lc] : mk<T>(0.0);
                acc[u] = mk<T>(0.0);
                any |= act[u];
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
            const T mysum = warp_sum_transposed<T, CPW>(acc, lane);
            __syncwarp();
            // the first lane of lane-group u finishes column u: w, pivot-row entry, norm down-date (squared)
            constexpr int FLS = 32 / CPW;
            bool flagged = false;
            double v1 = -1.0;
            int myc = 0x7fffffff;
            T wnew = mk<T>(0.0);
            const int myu = lane / FLS;
            const int mylc = g0 + myu * CTA_WARPS;
            const bool mine = ((lane % FLS) == 0) && (mylc < nloc) && (pos[mylc] < 0) && !(rank == win && mylc == plc);
            if (mine) {
                wnew = tauc * mysum;
                wcur[mylc] = wnew;
                myc = mylc * CS + rank;
                v1 = vn1[mylc];
                if (v1 != 0.0) {
                    const T ak = A[(size_t)mylc * ldc + k] - wnew;   // a'(k,c): v_k(k) = 1
                    const double temp = fmax(0.0, 1.0 - abs2_(ak) / v1);
                    const double q2 = v1 / vn2[mylc];
                    if (temp * q2 <= TOL3Z_C) flagged = true;
                    else { v1 *= temp; vn1[mylc] = v1; }
                }
            }
            unsigned fl = __ballot_sync(0xffffffffu, flagged);
            while (fl) {   // exact norm of a flagged column below row k (new reflector applied on the fly)
                const int src = __ffs(fl) - 1;
                fl &= fl - 1;
                const int lc = g0 + (src / FLS) * CTA_WARPS;
                const T wsel = mk<T>(__shfl_sync(0xffffffffu, real_(wnew), src), __shfl_sync(0xffffffffu, imag_(wnew), src));
                const T* col = A + (size_t)lc * ldc;
                double ss = 0.0;
                for (int r = k + 1 + lane; r < n; r += 32) {
                    T x = col[r];
                    fms(x, vcur[r], wsel);
                    ss += abs2_(x);
                }
                ss = warp_sum(ss);
                if (lane == src) { v1 = ss; vn1[lc] = ss; vn2[lc] = ss; }
            }
            if (mine) {
                v1 = nan_as_inf(v1);
                if (v1 > bv || (v1 == bv && myc < bi)) { bv = v1; bi = myc; }
            }
        }
        warp_argmax(bv, bi);
        PROF_MARK(3);                      // fused pass + finish + warp argmax
        // ---- speculative reflector for position k+1 from the local candidate (pending = H_k)
        select_and_publish(bv, bi, k + 1, vcur, wcur, true);
    }
    PROF_MARK(4);
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
    PROF_MARK(5);              // extraction of d, R + identity init + barrier

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
    PROF_MARK(6);              // Q phase
    cluster_barrier<CS>();   // every CTA is done reading reflectors from Lg before anyone overwrites them
    for (int lc = warp; lc < nloc; lc += CTA_WARPS) {
        const int c = lc * CS + rank;
        if (c >= n) continue;
        const T* col = A + (size_t)lc * ldc;
        T* dst = Lg + (long long)c * ldl;
        for (int r = lane; r < n; r += 32) dst[r] = col[r];
    }
    PROF_MARK(7);
#ifdef SLA_LDR_PROFILE
    if (prof && (lane == 0) && (warp == 0 || warp == 1)) {
        // [rank][warp0/warp1][8 phases]
        for (int i = 0; i < 8; ++i) p.prof[(rank * 2 + warp) * 8 + i] = pt[i];
    }
#endif
#undef PROF_MARK
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
