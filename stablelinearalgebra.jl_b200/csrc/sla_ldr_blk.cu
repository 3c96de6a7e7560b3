// Batched LDR factorisation, BLOCKED version for real matrices with 128 < n <= 256: column-pivoted Householder QR
// in the LAPACK ?laqps formulation (panel of NB reflectors, ONE pass over the trailing matrix per column, rank-NB
// trailing update), with the trailing update on the FP64 tensor cores.
//
// Replaces `ldr!(F, ws)` (src/LDR.jl:160-188: geqp3! -> d = |diag R'|, R = D^-1 R' P^T); Q is formed afterwards by
// sla_orgq.cu from the reflectors (column k of L) and the taus (p.Twork) this kernel leaves behind (`orgqr!`,
// src/qr.jl:95-109).  Greedy pivoting on down-dated partial column norms with LAPACK's cancellation safeguard, exactly
// as the unblocked register kernel (sla_ldr_reg.cu); what changes is WHERE the flops go:
//
//   * one cluster of 2 CTAs per matrix; column c lives in CTA c % 2.  A warp owns 8 columns and holds them as DMMA
//     ACCUMULATOR FRAGMENTS: lane (fi = lane / 4, fk = lane % 4) holds rows 8 s + 2 fk + {0, 1} of its column fi for
//     the 32 row tiles s -- tiles 16..31 (rows >= 128) in registers, tiles 0..15 in shared memory (thread-private
//     16-byte cells).  Rows retire in order whatever the input looks like, so the shared-memory traffic dies out
//     deterministically after 128 columns;
//   * per column only the GEMV g = v_k^T A runs over the (un-updated) trailing matrix: 2 FMAs per tile and lane and a
//     two-shuffle reduction inside the lane quad (the unblocked kernel: 2 FMAs + an eager rank-1 update per element
//     and an 18-shuffle transposed reduction).  f_k = tau (g - F h), h = V^T v_k comes with the candidate;
//   * every NB = 4 columns the trailing matrix gets A -= V F^T as ONE DMMA.8x8x4 per tile, straight on the
//     fragments (A operand = the lane's own f values, B operand = the reflectors where the candidate exchange left
//     them).  It is issued AFTER the next candidate has been published, i.e. in the shadow of the cluster exchange;
//   * pivot protocol as in the register kernel: every CTA picks its best column through packed integer keys, the
//     owner warp brings that column up to date (lazy panel update), builds the reflector it WOULD generate and ships
//     it with one DSMEM bulk copy per CTA that completes on the receiver's mbarrier -- the only cluster-wide
//     synchronisation per column; both CTAs then pick the same winner.  No physical column swaps: row k of R' is
//     final when column k has been processed and goes straight to R (original column order) in global memory.
#include <cooperative_groups.h>

#include "sla_cta.cuh"
#include "sla_kernels.h"

namespace cg = cooperative_groups;

namespace sla {

namespace {

constexpr double TOL3Z_B = 1.4901161193847656e-08;  // sqrt(eps), LAPACK ?laqps / ?laqp2
constexpr int BK_THREADS = 512, BK_WARPS = 16;
constexpr int BK_NMAX = 256;
constexpr int BK_VLEN = 264;      // a candidate slot: the vector (256 rows, zero from row n on), then its 64-byte meta
constexpr int BK_RING = 8;        // candidate slots of the last 8 columns stay valid (trailing update reads them)
constexpr int BK_RT = 14;         // row tiles held in registers (the highest rows: they live longest) ...
constexpr int BK_ST = 32 - BK_RT; // ... the others in shared memory (thread-private 16-byte cells)
constexpr int BK_GT = 2;          // register tiles per skip group (warp-uniform branches around dead tiles)
static_assert(BK_RT % BK_GT == 0, "whole groups");
#define SLA_KEEP_BRANCH_B asm volatile("")

__device__ __forceinline__ double nan_inf_b(double v) { return (v == v) ? v : INFINITY; }
__device__ __forceinline__ unsigned long long pack_key_b(double v, int c) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(nan_inf_b(v));
    return (b & ~0x3FFull) | (unsigned long long)(1023 - c);
}
__device__ __forceinline__ int key_col_b(unsigned long long key) { return 1023 - (int)(key & 0x3FFull); }

struct alignas(16) BMeta {   // 64 bytes, travels right behind its vector
    unsigned long long key;   // packed (squared down-dated norm, column); 0: no active column left in the sender
    double pad;
    double beta;              // (beta, tau) and the h pairs are read as 16-byte pairs
    double tau;
    double h[4];              // v_i^T v_cand for the reflectors i of the candidate's own panel
};
static_assert(sizeof(BMeta) == 64, "meta is 64 bytes");

constexpr size_t blk_al16(size_t o) { return (o + 15) & ~(size_t)15; }
struct BlkLayout {
    static constexpr size_t offMat = 0;                                                        // [16 warps][16 tiles][32 lanes] double2
    static constexpr size_t offSlot = offMat + (size_t)BK_WARPS * BK_ST * 32 * 16;               // [8][2][VLEN]
    static constexpr size_t offStage = offSlot + (size_t)BK_RING * 2 * BK_VLEN * 8;           // [2][VLEN]
    static constexpr size_t offColbuf = offStage + 2 * (size_t)BK_VLEN * 8;                   // register rows of the candidate column
    static constexpr size_t offKeys = offColbuf + (size_t)BK_RT * 8 * 8;                                    // [2][128]
    static constexpr size_t offTau = offKeys + 2 * 128 * 8;                                   // [256]
    static constexpr size_t offD = offTau + 256 * 8;                                          // [256]
    static constexpr size_t offWin = offD + 256 * 8;                                          // [8] winner rank per ring slot
    static constexpr size_t offMbar = blk_al16(offWin + 8 * 4);
    static constexpr size_t total = offMbar + 2 * 8 + 64;
};

__device__ __forceinline__ void tma_store_b(void* gdst, const void* ssrc, unsigned bytes) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_wait_read_b() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all_b() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_init_b(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_b(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_b(unsigned long long* bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ unsigned map_cluster_b(const void* local, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local)), "r"(rank));
    return r;
}
__device__ __forceinline__ void dsmem_bulk_copy_b(void* dst_local_addr, const void* src, unsigned bytes,
                                                  unsigned long long* bar_local_addr, unsigned rank) {
    const unsigned dst = map_cluster_b(dst_local_addr, rank), bar = map_cluster_b(bar_local_addr, rank);
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "r"(smem_u32(src)), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ double shx_b(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
__device__ __forceinline__ double shb_b(double v, int l) { return __shfl_sync(0xffffffffu, v, l); }
// sum 4 per-lane values over the warp; afterwards every lane of lane-group u (8 consecutive lanes) holds total u
__device__ __forceinline__ double reduce4_b(const double (&a)[4], int lane) {
    double k2[2];
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int i = 0; i < 2; ++i) { const double send = hi ? a[i] : a[i + 2], mine = hi ? a[i + 2] : a[i]; k2[i] = mine + shx_b(send, 16); }
    }
    double v;
    {
        const bool hi = lane & 8;
        const double send = hi ? k2[0] : k2[1], mine = hi ? k2[1] : k2[0];
        v = mine + shx_b(send, 8);
    }
    v += shx_b(v, 4);
    v += shx_b(v, 2);
    v += shx_b(v, 1);
    return v;
}

// -DSLA_BLK_PROFILE: per-phase cycle counters of cluster 0 (profiles/blk_phase_cycles.py): p.prof[0..15] = main-loop
// phases seen by warp 0 of CTA 0, p.prof[16..31] = phases of whichever warp builds the candidate (both CTAs added up)
#ifdef SLA_BLK_PROFILE
#define BPROF_ON const bool bp_on = (p.prof != nullptr) && (blockIdx.x < 2)
#define BPROF_DECL long long bp_t = clock64()
#define BPROF_MAIN(i) do { if (bp_on && tid == 0 && rank == 0) { long long t_ = clock64(); p.prof[i] += t_ - bp_t; bp_t = t_; } } while (0)
#define BPROF_OWN_START long long bo_t = clock64()
#define BPROF_OWN(i) do { if (bp_on && lane == 0) { long long t_ = clock64(); atomicAdd((unsigned long long*)&p.prof[16 + (i)], (unsigned long long)(t_ - bo_t)); bo_t = t_; } } while (0)
#else
#define BPROF_ON
#define BPROF_DECL
#define BPROF_MAIN(i)
#define BPROF_OWN_START
#define BPROF_OWN(i)
#endif

template <int NB> __global__ void __launch_bounds__(BK_THREADS, 1) ldr_blk_kernel(LdrArgs<double> p) {
    static_assert(NB == 4, "one DMMA.8x8x4 per tile and panel");
    using Lay = BlkLayout;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* const smat = reinterpret_cast<double2*>(smem_raw + Lay::offMat);
    double* const slots = reinterpret_cast<double*>(smem_raw + Lay::offSlot);       // [ring][rank][VLEN]
    double* const stage = reinterpret_cast<double*>(smem_raw + Lay::offStage);      // [2][VLEN]
    double* const colbuf = reinterpret_cast<double*>(smem_raw + Lay::offColbuf);
    unsigned long long* const ckey2 = reinterpret_cast<unsigned long long*>(smem_raw + Lay::offKeys);
    double* const taus = reinterpret_cast<double*>(smem_raw + Lay::offTau);
    double* const dvec = reinterpret_cast<double*>(smem_raw + Lay::offD);
    int* const wins = reinterpret_cast<int*>(smem_raw + Lay::offWin);
    unsigned long long* const mbar = reinterpret_cast<unsigned long long*>(smem_raw + Lay::offMbar);

    const int n = p.n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fi = lane >> 2, fk = lane & 3;
    const int rank = (int)cg::this_cluster().block_rank();
    const int b = blockIdx.x >> 1;
    double* const Lg = p.L + (long long)b * p.strideL;
    const int ldl = p.ldl;
    const double* const Ain = p.Ain ? p.Ain + (long long)b * p.strideAin : Lg;
    const int ldin = p.Ain ? p.ldain : ldl;
    double* const Rg = p.R + (long long)b * p.strideR;
    const int ldr = p.ldr;
    const bool bulk_ok = ((((uintptr_t)Lg) & 15) == 0) && (((size_t)ldl * 8) % 16 == 0) && (n % 2 == 0);
    const int n_al = (n + 1) & ~1;
    const int ntiles = (n + 7) >> 3;                    // row tiles that hold data
    const int mycol = ((warp + BK_WARPS * fi) << 1) + rank;   // local column warp + 16 fi of this CTA
    double2* const mytile = smat + (size_t)warp * BK_ST * 32 + lane;   // tile s of this lane: mytile[s * 32]
    auto slot_ptr = [&](int ring, int q) -> double* { return slots + ((size_t)ring * 2 + q) * BK_VLEN; };

    // ------------------------------------------------------------------ load: fragments, squared column norms
    double a[BK_RT][2];
    double ssq = 0.0;
    bool nz = false;
    {
        const bool vec_in = ((ldin & 1) == 0) && ((((uintptr_t)Ain) & 15) == 0);
        const double* src = Ain + (long long)mycol * ldin;
#pragma unroll
        for (int s = 0; s < 32; ++s) {
            const int r = 8 * s + 2 * fk;
            double2 x = make_double2(0.0, 0.0);
            if (mycol < n) {
                if (vec_in && r + 1 < n) x = *reinterpret_cast<const double2*>(src + r);
                else {
                    if (r < n) x.x = src[r];
                    if (r + 1 < n) x.y = src[r + 1];
                }
            }
            ssq = fma(x.x, x.x, ssq);
            ssq = fma(x.y, x.y, ssq);
            nz |= (x.x != 0.0) | (x.y != 0.0);
            if (s < BK_ST) mytile[s * 32] = x;
            else { a[s - BK_ST][0] = x.x; a[s - BK_ST][1] = x.y; }
        }
    }
    ssq += shx_b(ssq, 1);
    ssq += shx_b(ssq, 2);
    bool active = mycol < n;
    // squared partial column norm, replicated in the four lanes of the column's quad
    double vn1 = ssq, rvn1 = 1.0 / vn1, q12 = 1.0;
    {
        const unsigned nzq = __ballot_sync(0xffffffffu, nz);
        const bool colnz = (nzq >> (lane & ~3)) & 0xFu;
        const bool bad = active && ldr_norm2_out_of_range(vn1, colnz);
        if (__any_sync(0xffffffffu, bad) && lane == 0) ldr_raise_range(p.info, b);
    }
    int mypos = -1;          // pivot position of this quad's column once it has been chosen
    double f[NB];            // F(mycol, reflector i of the current panel)
#pragma unroll
    for (int i = 0; i < NB; ++i) f[i] = 0.0;
    unsigned winbits = 0;    // bit i: winner rank of the panel's column i

    for (int r = tid; r < BK_RING * 2 * BK_VLEN; r += BK_THREADS) slots[r] = 0.0;
    for (int r = tid; r < 2 * BK_VLEN; r += BK_THREADS) stage[r] = 0.0;
    auto rz_of = [&](int kpos) { return (kpos >= BK_RING ? kpos - BK_RING : 0) & ~1; };
    auto cbytes_of = [&](int kpos) { return (unsigned)((BK_NMAX - rz_of(kpos)) * 8 + (int)sizeof(BMeta)); };
    if (tid == 0) {
        mbar_init_b(&mbar[0], 1);
        mbar_init_b(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx_b(&mbar[0], 2 * cbytes_of(0));
        if (n > 1) mbar_expect_tx_b(&mbar[1], 2 * cbytes_of(1));
    }
    __syncthreads();
    cg::this_cluster().sync();   // the peer's slots are zeroed and its barriers armed before anyone sends; all inputs read

    BPROF_ON;
    // ------------------------------------------------------------------ candidate for position kpos
    // nlazy = reflectors not yet applied to the trailing matrix when this candidate is built
    auto publish = [&](int kpos, int nlazy) {
        BPROF_OWN_START;
        const int par = kpos & 1, ring = kpos & (BK_RING - 1);
        unsigned long long* const ckey = ckey2 + par * 128;
        if (fk == 0) ckey[warp * 8 + fi] = active ? pack_key_b(vn1, mycol) : 0ull;
        if (tid == 0) tma_wait_read_b();   // the TMA store is done reading the staging buffer that is rewritten below
        __syncthreads();
        const ulonglong2 q0 = *reinterpret_cast<const ulonglong2*>(ckey + lane * 2);
        const ulonglong2 q1 = *reinterpret_cast<const ulonglong2*>(ckey + 64 + lane * 2);
        unsigned long long km = q0.x > q0.y ? q0.x : q0.y;
        const unsigned long long k1 = q1.x > q1.y ? q1.x : q1.y;
        km = k1 > km ? k1 : km;
        const unsigned khi = (unsigned)(km >> 32);
        const unsigned mhi = __reduce_max_sync(0xffffffffu, khi);
        const unsigned mlo = __reduce_max_sync(0xffffffffu, khi == mhi ? (unsigned)km : 0u);
        const unsigned long long ck = ((unsigned long long)mhi << 32) | mlo;
        const int rz = rz_of(kpos);
        const unsigned cbytes = cbytes_of(kpos);
        double* const st = stage + (size_t)par * BK_VLEN;
        if (ck == 0ull) {   // no active column left in this CTA: still send (same sizes) so that the byte counts match
            if (warp == 0) {
                if (lane == 0) {
                    reinterpret_cast<BMeta*>(st + BK_NMAX)->key = 0ull;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
                    for (int q = 0; q < 2; ++q) dsmem_bulk_copy_b(slot_ptr(ring, rank) + rz, st + rz, cbytes, &mbar[par], q);
                }
            }
            return;
        }
        const int ci = key_col_b(ck);
        const int clc = ci >> 1;
        if (warp != (clc & (BK_WARPS - 1))) return;
        BPROF_OWN(0);   // key write + barrier + key scan
        // ---- this warp owns the CTA's best column: gather it (rows lane + 32 t), bring it up to date, build the reflector
        const int bfi = clc >> 4;
        if (fi == bfi) {
#pragma unroll
            for (int s = 0; s < BK_RT; ++s) *reinterpret_cast<double2*>(colbuf + 8 * s + 2 * fk) = make_double2(a[s][0], a[s][1]);
        }
        __syncwarp();
        double x[8];
        {
            const double* const mw = reinterpret_cast<const double*>(smat + (size_t)warp * BK_ST * 32);
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int r = lane + 32 * t;
                const double* src = (r < 8 * BK_ST) ? mw + ((size_t)(r >> 3) * 64 + (bfi * 4 + ((r & 7) >> 1)) * 2 + (r & 1))
                                                    : colbuf + (r - 8 * BK_ST);
                x[t] = *src;
            }
        }
        BPROF_OWN(1);   // gather
        // lazy panel update: x -= V(:, i) F(c, i) for the reflectors not yet applied to the trailing matrix
        for (int i = 0; i < nlazy; ++i) {
            const int ki = kpos - nlazy + i;
            const double* vi = slot_ptr(ki & (BK_RING - 1), wins[ki & (BK_RING - 1)]);
            double fli = f[0];
#pragma unroll
            for (int u = 1; u < NB; ++u) if (i == u) fli = f[u];
            fli = shb_b(fli, bfi * 4);             // F(candidate column, reflector i) lives in the owner quad
#pragma unroll
            for (int t = 0; t < 8; ++t) x[t] = fma(-vi[lane + 32 * t], fli, x[t]);
        }
        BPROF_OWN(2);   // lazy update
        // squared norm below kpos and the dots v_i^T x (rows > kpos) with the reflectors of the candidate's own panel
        const int jj = kpos & (NB - 1);
        double red[4];                                 // squared norm and at most NB - 1 = 3 dots
#pragma unroll
        for (int u = 0; u < 4; ++u) red[u] = 0.0;
        double xk = 0.0;
        const int tk = kpos >> 5;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int r = lane + 32 * t;
            if (r > kpos) red[0] = fma(x[t], x[t], red[0]);
            if (t == tk) xk = x[t];
        }
        for (int i = 0; i < jj; ++i) {
            const int ki = kpos - jj + i;
            const double* vi = slot_ptr(ki & (BK_RING - 1), wins[ki & (BK_RING - 1)]);
            double d = 0.0;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int r = lane + 32 * t;
                if (r > kpos) d = fma(vi[r], x[t], d);
            }
#pragma unroll
            for (int u = 0; u < NB - 1; ++u) if (i == u) red[1 + u] = d;
        }
        const double tot = reduce4_b(red, lane);       // lanes 8u..8u+7 hold total u
        BPROF_OWN(3);   // norm + dots + reduction
        const double xnorm2 = shb_b(tot, 0);
        const double alpha = shb_b(xk, kpos & 31);
        double tau, scal, beta;
        if (xnorm2 == 0.0) {
            tau = 0.0; scal = 0.0; beta = alpha;
        } else {
            // |beta| = sqrt(s), 1/|beta| from one rsqrt + Newton steps (no sqrt/div dependency chain)
            const double sq = fma(alpha, alpha, xnorm2);
            double rr = rsqrt(sq);
            double bb = sq * rr;
            bb = fma(0.5 * rr, fma(-bb, bb, sq), bb);
            double ib = rr;
            ib = fma(ib, fma(-bb, ib, 1.0), ib);
            const double aa = fabs(alpha);
            beta = -copysign(bb, alpha);
            tau = fma(aa, ib, 1.0);
            const double den = aa + bb;
            scal = copysign(__drcp_rn(den), alpha);
            if (!(sq > 1e-280 && sq < 1e280)) {
                beta = -copysign(sqrt(sq), alpha);
                tau = (beta - alpha) / beta;
                scal = 1.0 / (alpha - beta);
            }
        }
        BPROF_OWN(4);   // reflector scalars
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int r = lane + 32 * t;
            if (r < n_al && r >= rz) {
                double v = 0.0;
                if (r > kpos && r < n) v = x[t] * scal;
                else if (r == kpos) v = 1.0;
                st[r] = v;
            }
        }
        // meta: lanes 8, 16, 24 hold the dots with the panel's reflectors 0..2
        {
            BMeta* m = reinterpret_cast<BMeta*>(st + BK_NMAX);
            if (lane == 0) { m->key = ck; m->beta = beta; m->tau = tau; m->pad = 0.0; m->h[3] = 0.0; }
            const int u = (lane >> 3) - 1;
            if ((lane & 7) == 0 && u >= 0) {
                double hv = 0.0;
                if (u < jj) {
                    const int ki = kpos - jj + u;
                    hv = fma(scal, tot, slot_ptr(ki & (BK_RING - 1), wins[ki & (BK_RING - 1)])[kpos]);
                }
                m->h[u] = hv;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < 2; ++q) dsmem_bulk_copy_b(slot_ptr(ring, rank) + rz, st + rz, cbytes, &mbar[par], q);
        }
        BPROF_OWN(5);   // staging writes + fence + copy issue
    };

    publish(0, 0);
    BPROF_DECL;

    // ------------------------------------------------------------------ pivoted Householder QR, panels of NB columns
    for (int k0 = 0; k0 < n; k0 += NB) {
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const int k = k0 + j;
            if (k >= n) break;
            const int par = k & 1, ring = k & (BK_RING - 1);
            BPROF_MAIN(6);   // loop overhead
            mbar_wait_b(&mbar[par], (unsigned)((k >> 1) & 1));
            BPROF_MAIN(0);   // wait for the candidates
            if (tid == 0 && k + 2 < n) mbar_expect_tx_b(&mbar[par], 2 * cbytes_of(k + 2));
            const unsigned long long key0 = reinterpret_cast<const BMeta*>(slot_ptr(ring, 0) + BK_NMAX)->key;
            const unsigned long long key1 = reinterpret_cast<const BMeta*>(slot_ptr(ring, 1) + BK_NMAX)->key;
            const int win = key1 > key0 ? 1 : 0;
            const BMeta* const wm = reinterpret_cast<const BMeta*>(slot_ptr(ring, win) + BK_NMAX);
            const int pc = key_col_b(win ? key1 : key0);
            const double2 bt = *reinterpret_cast<const double2*>(&wm->beta);
            const double betak = bt.x, tauk = bt.y;
            double hh[NB];
            {
                const double2 h01 = *reinterpret_cast<const double2*>(&wm->h[0]);
                const double2 h23 = *reinterpret_cast<const double2*>(&wm->h[2]);
                hh[0] = h01.x; hh[1] = h01.y; hh[2] = h23.x; hh[3] = h23.y;
            }
            const double* const vcur = slot_ptr(ring, win);
            winbits = (winbits & ~(1u << j)) | ((unsigned)win << j);
            if (tid == 0) { taus[k] = tauk; dvec[k] = fabs(betak); wins[ring] = win; }
            const bool was_active = active;
            const bool is_pivot = (mycol == pc);
            if (is_pivot) { active = false; mypos = k; }
            const int q = k >> 3;                     // tile of row k; tiles below hold only rows < k (v is zero there)
            const int hold_fk = (k & 7) >> 1;         // the lane of a quad that holds row k
            // ---- GEMV over the live tiles: g = sum_r v(r) A(r, mycol); the element A(k, mycol) is picked up on the way
            double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
            double rowval = 0.0;
            if (k + 1 < n) {
                const double2* const v2 = reinterpret_cast<const double2*>(vcur) + fk;   // rows 8 s + 2 fk, +1 at v2[4 s]
                {
                    const int s_hi = ntiles < BK_ST ? ntiles : BK_ST;
                    int s = q;
                    for (; s + 1 < s_hi; s += 2) {
                        const double2 va = v2[4 * s], vb = v2[4 * s + 4];
                        const double2 xa = mytile[s * 32], xb = mytile[(s + 1) * 32];
                        g0 = fma(va.x, xa.x, g0); g1 = fma(va.y, xa.y, g1);
                        g2 = fma(vb.x, xb.x, g2); g3 = fma(vb.y, xb.y, g3);
                    }
                    if (s < s_hi) {
                        const double2 va = v2[4 * s];
                        const double2 xa = mytile[s * 32];
                        g0 = fma(va.x, xa.x, g0); g1 = fma(va.y, xa.y, g1);
                    }
                    if (q < BK_ST) { const double2 xr = mytile[q * 32]; rowval = (k & 1) ? xr.y : xr.x; }
                }
#pragma unroll
                for (int gq = 0; gq < BK_RT / BK_GT; ++gq) {
                    const int sb = BK_ST + BK_GT * gq;
                    if (q < sb + BK_GT && sb < ntiles) {
                        SLA_KEEP_BRANCH_B;
#pragma unroll
                        for (int u = 0; u < BK_GT; ++u) {
                            const double2 vv = v2[4 * (sb + u)];
                            if (u & 1) { g2 = fma(vv.x, a[BK_GT * gq + u][0], g2); g3 = fma(vv.y, a[BK_GT * gq + u][1], g3); }
                            else { g0 = fma(vv.x, a[BK_GT * gq + u][0], g0); g1 = fma(vv.y, a[BK_GT * gq + u][1], g1); }
                        }
                        if (q >= sb) {
#pragma unroll
                            for (int u = 0; u < BK_GT; ++u)
                                if (q == sb + u) rowval = (k & 1) ? a[BK_GT * gq + u][1] : a[BK_GT * gq + u][0];
                        }
                    }
                }
            }
            BPROF_MAIN(1);   // meta + GEMV
            double g = (g0 + g1) + (g2 + g3);
            g += shx_b(g, 1);
            g += shx_b(g, 2);
            rowval = shb_b(rowval, (lane & ~3) | hold_fk);
            // ---- f_k = tau (g - sum_{i<j} F(:, i) h_i);  row k of the updated matrix; down-date of the squared norm
            double fj = g, ahat = rowval;
#pragma unroll
            for (int i = 0; i < NB; ++i)
                if (i < j) {
                    fj = fma(-f[i], hh[i], fj);
                    const double vki = slot_ptr((k0 + i) & (BK_RING - 1), (winbits >> i) & 1u)[k];
                    ahat = fma(-vki, f[i], ahat);
                }
            fj *= tauk;
            f[j] = fj;
            ahat -= fj;                               // V(k, j) = 1
            double temp = 1.0;
            bool flagged = false;
            if (active && vn1 != 0.0) {
                temp = fmax(0.0, fma(-(ahat * ahat), rvn1, 1.0));
                flagged = temp * q12 <= TOL3Z_B;
                if (!flagged) vn1 *= temp;
            }
            if (__any_sync(0xffffffffu, flagged)) {
                // exact squared norm of the flagged columns below row k (panel reflectors 0..j applied on the fly)
                double ss = 0.0;
                if (flagged) {
                    auto tile_ss = [&](double2 xx, int s) {
                        const int r = 8 * s + 2 * fk;
#pragma unroll
                        for (int i = 0; i < NB; ++i)
                            if (i <= j) {
                                const double2 vv = *reinterpret_cast<const double2*>(slot_ptr((k0 + i) & (BK_RING - 1), (winbits >> i) & 1u) + r);
                                xx.x = fma(-vv.x, f[i], xx.x);
                                xx.y = fma(-vv.y, f[i], xx.y);
                            }
                        if (r > k) ss = fma(xx.x, xx.x, ss);
                        if (r + 1 > k) ss = fma(xx.y, xx.y, ss);
                    };
                    for (int s = q; s < BK_ST && s < ntiles; ++s) tile_ss(mytile[s * 32], s);
#pragma unroll
                    for (int u = 0; u < BK_RT; ++u)       // static register indices (a dynamic index would move `a` to local memory)
                        if (BK_ST + u >= q && BK_ST + u < ntiles) tile_ss(make_double2(a[u][0], a[u][1]), BK_ST + u);
                }
                ss += shx_b(ss, 1);
                ss += shx_b(ss, 2);
                if (flagged) { vn1 = ss; q12 = 1.0; temp = 1.0; }
            }
            BPROF_MAIN(2);   // reduction, f, row k, down-date (+ exact recomputation)
            // ---- candidate for the next column (the chain); everything below is off the critical path
            if (k + 1 < n) publish(k + 1, j + 1);
            BPROF_MAIN(3);   // publish (non-owner: barrier + key scan)
            q12 *= temp;
            rvn1 = 1.0 / vn1;
            // row k of R = D^-1 R' P^T, in original column order: final now
            if (fk == hold_fk && mycol < n) {
                const double dk = fabs(betak);
                if (is_pivot) Rg[(long long)mycol * ldr + k] = betak / dk;
                else if (was_active) Rg[(long long)mycol * ldr + k] = ahat / dk;
            }
            // reflector k -> column k of L (rows > k) by the CTA that built it, from its staging copy
            if (rank == win && k + 1 < n) {
                const double* src = stage + (size_t)par * BK_VLEN;
                const int r0 = (k + 1) & ~1;
                if (bulk_ok) {
                    if (tid == 0) tma_store_b(Lg + (long long)k * ldl + r0, src + r0, (unsigned)((n - r0) * 8));
                } else {
                    for (int r = k + 1 + tid; r < n; r += BK_THREADS) Lg[(long long)k * ldl + r] = src[r];
                }
            }
            BPROF_MAIN(4);   // R row, reflector store
            // ---- end of the panel: A -= V F^T on the tensor cores, one DMMA per live tile, in the exchange shadow
            if (j == NB - 1 && k + 1 < n) {
                const int s_lo = (k + 1) >> 3;
                double fs = f[0];
#pragma unroll
                for (int u = 1; u < NB; ++u) if (fk == u) fs = f[u];
                fs = -fs;
                const int ws = (winbits >> fk) & 1u;
                const double* const vr = slot_ptr((k0 + fk) & (BK_RING - 1), ws) + fi;   // V(row 8 s + fi, reflector fk)
                {
                    const int s_hi = ntiles < BK_ST ? ntiles : BK_ST;
                    for (int s = s_lo; s < s_hi; ++s) {
                        double2 acc = mytile[s * 32];
                        dmma884(acc.x, acc.y, fs, vr[8 * s]);
                        mytile[s * 32] = acc;
                    }
                }
#pragma unroll
                for (int gq = 0; gq < BK_RT / BK_GT; ++gq) {
                    const int sb = BK_ST + BK_GT * gq;
                    if (s_lo < sb + BK_GT && sb < ntiles) {
                        SLA_KEEP_BRANCH_B;
#pragma unroll
                        for (int u = 0; u < BK_GT; ++u) dmma884(a[BK_GT * gq + u][0], a[BK_GT * gq + u][1], fs, vr[8 * (sb + u)]);
                    }
                }
            }
            BPROF_MAIN(5);   // trailing update
        }
    }
    if (tid == 0) tma_wait_all_b();   // all reflectors have landed in global memory
    __syncthreads();

    // ------------------------------------------------------------------ zeros below the pivot position, d, taus
    if (mycol < n) {
        double* dst = Rg + (long long)mycol * ldr;
        for (int s = (mypos + 1) >> 3; s < ntiles; ++s) {
            const int r = 8 * s + 2 * fk;
            if (r > mypos && r < n) dst[r] = 0.0;
            if (r + 1 > mypos && r + 1 < n) dst[r + 1] = 0.0;
        }
    }
    if (rank == 0) {
        double* d = p.d + (long long)b * p.stride_d;
        double* tau_g = p.Twork + (long long)b * p.strideT;
        for (int r = tid; r < n; r += BK_THREADS) { d[r] = dvec[r]; tau_g[r] = taus[r]; }
    }
    cg::this_cluster().sync();   // no CTA exits while the peer may still read its staging buffer / write its slots
}

}  // namespace

template <> int ldr_blk_cluster_size<double>(int n) { return (n > 128 && n <= BK_NMAX) ? 2 : 0; }
template <> int ldr_blk_cluster_size<cplx>(int) { return 0; }

template <> cudaError_t ldr_blk_batched<double>(const LdrArgs<double>& p, cudaStream_t stream) {
    if (p.batch <= 0 || p.n <= 0) return cudaSuccess;
    if (ldr_blk_cluster_size<double>(p.n) == 0 || p.Twork == nullptr || p.strideT < p.n) return cudaErrorNotSupported;
    auto kern = ldr_blk_kernel<4>;
    const size_t smem = BlkLayout::total;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)p.batch * 2, 1, 1);
    cfg.blockDim = dim3(BK_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, p);
    ++g_launch_count;
    if (e != cudaSuccess) return e;
    return orgq_dmma_batched(p, stream);   // Q = H_0 ... H_{n-1} I on the tensor cores
}
template <> cudaError_t ldr_blk_batched<cplx>(const LdrArgs<cplx>&, cudaStream_t) { return cudaErrorNotSupported; }

}  // namespace sla
