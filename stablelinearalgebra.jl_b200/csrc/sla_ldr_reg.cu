// Batched LDR factorisation, register-resident version: one thread-block cluster owns one matrix and the
// matrix lives in the REGISTER FILES (and, in the hybrid shape, half of it in the shared memory) of its CTAs
// for the whole factorisation.
//
// Replaces `ldr!(F, ws)` (src/LDR.jl:160-188: geqp3! + triu/abs/ldiv!/mul_P! + orgqr!) for n <= 256.
// Same mathematics as LAPACK's ?laqp2/?larfg/?org2r, own formulation:
//   * a warp owns CPW columns (column-cyclic over CTAs, then over warps) as a CPW x RPL register tile per lane
//     (row r = lane + 32 t): the per-column pass is register math -- RPL shared-memory loads for v_k, a dot
//     product per column, a transposed warp reduction, an eager rank-1 update;
//   * speculative pivots: every CTA picks its best column through packed 64-bit keys (two 32-bit warp `redux`
//     reductions), the warp that holds it builds the reflector that column WOULD generate and ships it (vector,
//     then key/beta/tau behind it) with one DSMEM bulk copy per peer that completes on the peer's mbarrier --
//     the only cluster-wide synchronisation per column; all CTAs then pick the same winner;
//   * no physical column swaps; squared-norm down-dating with LAPACK's cancellation safeguard, the bookkeeping
//     replicated in every lane of a column's lane group (no divergence);
//   * off the critical path (behind the pivot scan): parking of the retired column, the TMA store of the
//     reflector, the reciprocals of the next down-dating;
//   * HYBRID shape (SPW > 0): a warp owns SPW more columns in shared memory (thread-private 16-byte pairs), so a
//     matrix needs half as many SMs and 64 matrices of N = 256 are resident in ONE wave instead of two;
//   * MODE 0 factorises and forms Q = H_0...H_{n-1} I in one launch (reflectors stream back through a cp.async
//     ring); MODE 1 only factorises; MODE 2 only forms Q, four reflectors per reduction (compact WY).
#include <cooperative_groups.h>

#include "sla_cta.cuh"
#include "sla_kernels.h"

namespace cg = cooperative_groups;

namespace sla {

namespace {

constexpr double TOL3Z_R = 1.4901161193847656e-08;  // sqrt(eps), LAPACK ?laqps / ?laqp2
constexpr int QKC_R = 8;                             // reflectors per chunk in the Q phase
// an empty volatile asm cannot be speculated: keeps a warp-uniform `if` a real branch (no if-conversion into selects)
#define SLA_KEEP_BRANCH asm volatile("")

// Pivot candidates are compared as packed 64-bit integer keys (no FP64 compares on the critical path):
// the bit pattern of a non-negative double is monotone, its 10 low mantissa bits are replaced by
// (1023 - column) so that ties go to the smaller column index.  0 = "no candidate".
__device__ __forceinline__ double nan_inf(double v) { return (v == v) ? v : INFINITY; }
__device__ __forceinline__ unsigned long long pack_key(double v, int c) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(nan_inf(v));
    return (b & ~0x3FFull) | (unsigned long long)(1023 - c);
}
__device__ __forceinline__ int key_col(unsigned long long key) { return 1023 - (int)(key & 0x3FFull); }

template <typename T> struct alignas(16) RMeta {   // 32 bytes for both element types (bulk copies need 16-byte multiples)
    unsigned long long key;   // packed (squared down-dated norm, column); 0: no active column left
    double beta;
    T tau;
};

// Shared-memory layout, fixed at compile time (sized for NMAX, not n) so that every address in the per-column
// chain is base + immediate.
template <typename T, int RPL, int CPW, int CS, int SPW> struct RegLayout {
    static constexpr int NMAX = 32 * RPL;
    static constexpr int VLEN = NMAX + 8;
    static constexpr size_t al16(size_t o) { return (o + 15) & ~(size_t)15; }
    // candidate reflectors [2][CS][VLEN] (a slot = the vector, then its RMeta at element n_al); the
    // Q-phase ring [2][QKC][NMAX] aliases this region
    static constexpr size_t candBytes = 2 * (size_t)CS * VLEN * sizeof(T), ringBytes = 2 * (size_t)QKC_R * NMAX * sizeof(T);
    static constexpr size_t offCand = 0;
    static constexpr size_t offRst = offCand + (candBytes > ringBytes ? candBytes : ringBytes);   // parked (raw) pivot columns [lc][NMAX]
    // SPW == 0: parking area for retired register columns.  SPW > 0 (hybrid): the shared-memory resident columns
    // [warp][SPW][RPL/2][32 lanes][2] (thread-private, 16-byte pairs); retired register columns go to global R
    static constexpr size_t offTau = offRst + (size_t)CTA_WARPS * (SPW > 0 ? SPW : CPW) * NMAX * sizeof(T);
    static constexpr size_t offStageV = al16(offTau + (size_t)NMAX * sizeof(T));   // [2][VLEN] staging of the published candidate
    static constexpr size_t offMbar = offStageV + 2 * (size_t)VLEN * sizeof(T);    // one mbarrier per column parity
    static constexpr size_t offD = al16(offMbar + 2 * sizeof(unsigned long long));
    static constexpr size_t offBeta = offD + (size_t)NMAX * sizeof(double);        // signed beta_k = R'(k, pivot column k)
    static constexpr size_t offKeys = al16(offBeta + (size_t)NMAX * sizeof(double));   // [2][16*CPW] pivot keys
    static constexpr size_t offPos = offKeys + 2 * (size_t)CTA_WARPS * (CPW + SPW) * sizeof(unsigned long long);   // pivot position per local column
    static constexpr size_t offWsc = al16(offPos + (size_t)CTA_WARPS * (CPW + SPW) * sizeof(int));   // per-warp w broadcast (hybrid)
    static constexpr size_t total = offWsc + (size_t)CTA_WARPS * 16 * sizeof(double) + 64;   // [warp][8 w | 4 row-k elements | 4 spare]
};

template <int CS> __device__ __forceinline__ void cl_barrier() {
    if constexpr (CS == 1) __syncthreads();
    else cg::this_cluster().sync();
}
template <int CS, typename P> __device__ __forceinline__ P* cl_peer(P* local, int rank) {
    if constexpr (CS == 1) return local;
    else return cg::this_cluster().map_shared_rank(local, rank);
}

__device__ __forceinline__ double shx(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
__device__ __forceinline__ cplx shx(cplx v, int o) { return cplx{shx(v.re, o), shx(v.im, o)}; }
__device__ __forceinline__ double shb(double v, int l) { return __shfl_sync(0xffffffffu, v, l); }
__device__ __forceinline__ cplx shb(cplx v, int l) { return cplx{shb(v.re, l), shb(v.im, l)}; }
// Sum G (power of two, <= 8) per-lane values over the warp with far fewer shuffles than G butterflies:
// each round halves the number of values a lane carries.  Afterwards every lane of lane-group u
// (32/G consecutive lanes) holds the total of a[u].
template <typename T, int G> __device__ __forceinline__ T reduce_transposed_inner(const T (&a)[G], int lane, int bit) {
    if constexpr (G == 1) {
        T v = a[0];
        for (int o = bit; o > 0; o >>= 1) v = v + shx(v, o);
        return v;
    } else {
        constexpr int H = G / 2;
        const bool hi = lane & bit;
        T keep[H];
#pragma unroll
        for (int i = 0; i < H; ++i) {
            T send = hi ? a[i] : a[i + H];
            T mine = hi ? a[i + H] : a[i];
            keep[i] = mine + shx(send, bit);
        }
        return reduce_transposed_inner<T, H>(keep, lane, bit >> 1);
    }
}
template <typename T, int G> __device__ __forceinline__ T reduce_transposed(const T (&a)[G], int lane) {
    return reduce_transposed_inner<T, G>(a, lane, 16);
}

template <typename T, int G> __device__ __forceinline__ T pickv(const T (&a)[G], int u) {
    T r = a[0];
#pragma unroll
    for (int i = 1; i < G; ++i) if (u == i) r = a[i];
    return r;
}

__device__ __forceinline__ void tma_store(void* gdst, const void* ssrc, unsigned bytes) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- mbarrier + DSMEM bulk copy (the candidate exchange of clusters with CS > 1) ----------------------
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
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
__device__ __forceinline__ unsigned map_cluster(const void* local, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local)), "r"(rank));
    return r;
}
// bulk copy local shared -> shared memory of CTA `rank`, completing `bytes` on that CTA's mbarrier
__device__ __forceinline__ void dsmem_bulk_copy(void* dst_local_addr, const void* src, unsigned bytes,
                                                unsigned long long* bar_local_addr, unsigned rank) {
    const unsigned dst = map_cluster(dst_local_addr, rank), bar = map_cluster(bar_local_addr, rank);
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "r"(smem_u32(src)), "r"(bytes), "r"(bar)
                 : "memory");
}

}  // namespace

// SPW > 0 is the HYBRID shape: besides its CPW register columns a warp owns SPW columns that live in shared
// memory (thread-private layout: a lane only ever touches its own rows), so that a matrix needs half as many
// SMs and twice as many matrices are resident -- one wave instead of two for 64 chains of N = 256.
// MODE 0: factorisation + Q formation in one launch.  MODE 1: factorisation only (reflectors -> L, taus -> p.Twork).
// MODE 2: Q formation only, from the reflectors and taus a MODE 1 launch left behind -- used to form Q with the
// register shape (4 SMs per matrix, no shared-memory half) after the hybrid shape has factorised in one wave.
template <typename T, int CS, int RPL, int CPW, int SPW, int MODE>
__global__ void __launch_bounds__(CTA_THREADS, 1) ldr_reg_kernel(LdrArgs<T> p) {
    using Lay = RegLayout<T, RPL, CPW, CS, SPW>;
    constexpr int NMAX = Lay::NMAX, VLEN = Lay::VLEN;
    constexpr int TOT = CPW + SPW;                // columns per warp (register + shared)
    constexpr int FLS = 32 / TOT;                 // lanes per lane-group; lane u*FLS leads column u of the warp
    constexpr bool CP = is_cplx<T>::value;
    static_assert(SPW == 0 || (!CP && RPL % 2 == 0), "hybrid shape: real only");
    constexpr int EPV = CP ? 1 : 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n;
    T* const vcand = reinterpret_cast<T*>(smem_raw + Lay::offCand);        // [2][CS][VLEN]
    T* const Rst = reinterpret_cast<T*>(smem_raw + Lay::offRst);           // [16*CPW][NMAX]   (SPW == 0)
    // hybrid: element (su, t) of this lane at scol[(su*RPL/2 + t/2)*64 + (t&1)]  (16-byte pairs of row tiles)
    T* const scol = Rst + ((size_t)(threadIdx.x >> 5) * (SPW > 0 ? SPW : 1) * RPL * 32) + (threadIdx.x & 31) * 2;
    [[maybe_unused]] const T* const swarp = Rst + ((size_t)(threadIdx.x >> 5) * (SPW > 0 ? SPW : 1) * RPL * 32);
    [[maybe_unused]] double* const wsc = reinterpret_cast<double*>(smem_raw + Lay::offWsc) + (threadIdx.x >> 5) * 16;
    auto sld = [&](int su, int t) -> T { return scol[(su * (RPL / 2) + (t >> 1)) * 64 + (t & 1)]; };
    auto sst = [&](int su, int t, T v) { scol[(su * (RPL / 2) + (t >> 1)) * 64 + (t & 1)] = v; };
    T* const taus = reinterpret_cast<T*>(smem_raw + Lay::offTau);
    T* const stage_v = reinterpret_cast<T*>(smem_raw + Lay::offStageV);                 // [2][VLEN] to be sent
    unsigned long long* const mbar = reinterpret_cast<unsigned long long*>(smem_raw + Lay::offMbar);
    double* const dvec = reinterpret_cast<double*>(smem_raw + Lay::offD);
    double* const bsg = reinterpret_cast<double*>(smem_raw + Lay::offBeta);
    int* const cpos = reinterpret_cast<int*>(smem_raw + Lay::offPos);
    unsigned long long* const ckey2 = reinterpret_cast<unsigned long long*>(smem_raw + Lay::offKeys);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int rank = 0;
    if constexpr (CS > 1) rank = (int)cg::this_cluster().block_rank();
    const int b = blockIdx.x / CS;
    T* Lg = p.L + (long long)b * p.strideL;
    const int ldl = p.ldl;
    const T* Ain = p.Ain ? p.Ain + (long long)b * p.strideAin : Lg;
    const int ldin = p.Ain ? p.ldain : ldl;
    const bool bulk_ok = ((((uintptr_t)Lg) & 15) == 0) && (((size_t)ldl * sizeof(T)) % 16 == 0) && (CP || (n % 2 == 0));

    // ------------------------------------------------------------------ load: columns -> registers
    // column u of this warp: local index lc = warp + 16 u, global c = lc*CS + rank
    // Slot u of a warp (u < CPW: registers, else shared memory) holds local column lc = warp + 16 * grp(u), global
    // c = lc*CS + rank.  In the hybrid shape the SHARED slots take the low column groups: a DQMC chain hands
    // over (B L) D with D sorted descending, so the low columns are pivoted (retired) first and the shared-memory
    // traffic dies out early; for Q = H_0...H_{n-1} I the low columns wake up last.  Any input is handled, this
    // only decides which half is the cheaper one.
    auto grp = [&](int u) { return (u + SPW) % TOT; };
    auto slot_of_grp = [&](int g) { return (g + CPW) % TOT; };
    auto colof = [&](int u) { return (warp + CTA_WARPS * grp(u)) * CS + rank; };
    T a[CPW][RPL];
    T* const tau_g = p.Twork + (long long)b * p.strideT;   // taus between a MODE 1 and a MODE 2 launch
#ifdef SLA_LDR_PROFILE
    long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bool prof = (p.prof != nullptr) && (blockIdx.x < CS);
    long long tq = 0;
#endif
    if constexpr (MODE != 2) {
    unsigned active = 0, nzmask = 0;
    double ssq[TOT];
#pragma unroll
    for (int u = 0; u < TOT; ++u) {
        const int c = colof(u);
        ssq[u] = 0.0;
        if (c < n) active |= 1u << u;
        const T* src = Ain + (long long)c * ldin;
#pragma unroll
        for (int t = 0; t < RPL; ++t) {
            const int r = lane + 32 * t;
            const T val = (c < n && r < n) ? src[r] : mk<T>(0.0);
            ssq[u] += abs2_(val);
            if (!is_zero(val)) nzmask |= 1u << u;
            if (u < CPW) a[u < CPW ? u : 0][t] = val;
            else sst(u - CPW, t, val);
        }
    }
    // squared partial column norms, replicated in every lane of the column's lane group (no divergence in the
    // down-dating): vn1 = current estimate, rvn1 = 1/vn1, q12 = vn1 / (value at the last exact computation)
    const int myu = lane / FLS;
    const bool leader = (lane % FLS) == 0;
    double vn1 = reduce_transposed<double, TOT>(ssq, lane);
    double rvn1 = 1.0 / vn1, q12 = 1.0;
    const int myc = colof(myu);
    {   // range guard of the squared norms (sla_kernels.h)
        nzmask = __reduce_or_sync(0xffffffffu, nzmask);
        const bool bad = ((active >> myu) & 1u) && ldr_norm2_out_of_range(vn1, (nzmask >> myu) & 1u);
        if (__any_sync(0xffffffffu, bad) && lane == 0) ldr_raise_range(p.info, b);
    }
    const int n_al = (n + EPV - 1) & ~(EPV - 1);
    constexpr int NKEY = CTA_WARPS * TOT;   // one key slot per (warp, column)
    constexpr int KPL = NKEY / 32;          // keys a lane scans
    static_assert(TOT == 2 || TOT == 4 || TOT == 8, "columns per warp");
    // the RMeta of a candidate travels right behind its vector (one bulk copy per peer)
    auto slot_meta = [&](T* slot) { return reinterpret_cast<RMeta<T>*>(slot + n_al); };
    // rows [rz, kpos) of a candidate are written as zeros: the slot still holds an older vector there and the
    // pass / the TMA read from an aligned start
    auto rz_of = [&](int kpos) { return (kpos >= 2 ? kpos - 2 : 0) & ~(EPV - 1); };
    for (int r = tid; r < 2 * VLEN; r += CTA_THREADS) stage_v[r] = mk<T>(0.0);
    auto xbytes = [&](int kpos) { return (unsigned)(CS * ((n_al - rz_of(kpos)) * sizeof(T) + sizeof(RMeta<T>))); };
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if constexpr (CS > 1) {
            mbar_expect_tx(&mbar[0], xbytes(0));
            if (n > 1) mbar_expect_tx(&mbar[1], xbytes(1));
        }
    }

    for (int r = tid; r < 2 * CS * VLEN; r += CTA_THREADS) vcand[r] = mk<T>(0.0);
    __syncthreads();
    cl_barrier<CS>();   // peers' candidate buffers are zeroed before anyone writes into them (also: all inputs read)

    // warp candidate -> shared; block barrier; the warp holding the CTA's best column builds the reflector
    // for position kpos from its registers and pushes it (plus norm, index, tau, beta) to every CTA.
#ifdef SLA_LDR_TRACE
#define TR(kk, ev) do { if (p.prof && blockIdx.x < CS && lane == 0 && (kk) >= 100 && (kk) < 104) p.prof[128 + ((((kk) - 100) * CS + rank) * 16 + warp) * 12 + (ev)] = clock64(); } while (0)
#else
#define TR(kk, ev) do { } while (0)
#endif
#ifdef SLA_LDR_PROFILE
#define SP_MARK(i) do { if (p.prof && blockIdx.x < CS && lane == 0) { long long t_ = clock64(); atomicAdd((unsigned long long*)&p.prof[64 + (i)], (unsigned long long)(t_ - tsp)); tsp = t_; } } while (0)
#else
#define SP_MARK(i) do { } while (0)
#endif
    auto select_and_publish = [&](int kpos) {
#ifdef SLA_LDR_PROFILE
        long long tsp = clock64();
#endif
        const int par = kpos & 1;
        unsigned long long* const ckey = ckey2 + par * NKEY;
        if (leader) ckey[warp * TOT + myu] = ((active >> myu) & 1u) ? pack_key(vn1, myc) : 0ull;
        TR(kpos - 1, 5);
        if (tid == 0) tma_wait_read();   // the TMA finished reading the slot that peers may overwrite from now on
        if (warp == 0) SP_MARK(0);   // key store
        __syncthreads();
        if (warp == 0) SP_MARK(1);   // wait for the slowest warp
        TR(kpos - 1, 6);
        // CTA-wide argmax of the NKEY keys: KPL keys per lane, then two 32-bit warp reductions (high word first)
        unsigned long long km;
        if constexpr (KPL == 1) {
            km = ckey[lane];
        } else {
            // 16-byte lane stride (conflict-free): keys (2 lane, 2 lane + 1) and, for 128 keys, (64 + 2 lane, 65 + 2 lane)
            const ulonglong2 q0 = *reinterpret_cast<const ulonglong2*>(ckey + lane * 2);
            km = q0.x > q0.y ? q0.x : q0.y;
            if constexpr (KPL == 4) {
                const ulonglong2 q1 = *reinterpret_cast<const ulonglong2*>(ckey + 64 + lane * 2);
                const unsigned long long k1 = q1.x > q1.y ? q1.x : q1.y;
                km = k1 > km ? k1 : km;
            }
        }
        const unsigned khi = (unsigned)(km >> 32);
        const unsigned mhi = __reduce_max_sync(0xffffffffu, khi);
        const unsigned mlo = __reduce_max_sync(0xffffffffu, khi == mhi ? (unsigned)km : 0u);
        const unsigned long long ck = ((unsigned long long)mhi << 32) | mlo;
        TR(kpos - 1, 7);
        const int rz = rz_of(kpos);
        const unsigned cbytes = (unsigned)((n_al - rz) * sizeof(T) + sizeof(RMeta<T>));   // vector tail + meta
        if (ck == 0ull) {   // no active column left in this CTA: still send (same sizes) so that the byte counts match
            if (warp == 0) {
                if constexpr (CS > 1) {
                    T* const st = stage_v + (size_t)par * VLEN;
                    if (lane == 0) {
                        slot_meta(st)->key = 0ull;
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
                        for (int q = 0; q < CS; ++q)
                            dsmem_bulk_copy(vcand + ((size_t)par * CS + rank) * VLEN + rz, st + rz, cbytes, &mbar[par], q);
                    }
                } else {
                    if (lane == 0) slot_meta(vcand + (size_t)par * VLEN)->key = 0ull;
                }
            }
            return;
        }
        const int ci = key_col(ck);
        const int clc = ci / CS;
        if (warp != (clc % CTA_WARPS)) return;
        const int ub = slot_of_grp(clc / CTA_WARPS);   // which of my columns
        SP_MARK(2);   // (winner warp) key scan
        T x[RPL];
#pragma unroll
        for (int t = 0; t < RPL; ++t) x[t] = a[0][t];
#pragma unroll
        for (int u = 1; u < CPW; ++u)
            if (ub == u) {
#pragma unroll
                for (int t = 0; t < RPL; ++t) x[t] = a[u][t];
            }
        if constexpr (SPW > 0) {
            if (ub >= CPW) {
#pragma unroll
                for (int t = 0; t < RPL; ++t) x[t] = sld(ub - CPW, t);
            }
        }
        double sq_[RPL];
        T xk = mk<T>(0.0);
        const int tk = kpos >> 5;
#pragma unroll
        for (int t = 0; t < RPL; ++t) {
            const int r = lane + 32 * t;
            sq_[t] = (r > kpos) ? abs2_(x[t]) : 0.0;
            if (t == tk) xk = x[t];
        }
#pragma unroll
        for (int st = RPL / 2; st > 0; st >>= 1)
#pragma unroll
            for (int t = 0; t < st; ++t) sq_[t] += sq_[t + st];
        const double xnorm2 = warp_sum(sq_[0]);
        const T alpha = shb(xk, kpos & 31);
        TR(kpos - 1, 8);
        SP_MARK(3);   // column select + norm reduce
        T tau, scal; double beta;
        if (xnorm2 == 0.0 && imag_(alpha) == 0.0) {
            tau = mk<T>(0.0); scal = mk<T>(0.0); beta = real_(alpha);
        } else {
            if constexpr (!CP) {
                // |beta| = sqrt(s), 1/|beta| from one rsqrt + Newton steps (no sqrt/div dependency chain)
                const double sq = fma(alpha, alpha, xnorm2);
                double r = rsqrt(sq);
                double bb = sq * r;
                bb = fma(0.5 * r, fma(-bb, bb, sq), bb);          // |beta| to < 1 ulp
                double ib = r;
                ib = fma(ib, fma(-bb, ib, 1.0), ib);              // 1/|beta| (r is already ~1 ulp; one correction)
                const double aa = fabs(alpha);
                beta = -copysign(bb, alpha);
                tau = fma(aa, ib, 1.0);                            // (beta - alpha)/beta = 1 + |alpha|/|beta|
                const double den = aa + bb;                        // alpha - beta = sign(alpha) (|alpha| + |beta|)
                double id = __drcp_rn(den);                        // correctly rounded reciprocal, no division slow path
                scal = copysign(id, alpha);
                if (!(sq > 1e-280 && sq < 1e280)) {                // out of the safe range of the fast path
                    beta = -copysign(sqrt(sq), alpha);
                    tau = (beta - alpha) / beta;
                    scal = 1.0 / (alpha - beta);
                }
            } else {
                beta = -copysign(sqrt(abs2_(alpha) + xnorm2), real_(alpha));
                tau = mk<T>((beta - real_(alpha)) / beta, -imag_(alpha) / beta);
                scal = mk<T>(1.0) / (alpha - mk<T>(beta));
            }
        }
        SP_MARK(4);   // scalar math
        TR(kpos - 1, 9);
        // CS > 1: stage locally, then ONE bulk DSMEM copy per peer (async proxy) that completes on the peer's
        // mbarrier -- no per-element remote stores, no cluster barrier
        T* const dstv = (CS > 1) ? stage_v + (size_t)par * VLEN : vcand + ((size_t)par * CS + rank) * VLEN;
#pragma unroll
        for (int t = 0; t < RPL; ++t) {
            const int r = lane + 32 * t;
            if (r < n && r >= rz) {
                T v = mk<T>(0.0);
                if (r > kpos) v = x[t] * scal;
                else if (r == kpos) v = mk<T>(1.0);
                dstv[r] = v;
            }
        }
        if (lane == 0) {
            RMeta<T> m;
            m.key = ck; m.beta = beta; m.tau = tau;
            *slot_meta(dstv) = m;
        }
        if constexpr (CS > 1) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < CS; ++q)
                    dsmem_bulk_copy(vcand + ((size_t)par * CS + rank) * VLEN + rz, dstv + rz, cbytes, &mbar[par], q);
            }
        }
        TR(kpos - 1, 10);
        SP_MARK(5);   // remote stores
    };

    select_and_publish(0);

#ifdef SLA_LDR_PROFILE
    tq = prof ? clock64() : 0;
#define PROF_MARK(i) do { if (prof) { long long t_ = clock64(); pt[i] += t_ - tq; tq = t_; } } while (0)
#else
#define PROF_MARK(i) do { } while (0)
#endif
    // ------------------------------------------------------------------ pivoted Householder QR
    int park_u = -1;   // column of this warp waiting to be parked
    for (int k = 0; k < n; ++k) {
        const int par = k & 1;
        PROF_MARK(0);
        // the ONLY cluster-wide synchronisation per column: wait until all CS candidates have landed here
        if constexpr (CS > 1) {
            mbar_wait(&mbar[par], (unsigned)((k >> 1) & 1));
            if (tid == 0 && k + 2 < n) mbar_expect_tx(&mbar[par], xbytes(k + 2));   // arm the next phase of this barrier
        } else {
            __syncthreads();
        }
        PROF_MARK(1);
        TR(k, 0);
        int win = 0;
        RMeta<T> wm = *slot_meta(vcand + (size_t)par * CS * VLEN);
#pragma unroll
        for (int q = 1; q < CS; ++q) {
            const RMeta<T> om = *slot_meta(vcand + ((size_t)par * CS + q) * VLEN);   // independent loads: issued together
            if (om.key > wm.key) { wm = om; win = q; }
        }
        const int pc = key_col(wm.key), plc = pc / CS;
        const T tauk = wm.tau;
        const T* vcur = vcand + ((size_t)par * CS + win) * VLEN;
        if (tid == 0) { taus[k] = tauk; dvec[k] = fabs(wm.beta); bsg[k] = wm.beta; }
        // ---- the owner warp retires the pivot column; it is parked (raw) in shared memory LATER, while the
        // warp would otherwise idle at the next mbarrier: retired columns are never touched again
        if (rank == win && warp == (plc % CTA_WARPS)) {
            park_u = slot_of_grp(plc / CTA_WARPS);
            active &= ~(1u << park_u);
        }
        // Deferred duties of column k, done off the critical path (after the pivot scan of column k+1):
        //  * park the retired column: all RPL row tiles, unmasked; the final pass masks rows >= its position
        //  * reflector k -> global (rows > k of column k of L) through the TMA, by thread 0 of the CTA that
        //    won, from its own staging copy (rewritten only after the next block barrier)
        auto deferred = [&](int kd, bool won) {
            if (park_u >= 0) {
                if constexpr (SPW == 0) {
                    T* dst = Rst + (size_t)(warp + CTA_WARPS * park_u) * NMAX;
#pragma unroll
                    for (int u = 0; u < CPW; ++u)
                        if (park_u == u) {
#pragma unroll
                            for (int t = 0; t < RPL; ++t) dst[lane + 32 * t] = a[u][t];
                        }
                } else if (park_u < CPW) {
                    // hybrid: a retired register column goes (raw) to its column of R in global memory; the final
                    // pass re-reads it with the same lane <-> row mapping.  Shared columns simply stay where they are.
                    T* dst = p.R + (long long)b * p.strideR + (long long)colof(park_u) * p.ldr;
#pragma unroll
                    for (int u = 0; u < CPW; ++u)
                        if (park_u == u) {
#pragma unroll
                            for (int t = 0; t < RPL; ++t) if (lane + 32 * t < n) dst[lane + 32 * t] = a[u][t];
                        }
                }
                if (lane == 0) cpos[warp + CTA_WARPS * park_u] = kd;
                park_u = -1;
            }
            if (won && kd + 1 < n) {
                const T* src = (CS > 1) ? stage_v + (size_t)(kd & 1) * VLEN : vcand + (size_t)(kd & 1) * VLEN;
                const int r0 = (kd + 1) & ~(EPV - 1);
                if (bulk_ok) {
                    if (tid == 0) tma_store(Lg + (long long)kd * ldl + r0, src + r0, (unsigned)((n - r0) * sizeof(T)));
                } else {
                    for (int r = kd + 1 + tid; r < n; r += CTA_THREADS) Lg[(long long)kd * ldl + r] = src[r];
                }
            }
        };
        if (k == n - 1) { deferred(k, false); break; }
        const bool won = (rank == win);
        TR(k, 1);
        PROF_MARK(2);
        // ---- pass: w = conj(tau) v^H a ; a -= v w ; down-date the squared norms.  Register columns never leave
        // the registers; the shared columns of the hybrid shape stream through 16-byte thread-private loads/stores
        double temp = 1.0;
        if (active) {
            const int tlo = k >> 5;   // row tiles below this hold only rows < k: v is zero there -> skipped
            T v[RPL];
            T dot[TOT];
            T akm = mk<T>(0.0);       // element (k, column myu) before the update
#pragma unroll
            for (int u = 0; u < TOT; ++u) dot[u] = mk<T>(0.0);
            T myw;
            // straight-line code (all loads first, full ILP) over the row tiles [T0, RPL), one variant per
            // pair of tiles so that tiles holding only rows < k are skipped; retired columns are skipped too
            // (warp-uniform branches), so the FP64 work shrinks with the active (n-k) x (n-k) block
#define SLA_PASS(T0)                                                                                   \
    {                                                                                                  \
        _Pragma("unroll") for (int t = (T0); t < RPL; ++t)                                             \
            v[t] = (lane + 32 * t < n) ? vcur[lane + 32 * t] : mk<T>(0.0);                             \
        _Pragma("unroll") for (int t = (T0); t < RPL && t < (T0) + 2; ++t)                             \
            if (t == tlo) {                                                                            \
                if constexpr (SPW > 0 && CPW == 4) {                                                   \
                    /* register slots: the owner lane of row k drops its 4 values into the warp's scratch (2 stores +  \
                       1 load instead of 8 shuffles); shared slots: straight from the owner lane's cell */               \
                    if (lane == (k & 31)) {                                                            \
                        *reinterpret_cast<double2*>(wsc + 8) = make_double2(real_(a[0][t]), real_(a[1][t]));  \
                        *reinterpret_cast<double2*>(wsc + 10) = make_double2(real_(a[2][t]), real_(a[3][t])); \
                    }                                                                                  \
                    __syncwarp();                                                                      \
                    if (myu < CPW) akm = mk<T>(wsc[8 + myu]);                                          \
                    else akm = swarp[(k & 31) * 2 + ((myu - CPW) * (RPL / 2) + (t >> 1)) * 64 + (t & 1)]; \
                } else {                                                                               \
                    _Pragma("unroll") for (int u = 0; u < CPW; ++u) {                                  \
                        const T e_ = shb(a[u][t], k & 31);                                             \
                        if (myu == u) akm = e_;                                                        \
                    }                                                                                  \
                }                                                                                      \
            }                                                                                          \
        _Pragma("unroll") for (int u = 0; u < CPW; ++u) {                                              \
            if ((active >> u) & 1u) {                                                                  \
                SLA_KEEP_BRANCH;                                                                       \
                T d2_ = mk<T>(0.0);                                                                    \
                _Pragma("unroll") for (int t = (T0); t < RPL; t += 2) {                                \
                    fmacc(dot[u], conj_(v[t]), a[u][t]);                                               \
                    if (t + 1 < RPL) fmacc(d2_, conj_(v[t + 1]), a[u][t + 1]);                         \
                }                                                                                      \
                dot[u] = dot[u] + d2_;                                                                 \
            }                                                                                          \
        }                                                                                              \
        if constexpr (SPW > 0) {                                                                       \
            _Pragma("unroll") for (int su = 0; su < SPW; ++su) {                                       \
                if ((active >> (CPW + su)) & 1u) {                                                     \
                    SLA_KEEP_BRANCH;                                                                   \
                    double d2_ = 0.0;                                                                  \
                    _Pragma("unroll") for (int t = (T0); t < RPL; t += 2) {                            \
                        const double2 x2_ = *reinterpret_cast<const double2*>(&scol[(su * (RPL / 2) + (t >> 1)) * 64]); \
                        dot[CPW + su] = fma(real_(v[t]), x2_.x, real_(dot[CPW + su]));                 \
                        d2_ = fma(real_(v[t + 1]), x2_.y, d2_);                                        \
                    }                                                                                  \
                    dot[CPW + su] = dot[CPW + su] + mk<T>(d2_);                                        \
                }                                                                                      \
            }                                                                                          \
        }                                                                                              \
        TR(k, 2);                                                                                      \
        myw = conj_(tauk) * reduce_transposed<T, TOT>(dot, lane);                                      \
        TR(k, 3);                                                                                      \
        [[maybe_unused]] double wv_[8];                                                                \
        if constexpr (SPW > 0) {   /* w of the 8 columns through a per-warp scratch: 1 store + 4 loads, not 16 shuffles */ \
            if ((lane % FLS) == 0) wsc[lane / FLS] = real_(myw);                                       \
            __syncwarp();                                                                              \
            _Pragma("unroll") for (int i_ = 0; i_ < 4; ++i_) {                                         \
                const double2 w2_ = *reinterpret_cast<const double2*>(wsc + 2 * i_);                   \
                wv_[2 * i_] = w2_.x; wv_[2 * i_ + 1] = w2_.y;                                          \
            }                                                                                          \
            __syncwarp();                                                                              \
        }                                                                                              \
        _Pragma("unroll") for (int u = 0; u < CPW; ++u) {                                              \
            T wu;                                                                                      \
            if constexpr (SPW > 0) wu = mk<T>(wv_[u]); else wu = shb(myw, u * FLS);                    \
            if ((active >> u) & 1u) {                                                                  \
                SLA_KEEP_BRANCH;                                                                       \
                _Pragma("unroll") for (int t = (T0); t < RPL; ++t) fms(a[u][t], v[t], wu);             \
            }                                                                                          \
        }                                                                                              \
        if constexpr (SPW > 0) {                                                                       \
            _Pragma("unroll") for (int su = 0; su < SPW; ++su) {                                       \
                const double wu = wv_[CPW + su];                                                       \
                if ((active >> (CPW + su)) & 1u) {                                                     \
                    SLA_KEEP_BRANCH;                                                                   \
                    _Pragma("unroll") for (int t = (T0); t < RPL; t += 2) {                            \
                        double2* q_ = reinterpret_cast<double2*>(&scol[(su * (RPL / 2) + (t >> 1)) * 64]); \
                        double2 x2_ = *q_;                                                             \
                        x2_.x = fma(-real_(v[t]), wu, x2_.x);                                          \
                        x2_.y = fma(-real_(v[t + 1]), wu, x2_.y);                                      \
                        *q_ = x2_;                                                                     \
                    }                                                                                  \
                }                                                                                      \
            }                                                                                          \
        }                                                                                              \
    }
            if (RPL > 6 && tlo >= 6) SLA_PASS(6)
            else if (RPL > 4 && tlo >= 4) SLA_PASS(4)
            else if (RPL > 2 && tlo >= 2) SLA_PASS(2)
            else SLA_PASS(0)
#undef SLA_PASS
            TR(k, 4);
            // element (k, column myu) after the update is a_k - w (v_k = 1): down-date the squared norm
            const bool cond = ((active >> myu) & 1u) && vn1 != 0.0;
            bool flagged = false;
            if (cond) {
                temp = fmax(0.0, fma(-abs2_(akm - myw), rvn1, 1.0));
                flagged = temp * q12 <= TOL3Z_R;
                if (!flagged) vn1 *= temp;
            }
            unsigned fl = __ballot_sync(0xffffffffu, flagged);
            while (fl) {   // exact squared norm of a flagged (already updated) column below row k
                const int uf = (__ffs(fl) - 1) / FLS;
                fl &= ~((FLS == 32 ? 0xffffffffu : ((1u << FLS) - 1u)) << (uf * FLS));
                double ss = 0.0;
#pragma unroll
                for (int u = 0; u < CPW; ++u)
                    if (uf == u) {
#pragma unroll
                        for (int t = 0; t < RPL; ++t) if (lane + 32 * t > k) ss += abs2_(a[u][t]);
                    }
                if constexpr (SPW > 0) {
                    if (uf >= CPW) {
#pragma unroll
                        for (int t = 0; t < RPL; ++t) if (lane + 32 * t > k) ss += abs2_(sld(uf - CPW, t));
                    }
                }
                ss = warp_sum(ss);
                if (myu == uf) { vn1 = ss; q12 = 1.0; temp = 1.0; }
            }
        }
        PROF_MARK(3);
        select_and_publish(k + 1);
        // off the critical path: the ratio and reciprocal the next down-dating needs, then the deferred duties
        q12 *= temp;
        rvn1 = 1.0 / vn1;
        deferred(k, won);
    }
    PROF_MARK(4);
    if (tid == 0) tma_wait_all();   // all reflectors have landed in global memory
    __syncthreads();

    // ------------------------------------------------------------------ d, R (original column order)
    {
        T* Rg = p.R + (long long)b * p.strideR;
#pragma unroll
        for (int u = 0; u < TOT; ++u) {   // the warp's own columns, same lane <-> row mapping as the parking
            const int lc = warp + CTA_WARPS * u;   // slot index (parking area, cpos)
            const int c = colof(u);
            if (c >= n) continue;
            T* dst = Rg + (long long)c * p.ldr;
            const int kp = cpos[lc];      // R'(:, kp) = [parked rows < kp ; beta_kp ; 0]
            const double bk = bsg[kp];
#pragma unroll
            for (int t = 0; t < RPL; ++t) {
                const int r = lane + 32 * t;
                if (r >= n) continue;
                T raw;
                if constexpr (SPW == 0) raw = Rst[(size_t)lc * NMAX + r];
                else raw = (u < CPW) ? dst[r] : sld(u - CPW, t);
                const T val = (r < kp) ? raw : mk<T>(r == kp ? bk : 0.0);
                dst[r] = val / dvec[r];
            }
        }
        if (rank == 0) {
            double* d = p.d + (long long)b * p.stride_d;
            for (int r = tid; r < n; r += CTA_THREADS) d[r] = dvec[r];
        }
    }
    if constexpr (MODE == 1) {
        if (rank == 0)
            for (int r = tid; r < n; r += CTA_THREADS) tau_g[r] = taus[r];
    }
    __threadfence();
    cl_barrier<CS>();   // every CTA's reflectors are in global memory and visible; parked columns consumed
    PROF_MARK(5);
    }   // MODE != 2
    if constexpr (MODE == 1) return;
    if constexpr (MODE == 2) {
        for (int r = tid; r < n; r += CTA_THREADS) taus[r] = tau_g[r];
        __syncthreads();
    }

    // ------------------------------------------------------------------ Q = H_0 ... H_{n-1} applied to I
    // (column c of Q is touched by the reflectors k <= c only: the shared slots hold the low columns, see colof)
#pragma unroll
    for (int u = 0; u < TOT; ++u) {
        const int c = colof(u);
#pragma unroll
        for (int t = 0; t < RPL; ++t) {
            const T e = mk<T>((lane + 32 * t == c && c < n) ? 1.0 : 0.0);
            if (u < CPW) a[u < CPW ? u : 0][t] = e;
            else sst(u - CPW, t, e);
        }
    }
    constexpr bool QBLK = (MODE == 2) && !CP && CS == 4 && RPL == 8 && CPW == 4 && SPW == 0;
    [[maybe_unused]] double* const gq = reinterpret_cast<double*>(smem_raw + Lay::offKeys);          // Gram entries [2][6]
    [[maybe_unused]] double* const ysc = reinterpret_cast<double*>(smem_raw + Lay::offStageV) + warp * 16;   // per-warp z / y
    {
        T* const ring = vcand;   // [2][QKC][NMAX]
        const int nchunk = (n + QKC_R - 1) / QKC_R;
        const bool vecq = (((size_t)ldl * sizeof(T)) % 16 == 0) && ((((uintptr_t)Lg) & 15) == 0) && (CP || (n % 2 == 0));
        auto load_chunk = [&](int ch, int buf) {
            T* dstb = ring + (size_t)buf * QKC_R * NMAX;
            // 64 threads per reflector, 16-byte pieces (no integer division in the index math)
            static_assert(CTA_THREADS == 64 * QKC_R, "one 64-thread group per reflector of a chunk");
            const int i = tid >> 6, kk = ch * QKC_R + i;
            if (kk < n) {
                for (int r = (tid & 63) * EPV; r < n; r += 64 * EPV) {
                    T* dst = dstb + (size_t)i * NMAX + r;
                    if (vecq) {
                        cp_async16(dst, Lg + (long long)kk * ldl + r, 16);
                    } else {
                        for (int t = 0; t < EPV; ++t) if (r + t < n) dst[t] = Lg[(long long)kk * ldl + r + t];
                    }
                }
            }
            cp_async_commit();
        };
        load_chunk(nchunk - 1, (nchunk - 1) & 1);
        for (int ch = nchunk - 1; ch >= 0; --ch) {
            if (ch > 0) load_chunk(ch - 1, (ch - 1) & 1);
            if (ch > 0) cp_async_wait<1>(); else cp_async_wait<0>();
            __syncthreads();
            const T* vb = ring + (size_t)(ch & 1) * QKC_R * NMAX;
            const int k0 = ch * QKC_R;
            const int kcnt = min(QKC_R, n - k0);
            // Blocked application (Q-only launches of the register shape): four reflectors at a time in compact-WY
            // form, Q <- Q - V (T (V^T Q)): ONE transposed reduction per four reflectors instead of four.
            [[maybe_unused]] int nq = 0;
            if constexpr (QBLK) {
                nq = kcnt / 4;
                T* vw = ring + (size_t)(ch & 1) * QKC_R * NMAX;
                const int rlo = ((k0 >> 5) & ~1) * 32;
                // (a) masks in place: rows < k -> 0, row k -> 1, rows >= n -> 0 (only rows >= rlo are ever read)
                if ((tid >> 6) < kcnt) {
                    const int i = tid >> 6, k = k0 + i;
                    for (int r = rlo + (tid & 63); r <= k; r += 64) vw[(size_t)i * NMAX + r] = mk<T>(r == k ? 1.0 : 0.0);
                    for (int r = n + (tid & 63); r < NMAX; r += 64) vw[(size_t)i * NMAX + r] = mk<T>(0.0);
                }
                // (b) Gram entries v_i^T v_j (i < j) of every quad, one warp each (masks applied on the fly)
                if (warp < 6 * nq) {
                    const int q = warp / 6, pi = warp % 6;
                    const int jj = (pi < 1) ? 1 : (pi < 3 ? 2 : 3), ii = pi - (jj == 1 ? 0 : (jj == 2 ? 1 : 3));
                    const int kj = k0 + 4 * q + jj;
                    const T* vi = vb + (size_t)(4 * q + ii) * NMAX;
                    const T* vj = vb + (size_t)(4 * q + jj) * NMAX;
                    double g = 0.0;
#pragma unroll
                    for (int t = 0; t < RPL; ++t) {
                        const int r = lane + 32 * t;
                        if (r >= kj && r < n) g = fma(real_(vi[r]), (r == kj) ? 1.0 : real_(vj[r]), g);
                    }
                    g = warp_sum(g);
                    if (lane == 0) gq[q * 6 + pi] = g;
                }
                __syncthreads();
            }
            // columns left of the whole chunk are untouched by it
            unsigned qact = 0;
#pragma unroll
            for (int u = 0; u < TOT; ++u) {
                const int c = colof(u);
                if ((c < n) && (c >= k0)) qact |= 1u << u;
            }
            if (qact) {
                const int tlo = k0 >> 5;
                for (int i = kcnt - 1; i >= (QBLK ? 4 * nq : 0); --i) {
                    const int k = k0 + i;
                    const T* vk = vb + (size_t)i * NMAX;
                    const T tk = taus[k];
                    T v[RPL];
                    T dot[TOT];
#pragma unroll
                    for (int u = 0; u < TOT; ++u) dot[u] = mk<T>(0.0);
#define SLA_QPASS(T0)                                                                                  \
    {                                                                                                  \
        _Pragma("unroll") for (int t = (T0); t < RPL; ++t) {                                           \
            const int r_ = lane + 32 * t;                                                              \
            v[t] = (r_ < n && r_ > k) ? vk[r_] : mk<T>(r_ == k ? 1.0 : 0.0);                           \
        }                                                                                              \
        _Pragma("unroll") for (int u = 0; u < CPW; ++u) {                                              \
            if ((qact >> u) & 1u) {                                                                    \
                SLA_KEEP_BRANCH;                                                                       \
                T d2_ = mk<T>(0.0);                                                                    \
                _Pragma("unroll") for (int t = (T0); t < RPL; t += 2) {                                \
                    fmacc(dot[u], conj_(v[t]), a[u][t]);                                               \
                    if (t + 1 < RPL) fmacc(d2_, conj_(v[t + 1]), a[u][t + 1]);                         \
                }                                                                                      \
                dot[u] = dot[u] + d2_;                                                                 \
            }                                                                                          \
        }                                                                                              \
        if constexpr (SPW > 0) {                                                                       \
            _Pragma("unroll") for (int su = 0; su < SPW; ++su) {                                       \
                if ((qact >> (CPW + su)) & 1u) {                                                       \
                    SLA_KEEP_BRANCH;                                                                   \
                    double d2_ = 0.0;                                                                  \
                    _Pragma("unroll") for (int t = (T0); t < RPL; t += 2) {                            \
                        const double2 x2_ = *reinterpret_cast<const double2*>(&scol[(su * (RPL / 2) + (t >> 1)) * 64]); \
                        dot[CPW + su] = fma(real_(v[t]), x2_.x, real_(dot[CPW + su]));                 \
                        d2_ = fma(real_(v[t + 1]), x2_.y, d2_);                                        \
                    }                                                                                  \
                    dot[CPW + su] = dot[CPW + su] + mk<T>(d2_);                                        \
                }                                                                                      \
            }                                                                                          \
        }                                                                                              \
        const T mydot = reduce_transposed<T, TOT>(dot, lane);                                          \
        const T myw = tk * mydot;                                                                      \
        [[maybe_unused]] double wv_[8];                                                                \
        if constexpr (SPW > 0) {   /* w of the 8 columns through a per-warp scratch: 1 store + 4 loads, not 16 shuffles */ \
            if ((lane % FLS) == 0) wsc[lane / FLS] = real_(myw);                                       \
            __syncwarp();                                                                              \
            _Pragma("unroll") for (int i_ = 0; i_ < 4; ++i_) {                                         \
                const double2 w2_ = *reinterpret_cast<const double2*>(wsc + 2 * i_);                   \
                wv_[2 * i_] = w2_.x; wv_[2 * i_ + 1] = w2_.y;                                          \
            }                                                                                          \
            __syncwarp();                                                                              \
        }                                                                                              \
        _Pragma("unroll") for (int u = 0; u < CPW; ++u) {                                              \
            T wu;                                                                                      \
            if constexpr (SPW > 0) wu = mk<T>(wv_[u]); else wu = shb(myw, u * FLS);                    \
            if ((qact >> u) & 1u) {                                                                    \
                SLA_KEEP_BRANCH;                                                                       \
                _Pragma("unroll") for (int t = (T0); t < RPL; ++t) fms(a[u][t], v[t], wu);             \
            }                                                                                          \
        }                                                                                              \
        if constexpr (SPW > 0) {                                                                       \
            _Pragma("unroll") for (int su = 0; su < SPW; ++su) {                                       \
                const double wu = wv_[CPW + su];                                                       \
                if ((qact >> (CPW + su)) & 1u) {                                                       \
                    SLA_KEEP_BRANCH;                                                                   \
                    _Pragma("unroll") for (int t = (T0); t < RPL; t += 2) {                            \
                        double2* q_ = reinterpret_cast<double2*>(&scol[(su * (RPL / 2) + (t >> 1)) * 64]); \
                        double2 x2_ = *q_;                                                             \
                        x2_.x = fma(-real_(v[t]), wu, x2_.x);                                          \
                        x2_.y = fma(-real_(v[t + 1]), wu, x2_.y);                                      \
                        *q_ = x2_;                                                                     \
                    }                                                                                  \
                }                                                                                      \
            }                                                                                          \
        }                                                                                              \
    }
                    if (RPL > 6 && tlo >= 6) SLA_QPASS(6)
                    else if (RPL > 4 && tlo >= 4) SLA_QPASS(4)
                    else if (RPL > 2 && tlo >= 2) SLA_QPASS(2)
                    else SLA_QPASS(0)
#undef SLA_QPASS
                }
                if constexpr (QBLK) {
                    for (int q = nq - 1; q >= 0; --q) {
                        const int kq = k0 + 4 * q;
                        const double* v0p = reinterpret_cast<const double*>(vb) + (size_t)(4 * q) * NMAX + lane;
                        // columns u < M of this warp lie left of the chunk (c < 64 M <= k0): skipped with their tiles
#define SLA_QBLOCK(M)                                                                                  \
    {                                                                                                  \
        double z[16];                                                                                  \
        _Pragma("unroll") for (int i_ = 0; i_ < 16; ++i_) z[i_] = 0.0;                                 \
        _Pragma("unroll") for (int t = 2 * (M); t < RPL; ++t) {                                        \
            const double w0 = v0p[32 * t], w1 = v0p[NMAX + 32 * t], w2 = v0p[2 * NMAX + 32 * t],       \
                         w3 = v0p[3 * NMAX + 32 * t];                                                  \
            _Pragma("unroll") for (int u = (M); u < 4; ++u) {                                          \
                z[4 * u + 0] = fma(w0, a[u][t], z[4 * u + 0]);                                         \
                z[4 * u + 1] = fma(w1, a[u][t], z[4 * u + 1]);                                         \
                z[4 * u + 2] = fma(w2, a[u][t], z[4 * u + 2]);                                         \
                z[4 * u + 3] = fma(w3, a[u][t], z[4 * u + 3]);                                         \
            }                                                                                          \
        }                                                                                              \
        const double zz = reduce_transposed<double, 16>(z, lane);   /* lanes 2g, 2g+1: z_g, g = 4u + j */ \
        if (!(lane & 1)) ysc[lane >> 1] = zz;                                                          \
        /* T of the quad (forward, column-wise ?larft) from the Gram entries and the taus */           \
        const double t0 = real_(taus[kq]), t1 = real_(taus[kq + 1]), t2 = real_(taus[kq + 2]),          \
                     t3 = real_(taus[kq + 3]);                                                         \
        const double g01 = gq[q * 6 + 0], g02 = gq[q * 6 + 1], g12 = gq[q * 6 + 2], g03 = gq[q * 6 + 3], \
                     g13 = gq[q * 6 + 4], g23 = gq[q * 6 + 5];                                         \
        const double T01 = -t1 * (t0 * g01);                                                           \
        const double T12 = -t2 * (t1 * g12), T02 = -t2 * (t0 * g02 + T01 * g12);                       \
        const double T23 = -t3 * (t2 * g23), T13 = -t3 * (t1 * g13 + T12 * g23),                       \
                     T03 = -t3 * (t0 * g03 + T01 * g13 + T02 * g23);                                   \
        __syncwarp();                                                                                  \
        {   /* y = T z for the column of this lane's 8-lane group; its first lane publishes it */      \
            const int ug = lane >> 3;                                                                  \
            const double2 za = *reinterpret_cast<const double2*>(ysc + 4 * ug);                        \
            const double2 zb = *reinterpret_cast<const double2*>(ysc + 4 * ug + 2);                    \
            const double y3 = t3 * zb.y;                                                               \
            const double y2 = fma(T23, zb.y, t2 * zb.x);                                               \
            const double y1 = fma(T13, zb.y, fma(T12, zb.x, t1 * za.y));                               \
            const double y0 = fma(T03, zb.y, fma(T02, zb.x, fma(T01, za.y, t0 * za.x)));               \
            __syncwarp();                                                                              \
            if ((lane & 7) == 0) {                                                                     \
                *reinterpret_cast<double2*>(ysc + 4 * ug) = make_double2(y0, y1);                      \
                *reinterpret_cast<double2*>(ysc + 4 * ug + 2) = make_double2(y2, y3);                  \
            }                                                                                          \
            __syncwarp();                                                                              \
        }                                                                                              \
        double y[16];                                                                                  \
        _Pragma("unroll") for (int u = (M); u < 4; ++u) {                                              \
            const double2 ya = *reinterpret_cast<const double2*>(ysc + 4 * u);                         \
            const double2 yb = *reinterpret_cast<const double2*>(ysc + 4 * u + 2);                     \
            y[4 * u] = ya.x; y[4 * u + 1] = ya.y; y[4 * u + 2] = yb.x; y[4 * u + 3] = yb.y;            \
        }                                                                                              \
        __syncwarp();                                                                                  \
        _Pragma("unroll") for (int t = 2 * (M); t < RPL; ++t) {                                        \
            const double w0 = v0p[32 * t], w1 = v0p[NMAX + 32 * t], w2 = v0p[2 * NMAX + 32 * t],       \
                         w3 = v0p[3 * NMAX + 32 * t];                                                  \
            _Pragma("unroll") for (int u = (M); u < 4; ++u) {                                          \
                double acc_ = a[u][t];                                                                 \
                acc_ = fma(-w0, y[4 * u + 0], acc_);                                                   \
                acc_ = fma(-w1, y[4 * u + 1], acc_);                                                   \
                acc_ = fma(-w2, y[4 * u + 2], acc_);                                                   \
                acc_ = fma(-w3, y[4 * u + 3], acc_);                                                   \
                a[u][t] = acc_;                                                                        \
            }                                                                                          \
        }                                                                                              \
    }
                        if (tlo >= 6) SLA_QBLOCK(3)
                        else if (tlo >= 4) SLA_QBLOCK(2)
                        else if (tlo >= 2) SLA_QBLOCK(1)
                        else SLA_QBLOCK(0)
#undef SLA_QBLOCK
                    }
                }
            }
            __syncthreads();
        }
    }
    PROF_MARK(6);
    cl_barrier<CS>();   // every CTA is done reading reflectors from Lg before anyone overwrites them
#ifdef SLA_LDR_PROFILE
    if (prof && (lane == 0) && (warp == 0 || warp == 1))
        for (int i = 0; i < 8; ++i) p.prof[(rank * 2 + warp) * 8 + i] = pt[i];
#endif
#undef PROF_MARK
#pragma unroll
    for (int u = 0; u < TOT; ++u) {
        const int c = colof(u);
        if (c < n) {
            T* dst = Lg + (long long)c * ldl;
#pragma unroll
            for (int t = 0; t < RPL; ++t) {
                const int r = lane + 32 * t;
                if (r < n) {
                    if (u < CPW) dst[r] = a[u < CPW ? u : 0][t];
                    else dst[r] = sld(u - CPW, t);
                }
            }
        }
    }
}

// ---- host side --------------------------------------------------------------------------------------
template <typename T, int CS, int RPL, int CPW, int SPW = 0, int MODE = 0> static cudaError_t launch_reg(const LdrArgs<T>& p, cudaStream_t stream) {
    auto kern = ldr_reg_kernel<T, CS, RPL, CPW, SPW, MODE>;
    const size_t smem = RegLayout<T, RPL, CPW, CS, SPW>::total;
    if (smem > 227 * 1024) return cudaErrorNotSupported;
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

// how many clusters of this shape the device holds at once (GPC fragmentation included); cached by the callers
template <typename T, int CS, int RPL, int CPW, int SPW = 0> static int max_clusters_of() {
    auto kern = ldr_reg_kernel<T, CS, RPL, CPW, SPW, 0>;
    const size_t smem = RegLayout<T, RPL, CPW, CS, SPW>::total;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS * 64, 1, 1);
    cfg.blockDim = dim3(CTA_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return nc;
}

// register tile shapes: real 32 doubles per lane, complex 16 complex per lane
template <typename T> struct RegShape;
template <> struct RegShape<double> {
    // n <= 64: RPL 2, CPW 4 (CS 1) | n <= 128: RPL 4, CPW 8 (CS 1) | n <= 256: RPL 8, CPW 4 (CS = ceil(n/64))
    static int cluster(int n) { return n <= 128 ? 1 : (n <= 256 ? (n <= 192 ? 4 : 4) : 0); }
};
template <> struct RegShape<cplx> {
    // n <= 64: RPL 2, CPW 4 (CS 1) | n <= 128: RPL 4, CPW 4 (CS 2) | n <= 256: RPL 8, CPW 2 (CS 8)
    static int cluster(int n) { return n <= 64 ? 1 : (n <= 128 ? 2 : (n <= 256 ? 8 : 0)); }
};

template <typename T> int ldr_reg_cluster_size(int n) { return RegShape<T>::cluster(n); }

template <> cudaError_t ldr_reg_batched<double>(const LdrArgs<double>& p, cudaStream_t s) {
    if (p.batch <= 0 || p.n <= 0) return cudaSuccess;
    if (p.n <= 64) return launch_reg<double, 1, 2, 4>(p, s);
    if (p.n <= 128) return launch_reg<double, 1, 4, 8>(p, s);
    if (p.n <= 256) {
        // 128 < n <= 256: factorise (MODE 1), then form Q on the tensor cores (sla_orgq.cu); SLA_Q_DMMA=0 or no tau
        // workspace: everything in one launch (MODE 0)
        static const bool qdmma = !(getenv("SLA_Q_DMMA") && atoi(getenv("SLA_Q_DMMA")) == 0);
        if (qdmma && p.Twork != nullptr && p.strideT >= p.n) {
            cudaError_t e = launch_reg<double, 4, 8, 4, 0, 1>(p, s);
            if (e != cudaSuccess) return e;
            return orgq_dmma_batched(p, s);
        }
        return launch_reg<double, 4, 8, 4>(p, s);
    }
    return cudaErrorNotSupported;
}
template <> cudaError_t ldr_reg_batched<cplx>(const LdrArgs<cplx>& p, cudaStream_t s) {
    if (p.batch <= 0 || p.n <= 0) return cudaSuccess;
    if (p.n <= 64) return launch_reg<cplx, 1, 2, 4>(p, s);
    if (p.n <= 128) return launch_reg<cplx, 2, 4, 4>(p, s);
    if (p.n <= 256) return launch_reg<cplx, 8, 8, 2>(p, s);
    return cudaErrorNotSupported;
}

// ---- hybrid shape (real, 128 < n <= 256): clusters of 2, half of the matrix in registers, half in shared memory
template <> cudaError_t ldr_hyb_batched<double>(const LdrArgs<double>& p, cudaStream_t s) {
    if (p.batch <= 0 || p.n <= 0) return cudaSuccess;
    if (p.n > 256) return cudaErrorNotSupported;
    // factorise with the hybrid shape (one wave), then form Q with the register shape, four reflectors per
    // reduction (compact WY): 1.12 -> 1.08 ms for 64 matrices.  SLA_HYB_SPLITQ=0: everything in the hybrid launch.
    static const bool splitq = !(getenv("SLA_HYB_SPLITQ") && atoi(getenv("SLA_HYB_SPLITQ")) == 0);
    if (splitq && p.Twork != nullptr && p.strideT >= p.n) {
        cudaError_t e = launch_reg<double, 2, 8, 4, 4, 1>(p, s);
        if (e != cudaSuccess) return e;
        // Q on the FP64 tensor cores (sla_orgq.cu: 0.215 ms for 64 matrices); SLA_Q_DMMA=0: the vector-unit MODE 2
        // launch of this kernel (0.275 ms)
        static const bool qdmma = !(getenv("SLA_Q_DMMA") && atoi(getenv("SLA_Q_DMMA")) == 0);
        if (qdmma) return orgq_dmma_batched(p, s);
        return launch_reg<double, 4, 8, 4, 0, 2>(p, s);
    }
    return launch_reg<double, 2, 8, 4, 4>(p, s);
}
template <> cudaError_t ldr_hyb_batched<cplx>(const LdrArgs<cplx>&, cudaStream_t) { return cudaErrorNotSupported; }
template <> int ldr_hyb_cluster_size<double>(int n) { return (n > 128 && n <= 256) ? 2 : 0; }
template <> int ldr_hyb_cluster_size<cplx>(int) { return 0; }
// resident matrices of the N <= 256 real shapes: {register kernel (clusters of 4), hybrid (clusters of 2)}
void ldr_n256_resident(int* reg_clusters, int* hyb_clusters) {
    static const int r = max_clusters_of<double, 4, 8, 4>(), h = max_clusters_of<double, 2, 8, 4, 4>();
    *reg_clusters = r;
    *hyb_clusters = h;
}

template int ldr_reg_cluster_size<double>(int);
template int ldr_reg_cluster_size<cplx>(int);

}  // namespace sla
