// extern "C" boundary of libsla_b200 (declared in include/sla_b200.h): composes the batched kernels
// (DMMA GEMM with fused diagonal scaling, LDR factorisation, LU/inverse with log|det| and sign,
// element-wise helpers) into the reference's operations.  Every composition cites the reference lines
// it reproduces; scratch use follows the reference's M / M' / M'' roles where that matters for the
// documented aliasing rules.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/sla_b200.h"
#include "sla_capi_f32.h"
#include "sla_kernels.h"

using namespace sla;

constexpr int SLA_N_LDR_ALGOS = 16;          // size of the per-algorithm counters (SLA_LDR_ALGO_* < 16)
static long long* g_ldr_prof = nullptr;   // diagnostics only (sla_debug_set_ldr_prof)

struct sla_handle_s {
    cudaStream_t stream;
    int device = 0;           // the device the handle was created on (sla_create); every call checks it is current
    int last_ldr_algo = 0;    // SLA_LDR_ALGO_* of the most recent LDR factorisation launched through this handle
    unsigned long long ldr_algo_launches[SLA_N_LDR_ALGOS] = {};   // factorisations launched per algorithm
    // fused driver only: a lower-priority side stream on which the group products (independent of the chain
    // state) are built while the chain's latency-bound LDR kernels leave SMs idle; created lazily
    cudaStream_t side = nullptr;
    cudaStream_t side_more[3] = {};   // further side streams: consecutive sub-chunks of group products go round-robin over them
    cudaEvent_t ev_start = nullptr;
    cudaEvent_t ev_chunk[64] = {};
    cudaEvent_t ev_r0 = nullptr, ev_r1 = nullptr;   // small problems: the R product of a stabilisation on its own stream
    int n_events = 0;
    // fused chain driver: the ~100-350 launches of one call captured once into a CUDA graph and replayed while the
    // arguments stay the same (small problems are bound by launch latency, and a multi-GPU host issues 8 x 200 launches
    // per step from one socket); SLA_GRAPH=0 switches it off
    struct ChainGraph {
        bool used = false;
        unsigned long long key[20] = {};
        cudaGraphExec_t exec = nullptr;
        unsigned long long kernels = 0;            // kernel launches inside the graph (sla_launch_count)
        unsigned long long algo_launches[SLA_N_LDR_ALGOS] = {};  // LDR factorisations per algorithm inside the graph
        int last_algo = 0;
    } graphs[4];
    int graph_next = 0;
    bool graph_broken = false;                     // a capture failed once on this handle: direct launches from then on
};

namespace {

#define CK(expr)                                            \
    do {                                                    \
        cudaError_t e__ = (expr);                           \
        if (e__ != cudaSuccess) return 1000 + (int)e__;     \
    } while (0)
#define CKI(expr)                 \
    do {                          \
        int r__ = (expr);         \
        if (r__ != 0) return r__; \
    } while (0)

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

constexpr int NSCAL = 8;  // per-matrix scalar slots in a workspace

// Workspace blob = LDRWorkspace (src/LDR.jl:51-70): M, M', M'' plus the LU/QR scratch and scalars.
template <typename T> struct Ws {
    T *M, *Mp, *Mpp, *X, *Tw;
    double *v1, *v2;  // n * batch each: prepared scaling multipliers (the reference's ws.v)
    double* sc;  // NSCAL * batch
    T* st;       // NSCAL * batch
    int* info;   // batch
    size_t tw_elems;
    static size_t bytes(int n, int batch) {
        size_t nn = (size_t)n * n * batch * sizeof(T);
        size_t b = 4 * align_up(nn);
        b += align_up(ldr_twork_elems<T>(n) * (size_t)batch * sizeof(T));
        b += 2 * align_up((size_t)n * batch * sizeof(double));
        b += align_up((size_t)NSCAL * batch * sizeof(double));
        b += align_up((size_t)NSCAL * batch * sizeof(T));
        b += align_up((size_t)batch * sizeof(int));
        return b;
    }
    Ws(void* blob, int n, int batch) {
        char* p = static_cast<char*>(blob);
        size_t nn = align_up((size_t)n * n * batch * sizeof(T));
        M = (T*)p; p += nn;
        Mp = (T*)p; p += nn;
        Mpp = (T*)p; p += nn;
        X = (T*)p; p += nn;
        tw_elems = ldr_twork_elems<T>(n);
        Tw = (T*)p; p += align_up(tw_elems * (size_t)batch * sizeof(T));
        v1 = (double*)p; p += align_up((size_t)n * batch * sizeof(double));
        v2 = (double*)p; p += align_up((size_t)n * batch * sizeof(double));
        sc = (double*)p; p += align_up((size_t)NSCAL * batch * sizeof(double));
        st = (T*)p; p += align_up((size_t)NSCAL * batch * sizeof(T));
        info = (int*)p;
    }
};

template <typename T> struct Ctx {
    sla_handle_t h;
    cudaStream_t s;
    int n, batch;
    long long nn;  // n*n
    Ws<T> w;
    Ctx(sla_handle_t h_, int n_, int batch_, void* ws) : h(h_), s(h_->stream), n(n_), batch(batch_), nn((long long)n_ * n_), w(ws, n_, batch_) {}

    int copy(T* dst, const T* src) const {
        if (dst == src) return 0;
        CK(cudaMemcpyAsync(dst, src, (size_t)nn * batch * sizeof(T), cudaMemcpyDeviceToDevice, s));
        return 0;
    }
    int copy_d(double* dst, const double* src) const {
        if (dst == src) return 0;
        CK(cudaMemcpyAsync(dst, src, (size_t)n * batch * sizeof(double), cudaMemcpyDeviceToDevice, s));
        return 0;
    }
    // C = rowscale * (op(A) op(B)) * colscale  (+ C)
    int gemm(T* C, int opa, const T* A, long long sA, int opb, const T* B, long long sB, const double* rs = nullptr,
             int rs_mode = SCALE_NONE, const double* cs = nullptr, int cs_mode = SCALE_NONE, int accumulate = 0,
             long long sC = -1) const {
        // the epilogue only multiplies: turn D^-1 / min(D,1) / 1/max(D,1) into multiplier vectors first
        if (rs && rs_mode != SCALE_MUL) {
            CK(scale_prep_batched(w.v1, rs, rs_mode, (long long)n * batch, s));
            rs = w.v1;
        }
        if (cs && cs_mode != SCALE_MUL) {
            CK(scale_prep_batched(w.v2, cs, cs_mode, (long long)n * batch, s));
            cs = w.v2;
        }
        GemmArgs<T> g{};
        g.A = A; g.B = B; g.C = C;
        g.strideA = sA; g.strideB = sB; g.strideC = (sC < 0) ? nn : sC;
        g.lda = g.ldb = g.ldc = n;
        g.M = g.N = g.K = n;
        g.rs = rs; g.stride_rs = n; g.rs_mode = SCALE_MUL; g.rs_div = 0; g.stride_rs_hi = 0;
        g.cs = cs; g.stride_cs = n; g.cs_mode = SCALE_MUL;
        g.accumulate = accumulate;
        g.batch = batch;
        CK(gemm_batched<T>(opa, opb, g, s));
        return 0;
    }
    // ldr! on buffer Lbuf (optionally reading its input from Ain): Q -> Lbuf, d, R0 -> Rout
    int ldr(T* Lbuf, double* d, T* Rout, const T* Ain = nullptr, long long sAin = 0) const {
        LdrArgs<T> a{};
        a.L = Lbuf; a.strideL = nn; a.ldl = n;
        a.Ain = Ain; a.strideAin = sAin; a.ldain = n;
        a.d = d; a.stride_d = n;
        a.R = Rout; a.strideR = nn; a.ldr = n;
        a.Twork = w.Tw; a.strideT = (long long)w.tw_elems;
        a.Qwork = w.X; a.strideQ = nn;   // X is the LU scratch: free during every factorisation
        a.jpvt_out = nullptr;
        a.prof = g_ldr_prof;
        a.n = n; a.batch = batch;
        // register-resident cluster kernel (n <= 256), else the shared-memory cluster kernel (matrix fits the
        // shared memory of <= 8 CTAs), else the L2-streaming kernel.  SLA_LDR_ALGO=reg|cluster|streaming forces one.
        static const int forced = [] {
            const char* e = getenv("SLA_LDR_ALGO");
            if (getenv("SLA_LDR_STREAMING")) return 3;
            if (!e) return 0;
            return e[0] == 'r' ? 1 : (e[0] == 'c' ? 2 : (e[0] == 's' ? 3 : (e[0] == 'h' ? 4 : (e[0] == 'b' ? 5 : (e[0] == 'm' ? 6 : (e[0] == 'p' ? 7 : (e[0] == 't' ? 8 : 0)))))));
        }();
        // SLA_LDR_PIVOT=panel: panel-restricted pivoting (sla_ldr_pp.cu) for every factorisation -- a different pivot
        // sequence than ?geqp3, hence a mode, never chosen by the dispatcher on its own
        static const bool panel_mode = [] { const char* e = getenv("SLA_LDR_PIVOT"); return e && e[0] == 'p'; }();
        // Cluster kernels minimise latency (one matrix per CS SMs, ~132/CS matrices in flight), the streaming
        // kernel maximises throughput (one matrix per SM, 148 in flight) but is ~4x (register kernel) / ~2.5x
        // (shared-memory cluster kernel) slower per wave (profiles/ldr_kernels_r1.txt): pick by wave count.
        const int cs_reg = ldr_reg_cluster_size<T>(n), cs_cl = ldr_cluster_size<T>(n);
        auto waves = [](int batch_, int conc) { return (batch_ + conc - 1) / (conc > 0 ? conc : 1); };
        const int w_stream = waves(batch, 148);
        const bool reg_pays = cs_reg > 0 && waves(batch, 132 / cs_reg) <= 4 * w_stream;
        const bool cl_pays = cs_cl > 0 && 2 * waves(batch, 132 / cs_cl) <= 5 * w_stream;
        // 128 < n <= 256, real: the hybrid shape (half the matrix in shared memory, clusters of 2) holds twice as many
        // matrices.  Measured at N=256: register-shape factorisation wave 0.56 ms, hybrid wave 0.71 ms, Q formation
        // (tensor cores, clusters of 4) 0.095 ms per register-shape wave: t_reg = 0.656 w_reg, t_hyb = 0.71 w_hyb +
        // 0.095 w_reg, i.e. the hybrid path wins when w_hyb < 0.79 w_reg (64 chains: 1 wave instead of 2).
        // SLA_LDR_ALGO=hyb forces it.
        const int cs_hyb = ldr_hyb_cluster_size<T>(n);
        bool hyb_pays = false;
        if (cs_hyb > 0 && forced == 0 && reg_pays) {
            int c_reg = 0, c_hyb = 0;
            ldr_n256_resident(&c_reg, &c_hyb);
            if (c_reg > 0 && c_hyb > 0) hyb_pays = 100 * waves(batch, c_hyb) < 79 * waves(batch, c_reg);
        }
        // a forced algorithm that does not apply to (T, n) is an error, never a silent fall-through to another kernel
        // real, 128 < n <= 256: the blocked kernel (?laqps panel of 4 reflectors, DMMA trailing update, clusters of 2) is
        // correct at every batch size but measured SLOWER than the hybrid shape (0.876 vs 0.688 ms per 64 x N=256
        // factorisation, profiles/ldr_blk_r2.txt: on B200 the DMMA and the FP64 FMA pipe have the same rate, and the lazy
        // panel update lengthens the one-warp candidate chain) -- opt-in: SLA_LDR_BLK=1 or SLA_LDR_ALGO=blk
        static const bool blk_on = getenv("SLA_LDR_BLK") && atoi(getenv("SLA_LDR_BLK")) != 0;
        const int cs_blk = ldr_blk_cluster_size<T>(n);
        // n > 512 (no on-chip kernel) and a batch that leaves SMs idle with one CTA per matrix: the multi-CTA streaming
        // kernel with as many CTAs per matrix as keep the batch in one wave (clusters of 8 / 4 / 2: ~14 / 33 / 66 resident;
        // SLA_LDR_MC=0 switches it off, SLA_LDR_ALGO=mc forces it with SLA_LDR_MC_CS CTAs, default 4)
        static const bool mc_on = !(getenv("SLA_LDR_MC") && atoi(getenv("SLA_LDR_MC")) == 0);
        static const int mc_forced_cs = getenv("SLA_LDR_MC_CS") ? atoi(getenv("SLA_LDR_MC_CS")) : 0;
        int cs_mc = 0;
        if (ldr_mc_supported<T>(n)) {
            if (batch <= 14) cs_mc = 8;
            else if (batch <= 33) cs_mc = 4;
            else if (batch <= 66) cs_mc = 2;
            if (forced == 6) cs_mc = (mc_forced_cs == 2 || mc_forced_cs == 4 || mc_forced_cs == 8) ? mc_forced_cs : 4;
        }
        // n <= 64: one CTA per matrix with the matrix in registers (sla_ldr_small.cu) -- the cluster kernels' per-column
        // machinery costs ~3000 cycles per column at this size, the small kernel's chain ~1300 (114 -> 61 us for one 64 x 64
        // matrix); SLA_LDR_SMALL=0 switches it off (1, 2: its earlier versions), SLA_LDR_ALGO=tiny forces it
        static const bool small_on = !(getenv("SLA_LDR_SMALL") && atoi(getenv("SLA_LDR_SMALL")) == 0);
        int algo;
        if (forced == 7 || (forced == 0 && panel_mode)) { if (!ldr_pp_supported<T>(n)) return SLA_ENOTIMPL; algo = SLA_LDR_ALGO_PANEL; }
        else if (forced == 6) { if (cs_mc == 0) return SLA_ENOTIMPL; algo = SLA_LDR_ALGO_MULTICTA; }
        else if (forced == 0 && mc_on && n > 512 && cs_mc > 0) algo = SLA_LDR_ALGO_MULTICTA;
        else if (forced == 5) { if (cs_blk == 0) return SLA_ENOTIMPL; algo = SLA_LDR_ALGO_BLOCKED; }
        else if (forced == 0 && blk_on && cs_blk > 0) algo = SLA_LDR_ALGO_BLOCKED;
        else if (forced == 4) { if (cs_hyb == 0) return SLA_ENOTIMPL; algo = SLA_LDR_ALGO_HYBRID; }
        else if (forced == 1) { if (cs_reg == 0) return SLA_ENOTIMPL; algo = SLA_LDR_ALGO_REG; }
        else if (forced == 2) { if (cs_cl == 0) return SLA_ENOTIMPL; algo = SLA_LDR_ALGO_CLUSTER; }
        else if (forced == 3) algo = SLA_LDR_ALGO_STREAMING;
        else if (forced == 8) { if (!ldr_small_supported<T>(n)) return SLA_ENOTIMPL; algo = SLA_LDR_ALGO_SMALL; }
        else if (small_on && ldr_small_supported<T>(n)) algo = SLA_LDR_ALGO_SMALL;
        else if (hyb_pays) algo = SLA_LDR_ALGO_HYBRID;
        else if (reg_pays) algo = SLA_LDR_ALGO_REG;
        else if (cl_pays && cs_reg == 0) algo = SLA_LDR_ALGO_CLUSTER;
        else algo = SLA_LDR_ALGO_STREAMING;
        a.info = w.info;
#ifdef SLA_LDR_INTERVALS   // diagnostics: time from "ldr! enqueued behind its input GEMM" to "ldr! finished" on the main stream
        static cudaEvent_t evs[2][256]; static int nev = 0; static bool made = false;
        if (!made) { for (int i = 0; i < 256; ++i) { cudaEventCreate(&evs[0][i]); cudaEventCreate(&evs[1][i]); } made = true; }
        const int ei = nev % 256;
        cudaEventRecord(evs[0][ei], s);
#endif
        switch (algo) {
            case SLA_LDR_ALGO_PANEL: CK(ldr_pp_batched<T>(a, s)); break;
            case SLA_LDR_ALGO_MULTICTA: CK(ldr_mc_batched<T>(a, cs_mc, s)); break;
            case SLA_LDR_ALGO_BLOCKED: CK(ldr_blk_batched<T>(a, s)); break;
            case SLA_LDR_ALGO_HYBRID: CK(ldr_hyb_batched<T>(a, s)); break;
            case SLA_LDR_ALGO_REG: CK(ldr_reg_batched<T>(a, s)); break;
            case SLA_LDR_ALGO_CLUSTER: CK(ldr_cluster_batched<T>(a, s)); break;
            case SLA_LDR_ALGO_SMALL: CK(ldr_small_batched<T>(a, s)); break;
            default: CK(ldr_batched<T>(a, s)); break;
        }
#ifdef SLA_LDR_INTERVALS
        cudaEventRecord(evs[1][ei], s);
        if (++nev % 200 == 0) {
            cudaStreamSynchronize(s);
            float tot = 0.f, mx = 0.f;
            for (int i = 0; i < 200; ++i) { float ms = 0.f; cudaEventElapsedTime(&ms, evs[0][i], evs[1][i]); tot += ms; mx = ms > mx ? ms : mx; }
            fprintf(stderr, "[sla] ldr! intervals on the main stream: mean %.3f ms, max %.3f ms over 200 calls\n", tot / 200, mx);
        }
#endif
        h->last_ldr_algo = algo;
        ++h->ldr_algo_launches[algo];
        return 0;
    }
    // LU of A (destroyed).  invert: A^-1 -> X (and back into A if copy_back)
    int lu(T* A, T* X, int invert, int copy_back, double* logdet, T* sign) const {
        LuArgs<T> a{};
        a.A = A; a.strideA = nn; a.lda = n;
        a.X = X; a.strideX = nn; a.ldx = n;
        a.logdet = logdet; a.sign = sign;
        a.ipiv_out = nullptr; a.info = w.info;
        a.invert = invert; a.copy_back = copy_back;
        a.n = n; a.batch = batch;
        CK(lu_batched<T>(a, s));
        return 0;
    }
    // ldiv_lu!(A, B, ws) (src/lu.jl:167-176): A (destroyed) := LU, B := A^-1 B
    int lu_solve(T* A, T* Brhs, double* logdet = nullptr, T* sign = nullptr) const {
        LuArgs<T> a{};
        a.A = A; a.strideA = nn; a.lda = n;
        a.X = w.X; a.strideX = nn; a.ldx = n;
        a.rhs = Brhs; a.strideRhs = nn; a.ldrhs = n;
        a.logdet = logdet; a.sign = sign;
        a.ipiv_out = nullptr; a.info = w.info;
        a.invert = 1; a.copy_back = 0;
        a.n = n; a.batch = batch;
        CK(lu_batched<T>(a, s));
        return 0;
    }
    // dst = src where src may be one matrix shared by the batch (stride 0)
    int copy_strided(T* dst, const T* src, long long ssrc) const {
        CK(scale_batched<T>(dst, nn, n, src, ssrc, n, OP_N, nullptr, 0, SCALE_NONE, nullptr, 0, SCALE_NONE, n, batch, s));
        return 0;
    }
    int scale(T* dst, const T* src, int op, const double* rs, int rs_mode, const double* cs, int cs_mode) const {
        CK(scale_batched<T>(dst, nn, n, src, nn, n, op, rs, n, rs_mode, cs, n, cs_mode, n, batch, s));
        return 0;
    }
    int logsum(double* out, const double* d, int mode) const {
        CK(logsum_batched(out, d, n, mode, n, batch, s));
        return 0;
    }
    double* sc(int k) const { return w.sc + (size_t)k * batch; }
    T* st(int k) const { return w.st + (size_t)k * batch; }
};

// ---------------------------------------------------------------------------------------------
template <typename T> int ldr_impl(sla_handle_t h, int n, int batch, T* L, double* d, T* R, void* ws) {
    Ctx<T> c(h, n, batch, ws);
    return c.ldr(L, d, R);
}

template <typename T>
int ldr_from_impl(sla_handle_t h, int n, int batch, const T* A, long long sA, T* L, double* d, T* R, void* ws) {
    Ctx<T> c(h, n, batch, ws);
    return c.ldr(L, d, R, A, sA);
}

template <typename T> int set_identity_impl(sla_handle_t h, int n, int batch, T* L, double* d, T* R) {
    CK(set_identity_batched<T>(L, (long long)n * n, n, n, batch, h->stream));
    CK(fill_batched(d, 1.0, (long long)n * batch, h->stream));
    CK(set_identity_batched<T>(R, (long long)n * n, n, n, batch, h->stream));
    return 0;
}

// copyto!(A, F, ws): overloaded_functions.jl:34-45
template <typename T>
int ldr_to_mat_impl(sla_handle_t h, int n, int batch, T* A, const T* L, const double* d, const T* R, void* ws) {
    Ctx<T> c(h, n, batch, ws);
    CKI(c.scale(c.w.M, R, OP_N, d, SCALE_MUL, nullptr, SCALE_NONE));  // M = D R      (:39-41)
    return c.gemm(A, OP_N, L, c.nn, OP_N, c.w.M, c.nn);               // A = L M      (:42)
}

// adjoint!(At, F, ws): overloaded_functions.jl:75-86
template <typename T>
int adjoint_impl(sla_handle_t h, int n, int batch, T* At, const T* L, const double* d, const T* R, void* ws) {
    Ctx<T> c(h, n, batch, ws);
    CKI(c.scale(c.w.M, L, OP_C, d, SCALE_MUL, nullptr, SCALE_NONE));  // M = D L^H    (:80-83)
    return c.gemm(At, OP_C, R, c.nn, OP_N, c.w.M, c.nn);              // At = R^H M   (:84)
}

// lmul!(U::Matrix, V::LDR): overloaded_functions.jl:126-145
template <typename T>
int lmul_mat_ldr_impl(sla_handle_t h, int n, int batch, const T* U, long long sU, T* L, double* d, T* R, void* ws) {
    Ctx<T> c(h, n, batch, ws);
    CKI(c.gemm(c.w.M, OP_N, U, sU, OP_N, L, c.nn, nullptr, SCALE_NONE, d, SCALE_MUL));  // (U Lv) Dv   (:133-135)
    CKI(c.ldr(L, d, c.w.Mp, c.w.M, c.nn));                                               // ldr!        (:138), R0 -> M'
    CKI(c.gemm(c.w.Mpp, OP_N, c.w.Mp, c.nn, OP_N, R, c.nn));                             // R0 Rv       (:141)
    return c.copy(R, c.w.Mpp);                                                           //             (:142)
}

// rmul!(U::LDR, V::Matrix): overloaded_functions.jl:234-253
template <typename T>
int rmul_ldr_mat_impl(sla_handle_t h, int n, int batch, T* L, double* d, T* R, const T* V, long long sV, void* ws) {
    Ctx<T> c(h, n, batch, ws);
    CKI(c.gemm(c.w.M, OP_N, R, c.nn, OP_N, V, sV, d, SCALE_MUL));  // Du (Ru V)      (:241-243)
    CKI(c.ldr(c.w.M, d, R));                                        // ldr!           (:246): L0 in M, R0 -> U.R
    CKI(c.gemm(c.w.Mp, OP_N, L, c.nn, OP_N, c.w.M, c.nn));          // Lu L0          (:249)
    return c.copy(L, c.w.Mp);                                       //                (:250)
}

// lmul!(U::LDR, V::LDR): overloaded_functions.jl:166-193
template <typename T>
int lmul_ldr_ldr_impl(sla_handle_t h, int n, int batch, const T* UL, const double* Ud, const T* UR, T* VL,
                      double* Vd, T* VR, void* ws) {
    Ctx<T> c(h, n, batch, ws);
    CKI(c.gemm(c.w.M, OP_N, UR, c.nn, OP_N, VL, c.nn, Ud, SCALE_MUL, Vd, SCALE_MUL));  // Du (Ru Lv) Dv  (:173-179)
    CKI(c.ldr(c.w.M, Vd, c.w.Mp));                                                      // ldr!           (:182)
    CKI(c.gemm(VL, OP_N, UL, c.nn, OP_N, c.w.M, c.nn));                                 // L1 = Lu L0     (:185-186)
    CKI(c.gemm(c.w.Mpp, OP_N, c.w.Mp, c.nn, OP_N, VR, c.nn));                           // R1 = R0 Rv     (:189)
    return c.copy(VR, c.w.Mpp);                                                         //                (:190)
}

// rmul!(U::LDR, V::LDR): overloaded_functions.jl:274-301
template <typename T>
int rmul_ldr_ldr_impl(sla_handle_t h, int n, int batch, T* UL, double* Ud, T* UR, const T* VL, const double* Vd,
                      const T* VR, void* ws) {
    Ctx<T> c(h, n, batch, ws);
    CKI(c.gemm(c.w.M, OP_N, UR, c.nn, OP_N, VL, c.nn, Ud, SCALE_MUL, Vd, SCALE_MUL));  // Du (Ru Lv) Dv  (:281-287)
    CKI(c.ldr(c.w.M, Ud, c.w.Mp));                                                      // ldr!           (:290)
    CKI(c.gemm(c.w.Mpp, OP_N, UL, c.nn, OP_N, c.w.M, c.nn));                            // L1 = Lu L0     (:293)
    CKI(c.copy(UL, c.w.Mpp));                                                           //                (:294)
    return c.gemm(UR, OP_N, c.w.Mp, c.nn, OP_N, VR, c.nn);                              // R1 = R0 Rv     (:297-298)
}

// lmul!(U::LDR, V::Matrix): overloaded_functions.jl:97-106
template <typename T>
int lmul_ldr_mat_impl(sla_handle_t h, int n, int batch, const T* L, const double* d, const T* R, T* V, void* ws) {
    Ctx<T> c(h, n, batch, ws);
    CKI(c.gemm(c.w.M, OP_N, R, c.nn, OP_N, V, c.nn, d, SCALE_MUL));  // Du (Ru V)
    return c.gemm(V, OP_N, L, c.nn, OP_N, c.w.M, c.nn);              // Lu (...)
}

// rmul!(U::Matrix, V::LDR): overloaded_functions.jl:205-214
template <typename T>
int rmul_mat_ldr_impl(sla_handle_t h, int n, int batch, T* U, const T* L, const double* d, const T* R, void* ws) {
    Ctx<T> c(h, n, batch, ws);
    CKI(c.gemm(c.w.M, OP_N, U, c.nn, OP_N, L, c.nn, nullptr, SCALE_NONE, d, SCALE_MUL));  // (U Lv) Dv
    return c.gemm(U, OP_N, c.w.M, c.nn, OP_N, R, c.nn);                                    // (...) Rv
}

template <typename T> int det_lu_impl(sla_handle_t h, int n, int batch, T* A, double* logdet, T* sign, void* ws) {
    Ctx<T> c(h, n, batch, ws);
    return c.lu(A, nullptr, 0, 0, logdet, sign);
}

template <typename T> int inv_lu_impl(sla_handle_t h, int n, int batch, T* A, double* logdet, T* sign, void* ws) {
    Ctx<T> c(h, n, batch, ws);
    return c.lu(A, c.w.X, 1, 1, logdet, sign);
}

// logabsdet(F, ws): overloaded_functions.jl:721-733
template <typename T>
int logabsdet_impl(sla_handle_t h, int n, int batch, const T* L, const double* d, const T* R, double* logdet,
                   T* sign, void* ws) {
    Ctx<T> c(h, n, batch, ws);
    CKI(c.copy(c.w.M, L));
    CKI(c.lu(c.w.M, nullptr, 0, 0, c.sc(0), c.st(0)));
    CKI(c.logsum(c.sc(1), d, SCALE_NONE));
    CKI(c.copy(c.w.M, R));
    CKI(c.lu(c.w.M, nullptr, 0, 0, c.sc(2), c.st(2)));
    CombineArgs<T> a{};
    a.lterm[0] = c.sc(0); a.lterm[1] = c.sc(1); a.lterm[2] = c.sc(2);
    a.lcoef[0] = a.lcoef[1] = a.lcoef[2] = 1.0; a.nl = 3;
    a.sterm[0] = c.st(0); a.sterm[1] = c.st(2); a.ns = 2;
    a.logdet = logdet; a.sign = sign; a.batch = batch;
    CK(combine_batched<T>(a, c.s));
    return 0;
}

// inv_IpA!(G, A, ws): exported_functions.jl:25-71
template <typename T>
int inv_IpA_impl(sla_handle_t h, int n, int batch, T* G, const T* L, const double* d, const T* R, double* logdet,
                 T* sign, void* ws) {
    Ctx<T> c(h, n, batch, ws);
    CKI(c.copy(c.w.Mp, R));                                          // :32-33
    CKI(c.lu(c.w.Mp, c.w.X, 1, 0, nullptr, c.st(0)));                // Ra^-1 -> X, sgn          (:34)
    CK(ipa_combine_batched<T>(c.w.M, c.w.X, L, d, n, batch, c.s));   // M = La D- + Ra^-1 D+^-1  (:37-57)
    CKI(c.logsum(c.sc(1), d, SCALE_MUL_MAX1));                       // log det D+               (:49-50)
    CKI(c.lu(c.w.M, c.w.Mpp, 1, 0, c.sc(2), c.st(2)));               // M^-1 -> M''              (:60-61)
    CKI(c.gemm(G, OP_N, c.w.X, c.nn, OP_N, c.w.Mpp, c.nn));          // G = Ra^-1 D+^-1 M^-1     (:64)
    CombineArgs<T> a{};
    a.lterm[0] = c.sc(1); a.lcoef[0] = -1.0; a.lterm[1] = c.sc(2); a.lcoef[1] = 1.0; a.nl = 2;  // :68
    a.sterm[0] = c.st(0); a.sterm[1] = c.st(2); a.ns = 2;                                        // :67
    a.logdet = logdet; a.sign = sign; a.batch = batch;
    CK(combine_batched<T>(a, c.s));
    return 0;
}

// shared prologue of inv_IpUV! / inv_UpV!: sgn det Lu, Rv^-1 Dv+^-1 -> X, Du+^-1 Lu^H -> M, log dets
template <typename T>
int uv_prologue(const Ctx<T>& c, const T* UL, const double* Ud, const T* VR, const double* Vd) {
    CKI(c.copy(c.w.M, UL));                                                               // :107 / :221
    CKI(c.lu(c.w.M, nullptr, 0, 0, nullptr, c.st(3)));                                    // sgn det Lu
    CKI(c.copy(c.w.Mp, VR));                                                              // :111-112
    CKI(c.lu(c.w.Mp, c.w.X, 1, 0, nullptr, c.st(0)));                                     // Rv^-1 -> X
    CKI(c.scale(c.w.X, c.w.X, OP_N, nullptr, SCALE_NONE, Vd, SCALE_DIV_MAX1));            // Rv^-1 Dv+^-1 (:124)
    CKI(c.logsum(c.sc(0), Vd, SCALE_MUL_MAX1));                                           // log det Dv+
    CKI(c.logsum(c.sc(1), Ud, SCALE_MUL_MAX1));                                           // log det Du+
    return c.scale(c.w.M, UL, OP_C, Ud, SCALE_DIV_MAX1, nullptr, SCALE_NONE);             // Du+^-1 Lu^H  (:136-137)
}

template <typename T> int uv_epilogue(const Ctx<T>& c, double* logdet, T* sign) {
    CombineArgs<T> a{};
    a.lterm[0] = c.sc(0); a.lcoef[0] = -1.0;   // -log det Dv+
    a.lterm[1] = c.sc(2); a.lcoef[1] = 1.0;    // +log det M^-1
    a.lterm[2] = c.sc(1); a.lcoef[2] = -1.0;   // -log det Du+
    a.nl = 3;
    a.sterm[0] = c.st(0); a.sconj[0] = 0;      // sgn det Rv^-1
    a.sterm[1] = c.st(2); a.sconj[1] = 0;      // sgn det M^-1
    a.sterm[2] = c.st(3); a.sconj[2] = 1;      // conj(sgn det Lu)
    a.ns = 3;
    a.logdet = logdet; a.sign = sign; a.batch = c.batch;
    CK(combine_batched<T>(a, c.s));
    return 0;
}

// inv_IpUV!(G, U, V, ws): exported_functions.jl:97-177
template <typename T>
int inv_IpUV_impl(sla_handle_t h, int n, int batch, T* G, const T* UL, const double* Ud, const T* UR, const T* VL,
                  const double* Vd, const T* VR, double* logdet, T* sign, void* ws) {
    Ctx<T> c(h, n, batch, ws);
    CKI(uv_prologue(c, UL, Ud, VR, Vd));
    CKI(c.gemm(G, OP_N, UR, c.nn, OP_N, VL, c.nn, Ud, SCALE_MUL_MIN1, Vd, SCALE_MUL_MIN1));  // Du- (Ru Lv) Dv-  (:146-155)
    CKI(c.gemm(G, OP_N, c.w.M, c.nn, OP_N, c.w.X, c.nn, nullptr, 0, nullptr, 0, 1));          // += M M'          (:158-162)
    CKI(c.lu(G, c.w.Mpp, 1, 0, c.sc(2), c.st(2)));                                            // M^-1 -> M''      (:165-166)
    CKI(c.gemm(c.w.Mp, OP_N, c.w.Mpp, c.nn, OP_N, c.w.M, c.nn));                              // M^-1 Du+^-1 Lu^H (:169)
    CKI(c.gemm(G, OP_N, c.w.X, c.nn, OP_N, c.w.Mp, c.nn));                                    // G                (:170)
    return uv_epilogue(c, logdet, sign);                                                      //                  (:173-174)
}

// inv_UpV!(G, U, V, ws): exported_functions.jl:211-289
template <typename T>
int inv_UpV_impl(sla_handle_t h, int n, int batch, T* G, const T* UL, const double* Ud, const T* UR, const T* VL,
                 const double* Vd, const T* VR, double* logdet, T* sign, void* ws) {
    Ctx<T> c(h, n, batch, ws);
    CKI(uv_prologue(c, UL, Ud, VR, Vd));
    CKI(c.gemm(G, OP_N, UR, c.nn, OP_N, c.w.X, c.nn, Ud, SCALE_MUL_MIN1));                     // Du- (Ru Rv^-1 Dv+^-1) (:260-261)
    CKI(c.gemm(G, OP_N, c.w.M, c.nn, OP_N, VL, c.nn, nullptr, 0, Vd, SCALE_MUL_MIN1, 1));      // += (Du+^-1 Lu^H Lv) Dv- (:269-274)
    CKI(c.lu(G, c.w.Mpp, 1, 0, c.sc(2), c.st(2)));                                             // M^-1             (:277-278)
    CKI(c.gemm(c.w.Mp, OP_N, c.w.X, c.nn, OP_N, c.w.Mpp, c.nn));                               // Rv^-1 Dv+^-1 M^-1 (:281)
    CKI(c.gemm(G, OP_N, c.w.Mp, c.nn, OP_N, c.w.M, c.nn));                                     // G                (:282)
    return uv_epilogue(c, logdet, sign);                                                       //                  (:285-286)
}

// inv_invUpV!(G, U, V, ws): exported_functions.jl:324-402
template <typename T>
int inv_invUpV_impl(sla_handle_t h, int n, int batch, T* G, const T* UL, const double* Ud, const T* UR,
                    const T* VL, const double* Vd, const T* VR, double* logdet, T* sign, void* ws) {
    Ctx<T> c(h, n, batch, ws);
    CKI(c.copy(c.w.Mp, UR));                                                                   // :334
    CKI(c.lu(c.w.Mp, nullptr, 0, 0, nullptr, c.st(3)));                                        // sgn det Ru       (:335)
    CKI(c.logsum(c.sc(1), Ud, SCALE_MUL_MIN1));                                                // log det Du-      (:343)
    CKI(c.scale(c.w.Mp, UR, OP_N, Ud, SCALE_MUL_MIN1, nullptr, SCALE_NONE));                   // Du- Ru           (:346-348)
    CKI(c.copy(c.w.Mpp, VR));                                                                  // :351-352
    CKI(c.lu(c.w.Mpp, c.w.X, 1, 0, nullptr, c.st(0)));                                         // Rv^-1 -> X       (:353)
    CKI(c.scale(c.w.X, c.w.X, OP_N, nullptr, SCALE_NONE, Vd, SCALE_DIV_MAX1));                 // Rv^-1 Dv+^-1     (:365)
    CKI(c.logsum(c.sc(0), Vd, SCALE_MUL_MAX1));                                                // log det Dv+      (:361)
    CKI(c.gemm(G, OP_N, c.w.Mp, c.nn, OP_N, VL, c.nn, nullptr, 0, Vd, SCALE_MUL_MIN1));        // (Du- Ru Lv) Dv-  (:373-374)
    CKI(c.gemm(G, OP_C, UL, c.nn, OP_N, c.w.X, c.nn, Ud, SCALE_DIV_MAX1, nullptr, 0, 1));      // += Du+^-1 (Lu^H Rv^-1 Dv+^-1) (:383-387)
    CKI(c.lu(G, c.w.Mpp, 1, 0, c.sc(2), c.st(2)));                                             // M^-1             (:390-391)
    CKI(c.gemm(c.w.M, OP_N, c.w.Mpp, c.nn, OP_N, c.w.Mp, c.nn));                               // M^-1 Du- Ru      (:394)
    CKI(c.gemm(G, OP_N, c.w.X, c.nn, OP_N, c.w.M, c.nn));                                      // G                (:395)
    CombineArgs<T> a{};
    a.lterm[0] = c.sc(0); a.lcoef[0] = -1.0;
    a.lterm[1] = c.sc(2); a.lcoef[1] = 1.0;
    a.lterm[2] = c.sc(1); a.lcoef[2] = 1.0;   // + log det Du-   (:399)
    a.nl = 3;
    a.sterm[0] = c.st(0); a.sterm[1] = c.st(2); a.sterm[2] = c.st(3); a.ns = 3;   // (:398)
    a.logdet = logdet; a.sign = sign; a.batch = batch;
    CK(combine_batched<T>(a, c.s));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// ldiv! / rdiv! family (src/overloaded_functions.jl:388-709) -- SURVEY.md section 8(f)-1
// ---------------------------------------------------------------------------------------------
template <typename T> int ldiv_lu_impl(sla_handle_t h, int n, int batch, T* A, T* B, double* logdet, T* sign, void* ws) {
    Ctx<T> c(h, n, batch, ws);
    return c.lu_solve(A, B, logdet, sign);
}

// ldiv!(U::LDR, V::Matrix): V := R^-1 D^-1 L^-1 V      (:388-400)
template <typename T>
int ldiv_ldr_mat_impl(sla_handle_t h, int n, int batch, const T* L, const double* d, const T* R, T* V, void* ws) {
    Ctx<T> c(h, n, batch, ws);
    CKI(c.copy(c.w.M, L));                                                  // :391-392
    CKI(c.lu_solve(c.w.M, V));                                              // Lu^-1 V           (:393)
    CKI(c.scale(V, V, OP_N, d, SCALE_DIV, nullptr, SCALE_NONE));            // Du^-1 ...         (:395)
    CKI(c.copy(c.w.M, R));                                                  // :396-397
    return c.lu_solve(c.w.M, V);                                            // Ru^-1 ...         (:398)
}

// ldiv!(U::LDR, V::LDR): V := U^-1 V, stabilised      (:437-464)
template <typename T>
int ldiv_ldr_ldr_impl(sla_handle_t h, int n, int batch, const T* UL, const double* Ud, const T* UR, T* VL,
                      double* Vd, T* VR, void* ws) {
    Ctx<T> c(h, n, batch, ws);
    CKI(c.gemm(c.w.Mpp, OP_C, UL, c.nn, OP_N, VL, c.nn, Ud, SCALE_DIV, Vd, SCALE_MUL));   // Du^-1 (Lu^H Lv) Dv  (:439-450)
    CKI(c.copy(c.w.M, UR));                                                                // :451-452
    CKI(c.lu_solve(c.w.M, c.w.Mpp));                                                       // Ru^-1 [...]         (:453)
    CKI(c.ldr(VL, Vd, c.w.Mp, c.w.Mpp, c.nn));                                             // ldr!                (:456)
    CKI(c.gemm(c.w.M, OP_N, c.w.Mp, c.nn, OP_N, VR, c.nn));                                // R0 Rv               (:459)
    return c.copy(VR, c.w.M);                                                              //                     (:460)
}

// ldiv!(U::Matrix, V::LDR): V := U^-1 V, stabilised   (:502-521)
template <typename T>
int ldiv_mat_ldr_impl(sla_handle_t h, int n, int batch, const T* U, long long sU, T* VL, double* Vd, T* VR, void* ws) {
    Ctx<T> c(h, n, batch, ws);
    CKI(c.scale(VL, VL, OP_N, nullptr, SCALE_NONE, Vd, SCALE_MUL));         // Lv Dv             (:508-509)
    CKI(c.copy_strided(c.w.M, U, sU));                                      // :510
    CKI(c.lu_solve(c.w.M, VL));                                             // U^-1 Lv Dv        (:511)
    CKI(c.ldr(VL, Vd, c.w.Mp));                                             // ldr!              (:514)
    CKI(c.gemm(c.w.M, OP_N, c.w.Mp, c.nn, OP_N, VR, c.nn));                 // R0 Rv             (:517)
    return c.copy(VR, c.w.M);                                               //                   (:518)
}

// rdiv!(U::Matrix, V::LDR): U := U R^-1 D^-1 L^-1     (:549-565)
template <typename T>
int rdiv_mat_ldr_impl(sla_handle_t h, int n, int batch, T* U, const T* L, const double* d, const T* R, void* ws) {
    Ctx<T> c(h, n, batch, ws);
    CK(set_identity_batched<T>(c.w.Mpp, c.nn, n, n, batch, c.s));           // :551
    CKI(c.copy(c.w.M, L));                                                  // :552-553
    CKI(c.lu_solve(c.w.M, c.w.Mpp));                                        // Lv^-1             (:554)
    CKI(c.scale(c.w.Mpp, c.w.Mpp, OP_N, d, SCALE_DIV, nullptr, SCALE_NONE)); // Dv^-1 Lv^-1      (:556)
    CKI(c.copy(c.w.M, R));                                                  // :557-558
    CKI(c.lu_solve(c.w.M, c.w.Mpp));                                        // Rv^-1 Dv^-1 Lv^-1 (:559)
    CKI(c.gemm(c.w.Mp, OP_N, U, c.nn, OP_N, c.w.Mpp, c.nn));                // U [...]           (:560)
    return c.copy(U, c.w.Mp);                                               //                   (:561)
}

// rdiv!(U::LDR, V::LDR): U := U V^-1, stabilised      (:603-637)
template <typename T>
int rdiv_ldr_ldr_impl(sla_handle_t h, int n, int batch, T* UL, double* Ud, T* UR, const T* VL, const double* Vd,
                      const T* VR, void* ws) {
    Ctx<T> c(h, n, batch, ws);
    CKI(c.copy(c.w.Mp, VR));                                                               // :606-607
    CKI(c.lu(c.w.Mp, c.w.X, 1, 0, nullptr, nullptr));                                      // Rv^-1 -> X          (:608)
    CKI(c.gemm(c.w.M, OP_N, UR, c.nn, OP_N, c.w.X, c.nn, Ud, SCALE_MUL, Vd, SCALE_DIV));   // Du (Ru Rv^-1) Dv^-1 (:611-620)
    CKI(c.ldr(c.w.M, Ud, c.w.Mp));                                                         // ldr!                (:623)
    CKI(c.gemm(c.w.Mpp, OP_N, UL, c.nn, OP_N, c.w.M, c.nn));                               // Lu L0               (:626)
    CKI(c.copy(UL, c.w.Mpp));                                                              //                     (:627)
    return c.gemm(UR, OP_N, c.w.Mp, c.nn, OP_C, VL, c.nn);                                 // R0 Lv^H             (:630-632)
}

// rdiv!(U::LDR, V::Matrix): U := U V^-1, stabilised   (:673-693)
template <typename T>
int rdiv_ldr_mat_impl(sla_handle_t h, int n, int batch, T* UL, double* Ud, T* UR, const T* V, long long sV, void* ws) {
    Ctx<T> c(h, n, batch, ws);
    CKI(c.copy_strided(c.w.Mp, V, sV));                                     // :680
    CKI(c.lu(c.w.Mp, c.w.X, 1, 0, nullptr, nullptr));                       // V^-1 -> X         (:681)
    CKI(c.gemm(c.w.M, OP_N, UR, c.nn, OP_N, c.w.X, c.nn, Ud, SCALE_MUL));   // Du (Ru V^-1)      (:682-683)
    CKI(c.ldr(c.w.M, Ud, UR));                                              // ldr!              (:686)
    CKI(c.gemm(c.w.Mp, OP_N, UL, c.nn, OP_N, c.w.M, c.nn));                 // Lu L0             (:689)
    return c.copy(UL, c.w.Mp);                                              //                   (:690)
}

// ---------------------------------------------------------------------------------------------
// fused chain driver
// ---------------------------------------------------------------------------------------------
constexpr size_t BBAR_BUDGET = (size_t)6 << 30;  // bytes for the group-product ping-pong buffers

template <typename T> struct ChainWs {
    double* ev;
    T *B0, *B1, *FL, *FR, *FR2, *UL, *UR, *UR2;
    double *Fd, *Ud2;
    void* ws;
    int gc;
    static int groups_per_chunk(int n, int batch, int ng) {
        size_t per_group = 2 * (size_t)n * n * batch * sizeof(T);
        size_t g = BBAR_BUDGET / per_group;
        if (g < 1) g = 1;
        return (int)(g > (size_t)ng ? ng : g);
    }
    static size_t bytes(int n, int batch, int L, int n_stab) {
        int ng = L / n_stab;
        int gc = groups_per_chunk(n, batch, ng);
        size_t nn = align_up((size_t)n * n * batch * sizeof(T));
        size_t b = align_up((size_t)batch * L * n * sizeof(double));
        b += 2 * align_up((size_t)gc * batch * n * n * sizeof(T));
        b += 6 * nn;
        b += 2 * align_up((size_t)n * batch * sizeof(double));
        b += Ws<T>::bytes(n, batch);
        return b;
    }
    ChainWs(void* blob, int n, int batch, int L, int n_stab) {
        int ng = L / n_stab;
        gc = groups_per_chunk(n, batch, ng);
        char* p = static_cast<char*>(blob);
        size_t nn = align_up((size_t)n * n * batch * sizeof(T));
        ev = (double*)p; p += align_up((size_t)batch * L * n * sizeof(double));
        size_t bb = align_up((size_t)gc * batch * n * n * sizeof(T));
        B0 = (T*)p; p += bb;
        B1 = (T*)p; p += bb;
        FL = (T*)p; p += nn; FR = (T*)p; p += nn; FR2 = (T*)p; p += nn;
        UL = (T*)p; p += nn; UR = (T*)p; p += nn; UR2 = (T*)p; p += nn;
        Fd = (double*)p; p += align_up((size_t)n * batch * sizeof(double));
        Ud2 = (double*)p; p += align_up((size_t)n * batch * sizeof(double));
        ws = p;
    }
};

// the low-priority side stream and its events (created once per handle, outside any graph capture)
static int ensure_side_stream(sla_handle_t h) {
    if (h->side != nullptr) return 0;
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, lo));
    for (int i = 0; i < 3; ++i) CK(cudaStreamCreateWithPriority(&h->side_more[i], cudaStreamNonBlocking, lo));
    CK(cudaEventCreateWithFlags(&h->ev_start, cudaEventDisableTiming));
    for (int i = 0; i < 64; ++i) CK(cudaEventCreateWithFlags(&h->ev_chunk[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_r0, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_r1, cudaEventDisableTiming));
    h->n_events = 64;
    return 0;
}

template <typename T>
int greens_chain_impl(sla_handle_t h, int n, int batch, int L, int n_stab, const T* expK, const signed char* fields,
                      double lam, int inverse, T* G, double* logdet, T* sign, T* Lout, double* dout, T* Rout,
                      void* blob, size_t blob_bytes) {
    if (blob_bytes < ChainWs<T>::bytes(n, batch, L, n_stab)) return -17;
    ChainWs<T> cw(blob, n, batch, L, n_stab);
    Ctx<T> c(h, n, batch, cw.ws);
    const int ng = L / n_stab, gc = cw.gc;
    const long long nn = c.nn;
    cudaStream_t s = c.s;

    CK(cudaMemsetAsync(c.w.info, 0, (size_t)batch * sizeof(int), s));
    CK(hs_exp_batched(cw.ev, fields, lam, (long long)batch * L * n, s));

    // two LDR states: F (IpA), or V (first half of the groups) and U (second half) for IpUV
    T *FL = cw.FL, *FR = cw.FR, *FR2 = cw.FR2;
    double* Fd = cw.Fd;
    CKI(set_identity_impl<T>(h, n, batch, FL, Fd, FR));
    T *UL = cw.UL, *UR = cw.UR, *UR2 = cw.UR2;
    double* Ud = cw.Ud2;
    const int half = ng / 2;
    if (inverse == SLA_INV_IPUV) CKI(set_identity_impl<T>(h, n, batch, UL, Ud, UR));

    // Group products are independent of the chain state: with SLA_OVERLAP (default on) they are built in
    // sub-chunks on a side stream and the chain waits per sub-chunk, so that GEMM CTAs fill the SMs the
    // cluster LDR kernels cannot use.
    static const bool overlap = [] { const char* e = getenv("SLA_OVERLAP"); return !(e && e[0] == '0'); }();
    static const int SUB0 = [] { const char* e = getenv("SLA_OVERLAP_SUB"); int v = e ? atoi(e) : 1; return v > 0 ? v : 1; }();   // groups per sub-chunk
    // measured on the C2 step (gpurun_out/r2_sides.log): 1 stream x 2 groups 31.69 ms, 2 streams x 1 group 31.43, 2 x 2 32.27,
    // 3 x 1 32.11, 4 x 1 32.62 -- small low-priority launches leave the SMs to the chain's cluster kernels sooner
    if (overlap) CKI(ensure_side_stream(h));
    // Small problems (C1: one 64 x 64 chain) are bound by the length of the launch chain, not by SM-time: the product R0 R
    // of a stabilisation is needed only by the NEXT R product and by the final inverse, so it leaves the chain for a stream
    // of its own (joined before the next ldr! overwrites R0, and at the end).  Measured without effect on the SM-time-bound
    // C2 step (profiles/launches_r2g_C2_step_summary.txt), hence the size gate; SLA_RFORK=0/1 overrides it.
    static const int rfork_env = getenv("SLA_RFORK") ? atoi(getenv("SLA_RFORK")) : -1;
    const bool rfork = overlap && (rfork_env >= 0 ? rfork_env != 0 : (long long)n * n * batch <= 64LL * 64 * 8);
    cudaStream_t rstream = rfork ? h->side_more[2] : s;
    bool r_pending = false;
    for (int g0 = 0; g0 < ng; g0 += gc) {
        const int gcur = (ng - g0 < gc) ? (ng - g0) : gc;
        const int SUB = SUB0 > (gcur + 63) / 64 ? SUB0 : (gcur + 63) / 64;   // at most 64 sub-chunks (events) per pass
        const int nsub = (gcur + SUB - 1) / SUB;
        const bool ov = overlap && nsub <= 64;
        // Consecutive sub-chunks are independent of each other: they go round-robin over SLA_SIDES (default 2) side streams so that the
        // partial last wave of one sub-chunk's GEMM launches is filled by the other's (SLA_SIDES=1: one side stream).
        // The second stream starts after sub-chunk 0, which the chain is waiting for, has the GPU to itself.
        static const int n_sides = [] { const char* e = getenv("SLA_SIDES"); int v = e ? atoi(e) : 2; return v < 1 ? 1 : (v > 4 ? 4 : v); }();
        cudaStream_t bs = ov ? h->side : s;
        if (ov) {
            CK(cudaEventRecord(h->ev_start, s));          // the buffers' previous consumers / exp(lam s) are done
            CK(cudaStreamWaitEvent(h->side, h->ev_start, 0));
        }
        // X_1 = diag(e_1) expK ; X_i = diag(e_i) (expK X_{i-1})      (test/runtests.jl:93-100)
        // group-major layout: product of (group gg, chain c) lives at index gg*batch + c
        T* fin = ((n_stab - 1) & 1) ? cw.B1 : cw.B0;
        for (int sc = 0; sc < nsub; ++sc) {
            const int ga = sc * SUB, gb = (ga + SUB < gcur) ? ga + SUB : gcur;
            T *cur = cw.B0, *nxt = cw.B1;
            if (ov && n_sides > 1 && nsub > 2) {
                const int si = sc % n_sides;
                bs = si == 0 ? h->side : h->side_more[si - 1];
                if (sc >= 1 && sc < n_sides) CK(cudaStreamWaitEvent(bs, h->ev_chunk[0], 0));   // (ev_chunk[0] follows ev_start)
            }
            for (int i = 0; i < n_stab; ++i) {
                const double* rs0 = cw.ev + ((long long)(g0 + ga) * n_stab + i) * n;
                if (i == 0) {
                    for (int gg = ga; gg < gb; ++gg)
                        CK(scale_batched<T>(cur + (long long)gg * batch * nn, nn, n, expK, 0, n, OP_N,
                                            cw.ev + ((long long)(g0 + gg) * n_stab) * n, (long long)L * n, SCALE_MUL,
                                            nullptr, 0, SCALE_NONE, n, batch, bs));
                    continue;
                }
                GemmArgs<T> g{};
                g.A = expK; g.strideA = 0; g.lda = n;
                g.B = cur + (long long)ga * batch * nn; g.strideB = nn; g.ldb = n;
                g.C = nxt + (long long)ga * batch * nn; g.strideC = nn; g.ldc = n;
                g.M = g.N = g.K = n;
                // row scaling exp(lam s) of (group gg = b / batch, chain c = b % batch), slice (g0+gg)*n_stab + i
                g.rs = rs0; g.rs_mode = SCALE_MUL; g.rs_div = batch; g.stride_rs_hi = (long long)n_stab * n;
                g.stride_rs = (long long)L * n;
                g.cs = nullptr; g.cs_mode = SCALE_NONE; g.stride_cs = 0;
                g.accumulate = 0; g.batch = batch * (gb - ga);
                CK(gemm_batched<T>(OP_N, OP_N, g, bs));
                T* t = cur; cur = nxt; nxt = t;
            }
            if (ov) CK(cudaEventRecord(h->ev_chunk[sc], bs));
        }
        // stabilised updates  F <- B_bar_g F   (lmul!, overloaded_functions.jl:126-145) with ping-pong R
        for (int gg = 0; gg < gcur; ++gg) {
            const int g = g0 + gg;
            if (ov && (gg % SUB) == 0) CK(cudaStreamWaitEvent(s, h->ev_chunk[gg / SUB], 0));
            const bool intoU = (inverse == SLA_INV_IPUV) && (g >= half);
            T*& XL = intoU ? UL : FL;
            T*& XR = intoU ? UR : FR;
            T*& XR2 = intoU ? UR2 : FR2;
            double* Xd = intoU ? Ud : Fd;
            const T* Bg = fin + (long long)gg * batch * nn;
            if (g == 0 || (intoU && g == half)) {
                // the state is still the identity: (B I) 1 = B and R0 I = R0 bit for bit, so the two GEMMs of lmul! are
                // skipped -- ldr! reads B_bar directly and writes R0 where R lives
                CKI(c.ldr(XL, Xd, XR, Bg, nn));
                continue;
            }
            CKI(c.gemm(c.w.M, OP_N, Bg, nn, OP_N, XL, nn, nullptr, SCALE_NONE, Xd, SCALE_MUL));
            if (r_pending) CK(cudaStreamWaitEvent(s, h->ev_r1, 0));   // the previous R product has consumed R0
            CKI(c.ldr(XL, Xd, c.w.Mp, c.w.M, nn));
            if (rfork) {
                CK(cudaEventRecord(h->ev_r0, s));
                CK(cudaStreamWaitEvent(rstream, h->ev_r0, 0));
                Ctx<T> cr = c;
                cr.s = rstream;
                CKI(cr.gemm(XR2, OP_N, c.w.Mp, nn, OP_N, XR, nn));
                CK(cudaEventRecord(h->ev_r1, rstream));
                r_pending = true;
            } else {
                CKI(c.gemm(XR2, OP_N, c.w.Mp, nn, OP_N, XR, nn));
            }
            T* t = XR; XR = XR2; XR2 = t;
        }
    }
    if (r_pending) CK(cudaStreamWaitEvent(s, h->ev_r1, 0));
    if (inverse == SLA_INV_IPA) {
        CKI(inv_IpA_impl<T>(h, n, batch, G, FL, Fd, FR, logdet, sign, cw.ws));
        if (Lout) CKI(c.copy(Lout, FL));
        if (dout) CKI(c.copy_d(dout, Fd));
        if (Rout) CKI(c.copy(Rout, FR));
        return 0;
    }
    return inv_IpUV_impl<T>(h, n, batch, G, UL, Ud, UR, FL, Fd, FR, logdet, sign, cw.ws);
}

// ---------------------------------------------------------------------------------------------
// fused sweep stepper: the caller pattern of inv_IpUV! over stacks of LDR factorisations
// ---------------------------------------------------------------------------------------------
template <typename T> struct SweepWs {
    double *ev, *evi;          // exp(+lam s), exp(-lam s): [batch][L][n]
    T *B0, *B1;                // group products, ping-pong: [ng][batch][n][n]
    T *SL, *SR; double* Sd;    // the stack of left factorisations F_left[0..ng]: [ng+1][batch] x (L, d, R)
    T *RL, *RR, *RR2; double* Rd;   // the right factorisation B(tau_g, 0) (ping-pong R)
    T *Ga, *Gb;                // wrapped / re-stabilised Green's function
    void* ws;
    static size_t bytes(int n, int batch, int L, int n_stab) {
        const int ng = L / n_stab;
        const size_t nn = align_up((size_t)n * n * batch * sizeof(T));
        const size_t nd = align_up((size_t)n * batch * sizeof(double));
        size_t b = 2 * align_up((size_t)batch * L * n * sizeof(double));
        b += 2 * align_up((size_t)ng * batch * n * n * sizeof(T));
        b += 2 * (size_t)(ng + 1) * nn + (size_t)(ng + 1) * nd;
        b += 3 * nn + nd;
        b += 2 * nn;
        b += Ws<T>::bytes(n, batch);
        return b;
    }
    SweepWs(void* blob, int n, int batch, int L, int n_stab) {
        const int ng = L / n_stab;
        char* p = static_cast<char*>(blob);
        const size_t nn = align_up((size_t)n * n * batch * sizeof(T));
        const size_t nd = align_up((size_t)n * batch * sizeof(double));
        const size_t ne = align_up((size_t)batch * L * n * sizeof(double));
        const size_t nb = align_up((size_t)ng * batch * n * n * sizeof(T));
        ev = (double*)p; p += ne; evi = (double*)p; p += ne;
        B0 = (T*)p; p += nb; B1 = (T*)p; p += nb;
        SL = (T*)p; p += (size_t)(ng + 1) * nn; SR = (T*)p; p += (size_t)(ng + 1) * nn;
        Sd = (double*)p; p += (size_t)(ng + 1) * nd;
        RL = (T*)p; p += nn; RR = (T*)p; p += nn; RR2 = (T*)p; p += nn;
        Rd = (double*)p; p += nd;
        Ga = (T*)p; p += nn; Gb = (T*)p; p += nn;
        ws = p;
    }
    T* sl(int g, int n, int batch) const { return (T*)((char*)SL + (size_t)g * align_up((size_t)n * n * batch * sizeof(T))); }
    T* sr(int g, int n, int batch) const { return (T*)((char*)SR + (size_t)g * align_up((size_t)n * n * batch * sizeof(T))); }
    double* sd(int g, int n, int batch) const { return (double*)((char*)Sd + (size_t)g * align_up((size_t)n * batch * sizeof(double))); }
};

// One upward sweep per chain, fields fixed (restated on the CPU in oracle/sweep.py):
//   setup  F_left[ng] = I;  F_left[g] = F_left[g+1] * B_bar_{g+1}   rmul!(LDR, Matrix)  overloaded_functions.jl:234-253
//          G(0) = inv_IpA!(F_left[0])                                                   exported_functions.jl:25-71
//   sweep  G <- B_l G B_l^-1 for every slice (two GEMMs with the diagonal factors fused into their epilogues);
//          every n_stab slices  F_right <- lmul!(B_bar_g, F_right)                      overloaded_functions.jl:126-145
//                               G' = inv_IpUV!(F_right, F_left[g])                      exported_functions.jl:97-177
//                               dG[g] = max|G' - G|,  G <- G'
template <typename T>
int sweep_impl(sla_handle_t h, int n, int batch, int L, int n_stab, const T* expK, const T* expKinv, const signed char* fields,
               double lam, T* G0, T* G, double* logdet, T* sign, double* dG, void* blob, size_t blob_bytes) {
    if (blob_bytes < SweepWs<T>::bytes(n, batch, L, n_stab)) return -17;
    SweepWs<T> sw(blob, n, batch, L, n_stab);
    Ctx<T> c(h, n, batch, sw.ws);
    const int ng = L / n_stab;
    const long long nn = c.nn;
    cudaStream_t s = c.s;
    CK(cudaMemsetAsync(c.w.info, 0, (size_t)batch * sizeof(int), s));
    CK(hs_exp_batched(sw.ev, fields, lam, (long long)batch * L * n, s));
    CK(hs_exp_batched(sw.evi, fields, -lam, (long long)batch * L * n, s));

    // ---- group products of every group at once (group-major: product of (group g, chain b) at index g*batch + b)
    T* fin = ((n_stab - 1) & 1) ? sw.B1 : sw.B0;
    {
        T *cur = sw.B0, *nxt = sw.B1;
        for (int i = 0; i < n_stab; ++i) {
            if (i == 0) {
                for (int g = 0; g < ng; ++g)
                    CK(scale_batched<T>(cur + (long long)g * batch * nn, nn, n, expK, 0, n, OP_N,
                                        sw.ev + ((long long)g * n_stab) * n, (long long)L * n, SCALE_MUL, nullptr, 0, SCALE_NONE,
                                        n, batch, s));
                continue;
            }
            GemmArgs<T> g{};
            g.A = expK; g.strideA = 0; g.lda = n;
            g.B = cur; g.strideB = nn; g.ldb = n;
            g.C = nxt; g.strideC = nn; g.ldc = n;
            g.M = g.N = g.K = n;
            g.rs = sw.ev + (long long)i * n; g.rs_mode = SCALE_MUL; g.rs_div = batch; g.stride_rs_hi = (long long)n_stab * n;
            g.stride_rs = (long long)L * n;
            g.batch = batch * ng;
            CK(gemm_batched<T>(OP_N, OP_N, g, s));
            T* t = cur; cur = nxt; nxt = t;
        }
    }
    // ---- setup: the stack of left factorisations, from the top of the chain downwards (no copies: slot g+1 -> slot g)
    CKI(set_identity_impl<T>(h, n, batch, sw.sl(ng, n, batch), sw.sd(ng, n, batch), sw.sr(ng, n, batch)));
    for (int g = ng - 1; g >= 0; --g) {
        const T* Bg = fin + (long long)g * batch * nn;
        CKI(c.gemm(c.w.M, OP_N, sw.sr(g + 1, n, batch), nn, OP_N, Bg, nn, sw.sd(g + 1, n, batch), SCALE_MUL));   // Du (Ru V)   (:241-243)
        CKI(c.ldr(c.w.M, sw.sd(g, n, batch), sw.sr(g, n, batch)));                                                // ldr!        (:246)
        CKI(c.gemm(sw.sl(g, n, batch), OP_N, sw.sl(g + 1, n, batch), nn, OP_N, c.w.M, nn));                        // Lu L0       (:249)
    }
    T *Gw = sw.Ga, *Gs = sw.Gb;
    CKI(inv_IpA_impl<T>(h, n, batch, Gw, sw.sl(0, n, batch), sw.sd(0, n, batch), sw.sr(0, n, batch), logdet, sign, sw.ws));
    if (G0) CKI(c.copy(G0, Gw));
    // ---- sweep
    T *RL = sw.RL, *RR = sw.RR, *RR2 = sw.RR2;
    CKI(set_identity_impl<T>(h, n, batch, RL, sw.Rd, RR));
    for (int g = 0; g < ng; ++g) {
        for (int i = 0; i < n_stab; ++i) {
            const long long l = (long long)g * n_stab + i;
            GemmArgs<T> a{};                 // T1 = diag(e_l) (expK G)
            a.A = expK; a.strideA = 0; a.lda = n;
            a.B = Gw; a.strideB = nn; a.ldb = n;
            a.C = c.w.M; a.strideC = nn; a.ldc = n;
            a.M = a.N = a.K = n;
            a.rs = sw.ev + l * n; a.stride_rs = (long long)L * n; a.rs_mode = SCALE_MUL;
            a.batch = batch;
            CK(gemm_batched<T>(OP_N, OP_N, a, s));
            GemmArgs<T> b2{};                // G = (T1 expK^-1) diag(1 / e_l)
            b2.A = c.w.M; b2.strideA = nn; b2.lda = n;
            b2.B = expKinv; b2.strideB = 0; b2.ldb = n;
            b2.C = Gw; b2.strideC = nn; b2.ldc = n;
            b2.M = b2.N = b2.K = n;
            b2.cs = sw.evi + l * n; b2.stride_cs = (long long)L * n; b2.cs_mode = SCALE_MUL;
            b2.batch = batch;
            CK(gemm_batched<T>(OP_N, OP_N, b2, s));
        }
        const T* Bg = fin + (long long)g * batch * nn;
        CKI(c.gemm(c.w.M, OP_N, Bg, nn, OP_N, RL, nn, nullptr, SCALE_NONE, sw.Rd, SCALE_MUL));   // (U Lv) Dv   (:133-135)
        CKI(c.ldr(RL, sw.Rd, c.w.Mp, c.w.M, nn));                                                // ldr!        (:138)
        CKI(c.gemm(RR2, OP_N, c.w.Mp, nn, OP_N, RR, nn));                                        // R0 Rv       (:141)
        { T* t = RR; RR = RR2; RR2 = t; }
        CKI(inv_IpUV_impl<T>(h, n, batch, Gs, RL, sw.Rd, RR, sw.sl(g + 1, n, batch), sw.sd(g + 1, n, batch),
                             sw.sr(g + 1, n, batch), logdet + (size_t)(g + 1) * batch, sign + (size_t)(g + 1) * batch, sw.ws));
        if (dG) CK(maxabsdiff_batched<T>(dG + (size_t)g * batch, Gs, Gw, nn, nn, batch, s));
        { T* t = Gw; Gw = Gs; Gs = t; }
    }
    return c.copy(G, Gw);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// extern "C"
// ---------------------------------------------------------------------------------------------
#define CHECK_COMMON(h, dtype, n, batch)                 \
    do {                                                 \
        if (!(h)) return -1;                             \
        if (!device_is_current(h)) return SLA_EDEVICE;   \
        if ((dtype) != SLA_F64 && (dtype) != SLA_C64) return -2; \
        if ((n) <= 0) return -3;                         \
        if (!n_supported(dtype, n)) return SLA_ENOTIMPL;  \
        if ((batch) < 0) return -4;                      \
        if ((batch) == 0) return 0;                      \
    } while (0)

// largest n: the QR / LU panels of one matrix must fit the shared memory of one CTA (sla_max_n)
static inline bool n_supported(int dtype, int n) {
    return dtype == SLA_F64 ? (ldr_twork_elems<double>(n) > 0 && lu_supported<double>(n))
                            : (ldr_twork_elems<cplx>(n) > 0 && lu_supported<cplx>(n));
}
// the handle's device must be the calling thread's current device (buffers, streams and kernel attributes are per device)
static inline bool device_is_current(sla_handle_t h) {
    int dev = -1;
    return cudaGetDevice(&dev) == cudaSuccess && dev == h->device;
}
// LAPACK-style info codes accumulate over the kernels of ONE entry point: cleared here, on the stream, and only
// ever raised from 0 by the kernels (the first failure is the one reported, as chklapackerror would throw it)
static inline int clear_info(sla_handle_t h, int dtype, int n, int batch, void* ws) {
    int* info = dtype == SLA_F64 ? Ws<double>(ws, n, batch).info : Ws<cplx>(ws, n, batch).info;
    cudaError_t e = cudaMemsetAsync(info, 0, (size_t)batch * sizeof(int), h->stream);
    return e == cudaSuccess ? 0 : 1000 + (int)e;
}
#define CLEAR_INFO(h, dtype, n, batch, ws) do { int r__ = clear_info(h, dtype, n, batch, ws); if (r__) return r__; } while (0)

#define DISPATCH(dtype, fn, ...) \
    ((dtype) == SLA_F64 ? fn<double>(__VA_ARGS__) : fn<cplx>(__VA_ARGS__))

// SLA_F32 / SLA_C32 are storage types served by promote -> FP64 entry point -> demote (sla_capi_f32.cu)
static inline bool is_f32(int dtype) { return dtype == SLA_F32 || dtype == SLA_C32; }
cudaStream_t sla_handle_stream(sla_handle_t h) { return h->stream; }

extern "C" {

int sla_version(void) { return 101; }

unsigned long long sla_launch_count(void) { return g_launch_count.load(); }

void sla_debug_set_ldr_prof(void* dev_buffer_128_int64) { g_ldr_prof = static_cast<long long*>(dev_buffer_128_int64); }

int sla_create(sla_handle_t* handle, void* stream) {
    if (!handle) return -1;
    int dev = 0;
    CK(cudaGetDevice(&dev));
    sla_handle_s* h = new (std::nothrow) sla_handle_s;
    if (!h) return -1;
    h->stream = static_cast<cudaStream_t>(stream);
    h->device = dev;
    *handle = h;
    return 0;
}

int sla_destroy(sla_handle_t handle) {
    if (handle) {
        for (auto& e : handle->graphs)
            if (e.exec) cudaGraphExecDestroy(e.exec);
        if (handle->side) cudaStreamDestroy(handle->side);
        for (int i = 0; i < 3; ++i) if (handle->side_more[i]) cudaStreamDestroy(handle->side_more[i]);
        if (handle->ev_start) cudaEventDestroy(handle->ev_start);
        for (int i = 0; i < handle->n_events; ++i) cudaEventDestroy(handle->ev_chunk[i]);
        if (handle->ev_r0) cudaEventDestroy(handle->ev_r0);
        if (handle->ev_r1) cudaEventDestroy(handle->ev_r1);
    }
    delete handle;
    return 0;
}

int sla_last_ldr_algo(sla_handle_t handle) { return handle ? handle->last_ldr_algo : -1; }

unsigned long long sla_ldr_algo_launches(sla_handle_t handle, int algo) {
    return (handle && algo >= 0 && algo < SLA_N_LDR_ALGOS) ? handle->ldr_algo_launches[algo] : 0ull;
}

int sla_set_stream(sla_handle_t handle, void* stream) {
    if (!handle) return -1;
    handle->stream = static_cast<cudaStream_t>(stream);
    return 0;
}

int sla_max_n(int dtype) {
    if (is_f32(dtype)) dtype -= 2;   // storage types compute in FP64
    if (dtype != SLA_F64 && dtype != SLA_C64) return 0;
    int lo = 1, hi = 1 << 14;   // n_supported is monotone in n
    while (lo < hi) { int mid = (lo + hi + 1) / 2; if (n_supported(dtype, mid)) lo = mid; else hi = mid - 1; }
    return lo;
}

size_t sla_workspace_bytes(int dtype, int n, int batch) {
    if (is_f32(dtype)) return f32_workspace_bytes(dtype, n, batch);
    if (n <= 0 || batch <= 0 || (dtype != SLA_F64 && dtype != SLA_C64) || !n_supported(dtype, n)) return 0;
    return dtype == SLA_F64 ? Ws<double>::bytes(n, batch) : Ws<cplx>::bytes(n, batch);
}

int* sla_workspace_info(void* ws, int dtype, int n, int batch) {
    if (is_f32(dtype)) dtype -= 2;   // the FP64 workspace is the head of the blob
    if (!ws || n <= 0 || batch <= 0) return nullptr;
    return dtype == SLA_F64 ? Ws<double>(ws, n, batch).info : Ws<cplx>(ws, n, batch).info;
}

#define TP(x) static_cast<T*>(x)
#define CTP(x) static_cast<const T*>(x)

int sla_ldr(sla_handle_t h, int dtype, int n, int batch, void* L, double* d, void* R, void* ws) {
    if (is_f32(dtype)) return f32_ldr(h, dtype, n, batch, L, d, R, ws);
    CHECK_COMMON(h, dtype, n, batch);
    if (!L) return -5; if (!d) return -6; if (!R) return -7; if (!ws) return -8; CLEAR_INFO(h, dtype, n, batch, ws);
    if (dtype == SLA_F64) return ldr_impl<double>(h, n, batch, (double*)L, d, (double*)R, ws);
    return ldr_impl<cplx>(h, n, batch, (cplx*)L, d, (cplx*)R, ws);
}

int sla_ldr_from(sla_handle_t h, int dtype, int n, int batch, const void* A, long long strideA, void* L, double* d,
                 void* R, void* ws) {
    if (is_f32(dtype)) return f32_ldr_from(h, dtype, n, batch, A, strideA, L, d, R, ws);
    CHECK_COMMON(h, dtype, n, batch);
    if (!A) return -5; if (!L) return -7; if (!d) return -8; if (!R) return -9; if (!ws) return -10; CLEAR_INFO(h, dtype, n, batch, ws);
    if (dtype == SLA_F64) return ldr_from_impl<double>(h, n, batch, (const double*)A, strideA, (double*)L, d, (double*)R, ws);
    return ldr_from_impl<cplx>(h, n, batch, (const cplx*)A, strideA, (cplx*)L, d, (cplx*)R, ws);
}

int sla_set_identity(sla_handle_t h, int dtype, int n, int batch, void* L, double* d, void* R) {
    if (is_f32(dtype)) return f32_set_identity(h, dtype, n, batch, L, d, R);
    CHECK_COMMON(h, dtype, n, batch);
    if (!L) return -5; if (!d) return -6; if (!R) return -7;
    if (dtype == SLA_F64) return set_identity_impl<double>(h, n, batch, (double*)L, d, (double*)R);
    return set_identity_impl<cplx>(h, n, batch, (cplx*)L, d, (cplx*)R);
}

int sla_copy(sla_handle_t h, int dtype, int n, int batch, void* Lo, double* dout, void* Ro, const void* Li,
             const double* din, const void* Ri) {
    if (is_f32(dtype)) return f32_copy(h, dtype, n, batch, Lo, dout, Ro, Li, din, Ri);
    CHECK_COMMON(h, dtype, n, batch);
    if (!Lo) return -5; if (!dout) return -6; if (!Ro) return -7; if (!Li) return -8; if (!din) return -9; if (!Ri) return -10;
    size_t es = dtype == SLA_F64 ? sizeof(double) : sizeof(cplx);
    size_t mb = (size_t)n * n * batch * es;
    CK(cudaMemcpyAsync(Lo, Li, mb, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(dout, din, (size_t)n * batch * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(Ro, Ri, mb, cudaMemcpyDeviceToDevice, h->stream));
    return 0;
}

int sla_ldr_to_mat(sla_handle_t h, int dtype, int n, int batch, void* A, const void* L, const double* d,
                   const void* R, void* ws) {
    if (is_f32(dtype)) return f32_ldr_to_mat(h, dtype, n, batch, A, L, d, R, ws, 0);
    CHECK_COMMON(h, dtype, n, batch);
    if (!A) return -5; if (!L) return -6; if (!d) return -7; if (!R) return -8; if (!ws) return -9; CLEAR_INFO(h, dtype, n, batch, ws);
    if (dtype == SLA_F64) return ldr_to_mat_impl<double>(h, n, batch, (double*)A, (const double*)L, d, (const double*)R, ws);
    return ldr_to_mat_impl<cplx>(h, n, batch, (cplx*)A, (const cplx*)L, d, (const cplx*)R, ws);
}

int sla_adjoint(sla_handle_t h, int dtype, int n, int batch, void* At, const void* L, const double* d, const void* R,
                void* ws) {
    if (is_f32(dtype)) return f32_ldr_to_mat(h, dtype, n, batch, At, L, d, R, ws, 1);
    CHECK_COMMON(h, dtype, n, batch);
    if (!At) return -5; if (!L) return -6; if (!d) return -7; if (!R) return -8; if (!ws) return -9; CLEAR_INFO(h, dtype, n, batch, ws);
    if (dtype == SLA_F64) return adjoint_impl<double>(h, n, batch, (double*)At, (const double*)L, d, (const double*)R, ws);
    return adjoint_impl<cplx>(h, n, batch, (cplx*)At, (const cplx*)L, d, (const cplx*)R, ws);
}

int sla_lmul_mat_ldr(sla_handle_t h, int dtype, int n, int batch, const void* U, long long strideU, void* L,
                     double* d, void* R, void* ws) {
    if (is_f32(dtype)) return f32_mat_ldr(h, dtype, n, batch, U, strideU, L, d, R, ws, 0);
    CHECK_COMMON(h, dtype, n, batch);
    if (!U) return -5; if (!L) return -7; if (!d) return -8; if (!R) return -9; if (!ws) return -10; CLEAR_INFO(h, dtype, n, batch, ws);
    if (dtype == SLA_F64) return lmul_mat_ldr_impl<double>(h, n, batch, (const double*)U, strideU, (double*)L, d, (double*)R, ws);
    return lmul_mat_ldr_impl<cplx>(h, n, batch, (const cplx*)U, strideU, (cplx*)L, d, (cplx*)R, ws);
}

int sla_rmul_ldr_mat(sla_handle_t h, int dtype, int n, int batch, void* L, double* d, void* R, const void* V,
                     long long strideV, void* ws) {
    if (is_f32(dtype)) return f32_ldr_mat(h, dtype, n, batch, L, d, R, V, strideV, ws, 0);
    CHECK_COMMON(h, dtype, n, batch);
    if (!L) return -5; if (!d) return -6; if (!R) return -7; if (!V) return -8; if (!ws) return -10; CLEAR_INFO(h, dtype, n, batch, ws);
    if (dtype == SLA_F64) return rmul_ldr_mat_impl<double>(h, n, batch, (double*)L, d, (double*)R, (const double*)V, strideV, ws);
    return rmul_ldr_mat_impl<cplx>(h, n, batch, (cplx*)L, d, (cplx*)R, (const cplx*)V, strideV, ws);
}

#define UV_NULLCHECK()                                                                           \
    if (!UL) return -5; if (!Ud) return -6; if (!UR) return -7; if (!VL) return -8; if (!Vd) return -9; \
    if (!VR) return -10;

int sla_lmul_ldr_ldr(sla_handle_t h, int dtype, int n, int batch, const void* UL, const double* Ud, const void* UR,
                     void* VL, double* Vd, void* VR, void* ws) {
    if (is_f32(dtype)) return f32_ldr_ldr(h, dtype, n, batch, (void*)UL, (double*)Ud, (void*)UR, VL, Vd, VR, ws, 0);
    CHECK_COMMON(h, dtype, n, batch);
    UV_NULLCHECK();
    if (!ws) return -11; CLEAR_INFO(h, dtype, n, batch, ws);
    if (UL == VL || UR == VR || Ud == Vd) return -8;  // U and V must be distinct (overloaded_functions.jl:179,185)
    if (dtype == SLA_F64)
        return lmul_ldr_ldr_impl<double>(h, n, batch, (const double*)UL, Ud, (const double*)UR, (double*)VL, Vd, (double*)VR, ws);
    return lmul_ldr_ldr_impl<cplx>(h, n, batch, (const cplx*)UL, Ud, (const cplx*)UR, (cplx*)VL, Vd, (cplx*)VR, ws);
}

int sla_rmul_ldr_ldr(sla_handle_t h, int dtype, int n, int batch, void* UL, double* Ud, void* UR, const void* VL,
                     const double* Vd, const void* VR, void* ws) {
    if (is_f32(dtype)) return f32_ldr_ldr(h, dtype, n, batch, UL, Ud, UR, (void*)VL, (double*)Vd, (void*)VR, ws, 2);
    CHECK_COMMON(h, dtype, n, batch);
    UV_NULLCHECK();
    if (!ws) return -11; CLEAR_INFO(h, dtype, n, batch, ws);
    if (UL == VL || UR == VR || Ud == Vd) return -8;
    if (dtype == SLA_F64)
        return rmul_ldr_ldr_impl<double>(h, n, batch, (double*)UL, Ud, (double*)UR, (const double*)VL, Vd, (const double*)VR, ws);
    return rmul_ldr_ldr_impl<cplx>(h, n, batch, (cplx*)UL, Ud, (cplx*)UR, (const cplx*)VL, Vd, (const cplx*)VR, ws);
}

int sla_lmul_ldr_mat(sla_handle_t h, int dtype, int n, int batch, const void* L, const double* d, const void* R,
                     void* V, void* ws) {
    if (is_f32(dtype)) return f32_dense(h, dtype, n, batch, V, L, d, R, ws, 0);
    CHECK_COMMON(h, dtype, n, batch);
    if (!L) return -5; if (!d) return -6; if (!R) return -7; if (!V) return -8; if (!ws) return -9; CLEAR_INFO(h, dtype, n, batch, ws);
    if (dtype == SLA_F64) return lmul_ldr_mat_impl<double>(h, n, batch, (const double*)L, d, (const double*)R, (double*)V, ws);
    return lmul_ldr_mat_impl<cplx>(h, n, batch, (const cplx*)L, d, (const cplx*)R, (cplx*)V, ws);
}

int sla_rmul_mat_ldr(sla_handle_t h, int dtype, int n, int batch, void* U, const void* L, const double* d,
                     const void* R, void* ws) {
    if (is_f32(dtype)) return f32_dense(h, dtype, n, batch, U, L, d, R, ws, 2);
    CHECK_COMMON(h, dtype, n, batch);
    if (!U) return -5; if (!L) return -6; if (!d) return -7; if (!R) return -8; if (!ws) return -9; CLEAR_INFO(h, dtype, n, batch, ws);
    if (dtype == SLA_F64) return rmul_mat_ldr_impl<double>(h, n, batch, (double*)U, (const double*)L, d, (const double*)R, ws);
    return rmul_mat_ldr_impl<cplx>(h, n, batch, (cplx*)U, (const cplx*)L, d, (const cplx*)R, ws);
}

int sla_det_lu(sla_handle_t h, int dtype, int n, int batch, void* A, double* logdet, void* sign, void* ws) {
    if (is_f32(dtype)) return f32_lu(h, dtype, n, batch, A, nullptr, logdet, sign, ws, 0);
    CHECK_COMMON(h, dtype, n, batch);
    if (!A) return -5; if (!ws) return -8; CLEAR_INFO(h, dtype, n, batch, ws);
    if (dtype == SLA_F64) return det_lu_impl<double>(h, n, batch, (double*)A, logdet, (double*)sign, ws);
    return det_lu_impl<cplx>(h, n, batch, (cplx*)A, logdet, (cplx*)sign, ws);
}

int sla_inv_lu(sla_handle_t h, int dtype, int n, int batch, void* A, double* logdet, void* sign, void* ws) {
    if (is_f32(dtype)) return f32_lu(h, dtype, n, batch, A, nullptr, logdet, sign, ws, 1);
    CHECK_COMMON(h, dtype, n, batch);
    if (!A) return -5; if (!ws) return -8; CLEAR_INFO(h, dtype, n, batch, ws);
    if (dtype == SLA_F64) return inv_lu_impl<double>(h, n, batch, (double*)A, logdet, (double*)sign, ws);
    return inv_lu_impl<cplx>(h, n, batch, (cplx*)A, logdet, (cplx*)sign, ws);
}

int sla_logabsdet(sla_handle_t h, int dtype, int n, int batch, const void* L, const double* d, const void* R,
                  double* logdet, void* sign, void* ws) {
    if (is_f32(dtype)) return f32_logabsdet(h, dtype, n, batch, L, d, R, logdet, sign, ws);
    CHECK_COMMON(h, dtype, n, batch);
    if (!L) return -5; if (!d) return -6; if (!R) return -7; if (!ws) return -10; CLEAR_INFO(h, dtype, n, batch, ws);
    if (dtype == SLA_F64) return logabsdet_impl<double>(h, n, batch, (const double*)L, d, (const double*)R, logdet, (double*)sign, ws);
    return logabsdet_impl<cplx>(h, n, batch, (const cplx*)L, d, (const cplx*)R, logdet, (cplx*)sign, ws);
}

int sla_inv_IpA(sla_handle_t h, int dtype, int n, int batch, void* G, const void* L, const double* d, const void* R,
                double* logdet, void* sign, void* ws) {
    if (is_f32(dtype)) return f32_inv_IpA(h, dtype, n, batch, G, L, d, R, logdet, sign, ws);
    CHECK_COMMON(h, dtype, n, batch);
    if (!G) return -5; if (!L) return -6; if (!d) return -7; if (!R) return -8; if (!ws) return -11; CLEAR_INFO(h, dtype, n, batch, ws);
    if (dtype == SLA_F64) return inv_IpA_impl<double>(h, n, batch, (double*)G, (const double*)L, d, (const double*)R, logdet, (double*)sign, ws);
    return inv_IpA_impl<cplx>(h, n, batch, (cplx*)G, (const cplx*)L, d, (const cplx*)R, logdet, (cplx*)sign, ws);
}

int sla_ldiv_lu(sla_handle_t h, int dtype, int n, int batch, void* A, void* B, double* logdet, void* sign, void* ws) {
    if (is_f32(dtype)) return f32_lu(h, dtype, n, batch, A, B, logdet, sign, ws, 2);
    CHECK_COMMON(h, dtype, n, batch);
    if (!A) return -5; if (!B) return -6; if (!ws) return -9; CLEAR_INFO(h, dtype, n, batch, ws);
    if (dtype == SLA_F64) return ldiv_lu_impl<double>(h, n, batch, (double*)A, (double*)B, logdet, (double*)sign, ws);
    return ldiv_lu_impl<cplx>(h, n, batch, (cplx*)A, (cplx*)B, logdet, (cplx*)sign, ws);
}

int sla_ldiv_ldr_mat(sla_handle_t h, int dtype, int n, int batch, const void* L, const double* d, const void* R,
                     void* V, void* ws) {
    if (is_f32(dtype)) return f32_dense(h, dtype, n, batch, V, L, d, R, ws, 1);
    CHECK_COMMON(h, dtype, n, batch);
    if (!L) return -5; if (!d) return -6; if (!R) return -7; if (!V) return -8; if (!ws) return -9; CLEAR_INFO(h, dtype, n, batch, ws);
    if (dtype == SLA_F64) return ldiv_ldr_mat_impl<double>(h, n, batch, (const double*)L, d, (const double*)R, (double*)V, ws);
    return ldiv_ldr_mat_impl<cplx>(h, n, batch, (const cplx*)L, d, (const cplx*)R, (cplx*)V, ws);
}

int sla_ldiv_ldr_ldr(sla_handle_t h, int dtype, int n, int batch, const void* UL, const double* Ud, const void* UR,
                     void* VL, double* Vd, void* VR, void* ws) {
    if (is_f32(dtype)) return f32_ldr_ldr(h, dtype, n, batch, (void*)UL, (double*)Ud, (void*)UR, VL, Vd, VR, ws, 1);
    CHECK_COMMON(h, dtype, n, batch);
    UV_NULLCHECK();
    if (!ws) return -11; CLEAR_INFO(h, dtype, n, batch, ws);
    if (UL == VL || UR == VR || Ud == Vd) return -8;
    if (dtype == SLA_F64)
        return ldiv_ldr_ldr_impl<double>(h, n, batch, (const double*)UL, Ud, (const double*)UR, (double*)VL, Vd, (double*)VR, ws);
    return ldiv_ldr_ldr_impl<cplx>(h, n, batch, (const cplx*)UL, Ud, (const cplx*)UR, (cplx*)VL, Vd, (cplx*)VR, ws);
}

int sla_ldiv_mat_ldr(sla_handle_t h, int dtype, int n, int batch, const void* U, long long strideU, void* L, double* d,
                     void* R, void* ws) {
    if (is_f32(dtype)) return f32_mat_ldr(h, dtype, n, batch, U, strideU, L, d, R, ws, 1);
    CHECK_COMMON(h, dtype, n, batch);
    if (!U) return -5; if (!L) return -7; if (!d) return -8; if (!R) return -9; if (!ws) return -10; CLEAR_INFO(h, dtype, n, batch, ws);
    if (dtype == SLA_F64) return ldiv_mat_ldr_impl<double>(h, n, batch, (const double*)U, strideU, (double*)L, d, (double*)R, ws);
    return ldiv_mat_ldr_impl<cplx>(h, n, batch, (const cplx*)U, strideU, (cplx*)L, d, (cplx*)R, ws);
}

int sla_rdiv_mat_ldr(sla_handle_t h, int dtype, int n, int batch, void* U, const void* L, const double* d,
                     const void* R, void* ws) {
    if (is_f32(dtype)) return f32_dense(h, dtype, n, batch, U, L, d, R, ws, 3);
    CHECK_COMMON(h, dtype, n, batch);
    if (!U) return -5; if (!L) return -6; if (!d) return -7; if (!R) return -8; if (!ws) return -9; CLEAR_INFO(h, dtype, n, batch, ws);
    if (dtype == SLA_F64) return rdiv_mat_ldr_impl<double>(h, n, batch, (double*)U, (const double*)L, d, (const double*)R, ws);
    return rdiv_mat_ldr_impl<cplx>(h, n, batch, (cplx*)U, (const cplx*)L, d, (const cplx*)R, ws);
}

int sla_rdiv_ldr_ldr(sla_handle_t h, int dtype, int n, int batch, void* UL, double* Ud, void* UR, const void* VL,
                     const double* Vd, const void* VR, void* ws) {
    if (is_f32(dtype)) return f32_ldr_ldr(h, dtype, n, batch, UL, Ud, UR, (void*)VL, (double*)Vd, (void*)VR, ws, 3);
    CHECK_COMMON(h, dtype, n, batch);
    UV_NULLCHECK();
    if (!ws) return -11; CLEAR_INFO(h, dtype, n, batch, ws);
    if (UL == VL || UR == VR || Ud == Vd) return -8;
    if (dtype == SLA_F64)
        return rdiv_ldr_ldr_impl<double>(h, n, batch, (double*)UL, Ud, (double*)UR, (const double*)VL, Vd, (const double*)VR, ws);
    return rdiv_ldr_ldr_impl<cplx>(h, n, batch, (cplx*)UL, Ud, (cplx*)UR, (const cplx*)VL, Vd, (const cplx*)VR, ws);
}

int sla_rdiv_ldr_mat(sla_handle_t h, int dtype, int n, int batch, void* L, double* d, void* R, const void* V,
                     long long strideV, void* ws) {
    if (is_f32(dtype)) return f32_ldr_mat(h, dtype, n, batch, L, d, R, V, strideV, ws, 1);
    CHECK_COMMON(h, dtype, n, batch);
    if (!L) return -5; if (!d) return -6; if (!R) return -7; if (!V) return -8; if (!ws) return -10; CLEAR_INFO(h, dtype, n, batch, ws);
    if (dtype == SLA_F64) return rdiv_ldr_mat_impl<double>(h, n, batch, (double*)L, d, (double*)R, (const double*)V, strideV, ws);
    return rdiv_ldr_mat_impl<cplx>(h, n, batch, (cplx*)L, d, (cplx*)R, (const cplx*)V, strideV, ws);
}

#define DEFINE_UV(name, impl, which)                                                                               \
    int name(sla_handle_t h, int dtype, int n, int batch, void* G, const void* UL, const double* Ud, const void* UR, \
             const void* VL, const double* Vd, const void* VR, double* logdet, void* sign, void* ws) {             \
        if (is_f32(dtype)) return f32_inv_uv(h, dtype, n, batch, G, UL, Ud, UR, VL, Vd, VR, logdet, sign, ws, which); \
        CHECK_COMMON(h, dtype, n, batch);                                                                          \
        if (!G) return -5;                                                                                         \
        if (!UL) return -6; if (!Ud) return -7; if (!UR) return -8; if (!VL) return -9; if (!Vd) return -10;       \
        if (!VR) return -11; if (!ws) return -14; CLEAR_INFO(h, dtype, n, batch, ws);                                                                  \
        if (dtype == SLA_F64)                                                                                      \
            return impl<double>(h, n, batch, (double*)G, (const double*)UL, Ud, (const double*)UR, (const double*)VL, \
                                Vd, (const double*)VR, logdet, (double*)sign, ws);                                 \
        return impl<cplx>(h, n, batch, (cplx*)G, (const cplx*)UL, Ud, (const cplx*)UR, (const cplx*)VL, Vd,         \
                          (const cplx*)VR, logdet, (cplx*)sign, ws);                                               \
    }
DEFINE_UV(sla_inv_IpUV, inv_IpUV_impl, 0)
DEFINE_UV(sla_inv_UpV, inv_UpV_impl, 1)
DEFINE_UV(sla_inv_invUpV, inv_invUpV_impl, 2)

int sla_gemm(sla_handle_t h, int dtype, int opa, int opb, int m, int n, int k, const void* A, int lda,
             long long strideA, const void* B, int ldb, long long strideB, void* C, int ldc, long long strideC,
             int batch, int accumulate) {
    if (!h) return -1;
    if (!device_is_current(h)) return SLA_EDEVICE;
    if (dtype != SLA_F64 && dtype != SLA_C64) return -2;
    if ((opa != SLA_OP_N && opa != SLA_OP_C)) return -3;
    if ((opb != SLA_OP_N && opb != SLA_OP_C)) return -4;
    if (m < 0) return -5; if (n < 0) return -6; if (k < 0) return -7;
    if (!A) return -8; if (!B) return -11; if (!C) return -14;
    if (batch < 0) return -17;
    if (m == 0 || n == 0 || batch == 0) return 0;
    if (dtype == SLA_F64) {
        GemmArgs<double> g{};
        g.A = (const double*)A; g.B = (const double*)B; g.C = (double*)C;
        g.strideA = strideA; g.strideB = strideB; g.strideC = strideC;
        g.lda = lda; g.ldb = ldb; g.ldc = ldc; g.M = m; g.N = n; g.K = k;
        g.accumulate = accumulate; g.batch = batch;
        CK(gemm_batched<double>(opa, opb, g, h->stream));
    } else {
        GemmArgs<cplx> g{};
        g.A = (const cplx*)A; g.B = (const cplx*)B; g.C = (cplx*)C;
        g.strideA = strideA; g.strideB = strideB; g.strideC = strideC;
        g.lda = lda; g.ldb = ldb; g.ldc = ldc; g.M = m; g.N = n; g.K = k;
        g.accumulate = accumulate; g.batch = batch;
        CK(gemm_batched<cplx>(opa, opb, g, h->stream));
    }
    return 0;
}

int sla_walker_sums(sla_handle_t h, int dtype, int batch, const double* logdet, const void* sign, double* out4) {
    if (!h) return -1;
    if (!device_is_current(h)) return SLA_EDEVICE;
    if (dtype != SLA_F64 && dtype != SLA_C64) return -2;
    if (batch < 0) return -3;
    if (!logdet) return -4; if (!sign) return -5; if (!out4) return -6;
    if (dtype == SLA_F64) CK(walker_sums_batched<double>(out4, logdet, (const double*)sign, batch, h->stream));
    else CK(walker_sums_batched<cplx>(out4, logdet, (const cplx*)sign, batch, h->stream));
    return 0;
}

size_t sla_greens_chain_workspace_bytes(int dtype, int n, int batch, int L, int n_stab) {
    if (is_f32(dtype)) return f32_chain_workspace_bytes(dtype, n, batch, L, n_stab);
    if (n <= 0 || batch <= 0 || L <= 0 || n_stab <= 0 || L % n_stab) return 0;
    if ((dtype != SLA_F64 && dtype != SLA_C64) || !n_supported(dtype, n)) return 0;
    return dtype == SLA_F64 ? ChainWs<double>::bytes(n, batch, L, n_stab) : ChainWs<cplx>::bytes(n, batch, L, n_stab);
}

int* sla_greens_chain_workspace_info(void* ws, int dtype, int n, int batch, int L, int n_stab) {
    if (is_f32(dtype)) dtype -= 2;
    if (!ws || n <= 0 || batch <= 0 || L <= 0 || n_stab <= 0 || L % n_stab) return nullptr;
    if (dtype == SLA_F64) return Ws<double>(ChainWs<double>(ws, n, batch, L, n_stab).ws, n, batch).info;
    return Ws<cplx>(ChainWs<cplx>(ws, n, batch, L, n_stab).ws, n, batch).info;
}

int sla_greens_chain(sla_handle_t h, int dtype, int n, int batch, int L, int n_stab, const void* expK,
                     const signed char* fields, double lam, int inverse, void* G, double* logdet, void* sign,
                     void* Lout, double* dout, void* Rout, void* ws, size_t ws_bytes) {
    if (is_f32(dtype)) return f32_greens_chain(h, dtype, n, batch, L, n_stab, expK, fields, lam, inverse, G, logdet, sign, Lout, dout, Rout, ws, ws_bytes);
    CHECK_COMMON(h, dtype, n, batch);
    if (L <= 0) return -5;
    if (n_stab <= 0 || L % n_stab) return -6;
    if (!expK) return -7; if (!fields) return -8;
    if (inverse != SLA_INV_IPA && inverse != SLA_INV_IPUV) return -10;
    if (inverse == SLA_INV_IPUV && (L / n_stab) < 2) return -10;
    if (!G) return -11; if (!logdet) return -12; if (!sign) return -13; if (!ws) return -17;
    auto direct = [&]() -> int {
        if (dtype == SLA_F64)
            return greens_chain_impl<double>(h, n, batch, L, n_stab, (const double*)expK, fields, lam, inverse, (double*)G,
                                             logdet, (double*)sign, (double*)Lout, dout, (double*)Rout, ws, ws_bytes);
        return greens_chain_impl<cplx>(h, n, batch, L, n_stab, (const cplx*)expK, fields, lam, inverse, (cplx*)G, logdet,
                                       (cplx*)sign, (cplx*)Lout, dout, (cplx*)Rout, ws, ws_bytes);
    };
    static const bool graphs_on = !(getenv("SLA_GRAPH") && atoi(getenv("SLA_GRAPH")) == 0);
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (!graphs_on || h->graph_broken || h->stream == nullptr /* legacy stream: not capturable */ ||
        cudaStreamIsCapturing(h->stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone)
        return direct();
    // ---- replay a graph captured for exactly these arguments, or capture one now
    unsigned long long key[20] = {};
    {
        int i = 0;
        key[i++] = (unsigned long long)dtype; key[i++] = (unsigned long long)n; key[i++] = (unsigned long long)batch;
        key[i++] = (unsigned long long)L; key[i++] = (unsigned long long)n_stab; key[i++] = (unsigned long long)(uintptr_t)expK;
        key[i++] = (unsigned long long)(uintptr_t)fields; memcpy(&key[i++], &lam, sizeof(double)); key[i++] = (unsigned long long)inverse;
        key[i++] = (unsigned long long)(uintptr_t)G; key[i++] = (unsigned long long)(uintptr_t)logdet; key[i++] = (unsigned long long)(uintptr_t)sign;
        key[i++] = (unsigned long long)(uintptr_t)Lout; key[i++] = (unsigned long long)(uintptr_t)dout; key[i++] = (unsigned long long)(uintptr_t)Rout;
        key[i++] = (unsigned long long)(uintptr_t)ws; key[i++] = (unsigned long long)ws_bytes; key[i++] = (unsigned long long)(uintptr_t)h->stream;
    }
    sla_handle_s::ChainGraph* g = nullptr;
    for (auto& e : h->graphs)
        if (e.used && memcmp(e.key, key, sizeof(key)) == 0) { g = &e; break; }
    if (g == nullptr) {
        if (ensure_side_stream(h) != 0) return direct();
        sla_handle_s::ChainGraph& e = h->graphs[h->graph_next];
        h->graph_next = (h->graph_next + 1) % 4;
        if (e.exec) { cudaGraphExecDestroy(e.exec); e.exec = nullptr; }
        e.used = false;
        const unsigned long long k0 = g_launch_count.load();
        unsigned long long a0[SLA_N_LDR_ALGOS];
        for (int i = 0; i < SLA_N_LDR_ALGOS; ++i) a0[i] = h->ldr_algo_launches[i];
        if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); h->graph_broken = true; return direct(); }
        const int rc = direct();
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
        // the launches were recorded, not executed: take them out of the counters again (a replay adds them back)
        const unsigned long long kernels = g_launch_count.load() - k0;
        g_launch_count -= kernels;
        for (int i = 0; i < SLA_N_LDR_ALGOS; ++i) { e.algo_launches[i] = h->ldr_algo_launches[i] - a0[i]; h->ldr_algo_launches[i] = a0[i]; }
        if (rc != 0 || ce != cudaSuccess || graph == nullptr) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            if (rc < 0 || rc == SLA_ENOTIMPL) return rc;     // an argument error: nothing was launched, report it
            h->graph_broken = true;
            return direct();
        }
        const cudaError_t ie = cudaGraphInstantiate(&e.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) { cudaGetLastError(); e.exec = nullptr; h->graph_broken = true; return direct(); }
        memcpy(e.key, key, sizeof(key));
        e.kernels = kernels;
        e.last_algo = h->last_ldr_algo;
        e.used = true;
        g = &e;
    }
    CK(cudaGraphLaunch(g->exec, h->stream));
    g_launch_count += g->kernels;
    for (int i = 0; i < SLA_N_LDR_ALGOS; ++i) h->ldr_algo_launches[i] += g->algo_launches[i];
    h->last_ldr_algo = g->last_algo;
    return 0;
}

size_t sla_sweep_workspace_bytes(int dtype, int n, int batch, int L, int n_stab) {
    if (is_f32(dtype)) return f32_sweep_workspace_bytes(dtype, n, batch, L, n_stab);
    if (n <= 0 || batch <= 0 || L <= 0 || n_stab <= 0 || L % n_stab) return 0;
    if ((dtype != SLA_F64 && dtype != SLA_C64) || !n_supported(dtype, n)) return 0;
    return dtype == SLA_F64 ? SweepWs<double>::bytes(n, batch, L, n_stab) : SweepWs<cplx>::bytes(n, batch, L, n_stab);
}

int* sla_sweep_workspace_info(void* ws, int dtype, int n, int batch, int L, int n_stab) {
    if (is_f32(dtype)) dtype -= 2;
    if (!ws || n <= 0 || batch <= 0 || L <= 0 || n_stab <= 0 || L % n_stab) return nullptr;
    if (dtype == SLA_F64) return Ws<double>(SweepWs<double>(ws, n, batch, L, n_stab).ws, n, batch).info;
    return Ws<cplx>(SweepWs<cplx>(ws, n, batch, L, n_stab).ws, n, batch).info;
}

int sla_sweep(sla_handle_t h, int dtype, int n, int batch, int L, int n_stab, const void* expK, const void* expKinv,
              const signed char* fields, double lam, void* G0, void* G, double* logdet, void* sign, double* dG, void* ws,
              size_t ws_bytes) {
    if (is_f32(dtype)) return f32_sweep(h, dtype, n, batch, L, n_stab, expK, expKinv, fields, lam, G0, G, logdet, sign, dG, ws, ws_bytes);
    CHECK_COMMON(h, dtype, n, batch);
    if (L <= 0) return -5;
    if (n_stab <= 0 || L % n_stab) return -6;
    if (!expK) return -7; if (!expKinv) return -8; if (!fields) return -9;
    if (!G) return -12; if (!logdet) return -13; if (!sign) return -14; if (!ws) return -16;
    if (dtype == SLA_F64)
        return sweep_impl<double>(h, n, batch, L, n_stab, (const double*)expK, (const double*)expKinv, fields, lam,
                                  (double*)G0, (double*)G, logdet, (double*)sign, dG, ws, ws_bytes);
    return sweep_impl<cplx>(h, n, batch, L, n_stab, (const cplx*)expK, (const cplx*)expKinv, fields, lam, (cplx*)G0,
                            (cplx*)G, logdet, (cplx*)sign, dG, ws, ws_bytes);
}

}  // extern "C"
