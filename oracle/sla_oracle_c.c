/* CPU oracle in C: restatement of the reference's hot path over the same Fortran-ABI ILP64 LAPACK/BLAS
 * entry points the reference ccall's (src/qr.jl:51-105, src/lu.jl:44-110, every mul!), here from the
 * OpenBLAS bundled with numpy (symbols scipy_?geqp3_64_ ...).
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/ (cross-check against the Python restatement) and by the
 * cpu_baseline / --impl reference legs of bench.py.  Never linked or called by the product.
 *
 * Functions follow the reference line by line:
 *   ldr          src/LDR.jl:160-188        lmul_mat_ldr   src/overloaded_functions.jl:126-145
 *   det_lu/inv_lu src/lu.jl:123-158        inv_IpA        src/exported_functions.jl:25-71
 *   inv_IpUV     src/exported_functions.jl:97-177
 *   eval_greens  the driver loop of test/runtests.jl:93-100,151-160 on Hubbard-like B matrices
 * Parity is pinned through tests/test_oracle_c.py (agreement with oracle/sla_oracle.py, which is pinned
 * by the reference's analytic tests).
 */
#include <complex.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t blasint;
typedef double complex zc;

extern void scipy_dgeqp3_64_(blasint*, blasint*, double*, blasint*, blasint*, double*, double*, blasint*, blasint*);
extern void scipy_dorgqr_64_(blasint*, blasint*, blasint*, double*, blasint*, double*, double*, blasint*, blasint*);
extern void scipy_dgetrf_64_(blasint*, blasint*, double*, blasint*, blasint*, blasint*);
extern void scipy_dgetri_64_(blasint*, double*, blasint*, blasint*, double*, blasint*, blasint*);
extern void scipy_dgemm_64_(const char*, const char*, blasint*, blasint*, blasint*, double*, double*, blasint*,
                            double*, blasint*, double*, double*, blasint*);
extern void scipy_zgeqp3_64_(blasint*, blasint*, zc*, blasint*, blasint*, zc*, zc*, blasint*, double*, blasint*);
extern void scipy_zungqr_64_(blasint*, blasint*, blasint*, zc*, blasint*, zc*, zc*, blasint*, blasint*);
extern void scipy_zgetrf_64_(blasint*, blasint*, zc*, blasint*, blasint*, blasint*);
extern void scipy_zgetri_64_(blasint*, zc*, blasint*, blasint*, zc*, blasint*, blasint*);
extern void scipy_zgemm_64_(const char*, const char*, blasint*, blasint*, blasint*, zc*, zc*, blasint*, zc*,
                            blasint*, zc*, zc*, blasint*);
extern void scipy_openblas_set_num_threads64_(int);
extern char* scipy_openblas_get_config64_(void);

/* ---- generic-by-macro implementation: T = double / double complex ---------------------------- */
#define DEFINE_ORACLE(SFX, T, GEQP3_CALL, ORGQR, GETRF, GETRI, GEMM, ABS, CONJ, ISCPLX)                          \
    typedef struct {                                                                                             \
        blasint n, lwork;                                                                                        \
        T *M, *Mp, *Mpp, *work, *tau;                                                                            \
        double *v, *rwork;                                                                                       \
        blasint *jpvt, *ipiv;                                                                                    \
    } ws_##SFX;                                                                                                  \
    static ws_##SFX* ws_new_##SFX(blasint n) {                                                                   \
        ws_##SFX* w = (ws_##SFX*)calloc(1, sizeof(ws_##SFX));                                                    \
        w->n = n;                                                                                                \
        w->lwork = 2 * n + (n + 1) * 64; /* >= geqp3 query (2n+(n+1)*32, src/qr.jl:36-67) and getri n*NB */      \
        w->M = (T*)malloc(sizeof(T) * n * n); w->Mp = (T*)malloc(sizeof(T) * n * n);                             \
        w->Mpp = (T*)malloc(sizeof(T) * n * n);                                                                  \
        w->work = (T*)malloc(sizeof(T) * w->lwork); w->tau = (T*)malloc(sizeof(T) * n);                          \
        w->v = (double*)malloc(sizeof(double) * n); w->rwork = (double*)malloc(sizeof(double) * 2 * n);          \
        w->jpvt = (blasint*)malloc(sizeof(blasint) * n); w->ipiv = (blasint*)malloc(sizeof(blasint) * n);        \
        return w;                                                                                                \
    }                                                                                                            \
    static void ws_free_##SFX(ws_##SFX* w) {                                                                     \
        free(w->M); free(w->Mp); free(w->Mpp); free(w->work); free(w->tau); free(w->v); free(w->rwork);          \
        free(w->jpvt); free(w->ipiv); free(w);                                                                   \
    }                                                                                                            \
    static void gemm_##SFX(blasint n, const char* ta, const char* tb, T* A, T* B, T* C) {                        \
        T one = 1.0, zero = 0.0;                                                                                 \
        GEMM(ta, tb, &n, &n, &n, &one, A, &n, B, &n, &zero, C, &n);                                              \
    }                                                                                                            \
    /* ldr!(F, ws): src/LDR.jl:160-188 */                                                                        \
    static void ldr_##SFX(blasint n, T* L, double* d, T* R, ws_##SFX* w) {                                       \
        blasint info = 0;                                                                                        \
        memset(w->jpvt, 0, sizeof(blasint) * n);                     /* qr.jl:77 */                              \
        GEQP3_CALL;                                                  /* LDR.jl:166 */                            \
        T* M = w->M;                                                                                             \
        for (blasint j = 0; j < n; ++j)                              /* LDR.jl:169-170 triu */                   \
            for (blasint i = 0; i < n; ++i) M[i + j * n] = (i <= j) ? L[i + j * n] : 0.0;                        \
        for (blasint i = 0; i < n; ++i) d[i] = ABS(M[i + i * n]);    /* :173-175 */                              \
        for (blasint j = 0; j < n; ++j)                              /* :178-179 */                              \
            for (blasint i = 0; i < n; ++i) M[i + j * n] /= d[i];                                                \
        for (blasint j = 0; j < n; ++j)                              /* :182 R[:,jpvt] = M */                    \
            memcpy(R + (w->jpvt[j] - 1) * n, M + j * n, sizeof(T) * n);                                          \
        ORGQR(&n, &n, &n, L, &n, w->tau, w->work, &w->lwork, &info); /* :185 */                                  \
    }                                                                                                            \
    /* det_lu!(A, ws): src/lu.jl:123-140 */                                                                      \
    static void det_lu_##SFX(blasint n, T* A, ws_##SFX* w, double* logdet, T* sgn) {                             \
        blasint info = 0;                                                                                        \
        memset(w->ipiv, 0, sizeof(blasint) * n);                                                                 \
        GETRF(&n, &n, A, &n, w->ipiv, &info);                                                                    \
        double ld = 0.0;                                                                                         \
        T s = 1.0;                                                                                               \
        for (blasint i = 0; i < n; ++i) {                                                                        \
            T a = A[i + i * n];                                                                                  \
            ld += log(ABS(a));                                                                                   \
            s *= a / ABS(a);                                                                                     \
            if (w->ipiv[i] != i + 1) s = -s;                                                                     \
        }                                                                                                        \
        *logdet = ld;                                                                                            \
        *sgn = s;                                                                                                \
    }                                                                                                            \
    /* inv_lu!(A, ws): src/lu.jl:149-158 */                                                                      \
    static void inv_lu_##SFX(blasint n, T* A, ws_##SFX* w, double* logdet, T* sgn) {                             \
        blasint info = 0;                                                                                        \
        det_lu_##SFX(n, A, w, logdet, sgn);                                                                      \
        GETRI(&n, A, &n, w->ipiv, w->work, &w->lwork, &info);                                                    \
        *logdet = -*logdet;                                                                                      \
        *sgn = CONJ(*sgn);                                                                                       \
    }                                                                                                            \
    /* lmul!(U::Matrix, V::LDR, ws): src/overloaded_functions.jl:126-145 */                                      \
    static void lmul_##SFX(blasint n, T* U, T* L, double* d, T* R, ws_##SFX* w) {                                \
        memcpy(w->Mp, R, sizeof(T) * n * n);                         /* :129-130 */                              \
        gemm_##SFX(n, "N", "N", U, L, w->M);                         /* :133 */                                  \
        for (blasint j = 0; j < n; ++j)                              /* :134-135 */                              \
            for (blasint i = 0; i < n; ++i) L[i + j * n] = w->M[i + j * n] * d[j];                               \
        ldr_##SFX(n, L, d, R, w);                                    /* :138 */                                  \
        gemm_##SFX(n, "N", "N", R, w->Mp, w->M);                     /* :141 */                                  \
        memcpy(R, w->M, sizeof(T) * n * n);                          /* :142 */                                  \
    }                                                                                                            \
    /* inv_IpA!(G, A, ws): src/exported_functions.jl:25-71 */                                                    \
    static void inv_IpA_##SFX(blasint n, T* G, T* L, double* d, T* R, ws_##SFX* w, double* logdet, T* sgn) {     \
        double ldR, ldM, ldDp = 0.0;                                                                             \
        T sR, sM;                                                                                                \
        memcpy(w->Mp, R, sizeof(T) * n * n);                                                                     \
        inv_lu_##SFX(n, w->Mp, w, &ldR, &sR);                        /* :32-34 */                                \
        for (blasint j = 0; j < n; ++j) {                                                                        \
            double dm = fmin(d[j], 1.0), dp = fmax(d[j], 1.0);       /* :37-46 */                                \
            ldDp += log(dp);                                         /* :49-50 */                                \
            for (blasint i = 0; i < n; ++i) {                                                                    \
                w->Mp[i + j * n] /= dp;                              /* :53-54 */                                \
                w->M[i + j * n] = L[i + j * n] * dm + w->Mp[i + j * n]; /* :41-42, :57 */                        \
            }                                                                                                    \
        }                                                                                                        \
        inv_lu_##SFX(n, w->M, w, &ldM, &sM);                         /* :60-61 */                                \
        gemm_##SFX(n, "N", "N", w->Mp, w->M, G);                     /* :64 */                                   \
        *sgn = sR * sM;                                              /* :67 */                                   \
        *logdet = -ldDp + ldM;                                       /* :68 */                                   \
    }                                                                                                            \
    /* inv_IpUV!(G, U, V, ws): src/exported_functions.jl:97-177 */                                               \
    static void inv_IpUV_##SFX(blasint n, T* G, T* Lu, double* du, T* Ru, T* Lv, double* dv, T* Rv, ws_##SFX* w, \
                               double* logdet, T* sgn) {                                                         \
        double ldLu, ldRv, ldM, ldDvp = 0.0, ldDup = 0.0;                                                        \
        T sLu, sRv, sM;                                                                                          \
        memcpy(w->M, Lu, sizeof(T) * n * n);                                                                     \
        det_lu_##SFX(n, w->M, w, &ldLu, &sLu);                       /* :107-108 */                              \
        memcpy(w->Mp, Rv, sizeof(T) * n * n);                                                                    \
        inv_lu_##SFX(n, w->Mp, w, &ldRv, &sRv);                      /* :111-113 */                              \
        for (blasint j = 0; j < n; ++j) {                                                                        \
            double dp = fmax(dv[j], 1.0);                                                                        \
            ldDvp += log(dp);                                        /* :116-121 */                              \
            for (blasint i = 0; i < n; ++i) w->Mp[i + j * n] /= dp;  /* :124 */                                  \
        }                                                                                                        \
        for (blasint i = 0; i < n; ++i) ldDup += log(fmax(du[i], 1.0)); /* :128-133 */                           \
        for (blasint j = 0; j < n; ++j)                              /* :136-137 */                              \
            for (blasint i = 0; i < n; ++i) w->M[i + j * n] = CONJ(Lu[j + i * n]) / fmax(du[i], 1.0);            \
        gemm_##SFX(n, "N", "N", Ru, Lv, G);                          /* :146 */                                  \
        for (blasint j = 0; j < n; ++j)                              /* :147, :155 */                            \
            for (blasint i = 0; i < n; ++i) G[i + j * n] = (fmin(du[i], 1.0) * G[i + j * n]) * fmin(dv[j], 1.0); \
        gemm_##SFX(n, "N", "N", w->M, w->Mp, w->Mpp);                /* :158 */                                  \
        for (blasint k = 0; k < n * n; ++k) G[k] += w->Mpp[k];       /* :161-162 */                              \
        inv_lu_##SFX(n, G, w, &ldM, &sM);                            /* :165-166 */                              \
        gemm_##SFX(n, "N", "N", G, w->M, w->Mpp);                    /* :169 */                                  \
        gemm_##SFX(n, "N", "N", w->Mp, w->Mpp, G);                   /* :170 */                                  \
        *sgn = sRv * sM * CONJ(sLu);                                 /* :173 */                                  \
        *logdet = -ldDvp + ldM - ldDup;                              /* :174 */                                  \
    }                                                                                                            \
    /* one evaluation for one chain; out: G (n*n), logdet, sgn, and the final (L,d,R) if wanted */               \
    static void eval_greens_##SFX(blasint n, blasint L_, blasint n_stab, const T* expK, const signed char* f,    \
                                  double lam, int inverse, T* G, double* logdet, T* sgn, double* d_out) {        \
        ws_##SFX* w = ws_new_##SFX(n);                                                                           \
        size_t nn = (size_t)n * n;                                                                               \
        T* B = (T*)malloc(sizeof(T) * nn); T* Bbar = (T*)malloc(sizeof(T) * nn); T* tmp = (T*)malloc(sizeof(T) * nn); \
        T* FL = (T*)calloc(nn, sizeof(T)); T* FR = (T*)calloc(nn, sizeof(T)); double* Fd = (double*)malloc(sizeof(double) * n); \
        T* UL = (T*)calloc(nn, sizeof(T)); T* UR = (T*)calloc(nn, sizeof(T)); double* Ud = (double*)malloc(sizeof(double) * n); \
        for (blasint i = 0; i < n; ++i) { FL[i + i * n] = 1.0; FR[i + i * n] = 1.0; Fd[i] = 1.0;                 \
                                          UL[i + i * n] = 1.0; UR[i + i * n] = 1.0; Ud[i] = 1.0; }               \
        blasint ng = L_ / n_stab;                                                                                \
        double ep = exp(lam), em = exp(-lam);                                                                    \
        for (blasint g = 0; g < ng; ++g) {                                                                       \
            for (blasint k = 0; k < (blasint)nn; ++k) Bbar[k] = 0.0;                                             \
            for (blasint i = 0; i < n; ++i) Bbar[i + i * n] = 1.0;                                               \
            for (blasint l = g * n_stab; l < (g + 1) * n_stab; ++l) { /* test/runtests.jl:93-100 */              \
                for (blasint j = 0; j < n; ++j)                                                                  \
                    for (blasint i = 0; i < n; ++i) B[i + j * n] = (f[l * n + i] > 0 ? ep : em) * expK[i + j * n]; \
                gemm_##SFX(n, "N", "N", B, Bbar, tmp);                                                           \
                memcpy(Bbar, tmp, sizeof(T) * nn);                                                               \
            }                                                                                                    \
            if (inverse == 1 && g >= ng / 2) lmul_##SFX(n, Bbar, UL, Ud, UR, w);                                 \
            else lmul_##SFX(n, Bbar, FL, Fd, FR, w);                 /* runtests.jl:154-156 */                   \
        }                                                                                                        \
        if (inverse == 0) inv_IpA_##SFX(n, G, FL, Fd, FR, w, logdet, sgn);                                       \
        else inv_IpUV_##SFX(n, G, UL, Ud, UR, FL, Fd, FR, w, logdet, sgn);                                       \
        if (d_out) memcpy(d_out, Fd, sizeof(double) * n);                                                        \
        free(B); free(Bbar); free(tmp); free(FL); free(FR); free(Fd); free(UL); free(UR); free(Ud);              \
        ws_free_##SFX(w);                                                                                        \
    }

#define ABS_D(x) fabs(x)
#define CONJ_D(x) (x)
#define ABS_Z(x) cabs(x)
#define CONJ_Z(x) conj(x)

DEFINE_ORACLE(d, double,
              scipy_dgeqp3_64_(&n, &n, L, &n, w->jpvt, w->tau, w->work, &w->lwork, &info),
              scipy_dorgqr_64_, scipy_dgetrf_64_, scipy_dgetri_64_, scipy_dgemm_64_, ABS_D, CONJ_D, 0)
DEFINE_ORACLE(z, zc,
              scipy_zgeqp3_64_(&n, &n, L, &n, w->jpvt, w->tau, w->work, &w->lwork, w->rwork, &info),
              scipy_zungqr_64_, scipy_zgetrf_64_, scipy_zgetri_64_, scipy_zgemm_64_, ABS_Z, CONJ_Z, 1)

/* ---- exported entry points -------------------------------------------------------------------- */
typedef struct {
    int dtype, inverse, tid, nthreads;
    blasint n, L, n_stab, chains;
    const void* expK;
    const signed char* fields;
    double lam;
    void* G;
    double* logdet;
    void* sgn;
} job_t;

static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    size_t nn = (size_t)j->n * j->n;
    for (blasint c = j->tid; c < j->chains; c += j->nthreads) { /* chains dealt round-robin */
        const signed char* f = j->fields + (size_t)c * j->L * j->n;
        if (j->dtype == 0)
            eval_greens_d(j->n, j->L, j->n_stab, (const double*)j->expK, f, j->lam, j->inverse,
                          (double*)j->G + c * nn, j->logdet + c, (double*)j->sgn + c, NULL);
        else
            eval_greens_z(j->n, j->L, j->n_stab, (const zc*)j->expK, f, j->lam, j->inverse, (zc*)j->G + c * nn,
                          j->logdet + c, (zc*)j->sgn + c, NULL);
    }
    return NULL;
}

/* dtype 0 = Float64, 1 = ComplexF64; inverse 0 = inv_IpA!, 1 = inv_IpUV!; column-major matrices;
 * fields int8 [chains][L][n]; nthreads worker threads x blas_threads BLAS threads each (the fair DQMC arrangement is
 * one chain per core x 1 BLAS thread; 1 worker x nproc BLAS threads is reported beside it for transparency). */
int oracle_eval_batch_bt(int dtype, long n, long L, long n_stab, long chains, const void* expK,
                         const signed char* fields, double lam, int inverse, void* G, double* logdet, void* sgn,
                         int nthreads, int blas_threads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > chains) nthreads = (int)chains;
    scipy_openblas_set_num_threads64_(blas_threads < 1 ? 1 : blas_threads);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    job_t* jobs = (job_t*)malloc(sizeof(job_t) * nthreads);
    for (int t = 0; t < nthreads; ++t) {
        job_t j = {dtype, inverse, t, nthreads, n, L, n_stab, chains, expK, fields, lam, G, logdet, sgn};
        jobs[t] = j;
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
    return 0;
}

int oracle_eval_batch(int dtype, long n, long L, long n_stab, long chains, const void* expK,
                      const signed char* fields, double lam, int inverse, void* G, double* logdet, void* sgn,
                      int nthreads) {
    return oracle_eval_batch_bt(dtype, n, L, n_stab, chains, expK, fields, lam, inverse, G, logdet, sgn, nthreads, 1);
}

const char* oracle_blas_config(void) { return scipy_openblas_get_config64_(); }
