/* A plain-C client of libsla_b200.so: calls the entry points with exactly the argument types the Julia shim's
 * `ccall` tuples declare (stablelinearalgebra.jl_b200/julia/StableLinearAlgebraB200.jl) -- Ptr{Cvoid} handle,
 * Cint sizes, CuPtr device pointers, Clonglong strides, Csize_t blob sizes -- so that a marshalling mistake in
 * the header shows up without Julia.  Compiled with gcc (NOT nvcc) against include/sla_b200.h and driven by
 * tests/test_abi.py::test_c_client (-m gpu).  Checks are self-contained properties, no oracle:
 *   1. sla_greens_chain with expK = I: every B_l is diagonal, so G = diag(1 / (1 + exp(lam * sum_l s_l,i))),
 *      log|det G| = -sum_i log(1 + exp(..)), sign = +1                   (test/runtests.jl:151-160 route 1);
 *   2. sla_ldr_from + sla_ldr_to_mat: L D R reproduces a dense random A  (test/runtests.jl:114-126);
 *   3. sla_inv_IpA + sla_gemm: G (I + A) = I                             (src/exported_functions.jl:25-71);
 *   4. sla_logabsdet(F) returns a finite log|det| and a unit sign        (overloaded_functions.jl:721-733).
 * Prints "ABI_C_OK" and exits 0 on success. */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sla_b200.h"

#define CU(x)                                                                      \
    do {                                                                           \
        cudaError_t e__ = (x);                                                     \
        if (e__ != cudaSuccess) {                                                  \
            fprintf(stderr, "CUDA error %d at %s:%d\n", (int)e__, __FILE__, __LINE__); \
            return 2;                                                              \
        }                                                                          \
    } while (0)
#define SLA(x)                                                                     \
    do {                                                                           \
        int r__ = (x);                                                             \
        if (r__ != 0) {                                                            \
            fprintf(stderr, "sla error %d at %s:%d\n", r__, __FILE__, __LINE__);   \
            return 3;                                                              \
        }                                                                          \
    } while (0)

static unsigned long long lcg_state = 88172645463325252ull;
static double lcg_uniform(void) {
    lcg_state = lcg_state * 6364136223846793005ull + 1442695040888963407ull;
    return (double)(lcg_state >> 11) / 9007199254740992.0;
}

static int check_info(const int* dinfo, int batch, const char* what) {
    int host[64];
    if (cudaMemcpy(host, dinfo, (size_t)batch * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
    for (int b = 0; b < batch; ++b)
        if (host[b] != 0) { fprintf(stderr, "%s: info[%d] = %d\n", what, b, host[b]); return 1; }
    return 0;
}

int main(void) {
    const int n = 48, batch = 5, L = 20, n_stab = 5;
    const double lam = 0.37;
    const size_t nn = (size_t)n * n;
    cudaStream_t stream;
    CU(cudaSetDevice(0));
    CU(cudaStreamCreate(&stream));
    sla_handle_t h = NULL;
    SLA(sla_create(&h, (void*)stream));
    if (sla_version() < 100) return 4;
    if (sla_max_n(SLA_F64) < 1024) return 4;

    /* ---------------------------------------------------------------- 1. fused chain, closed form */
    {
        double* expK = (double*)calloc(nn, sizeof(double));
        signed char* fields = (signed char*)malloc((size_t)batch * L * n);
        for (int i = 0; i < n; ++i) expK[(size_t)i * n + i] = 1.0;
        for (size_t i = 0; i < (size_t)batch * L * n; ++i) fields[i] = lcg_uniform() < 0.5 ? -1 : 1;
        const size_t blob_bytes = sla_greens_chain_workspace_bytes(SLA_F64, n, batch, L, n_stab);
        if (blob_bytes == 0) return 5;
        void *d_expK, *d_fields, *d_G, *d_logdet, *d_sign, *d_blob;
        CU(cudaMalloc(&d_expK, nn * 8));
        CU(cudaMalloc(&d_fields, (size_t)batch * L * n));
        CU(cudaMalloc(&d_G, nn * 8 * batch));
        CU(cudaMalloc(&d_logdet, 8 * batch));
        CU(cudaMalloc(&d_sign, 8 * batch));
        CU(cudaMalloc(&d_blob, blob_bytes));
        CU(cudaMemcpyAsync(d_expK, expK, nn * 8, cudaMemcpyHostToDevice, stream));
        CU(cudaMemcpyAsync(d_fields, fields, (size_t)batch * L * n, cudaMemcpyHostToDevice, stream));
        SLA(sla_greens_chain(h, SLA_F64, n, batch, L, n_stab, d_expK, (const signed char*)d_fields, lam, SLA_INV_IPA,
                             d_G, (double*)d_logdet, d_sign, NULL, NULL, NULL, d_blob, blob_bytes));
        double* G = (double*)malloc(nn * 8 * batch);
        double logdet[8], sign[8];
        CU(cudaMemcpyAsync(G, d_G, nn * 8 * batch, cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(logdet, d_logdet, 8 * batch, cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(sign, d_sign, 8 * batch, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        if (check_info(sla_greens_chain_workspace_info(d_blob, SLA_F64, n, batch, L, n_stab), batch, "greens_chain")) return 6;
        for (int b = 0; b < batch; ++b) {
            double ld = 0.0, err = 0.0;
            for (int i = 0; i < n; ++i) {
                int ssum = 0;
                for (int l = 0; l < L; ++l) ssum += fields[((size_t)b * L + l) * n + i];
                const double g = 1.0 / (1.0 + exp(lam * ssum));
                ld -= log(1.0 + exp(lam * ssum));
                for (int j = 0; j < n; ++j) {
                    const double want = (i == j) ? g : 0.0;
                    const double e = fabs(G[(size_t)b * nn + (size_t)j * n + i] - want);
                    if (e > err) err = e;
                }
            }
            if (!(err < 1e-12) || !(fabs(logdet[b] - ld) < 1e-10 * fmax(1.0, fabs(ld))) || sign[b] != 1.0) {
                fprintf(stderr, "chain %d: err %.3e logdet %.15g want %.15g sign %g\n", b, err, logdet[b], ld, sign[b]);
                return 7;
            }
        }
        cudaFree(d_expK); cudaFree(d_fields); cudaFree(d_G); cudaFree(d_logdet); cudaFree(d_sign); cudaFree(d_blob);
        free(expK); free(fields); free(G);
    }

    /* ---------------------------------------------------------------- 2.-4. op-level entry points */
    {
        const size_t ws_bytes = sla_workspace_bytes(SLA_F64, n, batch);
        if (ws_bytes == 0) return 8;
        double* A = (double*)malloc(nn * 8 * batch);
        for (size_t i = 0; i < nn * batch; ++i) A[i] = lcg_uniform() - 0.5;
        void *d_A, *d_L, *d_d, *d_R, *d_B, *d_G, *d_P, *d_ws, *d_logdet, *d_sign, *d_logdet2, *d_sign2;
        CU(cudaMalloc(&d_A, nn * 8 * batch));
        CU(cudaMalloc(&d_L, nn * 8 * batch));
        CU(cudaMalloc(&d_R, nn * 8 * batch));
        CU(cudaMalloc(&d_B, nn * 8 * batch));
        CU(cudaMalloc(&d_G, nn * 8 * batch));
        CU(cudaMalloc(&d_P, nn * 8 * batch));
        CU(cudaMalloc(&d_d, (size_t)n * 8 * batch));
        CU(cudaMalloc(&d_ws, ws_bytes));
        CU(cudaMalloc(&d_logdet, 8 * batch));
        CU(cudaMalloc(&d_sign, 8 * batch));
        CU(cudaMalloc(&d_logdet2, 8 * batch));
        CU(cudaMalloc(&d_sign2, 8 * batch));
        CU(cudaMemcpyAsync(d_A, A, nn * 8 * batch, cudaMemcpyHostToDevice, stream));
        SLA(sla_ldr_from(h, SLA_F64, n, batch, d_A, (long long)nn, d_L, (double*)d_d, d_R, d_ws));
        SLA(sla_ldr_to_mat(h, SLA_F64, n, batch, d_B, d_L, (const double*)d_d, d_R, d_ws));
        double* B = (double*)malloc(nn * 8 * batch);
        CU(cudaMemcpyAsync(B, d_B, nn * 8 * batch, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        if (check_info(sla_workspace_info(d_ws, SLA_F64, n, batch), batch, "ldr_from")) return 9;
        double err = 0.0;
        for (size_t i = 0; i < nn * batch; ++i) err = fmax(err, fabs(B[i] - A[i]));
        if (!(err < 1e-13 * n)) { fprintf(stderr, "L D R != A: %.3e\n", err); return 10; }

        SLA(sla_inv_IpA(h, SLA_F64, n, batch, d_G, d_L, (const double*)d_d, d_R, (double*)d_logdet, d_sign, d_ws));
        /* P = G * A, then G + P = G (I + A) must be the identity */
        SLA(sla_gemm(h, SLA_F64, SLA_OP_N, SLA_OP_N, n, n, n, d_G, n, (long long)nn, d_A, n, (long long)nn, d_P, n,
                     (long long)nn, batch, 0));
        SLA(sla_logabsdet(h, SLA_F64, n, batch, d_L, (const double*)d_d, d_R, (double*)d_logdet2, d_sign2, d_ws));
        double* G = (double*)malloc(nn * 8 * batch);
        double* P = (double*)malloc(nn * 8 * batch);
        double logdet[8], sign[8], logdet2[8], sign2[8];
        CU(cudaMemcpyAsync(G, d_G, nn * 8 * batch, cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(P, d_P, nn * 8 * batch, cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(logdet, d_logdet, 8 * batch, cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(sign, d_sign, 8 * batch, cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(logdet2, d_logdet2, 8 * batch, cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(sign2, d_sign2, 8 * batch, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        if (check_info(sla_workspace_info(d_ws, SLA_F64, n, batch), batch, "inv_IpA/logabsdet")) return 11;
        err = 0.0;
        for (int b = 0; b < batch; ++b)
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < n; ++i) {
                    const size_t o = (size_t)b * nn + (size_t)j * n + i;
                    err = fmax(err, fabs(G[o] + P[o] - (i == j ? 1.0 : 0.0)));
                }
        if (!(err < 1e-10)) { fprintf(stderr, "G (I + A) != I: %.3e\n", err); return 12; }
        for (int b = 0; b < batch; ++b) {
            if (fabs(fabs(sign[b]) - 1.0) > 0.0 || fabs(fabs(sign2[b]) - 1.0) > 0.0 || !isfinite(logdet[b]) || !isfinite(logdet2[b])) {
                fprintf(stderr, "matrix %d: sign %g / %g logdet %g / %g\n", b, sign[b], sign2[b], logdet[b], logdet2[b]);
                return 13;
            }
        }
        cudaFree(d_A); cudaFree(d_L); cudaFree(d_R); cudaFree(d_B); cudaFree(d_G); cudaFree(d_P); cudaFree(d_d);
        cudaFree(d_ws); cudaFree(d_logdet); cudaFree(d_sign); cudaFree(d_logdet2); cudaFree(d_sign2);
        free(A); free(B); free(G); free(P);
    }

    /* a handle refuses a call made with another current device only when there is one; argument errors are codes */
    if (sla_ldr(NULL, SLA_F64, n, batch, NULL, NULL, NULL, NULL) != -1) return 14;
    if (sla_ldr(h, 7, n, batch, NULL, NULL, NULL, NULL) != -2) return 15;
    SLA(sla_destroy(h));
    CU(cudaStreamDestroy(stream));
    printf("ABI_C_OK launches=%llu\n", sla_launch_count());
    return 0;
}
