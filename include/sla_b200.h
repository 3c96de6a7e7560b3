/* libsla_b200 -- C ABI of the B200-native stabilised matrix-product path of StableLinearAlgebra.jl.
 *
 * This is the drop-in boundary a Julia shim (or any FFI) binds: every entry point replaces one
 * exported method of the reference (file:line given per function, relative to the reference repo)
 * and operates on a BATCH of independent matrices/factorisations -- the batched generalisation of
 * the reference's serial 3-D API `ldrs!(Fs, A::AbstractArray{T,3}, ws)` (src/LDR.jl:249-257).
 *
 * Conventions
 *   - device pointers only; the caller (CuArray / torch tensor) owns every buffer including the
 *     workspace blob (sized by sla_workspace_bytes) -- the library never allocates in a hot call,
 *     mirroring the reference's "no allocations after construction" rule (src/LDR.jl:39).
 *   - matrices: column-major n x n, leading dimension n, batch stride n*n elements (= Julia
 *     CuArray{T,3} of size (n,n,batch));  d: n reals per matrix, batch stride n.
 *   - dtype: SLA_F64 (Float64) or SLA_C64 (ComplexF64, interleaved re,im); SLA_F32 / SLA_C32 as storage types.
 *     E = Float64 always.
 *   - logdet: double[batch];  sign: T[batch] (double, or (re,im) pair for SLA_C64) -- the
 *     Tuple{E,T} the reference's inverses return (src/exported_functions.jl:25).
 *   - every call is asynchronous on the handle's stream; a handle is bound to one stream and one
 *     workspace may be used by one in-flight call at a time (LDRWorkspace is not thread-safe either).
 *   - inputs are read-only unless documented in-place.  U and V may alias in the sla_inv_* calls
 *     (test/runtests.jl:280) but NOT in sla_lmul_ldr_ldr / sla_rmul_ldr_ldr
 *     (src/overloaded_functions.jl:179,185).  G must not alias the workspace.
 *   - return value: 0 ok; -k = argument k invalid; 1000+e = CUDA runtime error e; SLA_ENOTIMPL.
 *     LAPACK-style `info` (exact zero pivot) is written per matrix to the int array returned by
 *     sla_workspace_info() and is checked lazily by the host shim (chklapackerror, src/lu.jl:75).
 */
#ifndef SLA_B200_H
#define SLA_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLA_F64 0
#define SLA_C64 1
/* Float32 / ComplexF32 STORAGE types (src/qr.jl:29-32, src/lu.jl:26-29 generate all four element types): matrices and
 * signs are float / (float, float) in the caller's buffers; d and log|det| stay double (LDR{Float32,Float64}: a
 * chain's D leaves the Float32 range); every entry point promotes to FP64 inside the workspace, runs the FP64 kernels
 * and demotes.  Not supported by sla_gemm and sla_walker_sums (plain Float32 GEMMs are the caller's cuBLAS business). */
#define SLA_F32 2
#define SLA_C32 3
#define SLA_OP_N 0
#define SLA_OP_C 1 /* conjugate transpose (adjoint) */
#define SLA_INV_IPA 0
#define SLA_INV_IPUV 1
#define SLA_ENOTIMPL (-100) /* the requested (forced) kernel does not exist for this element type / size */
#define SLA_EDEVICE (-101)  /* the handle was created on another device than the calling thread's current one */
/* values of sla_last_ldr_algo(): which LDR factorisation kernel the dispatcher chose */
#define SLA_LDR_ALGO_NONE 0
#define SLA_LDR_ALGO_REG 1       /* register-resident cluster kernel (csrc/sla_ldr_reg.cu) */
#define SLA_LDR_ALGO_CLUSTER 2   /* shared-memory cluster kernel (csrc/sla_ldr_cluster.cu) */
#define SLA_LDR_ALGO_STREAMING 3 /* one CTA per matrix, L2-streaming (csrc/sla_ldr.cu) */
#define SLA_LDR_ALGO_HYBRID 4    /* registers + shared memory, clusters of 2 (csrc/sla_ldr_reg.cu) */
#define SLA_LDR_ALGO_BLOCKED 5   /* blocked panel + tensor-core trailing update (csrc/sla_ldr_blk.cu) */
#define SLA_LDR_ALGO_MULTICTA 6  /* L2-streaming, a cluster of 2-8 CTAs per matrix (csrc/sla_ldr_mc.cu) */
#define SLA_LDR_ALGO_PANEL 7     /* panel-restricted pivoting, opt-in mode SLA_LDR_PIVOT=panel (csrc/sla_ldr_pp.cu) */
#define SLA_LDR_ALGO_SMALL 8     /* n <= 64: one CTA per matrix, matrix in registers (csrc/sla_ldr_small.cu) */
/* per-matrix info codes (sla_workspace_info): k > 0 = U(k,k) exactly zero or not finite in an LU (src/lu.jl:75,91);
 * SLA_INFO_RANGE = a column norm of an ldr! input left the range in which squared norms are representable */
#define SLA_INFO_RANGE (-1)

#if defined(__GNUC__)
#define SLA_API __attribute__((visibility("default")))
#else
#define SLA_API
#endif

typedef struct sla_handle_s* sla_handle_t;

SLA_API int sla_version(void);
/* kernels launched by this library in this process so far (diagnostics; bench.py reports the per-step delta) */
SLA_API unsigned long long sla_launch_count(void);
/* diagnostics: device buffer of 128 int64 that the cluster LDR kernel fills with per-phase cycle counters
 * of cluster 0 (NULL switches it off, the default) */
SLA_API void sla_debug_set_ldr_prof(void* dev_buffer_128_int64);
/* stream: a cudaStream_t (0 = legacy default stream). */
SLA_API int sla_create(sla_handle_t* handle, void* stream);
SLA_API int sla_destroy(sla_handle_t handle);
SLA_API int sla_set_stream(sla_handle_t handle, void* stream);
/* diagnostics for the tests: SLA_LDR_ALGO_* of the last factorisation launched through this handle, and how many
 * factorisations each algorithm has served on it */
SLA_API int sla_last_ldr_algo(sla_handle_t handle);
SLA_API unsigned long long sla_ldr_algo_launches(sla_handle_t handle, int algo);

/* largest supported matrix size for an element type (the QR / LU panels of one matrix must fit one SM's shared
 * memory); larger n: sla_workspace_bytes() returns 0 and every entry point returns SLA_ENOTIMPL */
SLA_API int sla_max_n(int dtype);

/* LDRWorkspace equivalent (src/LDR.jl:51-70, ldr_workspace src/LDR.jl:267-282): bytes of the scratch
 * blob every call below takes as `ws` (M, M', M'', the LU/QR scratch, v and the per-matrix scalars). */
SLA_API size_t sla_workspace_bytes(int dtype, int n, int batch);
/* device pointer to the per-matrix LAPACK-style info codes (int[batch]) inside a workspace blob */
SLA_API int* sla_workspace_info(void* ws, int dtype, int n, int batch);

/* ldr!(F, ws)      src/LDR.jl:160-188 : factorise the matrices held in L in place -> (L=Q, d, R) */
SLA_API int sla_ldr(sla_handle_t h, int dtype, int n, int batch, void* L, double* d, void* R, void* ws);
/* ldr!(F, A, ws)   src/LDR.jl:122-126 / ldrs!(Fs, A3, ws) src/LDR.jl:249-257 ; strideA may be 0 */
SLA_API int sla_ldr_from(sla_handle_t h, int dtype, int n, int batch, const void* A, long long strideA, void* L,
                 double* d, void* R, void* ws);
/* ldr!(F, I)       src/LDR.jl:133-140 / copyto!(F, I) overloaded_functions.jl:56 */
SLA_API int sla_set_identity(sla_handle_t h, int dtype, int n, int batch, void* L, double* d, void* R);
/* ldr!(Fout, Fin)  src/LDR.jl:147-153 / copyto!(U::LDR, V::LDR) overloaded_functions.jl:63 */
SLA_API int sla_copy(sla_handle_t h, int dtype, int n, int batch, void* Lo, double* dout, void* Ro, const void* Li,
             const double* din, const void* Ri);
/* copyto!(A, F, ws)  overloaded_functions.jl:34-45 : A = L*D*R */
SLA_API int sla_ldr_to_mat(sla_handle_t h, int dtype, int n, int batch, void* A, const void* L, const double* d,
                   const void* R, void* ws);
/* adjoint!(At, F, ws) overloaded_functions.jl:75-86 : At = R^H D L^H */
SLA_API int sla_adjoint(sla_handle_t h, int dtype, int n, int batch, void* At, const void* L, const double* d,
                const void* R, void* ws);

/* lmul!(U::Matrix, V::LDR, ws)  overloaded_functions.jl:126-145 : V <- U*V, stabilised.  strideU 0 = shared U */
SLA_API int sla_lmul_mat_ldr(sla_handle_t h, int dtype, int n, int batch, const void* U, long long strideU, void* L,
                     double* d, void* R, void* ws);
/* rmul!(U::LDR, V::Matrix, ws)  overloaded_functions.jl:234-253 : U <- U*V, stabilised */
SLA_API int sla_rmul_ldr_mat(sla_handle_t h, int dtype, int n, int batch, void* L, double* d, void* R, const void* V,
                     long long strideV, void* ws);
/* lmul!(U::LDR, V::LDR, ws)     overloaded_functions.jl:166-193 : V <- U*V (U read-only, U != V) */
SLA_API int sla_lmul_ldr_ldr(sla_handle_t h, int dtype, int n, int batch, const void* UL, const double* Ud,
                     const void* UR, void* VL, double* Vd, void* VR, void* ws);
/* rmul!(U::LDR, V::LDR, ws)     overloaded_functions.jl:274-301 : U <- U*V (V read-only, U != V) */
SLA_API int sla_rmul_ldr_ldr(sla_handle_t h, int dtype, int n, int batch, void* UL, double* Ud, void* UR,
                     const void* VL, const double* Vd, const void* VR, void* ws);
/* lmul!(U::LDR, V::Matrix, ws)  overloaded_functions.jl:97-106 : V <- L D R V (dense, not stabilised) */
SLA_API int sla_lmul_ldr_mat(sla_handle_t h, int dtype, int n, int batch, const void* L, const double* d, const void* R,
                     void* V, void* ws);
/* rmul!(U::Matrix, V::LDR, ws)  overloaded_functions.jl:205-214 : U <- U L D R */
SLA_API int sla_rmul_mat_ldr(sla_handle_t h, int dtype, int n, int batch, void* U, const void* L, const double* d,
                     const void* R, void* ws);

/* det_lu!(A, ws)   src/lu.jl:123-140 : A <- LU factors; (log|det A|, sign det A) */
SLA_API int sla_det_lu(sla_handle_t h, int dtype, int n, int batch, void* A, double* logdet, void* sign, void* ws);
/* inv_lu!(A, ws)   src/lu.jl:149-158 : A <- A^-1; (log|det A^-1|, sign det A^-1) */
SLA_API int sla_inv_lu(sla_handle_t h, int dtype, int n, int batch, void* A, double* logdet, void* sign, void* ws);
/* logabsdet(F, ws) overloaded_functions.jl:721-733 */
SLA_API int sla_logabsdet(sla_handle_t h, int dtype, int n, int batch, const void* L, const double* d, const void* R,
                  double* logdet, void* sign, void* ws);

/* ldiv_lu!(A, B, ws)  src/lu.jl:167-176 : A <- LU factors, B <- A^-1 B; (log|det A^-1|, sign det A^-1) */
SLA_API int sla_ldiv_lu(sla_handle_t h, int dtype, int n, int batch, void* A, void* B, double* logdet, void* sign,
                        void* ws);
/* ldiv!(U::LDR, V::Matrix, ws)  overloaded_functions.jl:388-400 : V <- U^-1 V (dense) */
SLA_API int sla_ldiv_ldr_mat(sla_handle_t h, int dtype, int n, int batch, const void* L, const double* d, const void* R,
                             void* V, void* ws);
/* ldiv!(U::LDR, V::LDR, ws)     overloaded_functions.jl:437-464 : V <- U^-1 V, stabilised (U != V) */
SLA_API int sla_ldiv_ldr_ldr(sla_handle_t h, int dtype, int n, int batch, const void* UL, const double* Ud,
                             const void* UR, void* VL, double* Vd, void* VR, void* ws);
/* ldiv!(U::Matrix, V::LDR, ws)  overloaded_functions.jl:502-521 : V <- U^-1 V, stabilised; strideU 0 = shared U */
SLA_API int sla_ldiv_mat_ldr(sla_handle_t h, int dtype, int n, int batch, const void* U, long long strideU, void* L,
                             double* d, void* R, void* ws);
/* rdiv!(U::Matrix, V::LDR, ws)  overloaded_functions.jl:549-565 : U <- U V^-1 (dense) */
SLA_API int sla_rdiv_mat_ldr(sla_handle_t h, int dtype, int n, int batch, void* U, const void* L, const double* d,
                             const void* R, void* ws);
/* rdiv!(U::LDR, V::LDR, ws)     overloaded_functions.jl:603-637 : U <- U V^-1, stabilised (U != V) */
SLA_API int sla_rdiv_ldr_ldr(sla_handle_t h, int dtype, int n, int batch, void* UL, double* Ud, void* UR,
                             const void* VL, const double* Vd, const void* VR, void* ws);
/* rdiv!(U::LDR, V::Matrix, ws)  overloaded_functions.jl:673-693 : U <- U V^-1, stabilised; strideV 0 = shared V */
SLA_API int sla_rdiv_ldr_mat(sla_handle_t h, int dtype, int n, int batch, void* L, double* d, void* R, const void* V,
                             long long strideV, void* ws);

/* inv_IpA!(G, A, ws)       exported_functions.jl:25-71   : G = (I + A)^-1 */
SLA_API int sla_inv_IpA(sla_handle_t h, int dtype, int n, int batch, void* G, const void* L, const double* d,
                const void* R, double* logdet, void* sign, void* ws);
/* inv_IpUV!(G, U, V, ws)   exported_functions.jl:97-177  : G = (I + U V)^-1 */
SLA_API int sla_inv_IpUV(sla_handle_t h, int dtype, int n, int batch, void* G, const void* UL, const double* Ud,
                 const void* UR, const void* VL, const double* Vd, const void* VR, double* logdet, void* sign,
                 void* ws);
/* inv_UpV!(G, U, V, ws)    exported_functions.jl:211-289 : G = (U + V)^-1 */
SLA_API int sla_inv_UpV(sla_handle_t h, int dtype, int n, int batch, void* G, const void* UL, const double* Ud,
                const void* UR, const void* VL, const double* Vd, const void* VR, double* logdet, void* sign,
                void* ws);
/* inv_invUpV!(G, U, V, ws) exported_functions.jl:324-402 : G = (U^-1 + V)^-1 */
SLA_API int sla_inv_invUpV(sla_handle_t h, int dtype, int n, int batch, void* G, const void* UL, const double* Ud,
                   const void* UR, const void* VL, const double* Vd, const void* VR, double* logdet, void* sign,
                   void* ws);

/* mul!(C, A, B) on plain matrices (BLAS ?gemm through LinearAlgebra.mul!, e.g. test/runtests.jl:96-99):
 * C[b] = op(A[b]) * op(B[b]); stride 0 shares an operand across the batch; accumulate != 0: C += . */
SLA_API int sla_gemm(sla_handle_t h, int dtype, int opa, int opb, int m, int n, int k, const void* A, int lda,
             long long strideA, const void* B, int ldb, long long strideB, void* C, int ldc, long long strideC,
             int batch, int accumulate);

/* Fused driver for one stabilised G(tau=0) evaluation per chain (the loop of test/runtests.jl:93-100,
 * 151-160 on Hubbard-like B matrices): B_{c,l} = diag(exp(lam*s[c,l,:])) * expK; the group products
 * B_bar_g = B_{g*n_stab} ... B_{(g-1)n_stab+1} are built by GEMMs, F <- I, lmul!(B_bar_g, F) for all
 * groups, then inv_IpA!(G, F) (SLA_INV_IPA) or -- first half of the groups into V, second half into U --
 * inv_IpUV!(G, U, V) (SLA_INV_IPUV).
 *   expK   : n x n (shared by all chains);  fields: int8 [batch][L][n] in {-1,+1}
 *   G      : out n x n x batch;  logdet: double[batch];  sign: T[batch]
 *   Lout/dout/Rout: optional (may be NULL) final LDR factorisation F (SLA_INV_IPA only)
 *   ws     : blob of at least sla_greens_chain_workspace_bytes(...) bytes */
SLA_API size_t sla_greens_chain_workspace_bytes(int dtype, int n, int batch, int L, int n_stab);
/* local part of the cross-walker reduction: out4 = [sum logdet, sum Re sign, sum Im sign, batch] on
 * the device, in a fixed summation order; the caller hands out4 straight to ncclAllReduce(sum) */
SLA_API int sla_walker_sums(sla_handle_t h, int dtype, int batch, const double* logdet, const void* sign, double* out4);
/* device pointer to the per-chain info codes (int[batch]) inside a sla_greens_chain workspace blob */
SLA_API int* sla_greens_chain_workspace_info(void* ws, int dtype, int n, int batch, int L, int n_stab);
SLA_API int sla_greens_chain(sla_handle_t h, int dtype, int n, int batch, int L, int n_stab, const void* expK,
                     const signed char* fields, double lam, int inverse, void* G, double* logdet, void* sign,
                     void* Lout, double* dout, void* Rout, void* ws, size_t ws_bytes);

/* Fused DQMC sweep stepper: the caller pattern of inv_IpUV! (exported_functions.jl:97-177) over a
 * stack of LDR factorisations (the `ldrs` containers of src/LDR.jl:197-257), one upward sweep per chain, fields fixed:
 *   setup  F_left[ng] = I;  F_left[g] = rmul!(F_left[g+1], B_bar_{g+1})  (overloaded_functions.jl:234-253), all kept;
 *          G(0) = inv_IpA!(F_left[0])                                   -> G0 (optional), logdet[0][:], sign[0][:]
 *   sweep  G <- B_l G B_l^-1 for every slice l (B_l^-1 = expK^-1 diag(exp(-lam s_l)));  every n_stab slices
 *          F_right <- lmul!(B_bar_g, F_right) (overloaded_functions.jl:126-145),
 *          G' = inv_IpUV!(F_right, F_left[g])                           -> logdet[g][:], sign[g][:],
 *          dG[g-1][:] = max |G' - G| (the wrap error),  G <- G'.
 *   G: out, G(beta) (= G(0) up to rounding: the fields are not updated);  logdet: double[ng+1][batch];
 *   sign: T[ng+1][batch];  dG: double[ng][batch] or NULL;  expKinv: n x n, the inverse of expK. */
SLA_API size_t sla_sweep_workspace_bytes(int dtype, int n, int batch, int L, int n_stab);
SLA_API int* sla_sweep_workspace_info(void* ws, int dtype, int n, int batch, int L, int n_stab);
SLA_API int sla_sweep(sla_handle_t h, int dtype, int n, int batch, int L, int n_stab, const void* expK, const void* expKinv,
                      const signed char* fields, double lam, void* G0, void* G, double* logdet, void* sign, double* dG,
                      void* ws, size_t ws_bytes);

#ifdef __cplusplus
}
#endif
#endif /* SLA_B200_H */
