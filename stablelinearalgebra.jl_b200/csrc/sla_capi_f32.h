// Internal: SLA_F32 / SLA_C32 entry points (promote -> FP64 entry point -> demote), see sla_capi_f32.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "../../include/sla_b200.h"

cudaStream_t sla_handle_stream(sla_handle_t h);   // sla_capi.cu

namespace sla {
size_t f32_workspace_bytes(int dtype, int n, int batch);
size_t f32_chain_workspace_bytes(int dtype, int n, int batch, int L, int n_stab);
size_t f32_sweep_workspace_bytes(int dtype, int n, int batch, int L, int n_stab);
}  // namespace sla

extern "C" {
int f32_ldr(sla_handle_t h, int dtype, int n, int batch, void* L, double* d, void* R, void* ws);
int f32_ldr_from(sla_handle_t h, int dtype, int n, int batch, const void* A, long long strideA, void* L, double* d, void* R, void* ws);
int f32_ldr_to_mat(sla_handle_t h, int dtype, int n, int batch, void* A, const void* L, const double* d, const void* R, void* ws, int adjoint);
int f32_mat_ldr(sla_handle_t h, int dtype, int n, int batch, const void* U, long long strideU, void* L, double* d, void* R, void* ws, int which);
int f32_ldr_mat(sla_handle_t h, int dtype, int n, int batch, void* L, double* d, void* R, const void* V, long long strideV, void* ws, int which);
int f32_ldr_ldr(sla_handle_t h, int dtype, int n, int batch, void* UL, double* Ud, void* UR, void* VL, double* Vd, void* VR, void* ws, int which);
int f32_dense(sla_handle_t h, int dtype, int n, int batch, void* M, const void* L, const double* d, const void* R, void* ws, int which);
int f32_lu(sla_handle_t h, int dtype, int n, int batch, void* A, void* B, double* logdet, void* sign, void* ws, int which);
int f32_logabsdet(sla_handle_t h, int dtype, int n, int batch, const void* L, const double* d, const void* R, double* logdet, void* sign, void* ws);
int f32_inv_IpA(sla_handle_t h, int dtype, int n, int batch, void* G, const void* L, const double* d, const void* R, double* logdet, void* sign, void* ws);
int f32_inv_uv(sla_handle_t h, int dtype, int n, int batch, void* G, const void* UL, const double* Ud, const void* UR, const void* VL,
               const double* Vd, const void* VR, double* logdet, void* sign, void* ws, int which);
int f32_set_identity(sla_handle_t h, int dtype, int n, int batch, void* L, double* d, void* R);
int f32_copy(sla_handle_t h, int dtype, int n, int batch, void* Lo, double* dout, void* Ro, const void* Li, const double* din, const void* Ri);
int f32_greens_chain(sla_handle_t h, int dtype, int n, int batch, int L, int n_stab, const void* expK, const signed char* fields,
                     double lam, int inverse, void* G, double* logdet, void* sign, void* Lout, double* dout, void* Rout, void* ws, size_t ws_bytes);
int f32_sweep(sla_handle_t h, int dtype, int n, int batch, int L, int n_stab, const void* expK, const void* expKinv,
              const signed char* fields, double lam, void* G0, void* G, double* logdet, void* sign, double* dG, void* ws, size_t ws_bytes);
}
