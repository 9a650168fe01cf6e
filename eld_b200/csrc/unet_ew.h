// unet_ew.h - launchers of the non-GEMM U-Net kernels (unet_ew.cu).
#pragma once
#include "common.cuh"

namespace eld {
int launch_maxpool(eld_ctx* ctx, const void* in, int in_pitch, int in_c0, void* out, int C, int n, int Ho, int Wo, cudaStream_t st);
int launch_maxpool_bwd_code(eld_ctx* ctx, const void* code, const void* dskip, int s_pitch, int s_c0,
                            const void* dP, void* dZ, int C, int n, int Ho, int Wo, cudaStream_t st);
int launch_maxpool_bwd(eld_ctx* ctx, const void* A, int a_pitch, int a_c0, const void* dskip, int s_pitch, int s_c0,
                       const void* dP, void* dZ, int C, int n, int Ho, int Wo, cudaStream_t st);
int launch_colsum(eld_ctx* ctx, const void* g, int pitch, int c0, int C, size_t npix, float* out, cudaStream_t st);
int launch_head(eld_ctx* ctx, const void* a, const float* w, const float* b, float* out, const float* target, void* dz,
                float* dw, float* db, float* loss, int n, size_t plane, int cout, int l2_loss, cudaStream_t st);
int launch_clock_probe(eld_ctx* ctx, float* out_mhz, cudaStream_t st);
int launch_adam(eld_ctx* ctx, float* p, const float* g, float* m, float* v, size_t n, float lr, float b1, float b2,
                float eps, float wd, int step, float gscale, cudaStream_t st);
}  // namespace eld
