// mma_rate.cu - how many cycles does one tcgen05.mma.cta_group::1.kind::f16 (M=128, N, K=16) take when
// issued back to back from resident shared memory?  (decides whether 1-CTA tiles can exceed ~50% of peak)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../eld_b200/csrc mma_rate.cu -o mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace eld;

__global__ void __launch_bounds__(128, 1) k(int N, int iters, int same_acc, long long* out)
{
    extern __shared__ uint8_t raw[];
    const uint32_t r = ptx::smem_u32(raw);
    uint8_t* smem = raw + (((r + 1023u) & ~1023u) - r);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (64 * 1024) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) ptx::tmem_alloc(&slot, 512);
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = ptx::make_idesc_bf16(128, N, 0, 0);
        const uint64_t hi = ptx::make_smem_desc(0, 16, 1024, ptx::LAYOUT_SW128);
        const uint32_t a_lo = (uint32_t)hi | ((ptx::smem_u32(smem) & 0x3FFFFu) >> 4);
        const uint32_t b_lo = (uint32_t)hi | (((ptx::smem_u32(smem) + 16384) & 0x3FFFFu) >> 4);
        const uint32_t h32 = (uint32_t)(hi >> 32);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                ptx::umma_bf16_lohi(tm + (same_acc ? 0 : (j & 1) * N), a_lo + 2 * (j & 3), h32, b_lo + 2 * (j & 3), h32, idesc, true);
        }
        ptx::umma_commit(&bar);
        ptx::mbar_wait(&bar, 0);
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) ptx::tmem_dealloc(tm, 512);
}

int main()
{
    long long* d; cudaMalloc(&d, 148 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int grid : {1, 148}) for (int same : {1, 0}) for (int N : {32, 64, 128, 256}) {
        const int iters = 2000;
        k<<<grid, 128, 96 * 1024>>>(N, iters, same, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
        printf("grid %3d same_acc %d N %3d : %.1f cycles per MMA (M=128,K=16)  [%s]\n", grid, same, N, (double)h[0] / (iters * 8.0), cudaGetErrorString(e));
    }
    return 0;
}
