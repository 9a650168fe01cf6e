// mma_rate2.cu - issue-thread cost of the things that surround tcgen05.mma in a pipelined tile:
// tcgen05.commit, mbarrier try_wait on a completed phase, tcgen05.fence, and MN-major operands.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../eld_b200/csrc mma_rate2.cu -o mma_rate2
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace eld;

// mode bits: 1 = commit to a scratch barrier every iteration, 2 = try_wait on a completed phase every iteration,
//            4 = tcgen05.fence::after_thread_sync every iteration, 8 = MN-major A and B
__global__ void __launch_bounds__(128, 1) k(int N, int per_iter, int iters, int mode, long long* out)
{
    extern __shared__ uint8_t raw[];
    const uint32_t r = ptx::smem_u32(raw);
    uint8_t* smem = raw + (((r + 1023u) & ~1023u) - r);
    __shared__ uint64_t bar, scratch[4], done_bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (96 * 1024) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
    if (threadIdx.x == 0) {
        ptx::mbar_init(&bar, 1); ptx::mbar_init(&done_bar, 1);
        for (int i = 0; i < 4; ++i) ptx::mbar_init(&scratch[i], 1);
        ptx::fence_barrier_init();
    }
    if (threadIdx.x < 32) ptx::tmem_alloc(&slot, 512);
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x < 32 && ptx::elect_one()) {
        const bool mn = mode & 8;
        const uint32_t idesc = ptx::make_idesc_bf16(128, N, mn, mn);
        const uint64_t hi = mn ? ptx::make_smem_desc(0, 8192, 1024, ptx::LAYOUT_SW128) : ptx::make_smem_desc(0, 16, 1024, ptx::LAYOUT_SW128);
        const uint32_t a_lo = (uint32_t)hi | ((ptx::smem_u32(smem) & 0x3FFFFu) >> 4);
        const uint32_t b_lo = (uint32_t)hi | (((ptx::smem_u32(smem) + 32768) & 0x3FFFFu) >> 4);
        const uint32_t h32 = (uint32_t)(hi >> 32);
        const uint32_t kstep = mn ? 128u : 2u;    // MN-major: 16 k rows x 128 B = 2 KB
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (mode & 2) ptx::mbar_wait(&done_bar, 1);
            if (mode & 4) ptx::tc_fence_after();
            for (int j = 0; j < per_iter; ++j)
                ptx::umma_bf16_lohi(tm + (j & 1) * N, a_lo + kstep * (j & 3), h32, b_lo + kstep * (j & 3), h32, idesc, true);
            if (mode & 1) ptx::umma_commit(&scratch[i & 3]);
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
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
    const int iters = 1000;
    for (int N : {32, 64, 128, 256}) for (int per : {2, 4, 8}) for (int mode : {0, 1, 2, 4, 3, 7, 8, 9}) {
        k<<<148, 128, 100 * 1024>>>(N, per, iters, mode, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
        printf("N %3d  %d MMA/iter  mode %d%s%s%s%s : %.1f cycles per iteration, %.1f per MMA  [%s]\n", N, per, mode,
               (mode & 1) ? " commit" : "", (mode & 2) ? " wait" : "", (mode & 4) ? " fence" : "", (mode & 8) ? " MN-major" : "",
               (double)h[0] / iters, (double)h[0] / iters / per, cudaGetErrorString(e));
    }
    return 0;
}
