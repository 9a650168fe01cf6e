// mma_rate3.cu - fully unrolled issue groups (compile-time MMA count, immediate descriptor offsets):
// what does one pipelined "wait -> P x tcgen05.mma -> commit" iteration cost the issuing thread?
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace eld;

template <int P, int MODE>
__global__ void __launch_bounds__(128, 1) k(int N, int iters, long long* out)
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
    if (threadIdx.x < 32) {
        const uint32_t idesc = ptx::make_idesc_bf16(128, N, 0, 0);
        const uint64_t hi = ptx::make_smem_desc(0, 16, 1024, ptx::LAYOUT_SW128);
        const uint32_t a_lo = (uint32_t)hi | ((ptx::smem_u32(smem) & 0x3FFFFu) >> 4);
        const uint32_t b_lo = (uint32_t)hi | (((ptx::smem_u32(smem) + 32768) & 0x3FFFFu) >> 4);
        const uint32_t h32 = (uint32_t)(hi >> 32);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (MODE & 2) ptx::mbar_wait(&done_bar, 1);
            if (MODE & 4) ptx::tc_fence_after();
            if (ptx::elect_one()) {
#pragma unroll
                for (int j = 0; j < P; ++j)
                    ptx::umma_bf16_lohi(tm + (j & 1) * 256, a_lo + 2 * (j & 3) + 64 * (j >> 2), h32, b_lo + 2 * (j & 3), h32, idesc, true);
                if (MODE & 1) ptx::umma_commit(&scratch[i & 3]);
            }
            if (MODE & 8) __syncwarp();
        }
        if (ptx::elect_one()) ptx::umma_commit(&bar);
        ptx::mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) ptx::tmem_dealloc(tm, 512);
}

template <int P, int MODE>
void run(int N, long long* d)
{
    const int iters = 1000;
    cudaFuncSetAttribute(k<P, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
    k<P, MODE><<<148, 128, 100 * 1024>>>(N, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
    printf("N %3d  P %2d  mode %2d%s%s%s%s : %7.1f cycles per iteration, %6.1f per MMA  [%s]\n", N, P, MODE,
           (MODE & 1) ? " commit" : "", (MODE & 2) ? " wait" : "", (MODE & 4) ? " fence" : "", (MODE & 8) ? " syncwarp" : "",
           (double)h[0] / iters, (double)h[0] / iters / P, cudaGetErrorString(e));
}

template <int P> void runs(long long* d)
{
    for (int N : {32, 64, 128}) { run<P, 0>(N, d); run<P, 1>(N, d); run<P, 2>(N, d); run<P, 7>(N, d); run<P, 15>(N, d); }
}

int main()
{
    long long* d; cudaMalloc(&d, 148 * 8);
    runs<1>(d); runs<2>(d); runs<4>(d); runs<8>(d); runs<12>(d); runs<18>(d); runs<36>(d);
    return 0;
}
