// mma_rate5.cu - tcgen05.mma rate when every MMA reads a DIFFERENT A tile (operand footprint 4 KB x P), versus the
// 4-descriptor loop of mma_rate3 whose operands may stay in the tensor core's operand buffers.
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace eld;

template <int P, int ASTRIDE /*bytes between consecutive A tiles*/, int BSTRIDE>
__global__ void __launch_bounds__(128, 1) k(int N, int iters, long long* out)
{
    extern __shared__ uint8_t raw[];
    const uint32_t r = ptx::smem_u32(raw);
    uint8_t* smem = raw + (((r + 1023u) & ~1023u) - r);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (200 * 1024) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
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
        const uint32_t b_lo = (uint32_t)hi | (((ptx::smem_u32(smem) + 160 * 1024) & 0x3FFFFu) >> 4);
        const uint32_t h32 = (uint32_t)(hi >> 32);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (ptx::elect_one()) {
#pragma unroll
                for (int j = 0; j < P; ++j)
                    ptx::umma_bf16_lohi(tm + (j % 3) * 128, a_lo + (uint32_t)((j * ASTRIDE) >> 4), h32, b_lo + (uint32_t)(((j & 3) * BSTRIDE) >> 4), h32, idesc, true);
            }
            __syncwarp();
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

template <int P, int AS, int BS>
void run(const char* what, long long* d)
{
    cudaFuncSetAttribute(k<P, AS, BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    for (int N : {32, 64, 128}) {
        k<P, AS, BS><<<148, 128, 210 * 1024>>>(N, 1000, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
        printf("%-58s P %2d N %3d : %6.1f cycles per MMA  [%s]\n", what, P, N, (double)h[0] / 1000 / P, cudaGetErrorString(e));
    }
}

int main()
{
    long long* d; cudaMalloc(&d, 148 * 8);
    run<36, 32, 32>("A: 4 k-slices of ONE 128x64 tile (32 B apart)", d);     // like mma_rate3 (re-reads the same 16 KB)
    run<36, 4096, 32>("A: 36 distinct tiles, 4 KB apart (144 KB footprint)", d);
    run<36, 4096, 8192>("A: 36 distinct tiles; B: 4 distinct tiles 8 KB apart", d);
    run<18, 8192, 32>("A: 18 distinct tiles, 8 KB apart", d);
    run<36, 1280, 32>("A: overlapping views 1280 B apart (halo taps)", d);
    return 0;
}
