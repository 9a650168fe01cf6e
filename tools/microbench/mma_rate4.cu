// mma_rate4.cu - does the operand GEOMETRY of the full-halo tiles (unaligned start rows, 10-row group strides,
// MN-major views) change the tcgen05.mma rate?  Same harness as mma_rate3 (fully unrolled groups of P MMAs).
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace eld;

struct Geo { const char* name; int a_mn, b_mn; uint32_t a_lbo, a_sbo, a_layout, a_off, a_kstep; uint32_t b_lbo, b_sbo, b_layout, b_kstep; };

template <int P>
__global__ void __launch_bounds__(128, 1) k(int N, int iters, Geo g, long long* out)
{
    extern __shared__ uint8_t raw[];
    const uint32_t r = ptx::smem_u32(raw);
    uint8_t* smem = raw + (((r + 1023u) & ~1023u) - r);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (160 * 1024) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) ptx::tmem_alloc(&slot, 512);
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x < 32) {
        const uint32_t idesc = ptx::make_idesc_bf16(128, N, g.a_mn, g.b_mn);
        const uint64_t ahi = ptx::make_smem_desc(0, g.a_lbo, g.a_sbo, g.a_layout);
        const uint64_t bhi = ptx::make_smem_desc(0, g.b_lbo, g.b_sbo, g.b_layout);
        const uint32_t a_lo = (uint32_t)ahi | (((ptx::smem_u32(smem) + g.a_off) & 0x3FFFFu) >> 4);
        const uint32_t b_lo = (uint32_t)bhi | (((ptx::smem_u32(smem) + 98304) & 0x3FFFFu) >> 4);
        const uint32_t a_h = (uint32_t)(ahi >> 32), b_h = (uint32_t)(bhi >> 32);
        const uint32_t ak = g.a_kstep >> 4, bk = g.b_kstep >> 4;
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (ptx::elect_one()) {
#pragma unroll
                for (int j = 0; j < P; ++j)
                    ptx::umma_bf16_lohi(tm + (j & 1) * 256, a_lo + ak * (j & 3), a_h, b_lo + bk * (j & 3), b_h, idesc, true);
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

int main()
{
    long long* d; cudaMalloc(&d, 148 * 8);
    cudaFuncSetAttribute(k<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const uint32_t SW128 = ptx::LAYOUT_SW128, SW64 = ptx::LAYOUT_SW64;
    Geo geos[] = {
        // K-major reference (aligned atoms)
        { "K-major SW128 aligned (A sbo 1024)", 0, 0, 16, 1024, SW128, 0, 32, 16, 1024, SW128, 32 },
        { "K-major SW64 aligned (A sbo 512)", 0, 0, 16, 512, SW64, 0, 32, 16, 512, SW64, 32 },
        // conv full-halo views: 8-row groups 10 rows apart, start shifted by 11 rows
        { "K-major SW128 halo view (sbo 1280, +11 rows)", 0, 0, 16, 1280, SW128, 11 * 128, 32, 16, 1024, SW128, 32 },
        { "K-major SW64 halo view (sbo 640, +11 rows)", 0, 0, 16, 640, SW64, 11 * 64, 32, 16, 512, SW64, 32 },
        // wgrad: MN-major, A = halo box rows (pixels) x channels
        { "MN-major SW128 aligned (lbo 8192 sbo 1024)", 1, 1, 8192, 1024, SW128, 0, 2048, 8192, 1024, SW128, 2048 },
        { "MN-major SW128 wgrad kind2 (A lbo 13312 sbo 1280 +11 rows, kstep 20 rows)", 1, 1, 13312, 1280, SW128, 11 * 128, 20 * 128, 8192, 1024, SW128, 16 * 128 },
        { "MN-major SW128 wgrad kind1 (A lbo 1 row sbo 1280)", 1, 1, 128, 1280, SW128, 0, 20 * 128, 8192, 1024, SW128, 16 * 128 },
        { "MN-major SW64 wgrad kind0 (A lbo 64 sbo 640, B SW64 sbo 512)", 1, 1, 64, 640, SW64, 10 * 64, 20 * 64, 4096, 512, SW64, 16 * 64 },
    };
    for (const Geo& g : geos) for (int N : {32, 64, 128, 256}) {
        const int iters = 1000;
        k<12><<<148, 128, 180 * 1024>>>(N, iters, g, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
        printf("%-78s N %3d : %6.1f cycles per MMA  [%s]\n", g.name, N, (double)h[0] / iters / 12, cudaGetErrorString(e));
    }
    return 0;
}
