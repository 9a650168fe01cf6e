// first_conv.cuh - conv1_1 (4 -> 32 channels, 3x3, Unet.py:11) fprop and wgrad as tcgen05 tiles fed by a SOFTWARE im2col.
//
// The generic conv tile needs >= 32 input channels (one 64-byte TMA row per pixel), so round 1 ran conv1_1 on a
// zero-padded 32-channel copy of the input: a 134 MB pack pass, 18 MMAs per 128-pixel tile of which 7/8 multiplied
// zeros, and the same again for the weight gradient (216 us of a 4.1 ms step for 0.6 % of its FLOPs).  Here four
// builder warps read the fp32 NCHW frame directly (16 B per pixel) and write the im2col tile
//     A[128 pixels][k = tap*4 + c  (36 real, k = 36 is a column of ones, the rest zero)]       bf16, 128-byte rows, SW128
// into shared memory; that ONE tile is
//     fprop : the K-major A operand   D[128 px][32 co]  = A[px][k] * W[co][k]            3 MMAs (K = 48) per tile
//     wgrad : the MN-major A operand  D[k][32 co]      += A[px][k] * dZ[px][co]          8 MMAs (K = 128 pixels) per tile
// (the ones column makes row 36 of the wgrad accumulator the bias gradient).  No padded copy, no pack pass.
#pragma once
#include "umma.cuh"
#include <cuda_bf16.h>

namespace eld {

struct FirstConvParams {
    const float* x;             // f32 NCHW [n][4][H][W]
    int n_img, H, W;            // H % 8 == 0, W % 16 == 0
    int tiles_x, tiles_y;
    const uint8_t* w_img;       // fprop: 4 KB smem image of W[co][k] (K-major, SW128), k >= 36 zero
    const float* bias;          // fprop
    __nv_bfloat16* out;         // fprop: NHWC bf16, 32 channels at out_pitch
    int out_pitch;
    float* dw;                  // wgrad: f32 OIHW [32][4][3][3], accumulated into
    float* db;                  // wgrad: f32 [32]
    int stages;
};

constexpr int kFcThreads = 288;      // warps 0-3: im2col builders (one pixel each) | warp 4: MMA issuer | warps 5-8: epilogue
constexpr int kFcATile = 128 * 128;  // bytes
constexpr int kFcAcc = 4;            // fprop TMEM accumulator ring (4 x 32 columns)

__device__ __forceinline__ uint32_t fc_pack(float a, float b)
{
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// one row of the im2col tile: pixel (y, x) of image `img`, written as six swizzled 16-byte chunks
__device__ __forceinline__ void fc_build_row(const FirstConvParams& p, uint8_t* tile, int m, int img, int y, int x)
{
    const size_t plane = (size_t)p.H * p.W;
    const float* base = p.x + (size_t)img * 4 * plane;
    float v[9][4];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        const int yy = y + dy - 1;
        const bool oky = yy >= 0 && yy < p.H;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const int xx = x + dx - 1;
            const bool ok = oky && xx >= 0 && xx < p.W;
            const float* q = base + (size_t)(ok ? yy : 0) * p.W + (ok ? xx : 0);
#pragma unroll
            for (int c = 0; c < 4; ++c) v[dy * 3 + dx][c] = ok ? __ldg(q + c * plane) : 0.0f;
        }
    }
    uint8_t* row = tile + m * 128;
    const int sw = m & 7;
#pragma unroll
    for (int j = 0; j < 4; ++j) {                 // chunk j = taps 2j, 2j+1
        const uint4 c = make_uint4(fc_pack(v[2 * j][0], v[2 * j][1]), fc_pack(v[2 * j][2], v[2 * j][3]),
                                   fc_pack(v[2 * j + 1][0], v[2 * j + 1][1]), fc_pack(v[2 * j + 1][2], v[2 * j + 1][3]));
        *reinterpret_cast<uint4*>(row + ((j ^ sw) << 4)) = c;
    }
    // chunk 4 = tap 8, then k = 36: 1.0 (bias-gradient column of the wgrad; W[., 36] = 0 in the fprop operand), zeros
    *reinterpret_cast<uint4*>(row + ((4 ^ sw) << 4)) = make_uint4(fc_pack(v[8][0], v[8][1]), fc_pack(v[8][2], v[8][3]), 0x00003F80u, 0u);
    *reinterpret_cast<uint4*>(row + ((5 ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
}

__device__ __forceinline__ void fc_tile_coords(const FirstConvParams& p, int tile, int& img, int& y0, int& x0)
{
    const int txy = p.tiles_x * p.tiles_y;
    img = tile / txy;
    const int rem = tile - img * txy;
    const int ty = rem / p.tiles_x;
    y0 = ty * 8; x0 = (rem - ty * p.tiles_x) * 16;
}

// ---------------------------------------------------------------------------------------------------------------------
// fprop: a1_1 = lrelu(conv1_1(x) + b)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFcThreads, 1)
first_conv_fprop_kernel(const FirstConvParams p)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* w_s = smem;                                   // 4 KB weights
    uint8_t* a_s = smem + 4096;                            // ring of A tiles
    uint64_t* full = reinterpret_cast<uint64_t*>(a_s + (size_t)p.stages * kFcATile);
    uint64_t* empty = full + 8;
    uint64_t* tmem_full = empty + 8;
    uint64_t* tmem_empty = tmem_full + kFcAcc;
    uint64_t* w_full = tmem_empty + kFcAcc;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
    float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.n_img * p.tiles_x * p.tiles_y;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { ptx::mbar_init(&full[s], 128); ptx::mbar_init(&empty[s], 1); }
        for (int a = 0; a < kFcAcc; ++a) { ptx::mbar_init(&tmem_full[a], 1); ptx::mbar_init(&tmem_empty[a], 4); }
        ptx::mbar_init(w_full, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 5) ptx::tmem_alloc(tmem_slot, 128);
    if (threadIdx.x < 32) s_bias[threadIdx.x] = __ldg(p.bias + threadIdx.x);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::grid_dep_wait();          // PDL: x (noise kernel) and the packed weights (pack kernel) are complete past this point
    ptx::grid_dep_launch();

    if (warp < 4) {
        // ===================== im2col builders =====================
        const int m = threadIdx.x, py = m >> 4, px = m & 15;
        int s = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int img, y0, x0;
            fc_tile_coords(p, tile, img, y0, x0);
            ptx::mbar_wait(&empty[s], ph ^ 1u);
            fc_build_row(p, a_s + (size_t)s * kFcATile, m, img, y0 + py, x0 + px);
            ptx::fence_proxy_async();                      // generic-proxy stores -> visible to the tensor core's async proxy
            ptx::mbar_arrive(&full[s]);
            if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
    } else if (warp == 4) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            ptx::mbar_arrive_expect_tx(w_full, 4096u);
            ptx::bulk_load(w_s, p.w_img, 4096u, w_full);
        }
        __syncwarp();
        ptx::mbar_wait(w_full, 0);
        const uint32_t idesc = ptx::make_idesc_bf16(128, 32, 0, 0);
        const uint64_t desc_hi = ptx::make_smem_desc(0, 16, 1024, ptx::LAYOUT_SW128);
        const uint32_t hi = (uint32_t)(desc_hi >> 32);
        const uint32_t b_lo = (uint32_t)desc_hi | ((ptx::smem_u32(w_s) & 0x3FFFFu) >> 4);
        const uint32_t a_base = ptx::smem_u32(a_s);
        int s = 0;
        uint32_t ph = 0, acc = 0, acc_ph = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            ptx::mbar_wait(&tmem_empty[acc], acc_ph ^ 1u);
            ptx::mbar_wait(&full[s], ph);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                const uint32_t a_lo = (uint32_t)desc_hi | (((a_base + (uint32_t)s * kFcATile) & 0x3FFFFu) >> 4);
                const uint32_t d = tmem_base + acc * 32u;
                ptx::umma_bf16_lohi(d, a_lo, hi, b_lo, hi, idesc, false);
                ptx::umma_bf16_lohi(d, a_lo + 2u, hi, b_lo + 2u, hi, idesc, true);
                ptx::umma_bf16_lohi(d, a_lo + 4u, hi, b_lo + 4u, hi, idesc, true);
                ptx::umma_commit(&empty[s]);
                ptx::umma_commit(&tmem_full[acc]);
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1u; }
            if (++acc == kFcAcc) { acc = 0; acc_ph ^= 1u; }
        }
    } else {
        // ===================== epilogue: bias + LeakyReLU -> bf16 NHWC =====================
        const int q = warp & 3;                            // TMEM lane quarter of this warp
        const int m = q * 32 + lane, py = m >> 4, px = m & 15;
        const float4* sb4 = reinterpret_cast<const float4*>(s_bias);
        uint32_t acc = 0, acc_ph = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int img, y0, x0;
            fc_tile_coords(p, tile, img, y0, x0);
            ptx::mbar_wait(&tmem_full[acc], acc_ph);
            ptx::tc_fence_after();
            uint32_t r[32];
            ptx::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * 32u, r);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
            uint4* d4 = reinterpret_cast<uint4*>(p.out + ((size_t)(img * p.H + y0 + py) * p.W + (x0 + px)) * p.out_pitch);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const float4 b0 = sb4[2 * g], b1 = sb4[2 * g + 1];
                float v[8] = { __uint_as_float(r[8 * g]) + b0.x, __uint_as_float(r[8 * g + 1]) + b0.y,
                               __uint_as_float(r[8 * g + 2]) + b0.z, __uint_as_float(r[8 * g + 3]) + b0.w,
                               __uint_as_float(r[8 * g + 4]) + b1.x, __uint_as_float(r[8 * g + 5]) + b1.y,
                               __uint_as_float(r[8 * g + 6]) + b1.z, __uint_as_float(r[8 * g + 7]) + b1.w };
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.2f * v[j]);
                d4[g] = make_uint4(fc_pack(v[0], v[1]), fc_pack(v[2], v[3]), fc_pack(v[4], v[5]), fc_pack(v[6], v[7]));
            }
            if (++acc == kFcAcc) { acc = 0; acc_ph ^= 1u; }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 5) ptx::tmem_dealloc(tmem_base, 128);
}

// ---------------------------------------------------------------------------------------------------------------------
// wgrad: dW[co][c][tap] += sum_px dZ[px][co] * x[px + tap][c] ; db[co] += sum_px dZ[px][co]
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kFcQTile = 128 * 64;   // dZ tile: 128 pixel rows x 32 channels bf16 (SW64)

__global__ void __launch_bounds__(kFcThreads, 1)
first_conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmQ, const FirstConvParams p)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    // stage = [A tile 16 KB][dZ tile 8 KB]; one spare A-sized block behind the ring keeps the (ignored) second M block
    // of the last stage inside the allocation
    const int stage_bytes = kFcATile + kFcQTile;
    uint8_t* ring = smem;
    uint64_t* full_a = reinterpret_cast<uint64_t*>(ring + (size_t)p.stages * stage_bytes + kFcATile);
    uint64_t* full_q = full_a + 8;
    uint64_t* empty = full_q + 8;
    uint64_t* acc_full = empty + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.n_img * p.tiles_x * p.tiles_y;
    const int my_tiles = (int)blockIdx.x < total_tiles ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&tmQ);
        for (int s = 0; s < p.stages; ++s) { ptx::mbar_init(&full_a[s], 128); ptx::mbar_init(&full_q[s], 1); ptx::mbar_init(&empty[s], 1); }
        ptx::mbar_init(acc_full, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 5) ptx::tmem_alloc(tmem_slot, 32);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::grid_dep_wait();
    ptx::grid_dep_launch();

    if (warp < 4) {
        const int m = threadIdx.x, py = m >> 4, px = m & 15;
        int s = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int img, y0, x0;
            fc_tile_coords(p, tile, img, y0, x0);
            ptx::mbar_wait(&empty[s], ph ^ 1u);
            fc_build_row(p, ring + (size_t)s * stage_bytes, m, img, y0 + py, x0 + px);
            ptx::fence_proxy_async();
            ptx::mbar_arrive(&full_a[s]);
            if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
        // ===================== epilogue (after the last tile): rows 0..35 = dW, row 36 = db =====================
        if (my_tiles > 0 && warp < 2) {
            ptx::mbar_wait(acc_full, 0);
            ptx::tc_fence_after();
            uint32_t r[32];
            ptx::tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16), r);
            ptx::tmem_ld_wait();
            const int k = warp * 32 + lane;
            if (k < 36) {
                const int tap = k >> 2, c = k & 3;
#pragma unroll
                for (int co = 0; co < 32; ++co) atomicAdd(p.dw + (co * 4 + c) * 9 + tap, __uint_as_float(r[co]));
            } else if (k == 36) {
#pragma unroll
                for (int co = 0; co < 32; ++co) atomicAdd(p.db + co, __uint_as_float(r[co]));
            }
        }
    } else if (warp == 4) {
        // ===================== MMA issuer: D[k][co] += A^T (MN-major) * dZ (MN-major), K = 128 pixels per tile ==========
        const uint32_t idesc = ptx::make_idesc_bf16(128, 32, 1, 1);
        // A: 64 k-slots = one 128-byte M block per pixel row, 8-row groups 1024 B apart; M = 128 reads a second block
        // LBO bytes further on (the next stage's tile - finite or not, rows 64..127 of D are never read)
        const uint64_t a_desc = ptx::make_smem_desc(0, (uint32_t)stage_bytes, 1024, ptx::LAYOUT_SW128);
        const uint64_t b_desc = ptx::make_smem_desc(0, (uint32_t)kFcQTile, 512, ptx::LAYOUT_SW64);
        const uint32_t a_hi = (uint32_t)(a_desc >> 32), b_hi = (uint32_t)(b_desc >> 32);
        const uint32_t base = ptx::smem_u32(ring);
        int s = 0;
        uint32_t ph = 0;
        for (int i = 0; i < my_tiles; ++i) {
            ptx::mbar_wait(&full_a[s], ph);
            ptx::mbar_wait(&full_q[s], ph);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                const uint32_t st = base + (uint32_t)s * (uint32_t)stage_bytes;
                const uint32_t a_lo = (uint32_t)a_desc | ((st & 0x3FFFFu) >> 4);
                const uint32_t b_lo = (uint32_t)b_desc | (((st + kFcATile) & 0x3FFFFu) >> 4);
#pragma unroll
                for (int k = 0; k < 8; ++k)                 // 16 pixel rows per MMA: +2048 B in A, +1024 B in dZ
                    ptx::umma_bf16(tmem_base, ((uint64_t)a_hi << 32) | (a_lo + 128u * k), ((uint64_t)b_hi << 32) | (b_lo + 64u * k),
                                   idesc, (i | k) != 0 ? 1u : 0u);
                ptx::umma_commit(&empty[s]);
                if (i == my_tiles - 1) ptx::umma_commit(acc_full);
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
    } else if (warp == 5) {
        // ===================== TMA producer of the dZ tiles =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int img, y0, x0;
                fc_tile_coords(p, tile, img, y0, x0);
                ptx::mbar_wait(&empty[s], ph ^ 1u);
                ptx::mbar_arrive_expect_tx(&full_q[s], (uint32_t)kFcQTile);
                ptx::tma_load_5d(ring + (size_t)s * stage_bytes + kFcATile, &tmQ, &full_q[s], 0, x0, y0, img, 0);
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 5) ptx::tmem_dealloc(tmem_base, 32);
}

}  // namespace eld
