// unet_engine.cu - the UNetSeeInDark training / inference step as a fixed launch sequence over the
// tcgen05 tiles (unet_prims.cu) and the HBM-bound helpers (unet_ew.cu).
//
// Reference: models/arch/Unet.py:48-91 (forward), models/ELD_model.py:411-420,469-475 (L1 loss,
// backward, Adam).  Activations NHWC bf16; torch.cat (Unet.py:69,74,79,84) is free: the deconv and
// the encoder conv write disjoint channel ranges of one "cat" buffer.  Parameters and gradients are
// flat fp32 buffers in state_dict order (so released checkpoints map 1:1 and DDP all-reduces one buffer).
#include "common.cuh"
#include "unet_prims.h"
#include "unet_ew.h"
#include <cuda_bf16.h>
#include <cstring>
#include <vector>
#include <new>

namespace eld {

enum { L_CONV3 = 0, L_DECONV = 1, L_CONV1 = 2 };
struct Layer { const char* name; int type, cin, cout; size_t w_off, b_off, wf_off, wd_off; };

// state_dict order of UNetSeeInDark (Unet.py:11-46)
static const struct { const char* name; int type, cin, cout; } kLayers[] = {
    { "conv1_1", L_CONV3, 4, 32 },    { "conv1_2", L_CONV3, 32, 32 },   { "conv2_1", L_CONV3, 32, 64 },
    { "conv2_2", L_CONV3, 64, 64 },   { "conv3_1", L_CONV3, 64, 128 },  { "conv3_2", L_CONV3, 128, 128 },
    { "conv4_1", L_CONV3, 128, 256 }, { "conv4_2", L_CONV3, 256, 256 }, { "conv5_1", L_CONV3, 256, 512 },
    { "conv5_2", L_CONV3, 512, 512 }, { "upv6", L_DECONV, 512, 256 },   { "conv6_1", L_CONV3, 512, 256 },
    { "conv6_2", L_CONV3, 256, 256 }, { "upv7", L_DECONV, 256, 128 },   { "conv7_1", L_CONV3, 256, 128 },
    { "conv7_2", L_CONV3, 128, 128 }, { "upv8", L_DECONV, 128, 64 },    { "conv8_1", L_CONV3, 128, 64 },
    { "conv8_2", L_CONV3, 64, 64 },   { "upv9", L_DECONV, 64, 32 },     { "conv9_1", L_CONV3, 64, 32 },
    { "conv9_2", L_CONV3, 32, 32 },   { "conv10_1", L_CONV1, 32, 4 },
};
constexpr int kNumLayers = sizeof(kLayers) / sizeof(kLayers[0]);
enum { I_C11 = 0, I_C12, I_C21, I_C22, I_C31, I_C32, I_C41, I_C42, I_C51, I_C52, I_UP6, I_C61, I_C62, I_UP7,
       I_C71, I_C72, I_UP8, I_C81, I_C82, I_UP9, I_C91, I_C92, I_C10 };

constexpr int kGradBuckets = 4;
// first table entry (state_dict order without conv1_1 / conv10_1) of each bucket's tile range: upv6.., conv5_1.., conv2_1..,
// conv1_2.  The last bucket (conv1_1 + conv1_2, 42 KB) is all that is still in flight when backward ends.
constexpr int kBucketEntry0[kGradBuckets] = { 9, 7, 1, 0 };
constexpr int kBucketEntry1[kGradBuckets] = { 21, 9, 7, 1 };
constexpr int kBucketLayer0[kGradBuckets] = { 10 /*upv6*/, 8 /*conv5_1*/, 2 /*conv2_1*/, 0 /*conv1_1*/ };
constexpr int kBucketLayer1[kGradBuckets] = { 23, 10, 8, 2 };   // one past the last layer

struct PackEntry { unsigned long long src, dst_f, dst_d; int cout, cin, type; int pad; };
struct PackTable {
    PackEntry e[kNumLayers];
    int tile0[kNumLayers + 1];   // prefix sum of (cout/32 x cin/32) tiles per entry: one block per tile, whatever the layer
    int n;
    unsigned long long first_stage, first_dst, first_wf;
    int first_cin;
};

// flattened tile id -> (entry, tile inside the entry)
__device__ __forceinline__ int find_entry(const PackTable& T, int tile, int& local)
{
    int k = 0;
    while (k + 1 < T.n && tile >= T.tile0[k + 1]) ++k;
    local = tile - T.tile0[k];
    return k;
}

// packed_index (unet_prims.h) with everything that is uniform over a 32 x 32 tile hoisted: rows n0..n0+31 and K
// channels c0..c0+31 stay inside one n-tile and one channel chunk, so only the row / in-chunk channel vary.
struct PackTileBase { size_t base0, tap_stride; int r0, cc0, kc; };
__device__ __forceinline__ PackTileBase pack_tile_base(int rows, int ck, int taps, int n0, int c0)
{
    const int n_tile = rows <= 256 ? rows : 256;
    const int kc = (ck % 64 == 0) ? 64 : 32;
    const int kchunks = ck / kc;
    const int nt = n0 / n_tile, chunk = c0 / kc;
    PackTileBase b;
    b.tap_stride = (size_t)kchunks * n_tile * kc;
    b.base0 = ((size_t)nt * taps * kchunks + chunk) * ((size_t)n_tile * kc);
    b.r0 = n0 - nt * n_tile; b.cc0 = c0 - chunk * kc; b.kc = kc;
    return b;
}
__device__ __forceinline__ size_t pack_tile_index(const PackTileBase& b, int tap, int dn, int dc)
{
    const int r = b.r0 + dn, cc = b.cc0 + dc, rb = b.kc * 2;
    const int swz = rb == 128 ? (r & 7) : ((r >> 1) & 3);
    const int byte = r * rb + ((((cc * 2) >> 4) ^ swz) << 4) + ((cc * 2) & 15);
    return b.base0 + (size_t)tap * b.tap_stride + (size_t)(byte >> 1);
}

// one launch packs every layer's fp32 master weights into both bf16 GEMM operands (fprop + dgrad).
// A block moves a (32 x 32 x taps) tile through shared memory so that reads are 1 KB runs and writes are
// 64-byte runs in both destination layouts.
__global__ void __launch_bounds__(256)
pack_all_kernel(const float* __restrict__ params, __nv_bfloat16* __restrict__ packed, const __grid_constant__ PackTable T)
{
    __shared__ float tile[32][32 * 9 + 1];
    if ((int)blockIdx.x >= T.tile0[T.n]) {   // conv1_1: w[32][4][9] -> K-major operand [32 co][64 k], k = tap*4 + c (first_conv.cuh),
        const float* w1 = params + T.first_dst;   // 128-byte rows with the SW128 swizzle; k >= 36 are zeros (k = 36 meets the ones column)
        const int b = (int)blockIdx.x - T.tile0[T.n], nb = (int)gridDim.x - T.tile0[T.n];
        for (int i = b * 256 + threadIdx.x; i < 32 * 64; i += nb * 256) {
            const int co = i >> 6, k = i & 63, tap = k >> 2, c = k & 3;
            const float v = (k < 36 && c < T.first_cin) ? w1[(co * T.first_cin + c) * 9 + tap] : 0.0f;
            packed[T.first_wf + (size_t)co * 64 + ((((k >> 3) ^ (co & 7)) << 3) | (k & 7))] = __float2bfloat16_rn(v);
        }
        return;
    }
    int tl;
    const PackEntry& e = T.e[find_entry(T, (int)blockIdx.x, tl)];
    const float* w = params + e.src;
    __nv_bfloat16* of = packed + e.dst_f;
    __nv_bfloat16* od = packed + e.dst_d;
    if (e.type == L_CONV3) {              // w[co][ci][t]
        const int it = e.cin / 32;
        {
            const int co0 = (tl / it) * 32, ci0 = (tl % it) * 32;
            if ((reinterpret_cast<uintptr_t>(params) & 15) == 0) {     // layer offsets are multiples of 4 floats: 16-byte loads
                for (int i = threadIdx.x; i < 32 * 72; i += 256) {
                    const int co = i / 72, r4 = i - co * 72;
                    const float4 v = __ldg(reinterpret_cast<const float4*>(w + ((size_t)(co0 + co) * e.cin + ci0) * 9) + r4);
                    float* t4 = &tile[co][4 * r4];
                    t4[0] = v.x; t4[1] = v.y; t4[2] = v.z; t4[3] = v.w;
                }
            } else {
                for (int i = threadIdx.x; i < 32 * 288; i += 256) {
                    const int co = i / 288, r = i - co * 288;              // r = ci*9 + t
                    tile[co][r] = w[((size_t)(co0 + co) * e.cin + ci0) * 9 + r];
                }
            }
            __syncthreads();
            // eight adjacent K elements per thread = one 16-byte chunk of the swizzled row (chunks are what the swizzle permutes)
            const PackTileBase bf = pack_tile_base(e.cout, e.cin, 9, co0, ci0), bd = pack_tile_base(e.cin, e.cout, 9, ci0, co0);
            for (int i = threadIdx.x; i < 32 * 36; i += 256) {        // fprop: [co][t][ci]
                const int ci = (i & 3) * 8, t = (i >> 2) % 9, co = i / 36;
                uint32_t q[4];
#pragma unroll
                for (int e2 = 0; e2 < 4; ++e2) {
                    const __nv_bfloat162 v2 = __floats2bfloat162_rn(tile[co][(ci + 2 * e2) * 9 + t], tile[co][(ci + 2 * e2 + 1) * 9 + t]);
                    q[e2] = *reinterpret_cast<const uint32_t*>(&v2);
                }
                *reinterpret_cast<uint4*>(of + pack_tile_index(bf, t, co, ci)) = make_uint4(q[0], q[1], q[2], q[3]);
            }
            for (int i = threadIdx.x; i < 32 * 36; i += 256) {        // dgrad: [ci][8-t][co]
                const int co = (i & 3) * 8, t = (i >> 2) % 9, ci = i / 36;
                uint32_t q[4];
#pragma unroll
                for (int e2 = 0; e2 < 4; ++e2) {
                    const __nv_bfloat162 v2 = __floats2bfloat162_rn(tile[co + 2 * e2][ci * 9 + t], tile[co + 2 * e2 + 1][ci * 9 + t]);
                    q[e2] = *reinterpret_cast<const uint32_t*>(&v2);
                }
                *reinterpret_cast<uint4*>(od + pack_tile_index(bd, 8 - t, ci, co)) = make_uint4(q[0], q[1], q[2], q[3]);
            }
        }
    } else {                              // deconv wt[ci][co][s]
        const int ct = e.cout / 32;
        {
            const int ci0 = (tl / ct) * 32, co0 = (tl % ct) * 32;
            for (int i = threadIdx.x; i < 32 * 128; i += 256) {
                const int ci = i / 128, r = i - ci * 128;              // r = co*4 + s
                tile[ci][r] = w[((size_t)(ci0 + ci) * e.cout + co0) * 4 + r];
            }
            __syncthreads();
            for (int i = threadIdx.x; i < 32 * 128; i += 256) {       // fprop: [(s*cout + co)][ci]
                const int ci = i & 31, co = (i >> 5) & 31, sp = i >> 10;
                of[packed_index(4 * e.cout, e.cin, 1, sp * e.cout + co0 + co, 0, ci0 + ci)] = __float2bfloat16_rn(tile[ci][co * 4 + sp]);
            }
            for (int i = threadIdx.x; i < 32 * 128; i += 256) {       // dgrad: [ci][s][co]
                const int co = i & 31, sp = (i >> 5) & 3, ci = i >> 7;
                od[packed_index(e.cin, e.cout, 4, ci0 + ci, sp, co0 + co)] = __float2bfloat16_rn(tile[ci][co * 4 + sp]);
            }
        }
    }
}

// staging [tap][ci][co] -> PyTorch OIHW [co][ci][tap]; one block per (32 co x 32 ci) tile of one layer
__global__ void __launch_bounds__(256)
wgrad_permute_kernel(const float* __restrict__ gtmp, float* __restrict__ grads, const __grid_constant__ PackTable T,
                     int tile_begin, int tile_end)
{
    // blocks [0, tile_end - tile_begin) move the table tiles [tile_begin, tile_end) (one gradient bucket)
    __shared__ float tile[9][32][33];
    const int gtile = (int)blockIdx.x + tile_begin;
    if (gtile >= tile_end) return;
    int t;
    const PackEntry& e = T.e[find_entry(T, gtile, t)];
    if (e.type != L_CONV3) return;
    const int ct = e.cout / 32;
    {
        const int co0 = (t % ct) * 32, ci0 = (t / ct) * 32;
        // 16-byte accesses on both sides (layer offsets are multiples of 4 floats; the staging buffer is 1 KB aligned)
        for (int i = threadIdx.x; i < 9 * 32 * 8; i += 256) {
            const int co4 = i & 7, ci = (i >> 3) & 31, tap = i >> 8;
            const float4 v = __ldg(reinterpret_cast<const float4*>(gtmp + e.src + ((size_t)tap * e.cin + ci0 + ci) * e.cout + co0) + co4);
            float* t4 = &tile[tap][ci][4 * co4];
            t4[0] = v.x; t4[1] = v.y; t4[2] = v.z; t4[3] = v.w;
        }
        __syncthreads();
        if ((reinterpret_cast<uintptr_t>(grads) & 15) == 0) {
            for (int i = threadIdx.x; i < 32 * 72; i += 256) {
                const int co = i / 72, r4 = i - co * 72;           // r = ci*9 + tap, contiguous in OIHW
                float q[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int r = 4 * r4 + k, ci = r / 9, tap = r - ci * 9;
                    q[k] = tile[tap][ci][co];
                }
                *(reinterpret_cast<float4*>(grads + e.src + ((size_t)(co0 + co) * e.cin + ci0) * 9) + r4) = make_float4(q[0], q[1], q[2], q[3]);
            }
        } else {
            for (int i = threadIdx.x; i < 32 * 288; i += 256) {
                const int co = i / 288, r = i - co * 288;
                const int ci = r / 9, tap = r - ci * 9;
                grads[e.src + ((size_t)(co0 + co) * e.cin + ci0 + ci) * 9 + tap] = tile[tap][ci][co];
            }
        }
    }
}

}  // namespace eld

using namespace eld;

struct eld_unet {
    eld_ctx* ctx;
    int n, H, W;
    Layer L[kNumLayers];
    size_t n_params;
    char* ws;
    size_t ws_bytes;
    // activations / gradients (bf16), offsets in bytes into ws
    __nv_bfloat16 *a1_1, *cat9, *p1, *a2_1, *cat8, *p2, *a3_1, *cat7, *p3, *a4_1, *cat6, *p4, *a5_1, *a5_2,
        *a6_1, *a6_2, *a7_1, *a7_2, *a8_1, *a8_2, *a9_1, *a9_2;
    __nv_bfloat16 *dz9_2, *dz9_1, *dcat9, *dz8_2, *dz8_1, *dcat8, *dz7_2, *dz7_1, *dcat7, *dz6_2, *dz6_1, *dcat6,
        *dz5_2, *dz5_1, *dp4, *dz4_2, *dz4_1, *dp3, *dz3_2, *dz3_1, *dp2, *dz2_2, *dz2_1, *dp1, *dz1_2, *dz1_1;
    __nv_bfloat16 *pc1 = nullptr, *pc2 = nullptr, *pc3 = nullptr, *pc4 = nullptr;   // pool codes, 1 byte per pooled element (training)
    // sign words (1 bit per element) of the activations whose LeakyReLU' a data gradient applies (training): the dgrad
    // tiles read these instead of the activation itself
    struct SignBuf { const void* act; uint32_t* words; } signs[16];
    int n_signs = 0;
    __nv_bfloat16* packed;
    float* gtmp = nullptr;
    PackTable table;
    // gradient buckets in backward-completion order (data-parallel overlap): decoder, bottleneck, encoder
    cudaEvent_t bucket_ev[kGradBuckets] = { nullptr, nullptr, nullptr, nullptr };
    int cin0 = 4, cout_last = 4; // channels of the frame in / out: 4 = packed raw, 3 = sRGB (ELD_model.py:377-389)
    int l2_loss = 0;             // 0: nn.L1Loss (the reference default, losses.py:31-32), 1: nn.MSELoss (losses.py:33-34)
    // optional per-launch profile (CUDA events on the launch stream)
    bool profile = false;
    struct Rec { char name[32]; double flops, bytes; cudaEvent_t e0, e1; };
    std::vector<Rec> recs;
    size_t rec_used = 0;
};

static size_t layout(eld_unet* u, char* base, bool train)
{
    size_t off = 0;
    const size_t n = u->n;
    auto take = [&](__nv_bfloat16** p, int lvl, int ch) {
        const size_t h = u->H >> lvl, w = u->W >> lvl;
        const size_t bytes = (n * h * w * ch * 2 + 1023) & ~(size_t)1023;
        if (p) *p = reinterpret_cast<__nv_bfloat16*>(base + off);
        off += bytes;
    };
    take(&u->a1_1, 0, 32); take(&u->cat9, 0, 64); take(&u->p1, 1, 32);
    take(&u->a2_1, 1, 64); take(&u->cat8, 1, 128); take(&u->p2, 2, 64);
    take(&u->a3_1, 2, 128); take(&u->cat7, 2, 256); take(&u->p3, 3, 128);
    take(&u->a4_1, 3, 256); take(&u->cat6, 3, 512); take(&u->p4, 4, 256);
    take(&u->a5_1, 4, 512); take(&u->a5_2, 4, 512);
    take(&u->a6_1, 3, 256); take(&u->a6_2, 3, 256); take(&u->a7_1, 2, 128); take(&u->a7_2, 2, 128);
    take(&u->a8_1, 1, 64); take(&u->a8_2, 1, 64); take(&u->a9_1, 0, 32); take(&u->a9_2, 0, 32);
    if (train) {
        take(&u->dz9_2, 0, 32); take(&u->dz9_1, 0, 32); take(&u->dcat9, 0, 64);
        take(&u->dz8_2, 1, 64); take(&u->dz8_1, 1, 64); take(&u->dcat8, 1, 128);
        take(&u->dz7_2, 2, 128); take(&u->dz7_1, 2, 128); take(&u->dcat7, 2, 256);
        take(&u->dz6_2, 3, 256); take(&u->dz6_1, 3, 256); take(&u->dcat6, 3, 512);
        take(&u->dz5_2, 4, 512); take(&u->dz5_1, 4, 512); take(&u->dp4, 4, 256);
        take(&u->dz4_2, 3, 256); take(&u->dz4_1, 3, 256); take(&u->dp3, 3, 128);
        take(&u->dz3_2, 2, 128); take(&u->dz3_1, 2, 128); take(&u->dp2, 2, 64);
        take(&u->dz2_2, 1, 64); take(&u->dz2_1, 1, 64); take(&u->dp1, 1, 32);
        take(&u->dz1_2, 0, 32); take(&u->dz1_1, 0, 32);
        take(&u->pc1, 1, 16); take(&u->pc2, 2, 32); take(&u->pc3, 3, 64); take(&u->pc4, 4, 128);
        u->n_signs = 0;
        auto take_signs = [&](__nv_bfloat16* act, int lvl, int ch) {      // ch / 32 words per pixel = ch / 16 bf16-sized units
            __nv_bfloat16* w = nullptr;
            take(&w, lvl, ch / 16);
            u->signs[u->n_signs++] = { act, reinterpret_cast<uint32_t*>(w) };
        };
        take_signs(u->a1_1, 0, 32); take_signs(u->a2_1, 1, 64); take_signs(u->a3_1, 2, 128); take_signs(u->a4_1, 3, 256);
        take_signs(u->a5_1, 4, 512); take_signs(u->a5_2, 4, 512); take_signs(u->a6_1, 3, 256); take_signs(u->a6_2, 3, 256);
        take_signs(u->a7_1, 2, 128); take_signs(u->a7_2, 2, 128); take_signs(u->a8_1, 1, 64); take_signs(u->a8_2, 1, 64);
        take_signs(u->a9_1, 0, 32);
    }
    // packed weights
    size_t pk = 0;
    u->L[I_C11].wf_off = pk; pk += 32 * 9 * 32;          // conv1_1: input padded to 32 channels, fprop only
    for (int i = 0; i < kNumLayers; ++i) {
        Layer& l = u->L[i];
        if (i == I_C11 || l.type == L_CONV1) continue;
        const size_t cnt = (size_t)l.cin * l.cout * (l.type == L_CONV3 ? 9 : 4);
        l.wf_off = pk; pk += cnt;
        l.wd_off = pk; pk += cnt;
    }
    u->packed = reinterpret_cast<__nv_bfloat16*>(base + off);
    off += (pk * 2 + 1023) & ~(size_t)1023;
    if (train) {   // [tap][ci][co] staging of the conv3x3 weight gradients (same offsets as the fp32 parameters)
        u->gtmp = reinterpret_cast<float*>(base + off);
        off += (u->n_params * 4 + 1023) & ~(size_t)1023;
    }
    return off;
}

static void init_layers(eld_unet* u)
{
    size_t off = 0;
    for (int i = 0; i < kNumLayers; ++i) {
        Layer& l = u->L[i];
        l.name = kLayers[i].name; l.type = kLayers[i].type; l.cin = kLayers[i].cin; l.cout = kLayers[i].cout;
        if (i == 0) l.cin = u->cin0;
        if (i == kNumLayers - 1) l.cout = u->cout_last;
        const size_t ksz = l.type == L_CONV3 ? 9 : (l.type == L_DECONV ? 4 : 1);
        l.w_off = off; off += (size_t)l.cin * l.cout * ksz;
        l.b_off = off; off += l.cout;
        l.wf_off = l.wd_off = 0;
    }
    u->n_params = off;
}

static bool io_ok(int cin, int cout) { return (cin == 3 || cin == 4) && (cout == 3 || cout == 4); }

extern "C" size_t eld_unet_param_count_io(int cin, int cout)
{
    if (!io_ok(cin, cout)) return 0;
    eld_unet tmp{};
    tmp.cin0 = cin; tmp.cout_last = cout;
    init_layers(&tmp);
    return tmp.n_params;
}
extern "C" size_t eld_unet_param_count(void) { return eld_unet_param_count_io(4, 4); }

extern "C" int eld_unet_param_offset_io(const char* name, int is_bias, int cin, int cout, size_t* offset, size_t* count);
extern "C" int eld_unet_param_offset(const char* name, int is_bias, size_t* offset, size_t* count)
{
    return eld_unet_param_offset_io(name, is_bias, 4, 4, offset, count);
}
extern "C" int eld_unet_param_offset_io(const char* name, int is_bias, int cin, int cout, size_t* offset, size_t* count)
{
    ELD_REQUIRE(io_ok(cin, cout), "eld_unet_param_offset: channels in / out must be 3 or 4");
    eld_unet tmp{};
    tmp.cin0 = cin; tmp.cout_last = cout;
    init_layers(&tmp);
    for (int i = 0; i < kNumLayers; ++i) {
        if (strcmp(name, tmp.L[i].name) == 0) {
            const Layer& l = tmp.L[i];
            const size_t ksz = l.type == L_CONV3 ? 9 : (l.type == L_DECONV ? 4 : 1);
            if (offset) *offset = is_bias ? l.b_off : l.w_off;
            if (count) *count = is_bias ? (size_t)l.cout : (size_t)l.cin * l.cout * ksz;
            return ELD_OK;
        }
    }
    set_error("eld_unet_param_offset: unknown layer '%s'", name);
    return ELD_E_ARG;
}

extern "C" size_t eld_unet_workspace_bytes(int n, int h, int w, int train)
{
    eld_unet tmp{};
    tmp.n = n; tmp.H = h; tmp.W = w;
    init_layers(&tmp);
    return layout(&tmp, nullptr, train != 0) + 1024;
}

extern "C" int eld_unet_create_io(eld_ctx* ctx, int n, int h, int w, int train, void* workspace, size_t bytes,
                                  int cin, int cout, eld_unet** out);
extern "C" int eld_unet_create(eld_ctx* ctx, int n, int h, int w, int train, void* workspace, size_t bytes, eld_unet** out)
{
    return eld_unet_create_io(ctx, n, h, w, train, workspace, bytes, 4, 4, out);
}
extern "C" int eld_unet_create_io(eld_ctx* ctx, int n, int h, int w, int train, void* workspace, size_t bytes,
                                  int cin, int cout, eld_unet** out)
{
    ELD_REQUIRE(ctx && out && workspace, "eld_unet_create: NULL argument");
    ELD_REQUIRE(io_ok(cin, cout), "eld_unet_create: channels in / out must be 3 (sRGB) or 4 (packed raw); got %d -> %d", cin, cout);
    ELD_REQUIRE(n > 0 && h > 0 && w > 0, "eld_unet_create: bad shape");
    ELD_REQUIRE(h % 16 == 0 && w % 16 == 0, "eld_unet_create: H and W must be multiples of 16 (four 2x2 pools, like the reference); got %dx%d", h, w);
    ELD_CHECK_CUDA(cudaSetDevice(ctx->device));
    ELD_REQUIRE(!train || (h % 128 == 0 && w % 256 == 0),
                "eld_unet_create: TRAINING needs H %% 128 == 0 and W %% 256 == 0 (whole 8x16 / 8x8 gradient tiles at 1/16 scale); got %dx%d", h, w);
    eld_unet* u = new (std::nothrow) eld_unet();
    ELD_REQUIRE(u, "eld_unet_create: out of host memory");
    u->ctx = ctx; u->n = n; u->H = h; u->W = w; u->cin0 = cin; u->cout_last = cout;
    init_layers(u);
    char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    const size_t need = layout(u, base, train != 0) + (base - static_cast<char*>(workspace));
    if (need > bytes) {
        set_error("eld_unet_create: workspace %zu bytes < required %zu", bytes, need);
        delete u;
        return ELD_E_WORKSPACE;
    }
    u->ws = base; u->ws_bytes = bytes;
    if (!train) u->dz9_2 = nullptr;
    int k = 0;
    for (int i = 0; i < kNumLayers; ++i) {
        const Layer& l = u->L[i];
        if (i == I_C11 || l.type == L_CONV1) continue;
        u->table.tile0[k] = k == 0 ? 0 : u->table.tile0[k - 1] + (u->table.e[k - 1].cout / 32) * (u->table.e[k - 1].cin / 32);
        u->table.e[k++] = PackEntry{ l.w_off, l.wf_off, l.wd_off, l.cout, l.cin, l.type, 0 };
    }
    u->table.tile0[k] = u->table.tile0[k - 1] + (u->table.e[k - 1].cout / 32) * (u->table.e[k - 1].cin / 32);
    u->table.n = k;
    u->table.first_stage = u->n_params;
    u->table.first_dst = u->L[I_C11].w_off;
    u->table.first_wf = u->L[I_C11].wf_off;
    u->table.first_cin = u->cin0;
    // conv1_1's operand image: zero once (pack_all rewrites all of it every step anyway)
    ELD_CHECK_CUDA(cudaMemset(u->packed + u->L[I_C11].wf_off, 0, 32 * 9 * 32 * sizeof(__nv_bfloat16)));
    // opt in to large dynamic shared memory once (not inside a captured region)
    { int rc = init_gemm_kernels(ctx); if (rc != ELD_OK) { delete u; return rc; } }
    *out = u;
    return ELD_OK;
}

extern "C" void eld_unet_destroy(eld_unet* u)
{
    if (!u) return;
    for (int k = 0; k < kGradBuckets; ++k) if (u->bucket_ev[k]) cudaEventDestroy(u->bucket_ev[k]);
    for (auto& r : u->recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    delete u;
}

extern "C" int eld_unet_grad_buckets_io(int cin, int cout, size_t* offsets, int max_offsets);
extern "C" int eld_unet_grad_buckets(size_t* offsets, int max_offsets) { return eld_unet_grad_buckets_io(4, 4, offsets, max_offsets); }
extern "C" int eld_unet_grad_buckets_io(int cin, int cout, size_t* offsets, int max_offsets)
{
    ELD_REQUIRE(io_ok(cin, cout), "eld_unet_grad_buckets: channels in / out must be 3 or 4");
    eld_unet tmp{};
    tmp.cin0 = cin; tmp.cout_last = cout;
    init_layers(&tmp);
    if (offsets) {
        ELD_REQUIRE(max_offsets >= 2 * kGradBuckets, "eld_unet_grad_buckets: need room for %d (offset, count) pairs", kGradBuckets);
        for (int k = 0; k < kGradBuckets; ++k) {
            const size_t b = tmp.L[kBucketLayer0[k]].w_off;
            const size_t e = kBucketLayer1[k] == kNumLayers ? tmp.n_params : tmp.L[kBucketLayer1[k]].w_off;
            offsets[2 * k] = b; offsets[2 * k + 1] = e - b;
        }
    }
    return kGradBuckets;
}

extern "C" int eld_unet_bucket_events(eld_unet* u, int enable)
{
    ELD_REQUIRE(u, "eld_unet_bucket_events: NULL");
    ELD_CHECK_CUDA(cudaSetDevice(u->ctx->device));
    for (int k = 0; k < kGradBuckets; ++k) {
        if (enable && !u->bucket_ev[k]) ELD_CHECK_CUDA(cudaEventCreateWithFlags(&u->bucket_ev[k], cudaEventDisableTiming));
        if (!enable && u->bucket_ev[k]) { cudaEventDestroy(u->bucket_ev[k]); u->bucket_ev[k] = nullptr; }
    }
    return ELD_OK;
}

extern "C" int eld_unet_wait_bucket(eld_unet* u, int bucket, void* stream)
{
    ELD_REQUIRE(u && bucket >= 0 && bucket < kGradBuckets, "eld_unet_wait_bucket: bad bucket %d", bucket);
    ELD_REQUIRE(u->bucket_ev[bucket], "eld_unet_wait_bucket: call eld_unet_bucket_events(u, 1) before the step");
    ELD_CHECK_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), u->bucket_ev[bucket], 0));
    return ELD_OK;
}

#define TRY(expr) do { int _rc = (expr); if (_rc != ELD_OK) return _rc; } while (0)

namespace {

struct Scope {
    eld_unet* u; cudaStream_t st; eld_unet::Rec* r = nullptr;
    Scope(eld_unet* u_, cudaStream_t st_, const char* layer, const char* what, double flops, double bytes) : u(u_), st(st_)
    {
        if (!u->profile) return;
        if (u->rec_used == u->recs.size()) {
            eld_unet::Rec n{};
            cudaEventCreate(&n.e0); cudaEventCreate(&n.e1);
            u->recs.push_back(n);
        }
        r = &u->recs[u->rec_used++];
        snprintf(r->name, sizeof(r->name), "%s.%s", layer, what);
        r->flops = flops; r->bytes = bytes;
        cudaEventRecord(r->e0, st);
    }
    ~Scope() { if (r) cudaEventRecord(r->e1, st); }
};

struct Runner {
    eld_unet* u;
    const float* params;
    cudaStream_t st;
    eld_ctx* ctx() const { return u->ctx; }
    const __nv_bfloat16* wf(int i) const { return u->packed + u->L[i].wf_off; }
    const __nv_bfloat16* wd(int i) const { return u->packed + u->L[i].wd_off; }
    const float* bias(int i) const { return params + u->L[i].b_off; }

    // sign words of activation `act` (training engines; ELD_MASK_FROM_ACT=1 keeps the masks that re-read the activations: A/B)
    uint32_t* sign_of(const void* act) const
    {
        static const bool off = getenv("ELD_MASK_FROM_ACT") != nullptr;
        if (off || !u->dz9_2) return nullptr;
        for (int i = 0; i < u->n_signs; ++i) if (u->signs[i].act == act) return u->signs[i].words;
        return nullptr;
    }
    // pool_dst != nullptr: MaxPool2d(2) of the output fused into the tile's epilogue (pooled tensor has cout channels)
    int conv(int li, const void* x, int xp, int xc0, void* y, int yp, int yc0, int lvl, void* pool_dst = nullptr, void* pool_code = nullptr) const
    {
        const Layer& l = u->L[li];
        GemmOp op{};
        op.a = x; op.a_pitch = xp; op.a_c0 = xc0; op.a_mode = A_CONV; op.taps = 9; op.cin = l.cin;
        op.n_img = u->n; op.H = u->H >> lvl; op.W = u->W >> lvl;
        op.b = wf(li); op.n_total = l.cout; op.cout = l.cout;
        op.epi_mode = EPI_STORE; op.act = ACT_LRELU; op.out = y; op.out_pitch = yp; op.out_c0 = yc0; op.bias = bias(li);
        op.pool_out = pool_dst; op.pool_pitch = l.cout; op.pool_code = pool_code;
        if (yc0 == 0 && yp == l.cout) op.sign_out = sign_of(y);
        const double px = (double)u->n * op.H * op.W;
        Scope sc(u, st, l.name, "fprop", 2.0 * px * l.cout * 9 * l.cin, px * 2 * (l.cin + l.cout) + 18.0 * l.cin * l.cout);
        return launch_conv_gemm(ctx(), op, st);
    }
    int deconv(int li, const void* x, int xp, void* y, int yp, int lvl_in) const
    {
        const Layer& l = u->L[li];
        GemmOp op{};
        op.a = x; op.a_pitch = xp; op.a_c0 = 0; op.a_mode = A_CONV; op.taps = 1; op.cin = l.cin;
        op.n_img = u->n; op.H = u->H >> lvl_in; op.W = u->W >> lvl_in;
        op.b = wf(li); op.n_total = 4 * l.cout; op.cout = l.cout;
        op.epi_mode = EPI_SHUFFLE; op.act = ACT_NONE; op.out = y; op.out_pitch = yp; op.out_c0 = 0; op.bias = bias(li);
        const double px = (double)u->n * op.H * op.W;
        Scope sc(u, st, l.name, "fprop", 2.0 * px * 4 * l.cout * l.cin, px * 2 * (l.cin + 4 * l.cout) + 8.0 * l.cin * l.cout);
        return launch_conv_gemm(ctx(), op, st);
    }
    // data gradient of a conv: dz [cout] -> d(input) [cin channels at dxc0], optional lrelu' mask from `act_src`
    // dx2 != nullptr: the gradient of a concat input is stored as two PLANAR halves (dx = [up], dx2 = [skip], pitch cin/2
    // each) - every consumer reads exactly one half, and an interleaved buffer made each of them move whole 128-byte
    // lines for 64 useful bytes (ncu r01: level-1 pool.bwd 570 MB for 300, upv9 dgrad/wgrad 336 MB for 200)
    int conv_dgrad(int li, const void* dz, void* dx, int dxp, int dxc0, const void* act_src, int asp, int asc0, int lvl,
                   void* dx2 = nullptr) const
    {
        const Layer& l = u->L[li];
        GemmOp op{};
        op.a = dz; op.a_pitch = l.cout; op.a_c0 = 0; op.a_mode = A_CONV; op.taps = 9; op.cin = l.cout;
        op.n_img = u->n; op.H = u->H >> lvl; op.W = u->W >> lvl;
        op.b = wd(li); op.n_total = l.cin; op.cout = l.cin;
        op.epi_mode = EPI_STORE; op.act = act_src ? ACT_MASK : ACT_NONE;
        op.out = dx; op.out_pitch = dxp; op.out_c0 = dxc0; op.bias = nullptr;
        if (dx2) { op.out2 = dx2; op.out2_pitch = dxp; op.out_split = l.cin / 2; }
        op.aux = act_src; op.aux_pitch = asp; op.aux_c0 = asc0;
        if (act_src && asc0 == 0 && asp == l.cin) op.aux_sign = sign_of(act_src);
        const double px = (double)u->n * op.H * op.W;
        Scope sc(u, st, l.name, "dgrad", 2.0 * px * l.cout * 9 * l.cin, px * 2 * (l.cin * (act_src ? 2 : 1) + l.cout) + 18.0 * l.cin * l.cout);
        return launch_conv_gemm(ctx(), op, st);
    }
    int deconv_dgrad(int li, const void* dy, int dyp, void* dx, const void* act_src, int lvl_in) const
    {
        const Layer& l = u->L[li];
        GemmOp op{};
        op.a = dy; op.a_pitch = dyp; op.a_c0 = 0; op.a_mode = A_GATHER; op.taps = 4; op.cin = l.cout;
        op.n_img = u->n; op.H = u->H >> lvl_in; op.W = u->W >> lvl_in;
        op.b = wd(li); op.n_total = l.cin; op.cout = l.cin;
        op.epi_mode = EPI_STORE; op.act = ACT_MASK; op.out = dx; op.out_pitch = l.cin; op.out_c0 = 0;
        op.aux = act_src; op.aux_pitch = l.cin; op.aux_c0 = 0; op.aux_sign = sign_of(act_src);
        const double px = (double)u->n * op.H * op.W;
        Scope sc(u, st, l.name, "dgrad", 2.0 * px * 4 * l.cout * l.cin, px * 2 * (2 * l.cin + 4 * l.cout) + 8.0 * l.cin * l.cout);
        return launch_conv_gemm(ctx(), op, st);
    }
    int conv_wgrad(int li, const void* x, int xp, int xc0, const void* dz, float* grads, int lvl) const
    {
        const Layer& l = u->L[li];
        WgradOp op{};
        op.mode = WG_CONV; op.p = x; op.p_pitch = xp; op.p_c0 = xc0; op.p_ch = l.cin;
        op.q = dz; op.q_pitch = l.cout; op.q_c0 = 0; op.q_ch = l.cout;
        op.n_img = u->n; op.H = u->H >> lvl; op.W = u->W >> lvl;
        op.dw = u->gtmp + l.w_off; op.out_tco = 1;
        op.db = grads + l.b_off;                     // bias gradient fused into the same launch
        const double px = (double)u->n * op.H * op.W;
        Scope sc(u, st, l.name, "wgrad", 2.0 * px * l.cout * 9 * l.cin, px * 2 * (l.cin + l.cout) + 36.0 * l.cin * l.cout);
        return launch_wgrad(ctx(), op, st);
    }
    int deconv_wgrad(int li, const void* x, const void* dy, int dyp, float* grads, int lvl_in) const
    {
        const Layer& l = u->L[li];
        WgradOp op{};
        op.mode = WG_DECONV; op.p = dy; op.p_pitch = dyp; op.p_c0 = 0; op.p_ch = l.cout;
        op.q = x; op.q_pitch = l.cin; op.q_c0 = 0; op.q_ch = l.cin;
        op.n_img = u->n; op.H = u->H >> lvl_in; op.W = u->W >> lvl_in; op.dw = grads + l.w_off;
        static const bool fuse_bias = getenv("ELD_DECONV_COLSUM") == nullptr;    // (A/B: the stand-alone column-sum kernel)
        op.db = fuse_bias ? grads + l.b_off : nullptr;           // bias gradient = column sums of the d(up) boxes, same launch
        const double px = (double)u->n * op.H * op.W;
        {
            Scope sc(u, st, l.name, "wgrad", 2.0 * px * 4 * l.cout * l.cin, px * 2 * (l.cin + 4 * l.cout) + 16.0 * l.cin * l.cout);
            TRY(launch_wgrad(ctx(), op, st));
        }
        if (fuse_bias) return ELD_OK;
        Scope sc(u, st, l.name, "bgrad", 0.0, px * 8 * l.cout);
        return launch_colsum(ctx(), dy, dyp, 0, l.cout, (size_t)u->n * op.H * op.W * 4, grads + l.b_off, st);
    }
    int pool(const void* in, int pitch, int c0, void* out, int C, int lvl_out) const
    {
        const double pxo = (double)u->n * (u->H >> lvl_out) * (u->W >> lvl_out);
        Scope sc(u, st, "pool", "fwd", 0.0, pxo * C * 2 * 5);
        return launch_maxpool(ctx(), in, pitch, c0, out, C, u->n, u->H >> lvl_out, u->W >> lvl_out, st);
    }
    // A = the interleaved concat buffer's skip half (pitch 2C, offset C); dskip = the PLANAR skip half of its gradient
    // code != nullptr: the forward tile left the argmax + sign code, the activation is not read again
    int pool_bwd(const void* A, int pitch, int c0, const void* dskip, const void* dP, void* dZ, int C, int lvl_out, const void* code) const
    {
        const double pxo = (double)u->n * (u->H >> lvl_out) * (u->W >> lvl_out);
        Scope sc(u, st, "pool", "bwd", 0.0, pxo * C * (code ? 2 * 9 + 1 : 2 * 13));
        if (code) return launch_maxpool_bwd_code(ctx(), code, dskip, C, 0, dP, dZ, C, u->n, u->H >> lvl_out, u->W >> lvl_out, st);
        return launch_maxpool_bwd(ctx(), A, pitch, c0, dskip, C, 0, dP, dZ, C, u->n, u->H >> lvl_out, u->W >> lvl_out, st);
    }
    // second (skip) half of a planar concat gradient: [n][h][w][C] right behind the up half
    __nv_bfloat16* skip_half(__nv_bfloat16* dcat, int lvl, int C) const
    {
        return dcat + (size_t)u->n * (u->H >> lvl) * (u->W >> lvl) * C;
    }

    // gradients of bucket k are complete in the staging area: move its conv tiles to the PyTorch layout and mark the
    // bucket final (a data-parallel caller all-reduces it on a side stream while the rest of backward runs)
    int finish_bucket(int k, float* g) const
    {
        // Single-GPU steps (no bucket events) move all conv gradients with ONE launch at the end: a permute launch over a
        // few dozen tiles is latency-bound (~15 us whatever its size), four of them cost 61 us against 27 us for one.
        // Data-parallel steps (bucket events on): the bucket's conv gradients move to the PyTorch layout right here, on
        // the compute stream, then an event marks the bucket final.  (Running this permute on the caller's communication
        // stream instead was measured at 2 GPUs: +0.10 ms per step - any foreign kernel that overlaps the persistent
        // one-CTA-per-SM tiles delays some of their CTAs, and a tile kernel is as slow as its slowest CTA.)
        const bool per_bucket = u->bucket_ev[0] != nullptr;
        if (!per_bucket && k != kGradBuckets - 1) return ELD_OK;
        const int t0 = per_bucket ? u->table.tile0[kBucketEntry0[k]] : 0;
        const int t1 = per_bucket ? u->table.tile0[kBucketEntry1[k]] : u->table.tile0[u->table.n];
        {
            Scope sc(u, st, "weights", "gperm", 0.0, 0.0);
            wgrad_permute_kernel<<<t1 - t0, 256, 0, st>>>(u->gtmp, g, u->table, t0, t1);
            ELD_CHECK_CUDA(cudaGetLastError());
            count_launch(ctx());
        }
        if (u->bucket_ev[k]) ELD_CHECK_CUDA(cudaEventRecord(u->bucket_ev[k], st));
        return ELD_OK;
    }

    int pack() const
    {
        Scope sc(u, st, "weights", "pack", 0.0, (double)u->n_params * 8);
        pack_all_kernel<<<u->table.tile0[u->table.n] + 2, 256, 0, st>>>(params, u->packed, u->table);
        ELD_CHECK_CUDA(cudaGetLastError());
        count_launch(ctx());
        return ELD_OK;
    }

    int forward(const float* x) const
    {
        eld_unet* U = u;
        static const bool fuse_pool = getenv("ELD_NO_FUSED_POOL") == nullptr;   // (A/B: the stand-alone pool kernel)
        TRY(pack());
        {
            // conv1_1 (4 -> 32): software-im2col tile straight from the fp32 NCHW frame (first_conv.cuh)
            const double px = (double)U->n * U->H * U->W;
            Scope sc(u, st, "conv1_1", "fprop", 2.0 * px * 32 * 9 * U->cin0, px * (4 * U->cin0 + 64));
            TRY(launch_first_conv(ctx(), x, U->cin0, wf(I_C11), bias(I_C11), U->a1_1, 32, U->n, U->H, U->W, st, sign_of(U->a1_1)));
        }
        if (fuse_pool) { TRY(conv(I_C12, U->a1_1, 32, 0, U->cat9, 64, 32, 0, U->p1, U->pc1)); } else { TRY(conv(I_C12, U->a1_1, 32, 0, U->cat9, 64, 32, 0)); TRY(pool(U->cat9, 64, 32, U->p1, 32, 1)); }      // + pool (Unet.py:51)
        TRY(conv(I_C21, U->p1, 32, 0, U->a2_1, 64, 0, 1));
        if (fuse_pool) { TRY(conv(I_C22, U->a2_1, 64, 0, U->cat8, 128, 64, 1, U->p2, U->pc2)); } else { TRY(conv(I_C22, U->a2_1, 64, 0, U->cat8, 128, 64, 1)); TRY(pool(U->cat8, 128, 64, U->p2, 64, 2)); }     // + pool (Unet.py:55)
        TRY(conv(I_C31, U->p2, 64, 0, U->a3_1, 128, 0, 2));
        if (fuse_pool) { TRY(conv(I_C32, U->a3_1, 128, 0, U->cat7, 256, 128, 2, U->p3, U->pc3)); } else { TRY(conv(I_C32, U->a3_1, 128, 0, U->cat7, 256, 128, 2)); TRY(pool(U->cat7, 256, 128, U->p3, 128, 3)); }   // + pool (Unet.py:59)
        TRY(conv(I_C41, U->p3, 128, 0, U->a4_1, 256, 0, 3));
        if (fuse_pool) { TRY(conv(I_C42, U->a4_1, 256, 0, U->cat6, 512, 256, 3, U->p4, U->pc4)); } else { TRY(conv(I_C42, U->a4_1, 256, 0, U->cat6, 512, 256, 3)); TRY(pool(U->cat6, 512, 256, U->p4, 256, 4)); }   // + pool (Unet.py:63)
        TRY(conv(I_C51, U->p4, 256, 0, U->a5_1, 512, 0, 4));
        TRY(conv(I_C52, U->a5_1, 512, 0, U->a5_2, 512, 0, 4));
        TRY(deconv(I_UP6, U->a5_2, 512, U->cat6, 512, 4));
        TRY(conv(I_C61, U->cat6, 512, 0, U->a6_1, 256, 0, 3)); TRY(conv(I_C62, U->a6_1, 256, 0, U->a6_2, 256, 0, 3));
        TRY(deconv(I_UP7, U->a6_2, 256, U->cat7, 256, 3));
        TRY(conv(I_C71, U->cat7, 256, 0, U->a7_1, 128, 0, 2)); TRY(conv(I_C72, U->a7_1, 128, 0, U->a7_2, 128, 0, 2));
        TRY(deconv(I_UP8, U->a7_2, 128, U->cat8, 128, 2));
        TRY(conv(I_C81, U->cat8, 128, 0, U->a8_1, 64, 0, 1));  TRY(conv(I_C82, U->a8_1, 64, 0, U->a8_2, 64, 0, 1));
        TRY(deconv(I_UP9, U->a8_2, 64, U->cat9, 64, 1));
        TRY(conv(I_C91, U->cat9, 64, 0, U->a9_1, 32, 0, 0));   TRY(conv(I_C92, U->a9_1, 32, 0, U->a9_2, 32, 0, 0));
        return ELD_OK;
    }

    int backward(const float* x, float* g) const
    {
        eld_unet* U = u;
        // the fused pools leave their codes; ELD_POOL_BWD_FROM_ACT=1 keeps the backward that re-reads the activations (A/B)
        static const bool coded = getenv("ELD_NO_FUSED_POOL") == nullptr && getenv("ELD_POOL_BWD_FROM_ACT") == nullptr;
        TRY(conv_wgrad(I_C92, U->a9_1, 32, 0, U->dz9_2, g, 0));
        TRY(conv_dgrad(I_C92, U->dz9_2, U->dz9_1, 32, 0, U->a9_1, 32, 0, 0));
        TRY(conv_wgrad(I_C91, U->cat9, 64, 0, U->dz9_1, g, 0));
        TRY(conv_dgrad(I_C91, U->dz9_1, U->dcat9, 32, 0, nullptr, 0, 0, 0, skip_half(U->dcat9, 0, 32)));
        TRY(deconv_wgrad(I_UP9, U->a8_2, U->dcat9, 32, g, 1));
        TRY(deconv_dgrad(I_UP9, U->dcat9, 32, U->dz8_2, U->a8_2, 1));
        TRY(conv_wgrad(I_C82, U->a8_1, 64, 0, U->dz8_2, g, 1));
        TRY(conv_dgrad(I_C82, U->dz8_2, U->dz8_1, 64, 0, U->a8_1, 64, 0, 1));
        TRY(conv_wgrad(I_C81, U->cat8, 128, 0, U->dz8_1, g, 1));
        TRY(conv_dgrad(I_C81, U->dz8_1, U->dcat8, 64, 0, nullptr, 0, 0, 1, skip_half(U->dcat8, 1, 64)));
        TRY(deconv_wgrad(I_UP8, U->a7_2, U->dcat8, 64, g, 2));
        TRY(deconv_dgrad(I_UP8, U->dcat8, 64, U->dz7_2, U->a7_2, 2));
        TRY(conv_wgrad(I_C72, U->a7_1, 128, 0, U->dz7_2, g, 2));
        TRY(conv_dgrad(I_C72, U->dz7_2, U->dz7_1, 128, 0, U->a7_1, 128, 0, 2));
        TRY(conv_wgrad(I_C71, U->cat7, 256, 0, U->dz7_1, g, 2));
        TRY(conv_dgrad(I_C71, U->dz7_1, U->dcat7, 128, 0, nullptr, 0, 0, 2, skip_half(U->dcat7, 2, 128)));
        TRY(deconv_wgrad(I_UP7, U->a6_2, U->dcat7, 128, g, 3));
        TRY(deconv_dgrad(I_UP7, U->dcat7, 128, U->dz6_2, U->a6_2, 3));
        TRY(conv_wgrad(I_C62, U->a6_1, 256, 0, U->dz6_2, g, 3));
        TRY(conv_dgrad(I_C62, U->dz6_2, U->dz6_1, 256, 0, U->a6_1, 256, 0, 3));
        TRY(conv_wgrad(I_C61, U->cat6, 512, 0, U->dz6_1, g, 3));
        TRY(conv_dgrad(I_C61, U->dz6_1, U->dcat6, 256, 0, nullptr, 0, 0, 3, skip_half(U->dcat6, 3, 256)));
        TRY(deconv_wgrad(I_UP6, U->a5_2, U->dcat6, 256, g, 4));
        TRY(finish_bucket(0, g));
        TRY(deconv_dgrad(I_UP6, U->dcat6, 256, U->dz5_2, U->a5_2, 4));
        // bottleneck + encoder
        TRY(conv_wgrad(I_C52, U->a5_1, 512, 0, U->dz5_2, g, 4));
        TRY(conv_dgrad(I_C52, U->dz5_2, U->dz5_1, 512, 0, U->a5_1, 512, 0, 4));
        TRY(conv_wgrad(I_C51, U->p4, 256, 0, U->dz5_1, g, 4));
        TRY(finish_bucket(1, g));
        TRY(conv_dgrad(I_C51, U->dz5_1, U->dp4, 256, 0, nullptr, 0, 0, 4));
        TRY(pool_bwd(U->cat6, 512, 256, skip_half(U->dcat6, 3, 256), U->dp4, U->dz4_2, 256, 4, coded ? U->pc4 : nullptr));
        TRY(conv_wgrad(I_C42, U->a4_1, 256, 0, U->dz4_2, g, 3));
        TRY(conv_dgrad(I_C42, U->dz4_2, U->dz4_1, 256, 0, U->a4_1, 256, 0, 3));
        TRY(conv_wgrad(I_C41, U->p3, 128, 0, U->dz4_1, g, 3));
        TRY(conv_dgrad(I_C41, U->dz4_1, U->dp3, 128, 0, nullptr, 0, 0, 3));
        TRY(pool_bwd(U->cat7, 256, 128, skip_half(U->dcat7, 2, 128), U->dp3, U->dz3_2, 128, 3, coded ? U->pc3 : nullptr));
        TRY(conv_wgrad(I_C32, U->a3_1, 128, 0, U->dz3_2, g, 2));
        TRY(conv_dgrad(I_C32, U->dz3_2, U->dz3_1, 128, 0, U->a3_1, 128, 0, 2));
        TRY(conv_wgrad(I_C31, U->p2, 64, 0, U->dz3_1, g, 2));
        TRY(conv_dgrad(I_C31, U->dz3_1, U->dp2, 64, 0, nullptr, 0, 0, 2));
        TRY(pool_bwd(U->cat8, 128, 64, skip_half(U->dcat8, 1, 64), U->dp2, U->dz2_2, 64, 2, coded ? U->pc2 : nullptr));
        TRY(conv_wgrad(I_C22, U->a2_1, 64, 0, U->dz2_2, g, 1));
        TRY(conv_dgrad(I_C22, U->dz2_2, U->dz2_1, 64, 0, U->a2_1, 64, 0, 1));
        TRY(conv_wgrad(I_C21, U->p1, 32, 0, U->dz2_1, g, 1));
        TRY(finish_bucket(2, g));
        TRY(conv_dgrad(I_C21, U->dz2_1, U->dp1, 32, 0, nullptr, 0, 0, 1));
        TRY(pool_bwd(U->cat9, 64, 32, skip_half(U->dcat9, 0, 32), U->dp1, U->dz1_2, 32, 1, coded ? U->pc1 : nullptr));
        TRY(conv_wgrad(I_C12, U->a1_1, 32, 0, U->dz1_2, g, 0));
        TRY(conv_dgrad(I_C12, U->dz1_2, U->dz1_1, 32, 0, U->a1_1, 32, 0, 0));
        {
            const double px = (double)U->n * U->H * U->W;
            Scope sc(u, st, "conv1_1", "wgrad", 2.0 * px * 32 * 9 * U->cin0, px * (4 * U->cin0 + 64));
            TRY(launch_first_conv_wgrad(ctx(), x, U->cin0, U->dz1_1, 32, g + U->L[I_C11].w_off, g + U->L[I_C11].b_off, U->n, U->H, U->W, st));
        }
        return finish_bucket(3, g);
    }
};

}  // namespace

extern "C" int eld_unet_forward(eld_unet* u, const float* params, const float* x, float* out, void* stream)
{
    ELD_REQUIRE(u && params && x && out, "eld_unet_forward: NULL argument");
    ELD_CHECK_CUDA(cudaSetDevice(u->ctx->device));
    Runner r{ u, params, static_cast<cudaStream_t>(stream) };
    TRY(r.forward(x));
    const double hpx = (double)u->n * u->H * u->W;
    Scope sc(u, r.st, "conv10_1", "fprop", 2.0 * hpx * 128, hpx * (64 + 16));
    return launch_head(u->ctx, u->a9_2, params + u->L[I_C10].w_off, params + u->L[I_C10].b_off, out, nullptr, nullptr,
                       nullptr, nullptr, nullptr, u->n, (size_t)u->H * u->W, u->cout_last, 0, r.st);
}

extern "C" int eld_unet_train_step(eld_unet* u, const float* params, const float* x, const float* target,
                                   float* out, float* grads, float* loss, void* stream)
{
    ELD_REQUIRE(u && params && x && target && out && grads && loss, "eld_unet_train_step: NULL argument");
    ELD_REQUIRE(u->dz9_2 != nullptr, "eld_unet_train_step: the eld_unet was created with train = 0");
    ELD_CHECK_CUDA(cudaSetDevice(u->ctx->device));
    Runner r{ u, params, static_cast<cudaStream_t>(stream) };
    ELD_CHECK_CUDA(cudaMemsetAsync(grads, 0, u->n_params * sizeof(float), r.st));
    ELD_CHECK_CUDA(cudaMemsetAsync(u->gtmp, 0, u->n_params * sizeof(float), r.st));
    ELD_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), r.st));
    TRY(r.forward(x));
    {
        const double hpx = (double)u->n * u->H * u->W;
        Scope sc(u, r.st, "conv10_1", "fwd+loss+bwd", 6.0 * hpx * 128, hpx * (64 + 16 + 16 + 64));
        TRY(launch_head(u->ctx, u->a9_2, params + u->L[I_C10].w_off, params + u->L[I_C10].b_off, out, target, u->dz9_2,
                        grads + u->L[I_C10].w_off, grads + u->L[I_C10].b_off, loss, u->n, (size_t)u->H * u->W, u->cout_last, u->l2_loss, r.st));
    }
    return r.backward(x, grads);
}

extern "C" int eld_unet_backward(eld_unet* u, const float* params, const float* x, const float* dout, float* grads, void* stream)
{
    ELD_REQUIRE(u && params && x && dout && grads, "eld_unet_backward: NULL argument");
    ELD_REQUIRE(u->dz9_2 != nullptr, "eld_unet_backward: the eld_unet was created with train = 0");
    ELD_CHECK_CUDA(cudaSetDevice(u->ctx->device));
    Runner r{ u, params, static_cast<cudaStream_t>(stream) };
    ELD_CHECK_CUDA(cudaMemsetAsync(grads, 0, u->n_params * sizeof(float), r.st));
    ELD_CHECK_CUDA(cudaMemsetAsync(u->gtmp, 0, u->n_params * sizeof(float), r.st));
    {
        const double hpx = (double)u->n * u->H * u->W;
        Scope sc(u, r.st, "conv10_1", "bwd", 4.0 * hpx * 128, hpx * (64 + 16 + 64));
        // the head re-forms `out` into scratch (dz1_1 is not written before the very end of backward) and back-propagates dout
        TRY(launch_head(u->ctx, u->a9_2, params + u->L[I_C10].w_off, params + u->L[I_C10].b_off,
                        reinterpret_cast<float*>(u->dz1_1), dout, u->dz9_2,
                        grads + u->L[I_C10].w_off, grads + u->L[I_C10].b_off, nullptr, u->n, (size_t)u->H * u->W, u->cout_last, 2, r.st));
    }
    return r.backward(x, grads);
}

extern "C" int eld_adam_step(eld_ctx* ctx, float* params, const float* grads, float* m, float* v, size_t n,
                             float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                             float grad_scale, void* stream)
{
    ELD_REQUIRE(ctx && params && grads && m && v, "eld_adam_step: NULL argument");
    ELD_REQUIRE(step >= 1, "eld_adam_step: step counts from 1");
    ELD_CHECK_CUDA(cudaSetDevice(ctx->device));
    return launch_adam(ctx, params, grads, m, v, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale,
                       static_cast<cudaStream_t>(stream));
}

extern "C" int eld_unet_set_loss(eld_unet* u, int kind)
{
    ELD_REQUIRE(u && (kind == 0 || kind == 1), "eld_unet_set_loss: kind must be 0 (L1) or 1 (MSE)");
    u->l2_loss = kind;
    return ELD_OK;
}

extern "C" int eld_clock_probe(eld_ctx* ctx, float* out_mhz_device, void* stream)
{
    ELD_REQUIRE(ctx && out_mhz_device, "eld_clock_probe: NULL argument");
    ELD_CHECK_CUDA(cudaSetDevice(ctx->device));
    return launch_clock_probe(ctx, out_mhz_device, static_cast<cudaStream_t>(stream));
}

extern "C" int eld_unet_profile(eld_unet* u, int enable)
{
    ELD_REQUIRE(u, "eld_unet_profile: NULL");
    u->profile = enable != 0;
    u->rec_used = 0;
    return ELD_OK;
}

/* Synchronises the device; fills up to `max` records of the launches since eld_unet_profile(u, 1). */
extern "C" int eld_unet_profile_read(eld_unet* u, int max, char* names32, float* ms, double* flops, double* bytes, int* count)
{
    ELD_REQUIRE(u && count, "eld_unet_profile_read: NULL");
    ELD_CHECK_CUDA(cudaDeviceSynchronize());
    int n = (int)u->rec_used < max ? (int)u->rec_used : max;
    for (int i = 0; i < n; ++i) {
        float t = 0.f;
        ELD_CHECK_CUDA(cudaEventElapsedTime(&t, u->recs[i].e0, u->recs[i].e1));
        if (names32) memcpy(names32 + 32 * i, u->recs[i].name, 32);
        if (ms) ms[i] = t;
        if (flops) flops[i] = u->recs[i].flops;
        if (bytes) bytes[i] = u->recs[i].bytes;
    }
    *count = n;
    u->rec_used = 0;
    return ELD_OK;
}
