// wgrad_conv.cuh - conv3x3 weight gradient, second generation ("full halo").
//
//   dW[co][ci][kh][kw] += sum over pixels  dZ[p][co] * X[p + (kh-1, kw-1)][ci]
//
// GEMM with pixels as the reduction dimension; both operands are MN-major UMMA operands taken straight
// from NHWC tensors (see wgrad_umma.cuh for the first generation, still used for the deconvs).  New here:
//   * the pixel chunk is an 8x8 patch; X is loaded ONCE per chunk as a {ch,10,10} halo box and every filter
//     tap is the same smem tile read from a shifted start row (the UMMA swizzle is a function of absolute
//     smem address bits - measured - so any 16-byte aligned row start and any 8-row-group stride work);
//   * a CTA keeps up to G accumulators (M tiles) in TMEM, so dZ is loaded once per chunk for all of them;
//   * the MMA issuer's per-instruction work is a handful of adds (descriptor halves precomputed).
// TMA requests per pixel drop 3-6x against the first generation.
#pragma once
#include "umma.cuh"
#include "unet_prims.h"
#include <cuda_bf16.h>

namespace eld {

struct Wgrad2Params {
    int n_img, H, W;
    int chunks_x, chunks_y;     // W/8, H/8
    int cin, cout, p_c0, q_c0;
    int box_ch, q_box_ch;       // 32 or 64
    int n_tile, n_tiles;
    int kind;                   // 0: cin == 32 (M tile = filter row kh, blocks = kw 0..3), 1: cin == 64 (M tile = tap pair),
                                // 2: cin >= 128 (M tile = one tap of one 128-channel pair)
    int G, groups, cps;         // M tiles per CTA, CTA groups per channel pair, channel pairs
    int m_tiles;                // M tiles per channel pair (3, 5 or 9)
    int p_boxes, q_boxes;
    int ksplit, stages, tmem_cols;
    int cps_stage;              // 8x8-pixel chunks per pipeline stage (1..4)
    float* dw;
    int out_tco;                // 0: dw is OIHW [co][ci][3][3], scalar red.add; 1: dw is [tap][ci][co], 16-byte vector red.add
    long long* prof;            // PROF instantiation only
    int dbg;                    // experiments only (ELD_CONV_DBG): 8 = skip the TMA loads
    float* db;                  // optional bias gradient: db[co] += sum over pixels of dZ (column sums of the Q tiles,
                                // computed by the otherwise idle epilogue warps of the cp == 0, grp == 0 CTAs)
};

constexpr int kWg2Threads = 192;
constexpr int kWg2MaxG = 8;

__device__ __forceinline__ int tap_row_offset(int t) { return (t / 3) * 10 + (t % 3); }

struct Wg2Issue {
    uint64_t *full, *empty, *acc_full;
    int nchunks, stages, cps;
    uint32_t smem_base, stage_bytes, chunk_bytes, p_bytes, a_hi, b_hi, b_lo_c, idesc;
};

// How the G accumulators (M tiles) of a CTA map to start offsets inside the X halo box.  Everything that does not
// depend on the pipeline stage is a compile-time constant, so one tcgen05.mma costs the issuing thread two
// uniform-datapath adds with immediates (measured: descriptors assembled from vector registers cost ~35 cycles per
// MMA in R2UR transfers - as much as an N = 32 MMA takes on the tensor pipe).
enum { WV_K0 = 0,       // cin == 32: M tile g = filter row g (blocks kw 0..3), G = 3
       WV_K1 = 1,       // cin == 64: M tile = tap pair (2 mt, 2 mt + 1), mt = MT0 + g
       WV_K2_ROW = 2,   // cin >= 128: taps mt0 .. mt0 + G - 1 inside ONE filter row (mt0 % 3 == 0): +1 pixel row per g
       WV_K2_PAIR = 3 };// cin >= 128, G <= 2: second tap `delta` descriptor units after the first (run-time)

template <int V, int MT0, int RBP>
__host__ __device__ constexpr uint32_t wg2_a_const(int g)
{
    if (V == WV_K0) return ((uint32_t)(RBP >> 4) << 16) + (uint32_t)(((MT0 + g) * 10 * RBP) >> 4);
    if (V == WV_K1) {
        const int t0 = 2 * (MT0 + g), t1 = t0 + 1;
        const int off = (t0 / 3) * 10 + (t0 % 3);
        const int lbo = (t1 <= 8 ? ((t1 / 3) * 10 + (t1 % 3)) - off : 1) * RBP;
        return ((uint32_t)(lbo >> 4) << 16) + (uint32_t)((off * RBP) >> 4);
    }
    return (uint32_t)((g * RBP) >> 4);      // WV_K2_ROW: relative to the CTA's first tap
}

// The MMA issuer's loop over pipeline stages (W.cps chunks each).  The readiness of the NEXT stage is probed
// (mbarrier.try_wait, ~100 cycles round trip) before this stage's MMAs are issued, so the probe hides behind the issue work.
template <int V, int MT0, int G, int NT, int RBP, int RBQ, bool PROF>
__device__ __forceinline__ void wg2_issue_loop(const Wg2Issue& W, uint32_t a_cta, uint32_t delta, uint32_t tmem_base,
                                               long long& pw0, long long& pw_issue, long long& pw_commit)
{
    constexpr uint32_t a_kstep = (uint32_t)(20 * RBP) >> 4;      // 16 pixels = 2 patch rows of the 10-wide halo box
    constexpr uint32_t b_kstep = (uint32_t)(16 * RBQ) >> 4;      // 2 rows of the dense 8-wide box
    constexpr uint32_t idesc = ptx::make_idesc_bf16(128, (uint32_t)NT, 1, 1);   // both operands MN-major
    uint32_t st_addr = W.smem_base;
    int s = 0;
    uint32_t ph = 0;
    bool ready = W.nchunks > 0 && ptx::mbar_try_wait(&W.full[0], 0);
    for (int i = 0; i < W.nchunks; i += W.cps) {
        const int n_in = min(W.cps, W.nchunks - i);
        if (!ready) {
            if (PROF) { const long long t_ = clock64(); ptx::mbar_wait(&W.full[s], ph); pw0 += clock64() - t_; }
            else ptx::mbar_wait(&W.full[s], ph);
        }
        ptx::tc_fence_after();
        int s1 = s + 1;
        uint32_t ph1 = ph, st1 = st_addr + W.stage_bytes;
        if (s1 == W.stages) { s1 = 0; ph1 ^= 1u; st1 = W.smem_base; }
        ready = (i + W.cps < W.nchunks) && ptx::mbar_try_wait(&W.full[s1], ph1);
        if (ptx::elect_one()) {
            long long ta_ = 0, tb_ = 0;
            if (PROF) ta_ = clock64();
            uint32_t a_st = a_cta + ((st_addr & 0x3FFFFu) >> 4);
            uint32_t b_lo0 = W.b_lo_c + (((st_addr + W.p_bytes) & 0x3FFFFu) >> 4);
            uint32_t flag0 = i != 0 ? 1u : 0u;
            for (int c = 0; c < n_in; ++c) {
                // g outermost (4 consecutive K steps per accumulator): measured faster than k-outermost on B200
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const uint32_t a_lo0 = a_st + (V == WV_K2_PAIR ? (uint32_t)g * delta : wg2_a_const<V, MT0, RBP>(g));
                    const uint32_t d_tm = tmem_base + (uint32_t)(g * NT);
                    ptx::umma_bf16(d_tm, ((uint64_t)W.a_hi << 32) | a_lo0, ((uint64_t)W.b_hi << 32) | b_lo0, idesc, flag0);
#pragma unroll
                    for (int k = 1; k < 4; ++k)
                        ptx::umma_bf16_lohi(d_tm, a_lo0 + k * a_kstep, W.a_hi, b_lo0 + k * b_kstep, W.b_hi, idesc, true);
                }
                flag0 = 1u;
                a_st += W.chunk_bytes >> 4;
                b_lo0 += W.chunk_bytes >> 4;
            }
            if (PROF) tb_ = clock64();
            ptx::umma_commit(&W.empty[s]);
            if (i + W.cps >= W.nchunks) ptx::umma_commit(W.acc_full);
            if (PROF) { pw_issue += tb_ - ta_; pw_commit += clock64() - tb_; }
        }
        __syncwarp();
        s = s1; ph = ph1; st_addr = st1;
    }
}

// RBP / RBQ: bytes per pixel row of the X (P) and dZ (Q) boxes (64 or 128) - compile-time so that the K-step
// descriptor increments of the MMA issue loop are immediates (uniform-datapath adds instead of R2UR chains).
// NT: the N tile (32 / 64 / 128 / 256); the dZ row bytes follow from it (32-channel boxes only for NT == 32).
template <bool PROF, int RBP, int NT>
__global__ void __launch_bounds__(kWg2Threads, 1)
wgrad_conv_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmQ,
                  const Wgrad2Params p)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    long long pw0 = 0, pt0 = 0, pt1 = 0, pw_issue = 0, pw_commit = 0;
    if (PROF) pt0 = clock64();

    constexpr int RBQ = NT == 32 ? 64 : 128;
    constexpr int rb_p = RBP, rb_q = RBQ;
    const int p_box = (100 * rb_p + 1023) & ~1023;
    const int q_box = 64 * rb_q;
    const int p_bytes = p.p_boxes * p_box;
    // one pipeline stage = p.cps consecutive 8x8-pixel chunks: the barrier round trip (wait -> MMAs -> commit) costs the
    // issuing thread ~450 cycles, more than the 12-20 small-N MMAs of ONE chunk of a thin layer take on the tensor pipe
    const int chunk_bytes = p_bytes + p.q_boxes * q_box;
    const int stage_bytes = p.cps_stage * chunk_bytes;
    const uint32_t tx_bytes = (uint32_t)(p.p_boxes * 100 * rb_p + p.q_boxes * q_box);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* empty = full + p.stages;
    uint64_t* acc_full = empty + p.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int item = blockIdx.x;
    const int ks = item % p.ksplit; item /= p.ksplit;
    const int nt = item % p.n_tiles; item /= p.n_tiles;
    const int grp = item % p.groups;
    const int cp = item / p.groups;
    const int mt0 = grp * p.G;
    const int g_cnt = min(p.G, p.m_tiles - mt0);

    const int total_chunks = p.n_img * p.chunks_y * p.chunks_x;
    const int per = (total_chunks + p.ksplit - 1) / p.ksplit;
    const int ch_begin = ks * per;
    const int nchunks = max(0, min(total_chunks, ch_begin + per) - ch_begin);

    // every CTA of one (n tile, K split) streams the same dZ tiles: each sums its own slice of the columns
    // (one CTA doing all of them was the straggler of the wide layers)
    const bool do_bias = p.db != nullptr;
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmP);
        ptx::prefetch_tmap(&tmQ);
        for (int s = 0; s < p.stages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], do_bias ? 5 : 1); }
        ptx::mbar_init(acc_full, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::grid_dep_wait();       // PDL: the prologue above overlapped the previous kernel's tail
    ptx::grid_dep_launch();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0 && nchunks > 0) {
            const int pc0 = p.p_c0 + cp * 128;                 // kind 2: this channel pair; kinds 0/1: cp == 0
            const int qc0 = p.q_c0 + nt * p.n_tile;
            const int cxy = p.chunks_x * p.chunks_y;
            int img = ch_begin / cxy;
            int rem = ch_begin - img * cxy;
            int cy = rem / p.chunks_x, cx = rem - cy * p.chunks_x;
            int s = 0;
            uint32_t ph = 0;
            uint8_t* sa = smem;
            for (int i = 0; i < nchunks; i += p.cps_stage) {
                const int n_in = min(p.cps_stage, nchunks - i);
                if (PROF) { const long long t_ = clock64(); ptx::mbar_wait(&empty[s], ph ^ 1u); pw0 += clock64() - t_; }
                else ptx::mbar_wait(&empty[s], ph ^ 1u);
                if (p.dbg & 8) ptx::mbar_arrive(&full[s]);        // experiments only: no loads
                else ptx::mbar_arrive_expect_tx(&full[s], (uint32_t)n_in * tx_bytes);
                for (int c = 0; c < n_in; ++c) {
                    const int x0 = cx * 8, y0 = cy * 8;
                    uint8_t* sc = sa + c * chunk_bytes;
                    if (!(p.dbg & 8)) {
                        for (int b = 0; b < p.p_boxes; ++b)
                            ptx::tma_load_5d(sc + b * p_box, &tmP, &full[s], pc0 + b * p.box_ch, x0 - 1, y0 - 1, img, 0);
                        for (int b = 0; b < p.q_boxes; ++b)
                            ptx::tma_load_5d(sc + p_bytes + b * q_box, &tmQ, &full[s], qc0 + b * p.q_box_ch, x0, y0, img, 0);
                    }
                    if (++cx == p.chunks_x) { cx = 0; if (++cy == p.chunks_y) { cy = 0; ++img; } }
                }
                sa += stage_bytes;
                if (++s == p.stages) { s = 0; ph ^= 1u; sa = smem; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t a_layout = p.box_ch == 64 ? ptx::LAYOUT_SW128 : ptx::LAYOUT_SW64;
        const uint32_t b_layout = p.q_box_ch == 64 ? ptx::LAYOUT_SW128 : ptx::LAYOUT_SW64;
        const uint32_t a_hi = (uint32_t)(ptx::make_smem_desc(0, 0, 10u * rb_p, a_layout) >> 32);
        const uint64_t b_desc0 = ptx::make_smem_desc(0, (uint32_t)q_box, 8u * rb_q, b_layout);
        const uint32_t b_hi = (uint32_t)(b_desc0 >> 32);
        const uint32_t b_lo_c = (uint32_t)b_desc0;
        const uint32_t smem_base = ptx::smem_u32(smem);
        const Wg2Issue W{ full, empty, acc_full, nchunks, p.stages, p.cps_stage, smem_base, (uint32_t)stage_bytes, (uint32_t)chunk_bytes, (uint32_t)p_bytes,
                          a_hi, b_hi, b_lo_c, 0u };
#define ELD_WG2_GO(V_, MT0_, G_, A_CTA_, DELTA_) \
        wg2_issue_loop<V_, MT0_, G_, NT, RBP, RBQ, PROF>(W, (A_CTA_), (DELTA_), tmem_base, pw0, pw_issue, pw_commit)
        if constexpr (RBP == 64) {                 // kind 0 (cin == 32): three filter rows, one CTA
            ELD_WG2_GO(WV_K0, 0, 3, 0u, 0u);
        } else if (p.kind == 1) {                  // cin == 64: five tap pairs, split over CTA groups when NT > 64
            if constexpr (NT <= 64) ELD_WG2_GO(WV_K1, 0, 5, 0u, 0u);
            else if constexpr (NT == 128) { if (mt0 == 0) ELD_WG2_GO(WV_K1, 0, 3, 0u, 0u); else ELD_WG2_GO(WV_K1, 3, 2, 0u, 0u); }
            else { if (mt0 == 0) ELD_WG2_GO(WV_K1, 0, 2, 0u, 0u); else if (mt0 == 2) ELD_WG2_GO(WV_K1, 2, 2, 0u, 0u); else ELD_WG2_GO(WV_K1, 4, 1, 0u, 0u); }
        } else {                                   // cin >= 128: one tap of one 128-channel pair per M tile
            const uint32_t a_cta = ((uint32_t)(p_box >> 4) << 16) + (uint32_t)((tap_row_offset(mt0) * rb_p) >> 4);
            if constexpr (NT <= 128) ELD_WG2_GO(WV_K2_ROW, 0, 3, a_cta, 0u);          // taps mt0, mt0+1, mt0+2 (one filter row)
            else {
                const uint32_t delta = (uint32_t)(((tap_row_offset(min(mt0 + 1, 8)) - tap_row_offset(mt0)) * rb_p) >> 4);
                if (g_cnt == 2) ELD_WG2_GO(WV_K2_PAIR, 0, 2, a_cta, delta); else ELD_WG2_GO(WV_K2_PAIR, 0, 1, a_cta, 0u);
            }
        }
#undef ELD_WG2_GO
    } else if (nchunks > 0) {
        const int q = warp & 3;
        const int r = q * 32 + lane;                      // accumulator row
        if (do_bias) {
            // ---- bias gradient while the MMAs run: column sums of the dZ (Q) tile of every chunk ----
            const int et = (warp - 2) * 32 + lane;        // 0..127
            const int words_all = p.n_tile >> 1;          // 32-bit words (channel pairs) per pixel row: 16..128
            const int n_share = p.cps * p.groups, share = cp * p.groups + grp;
            const int wpi = (words_all + n_share - 1) / n_share;
            const int w_begin = share * wpi;
            const int words = max(0, min(words_all, w_begin + wpi) - w_begin);   // this CTA's slice
            const int rgroups = words > 0 ? 128 / words : 1;   // threads sharing a word split the 64 pixel rows
            const bool active = words > 0 && et < words * rgroups;
            const int wd = w_begin + (words > 0 ? et % words : 0), rg = words > 0 ? et / words : 0;
            const int wpb = p.q_box_ch >> 1;              // words per row of one box
            const int bx = wd / wpb, wb = wd - bx * wpb;
            const uint32_t col_byte = (uint32_t)(wb & 3) * 4u;
            const uint32_t chunk16 = (uint32_t)wb >> 2;
            float s0 = 0.f, s1 = 0.f;
            int s = 0;
            uint32_t ph = 0;
            const uint8_t* qb = smem + p_bytes + (size_t)bx * q_box;
            for (int i = 0; i < nchunks; i += p.cps_stage) {
                const int n_in = min(p.cps_stage, nchunks - i);
                ptx::mbar_wait(&full[s], ph);
                if (active) for (int c = 0; c < n_in; ++c) {
                    const uint8_t* base = qb + (size_t)s * stage_bytes + (size_t)c * chunk_bytes;
                    for (int row = rg; row < 64; row += rgroups) {
                        const uint32_t swz = (rb_q == 128) ? (uint32_t)(row & 7) : (uint32_t)((row >> 1) & 3);
                        const uint32_t v = *reinterpret_cast<const uint32_t*>(base + row * rb_q + ((chunk16 ^ swz) << 4) + col_byte);
                        s0 += __uint_as_float(v << 16);
                        s1 += __uint_as_float(v & 0xFFFF0000u);
                    }
                }
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&empty[s]);
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
            if (active) {
                const int co = nt * p.n_tile + 2 * wd;
                atomicAdd(p.db + co, s0);
                atomicAdd(p.db + co + 1, s1);
            }
        }
        // ===================== epilogue: TMEM -> red.add into dW =====================
        ptx::mbar_wait(acc_full, 0);
        ptx::tc_fence_after();
        if (PROF) pt1 = clock64();
        for (int g = 0; g < g_cnt; ++g) {
            const int mt = mt0 + g;
            int tap, ci;
            if (p.kind == 0) { const int kw = r >> 5; tap = kw < 3 ? mt * 3 + kw : 9; ci = r & 31; }
            else if (p.kind == 1) { tap = 2 * mt + (r >> 6); ci = r & 63; }
            else { tap = mt; ci = cp * 128 + r; }
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * p.n_tile);
            for (int c32 = 0; c32 < p.n_tile / 32; ++c32) {
                uint32_t v[32];
                ptx::tmem_ld32(t_addr + c32 * 32, v);
                ptx::tmem_ld_wait();
                if (tap <= 8) {
                    if (p.out_tco) {
                        float* dst = p.dw + ((size_t)tap * p.cin + ci) * p.cout + nt * p.n_tile + c32 * 32;
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
                                         :: "l"(dst + j), "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])),
                                            "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3])) : "memory");
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int co = nt * p.n_tile + c32 * 32 + j;
                            atomicAdd(p.dw + ((size_t)co * p.cin + ci) * 9 + tap, __uint_as_float(v[j]));
                        }
                    }
                }
            }
        }
    }
    if (PROF && lane == 0 && warp <= 2) {
        // slots: 0-1 producer (total, wait empty) | 2-3 MMA (total, wait full) | 4-5 epilogue warp 2 (total, time in the red.add phase)
        long long* o = p.prof + (size_t)blockIdx.x * 8 + warp * 2;
        const long long t = clock64();
        o[0] = t - pt0; o[1] = warp == 2 ? t - pt1 : pw0;
        if (warp == 1) { p.prof[(size_t)blockIdx.x * 8 + 6] = pw_issue; p.prof[(size_t)blockIdx.x * 8 + 7] = pw_commit; }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace eld
