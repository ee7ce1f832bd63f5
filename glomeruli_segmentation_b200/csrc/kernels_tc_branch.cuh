// tcgen05 / TMEM kernel of the ESP "split-transform-merge" stage (sm_100a): fp16 operands, fp32 accumulation in
// tensor memory, fp32 epilogue.  ESPNET_MODE_F16TC of include/espnet_b200.h.
//
// Reference semantics: DilatedParllelResidualBlockB.forward (Model.py:187-214) and DownSamplerB.forward
// (Model.py:144-160) after the c1 reduce: d_k = Conv3x3(dilation k, pad k)(o1) for k = 1,2,4,8,16, HFF prefix sums,
// concat, residual add before BN, BN(eval, eps 1e-3) folded to scale/shift, PReLU.
//
// Work item = one M = 128 MMA tile = 8 columns x 16 rows of output pixels (TMEM lane l <-> row l/8, column l%8).
//   * A operand: the reduced map o1 in fp16 "chunk-plane" layout [B][kc][H][W][8 ch] (tc_common.cuh).  TMA brings the
//     tile plus the 16-pixel halo of the d = 16 branch (48 rows x 40 columns) into shared memory, two K chunks
//     (= one K = 16 MMA step) per pipeline stage, two stages.  The tensor map presents [W][8 ch] as ONE dimension of
//     32-bit words so that box rows are 640 B contiguous runs.  The zero padding of every conv is TMA's out-of-bounds
//     fill.  Tap (ky,kx) of dilation d is the SAME buffer read through a UMMA descriptor whose start address is moved
//     by ((ky-1)*d*40 + (kx-1)*d) * 16 B: the implicit-GEMM window shift costs no data movement.
//   * B operand: all five branches' weights [tap][kc][br][NOUT][8] fp16, resident in shared memory for the whole
//     persistent CTA (one cp.async.bulk at kernel start; 90 KB at level 3, 22.5 KB at level 2).  The branches are
//     contiguous along N: the centre tap reads the SAME A window for every dilation, so it is ONE MMA of N = 5 NOUT
//     into the five adjacent accumulators instead of five (small-N MMAs cost a fixed ~57 cycles each on this part).
//   * Taps whose whole 8x16 window lies in the zero padding (d = 8, 16 near the map border) CAN be predicated off
//     (ESPNET_TC_NOSKIP=0); measured slower than issuing them, see the macro.
//   * D: five accumulators of NOUT fp32 columns per tile, THREE tiles deep in TMEM, so that the epilogue of tile i
//     runs while TMA and the tensor core work on tiles i+1 and i+2.
//   * MMA order per tile: K step outer, centre tap then (branch, tap) inner -- stage s only needs the K chunks of step s, so the TMA
//     of the next tile's first half overlaps the second half's MMAs.
//   * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane of the converged warp), warp 2 = TMEM allocator,
//     warp 3 = L2 prefetcher of the residual, warps 4..19 = epilogue: TMEM -> registers, HFF sums, residual + BN +
//     PReLU, planar fp32 stores; warp e works on TMEM lanes 32*(e%4).. and on accumulator column group e/4.
#pragma once
#include "kernels_fp32.cuh"
#include "tc_common.cuh"

namespace espnet {

constexpr int kTcTileW = 8;                          // output tile: 8 columns x 16 rows = 128 pixels = MMA M
constexpr int kTcTileH = 16;
constexpr int kTcHalo2 = 16;                         // largest dilation
constexpr int kTcBoxW = kTcTileW + 2 * kTcHalo2;     // 40
constexpr int kTcBoxH = kTcTileH + 2 * kTcHalo2;     // 48
constexpr int kTcPlane2 = kTcBoxH * kTcBoxW * 16;    // bytes of one K chunk plane of the box: 30720
constexpr int kTcStage = 2 * kTcPlane2;              // one pipeline stage = 2 K chunks = one K = 16 MMA step: 61440
constexpr int kTcThreads = 640;   // 4 control warps + 16 epilogue warps
constexpr int kTcAccStages = 3;
#ifndef ESPNET_TC_PREFETCH_LEAD
#define ESPNET_TC_PREFETCH_LEAD 3
#endif
constexpr int kTcPrefetchLead = ESPNET_TC_PREFETCH_LEAD;
#ifndef ESPNET_TC_NOSKIP
// 0: predicate off the taps whose window lies entirely in the zero padding (8 % of the level-3 MMAs).  Measured SLOWER
// (0.976 -> 1.006 ms per 9 level-3 launches): a predicated-off UTCHMMA still costs the issuing thread its descriptor
// arithmetic and the predicate tests lengthen the single-thread issue loop, so the default issues every tap.
#define ESPNET_TC_NOSKIP 1
#endif

struct BranchTcParams {
    const __half* w;        // [9][NKC][5][NOUT][8] fp16 (d1, d2, d4, d8, d16); split variants: see BranchTcCfg
    const float* res;       // [B,C,H,W] residual input or nullptr
    const float *s, *t, *a; // BN scale/shift + PReLU slope (C)
    float* out;             // [B,C,H,W] or nullptr
    const float *s2, *t2, *a2;
    float* out2;            // [B,C2,H,W] or nullptr (second BR straight into the following concat buffer)
    int C2, c2_off;
    int B, H, W;
    int dbg;                // TIMING EXPERIMENTS ONLY (wrong results; honoured only when the library is built with
                            // -DESPNET_TC_TIMING_EXPERIMENTS=1): bit 0 = stream the split weights only for the first tile, bit 1 = load
                            // the A boxes only for the first tile, bit 2 = epilogue without global loads / stores (option "dbg")
};
#ifndef ESPNET_TC_TIMING_EXPERIMENTS
#define ESPNET_TC_TIMING_EXPERIMENTS 0
#endif
__device__ __forceinline__ int tc_dbg(const BranchTcParams& p) { return ESPNET_TC_TIMING_EXPERIMENTS ? p.dbg : 0; }

template <int NKC, int NOUT, bool SPLIT = false>
struct BranchTcCfg {
    static constexpr int KS = NKC / 2;                            // K = 16 steps per tile
    static constexpr int W_BRANCH = 9 * NKC * NOUT * 16;          // bytes of one branch's weights (plain layout)
    static constexpr bool W_STREAM = SPLIT && KS > 1;             // split weights do not fit: streamed in four units per K step
    static constexpr bool MERGE = SPLIT && KS == 1;               // one K step: [tap][2][br][hi NOUT | lo NOUT][8], resident
    static constexpr int NB = MERGE ? 2 * NOUT : NOUT;            // weight rows = accumulator columns per branch
    // streamed split layout: per K step four units [tap][2 chunks][branches of the group][NOUT][8]: W_hi group A (d1, d2),
    // W_hi group B (d4, d8, d16), W_lo group A, W_lo group B (see the kernel comment for why two branch groups)
    static constexpr int GA = 2, GB = 3;
    static constexpr int UA = 9 * 2 * GA * NOUT * 16, UB = 9 * 2 * GB * NOUT * 16;
    static constexpr int OFF_HIA = 0, OFF_HIB = UA, OFF_LOA = UA + UB, OFF_LOB = 2 * UA + UB;
    static constexpr int W_KSTEP = 2 * (UA + UB);                 // bytes of one K step's hi + lo weights
    static constexpr int W_BYTES = SPLIT ? W_KSTEP : 5 * W_BRANCH;
    static constexpr int ACC_COLS = 5 * NB;                       // TMEM columns per tile
    static constexpr int TMEM_COLS = (kTcAccStages * ACC_COLS <= 256) ? 256 : 512;
    static constexpr int EP_BYTES = 2 * 128 * 16;                 // two float4 tables of 128 channels
    static constexpr size_t SMEM = 1024 + 2 * (size_t)kTcStage + (size_t)W_BYTES + EP_BYTES + 256;
    static_assert(kTcAccStages * ACC_COLS <= 512, "TMEM columns");
    static_assert(!MERGE || W_BYTES == 5 * 9 * 2 * 2 * NOUT * 16, "merged layout size");
};

// VAR: 0 = DownSamplerB (no residual, writes out and out2), 1 = ESP block (residual, out), 2 = last ESP block of a level
// (residual, writes only the BR'd copy into the following concat buffer)
//
// SPLIT = true is the fp32-equivalent variant (ESPNET_MODE_FP32 with fp32_impl = tensor cores): both operands are
// 3-term fp16 splits, a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo with a_hi = fp16(a/4), a_lo = fp16(a/4 - a_hi),
// w_hi = fp16(4w), w_lo = fp16(4w - w_hi): 22-bit mantissa products, fp32 accumulation in TMEM.  o1 arrives as two
// chunk-plane tensors (hi crops [0,B), lo crops [B,2B) of one tensor map); per K step the hi half-box goes to stage
// 0 and the lo half-box to stage 1.  Two K steps (level 3): hi + lo weights of all branches (180 KB) do not fit next to the
// boxes, they stream per K step.  The two MMAs that share A_hi -- A_hi x W_hi and A_hi x W_lo -- are issued back to back
// with the A COLLECTOR (tc::umma_f16_keep_a / umma_f16_reuse_a): the second one does not re-read its 4 KB window from shared
// memory, 76 instead of 114 cycles per pair.  That needs W_hi(k) and W_lo(k) at the same time and W_hi(k) again for
// A_lo x W_hi, so a plain two-unit ring would leave no time to fetch W_hi(k+1).  The branches are therefore split in two
// groups, A = (d1, d2) and B = (d4, d8, d16), each with its own W_hi / W_lo unit, and a K step runs four phases
//     P0: A_hi x {W_hi, W_lo}(A)   P1: A_hi x {W_hi, W_lo}(B)   P2: A_lo x W_hi(A)   P3: A_lo x W_hi(B)
// so that every unit is idle for at least one phase before it is needed again: W_lo(A) after P0, W_lo(B) and the A_hi box
// after P1, W_hi(A) after P2 (refilled during P3), W_hi(B) and the A_lo box after P3 (refilled during the next P0).  The
// producer refills in exactly that order.  (41 + 82) MMAs per K step as before, 2 x 41 of them at the pair rate.
// With a single K step (level 2, Cfg::MERGE) TMEM has room for separate hi / lo accumulators: the weights are resident
// as [W_hi | W_lo] rows, A_hi x [W_hi | W_lo] is ONE N = 2 NOUT MMA, A_lo x W_hi a second one, and the epilogue adds the
// two accumulator halves -- 2x the MMAs of the plain variant.
template <int NKC, int NOUT, int CO1, int CO, int VAR, bool SPLIT>
__global__ void __launch_bounds__(kTcThreads, 1) esp_branch_tc_kernel(const __grid_constant__ CUtensorMap tmap, const BranchTcParams p) {
    using Cfg = BranchTcCfg<NKC, NOUT, SPLIT>;
    constexpr int C = CO1 + 4 * CO;
    constexpr int KS = Cfg::KS;
    static_assert(C <= 128 && CO1 <= NOUT && CO <= NOUT && (NOUT == 16 || NOUT == 32), "channel counts");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [A stage 0 | A stage 1 | weights | epilogue params | barriers | tmem slot]
    uint8_t* abuf = smem_raw;
    uint8_t* wbuf = abuf + 2 * kTcStage;
    float4* sep4 = reinterpret_cast<float4*>(wbuf + Cfg::W_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sep4) + Cfg::EP_BYTES);
    uint64_t* a_full = bars + 0;      // [2]
    uint64_t* a_empty = bars + 2;     // [2]
    uint64_t* w_full = bars + 4;      // [4] streamed split weights: hi A, hi B, lo A, lo B (other variants: only [0])
    uint64_t* w_empty = bars + 8;     // [4] (streamed split weights only)
    uint64_t* acc_full = bars + 12;   // [3]
    uint64_t* acc_empty = bars + 15;  // [3]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
    uint32_t* epi_done = tmem_slot + 1;   // tiles finished by epilogue warp 0 (throttle of the L2 prefetcher)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = p.H, W = p.W;
    const int tiles_x = (W + kTcTileW - 1) / kTcTileW, tiles_y = (H + kTcTileH - 1) / kTcTileH;
    const int total_tiles = p.B * tiles_x * tiles_y;
    const int my_tiles = ((int)blockIdx.x < total_tiles) ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    pdl_trigger();
    if (tid == 0) {
        if ((tc::smem_addr(abuf) & 127u) != 0) __trap();   // TMA destination alignment
        tc::mbar_init(a_full + 0, 1); tc::mbar_init(a_full + 1, 1);
        tc::mbar_init(a_empty + 0, 1); tc::mbar_init(a_empty + 1, 1);
        for (int u = 0; u < 4; ++u) { tc::mbar_init(w_full + u, 1); tc::mbar_init(w_empty + u, 1); }
        for (int s = 0; s < kTcAccStages; ++s) { tc::mbar_init(acc_full + s, 1); tc::mbar_init(acc_empty + s, 16); }
        *epi_done = 0;
        tc::mbar_fence_init();
        tc::tma_prefetch_desc(&tmap);
    }
    if (warp == 2) tc::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    // epilogue params as float4 (scale, shift, slope, -): [0,128) own BN + PReLU, [128,256) the second BR
    for (int i = tid; i < C; i += kTcThreads) {
        sep4[i] = make_float4(p.s[i], p.t[i], p.a[i], 0.f);
        if (VAR != 1) sep4[128 + i] = make_float4(p.s2[p.c2_off + i], p.t2[p.c2_off + i], p.a2[p.c2_off + i], 0.f);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();   // everything above touched only weights / barriers / TMEM; the input of the kernel before is read below

    if (warp == 0) {
        // ===== producer =====
        if (lane == 0) {
            if constexpr (!SPLIT) {
                // weights once, then one half-box (2 K chunks) per pipeline step
                tc::mbar_expect_tx(w_full, Cfg::W_BYTES);
                tc::bulk_g2s(wbuf, p.w, Cfg::W_BYTES, w_full);
                int c = 0;   // chunk counter: chunk c of this CTA = (tile c / KS, K step c % KS), stage c & 1
                for (int it = 0; it < my_tiles; ++it) {
                    const int tile = (int)blockIdx.x + it * (int)gridDim.x;
                    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks, ++c) {
                        const int s = c & 1;
                        tc::mbar_wait(a_empty + s, (uint32_t)(((c >> 1) & 1) ^ 1));
                        tc::mbar_expect_tx(a_full + s, kTcStage);
                        tc::tma_load_4d(abuf + s * kTcStage, &tmap, a_full + s, 4 * (tx * kTcTileW - kTcHalo2), ty * kTcTileH - kTcHalo2, 2 * ks, b);
                    }
                }
            } else if constexpr (Cfg::MERGE) {
                // merged weights once; per tile A_hi -> stage 0, then A_lo -> stage 1 (the order the MMA issuer needs them)
                tc::mbar_expect_tx(w_full, Cfg::W_BYTES);
                tc::bulk_g2s(wbuf, p.w, Cfg::W_BYTES, w_full);
                for (int it = 0; it < my_tiles; ++it) {
                    const int tile = (int)blockIdx.x + it * (int)gridDim.x;
                    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
                    const int cx = 4 * (tx * kTcTileW - kTcHalo2), cy = ty * kTcTileH - kTcHalo2;
                    const uint32_t par = (uint32_t)(it & 1);
                    tc::mbar_wait(a_empty + 0, par ^ 1);
                    tc::mbar_expect_tx(a_full + 0, kTcStage);
                    tc::tma_load_4d(abuf, &tmap, a_full + 0, cx, cy, 0, b);
                    tc::mbar_wait(a_empty + 1, par ^ 1);
                    tc::mbar_expect_tx(a_full + 1, kTcStage);
                    tc::tma_load_4d(abuf + kTcStage, &tmap, a_full + 1, cx, cy, 0, b + p.B);
                }
            } else {
                // per K step n, in the order in which the MMA issuer frees the buffers during the previous K step:
                // W_lo(A), W_lo(B), A_hi -> stage 0, W_hi(A), W_hi(B), A_lo -> stage 1
                const uint8_t* wg = reinterpret_cast<const uint8_t*>(p.w);
                const bool dw = tc_dbg(p) & 1, da = tc_dbg(p) & 2;
                int n = 0;
                auto load_w = [&](int u, int off, int bytes, int ks, uint32_t par) {
                    if (dw && n >= KS) return;
                    tc::mbar_wait(w_empty + u, par ^ 1);
                    tc::mbar_expect_tx(w_full + u, bytes);
                    tc::bulk_g2s(wbuf + off, wg + (size_t)ks * Cfg::W_KSTEP + off, bytes, w_full + u);
                };
                for (int it = 0; it < my_tiles; ++it) {
                    const int tile = (int)blockIdx.x + it * (int)gridDim.x;
                    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
                    const int cx = 4 * (tx * kTcTileW - kTcHalo2), cy = ty * kTcTileH - kTcHalo2;
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks, ++n) {
                        const uint32_t par = (uint32_t)(n & 1);
                        load_w(2, Cfg::OFF_LOA, Cfg::UA, ks, par);
                        load_w(3, Cfg::OFF_LOB, Cfg::UB, ks, par);
                        if (!(da && n >= KS)) {
                            tc::mbar_wait(a_empty + 0, par ^ 1);
                            tc::mbar_expect_tx(a_full + 0, kTcStage);
                            tc::tma_load_4d(abuf, &tmap, a_full + 0, cx, cy, 2 * ks, b);
                        }
                        load_w(0, Cfg::OFF_HIA, Cfg::UA, ks, par);
                        load_w(1, Cfg::OFF_HIB, Cfg::UB, ks, par);
                        if (!(da && n >= KS)) {
                            tc::mbar_wait(a_empty + 1, par ^ 1);
                            tc::mbar_expect_tx(a_full + 1, kTcStage);
                            tc::tma_load_4d(abuf + kTcStage, &tmap, a_full + 1, cx, cy, 2 * ks, b + p.B);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp runs the control flow converged, one elected lane issues =====
        {
            constexpr int NB = Cfg::NB;
            // Descriptors as (hi, lo) words: hi = SBO | version is constant per operand, lo = start>>4 | LBO<<12 and
            // every window shift / tap is an ADD on lo in 16 B units.  One thread issues everything, so the per-MMA
            // instruction count IS the issue rate: taps are unrolled with constant offsets.
            constexpr uint32_t a_hi = (uint32_t)((kTcBoxW * 16) >> 4) | (1u << 14);
            constexpr uint32_t b_hi = (uint32_t)(128 >> 4) | (1u << 14);
            const uint32_t a_lo0 = (tc::smem_addr(abuf) >> 4) + (uint32_t)(kTcHalo2 * kTcBoxW + kTcHalo2) + ((uint32_t)(kTcPlane2 >> 4) << 16);
            const uint32_t b_lo0 = (tc::smem_addr(wbuf) >> 4) + ((uint32_t)((5 * NB * 16) >> 4) << 16);   // K chunk stride = 5 NB rows
            // All MMAs of one (K step, operand pair): the centre tap of the five branches as one N = 5 NB MMA (CENTER) or
            // five N = NBR ones, then the 8 outer taps per branch, N = NBR columns at accumulator column br * NB.
            // a_lo_s = A stage, b_lo_s = weights of tap 0 / branch 0 of this K step, b_tap = tap stride (16 B units),
            // fresh = first MMAs of the tile (the centre tap overwrites the accumulators), vm = tap validity bits of
            // this tile (4 per branch: top, bottom, left, right window not entirely in the padding).  Afterwards up to
            // three mbarriers receive the completion of everything issued so far (tcgen05.commit).
            auto issue_step = [&](auto nbr_tag, auto center_tag, uint32_t a_lo_s, uint32_t b_lo_s, uint32_t b_tap, uint32_t d_tile,
                                  bool fresh, uint32_t vm, uint64_t* c0, uint64_t* c1, uint64_t* c2) {
                constexpr int NBR = decltype(nbr_tag)::value;
                constexpr bool CENTER = decltype(center_tag)::value != 0;
                constexpr uint32_t idesc = tc::umma_idesc_f16(NBR);
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    const uint64_t adesc_c = ((uint64_t)a_hi << 32) | (uint64_t)a_lo_s;
                    if constexpr (CENTER) {
                        constexpr uint32_t idesc_c = tc::umma_idesc_f16(5 * NB);
                        tc::umma_f16(d_tile, adesc_c, ((uint64_t)b_hi << 32) | (uint64_t)(b_lo_s + 4u * b_tap), idesc_c, fresh ? 0u : 1u);
                    }
#pragma unroll 1
                    for (int br = 0; br < 5; ++br) {
                        const int d = 1 << br, dp = d * kTcBoxW;
                        const uint32_t b_lo = b_lo_s + (uint32_t)(br * NB);
                        const uint32_t d_tmem = d_tile + (uint32_t)(br * NB);
                        const uint32_t m = vm >> (4 * br);
                        if constexpr (!CENTER)
                            tc::umma_f16(d_tmem, adesc_c, ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + 4u * b_tap), idesc, fresh ? 0u : 1u);
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            if (tap == 4) continue;
                            const int ky = tap / 3 - 1, kx = tap % 3 - 1;           // compile-time after unrolling
                            const uint32_t need = (ky < 0 ? 1u : ky > 0 ? 2u : 0u) | (kx < 0 ? 4u : kx > 0 ? 8u : 0u);
                            const uint64_t adesc = ((uint64_t)a_hi << 32) | (uint64_t)(a_lo_s + (uint32_t)(ky * dp + kx * d));
                            const uint64_t bdesc = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + (uint32_t)tap * b_tap);
                            // predicated off when the window lies entirely in the zero padding
                            tc::umma_f16_acc_if((m & need) == need ? 1u : 0u, d_tmem, adesc, bdesc, idesc);
                        }
                    }
                    if (c0) tc::umma_commit(c0);
                    if (c1) tc::umma_commit(c1);
                    if (c2) tc::umma_commit(c2);
                }
                __syncwarp();
            };
            // tap validity of a tile: bit 4 br + {0: ky = 0, 1: ky = 2, 2: kx = 0, 3: kx = 2}
            auto tap_mask = [&](int it) {
                const int tile = (int)blockIdx.x + it * (int)gridDim.x;
                const int x0 = (tile % tiles_x) * kTcTileW, y0 = ((tile / tiles_x) % tiles_y) * kTcTileH;
                uint32_t vm = 0;
#pragma unroll
                for (int br = 0; br < 5; ++br) {
                    const int d = 1 << br;
                    vm |= (uint32_t)(y0 - d + kTcTileH - 1 >= 0) << (4 * br);
                    vm |= (uint32_t)(y0 + d < H) << (4 * br + 1);
                    vm |= (uint32_t)(x0 - d + kTcTileW - 1 >= 0) << (4 * br + 2);
                    vm |= (uint32_t)(x0 + d < W) << (4 * br + 3);
                }
#if ESPNET_TC_NOSKIP
                vm = 0xFFFFFu;
#endif
                return vm;
            };
            if constexpr (!SPLIT) {
                tc::mbar_wait(w_full, 0);
                int c = 0;
                for (int it = 0; it < my_tiles; ++it) {
                    const int as = it % kTcAccStages;
                    const uint32_t vm = tap_mask(it);
                    tc::mbar_wait(acc_empty + as, (uint32_t)(((it / kTcAccStages) & 1) ^ 1));
                    const uint32_t d_tile = tmem_base + (uint32_t)(as * Cfg::ACC_COLS);
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks, ++c) {
                        const int s = c & 1;
                        tc::mbar_wait(a_full + s, (uint32_t)((c >> 1) & 1));
                        __syncwarp();
                        // the half-box is reusable once these MMAs have read it; after the last K step the tile is complete
                        issue_step(IntTag<NOUT>(), IntTag<1>(), a_lo0 + (uint32_t)(s * (kTcStage >> 4)), b_lo0 + (uint32_t)(2 * ks * 5 * NOUT),
                                   (uint32_t)(NKC * 5 * NOUT), d_tile, ks == 0, vm, a_empty + s, ks == KS - 1 ? acc_full + as : nullptr, nullptr);
                    }
                }
            } else if constexpr (Cfg::MERGE) {
                constexpr uint32_t TAP = 2 * 5 * NB;
                tc::mbar_wait(w_full, 0);
                for (int it = 0; it < my_tiles; ++it) {
                    const int as = it % kTcAccStages;
                    const uint32_t par = (uint32_t)(it & 1);
                    const uint32_t vm = tap_mask(it);
                    tc::mbar_wait(acc_empty + as, (uint32_t)(((it / kTcAccStages) & 1) ^ 1));
                    const uint32_t d_tile = tmem_base + (uint32_t)(as * Cfg::ACC_COLS);
                    // (1) A_hi x [W_hi | W_lo]: N = 2 NOUT per branch, both accumulator halves
                    tc::mbar_wait(a_full + 0, par);
                    __syncwarp();
                    issue_step(IntTag<NB>(), IntTag<1>(), a_lo0, b_lo0, TAP, d_tile, true, vm, a_empty + 0, nullptr, nullptr);
                    // (2) A_lo x W_hi: N = NOUT per branch into the hi halves
                    tc::mbar_wait(a_full + 1, par);
                    __syncwarp();
                    issue_step(IntTag<NOUT>(), IntTag<0>(), a_lo0 + (uint32_t)(kTcStage >> 4), b_lo0, TAP, d_tile, false, vm, a_empty + 1,
                               acc_full + as, nullptr);
                }
            } else {
                // One phase = one branch group against one A half-box.  PAIR: A_hi x W_hi keeps A in the collector, A_hi x W_lo
                // reuses it; otherwise single MMAs (A_lo x W_hi).  The centre tap of the group's branches is one N = GN * NOUT
                // MMA (the group's branches are adjacent rows in the unit and adjacent accumulator columns).
                auto issue_group = [&](auto gn_tag, auto pair_tag, uint32_t a_lo_s, uint32_t bh, uint32_t bl, int br0, uint32_t d_tile, bool fresh,
                                       uint64_t* c0, uint64_t* c1, uint64_t* c2) {
                    constexpr int GN = decltype(gn_tag)::value;
                    constexpr bool PAIR = decltype(pair_tag)::value != 0;
                    constexpr uint32_t idesc = tc::umma_idesc_f16(NOUT), idesc_c = tc::umma_idesc_f16(GN * NOUT);
                    constexpr uint32_t TAP = 2 * GN * NOUT;                      // tap stride inside a unit (16 B units)
                    constexpr uint32_t lbo = ((uint32_t)((GN * NOUT * 16) >> 4) << 16);   // K chunk stride = GN * NOUT rows
                    tc::tc_fence_after();
                    if (tc::elect_one()) {
                        const uint64_t adesc_c = ((uint64_t)a_hi << 32) | (uint64_t)a_lo_s;
                        const uint32_t d_g = d_tile + (uint32_t)(br0 * NOUT);
                        {
                            const uint64_t bdesc = ((uint64_t)b_hi << 32) | (uint64_t)(bh + lbo + 4u * TAP);
                            if constexpr (PAIR) {
                                tc::umma_f16_keep_a(d_g, adesc_c, bdesc, idesc_c, fresh ? 0u : 1u);
                                tc::umma_f16_reuse_a(d_g, adesc_c, ((uint64_t)b_hi << 32) | (uint64_t)(bl + lbo + 4u * TAP), idesc_c);
                            } else {
                                tc::umma_f16(d_g, adesc_c, bdesc, idesc_c, fresh ? 0u : 1u);
                            }
                        }
#pragma unroll 1
                        for (int g = 0; g < GN; ++g) {
                            const int d = 1 << (br0 + g), dp = d * kTcBoxW;
                            const uint32_t d_tmem = d_g + (uint32_t)(g * NOUT);
                            const uint32_t bh_g = bh + lbo + (uint32_t)(g * NOUT), bl_g = bl + lbo + (uint32_t)(g * NOUT);
#pragma unroll
                            for (int tap = 0; tap < 9; ++tap) {
                                if (tap == 4) continue;
                                const int ky = tap / 3 - 1, kx = tap % 3 - 1;       // compile-time after unrolling
                                const uint64_t adesc = ((uint64_t)a_hi << 32) | (uint64_t)(a_lo_s + (uint32_t)(ky * dp + kx * d));
                                const uint64_t bdesc = ((uint64_t)b_hi << 32) | (uint64_t)(bh_g + (uint32_t)tap * TAP);
                                if constexpr (PAIR) {
                                    tc::umma_f16_keep_a(d_tmem, adesc, bdesc, idesc, 1u);
                                    tc::umma_f16_reuse_a(d_tmem, adesc, ((uint64_t)b_hi << 32) | (uint64_t)(bl_g + (uint32_t)tap * TAP), idesc);
                                } else {
                                    tc::umma_f16(d_tmem, adesc, bdesc, idesc, 1u);
                                }
                            }
                        }
                        if (c0) tc::umma_commit(c0);
                        if (c1) tc::umma_commit(c1);
                        if (c2) tc::umma_commit(c2);
                    }
                    __syncwarp();
                };
                const uint32_t a_st0 = a_lo0, a_st1 = a_lo0 + (uint32_t)(kTcStage >> 4);
                const uint32_t w0 = tc::smem_addr(wbuf) >> 4;
                const uint32_t w_hia = w0 + (uint32_t)(Cfg::OFF_HIA >> 4), w_hib = w0 + (uint32_t)(Cfg::OFF_HIB >> 4);
                const uint32_t w_loa = w0 + (uint32_t)(Cfg::OFF_LOA >> 4), w_lob = w0 + (uint32_t)(Cfg::OFF_LOB >> 4);
                int n = 0;
                for (int it = 0; it < my_tiles; ++it) {
                    const int as = it % kTcAccStages;
                    tc::mbar_wait(acc_empty + as, (uint32_t)(((it / kTcAccStages) & 1) ^ 1));
                    const uint32_t d_tile = tmem_base + (uint32_t)(as * Cfg::ACC_COLS);
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks, ++n) {
                        const uint32_t par = (uint32_t)(n & 1);
                        const bool ww = !((tc_dbg(p) & 1) && n >= KS), wa = !((tc_dbg(p) & 2) && n >= KS);
                        // P0: A_hi x {W_hi, W_lo}, group A
                        if (ww) tc::mbar_wait(w_full + 2, par);
                        if (wa) tc::mbar_wait(a_full + 0, par);
                        if (ww) tc::mbar_wait(w_full + 0, par);
                        __syncwarp();
                        issue_group(IntTag<Cfg::GA>(), IntTag<1>(), a_st0, w_hia, w_loa, 0, d_tile, ks == 0, w_empty + 2, nullptr, nullptr);
                        // P1: A_hi x {W_hi, W_lo}, group B
                        if (ww) tc::mbar_wait(w_full + 3, par);
                        if (ww) tc::mbar_wait(w_full + 1, par);
                        __syncwarp();
                        issue_group(IntTag<Cfg::GB>(), IntTag<1>(), a_st0, w_hib, w_lob, Cfg::GA, d_tile, ks == 0, w_empty + 3, a_empty + 0, nullptr);
                        // P2: A_lo x W_hi, group A
                        if (wa) tc::mbar_wait(a_full + 1, par);
                        __syncwarp();
                        issue_group(IntTag<Cfg::GA>(), IntTag<0>(), a_st1, w_hia, 0u, 0, d_tile, false, w_empty + 0, nullptr, nullptr);
                        // P3: A_lo x W_hi, group B
                        issue_group(IntTag<Cfg::GB>(), IntTag<0>(), a_st1, w_hib, 0u, Cfg::GA, d_tile, false, w_empty + 1, a_empty + 1,
                                    ks == KS - 1 ? acc_full + as : nullptr);
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ===== L2 prefetcher for the residual input: the epilogue's loads are the latency-bound part of this kernel
        // (16 warps x 5*GW loads in flight per SM), so this otherwise idle warp pulls the residual tile of the tile
        // kTcPrefetchLead ahead into L2: 128 channels x 16 rows x 32 B segments = 64 warp-wide prefetches per tile.
        if (VAR != 0) {
            const size_t plane = (size_t)H * W;
            auto prefetch_tile = [&](int it) {
                const int tile = (int)blockIdx.x + it * (int)gridDim.x;
                const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
                const int r = lane & 15, y = ty * kTcTileH + r, x = tx * kTcTileW;
                if (y >= H) return;
                const float* base = p.res + (size_t)b * C * plane + (size_t)y * W + x;
#pragma unroll 4
                for (int c = lane >> 4; c < C; c += 2)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)c * plane));
            };
            // Throttle on a monotonic progress counter written by the epilogue (an mbarrier can NOT be observed by a
            // thread that may lag more than one phase behind: it would wait for a phase that never comes).
            for (int it = 0; it < my_tiles; ++it) {
                while (*reinterpret_cast<volatile uint32_t*>(epi_done) + (uint32_t)kTcPrefetchLead <= (uint32_t)it) __nanosleep(200);
                prefetch_tile(it);
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: 16 warps; thread <-> pixel (TMEM lane), so a warp touches 4 rows x 32 B per channel plane.
        // Warp e works on TMEM lanes 32*(e%4).. and on accumulator column group G = e/4 of every tile: columns
        // [GW*G, GW*G + GW) of all five branches (GW = NOUT/4), i.e. up to 5*GW concat channels.  G is dispatched to a
        // COMPILE-TIME constant so that which channels exist is known statically and the group is straight-line code.
        // Per tile: all residual loads first (the epilogue is the HBM side of this kernel: 16 warps x 5*GW loads in
        // flight per SM), then the accumulators: per branch one TMEM load, the HFF prefix sum add1..add4
        // (Model.py:152-155,203-206), residual add BEFORE BN (Model.py:211-212), folded BN + PReLU, optional second BR.
        constexpr bool HAS_RES = VAR != 0, HAS_OUT = VAR != 2, HAS_OUT2 = VAR != 1;
        constexpr int GW = NOUT / 4;
        const int e = warp - 4, q = e & 3, gsel = e >> 2;
        const int row = 4 * q + (lane >> 3), col = lane & 7;
        const size_t plane = (size_t)H * W;
        const uint32_t plane_b = (uint32_t)(plane * sizeof(float));   // host guarantees C * plane * 4 < 2^32
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = (int)blockIdx.x + it * (int)gridDim.x;
            const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
            const int y = ty * kTcTileH + row, x = tx * kTcTileW + col;
            const bool valid = (y < H) && (x < W) && !(tc_dbg(p) & 4);
            const size_t pix = valid ? (size_t)y * W + x : 0;
            // byte pointers to channel 0 of this pixel; a channel is "+ ch * plane_b" = one IMAD.WIDE.U32
            const char* res_b = HAS_RES ? reinterpret_cast<const char*>(p.res + (size_t)b * C * plane + pix) : nullptr;
            char* out_b = HAS_OUT ? reinterpret_cast<char*>(p.out + (size_t)b * C * plane + pix) : nullptr;
            char* out2_b = HAS_OUT2 ? reinterpret_cast<char*>(p.out2 + ((size_t)b * p.C2 + p.c2_off) * plane + pix) : nullptr;
            const int as = it % kTcAccStages;
            const uint32_t t0 = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(as * Cfg::ACC_COLS);
            auto group = [&](auto gtag) {
                constexpr int G = decltype(gtag)::value;
                float rv[5][GW];
#pragma unroll
                for (int br = 0; br < 5; ++br)
#pragma unroll
                    for (int jj = 0; jj < GW; ++jj) rv[br][jj] = 0.f;
                if (HAS_RES && valid) {
#pragma unroll
                    for (int br = 0; br < 5; ++br) {
                        const int ch0 = br == 0 ? 0 : CO1 + (br - 1) * CO;
                        const int cnt = br == 0 ? CO1 : CO;
#pragma unroll
                        for (int jj = 0; jj < GW; ++jj)
                            if (GW * G + jj < cnt)
                                rv[br][jj] = __ldg(reinterpret_cast<const float*>(res_b + (uint64_t)plane_b * (uint32_t)(ch0 + GW * G + jj)));
                    }
                }
                tc::mbar_wait(acc_full + as, (uint32_t)((it / kTcAccStages) & 1));
                tc::tc_fence_after();
                float run[GW];
#pragma unroll
                for (int br = 0; br < 5; ++br) {
                    const int ch0 = br == 0 ? 0 : CO1 + (br - 1) * CO;
                    const int cnt = br == 0 ? CO1 : CO;
                    uint32_t r[GW], r2[GW];
                    __syncwarp();    // tcgen05.ld is .sync.aligned: the lanes diverge on `valid` around the stores
                    if constexpr (GW == 8) tc::tmem_ld8_nowait(t0 + (uint32_t)(br * Cfg::NB + GW * G), r);
                    else tc::tmem_ld4_nowait(t0 + (uint32_t)(br * Cfg::NB + GW * G), r);
                    if constexpr (Cfg::MERGE) {      // lo-weight half of the accumulator pair
                        if constexpr (GW == 8) tc::tmem_ld8_nowait(t0 + (uint32_t)(br * Cfg::NB + NOUT + GW * G), r2);
                        else tc::tmem_ld4_nowait(t0 + (uint32_t)(br * Cfg::NB + NOUT + GW * G), r2);
                    }
                    tc::tmem_ld_wait();
                    float o[GW], o2[GW];
#pragma unroll
                    for (int jj = 0; jj < GW; ++jj) {
                        const float d = Cfg::MERGE ? __uint_as_float(r[jj]) + __uint_as_float(r2[jj]) : __uint_as_float(r[jj]);
                        run[jj] = br <= 1 ? d : run[jj] + d;
                        o[jj] = 0.f; o2[jj] = 0.f;
                        if (GW * G + jj < cnt) {
                            const float4 q1 = sep4[ch0 + GW * G + jj];
                            o[jj] = bn_prelu(run[jj] + rv[br][jj], q1.x, q1.y, q1.z);
                            if (HAS_OUT2) {
                                const float4 q2 = sep4[128 + ch0 + GW * G + jj];
                                o2[jj] = bn_prelu(o[jj], q2.x, q2.y, q2.z);
                            }
                        }
                    }
                    if (valid) {
#pragma unroll
                        for (int jj = 0; jj < GW; ++jj) {
                            if (GW * G + jj >= cnt) continue;
                            if (HAS_OUT) *reinterpret_cast<float*>(out_b + (uint64_t)plane_b * (uint32_t)(ch0 + GW * G + jj)) = o[jj];
                            if (HAS_OUT2) *reinterpret_cast<float*>(out2_b + (uint64_t)plane_b * (uint32_t)(ch0 + GW * G + jj)) = o2[jj];
                        }
                    }
                }
            };
            if (gsel == 0) group(IntTag<0>());
            else if (gsel == 1) group(IntTag<1>());
            else if (gsel == 2) group(IntTag<2>());
            else group(IntTag<3>());
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                tc::mbar_arrive(acc_empty + as);
                if (e == 0) *reinterpret_cast<volatile uint32_t*>(epi_done) = (uint32_t)(it + 1);
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace espnet
