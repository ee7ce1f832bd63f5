// tcgen05 kernel of DownSamplerB's c1 (Model.py:135,145: C(nIn, n, 3, 2) = 3x3, stride 2, pad 1, no bias), F16TC mode.
//
// Implicit GEMM: M = 128 output pixels (8 columns x 16 rows, TMEM lane l <-> row l/8, column l%8), N = NOUT (12 -> 16,
// 25 -> 32), K = 9 taps x CIN (19 -> 32, 131 -> 144 input channels, padded to K = 16 steps).
// A stride-2 tap reads every second input column, so the loader warps split the staged input by COLUMN PARITY while they
// convert it from planar fp32 to the fp16 K-major chunk-plane operand layout (tc_common.cuh):
//     region = input rows 2*oy0-1 .. 2*oy0+31 (33) x columns 2*ox0-1 .. 2*ox0+15 (17), zero outside the map (= pad 1)
//     smem   = [K chunk][parity][33 rows][9 entries][8 ch]        parity 1: columns 2*ox0-1, +1, ..   parity 0: 2*ox0, +2, ..
// Output column x then finds tap kx at parity (kx != 1), entry (x - ox0) + (kx == 2): eight consecutive output pixels are
// eight consecutive 16 B entries = one UMMA core matrix, the next output row is two input rows further (SBO = 2*9*16 B),
// and tap ky is a start-address shift of ky rows.  One pipeline stage = one K = 16 step (16 input channels of the region,
// 19 KB); the weights of all taps and K steps stay resident in shared memory (81 KB at level 3).
//   warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4..11 = loaders, warps 12..15 = epilogue (fp16 chunk-plane o1h).
#pragma once
#include "kernels_fp32.cuh"
#include "tc_common.cuh"
#include "kernels_tc_reduce.cuh"

namespace espnet {

constexpr int kDownThreads = 512;
constexpr int kDownRows = 33, kDownCols = 17, kDownPitch = 9;
constexpr int kDownParBytes = kDownRows * kDownPitch * 16;      // one parity plane of one K chunk: 4752
constexpr int kDownChunkBytes = 2 * kDownParBytes;              // 9504
constexpr int kDownStageBytes = 2 * kDownChunkBytes;            // 19008: two K chunks = one K = 16 step

// SPLIT = fp32-equivalent variant (3-term fp16 operand splits, see kernels_tc_branch.cuh): a stage holds the hi and lo
// copies of the region AND the hi / lo weights of its K step (hi + lo of all K steps, 162 KB at level 3, cannot stay
// resident), streamed by a producer thread with cp.async.bulk; 3 MMAs per tap.
template <int CIN, int NOUT, bool SPLIT = false>
struct DownTcCfg {
    static constexpr int KS = (CIN + 15) / 16;                  // K = 16 steps
    static constexpr int WK = 9 * 2 * NOUT * 16;                // weights of one K step: [tap][2 chunks][NOUT][8] fp16
    static constexpr int W_PART = KS * WK;                      // one (hi or lo) copy of all K steps
    static constexpr int W_RESIDENT = SPLIT ? 0 : W_PART;
    static constexpr int STAGES = SPLIT ? 3 : 4;
    static constexpr int STAGE_BYTES = SPLIT ? 2 * kDownStageBytes + 2 * WK : kDownStageBytes;
    static constexpr size_t SMEM = 1024 + (size_t)STAGES * STAGE_BYTES + W_RESIDENT + 256;
};

template <int CIN, int NOUT, int NKC, bool SPLIT>
__global__ void __launch_bounds__(kDownThreads, 1) reduce3x3s2_tc_kernel(const float* __restrict__ in, const __half* __restrict__ w,
                                                                        __half* __restrict__ o1h, int B, int Hi, int Wi) {
    using Cfg = DownTcCfg<CIN, NOUT, SPLIT>;
    constexpr int KS = Cfg::KS;
    constexpr int kDownStages = Cfg::STAGES;
    constexpr int STAGE = Cfg::STAGE_BYTES;
    static_assert(NKC * 8 == NOUT, "shapes");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* abuf = smem_raw;
    uint8_t* wbuf = abuf + kDownStages * STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(wbuf + Cfg::W_RESIDENT);
    uint64_t* a_full = bars + 0;                    // [4] 8 loader warps arrive
    uint64_t* a_empty = bars + kDownStages;         // [4] MMA commit
    uint64_t* acc_full = bars + 2 * kDownStages;    // [2]
    uint64_t* acc_empty = acc_full + 2;             // [2] 4 epilogue warps arrive
    uint64_t* w_full = acc_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Ho = Hi >> 1, Wo = Wi >> 1;
    const int tiles_x = (Wo + 7) / 8, tiles_y = (Ho + 15) / 16;
    const int total_tiles = B * tiles_x * tiles_y;
    const int my_tiles = ((int)blockIdx.x < total_tiles) ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    pdl_trigger();
    if (tid == 0) {
        for (int s = 0; s < kDownStages; ++s) { tc::mbar_init(a_full + s, SPLIT ? 9 : 8); tc::mbar_init(a_empty + s, 1); }   // 8 loader warps (+ the weight producer)
        for (int s = 0; s < 2; ++s) { tc::mbar_init(acc_full + s, 1); tc::mbar_init(acc_empty + s, 4); }
        tc::mbar_init(w_full, 1);
        tc::mbar_fence_init();
    }
    if (warp == 2) tc::tmem_alloc(tmem_slot, 64);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();   // everything above touched only weights / barriers / TMEM; the input of the kernel before is read below

    if (warp == 0) {
        if (lane == 0) {
            if constexpr (!SPLIT) {     // weights once per CTA
                tc::mbar_expect_tx(w_full, Cfg::W_PART);
                tc::bulk_g2s(wbuf, w, Cfg::W_PART, w_full);
            } else {                    // hi / lo weights of every K step into its stage
                const uint8_t* wg = reinterpret_cast<const uint8_t*>(w);
                int c = 0;
                for (int it = 0; it < my_tiles; ++it)
                    for (int ks = 0; ks < KS; ++ks, ++c) {
                        const int s = c % kDownStages;
                        tc::mbar_wait(a_empty + s, (uint32_t)(((c / kDownStages) & 1) ^ 1));
                        uint8_t* dst = abuf + s * STAGE + 2 * kDownStageBytes;
                        tc::mbar_expect_tx(a_full + s, 2 * Cfg::WK);
                        tc::bulk_g2s(dst, wg + (size_t)ks * Cfg::WK, Cfg::WK, a_full + s);
                        tc::bulk_g2s(dst + Cfg::WK, wg + (size_t)Cfg::W_PART + (size_t)ks * Cfg::WK, Cfg::WK, a_full + s);
                    }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: converged warp, one elected lane issues (tc_common.cuh: elect_one) =====
        {
            constexpr uint32_t idesc = tc::umma_idesc_f16(NOUT);
            constexpr uint32_t a_hi = (uint32_t)((2 * kDownPitch * 16) >> 4) | (1u << 14);   // SBO: next output row = 2 input rows
            constexpr uint32_t b_hi = (uint32_t)(128 >> 4) | (1u << 14);
            const uint32_t a_lo0 = (tc::smem_addr(abuf) >> 4) + ((uint32_t)(kDownChunkBytes >> 4) << 16);
            const uint32_t b_lo0 = (tc::smem_addr(SPLIT ? abuf : wbuf) >> 4) + ((uint32_t)((NOUT * 16) >> 4) << 16);
            if constexpr (!SPLIT) tc::mbar_wait(w_full, 0);
            int c = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int as = it & 1;
                tc::mbar_wait(acc_empty + as, (uint32_t)(((it >> 1) & 1) ^ 1));
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * 32);
#pragma unroll 1
                for (int ks = 0; ks < KS; ++ks, ++c) {
                    const int s = c % kDownStages;
                    tc::mbar_wait(a_full + s, (uint32_t)((c / kDownStages) & 1));
                    __syncwarp();
                    tc::tc_fence_after();
                    const uint32_t a_lo_s = a_lo0 + (uint32_t)(s * (STAGE >> 4));
                    // plain: resident weights of K step ks; split: the stage's own [W_hi | W_lo] behind the two region copies
                    const uint32_t b_lo_s = SPLIT ? b_lo0 + (uint32_t)((s * STAGE + 2 * kDownStageBytes) >> 4) : b_lo0 + (uint32_t)(ks * 9 * 2 * NOUT);
                    if (tc::elect_one()) {
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const int ky = tap / 3, kx = tap % 3;
                        // parity plane (kx != 1), row shift ky, entry shift (kx == 2)
                        const uint32_t aoff = (uint32_t)((kx != 1 ? (kDownParBytes >> 4) : 0) + ky * kDownPitch + (kx == 2 ? 1 : 0));
                        const uint64_t adesc = ((uint64_t)a_hi << 32) | (uint64_t)(a_lo_s + aoff);
                        const uint64_t bdesc = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo_s + (uint32_t)(tap * 2 * NOUT));
                        tc::umma_f16(d_tmem, adesc, bdesc, idesc, (ks | tap) != 0 ? 1u : 0u);
                        if constexpr (SPLIT) {
                            tc::umma_f16(d_tmem, adesc + (uint64_t)(kDownStageBytes >> 4), bdesc, idesc, 1u);     // A_lo x W_hi
                            tc::umma_f16(d_tmem, adesc, bdesc + (uint64_t)(Cfg::WK >> 4), idesc, 1u);            // A_hi x W_lo
                        }
                    }
                    tc::umma_commit(a_empty + s);
                    if (ks == KS - 1) tc::umma_commit(acc_full + as);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 4 && warp < 12) {
        // ===== loaders: 256 threads; task t = (K chunk t / 561, region position t % 561), lanes walk region columns.
        // The global loads of chunk c+1 (next K step, possibly of the next tile) are issued BEFORE chunk c is converted
        // and stored (register double buffer): the loaders are the HBM side of this kernel and were latency-bound. =====
        constexpr int POS = kDownRows * kDownCols;             // 561
        constexpr int TASKS = 2 * POS;                          // per stage
        constexpr int NT = (TASKS + 255) / 256;                 // 5
        const int lt = tid - 128;
        const size_t iplane = (size_t)Hi * Wi;
        int rr[NT], cq[NT], kk[NT], soff[NT];                   // task geometry inside the region: tile independent
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            const int t = lt + 256 * i;
            soff[i] = -1; rr[i] = 0; cq[i] = 0; kk[i] = 0;
            if (t < TASKS) {
                const int k = t / POS, pos = t - k * POS;
                const int r = pos / kDownCols, cc = pos - r * kDownCols;
                rr[i] = r; cq[i] = cc; kk[i] = k;
                soff[i] = k * kDownChunkBytes + ((cc & 1) ? 0 : kDownParBytes) + (r * kDownPitch + (cc >> 1)) * 16;
            }
        }
        const int total = my_tiles * KS;
        // The loaders are instruction-bound (ncu: ~24 SASS instructions per loaded value with naive pointer arithmetic), so
        // addresses are a per-crop 64-bit base + ONE IMAD.WIDE.U32 per load (32-bit plane stride x constant channel index;
        // the host guarantees CIN * Hi * Wi * 4 < 2^32), and the channel bound is only tested in the last, partial K step.
        const uint32_t plane_b = (uint32_t)(iplane * sizeof(float));
        auto issue = [&](int c, float (&v)[NT][8]) {
            const int it = c / KS, ks = c - it * KS;
            const int tile = (int)blockIdx.x + it * (int)gridDim.x;
            const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
            const int y_in0 = 2 * (ty * 16) - 1, x_in0 = 2 * (tx * 8) - 1;
            const char* src_b = reinterpret_cast<const char*>(in + (size_t)b * CIN * iplane);
            const bool full = 16 * ks + 16 <= CIN;              // warp-uniform: every channel of this K step exists
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                const int yi = y_in0 + rr[i], xi = x_in0 + cq[i];
                const bool inb = soff[i] >= 0 && yi >= 0 && yi < Hi && xi >= 0 && xi < Wi;
                const int ch0 = 16 * ks + 8 * kk[i];
                const char* sp = src_b + (uint64_t)plane_b * (uint32_t)ch0 + (uint32_t)((inb ? yi * Wi + xi : 0) * 4);
                if (full) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        v[i][j] = inb ? __ldg(reinterpret_cast<const float*>(sp + (uint64_t)plane_b * (uint32_t)j)) : 0.f;
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        v[i][j] = (inb && ch0 + j < CIN) ? __ldg(reinterpret_cast<const float*>(sp + (uint64_t)plane_b * (uint32_t)j)) : 0.f;
                }
            }
        };
        auto store = [&](int c, const float (&v)[NT][8]) {
            const int s = c % kDownStages;
            tc::mbar_wait(a_empty + s, (uint32_t)(((c / kDownStages) & 1) ^ 1));
            uint8_t* dst = abuf + s * STAGE;
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                if (soff[i] < 0) continue;
                __half2 h[4], l[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if constexpr (SPLIT) split_f16x2(v[i][2 * j], v[i][2 * j + 1], h[j], l[j]);
                    else h[j] = __floats2half2_rn(v[i][2 * j], v[i][2 * j + 1]);
                }
                uint4 u;
                u.x = *reinterpret_cast<uint32_t*>(&h[0]); u.y = *reinterpret_cast<uint32_t*>(&h[1]);
                u.z = *reinterpret_cast<uint32_t*>(&h[2]); u.w = *reinterpret_cast<uint32_t*>(&h[3]);
                *reinterpret_cast<uint4*>(dst + soff[i]) = u;
                if constexpr (SPLIT) {
                    u.x = *reinterpret_cast<uint32_t*>(&l[0]); u.y = *reinterpret_cast<uint32_t*>(&l[1]);
                    u.z = *reinterpret_cast<uint32_t*>(&l[2]); u.w = *reinterpret_cast<uint32_t*>(&l[3]);
                    *reinterpret_cast<uint4*>(dst + kDownStageBytes + soff[i]) = u;
                }
            }
            tc::fence_proxy_async();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(a_full + s);
        };
        float va[NT][8], vb[NT][8];
        if (total > 0) issue(0, va);
#pragma unroll 1
        for (int c = 0; c < total; c += 2) {
            if (c + 1 < total) issue(c + 1, vb);
            store(c, va);
            if (c + 1 < total) {
                if (c + 2 < total) issue(c + 2, va);
                store(c + 1, vb);
            }
        }
    } else if (warp >= 12) {
        // ===== epilogue: TMEM -> fp16 chunk-plane o1h [B][kc][Ho][Wo][8] =====
        const int q = warp & 3;
        const int row = 4 * q + (lane >> 3), col = lane & 7;
        const size_t oplane = (size_t)Ho * Wo;
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = (int)blockIdx.x + it * (int)gridDim.x;
            const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
            const int y = ty * 16 + row, x = tx * 8 + col;
            const int as = it & 1;
            tc::mbar_wait(acc_full + as, (uint32_t)((it >> 1) & 1));
            tc::tc_fence_after();
            float v[NOUT];
            const uint32_t t0 = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(as * 32);
            if constexpr (NOUT == 32) tc::tmem_ld32(t0, v); else tc::tmem_ld16(t0, v);
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(acc_empty + as);
            if (y < Ho && x < Wo) store_o1_chunks<NOUT, NKC, SPLIT>(o1h, B, b, oplane, (size_t)y * Wo + x, v);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem_base, 64);
}

// ------------------------------------------------------------------------------------------------------------------------
// TMA-staged variant (default when the input row pitch is a multiple of 16 B, i.e. Wi % 4 == 0).
//
// The loader warps of the kernel above are INSTRUCTION bound (ncu, profiles/r01_reduce3x3s2_tc3_v1.txt: 14-24 SASS
// instructions per loaded value -- per-load address arithmetic, image-border and channel-bound tests -- at 0.37-0.40 of the HBM
// roof).  Here one elected thread hands the whole 33 x 17 region of one K = 16 step to TMA instead: a 4-D tiled load
// {20 columns, 33 rows, 16 channels, 1 crop} of the planar fp32 input into a shared-memory staging slot.  TMA wants the box's
// innermost start coordinate AND extent on 16 B boundaries (a start at column 2*ox0 - 1 raises "illegal instruction"), so the
// box starts three columns early, at 2*ox0 - 4, and region column cc is box column cc + 3: 20 columns = exactly 80 B.  TMA's out-of-bounds zero fill IS the convolution padding and the channel padding (CIN = 19 / 131 -> 32 /
// 144), so the loader warps are left with one LDS per value, the fp16 (hi / lo) conversion and one STS.128 per 8 channels
// into the same parity-split K-major operand stage as above: no global address arithmetic, no bounds tests.
//   warp 0 = TMA producer of the staging ring, warp 1 = MMA issuer, warp 2 = TMEM allocator, warp 3 = producer of the
//   streamed split weights, warps 4..11 = converters (staging -> operand stage), warps 12..15 = epilogue.
constexpr int kDownBoxCols = 20;                                                  // columns 2*ox0-4 .. 2*ox0+15: 16 B aligned start and size
constexpr int kDownBoxSkip = 3;                                                   // region column 0 (= 2*ox0-1) is box column 3
constexpr int kDownStgBytes = 16 * kDownRows * kDownBoxCols * 4;                  // one staging slot: 42240 B

template <int CIN, int NOUT, bool SPLIT = false>
struct DownTmaCfg {
    static constexpr int KS = (CIN + 15) / 16;
    static constexpr int WK = 9 * 2 * NOUT * 16;
    static constexpr int W_PART = KS * WK;
    static constexpr int W_RESIDENT = SPLIT ? 0 : W_PART;
    static constexpr int STAGES = 2;                                             // operand stages
    static constexpr int SLOTS = 2;                                              // fp32 staging slots
    static constexpr int STAGE_BYTES = SPLIT ? 2 * kDownStageBytes + 2 * WK : kDownStageBytes;
    static constexpr size_t SMEM = 1024 + (size_t)STAGES * STAGE_BYTES + W_RESIDENT + (size_t)SLOTS * kDownStgBytes + 256;
    static_assert(SMEM <= 227 * 1024, "shared memory");
    static_assert((STAGES * STAGE_BYTES + W_RESIDENT) % 128 == 0, "TMA destination alignment");
};

template <int CIN, int NOUT, int NKC, bool SPLIT>
__global__ void __launch_bounds__(kDownThreads, 1) reduce3x3s2_tma_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                         const __grid_constant__ CUtensorMap tmap_tail, const __half* __restrict__ w,
                                                                         __half* __restrict__ o1h, int B, int Hi, int Wi) {
    using Cfg = DownTmaCfg<CIN, NOUT, SPLIT>;
    constexpr int KS = Cfg::KS;
    // The TMA engine walks a box row by row and 80 B rows keep it busier than the bytes suggest (~2.4 k cycles per 528-row box,
    // measured: the chunk period did not change when 13 of the 16 channels of the last K step were out of bounds).  The last K
    // step only has TAIL = CIN % 16 real channels (3 for both DownSamplers), so it goes through a second map whose box holds
    // just those: 99 rows instead of 528; the converters fill the missing channels with zeros.
    constexpr int TAIL = CIN % 16;
    constexpr int NST = Cfg::STAGES, NSL = Cfg::SLOTS;
    constexpr int STAGE = Cfg::STAGE_BYTES;
    static_assert(NKC * 8 == NOUT, "shapes");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* abuf = smem_raw;
    uint8_t* wbuf = abuf + NST * STAGE;
    uint8_t* stg = wbuf + Cfg::W_RESIDENT;
    uint64_t* bars = reinterpret_cast<uint64_t*>(stg + NSL * kDownStgBytes);
    uint64_t* a_full = bars + 0;                 // [NST] 8 converter warps (+ the weight producer when SPLIT)
    uint64_t* a_empty = a_full + NST;            // [NST] MMA commit
    uint64_t* s_full = a_empty + NST;            // [NSL] TMA transaction bytes
    uint64_t* s_empty = s_full + NSL;            // [NSL] 8 converter warps
    uint64_t* acc_full = s_empty + NSL;          // [2]
    uint64_t* acc_empty = acc_full + 2;          // [2] 4 epilogue warps
    uint64_t* w_full = acc_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Ho = Hi >> 1, Wo = Wi >> 1;
    const int tiles_x = (Wo + 7) / 8, tiles_y = (Ho + 15) / 16;
    const int total_tiles = B * tiles_x * tiles_y;
    const int my_tiles = ((int)blockIdx.x < total_tiles) ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int total = my_tiles * KS;

    pdl_trigger();
    if (tid == 0) {
        if ((tc::smem_addr(stg) & 127u) != 0) __trap();   // TMA destination alignment
        for (int s = 0; s < NST; ++s) { tc::mbar_init(a_full + s, SPLIT ? 9 : 8); tc::mbar_init(a_empty + s, 1); }
        for (int s = 0; s < NSL; ++s) { tc::mbar_init(s_full + s, 1); tc::mbar_init(s_empty + s, 8); }
        for (int s = 0; s < 2; ++s) { tc::mbar_init(acc_full + s, 1); tc::mbar_init(acc_empty + s, 4); }
        tc::mbar_init(w_full, 1);
        tc::mbar_fence_init();
        tc::tma_prefetch_desc(&tmap);
        if (TAIL) tc::tma_prefetch_desc(&tmap_tail);
    }
    if (warp == 2) tc::tmem_alloc(tmem_slot, 64);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();   // everything above touched only weights / barriers / TMEM; the input of the kernel before is read below

    if (warp == 0) {
        // ===== TMA producer: one 33 x 20 x 16-channel fp32 box per K step into the staging ring =====
        if (lane == 0) {
            if constexpr (!SPLIT) {     // resident weights once per CTA
                tc::mbar_expect_tx(w_full, Cfg::W_PART);
                tc::bulk_g2s(wbuf, w, Cfg::W_PART, w_full);
            }
            int c = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int tile = (int)blockIdx.x + it * (int)gridDim.x;
                const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
                const int y_in0 = 2 * (ty * 16) - 1, x_in0 = 2 * (tx * 8) - 1 - kDownBoxSkip;
                for (int ks = 0; ks < KS; ++ks, ++c) {
                    const int sl = c % NSL;
                    tc::mbar_wait(s_empty + sl, (uint32_t)(((c / NSL) & 1) ^ 1));
                    if (TAIL != 0 && ks == KS - 1) {
                        tc::mbar_expect_tx(s_full + sl, TAIL * kDownRows * kDownBoxCols * 4);
                        tc::tma_load_4d(stg + sl * kDownStgBytes, &tmap_tail, s_full + sl, x_in0, y_in0, 16 * ks, b);
                    } else {
                        tc::mbar_expect_tx(s_full + sl, kDownStgBytes);
                        tc::tma_load_4d(stg + sl * kDownStgBytes, &tmap, s_full + sl, x_in0, y_in0, 16 * ks, b);
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ===== split mode: hi / lo weights of every K step into its operand stage =====
        if constexpr (SPLIT) {
            if (lane == 0) {
                const uint8_t* wg = reinterpret_cast<const uint8_t*>(w);
                int c = 0;
                for (int it = 0; it < my_tiles; ++it)
                    for (int ks = 0; ks < KS; ++ks, ++c) {
                        const int s = c % NST;
                        tc::mbar_wait(a_empty + s, (uint32_t)(((c / NST) & 1) ^ 1));
                        uint8_t* dst = abuf + s * STAGE + 2 * kDownStageBytes;
                        tc::mbar_expect_tx(a_full + s, 2 * Cfg::WK);
                        tc::bulk_g2s(dst, wg + (size_t)ks * Cfg::WK, Cfg::WK, a_full + s);
                        tc::bulk_g2s(dst + Cfg::WK, wg + (size_t)Cfg::W_PART + (size_t)ks * Cfg::WK, Cfg::WK, a_full + s);
                    }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (identical to reduce3x3s2_tc_kernel) =====
        constexpr uint32_t idesc = tc::umma_idesc_f16(NOUT);
        constexpr uint32_t a_hi = (uint32_t)((2 * kDownPitch * 16) >> 4) | (1u << 14);   // SBO: next output row = 2 input rows
        constexpr uint32_t b_hi = (uint32_t)(128 >> 4) | (1u << 14);
        const uint32_t a_lo0 = (tc::smem_addr(abuf) >> 4) + ((uint32_t)(kDownChunkBytes >> 4) << 16);
        const uint32_t b_lo0 = (tc::smem_addr(SPLIT ? abuf : wbuf) >> 4) + ((uint32_t)((NOUT * 16) >> 4) << 16);
        if constexpr (!SPLIT) tc::mbar_wait(w_full, 0);
        int c = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const int as = it & 1;
            tc::mbar_wait(acc_empty + as, (uint32_t)(((it >> 1) & 1) ^ 1));
            const uint32_t d_tmem = tmem_base + (uint32_t)(as * 32);
#pragma unroll 1
            for (int ks = 0; ks < KS; ++ks, ++c) {
                const int s = c % NST;
                tc::mbar_wait(a_full + s, (uint32_t)((c / NST) & 1));
                __syncwarp();
                tc::tc_fence_after();
                const uint32_t a_lo_s = a_lo0 + (uint32_t)(s * (STAGE >> 4));
                const uint32_t b_lo_s = SPLIT ? b_lo0 + (uint32_t)((s * STAGE + 2 * kDownStageBytes) >> 4) : b_lo0 + (uint32_t)(ks * 9 * 2 * NOUT);
                if (tc::elect_one()) {
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const int ky = tap / 3, kx = tap % 3;
                        const uint32_t aoff = (uint32_t)((kx != 1 ? (kDownParBytes >> 4) : 0) + ky * kDownPitch + (kx == 2 ? 1 : 0));
                        const uint64_t adesc = ((uint64_t)a_hi << 32) | (uint64_t)(a_lo_s + aoff);
                        const uint64_t bdesc = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo_s + (uint32_t)(tap * 2 * NOUT));
                        tc::umma_f16(d_tmem, adesc, bdesc, idesc, (ks | tap) != 0 ? 1u : 0u);
                        if constexpr (SPLIT) {
                            tc::umma_f16(d_tmem, adesc + (uint64_t)(kDownStageBytes >> 4), bdesc, idesc, 1u);     // A_lo x W_hi
                            tc::umma_f16(d_tmem, adesc, bdesc + (uint64_t)(Cfg::WK >> 4), idesc, 1u);            // A_hi x W_lo
                        }
                    }
                    tc::umma_commit(a_empty + s);
                    if (ks == KS - 1) tc::umma_commit(acc_full + as);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4 && warp < 12) {
        // ===== converters: 256 threads; task t = (K chunk t / 561, region position t % 561): 8 LDS (one per channel of the
        // chunk, lanes walk region columns), fp16 / split conversion, one STS.128 (two when SPLIT) =====
        constexpr int POS = kDownRows * kDownCols;             // 561
        constexpr int TASKS = 2 * POS;
        constexpr int NT = (TASKS + 255) / 256;                 // 5
        constexpr int CH_STRIDE = kDownRows * kDownBoxCols;     // floats between channels of the staging box
        const int lt = tid - 128;
        int soff[NT], foff[NT], kk[NT];
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            const int t = lt + 256 * i;
            soff[i] = -1; foff[i] = 0; kk[i] = 0;
            if (t < TASKS) {
                const int k = t / POS, pos = t - k * POS;
                kk[i] = k;
                const int r = pos / kDownCols, cc = pos - r * kDownCols;
                soff[i] = k * kDownChunkBytes + ((cc & 1) ? 0 : kDownParBytes) + (r * kDownPitch + (cc >> 1)) * 16;
                foff[i] = (8 * k * kDownRows + r) * kDownBoxCols + cc + kDownBoxSkip;
            }
        }
#pragma unroll 1
        for (int c = 0; c < total; ++c) {
            const int sl = c % NSL, s = c % NST;
            tc::mbar_wait(s_full + sl, (uint32_t)((c / NSL) & 1));
            const float* src = reinterpret_cast<const float*>(stg + sl * kDownStgBytes);
            float v[NT][8];
            const bool tail = TAIL != 0 && (c % KS) == KS - 1;        // warp-uniform: only TAIL channels of this K step were loaded
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                if (soff[i] < 0) continue;
                if (!tail) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[i][j] = src[foff[i] + j * CH_STRIDE];
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[i][j] = (8 * kk[i] + j < TAIL) ? src[foff[i] + j * CH_STRIDE] : 0.f;
                }
            }
            tc::mbar_wait(a_empty + s, (uint32_t)(((c / NST) & 1) ^ 1));
            uint8_t* dst = abuf + s * STAGE;
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                if (soff[i] < 0) continue;
                __half2 h[4], l[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if constexpr (SPLIT) split_f16x2(v[i][2 * j], v[i][2 * j + 1], h[j], l[j]);
                    else h[j] = __floats2half2_rn(v[i][2 * j], v[i][2 * j + 1]);
                }
                uint4 u;
                u.x = *reinterpret_cast<uint32_t*>(&h[0]); u.y = *reinterpret_cast<uint32_t*>(&h[1]);
                u.z = *reinterpret_cast<uint32_t*>(&h[2]); u.w = *reinterpret_cast<uint32_t*>(&h[3]);
                *reinterpret_cast<uint4*>(dst + soff[i]) = u;
                if constexpr (SPLIT) {
                    u.x = *reinterpret_cast<uint32_t*>(&l[0]); u.y = *reinterpret_cast<uint32_t*>(&l[1]);
                    u.z = *reinterpret_cast<uint32_t*>(&l[2]); u.w = *reinterpret_cast<uint32_t*>(&l[3]);
                    *reinterpret_cast<uint4*>(dst + kDownStageBytes + soff[i]) = u;
                }
            }
            tc::fence_proxy_async();
            __syncwarp();
            if (lane == 0) { tc::mbar_arrive(s_empty + sl); tc::mbar_arrive(a_full + s); }   // the slot's values sit in registers / the stage
        }
    } else if (warp >= 12) {
        // ===== epilogue: TMEM -> fp16 chunk-plane o1h [B][kc][Ho][Wo][8] =====
        const int q = warp & 3;
        const int row = 4 * q + (lane >> 3), col = lane & 7;
        const size_t oplane = (size_t)Ho * Wo;
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = (int)blockIdx.x + it * (int)gridDim.x;
            const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
            const int y = ty * 16 + row, x = tx * 8 + col;
            const int as = it & 1;
            tc::mbar_wait(acc_full + as, (uint32_t)((it >> 1) & 1));
            tc::tc_fence_after();
            float v[NOUT];
            const uint32_t t0 = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(as * 32);
            if constexpr (NOUT == 32) tc::tmem_ld32(t0, v); else tc::tmem_ld16(t0, v);
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(acc_empty + as);
            if (y < Ho && x < Wo) store_o1_chunks<NOUT, NKC, SPLIT>(o1h, B, b, oplane, (size_t)y * Wo + x, v);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem_base, 64);
}

}  // namespace espnet
