// CTA-pair (tcgen05 cta_group::2) variant of the ESP branch stage kernel (kernels_tc_branch.cuh has the semantics, the
// operand layouts and the single-CTA kernel this one is derived from).
//
// Why pairs: with N <= 32 accumulator columns per (branch, tap) a single-CTA M = 128 MMA occupies the tensor pipe for a
// fixed ~57 cycles however small N is, and the 3-term split variants are bound by exactly that (profiles/).
// One cta_group::2 MMA covers M = 256 = the tiles of BOTH CTAs of a cluster of two in ~39 cycles (measured), i.e.
// ~2.9x the MMA throughput per SM pair.  Each CTA keeps its own TMA ring, accumulators and epilogue; only the issue of
// the MMAs is centralised in the leader (cluster rank 0):
//   * A: every CTA's producer loads the halo box of ITS tile into ITS shared memory; the completion bytes of both CTAs
//     are counted on the LEADER's a_full barrier (cp.async.bulk.tensor .cta_group::2 with the mapa'd barrier address).
//   * B: the N rows of every MMA are split between the CTAs (rows [0, N/2) from rank 0, [N/2, N) from rank 1, at the
//     same shared-memory offsets), so each CTA holds HALF of the weights -- all 3-term split units of level 3 become
//     resident (92 KB) and the weight streaming ring of the single-CTA kernel disappears.
//   * D: each CTA's TMEM receives the 128 x N accumulator of its own tile.
//   * tcgen05.commit ... multicast::cluster arrives on the barrier at the same offset in both CTAs (a_empty, acc_full);
//     the peer's epilogue warps release accumulator stages by remote arrives on the leader's acc_empty barrier.
// Weight image per CTA and unit: [9 taps][2 K chunks][5 * HB rows][8] fp16 with HB = NB / 2; outer taps: rows
// br * HB + i = accumulator column rank * HB + i of branch br; centre tap (ONE N = 5 NB MMA): row R = column
// rank * 5 HB + R of the concatenated 5 NB columns.  Units: plain KS (one per K step); split 2 KS (hi then lo);
// level-2 split ("merge", one K step): unit 0 = [W_hi | W_lo] column pairs, unit 1 = [W_hi | 0] for the A_lo term.
#pragma once
#include "kernels_tc_branch.cuh"

namespace espnet {

template <int NKC, int NOUT, bool SPLIT = false>
struct BranchPairCfg {
    static constexpr int KS = NKC / 2;
    static constexpr bool MERGE = SPLIT && KS == 1;
    static constexpr int NB = MERGE ? 2 * NOUT : NOUT;            // accumulator columns per branch
    static constexpr int HB = NB / 2;                             // weight rows per branch held by one CTA
    static constexpr int W_UNIT = 9 * 2 * 5 * HB * 16;            // bytes per CTA
    static constexpr int UNITS = SPLIT ? (MERGE ? 2 : 2 * KS) : KS;
    static constexpr int W_BYTES = UNITS * W_UNIT;                // per CTA, resident
    static constexpr int ACC_COLS = 5 * NB;
    static constexpr int TMEM_COLS = (kTcAccStages * ACC_COLS <= 256) ? 256 : 512;
    static constexpr int EP_BYTES = 2 * 128 * 16;
    static constexpr size_t SMEM = 1024 + 2 * (size_t)kTcStage + (size_t)W_BYTES + EP_BYTES + 256;
    static_assert(kTcAccStages * ACC_COLS <= 512, "TMEM columns");
    static_assert(SMEM <= 232448, "shared memory");
};

// accumulator column (within the branch's NB columns) of output channel j; `lo` selects the W_lo half of a merged pair
template <int NOUT, bool MERGE>
__host__ __device__ constexpr int pair_acc_col(int j, bool lo) {
    if (!MERGE) return j;
    // [hi 0..NOUT/2) | lo 0..NOUT/2) | hi NOUT/2.. | lo NOUT/2..): rank 0's rows then rank 1's rows
    return (j < NOUT / 2 ? j : j + NOUT / 2) + (lo ? NOUT / 2 : 0);
}

template <int NKC, int NOUT, int CO1, int CO, int VAR, bool SPLIT>
__global__ void __launch_bounds__(kTcThreads, 1) esp_branch_tc_pair_kernel(const __grid_constant__ CUtensorMap tmap, const BranchTcParams p) {
    using Cfg = BranchPairCfg<NKC, NOUT, SPLIT>;
    constexpr int C = CO1 + 4 * CO;
    constexpr int KS = Cfg::KS, NB = Cfg::NB, HB = Cfg::HB;
    static_assert(C <= 128 && CO1 <= NOUT && CO <= NOUT && (NOUT == 16 || NOUT == 32), "channel counts");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* abuf = smem_raw;
    uint8_t* wbuf = abuf + 2 * kTcStage;
    float4* sep4 = reinterpret_cast<float4*>(wbuf + Cfg::W_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sep4) + Cfg::EP_BYTES);
    uint64_t* a_full = bars + 0;      // [2] used in the leader: bytes of both CTAs
    uint64_t* a_empty = bars + 2;     // [2] both CTAs (multicast commit)
    uint64_t* w_full = bars + 4;      // [1] own weights
    uint64_t* acc_full = bars + 5;    // [3] both CTAs (multicast commit)
    uint64_t* acc_empty = bars + 8;   // [3] used in the leader: 16 epilogue warps of each CTA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);
    uint32_t* epi_done = tmem_slot + 1;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = tc::cluster_ctarank();
    const int H = p.H, W = p.W;
    const int tiles_x = (W + kTcTileW - 1) / kTcTileW, tiles_y = (H + kTcTileH - 1) / kTcTileH;
    const int total_tiles = p.B * tiles_x * tiles_y;
    const int total_pairs = (total_tiles + 1) / 2;
    const int pair = (int)blockIdx.x >> 1, npairs = (int)gridDim.x >> 1;
    const int my_iters = pair < total_pairs ? (total_pairs - pair + npairs - 1) / npairs : 0;   // identical in both CTAs
    // tile of this CTA in pair iteration `it`; an odd tile count leaves one dummy slot that repeats the last tile
    auto tile_of = [&](int it, bool& real) {
        const int t = 2 * (pair + it * npairs) + (int)rank;
        real = t < total_tiles;
        return real ? t : total_tiles - 1;
    };

    if (tid == 0) {
        if ((tc::smem_addr(abuf) & 127u) != 0) __trap();
        tc::mbar_init(a_full + 0, 1); tc::mbar_init(a_full + 1, 1);
        tc::mbar_init(a_empty + 0, 1); tc::mbar_init(a_empty + 1, 1);
        tc::mbar_init(w_full, 1);
        for (int s = 0; s < kTcAccStages; ++s) { tc::mbar_init(acc_full + s, 1); tc::mbar_init(acc_empty + s, 32); }
        *epi_done = 0;
        tc::mbar_fence_init();
        tc::tma_prefetch_desc(&tmap);
        // this CTA's half of the weights, resident for the whole kernel
        tc::mbar_expect_tx(w_full, Cfg::W_BYTES);
        tc::bulk_g2s(wbuf, reinterpret_cast<const uint8_t*>(p.w) + (size_t)rank * Cfg::W_BYTES, Cfg::W_BYTES, w_full);
    }
    if (warp == 2) tc::tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
    for (int i = tid; i < C; i += kTcThreads) {
        sep4[i] = make_float4(p.s[i], p.t[i], p.a[i], 0.f);
        if (VAR != 1) sep4[128 + i] = make_float4(p.s2[p.c2_off + i], p.t2[p.c2_off + i], p.a2[p.c2_off + i], 0.f);
    }
    __syncthreads();                       // barrier init visible to this CTA's waiters
    tc::mbar_wait(w_full, 0);              // own weights have landed
    tc::tc_fence_before();
    tc::cluster_sync();                    // both CTAs: barriers initialised, weights resident, TMEM allocated
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== producer (both CTAs): own halo boxes, completion counted on the leader's barrier =====
        if (lane == 0) {
            const uint32_t full0 = tc::mapa(a_full + 0, 0), full1 = tc::mapa(a_full + 1, 0);
            auto load = [&](int s, uint32_t parity, int cx, int cy, int kc, int b) {
                tc::mbar_wait(a_empty + s, parity ^ 1);
                if (rank == 0) tc::mbar_expect_tx(a_full + s, 2 * kTcStage);
                tc::tma_load_4d_pair(abuf + s * kTcStage, &tmap, s ? full1 : full0, cx, cy, kc, b);
            };
            int n = 0;
            for (int it = 0; it < my_iters; ++it) {
                bool real;
                const int tile = tile_of(it, real);
                const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
                const int cx = 4 * (tx * kTcTileW - kTcHalo2), cy = ty * kTcTileH - kTcHalo2;
                if constexpr (!SPLIT) {
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks, ++n) load(n & 1, (uint32_t)((n >> 1) & 1), cx, cy, 2 * ks, b);
                } else if constexpr (Cfg::MERGE) {
                    load(0, (uint32_t)(it & 1), cx, cy, 0, b);             // A_hi
                    load(1, (uint32_t)(it & 1), cx, cy, 0, b + p.B);       // A_lo
                } else {
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks, ++n) {
                        load(1, (uint32_t)(n & 1), cx, cy, 2 * ks, b + p.B);   // A_lo
                        load(0, (uint32_t)(n & 1), cx, cy, 2 * ks, b);         // A_hi
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: leader CTA only; the warp runs converged, one elected lane issues =====
        if (rank == 0) {
            constexpr uint32_t idesc = tc::umma_idesc_f16_pair(NB), idesc_c = tc::umma_idesc_f16_pair(5 * NB);
            constexpr uint32_t a_hi = (uint32_t)((kTcBoxW * 16) >> 4) | (1u << 14);
            constexpr uint32_t b_hi = (uint32_t)(128 >> 4) | (1u << 14);
            constexpr uint32_t TAP = 2 * 5 * HB, UNIT = (uint32_t)(Cfg::W_UNIT >> 4);
            const uint32_t a_lo0 = (tc::smem_addr(abuf) >> 4) + (uint32_t)(kTcHalo2 * kTcBoxW + kTcHalo2) + ((uint32_t)(kTcPlane2 >> 4) << 16);
            const uint32_t a_lo1 = a_lo0 + (uint32_t)(kTcStage >> 4);
            const uint32_t b_lo0 = (tc::smem_addr(wbuf) >> 4) + ((uint32_t)((5 * HB * 16) >> 4) << 16);
            // 41 MMAs: centre tap of all branches (N = 5 NB), then 8 outer taps per branch (N = NB)
            auto issue_step = [&](uint32_t a_lo_s, uint32_t b_lo_s, uint32_t d_tile, bool fresh, uint64_t* c0, uint64_t* c1, uint64_t* c2) {
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    tc::umma_f16_pair(d_tile, ((uint64_t)a_hi << 32) | (uint64_t)a_lo_s, ((uint64_t)b_hi << 32) | (uint64_t)(b_lo_s + 4u * TAP),
                                      idesc_c, fresh ? 0u : 1u);
#pragma unroll 1
                    for (int br = 0; br < 5; ++br) {
                        const int d = 1 << br, dp = d * kTcBoxW;
                        const uint32_t b_lo = b_lo_s + (uint32_t)(br * HB);
                        const uint32_t d_tmem = d_tile + (uint32_t)(br * NB);
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            if (tap == 4) continue;
                            const int ky = tap / 3 - 1, kx = tap % 3 - 1;
                            const uint64_t adesc = ((uint64_t)a_hi << 32) | (uint64_t)(a_lo_s + (uint32_t)(ky * dp + kx * d));
                            const uint64_t bdesc = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + (uint32_t)tap * TAP);
                            tc::umma_f16_pair(d_tmem, adesc, bdesc, idesc, 1u);
                        }
                    }
                    if (c0) tc::umma_commit_pair(c0);
                    if (c1) tc::umma_commit_pair(c1);
                    if (c2) tc::umma_commit_pair(c2);
                }
                __syncwarp();
            };
            int n = 0;
            for (int it = 0; it < my_iters; ++it) {
                const int as = it % kTcAccStages;
                tc::mbar_wait(acc_empty + as, (uint32_t)(((it / kTcAccStages) & 1) ^ 1));
                const uint32_t d_tile = tmem_base + (uint32_t)(as * Cfg::ACC_COLS);
                if constexpr (!SPLIT) {
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks, ++n) {
                        const int s = n & 1;
                        tc::mbar_wait(a_full + s, (uint32_t)((n >> 1) & 1));
                        __syncwarp();
                        issue_step(s ? a_lo1 : a_lo0, b_lo0 + (uint32_t)ks * UNIT, d_tile, ks == 0, a_empty + s,
                                   ks == KS - 1 ? acc_full + as : nullptr, nullptr);
                    }
                } else if constexpr (Cfg::MERGE) {
                    const uint32_t par = (uint32_t)(it & 1);
                    tc::mbar_wait(a_full + 0, par);                         // A_hi x [W_hi | W_lo]
                    __syncwarp();
                    issue_step(a_lo0, b_lo0, d_tile, true, a_empty + 0, nullptr, nullptr);
                    tc::mbar_wait(a_full + 1, par);                         // A_lo x [W_hi | 0]
                    __syncwarp();
                    issue_step(a_lo1, b_lo0 + UNIT, d_tile, false, a_empty + 1, acc_full + as, nullptr);
                } else {
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks, ++n) {
                        const uint32_t par = (uint32_t)(n & 1);
                        const uint32_t w_hi = b_lo0 + (uint32_t)ks * UNIT, w_lo = b_lo0 + (uint32_t)(KS + ks) * UNIT;
                        tc::mbar_wait(a_full + 1, par);                     // A_lo x W_hi
                        __syncwarp();
                        issue_step(a_lo1, w_hi, d_tile, ks == 0, a_empty + 1, nullptr, nullptr);
                        tc::mbar_wait(a_full + 0, par);                     // A_hi x W_hi, A_hi x W_lo
                        __syncwarp();
                        issue_step(a_lo0, w_hi, d_tile, false, nullptr, nullptr, nullptr);
                        issue_step(a_lo0, w_lo, d_tile, false, a_empty + 0, ks == KS - 1 ? acc_full + as : nullptr, nullptr);
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ===== L2 prefetcher of the residual (see the single-CTA kernel) =====
        if (VAR != 0) {
            const size_t plane = (size_t)H * W;
            for (int it = 0; it < my_iters; ++it) {
                while (*reinterpret_cast<volatile uint32_t*>(epi_done) + (uint32_t)kTcPrefetchLead <= (uint32_t)it) __nanosleep(200);
                bool real;
                const int tile = tile_of(it, real);
                if (!real) continue;
                const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
                const int r = lane & 15, y = ty * kTcTileH + r, x = tx * kTcTileW;
                if (y >= H) continue;
                const float* base = p.res + (size_t)b * C * plane + (size_t)y * W + x;
#pragma unroll 4
                for (int c = lane >> 4; c < C; c += 2)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)c * plane));
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue (both CTAs, own tile): identical to the single-CTA kernel except for the accumulator column map
        // of the merged variant and the remote release of the accumulator stage =====
        constexpr bool HAS_RES = VAR != 0, HAS_OUT = VAR != 2, HAS_OUT2 = VAR != 1;
        constexpr int GW = NOUT / 4;
        const int e = warp - 4, q = e & 3, gsel = e >> 2;
        const int row = 4 * q + (lane >> 3), col = lane & 7;
        const size_t plane = (size_t)H * W;
        const uint32_t plane_b = (uint32_t)(plane * sizeof(float));
        const uint32_t empty0 = tc::mapa(acc_empty, 0);                     // leader's acc_empty[0]
        for (int it = 0; it < my_iters; ++it) {
            bool real;
            const int tile = tile_of(it, real);
            const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
            const int y = ty * kTcTileH + row, x = tx * kTcTileW + col;
            const bool valid = real && (y < H) && (x < W);
            const size_t pix = valid ? (size_t)y * W + x : 0;
            const char* res_b = HAS_RES ? reinterpret_cast<const char*>(p.res + (size_t)b * C * plane + pix) : nullptr;
            char* out_b = HAS_OUT ? reinterpret_cast<char*>(p.out + (size_t)b * C * plane + pix) : nullptr;
            char* out2_b = HAS_OUT2 ? reinterpret_cast<char*>(p.out2 + ((size_t)b * p.C2 + p.c2_off) * plane + pix) : nullptr;
            const int as = it % kTcAccStages;
            const uint32_t t0 = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(as * Cfg::ACC_COLS);
            auto group = [&](auto gtag) {
                constexpr int G = decltype(gtag)::value;
                constexpr int col_hi = pair_acc_col<NOUT, Cfg::MERGE>(GW * G, false), col_lo = pair_acc_col<NOUT, Cfg::MERGE>(GW * G, true);
                float rv[5][GW];
#pragma unroll
                for (int br = 0; br < 5; ++br)
#pragma unroll
                    for (int jj = 0; jj < GW; ++jj) rv[br][jj] = 0.f;
#ifdef ESPNET_TC_EXP
                if (HAS_RES && valid && !(ESPNET_TC_EXP & 4)) {
#else
                if (HAS_RES && valid) {
#endif
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
#ifdef ESPNET_TC_EXP
                    if (ESPNET_TC_EXP & 4) break;      // timing experiment: no epilogue math
#endif
                    const int ch0 = br == 0 ? 0 : CO1 + (br - 1) * CO;
                    const int cnt = br == 0 ? CO1 : CO;
                    uint32_t r[GW], r2[GW];
                    __syncwarp();
                    if constexpr (GW == 8) tc::tmem_ld8_nowait(t0 + (uint32_t)(br * NB + col_hi), r);
                    else tc::tmem_ld4_nowait(t0 + (uint32_t)(br * NB + col_hi), r);
                    if constexpr (Cfg::MERGE) {
                        if constexpr (GW == 8) tc::tmem_ld8_nowait(t0 + (uint32_t)(br * NB + col_lo), r2);
                        else tc::tmem_ld4_nowait(t0 + (uint32_t)(br * NB + col_lo), r2);
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
                tc::mbar_arrive_cluster(empty0 + (uint32_t)(as * sizeof(uint64_t)));
                if (e == 0) *reinterpret_cast<volatile uint32_t*>(epi_done) = (uint32_t)(it + 1);
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::cluster_sync();                    // no CTA leaves while its peer may still touch its shared memory / barriers
    if (warp == 2) tc::tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace espnet
