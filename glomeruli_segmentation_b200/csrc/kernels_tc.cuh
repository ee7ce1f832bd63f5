// tcgen05 / TMEM kernels of the ESP "split-transform-merge" stage (sm_100a): fp16 operands, fp32 accumulation
// in tensor memory, fp32 epilogue.  ESPNET_MODE_F16TC of include/espnet_b200.h.
//
// Reference semantics: DilatedParllelResidualBlockB.forward (Model.py:187-214) and DownSamplerB.forward
// (Model.py:144-160) after the c1 reduce: d_k = Conv3x3(dilation k, pad k)(o1), HFF prefix sums, concat,
// residual add before BN, BN(eval, eps 1e-3) folded to scale/shift, PReLU.
//
// Implicit GEMM per CTA tile of 16x16 output pixels (two M = 128 MMA tiles of 8 columns x 16 rows):
//   * A operand: the reduced map o1 in fp16 "chunk-plane" layout [B][kc][H][W][8 ch] (tc_common.cuh).  One 5-D TMA
//     box {8, 48, 48, NKC, 1} brings the tile plus the 16-pixel halo of the d = 16 branch into shared memory; the
//     zero padding of every conv is TMA's out-of-bounds fill.  A tap (ky,kx) of dilation d is the SAME buffer read
//     through a UMMA descriptor whose start address is moved by ((ky-1)*d*48 + (kx-1)*d) * 16 B.
//   * B operand: per-branch weights [tap][kc][NOUT][8] fp16, streamed branch by branch with cp.async.bulk into a
//     2-deep ring (all five branches do not fit next to the 144 KB activation box at level 3).
//   * D: five accumulators of NOUT fp32 columns per MMA tile in TMEM; the HFF sums are formed in the epilogue.
//   * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warp 2 = TMEM allocator,
//     warps 4..11 = epilogue (TMEM -> registers -> residual + BN + PReLU -> coalesced planar fp32 stores).
#pragma once
#include "kernels_fp32.cuh"
#include "tc_common.cuh"

namespace espnet {

inline bool tc_path_available() { return true; }

constexpr int kTcTile = 16;                        // output tile edge (pixels)
constexpr int kTcHalo = 16;                        // largest dilation
constexpr int kTcBox = kTcTile + 2 * kTcHalo;      // 48: staged region edge
constexpr int kTcPlaneBytes = kTcBox * kTcBox * 16;
constexpr int kTcThreads = 384;

// ------------------------------------------------------------------------------------------------
// fp16 chunk-plane stores for the reduce kernels: o1h[((b*NKC + kc)*HW + pix)*8 + j] = half(acc[kc*8+j])
// ------------------------------------------------------------------------------------------------
template <int CO, int NKC>
__device__ __forceinline__ void store_o1_f16(__half* __restrict__ o1h, int b, size_t HW, size_t pix, const float (&acc)[CO]) {
#pragma unroll
    for (int kc = 0; kc < NKC; ++kc) {
        __half2 h[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c0 = kc * 8 + 2 * j, c1 = c0 + 1;
            const float v0 = c0 < CO ? acc[c0 < CO ? c0 : 0] : 0.f;
            const float v1 = c1 < CO ? acc[c1 < CO ? c1 : 0] : 0.f;
            h[j] = __floats2half2_rn(v0, v1);
        }
        uint4 u;
        u.x = *reinterpret_cast<uint32_t*>(&h[0]); u.y = *reinterpret_cast<uint32_t*>(&h[1]);
        u.z = *reinterpret_cast<uint32_t*>(&h[2]); u.w = *reinterpret_cast<uint32_t*>(&h[3]);
        *reinterpret_cast<uint4*>(o1h + (((size_t)b * NKC + kc) * HW + pix) * 8) = u;
    }
}

// ESP reduce (Model.py:178,192): 1x1 conv CIN -> CO on CUDA cores in fp32, result rounded once to fp16.
template <int CIN, int CO, int NKC>
__global__ void __launch_bounds__(256) reduce1x1_f16_kernel(const float* __restrict__ in, const float* __restrict__ w /*[CIN][pad4(CO)]*/,
                                                            __half* __restrict__ o1h, int B, int HW) {
    constexpr int CP = pad4(CO);
    constexpr int G = 4;
    static_assert(CIN % (2 * G) == 0, "channel groups are processed in ping-pong pairs");
    extern __shared__ __align__(16) float smem[];
    copy_to_smem(smem, w, CIN * CP);
    __syncthreads();
    const int chunks = (HW + 127) / 128;
    const int total = B * chunks;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int k = warp;; k += nwarp) {
        const int item = k * gridDim.x + blockIdx.x;
        if (item >= total) break;
        const int b = item / chunks;
        const int p0 = (item % chunks) * 128 + lane;
        bool ok[kRows];
#pragma unroll
        for (int r = 0; r < kRows; ++r) ok[r] = (p0 + 32 * r) < HW;
        float acc[kRows][CO];
#pragma unroll
        for (int r = 0; r < kRows; ++r)
#pragma unroll
            for (int j = 0; j < CO; ++j) acc[r][j] = 0.f;
        const float* src = in + (size_t)b * CIN * HW + p0;
        auto load = [&](float (&a)[G][kRows], int g) {
#pragma unroll
            for (int j = 0; j < G; ++j)
#pragma unroll
                for (int r = 0; r < kRows; ++r) a[j][r] = ok[r] ? __ldg(src + (size_t)(g * G + j) * HW + 32 * r) : 0.f;
        };
        float a0[G][kRows], a1[G][kRows];
        load(a0, 0);
#pragma unroll 1
        for (int g = 0; g < CIN / G; g += 2) {
            load(a1, g + 1);
#pragma unroll
            for (int j = 0; j < G; ++j) fma_tile<CO>(acc, a0[j], smem + (g * G + j) * CP);
            if (g + 2 < CIN / G) load(a0, g + 2);
#pragma unroll
            for (int j = 0; j < G; ++j) fma_tile<CO>(acc, a1[j], smem + ((g + 1) * G + j) * CP);
        }
#pragma unroll
        for (int r = 0; r < kRows; ++r)
            if (ok[r]) store_o1_f16<CO, NKC>(o1h, b, (size_t)HW, (size_t)(p0 + 32 * r), acc[r]);
    }
}

// DownSamplerB reduce (Model.py:135,145): 3x3 stride-2 conv CIN -> CO, fp32 CUDA cores, fp16 chunk-plane output.
template <int CIN, int CO, int NKC>
__global__ void __launch_bounds__(kHeavyThreads, 1) reduce3x3s2_f16_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                                           __half* __restrict__ o1h, int B, int Hi, int Wi) {
    constexpr int CP = pad4(CO);
    constexpr int G = 4;
    extern __shared__ __align__(16) float smem[];
    copy_to_smem(smem, w, 9 * CIN * CP);
    for (int i = threadIdx.x; i < G * CP; i += blockDim.x) smem[9 * CIN * CP + i] = 0.f;
    __syncthreads();
    const int Ho = Hi >> 1, Wo = Wi >> 1;
    const size_t iplane = (size_t)Hi * Wi;
    const TileIter it(B, Ho, Wo);
    const int warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int k = warp;; k += nwarp) {
        const int item = k * gridDim.x + blockIdx.x;
        if (item >= it.total) break;
        int b, y0, x;
        it.decode(item, b, y0, x);
        float acc[kRows][CO];
#pragma unroll
        for (int r = 0; r < kRows; ++r)
#pragma unroll
            for (int j = 0; j < CO; ++j) acc[r][j] = 0.f;
        Conv3x3Pipe<CIN, CO, G, 2> pipe{in + (size_t)b * CIN * iplane, iplane, Hi, Wi, Wi, Ho, Wo, y0, x, 1};
        pipe.run(acc, smem);
        if (x < Wo) {
#pragma unroll
            for (int r = 0; r < kRows; ++r)
                if (y0 + r < Ho) store_o1_f16<CO, NKC>(o1h, b, (size_t)Ho * Wo, (size_t)(y0 + r) * Wo + x, acc[r]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The tensor-core branch kernel
// ------------------------------------------------------------------------------------------------
struct BranchTcParams {
    const __half* w;        // [5][9][NKC][NOUT][8] fp16 (d1, d2, d4, d8, d16)
    const float* res;       // [B,C,H,W] residual input or nullptr
    const float *s, *t, *a; // BN scale/shift + PReLU slope (C)
    float* out;             // [B,C,H,W] or nullptr
    const float *s2, *t2, *a2;
    float* out2;            // [B,C2,H,W] or nullptr (second BR straight into the following concat buffer)
    int C2, c2_off;
    int B, H, W;
};

template <int NKC, int NOUT>
struct BranchTcCfg {
    static constexpr int A_BYTES = NKC * kTcPlaneBytes;
    static constexpr int W_BYTES = 9 * NKC * NOUT * 16;          // one branch
    static constexpr int ACC_COLS = 5 * NOUT;                     // per MMA tile
    static constexpr int TMEM_COLS = (2 * ACC_COLS <= 256) ? 256 : 512;
    static constexpr int EP_FLOATS = 6 * 128;
    static constexpr size_t SMEM = 1024 + (size_t)A_BYTES + 2 * (size_t)W_BYTES + EP_FLOATS * 4 + 128;
};

template <int NKC, int NOUT, int CO1, int CO>
__global__ void __launch_bounds__(kTcThreads, 1) esp_branch_tc_kernel(const __grid_constant__ CUtensorMap tmap, const BranchTcParams p) {
    using Cfg = BranchTcCfg<NKC, NOUT>;
    constexpr int C = CO1 + 4 * CO;
    static_assert(C <= 128 && CO1 <= NOUT && CO <= NOUT, "channel counts");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [A box | W ring | epilogue params | barriers | tmem slot]; the dynamic segment is 1024 B aligned already
    uint8_t* abuf = smem_raw;
    uint8_t* wbuf = abuf + Cfg::A_BYTES;
    float* sep = reinterpret_cast<float*>(wbuf + 2 * Cfg::W_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sep + Cfg::EP_FLOATS);
    uint64_t* a_full = bars + 0;
    uint64_t* a_empty = bars + 1;
    uint64_t* w_full = bars + 2;    // [2]
    uint64_t* w_empty = bars + 4;   // [2]
    uint64_t* acc_full = bars + 6;
    uint64_t* acc_empty = bars + 7;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = p.H, W = p.W;
    const int tiles_x = (W + kTcTile - 1) / kTcTile, tiles_y = (H + kTcTile - 1) / kTcTile;
    const int total_tiles = p.B * tiles_x * tiles_y;
    const int my_tiles = ((int)blockIdx.x < total_tiles) ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (tid == 0) {
        if ((tc::smem_addr(abuf) & 127u) != 0) __trap();   // TMA destination alignment
        tc::mbar_init(a_full, 1); tc::mbar_init(a_empty, 1);
        tc::mbar_init(w_full + 0, 1); tc::mbar_init(w_full + 1, 1);
        tc::mbar_init(w_empty + 0, 1); tc::mbar_init(w_empty + 1, 1);
        tc::mbar_init(acc_full, 1); tc::mbar_init(acc_empty, 8);
        tc::mbar_fence_init();
        tc::tma_prefetch_desc(&tmap);
    }
    if (warp == 2) tc::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    for (int i = tid; i < C; i += kTcThreads) {
        sep[i] = p.s[i]; sep[128 + i] = p.t[i]; sep[256 + i] = p.a[i];
        if (p.out2) { sep[384 + i] = p.s2[p.c2_off + i]; sep[512 + i] = p.t2[p.c2_off + i]; sep[640 + i] = p.a2[p.c2_off + i]; }
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== producer: activation box per tile, weights per (tile, branch) =====
        if (lane == 0) {
            for (int it = 0; it < my_tiles; ++it) {
                const int tile = (int)blockIdx.x + it * (int)gridDim.x;
                const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
                tc::mbar_wait(a_empty, (uint32_t)((it & 1) ^ 1));
                tc::mbar_expect_tx(a_full, Cfg::A_BYTES);
                tc::tma_load_5d(abuf, &tmap, a_full, 0, tx * kTcTile - kTcHalo, ty * kTcTile - kTcHalo, 0, b);
                for (int br = 0; br < 5; ++br) {
                    const int wi = it * 5 + br, buf = wi & 1;
                    tc::mbar_wait(w_empty + buf, (uint32_t)(((wi >> 1) & 1) ^ 1));
                    tc::mbar_expect_tx(w_full + buf, Cfg::W_BYTES);
                    tc::bulk_g2s(wbuf + buf * Cfg::W_BYTES, reinterpret_cast<const uint8_t*>(p.w) + (size_t)br * Cfg::W_BYTES, Cfg::W_BYTES,
                                 w_full + buf);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = tc::umma_idesc_f16(NOUT);
            const uint32_t a_addr = tc::smem_addr(abuf), w_addr = tc::smem_addr(wbuf);
            for (int it = 0; it < my_tiles; ++it) {
                tc::mbar_wait(a_full, (uint32_t)(it & 1));
                tc::mbar_wait(acc_empty, (uint32_t)((it & 1) ^ 1));
                tc::tc_fence_after();
                for (int br = 0; br < 5; ++br) {
                    const int wi = it * 5 + br, buf = wi & 1;
                    tc::mbar_wait(w_full + buf, (uint32_t)((wi >> 1) & 1));
                    tc::tc_fence_after();
                    const int d = 1 << br;
#pragma unroll 1
                    for (int m = 0; m < 2; ++m) {
                        const uint32_t d_tmem = tmem_base + (uint32_t)(m * Cfg::ACC_COLS + br * NOUT);
#pragma unroll 1
                        for (int tap = 0; tap < 9; ++tap) {
                            const int ky = tap / 3, kx = tap - 3 * ky;
                            const int pos = (kTcHalo + (ky - 1) * d) * kTcBox + kTcHalo + 8 * m + (kx - 1) * d;
#pragma unroll
                            for (int ks = 0; ks < NKC / 2; ++ks) {
                                const uint64_t adesc = tc::umma_desc(a_addr + (uint32_t)(2 * ks * kTcPlaneBytes + pos * 16), kTcPlaneBytes, kTcBox * 16);
                                const uint64_t bdesc = tc::umma_desc(w_addr + (uint32_t)(buf * Cfg::W_BYTES + tap * (NKC * NOUT * 16) + 2 * ks * (NOUT * 16)),
                                                                     NOUT * 16, 128);
                                tc::umma_f16(d_tmem, adesc, bdesc, idesc, (tap | ks) != 0 ? 1u : 0u);
                            }
                        }
                    }
                    tc::umma_commit(w_empty + buf);     // weight slot reusable once these MMAs have read it
                }
                tc::umma_commit(a_empty);               // activation box reusable
                tc::umma_commit(acc_full);              // accumulators complete
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: 8 warps; warp e handles TMEM lanes 32*(e%4).. of MMA tile e/4 =====
        const int e = warp - 4, q = e & 3, m = e >> 2;
        const int row = 4 * q + (lane >> 3), col = 8 * m + (lane & 7);
        const size_t plane = (size_t)H * W;
        const float* __restrict__ res = p.res;
        float* __restrict__ out = p.out;
        float* __restrict__ out2 = p.out2;
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = (int)blockIdx.x + it * (int)gridDim.x;
            const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
            const int y = ty * kTcTile + row, x = tx * kTcTile + col;
            const bool valid = (y < H) && (x < W);
            const size_t pix = (size_t)y * W + x;
            tc::mbar_wait(acc_full, (uint32_t)(it & 1));
            tc::tc_fence_after();
            const uint32_t t0 = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(m * Cfg::ACC_COLS);

            auto emit = [&](const float (&v)[NOUT], int ch0, int cnt) {
                if (!valid) return;
#pragma unroll
                for (int j0 = 0; j0 < NOUT; j0 += 8) {
                    float rv[8];
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        rv[jj] = 0.f;
                        if (j0 + jj < cnt && res != nullptr) rv[jj] = __ldg(res + ((size_t)b * C + ch0 + j0 + jj) * plane + pix);
                    }
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        if (j0 + jj >= cnt) continue;
                        const int ch = ch0 + j0 + jj;
                        const float o = bn_prelu(v[j0 + jj] + rv[jj], sep[ch], sep[128 + ch], sep[256 + ch]);
                        if (out != nullptr) out[((size_t)b * C + ch) * plane + pix] = o;
                        if (out2 != nullptr)
                            out2[((size_t)b * p.C2 + p.c2_off + ch) * plane + pix] = bn_prelu(o, sep[384 + ch], sep[512 + ch], sep[640 + ch]);
                    }
                }
            };

            float acc[NOUT], v[NOUT];
            if constexpr (NOUT == 32) tc::tmem_ld32(t0, v); else tc::tmem_ld16(t0, v);
            emit(v, 0, CO1);
#pragma unroll 1
            for (int br = 1; br < 5; ++br) {
                if constexpr (NOUT == 32) tc::tmem_ld32(t0 + (uint32_t)(br * NOUT), v); else tc::tmem_ld16(t0 + (uint32_t)(br * NOUT), v);
                if (br == 1) {
#pragma unroll
                    for (int j = 0; j < NOUT; ++j) acc[j] = v[j];
                } else {
#pragma unroll
                    for (int j = 0; j < NOUT; ++j) acc[j] += v[j];
                }
                emit(acc, CO1 + (br - 1) * CO, CO);
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(acc_empty);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// Hardware self-test of the operand convention: one M = 128 tile, one tap shifted by (dy, dx), all K chunks.
//   D[l][n] = sum_{kc,j} A[kc][16 + dy + l/8][16 + dx + l%8][j] * Bw[kc][n][j]
// The activation region comes either from a pre-padded global copy through plain stores (+ fence.proxy.async)
// or through the same 5-D TMA box (with out-of-bounds zero fill) the real kernel uses.
// ------------------------------------------------------------------------------------------------
template <int NOUT>
__global__ void __launch_bounds__(128, 1) tc_selftest_kernel(const __grid_constant__ CUtensorMap tmap, const __half* __restrict__ a_padded,
                                                             const __half* __restrict__ bw, float* __restrict__ dout, int nkc, int dy, int dx,
                                                             int use_tma, int ox, int oy) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* abuf = smem_raw;
    uint8_t* wbuf = abuf + 4 * kTcPlaneBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(wbuf + 4 * NOUT * 16);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        tc::mbar_init(bars + 0, 1); tc::mbar_init(bars + 1, 1);
        tc::mbar_fence_init();
    }
    if (warp == 0) tc::tmem_alloc(tmem_slot, 32);
    // weights: [nkc][NOUT][8] halves, plain copy
    for (int i = tid; i < nkc * NOUT * 2; i += 128) reinterpret_cast<uint64_t*>(wbuf)[i] = reinterpret_cast<const uint64_t*>(bw)[i];
    if (!use_tma)
        for (int i = tid; i < nkc * kTcPlaneBytes / 16; i += 128) reinterpret_cast<uint4*>(abuf)[i] = reinterpret_cast<const uint4*>(a_padded)[i];
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) {
        if (use_tma) {
            tc::mbar_expect_tx(bars + 0, (uint32_t)(nkc * kTcPlaneBytes));
            tc::tma_load_5d(abuf, &tmap, bars + 0, 0, ox - kTcHalo, oy - kTcHalo, 0, 0);
            tc::mbar_wait(bars + 0, 0);
        }
        tc::tc_fence_after();
        constexpr uint32_t idesc = tc::umma_idesc_f16(NOUT);
        const int pos = (kTcHalo + dy) * kTcBox + kTcHalo + dx;
        for (int ks = 0; ks < nkc / 2; ++ks) {
            const uint64_t adesc = tc::umma_desc(tc::smem_addr(abuf) + (uint32_t)(2 * ks * kTcPlaneBytes + pos * 16), kTcPlaneBytes, kTcBox * 16);
            const uint64_t bdesc = tc::umma_desc(tc::smem_addr(wbuf) + (uint32_t)(2 * ks * NOUT * 16), NOUT * 16, 128);
            tc::umma_f16(tmem_base, adesc, bdesc, idesc, ks != 0 ? 1u : 0u);
        }
        tc::umma_commit(bars + 1);
    }
    tc::mbar_wait(bars + 1, 0);
    tc::tc_fence_after();
    float v[NOUT];
    const uint32_t t0 = tmem_base + ((uint32_t)(32 * warp) << 16);
    if constexpr (NOUT == 32) tc::tmem_ld32(t0, v); else tc::tmem_ld16(t0, v);
#pragma unroll
    for (int n = 0; n < NOUT; ++n) dout[(size_t)(32 * warp + lane) * NOUT + n] = v[n];
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 32);
}

}  // namespace espnet
