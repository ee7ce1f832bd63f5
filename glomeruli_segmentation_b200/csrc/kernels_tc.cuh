// tcgen05 / TMEM kernels of the ESP "split-transform-merge" stage (sm_100a): fp16 operands, fp32 accumulation
// in tensor memory, fp32 epilogue.  ESPNET_MODE_F16TC of include/espnet_b200.h.
//
// Reference semantics: DilatedParllelResidualBlockB.forward (Model.py:187-214) and DownSamplerB.forward
// (Model.py:144-160) after the c1 reduce: d_k = Conv3x3(dilation k, pad k)(o1), HFF prefix sums, concat,
// residual add before BN, BN(eval, eps 1e-3) folded to scale/shift, PReLU.
//
// This file: the fp16 "chunk-plane" producers (reduce kernels) and the hardware self-test of the operand convention;
// the branch-stage kernel itself is in kernels_tc_branch.cuh.
#pragma once
#include "kernels_fp32.cuh"
#include "tc_common.cuh"
#include "kernels_tc_branch.cuh"
#include "kernels_tc_reduce.cuh"

namespace espnet {

inline bool tc_path_available() { return true; }

constexpr int kTcTile = 16;                        // output tile edge (pixels)
constexpr int kTcHalo = 16;                        // largest dilation
constexpr int kTcBox = kTcTile + 2 * kTcHalo;      // 48: staged region edge
constexpr int kTcPlaneBytes = kTcBox * kTcBox * 16;

// ------------------------------------------------------------------------------------------------
// fp16 chunk-plane stores for the reduce kernels: o1h[((b*NKC + kc)*HW + pix)*8 + j] = half(acc[kc*8+j])
// ------------------------------------------------------------------------------------------------
// SPLIT: the hi / lo pair (fp16(a/4) and its fp16 remainder) as crops [0,B) and [B,2B) of one tensor (fp32-equivalent path)
template <int CO, int NKC, bool SPLIT = false>
__device__ __forceinline__ void store_o1_f16(__half* __restrict__ o1h, int B, int b, size_t HW, size_t pix, const float (&acc)[CO]) {
    float v[NKC * 8];
#pragma unroll
    for (int c = 0; c < NKC * 8; ++c) v[c] = c < CO ? acc[c < CO ? c : 0] : 0.f;
    store_o1_chunks<NKC * 8, NKC, SPLIT>(o1h, B, b, HW, pix, v);
}

// ESP reduce (Model.py:178,192): 1x1 conv CIN -> CO on CUDA cores in fp32, result rounded once to fp16.
template <int CIN, int CO, int NKC, bool SPLIT>
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
            if (ok[r]) store_o1_f16<CO, NKC, SPLIT>(o1h, B, b, (size_t)HW, (size_t)(p0 + 32 * r), acc[r]);
    }
}

// DownSamplerB reduce (Model.py:135,145): 3x3 stride-2 conv CIN -> CO, fp32 CUDA cores, fp16 chunk-plane output.
template <int CIN, int CO, int NKC, bool SPLIT>
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
                if (y0 + r < Ho) store_o1_f16<CO, NKC, SPLIT>(o1h, B, b, (size_t)Ho * Wo, (size_t)(y0 + r) * Wo + x, acc[r]);
        }
    }
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
            tc::tma_load_4d(abuf, &tmap, bars + 0, 4 * (ox - kTcHalo), oy - kTcHalo, 0, 0);
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
