// TMA-staged fp32 ESP branch kernel (sm_100a).
//
// Same arithmetic and register tile as esp_branch_kernel (kernels_fp32.cuh) but the activations no longer
// come from per-thread global loads: one elected thread streams [CH channels] x 64 x 64 boxes of the reduced
// map o1 -- the CTA's 32x32 output tile plus the 16-pixel halo the d=16 branch needs -- into shared memory
// with cp.async.bulk.tensor (TMA), NS stages deep, completion tracked by mbarriers.  TMA zero-fills
// everything outside the map, so the conv zero padding costs no predicates, and the hot loop is pure
// LDS + FFMA: per (tap, channel) 4 conflict-free LDS.32 (lanes = consecutive x) + ceil(CO/4) broadcast LDS.128
// for 4*CO FFMA.  The CTA is persistent; the (tile, branch, channel-chunk) stage sequence is flattened so the
// pipeline never drains between tiles.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels_fp32.cuh"

namespace espnet {

constexpr int kTmaTile = 32;              // output tile edge
constexpr int kTmaHalo = 16;              // largest dilation
constexpr int kTmaBox = kTmaTile + 2 * kTmaHalo;   // 64

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

template <int N, int CO1, int CO, int CH, int NS>
struct BranchTmaCfg {
    static constexpr int C = CO1 + 4 * CO;
    static constexpr int CP1 = pad4(CO1), CP = pad4(CO);
    static constexpr int W1 = 9 * N * CP1, WC = 4 * 9 * N * CP;
    static constexpr int STAGE_FLOATS = CH * kTmaBox * kTmaBox;
    static constexpr int NCH = (N + CH - 1) / CH;        // channel chunks per branch
    static constexpr size_t SMEM = (size_t)(NS * STAGE_FLOATS + W1 + WC + 6 * C) * sizeof(float) + NS * sizeof(uint64_t) + 128;
};

// acc[r][j] += sum over the chunk's channels and the 9 taps, activations from the staged box.
template <int N, int CO>
__device__ __forceinline__ void tma_stage_compute(float (&acc)[kRows][CO], const float* __restrict__ box, int nc, int c0, int d,
                                                  int ly, int lx, int gy0, int H, const float* __restrict__ wsm) {
    constexpr int CP = pad4(CO);
#if ESPNET_TAPROLL == 1
    // rolled tap rows: the hot body is 3 (tap, channel) steps = ~370 instructions (~6 KB), it stays in the instruction
    // cache (the fully unrolled 9-tap body is ~18 KB per instantiation and ncu showed stall_no_instruction on top)
#pragma unroll 1
    for (int cc = 0; cc < nc; ++cc) {
        const float* bc = box + cc * (kTmaBox * kTmaBox) + (ly + kTmaHalo) * kTmaBox + lx + kTmaHalo;
        const float* wc = wsm + (size_t)(c0 + cc) * CP;
#pragma unroll 1
        for (int ky = 0; ky < 3; ++ky) {
            const int dy = (ky - 1) * d;
            if (gy0 + kRows - 1 + dy < 0 || gy0 + dy >= H) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float* s = bc + dy * kTmaBox + (kx - 1) * d;
                float a[kRows];
#pragma unroll
                for (int r = 0; r < kRows; ++r) a[r] = s[r * kTmaBox];
                fma_tile<CO>(acc, a, wc + (size_t)(ky * 3 + kx) * N * CP);
            }
        }
    }
    return;
#elif ESPNET_TAPROLL == 2
    // rolled tap rows, both channels of the stage inside the body: 6 steps = ~740 instructions (~12 KB)
#pragma unroll 1
    for (int ky = 0; ky < 3; ++ky) {
        const int dy = (ky - 1) * d;
        if (gy0 + kRows - 1 + dy < 0 || gy0 + dy >= H) continue;
#pragma unroll 2
        for (int cc = 0; cc < nc; ++cc) {
            const float* bc = box + cc * (kTmaBox * kTmaBox) + (ly + kTmaHalo) * kTmaBox + lx + kTmaHalo;
            const float* wc = wsm + (size_t)(c0 + cc) * CP;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float* s = bc + dy * kTmaBox + (kx - 1) * d;
                float a[kRows];
#pragma unroll
                for (int r = 0; r < kRows; ++r) a[r] = s[r * kTmaBox];
                fma_tile<CO>(acc, a, wc + (size_t)(ky * 3 + kx) * N * CP);
            }
        }
    }
    return;
#endif
#pragma unroll 1
    for (int cc = 0; cc < nc; ++cc) {
        const float* bc = box + cc * (kTmaBox * kTmaBox) + (ly + kTmaHalo) * kTmaBox + lx + kTmaHalo;
        const float* wc = wsm + (size_t)(c0 + cc) * CP;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int dy = (ky - 1) * d;
            // the warp's 4 rows all fall into the zero padding above / below the map: nothing to add
            if (gy0 + kRows - 1 + dy < 0 || gy0 + dy >= H) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float* s = bc + dy * kTmaBox + (kx - 1) * d;
                float a[kRows];
#pragma unroll
                for (int r = 0; r < kRows; ++r) a[r] = s[r * kTmaBox];
                fma_tile<CO>(acc, a, wc + (size_t)(ky * 3 + kx) * N * CP);
            }
        }
    }
}

template <int N, int CO1, int CO, int CH, int NS>
__global__ void __launch_bounds__(256, 1) esp_branch_tma_kernel(const __grid_constant__ CUtensorMap tmap, const BranchParams p) {
    using Cfg = BranchTmaCfg<N, CO1, CO, CH, NS>;
    constexpr int C = Cfg::C, CP = Cfg::CP, W1 = Cfg::W1, WC = Cfg::WC;
    constexpr int NCH = Cfg::NCH, SPT = 5 * NCH;
    constexpr uint32_t STAGE_BYTES = Cfg::STAGE_FLOATS * sizeof(float);

    // No integer casts on this pointer: everything derived from it must stay in the shared address space so
    // that the hot loop is LDS with 32-bit immediate-offset addressing (a generic LD costs 64-bit address math).
    extern __shared__ __align__(1024) float smem_f[];
    float* stages = smem_f;
    if ((smem_u32(smem_f) & 127u) != 0) __trap();   // TMA destinations need 128 B alignment
    float* sw1 = stages + NS * Cfg::STAGE_FLOATS;
    float* swc = sw1 + W1;
    float* sep = swc + WC;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sep + 6 * C);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = p.H, W = p.W;
    const int tiles_x = (W + kTmaTile - 1) / kTmaTile, tiles_y = (H + kTmaTile - 1) / kTmaTile;
    const int total_tiles = p.B * tiles_x * tiles_y;
    const int my_tiles = ((int)blockIdx.x < total_tiles) ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int total_stages = my_tiles * SPT;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(bars + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    copy_to_smem(sw1, p.w_d1, W1);
    copy_to_smem(swc, p.w_chain, WC);
    for (int i = tid; i < C; i += blockDim.x) {
        sep[i] = p.s[i]; sep[C + i] = p.t[i]; sep[2 * C + i] = p.a[i];
        if (p.out2) {
            sep[3 * C + i] = p.s2[p.c2_off + i]; sep[4 * C + i] = p.t2[p.c2_off + i]; sep[5 * C + i] = p.a2[p.c2_off + i];
        }
    }
    __syncthreads();

    // producer: stage g of this CTA = (tile g / SPT, branch (g % SPT) / NCH, chunk g % NCH)
    auto issue = [&](int g) {
        const int tk = g / SPT, c = (g % SPT) % NCH;
        const int tile = (int)blockIdx.x + tk * (int)gridDim.x;
        const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
        const int s = g % NS;
        mbar_expect_tx(bars + s, STAGE_BYTES);
        tma_load_4d(stages + (size_t)s * Cfg::STAGE_FLOATS, &tmap, bars + s, tx * kTmaTile - kTmaHalo, ty * kTmaTile - kTmaHalo, c * CH, b);
    };
    if (tid == 0)
        for (int g = 0; g < NS - 1 && g < total_stages; ++g) issue(g);

    const size_t plane = (size_t)H * W;
    const float* __restrict__ res = p.res;
    float* __restrict__ out = p.out;
    float* __restrict__ out2 = p.out2;
    const int ly = warp * kRows, lx = lane;

    int g = 0;
    for (int tk = 0; tk < my_tiles; ++tk) {
        const int tile = (int)blockIdx.x + tk * (int)gridDim.x;
        const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
        const int y0 = ty * kTmaTile + ly, x = tx * kTmaTile + lx;
        const size_t pix0 = (size_t)y0 * W + x;

        auto emit = [&](auto& acc, int ch0, auto cnt_tag) {
            constexpr int CNT = decltype(cnt_tag)::value;
            constexpr int CHN = 8;
            if (x >= W) return;
#pragma unroll
            for (int j0 = 0; j0 < CNT; j0 += CHN) {
                float rv[CHN][kRows];
#pragma unroll
                for (int jj = 0; jj < CHN; ++jj)
#pragma unroll
                    for (int r = 0; r < kRows; ++r) {
                        rv[jj][r] = 0.f;
                        if (j0 + jj < CNT && res != nullptr && y0 + r < H)
                            rv[jj][r] = __ldg(res + ((size_t)b * C + ch0 + j0 + jj) * plane + pix0 + (size_t)r * W);
                    }
#pragma unroll
                for (int jj = 0; jj < CHN; ++jj) {
                    if (j0 + jj >= CNT) continue;
                    const int ch = ch0 + j0 + jj;
                    const float s1 = sep[ch], t1 = sep[C + ch], a1 = sep[2 * C + ch];
                    float s2 = 0.f, t2 = 0.f, a2 = 0.f;
                    if (out2 != nullptr) { s2 = sep[3 * C + ch]; t2 = sep[4 * C + ch]; a2 = sep[5 * C + ch]; }
#pragma unroll
                    for (int r = 0; r < kRows; ++r) {
                        if (y0 + r >= H) continue;
                        const float v = bn_prelu(acc[r][j0 + jj] + rv[jj][r], s1, t1, a1);
                        if (out != nullptr) out[((size_t)b * C + ch) * plane + pix0 + (size_t)r * W] = v;
                        if (out2 != nullptr)
                            out2[((size_t)b * p.C2 + p.c2_off + ch) * plane + pix0 + (size_t)r * W] = bn_prelu(v, s2, t2, a2);
                    }
                }
            }
        };

        // one pipeline step: refill the stage consumed one step ago, wait for this step's box, compute, release
        auto step = [&](auto& acc, auto co_tag, int c, int d, const float* wsm) {
            constexpr int COx = decltype(co_tag)::value;
            if (tid == 0 && g + NS - 1 < total_stages) issue(g + NS - 1);
            mbar_wait(bars + (g % NS), (uint32_t)((g / NS) & 1));
            const int c0 = c * CH;
            const int nc = (N - c0) < CH ? (N - c0) : CH;
            tma_stage_compute<N, COx>(acc, stages + (size_t)(g % NS) * Cfg::STAGE_FLOATS, nc, c0, d, ly, lx, y0, H, wsm);
            __syncthreads();
            ++g;
        };

        {
            float acc[kRows][CO1];
#pragma unroll
            for (int r = 0; r < kRows; ++r)
#pragma unroll
                for (int j = 0; j < CO1; ++j) acc[r][j] = 0.f;
#pragma unroll 1
            for (int c = 0; c < NCH; ++c) step(acc, IntTag<CO1>(), c, 1, sw1);
            emit(acc, 0, IntTag<CO1>());
        }
        {
            float acc[kRows][CO];
#pragma unroll
            for (int r = 0; r < kRows; ++r)
#pragma unroll
                for (int j = 0; j < CO; ++j) acc[r][j] = 0.f;
#pragma unroll 1
            for (int br = 0; br < 4; ++br) {
#pragma unroll 1
                for (int c = 0; c < NCH; ++c) step(acc, IntTag<CO>(), c, 2 << br, swc + (size_t)br * 9 * N * CP);
                emit(acc, CO1 + br * CO, IntTag<CO>());
            }
        }
    }
}

}  // namespace espnet
