// tcgen05 kernel of the ESP block's 1x1 reduce (Model.py:178,192: c1 = C(nIn, n, 1, 1)), ESPNET_MODE_F16TC.
//
// o1[px][co] = sum_ci x[px][ci] * W1[co][ci] as a plain GEMM: M = 128 consecutive pixels of one crop, N = NOUT
// (12 -> 16, 25 -> 32 output channels), K = CIN (64 / 128).  The input is the previous block's planar fp32 output; the
// loader warps read it coalesced (lanes = consecutive pixels of one channel plane), round once to fp16 and store it as
// the K-major chunk-plane A operand [kc][128 px][8 ch] (tc_common.cuh) -- the transpose is free because a thread owns a
// pixel and writes one 16 B chunk per 8 channels.  The result leaves TMEM as fp16 chunk-plane o1h [B][kc][HW][8], the
// layout the branch kernel's TMA box reads.  HBM-bound: 4*CIN + 2*8*NKC bytes per pixel.
//   warps 0..15 = loaders in TWO groups of 8 (pixel quarter x channel half): group g owns A stage g and the tiles of
//   parity g, so two tiles' worth of global loads (2 x 64 KB at level 3) are in flight per SM -- one group's load latency
//   hides the other's convert + store phase; warps 16..19 = epilogue, warp 20 = MMA issuer, warp 21 = TMEM allocator;
//   accumulators are two tiles deep.
#pragma once
#include "kernels_fp32.cuh"
#include "tc_common.cuh"

namespace espnet {

constexpr int kRedThreads = 704;

// SPLIT: fp32-equivalent variant, both operands as 3-term fp16 splits (see kernels_tc_branch.cuh): the A stage holds
// [hi planes | lo planes], the weights [hi | lo], three MMAs per K step, and the result is written as the hi / lo
// chunk-plane pair (crops [0,B) and [B,2B) of one tensor) the split branch kernel reads.
constexpr float kSplitScaleA = 0.25f;   // activations are stored as fp16(a/4) (+ remainder), weights as fp16(4w) (+ remainder)

__device__ __forceinline__ void split_f16x2(float a, float b, __half2& hi, __half2& lo) {
    const float sa = a * kSplitScaleA, sb = b * kSplitScaleA;
    hi = __floats2half2_rn(sa, sb);
    const float2 hf = __half22float2(hi);
    lo = __floats2half2_rn(sa - hf.x, sb - hf.y);
}
// the same with the scale and the remainder as packed fp32x2 instructions (FMUL2 / FADD2, bit-identical): 6 instead of 8
// instructions per pair where the two inputs already sit in an aligned register pair (the 1x1 reduce loaders: measured
// 0.45 -> 0.43 ms; in the 3x3-s2 loaders the pairing costs moves and is slower)
__device__ __forceinline__ void split_f16x2_packed(float a, float b, __half2& hi, __half2& lo) {
    unsigned long long v, s, h, d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a), "f"(b));
    asm("mov.b64 %0, {%1, %1};" : "=l"(s) : "f"(kSplitScaleA));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(v) : "l"(v), "l"(s));
    float sa, sb;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(sa), "=f"(sb) : "l"(v));
    hi = __floats2half2_rn(sa, sb);
    const float2 hf = __half22float2(hi);
    asm("mov.b64 %0, {%1, %2};" : "=l"(h) : "f"(hf.x), "f"(hf.y));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(v), "l"(h));
    float da, db;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(da), "=f"(db) : "l"(d));
    lo = __floats2half2_rn(da, db);
}

template <int CIN, int NOUT, bool SPLIT = false>
struct ReduceTcCfg {
    static constexpr int KC = CIN / 8;                 // input K chunks
    static constexpr int A_PART = KC * 128 * 16;       // bytes of one (hi or lo) operand copy
    static constexpr int A_STAGE = (SPLIT ? 2 : 1) * A_PART;
    static constexpr int W_PART = KC * NOUT * 16;
    static constexpr int W_BYTES = (SPLIT ? 2 : 1) * W_PART;
    static constexpr size_t SMEM = 1024 + 2 * (size_t)A_STAGE + W_BYTES + 256;
};

// accumulator row (NOUT fp32) -> fp16 chunk-plane o1h [B][kc][HW][8], or the hi / lo pair [2B][kc][HW][8] when SPLIT
template <int NOUT, int NKC, bool SPLIT>
__device__ __forceinline__ void store_o1_chunks(__half* __restrict__ o1h, int B, int b, size_t HW, size_t pix, const float (&v)[NOUT]) {
#pragma unroll
    for (int kc = 0; kc < NKC; ++kc) {
        __half2 h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if constexpr (SPLIT) split_f16x2(v[8 * kc + 2 * j], v[8 * kc + 2 * j + 1], h[j], l[j]);
            else h[j] = __floats2half2_rn(v[8 * kc + 2 * j], v[8 * kc + 2 * j + 1]);
        }
        uint4 u;
        u.x = *reinterpret_cast<uint32_t*>(&h[0]); u.y = *reinterpret_cast<uint32_t*>(&h[1]);
        u.z = *reinterpret_cast<uint32_t*>(&h[2]); u.w = *reinterpret_cast<uint32_t*>(&h[3]);
        *reinterpret_cast<uint4*>(o1h + (((size_t)b * NKC + kc) * HW + pix) * 8) = u;
        if constexpr (SPLIT) {
            u.x = *reinterpret_cast<uint32_t*>(&l[0]); u.y = *reinterpret_cast<uint32_t*>(&l[1]);
            u.z = *reinterpret_cast<uint32_t*>(&l[2]); u.w = *reinterpret_cast<uint32_t*>(&l[3]);
            *reinterpret_cast<uint4*>(o1h + (((size_t)(B + b) * NKC + kc) * HW + pix) * 8) = u;
        }
    }
}

// w: [CIN/8][NOUT][8] fp16 (SPLIT: hi copy then lo copy), element (kc, n, j) = W1[co = n][ci = 8 kc + j] (zero for n >= CO)
template <int CIN, int NOUT, int NKC, bool SPLIT>
__global__ void __launch_bounds__(kRedThreads, 1) reduce1x1_tc_kernel(const float* __restrict__ in, const __half* __restrict__ w,
                                                                      __half* __restrict__ o1h, int B, int HW, int reverse) {
    using Cfg = ReduceTcCfg<CIN, NOUT, SPLIT>;
    constexpr int KC = Cfg::KC;
    static_assert(CIN % 16 == 0 && NKC * 8 == NOUT, "shapes");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* abuf = smem_raw;
    uint8_t* wbuf = abuf + 2 * Cfg::A_STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(wbuf + Cfg::W_BYTES);
    uint64_t* a_full = bars + 0;      // [2] 8 loader warps arrive
    uint64_t* a_empty = bars + 2;     // [2] MMA commit
    uint64_t* acc_full = bars + 4;    // [2]
    uint64_t* acc_empty = bars + 6;   // [2] 4 epilogue warps arrive
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunks = (HW + 127) / 128;
    const int total_tiles = B * chunks;
    const int my_tiles = ((int)blockIdx.x < total_tiles) ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    pdl_trigger();
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(a_full + s, 8); tc::mbar_init(a_empty + s, 1);
            tc::mbar_init(acc_full + s, 1); tc::mbar_init(acc_empty + s, 4);
        }
        tc::mbar_fence_init();
    }
    if (warp == 21) tc::tmem_alloc(tmem_slot, 64);
    for (int i = tid; i < Cfg::W_BYTES / 16; i += kRedThreads) reinterpret_cast<uint4*>(wbuf)[i] = __ldg(reinterpret_cast<const uint4*>(w) + i);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();   // everything above touched only weights / barriers / TMEM; the input of the kernel before is read below

    if (warp == 20) {
        // ===== MMA issuer: converged warp, one elected lane issues (tc_common.cuh: elect_one) =====
        {
            constexpr uint32_t idesc = tc::umma_idesc_f16(NOUT);
            constexpr uint32_t a_hi = (uint32_t)(128 >> 4) | (1u << 14);
            constexpr uint32_t b_hi = (uint32_t)(128 >> 4) | (1u << 14);
            const uint32_t a_lo0 = (tc::smem_addr(abuf) >> 4) + ((uint32_t)(2048 >> 4) << 16);
            const uint32_t b_lo0 = (tc::smem_addr(wbuf) >> 4) + ((uint32_t)((NOUT * 16) >> 4) << 16);
            for (int it = 0; it < my_tiles; ++it) {
                const int s = it & 1;
                tc::mbar_wait(acc_empty + s, (uint32_t)(((it >> 1) & 1) ^ 1));
                tc::mbar_wait(a_full + s, (uint32_t)((it >> 1) & 1));
                __syncwarp();
                tc::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(s * 32);
                if (tc::elect_one()) {
#pragma unroll
                for (int ks = 0; ks < KC / 2; ++ks) {
                    const uint32_t al = a_lo0 + (uint32_t)(s * (Cfg::A_STAGE >> 4) + 2 * ks * (2048 >> 4));
                    const uint32_t bl = b_lo0 + (uint32_t)(2 * ks * NOUT);
                    tc::umma_f16(d_tmem, ((uint64_t)a_hi << 32) | al, ((uint64_t)b_hi << 32) | bl, idesc, ks != 0 ? 1u : 0u);
                    if constexpr (SPLIT) {
                        tc::umma_f16(d_tmem, ((uint64_t)a_hi << 32) | (al + (uint32_t)(Cfg::A_PART >> 4)), ((uint64_t)b_hi << 32) | bl, idesc, 1u);   // lo x W_hi
                        tc::umma_f16(d_tmem, ((uint64_t)a_hi << 32) | al, ((uint64_t)b_hi << 32) | (bl + (uint32_t)(Cfg::W_PART >> 4)), idesc, 1u);   // hi x W_lo
                    }
                }
                tc::umma_commit(a_empty + s);
                tc::umma_commit(acc_full + s);
                }
                __syncwarp();
            }
        }
    } else if (warp < 16) {
        // ===== loaders: group = warp / 8 (tiles of that parity), warp lw of the group -> pixels 32*(lw%4) + lane,
        // K chunks [(lw/4) * KC/2, +KC/2) =====
        const int lw = warp & 7, px = 32 * (lw & 3) + lane, kc0 = (lw >> 2) * (KC / 2);
        const uint32_t plane_b = (uint32_t)HW * 4u;
        for (int it = warp >> 3; it < my_tiles; it += 2) {
            const int t_lin = (int)blockIdx.x + it * (int)gridDim.x;
            const int tile = reverse ? total_tiles - 1 - t_lin : t_lin;     // see the launcher: walk against the producer's order
            const int b = tile / chunks, p = (tile % chunks) * 128 + px;
            const bool ok = p < HW;
            const char* src = reinterpret_cast<const char*>(in + (size_t)b * CIN * HW + (ok ? p : 0));
            const int s = it & 1;
            uint8_t* dst = abuf + s * Cfg::A_STAGE + px * 16;
            // issue this tile's global loads before waiting for the smem slot: the slot wait is hidden behind them
            float v[KC / 2][8];
#pragma unroll
            for (int k = 0; k < KC / 2; ++k)
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    v[k][j] = ok ? __ldg(reinterpret_cast<const float*>(src + (uint64_t)plane_b * (uint32_t)(8 * (kc0 + k) + j))) : 0.f;
            tc::mbar_wait(a_empty + s, (uint32_t)(((it >> 1) & 1) ^ 1));
#pragma unroll
            for (int k = 0; k < KC / 2; ++k) {
                __half2 h[4], l[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if constexpr (SPLIT) split_f16x2_packed(v[k][2 * j], v[k][2 * j + 1], h[j], l[j]);
                    else h[j] = __floats2half2_rn(v[k][2 * j], v[k][2 * j + 1]);
                }
                uint4 u;
                u.x = *reinterpret_cast<uint32_t*>(&h[0]); u.y = *reinterpret_cast<uint32_t*>(&h[1]);
                u.z = *reinterpret_cast<uint32_t*>(&h[2]); u.w = *reinterpret_cast<uint32_t*>(&h[3]);
                *reinterpret_cast<uint4*>(dst + (kc0 + k) * 2048) = u;
                if constexpr (SPLIT) {
                    u.x = *reinterpret_cast<uint32_t*>(&l[0]); u.y = *reinterpret_cast<uint32_t*>(&l[1]);
                    u.z = *reinterpret_cast<uint32_t*>(&l[2]); u.w = *reinterpret_cast<uint32_t*>(&l[3]);
                    *reinterpret_cast<uint4*>(dst + Cfg::A_PART + (kc0 + k) * 2048) = u;
                }
            }
            tc::fence_proxy_async();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(a_full + s);
        }
    } else if (warp < 20) {
        // ===== epilogue: TMEM -> fp16 chunk-plane o1h =====
        const int q = warp & 3, px = 32 * q + lane;
        for (int it = 0; it < my_tiles; ++it) {
            const int t_lin = (int)blockIdx.x + it * (int)gridDim.x;
            const int tile = reverse ? total_tiles - 1 - t_lin : t_lin;
            const int b = tile / chunks, p = (tile % chunks) * 128 + px;
            const int s = it & 1;
            tc::mbar_wait_backoff(acc_full + s, (uint32_t)((it >> 1) & 1));
            tc::tc_fence_after();
            float v[NOUT];
            const uint32_t t0 = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(s * 32);
            if constexpr (NOUT == 32) tc::tmem_ld32(t0, v); else tc::tmem_ld16(t0, v);
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(acc_empty + s);
            if (p < HW) store_o1_chunks<NOUT, NKC, SPLIT>(o1h, B, b, (size_t)HW, (size_t)p, v);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 21) tc::tmem_dealloc(tmem_base, 64);
}

}  // namespace espnet
