// S10, register-tiled: `conv` CBR 3x3 (NC+19)->NC on cat[comb, output0_cat] (Model.py:332,375) + classifier ConvT k2 s2
// (Model.py:339,377) -> logits [B,NC,H,W] with the fused arg-max (VisualizeResults_iou.py:128) and softmax-accumulate
// (ensemble extension) epilogues.  Same arithmetic as dec_c_kernel (kernels_fp32.cuh) but one thread owns FOUR consecutive
// half-resolution pixels: per (input channel, tap row) it loads 6 activations (one 16 B vector + the two neighbours) for
// 3 x 4 x NC FMAs instead of one load per NC FMAs, and writes 8 consecutive full-resolution pixels per row (two 16 B logit
// stores per class, one 8 B mask store).
#pragma once
#include "kernels_fp32.cuh"

namespace espnet {

// UNR = input channels whose loads are in flight per thread (2 for large launches, 8 for small, latency-bound ones; same order).
template <int NC, int UNR>
__global__ void __launch_bounds__(256) dec_c4_kernel(const DecCParams<NC> p) {
    constexpr int CI = NC + 19;
    // weights of one (input channel, tap row): 3 kx x NC values padded to WR floats -> LDS.128 broadcasts
    constexpr int WR = (3 * NC + 3) & ~3;
    __shared__ __align__(16) float sw[CI * 3 * WR];
    __shared__ float swt[NC * NC * 4];
    __shared__ float sb[3 * NC];
    pdl_trigger();
    for (int i = threadIdx.x; i < CI * 3 * WR; i += 256) {
        const int r = i / WR, k = i - r * WR;
        sw[i] = k < 3 * NC ? p.w[r * 3 * NC + k] : 0.f;
    }
    for (int i = threadIdx.x; i < NC * NC * 4; i += 256) swt[i] = p.wt[i];
    for (int i = threadIdx.x; i < NC; i += 256) { sb[i] = p.s[i]; sb[NC + i] = p.t[i]; sb[2 * NC + i] = p.a[i]; }
    __syncthreads();
    pdl_wait();
    const int H2 = p.H2, W2 = p.W2;
    const int x0 = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4;
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int b = blockIdx.z;
    if (x0 >= W2 || y >= H2) return;
    const size_t plane = (size_t)H2 * W2;
    float acc[4][NC];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < NC; ++j) acc[q][j] = 0.f;
    const bool has_l = x0 > 0, has_r = x0 + 4 < W2;
#pragma unroll 1
    for (int ky = 0; ky < 3; ++ky) {
        const int yy = y + ky - 1;
        if (yy < 0 || yy >= H2) continue;                 // zero padding row (uniform per warp)
        const size_t roff = (size_t)yy * W2 + x0;
#pragma unroll UNR
        for (int ci = 0; ci < CI; ++ci) {
            const float* row = (ci < NC ? p.comb + ((size_t)b * NC + ci) * plane : p.out0cat + ((size_t)b * 19 + (ci - NC)) * plane) + roff;
            const float4 c = __ldg(reinterpret_cast<const float4*>(row));
            float in[6];
            in[0] = has_l ? __ldg(row - 1) : 0.f;
            in[1] = c.x; in[2] = c.y; in[3] = c.z; in[4] = c.w;
            in[5] = has_r ? __ldg(row + 4) : 0.f;
            float wr[WR];
#pragma unroll
            for (int k = 0; k < WR / 4; ++k) {
                const float4 w4 = *reinterpret_cast<const float4*>(sw + (ci * 3 + ky) * WR + 4 * k);
                wr[4 * k] = w4.x; wr[4 * k + 1] = w4.y; wr[4 * k + 2] = w4.z; wr[4 * k + 3] = w4.w;
            }
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
                for (int j = 0; j + 1 < NC; j += 2)       // class pairs: one packed FFMA2 per pixel
#pragma unroll
                    for (int q = 0; q < 4; ++q) ffma2(acc[q][j], acc[q][j + 1], in[q + kx], wr[kx * NC + j], wr[kx * NC + j + 1]);
                if (NC & 1) {
                    const float wv = wr[kx * NC + NC - 1];
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[q][NC - 1] = fmaf(in[q + kx], wv, acc[q][NC - 1]);
                }
            }
        }
    }
    const int W = 2 * W2;
    const size_t fplane = 4 * plane;
    // lg[r][o][8]: two output rows x classes x 8 consecutive full-resolution pixels
    float lg[2][NC][8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float v[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) v[j] = bn_prelu(acc[q][j], sb[j], sb[NC + j], sb[2 * NC + j]);
#pragma unroll
        for (int o = 0; o < NC; ++o) {
            float r00 = 0.f, r01 = 0.f, r10 = 0.f, r11 = 0.f;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const float* wq = swt + (c * NC + o) * 4;
                r00 = fmaf(v[c], wq[0], r00); r01 = fmaf(v[c], wq[1], r01);
                r10 = fmaf(v[c], wq[2], r10); r11 = fmaf(v[c], wq[3], r11);
            }
            lg[0][o][2 * q] = r00; lg[0][o][2 * q + 1] = r01;
            lg[1][o][2 * q] = r10; lg[1][o][2 * q + 1] = r11;
        }
    }
    const size_t base = (size_t)(2 * y) * W + 2 * x0;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        if (p.logits) {
#pragma unroll
            for (int o = 0; o < NC; ++o) {
                float* d = p.logits + ((size_t)b * NC + o) * fplane + base + (size_t)r * W;
                *reinterpret_cast<float4*>(d) = make_float4(lg[r][o][0], lg[r][o][1], lg[r][o][2], lg[r][o][3]);
                *reinterpret_cast<float4*>(d + 4) = make_float4(lg[r][o][4], lg[r][o][5], lg[r][o][6], lg[r][o][7]);
            }
        }
        unsigned char am[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float col[NC];
#pragma unroll
            for (int o = 0; o < NC; ++o) col[o] = lg[r][o][k];
            am[k] = (unsigned char)argmax_first<NC>(col);
        }
        if (p.prob_acc) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float m = lg[r][0][k];
#pragma unroll
                for (int o = 1; o < NC; ++o) m = fmaxf(m, lg[r][o][k]);
                float sum = 0.f;
#pragma unroll
                for (int o = 0; o < NC; ++o) { lg[r][o][k] = expf(lg[r][o][k] - m); sum += lg[r][o][k]; }
                const float inv = 1.f / sum;
#pragma unroll
                for (int o = 0; o < NC; ++o) lg[r][o][k] *= inv;
            }
#pragma unroll
            for (int o = 0; o < NC; ++o) {
                float* d = p.prob_acc + ((size_t)b * NC + o) * fplane + base + (size_t)r * W;
                if (!p.prob_init) {
                    const float4 t0 = *reinterpret_cast<const float4*>(d), t1 = *reinterpret_cast<const float4*>(d + 4);
                    lg[r][o][0] += t0.x; lg[r][o][1] += t0.y; lg[r][o][2] += t0.z; lg[r][o][3] += t0.w;
                    lg[r][o][4] += t1.x; lg[r][o][5] += t1.y; lg[r][o][6] += t1.z; lg[r][o][7] += t1.w;
                }
                *reinterpret_cast<float4*>(d) = make_float4(lg[r][o][0], lg[r][o][1], lg[r][o][2], lg[r][o][3]);
                *reinterpret_cast<float4*>(d + 4) = make_float4(lg[r][o][4], lg[r][o][5], lg[r][o][6], lg[r][o][7]);
            }
            if (p.mask_from_prob) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float col[NC];
#pragma unroll
                    for (int o = 0; o < NC; ++o) col[o] = lg[r][o][k];
                    am[k] = (unsigned char)argmax_first<NC>(col);
                }
            }
        }
        if (p.mask) {
            uint2 pk;
            pk.x = (unsigned)am[0] | ((unsigned)am[1] << 8) | ((unsigned)am[2] << 16) | ((unsigned)am[3] << 24);
            pk.y = (unsigned)am[4] | ((unsigned)am[5] << 8) | ((unsigned)am[6] << 16) | ((unsigned)am[7] << 24);
            *reinterpret_cast<uint2*>(p.mask + (size_t)b * fplane + base + (size_t)r * W) = pk;
        }
    }
}

}  // namespace espnet

namespace espnet {

// S7, 4 pixels per thread (plane % 4 == 0): b3 output -> encoder.classifier 1x1 (Model.py:271,302) [+ br BN + up_l3 ConvT for
// the full net, Model.py:331,334,370].  One 16 B load per channel plane feeds 4 x NC FMAs; same arithmetic as head3_kernel.
// UNR = channel planes whose loads are in flight per thread: 8 for large launches (bandwidth), 32 for small ones.  Small launches
// also use 64-thread blocks (blockDim.x is a run-time value here): at batch 1 the 1 024 threads of this kernel would otherwise sit
// on 4 SMs and pull the whole 4 MB input through 4 SMs' L2 ports.  The summation order per pixel never changes.
template <int NC, int UNR>
__global__ void __launch_bounds__(256) head3v_kernel(const Head3Params<NC> p) {
    __shared__ float sw[256 * NC];
    __shared__ float swt[NC * NC * 4];
    __shared__ float sbn[2 * NC];
    pdl_trigger();
    for (int i = threadIdx.x; i < 256 * NC; i += blockDim.x) sw[i] = p.w[i];
    if (p.up_out) {
        for (int i = threadIdx.x; i < NC * NC * 4; i += blockDim.x) swt[i] = p.wt[i];
        for (int i = threadIdx.x; i < NC; i += blockDim.x) { sbn[i] = p.bn_s[i]; sbn[NC + i] = p.bn_t[i]; }
    }
    __syncthreads();
    pdl_wait();
    const size_t plane = (size_t)p.H8 * p.W8;
    const size_t n4 = (size_t)p.B * plane / 4;
    for (size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += (size_t)gridDim.x * blockDim.x) {
        const size_t i = 4 * i4;
        const int b = (int)(i / plane);
        const size_t pix = i % plane;
        const float* src = p.in + (size_t)b * 256 * plane + pix;
        float acc[4][NC];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < NC; ++j) acc[q][j] = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < 256; c0 += UNR) {
            float4 a[UNR];                                 // all loads of a batch first, then the FMAs in channel order
#pragma unroll
            for (int u = 0; u < UNR; ++u) a[u] = __ldg(reinterpret_cast<const float4*>(src + (size_t)(c0 + u) * plane));
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
#pragma unroll
                for (int j = 0; j < NC; ++j) {
                    const float wv = sw[(c0 + u) * NC + j];
                    acc[0][j] = fmaf(a[u].x, wv, acc[0][j]); acc[1][j] = fmaf(a[u].y, wv, acc[1][j]);
                    acc[2][j] = fmaf(a[u].z, wv, acc[2][j]); acc[3][j] = fmaf(a[u].w, wv, acc[3][j]);
                }
            }
        }
        if (p.enc_out) {
#pragma unroll
            for (int j = 0; j < NC; ++j)
                *reinterpret_cast<float4*>(p.enc_out + ((size_t)b * NC + j) * plane + pix) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
        }
        if (p.up_out) {
            const int W4 = 2 * p.W8;
            const size_t plane4 = 4 * plane;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int y = (int)((pix + q) / p.W8), x = (int)((pix + q) % p.W8);
                float v[NC];
#pragma unroll
                for (int j = 0; j < NC; ++j) v[j] = fmaf(acc[q][j], sbn[j], sbn[NC + j]);
#pragma unroll
                for (int o = 0; o < NC; ++o) {
                    float r00 = 0.f, r01 = 0.f, r10 = 0.f, r11 = 0.f;
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const float* wq = swt + (c * NC + o) * 4;
                        r00 = fmaf(v[c], wq[0], r00); r01 = fmaf(v[c], wq[1], r01);
                        r10 = fmaf(v[c], wq[2], r10); r11 = fmaf(v[c], wq[3], r11);
                    }
                    float* d = p.up_out + ((size_t)b * NC + o) * plane4 + (size_t)(2 * y) * W4 + 2 * x;
                    *reinterpret_cast<float2*>(d) = make_float2(r00, r01);
                    *reinterpret_cast<float2*>(d + W4) = make_float2(r10, r11);
                }
            }
        }
    }
}

// S8 + first half of S9, 4 pixels per thread (plane % 4 == 0): level3_C 1x1 131->NC (Model.py:330,372), cat with up_l3's output
// (Model.py:373), combine_l2_l3[0] BR(2NC).  Same arithmetic as dec_a_kernel.
template <int NC, int UNR>
__global__ void __launch_bounds__(256) dec_av_kernel(const DecAParams<NC> p) {
    __shared__ float sw[131 * NC];
    __shared__ float sb[6 * NC];
    pdl_trigger();
    for (int i = threadIdx.x; i < 131 * NC; i += blockDim.x) sw[i] = p.w[i];
    for (int i = threadIdx.x; i < 2 * NC; i += blockDim.x) { sb[i] = p.s[i]; sb[2 * NC + i] = p.t[i]; sb[4 * NC + i] = p.a[i]; }
    __syncthreads();
    pdl_wait();
    const size_t plane = (size_t)p.H4 * p.W4;
    const size_t n4 = (size_t)p.B * plane / 4;
    for (size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += (size_t)gridDim.x * blockDim.x) {
        const size_t i = 4 * i4;
        const int b = (int)(i / plane);
        const size_t pix = i % plane;
        const float* src = p.out1cat + (size_t)b * 131 * plane + pix;
        float acc[4][NC];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < NC; ++j) acc[q][j] = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < 131; c0 += UNR) {
            float4 a[UNR];                                 // all loads of a batch first, then the FMAs in channel order
#pragma unroll
            for (int u = 0; u < UNR; ++u)
                a[u] = c0 + u < 131 ? __ldg(reinterpret_cast<const float4*>(src + (size_t)(c0 + u) * plane)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                if (c0 + u >= 131) break;
#pragma unroll
                for (int j = 0; j < NC; ++j) {
                    const float wv = sw[(c0 + u) * NC + j];
                    acc[0][j] = fmaf(a[u].x, wv, acc[0][j]); acc[1][j] = fmaf(a[u].y, wv, acc[1][j]);
                    acc[2][j] = fmaf(a[u].z, wv, acc[2][j]); acc[3][j] = fmaf(a[u].w, wv, acc[3][j]);
                }
            }
        }
        float* d = p.tout + (size_t)b * 2 * NC * plane + pix;
#pragma unroll
        for (int j = 0; j < NC; ++j)
            *reinterpret_cast<float4*>(d + (size_t)j * plane) =
                make_float4(bn_prelu(acc[0][j], sb[j], sb[2 * NC + j], sb[4 * NC + j]), bn_prelu(acc[1][j], sb[j], sb[2 * NC + j], sb[4 * NC + j]),
                            bn_prelu(acc[2][j], sb[j], sb[2 * NC + j], sb[4 * NC + j]), bn_prelu(acc[3][j], sb[j], sb[2 * NC + j], sb[4 * NC + j]));
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p.up3 + ((size_t)b * NC + j) * plane + pix));
            const float s = sb[NC + j], t = sb[3 * NC + j], a = sb[5 * NC + j];
            *reinterpret_cast<float4*>(d + (size_t)(NC + j) * plane) =
                make_float4(bn_prelu(v.x, s, t, a), bn_prelu(v.y, s, t, a), bn_prelu(v.z, s, t, a), bn_prelu(v.w, s, t, a));
        }
    }
}

// S9, register-tiled like dec_c4_kernel: combine_l2_l3 CBR 3x3 2NC->NC (Model.py:330,373) + up_l2 ConvT k2 s2 + BR
// (Model.py:335,374) -> comb [B,NC,H2,W2].  One thread owns FOUR consecutive quarter-resolution pixels: per (input channel,
// tap row) 3 loads (one 16 B vector + the two neighbours) and 4 LDS.128 weight broadcasts feed 12 NC FMAs (the one-pixel
// dec_b_kernel issues one scalar LDS per FMA and is LSU-bound), and each class row leaves as two 16 B stores.
template <int NC, int UNR>
__global__ void __launch_bounds__(256) dec_b4_kernel(const DecBParams<NC> p) {
    constexpr int CI = 2 * NC;
    constexpr int WR = (3 * NC + 3) & ~3;
    __shared__ __align__(16) float sw[CI * 3 * WR];
    __shared__ float swt[NC * NC * 4];
    __shared__ float sb[6 * NC];
    pdl_trigger();
    for (int i = threadIdx.x; i < CI * 3 * WR; i += 256) {
        const int r = i / WR, k = i - r * WR;
        sw[i] = k < 3 * NC ? p.w[r * 3 * NC + k] : 0.f;
    }
    for (int i = threadIdx.x; i < NC * NC * 4; i += 256) swt[i] = p.wt[i];
    for (int i = threadIdx.x; i < NC; i += 256) {
        sb[i] = p.s[i]; sb[NC + i] = p.t[i]; sb[2 * NC + i] = p.a[i];
        sb[3 * NC + i] = p.s2[i]; sb[4 * NC + i] = p.t2[i]; sb[5 * NC + i] = p.a2[i];
    }
    __syncthreads();
    pdl_wait();
    const int H4 = p.H4, W4 = p.W4;
    const int x0 = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4;
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int b = blockIdx.z;
    if (x0 >= W4 || y >= H4) return;
    const size_t plane = (size_t)H4 * W4;
    float acc[4][NC];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < NC; ++j) acc[q][j] = 0.f;
    const bool has_l = x0 > 0, has_r = x0 + 4 < W4;
#pragma unroll 1
    for (int ky = 0; ky < 3; ++ky) {
        const int yy = y + ky - 1;
        if (yy < 0 || yy >= H4) continue;                 // zero padding row (uniform per warp)
        const float* rowb = p.tin + (size_t)b * CI * plane + (size_t)yy * W4 + x0;
#pragma unroll UNR
        for (int ci = 0; ci < CI; ++ci) {
            const float* row = rowb + (size_t)ci * plane;
            const float4 c = __ldg(reinterpret_cast<const float4*>(row));
            float in[6];
            in[0] = has_l ? __ldg(row - 1) : 0.f;
            in[1] = c.x; in[2] = c.y; in[3] = c.z; in[4] = c.w;
            in[5] = has_r ? __ldg(row + 4) : 0.f;
            float wr[WR];
#pragma unroll
            for (int k = 0; k < WR / 4; ++k) {
                const float4 w4 = *reinterpret_cast<const float4*>(sw + (ci * 3 + ky) * WR + 4 * k);
                wr[4 * k] = w4.x; wr[4 * k + 1] = w4.y; wr[4 * k + 2] = w4.z; wr[4 * k + 3] = w4.w;
            }
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int j = 0; j < NC; ++j) {
                    const float wv = wr[kx * NC + j];
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[q][j] = fmaf(in[q + kx], wv, acc[q][j]);
                }
        }
    }
    const int W2 = 2 * W4;
    const size_t plane2 = 4 * plane;
    float v[4][NC];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < NC; ++j) v[q][j] = bn_prelu(acc[q][j], sb[j], sb[NC + j], sb[2 * NC + j]);
#pragma unroll
    for (int o = 0; o < NC; ++o) {
        float r0[8], r1[8];            // two output rows x 8 consecutive half-resolution pixels
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float r[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const float* wq = swt + (c * NC + o) * 4;
#pragma unroll
                for (int k = 0; k < 4; ++k) r[k] = fmaf(v[q][c], wq[k], r[k]);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) r[k] = bn_prelu(r[k], sb[3 * NC + o], sb[4 * NC + o], sb[5 * NC + o]);
            r0[2 * q] = r[0]; r0[2 * q + 1] = r[1];
            r1[2 * q] = r[2]; r1[2 * q + 1] = r[3];
        }
        float* d = p.comb + ((size_t)b * NC + o) * plane2 + (size_t)(2 * y) * W2 + 2 * x0;
        *reinterpret_cast<float4*>(d) = make_float4(r0[0], r0[1], r0[2], r0[3]);
        *reinterpret_cast<float4*>(d + 4) = make_float4(r0[4], r0[5], r0[6], r0[7]);
        *reinterpret_cast<float4*>(d + W2) = make_float4(r1[0], r1[1], r1[2], r1[3]);
        *reinterpret_cast<float4*>(d + W2 + 4) = make_float4(r1[4], r1[5], r1[6], r1[7]);
    }
}

}  // namespace espnet
