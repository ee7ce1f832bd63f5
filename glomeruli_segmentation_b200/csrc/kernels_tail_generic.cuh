// Heads / decoder tail for an ARBITRARY number of classes (Model.py:246,311 accept any int; the reference default is 20, the
// shipped checkpoints use 5).  The kernels in kernels_fp32.cuh / kernels_dec.cuh are compile-time specialised for classes = 5 and
// 20; these are the same stages with the class count as a run-time value `nc` (1..kMaxClasses): weights in dynamic shared
// memory, per-thread class vectors in (L1-resident) local arrays.  Same arithmetic and summation order as the specialised
// scalar kernels, so for nc = 5 the results are bit-identical to them (tests/test_gpu_parity.py).
//   g_head3_kernel   S7: encoder.classifier 1x1 (+ br BN + up_l3 ConvT)          Model.py:271,331,334
//   g_dec_a_kernel   S8: level3_C 1x1 + cat + combine_l2_l3[0] BR                  Model.py:330,372-373
//   g_dec_b_kernel   S9: combine_l2_l3[1] CBR 3x3 + up_l2 ConvT + BR               Model.py:335,337
//   g_dec_c_kernel   S10: conv CBR 3x3 + classifier ConvT + arg-max / softmax      Model.py:332,339,375-377
//   g_upsample8_argmax_kernel  ESPNet-C: x8 bilinear + arg-max                     VisualizeResults_iou.py:125-128,258-261
#pragma once
#include "kernels_fp32.cuh"

namespace espnet {

constexpr int kMaxClasses = 48;   // dynamic shared memory of g_dec_b_kernel: (22 nc^2 + 6 nc) floats <= 227 KB

__device__ __forceinline__ int g_argmax_first(const float* v, int nc) {
    int best = 0;
    float bv = v[0];
    for (int j = 1; j < nc; ++j)
        if (v[j] > bv) { bv = v[j]; best = j; }
    return best;
}

__global__ void __launch_bounds__(256) g_head3_kernel(const Head3Params<0> p, int nc) {
    extern __shared__ float g_smem[];
    float* sw = g_smem;                      // [256][nc]
    float* swt = sw + 256 * nc;              // [nc][nc][4]
    float* sbn = swt + nc * nc * 4;          // [2 nc]
    for (int i = threadIdx.x; i < 256 * nc; i += 256) sw[i] = p.w[i];
    if (p.up_out) {
        for (int i = threadIdx.x; i < nc * nc * 4; i += 256) swt[i] = p.wt[i];
        for (int i = threadIdx.x; i < nc; i += 256) { sbn[i] = p.bn_s[i]; sbn[nc + i] = p.bn_t[i]; }
    }
    __syncthreads();
    const size_t plane = (size_t)p.H8 * p.W8;
    const size_t n = (size_t)p.B * plane;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const int b = (int)(i / plane);
        const size_t pix = i % plane;
        const float* src = p.in + (size_t)b * 256 * plane + pix;
        float acc[kMaxClasses];
        for (int j = 0; j < nc; ++j) acc[j] = 0.f;
        for (int ci = 0; ci < 256; ++ci) {
            const float a = __ldg(src + (size_t)ci * plane);
            for (int j = 0; j < nc; ++j) acc[j] = fmaf(a, sw[ci * nc + j], acc[j]);
        }
        if (p.enc_out)
            for (int j = 0; j < nc; ++j) p.enc_out[((size_t)b * nc + j) * plane + pix] = acc[j];
        if (p.up_out) {
            const int y = (int)(pix / p.W8), x = (int)(pix % p.W8);
            const int W4 = 2 * p.W8;
            const size_t plane4 = 4 * plane;
            for (int j = 0; j < nc; ++j) acc[j] = fmaf(acc[j], sbn[j], sbn[nc + j]);
            for (int o = 0; o < nc; ++o) {
                float r00 = 0.f, r01 = 0.f, r10 = 0.f, r11 = 0.f;
                for (int c = 0; c < nc; ++c) {
                    const float* wq = swt + (c * nc + o) * 4;
                    r00 = fmaf(acc[c], wq[0], r00); r01 = fmaf(acc[c], wq[1], r01);
                    r10 = fmaf(acc[c], wq[2], r10); r11 = fmaf(acc[c], wq[3], r11);
                }
                float* d = p.up_out + ((size_t)b * nc + o) * plane4 + (size_t)(2 * y) * W4 + 2 * x;
                d[0] = r00; d[1] = r01; d[W4] = r10; d[W4 + 1] = r11;
            }
        }
    }
}

__global__ void __launch_bounds__(256) g_dec_a_kernel(const DecAParams<0> p, int nc) {
    extern __shared__ float g_smem[];
    float* sw = g_smem;                      // [131][nc]
    float* sb = sw + 131 * nc;               // [6 nc]
    for (int i = threadIdx.x; i < 131 * nc; i += 256) sw[i] = p.w[i];
    for (int i = threadIdx.x; i < 2 * nc; i += 256) { sb[i] = p.s[i]; sb[2 * nc + i] = p.t[i]; sb[4 * nc + i] = p.a[i]; }
    __syncthreads();
    const size_t plane = (size_t)p.H4 * p.W4;
    const size_t n = (size_t)p.B * plane;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const int b = (int)(i / plane);
        const size_t pix = i % plane;
        const float* src = p.out1cat + (size_t)b * 131 * plane + pix;
        float acc[kMaxClasses];
        for (int j = 0; j < nc; ++j) acc[j] = 0.f;
        for (int ci = 0; ci < 131; ++ci) {
            const float a = __ldg(src + (size_t)ci * plane);
            for (int j = 0; j < nc; ++j) acc[j] = fmaf(a, sw[ci * nc + j], acc[j]);
        }
        float* d = p.tout + (size_t)b * 2 * nc * plane + pix;
        for (int j = 0; j < nc; ++j) d[(size_t)j * plane] = bn_prelu(acc[j], sb[j], sb[2 * nc + j], sb[4 * nc + j]);
        for (int j = 0; j < nc; ++j) {
            const float v = __ldg(p.up3 + ((size_t)b * nc + j) * plane + pix);
            d[(size_t)(nc + j) * plane] = bn_prelu(v, sb[nc + j], sb[3 * nc + j], sb[5 * nc + j]);
        }
    }
}

__global__ void __launch_bounds__(256) g_dec_b_kernel(const DecBParams<0> p, int nc) {
    extern __shared__ float g_smem[];
    float* sw = g_smem;                      // [2 nc][9][nc]
    float* swt = sw + 2 * nc * 9 * nc;       // [nc][nc][4]
    float* sb = swt + nc * nc * 4;           // [6 nc]
    for (int i = threadIdx.x; i < 2 * nc * 9 * nc; i += 256) sw[i] = p.w[i];
    for (int i = threadIdx.x; i < nc * nc * 4; i += 256) swt[i] = p.wt[i];
    for (int i = threadIdx.x; i < nc; i += 256) {
        sb[i] = p.s[i]; sb[nc + i] = p.t[i]; sb[2 * nc + i] = p.a[i];
        sb[3 * nc + i] = p.s2[i]; sb[4 * nc + i] = p.t2[i]; sb[5 * nc + i] = p.a2[i];
    }
    __syncthreads();
    const int H4 = p.H4, W4 = p.W4;
    const size_t plane = (size_t)H4 * W4;
    const size_t n = (size_t)p.B * plane;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const int b = (int)(i / plane);
        const int pix = (int)(i % plane);
        const int y = pix / W4, x = pix % W4;
        const float* src = p.tin + (size_t)b * 2 * nc * plane;
        float acc[kMaxClasses];
        for (int j = 0; j < nc; ++j) acc[j] = 0.f;
        for (int tap = 0; tap < 9; ++tap) {
            const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
            if (yy < 0 || yy >= H4 || xx < 0 || xx >= W4) continue;
            const size_t o = (size_t)yy * W4 + xx;
            for (int ci = 0; ci < 2 * nc; ++ci) {
                const float a = __ldg(src + (size_t)ci * plane + o);
                const float* wr = sw + (ci * 9 + tap) * nc;
                for (int j = 0; j < nc; ++j) acc[j] = fmaf(a, wr[j], acc[j]);
            }
        }
        for (int j = 0; j < nc; ++j) acc[j] = bn_prelu(acc[j], sb[j], sb[nc + j], sb[2 * nc + j]);
        const int W2 = 2 * W4;
        const size_t plane2 = 4 * plane;
        for (int o = 0; o < nc; ++o) {
            float r[4] = {0.f, 0.f, 0.f, 0.f};
            for (int c = 0; c < nc; ++c) {
                const float* wq = swt + (c * nc + o) * 4;
#pragma unroll
                for (int q = 0; q < 4; ++q) r[q] = fmaf(acc[c], wq[q], r[q]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) r[q] = bn_prelu(r[q], sb[3 * nc + o], sb[4 * nc + o], sb[5 * nc + o]);
            float* d = p.comb + ((size_t)b * nc + o) * plane2 + (size_t)(2 * y) * W2 + 2 * x;
            d[0] = r[0]; d[1] = r[1]; d[W2] = r[2]; d[W2 + 1] = r[3];
        }
    }
}

__global__ void __launch_bounds__(256) g_dec_c_kernel(const DecCParams<0> p, int nc) {
    extern __shared__ float g_smem[];
    const int CI = nc + 19;
    float* sw = g_smem;                      // [CI][9][nc]
    float* swt = sw + CI * 9 * nc;           // [nc][nc][4]
    float* sb = swt + nc * nc * 4;           // [3 nc]
    for (int i = threadIdx.x; i < CI * 9 * nc; i += 256) sw[i] = p.w[i];
    for (int i = threadIdx.x; i < nc * nc * 4; i += 256) swt[i] = p.wt[i];
    for (int i = threadIdx.x; i < nc; i += 256) { sb[i] = p.s[i]; sb[nc + i] = p.t[i]; sb[2 * nc + i] = p.a[i]; }
    __syncthreads();
    const int H2 = p.H2, W2 = p.W2;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int b = blockIdx.z;
    if (x >= W2 || y >= H2) return;
    const size_t plane = (size_t)H2 * W2;
    float acc[kMaxClasses];
    for (int j = 0; j < nc; ++j) acc[j] = 0.f;
    for (int tap = 0; tap < 9; ++tap) {
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        if (yy < 0 || yy >= H2 || xx < 0 || xx >= W2) continue;
        const size_t o = (size_t)yy * W2 + xx;
        const float* s0 = p.comb + (size_t)b * nc * plane + o;
        for (int ci = 0; ci < nc; ++ci) {
            const float a = __ldg(s0 + (size_t)ci * plane);
            const float* wr = sw + (ci * 9 + tap) * nc;
            for (int j = 0; j < nc; ++j) acc[j] = fmaf(a, wr[j], acc[j]);
        }
        const float* s1 = p.out0cat + (size_t)b * 19 * plane + o;
        for (int ci = 0; ci < 19; ++ci) {
            const float a = __ldg(s1 + (size_t)ci * plane);
            const float* wr = sw + ((nc + ci) * 9 + tap) * nc;
            for (int j = 0; j < nc; ++j) acc[j] = fmaf(a, wr[j], acc[j]);
        }
    }
    for (int j = 0; j < nc; ++j) acc[j] = bn_prelu(acc[j], sb[j], sb[nc + j], sb[2 * nc + j]);
    const int W = 2 * W2;
    const size_t fplane = 4 * plane;
    const size_t base = (size_t)(2 * y) * W + 2 * x;
    const size_t qoff[4] = {0, 1, (size_t)W, (size_t)W + 1};
    // one output pixel of the 2x2 group at a time: logits -> (softmax accumulate) -> arg-max
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {
        float lg[kMaxClasses];
        for (int o = 0; o < nc; ++o) {
            float v = 0.f;
            for (int c = 0; c < nc; ++c) v = fmaf(acc[c], swt[(c * nc + o) * 4 + q], v);
            lg[o] = v;
        }
        if (p.logits)
            for (int o = 0; o < nc; ++o) p.logits[((size_t)b * nc + o) * fplane + base + qoff[q]] = lg[o];
        int am = g_argmax_first(lg, nc);
        if (p.prob_acc) {
            float m = lg[0];
            for (int o = 1; o < nc; ++o) m = fmaxf(m, lg[o]);
            float sum = 0.f;
            for (int o = 0; o < nc; ++o) { lg[o] = expf(lg[o] - m); sum += lg[o]; }
            const float inv = 1.f / sum;
            for (int o = 0; o < nc; ++o) {
                float* d = p.prob_acc + ((size_t)b * nc + o) * fplane + base + qoff[q];
                float pr = lg[o] * inv;
                if (!p.prob_init) pr += *d;
                *d = pr;
                lg[o] = pr;
            }
            if (p.mask_from_prob) am = g_argmax_first(lg, nc);
        }
        if (p.mask) p.mask[(size_t)b * fplane + base + qoff[q]] = (unsigned char)am;
    }
}

__global__ void __launch_bounds__(256) g_upsample8_argmax_kernel(const float* __restrict__ enc, int nc, int B, int H8, int W8,
                                                                 unsigned char* __restrict__ mask) {
    const int H = 8 * H8, W = 8 * W8;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int b = blockIdx.z;
    if (x >= W || y >= H) return;
    // area_pixel_compute_source_index(scale = 1/8, align_corners = False): src = (dst + 0.5) / 8 - 0.5, clamped at 0
    float sy = 0.125f * ((float)y + 0.5f) - 0.5f; sy = sy < 0.f ? 0.f : sy;
    const int y0 = (int)sy, y1 = y0 + (y0 < H8 - 1 ? 1 : 0);
    const float ly1 = sy - (float)y0, ly0 = 1.f - ly1;
    float sx = 0.125f * ((float)x + 0.5f) - 0.5f; sx = sx < 0.f ? 0.f : sx;
    const int x0 = (int)sx, x1 = x0 + (x0 < W8 - 1 ? 1 : 0);
    const float lx1 = sx - (float)x0, lx0 = 1.f - lx1;
    const size_t plane = (size_t)H8 * W8;
    int best = 0;
    float bv = 0.f;
    for (int c = 0; c < nc; ++c) {
        const float* s = enc + ((size_t)b * nc + c) * plane;
        const float v00 = __ldg(s + (size_t)y0 * W8 + x0), v01 = __ldg(s + (size_t)y0 * W8 + x1);
        const float v10 = __ldg(s + (size_t)y1 * W8 + x0), v11 = __ldg(s + (size_t)y1 * W8 + x1);
        const float v = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
        if (c == 0 || v > bv) { bv = v; best = c; }
    }
    mask[(size_t)b * H * W + (size_t)y * W + x] = (unsigned char)best;
}

}  // namespace espnet
