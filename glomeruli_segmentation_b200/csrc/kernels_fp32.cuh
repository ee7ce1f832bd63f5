// fp32 CUDA-core kernels of the ESPNet inference path (sm_100a).
//
// Layout: activations are planar fp32 [B][C][H][W] (the reference's NCHW, Model.py) so that the 32
// lanes of a warp walk consecutive x of one channel plane: every global access is a coalesced 128 B
// line, dilated taps are plain address shifts, and concat (Model.py:157,208,350,359,368) is a
// channel offset into a wider buffer.  Every heavy kernel uses the same register tile: one thread
// owns 4 pixels (4 consecutive rows at one column) x all output channels of a branch, activations
// come from global/L1 (1 coalesced load per pixel-row), weights are warp-broadcast LDS.128 from
// shared memory, so one (tap, input-channel) step is 4 LDG + ceil(CO/4) LDS for 4*CO FFMA.
// These kernels carry the 1e-3 logit bar (true fp32 FMA); they are FMA-bound, not HBM-bound
// (SURVEY.md 8(d) caveat).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace espnet {

constexpr int kRows = 4;  // pixels (rows) per thread in the register tile
#ifndef ESPNET_HEAVY_THREADS
#define ESPNET_HEAVY_THREADS 256
#endif
#ifndef ESPNET_PIPE
#define ESPNET_PIPE 0
#endif
#ifndef ESPNET_TAPROLL
#define ESPNET_TAPROLL 0
#endif
#ifndef ESPNET_G25
#define ESPNET_G25 5
#endif
#ifndef ESPNET_G12
#define ESPNET_G12 4
#endif
constexpr int kHeavyThreads = ESPNET_HEAVY_THREADS;   // reduce3x3 / branch kernels, 1 CTA per SM

template <int V>
struct IntTag { static constexpr int value = V; };

// Programmatic dependent launch.  The kernels of one forward are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization (espnet_api.cu: launch_k): a kernel's CTAs may become resident
// while the kernel before it in the stream is still draining.  Contract for every kernel launched that way:
//   * pdl_trigger() first thing (all threads), so that the kernel AFTER it may be scheduled as soon as every CTA of
//     this one has started;
//   * before pdl_wait() only immutable data is touched (packed weights, BN tables, tensor maps) and nothing is
//     written to global memory: barrier init, TMEM allocation, weight staging;
//   * pdl_wait() returns when the kernel before it has completed and its writes are visible; since that kernel
//     did not finish its own pdl_wait() before ITS predecessor completed, everything after the waits is ordered exactly
//     as in a plain stream.
// Both are no-ops in a kernel launched without the attribute.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ float bn_prelu(float v, float s, float t, float a) {
    v = fmaf(v, s, t);
    return v >= 0.f ? v : a * v;
}

// acc[r][j] += a[r] * w[j], j < CO; w is a 16 B aligned shared-memory row of round_up(CO,4) floats,
// identical for all lanes (broadcast).
template <int CO>
__device__ __forceinline__ void fma_tile(float (&acc)[kRows][CO], const float (&a)[kRows], const float* __restrict__ w) {
#pragma unroll
    for (int j4 = 0; j4 < CO / 4; ++j4) {
        const float4 wv = *reinterpret_cast<const float4*>(w + 4 * j4);
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            acc[r][4 * j4 + 0] = fmaf(a[r], wv.x, acc[r][4 * j4 + 0]);
            acc[r][4 * j4 + 1] = fmaf(a[r], wv.y, acc[r][4 * j4 + 1]);
            acc[r][4 * j4 + 2] = fmaf(a[r], wv.z, acc[r][4 * j4 + 2]);
            acc[r][4 * j4 + 3] = fmaf(a[r], wv.w, acc[r][4 * j4 + 3]);
        }
    }
#pragma unroll
    for (int j = (CO / 4) * 4; j < CO; ++j) {
        const float wv = w[j];
#pragma unroll
        for (int r = 0; r < kRows; ++r) acc[r][j] = fmaf(a[r], wv, acc[r][j]);
    }
}

__host__ __device__ constexpr int pad4(int n) { return (n + 3) & ~3; }

// d0 = fma(a, b0, d0), d1 = fma(a, b1, d1) as ONE packed instruction (sm_100 FFMA2 with a broadcast scalar operand): two
// IEEE fp32 FMAs, bit-identical to two fmaf, half the issue slots -- for the issue-bound CUDA-core convolutions.
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a, float b0, float b1) {
    unsigned long long d, av, bv;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(d0), "f"(d1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
    asm("mov.b64 %0, {%1, %2};" : "=l"(bv) : "f"(b0), "f"(b1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(av), "l"(bv));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}

__device__ __forceinline__ void copy_to_smem(float* dst, const float* __restrict__ src, int n) {
    // n is a multiple of 4 and both pointers are 16 B aligned (the packer guarantees it)
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = threadIdx.x; i < n / 4; i += blockDim.x) d4[i] = __ldg(s4 + i);
}

// ------------------------------------------------------------------------------------------------
// S1 stem: P0 normalise (VisualizeResults_iou.py:107-119) + level1 CBR 3x3 s2 (Model.py:253) +
// sample1 avg-pool (Model.py:254) + cat + b1 BR (Model.py:257,350)  ->  output0_cat [B,19,H/2,W/2]
// plus the raw pooled image inp1 [B,3,H/2,W/2] (sample2's second pool reads it).
// ------------------------------------------------------------------------------------------------
struct StemParams {
    const void* x;
    int in_fmt, B, H, W;
    float mean[3], stdv[3];
    const int32_t* origins;
    int slide_h, slide_w;
    const float* w1;                    // [27][16]: (ci*9 + ky*3 + kx) x co
    const float *l1_s, *l1_t, *l1_a;    // level1 BN scale/shift, PReLU slope (16)
    const float *b1_s, *b1_t, *b1_a;    // b1 (19)
    const float *b2_s, *b2_t, *b2_a;    // b2 (131): channels 128..130 = the twice-pooled image (Model.py:255,348,359)
    float* out0cat;
    float* out1cat;                     // [B,131,H/4,W/4]: this kernel writes channels 128..130
};

// Block = 32 x 16 outputs (thread: column lane, rows ty and ty + 8).  The 65 x 33 input region is staged ONCE in shared
// memory, already normalised (u8 inputs go through a 3 x 256 table built with the reference's three separate fp32
// roundings, so the 54 divisions per output of a direct evaluation disappear), split by column parity so that the
// stride-2 window reads of a warp are bank-conflict free.  Accumulation order (channel, then tap) is fixed.
//
// sample2's SECOND AvgPool2d(3,2,1) + b2's BR on cat channels 128..130 (Model.py:255,348,359) ride along: the block's 16 x 8
// quarter-resolution outputs need the once-pooled image on its 32 x 16 tile plus one halo row above and one halo column to
// the left, i.e. an input region that starts two rows / columns earlier (67 x 35 instead of 65 x 33).  The once-pooled tile
// goes through shared memory; the halo values are computed by the same 9-term sum in the same order as the neighbouring
// block computes them, so the result is bit-identical to pooling a stored tensor (which no longer exists in HBM).
constexpr int kStemTW = 32, kStemTH = 16;
constexpr int kStemRW = 2 * kStemTW + 3, kStemRH = 2 * kStemTH + 3;      // 67 x 35 input region, origin (2 x0 - 3, 2 y0 - 3)
constexpr int kStemPE = 34, kStemPO = 33;                               // even / odd column counts

template <int FMT>
__global__ void __launch_bounds__(256, 4) stem_kernel(const StemParams p) {
    __shared__ __align__(16) float sw[27 * 16];
    __shared__ float sp[3 * 16 + 3 * 19];
    __shared__ float lut[FMT == 0 ? 1 : 3 * 256];
    __shared__ float se[3][kStemRH][kStemPE];      // region columns 0, 2, 4, ..
    __shared__ float so[3][kStemRH][kStemPO];      // region columns 1, 3, 5, ..
    __shared__ float sp1[3][kStemTH + 1][kStemTW + 1];   // once-pooled image: [0] = halo row / column
    const int tid = threadIdx.x;
    pdl_trigger();
    for (int i = tid; i < 27 * 16; i += 256) sw[i] = p.w1[i];
    for (int i = tid; i < 16; i += 256) { sp[i] = p.l1_s[i]; sp[16 + i] = p.l1_t[i]; sp[32 + i] = p.l1_a[i]; }
    for (int i = tid; i < 19; i += 256) { sp[48 + i] = p.b1_s[i]; sp[67 + i] = p.b1_t[i]; sp[86 + i] = p.b1_a[i]; }
    if (FMT != 0) {
        // three separate fp32 roundings, exactly as numpy does them (VisualizeResults_iou.py:107-119)
        for (int i = tid; i < 3 * 256; i += 256) {
            const int c = i >> 8;
            lut[i] = __fdiv_rn(__fdiv_rn(__fsub_rn((float)(i & 255), p.mean[c]), p.stdv[c]), 255.f);
        }
    }
    __syncthreads();
    pdl_wait();
    const int H2 = p.H >> 1, W2 = p.W >> 1;
    const int b = blockIdx.z;
    const int ry0 = 2 * (int)blockIdx.y * kStemTH - 3, rx0 = 2 * (int)blockIdx.x * kStemTW - 3;   // region origin in the crop

    auto put = [&](int c, int r, int j, float v) {
        if (j & 1) so[c][r][j >> 1] = v; else se[c][r][j >> 1] = v;
    };
    if (FMT == 0) {
        // One warp per region row (rows w, w + 8, ..), lanes on columns lane + 32 k.  Everything that depends only on the
        // column -- bounds test, source offset, the parity-split destination -- is computed once per thread; the loops over
        // channel and row are fully unrolled so that row addresses are immediate offsets (the staging used to be a third of
        // the kernel's instructions).  Asynchronous 4-byte copies, all in flight at once; src-size 0 = zero fill = conv / pool
        // zero padding of the NORMALISED tensor (Model.py:20,230).
        const float* x = reinterpret_cast<const float*>(p.x);
        const int ln = tid & 31, wp = tid >> 5;
        const uint32_t pitch = (ln & 1) ? kStemPO * 4u : kStemPE * 4u;           // bytes per region row of this lane's parity
        const uint32_t dbase = (uint32_t)__cvta_generic_to_shared((ln & 1) ? &so[0][0][0] : &se[0][0][0]) + (uint32_t)(ln >> 1) * 4u;
        bool cok[3];
        int xoff[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int xx = rx0 + ln + 32 * k;
            cok[k] = ln + 32 * k < kStemRW && xx >= 0 && xx < p.W;
            xoff[k] = cok[k] ? xx : 0;
        }
        const size_t cplane = (size_t)p.H * p.W;
        const float* xb = x + (size_t)b * 3 * cplane;
#pragma unroll
        for (int i = 0; i < (kStemRH + 7) / 8; ++i) {
            const int r = wp + 8 * i;
            if (r >= kStemRH) break;                                             // warp-uniform
            const int yy = ry0 + r;
            const bool row_ok = yy >= 0 && yy < p.H;
            const float* xr = xb + (size_t)(row_ok ? yy : 0) * p.W;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const uint32_t drow = dbase + (uint32_t)(c * kStemRH + r) * pitch;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (32 * k >= kStemRW) continue;
                    if (32 * (k + 1) > kStemRW && ln + 32 * k >= kStemRW) continue;
                    const bool ok = row_ok && cok[k];
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(drow + 64u * k),
                                 "l"(ok ? xr + c * cplane + xoff[k] : x), "r"(ok ? 4 : 0) : "memory");
                }
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else {
        // u8 HWC: one task = one region pixel (3 consecutive bytes, a warp reads 96 contiguous bytes per load); all byte
        // loads of a thread are issued first (phase 1), then table look-ups and the parity-split stores (phase 2).  This phase
        // is bound by the latency of the byte loads, not by instruction count: a column-per-thread mapping with half the index
        // arithmetic but 12 instead of 10 pixels per thread (201 of 256 threads busy) was 12 % SLOWER (ncu: 226 -> 253 us).
        const unsigned char* x = reinterpret_cast<const unsigned char*>(p.x);
        long long ox = 0, oy = 0, pitch_px = p.W, rows = p.H, row0 = (long long)b * p.H;
        if (FMT == 2) { ox = p.origins[2 * b]; oy = p.origins[2 * b + 1]; pitch_px = p.slide_w; rows = p.slide_h; row0 = 0; }
        constexpr int kPix = kStemRH * kStemRW, kIters = (kPix + 255) / 256;
        unsigned int bgr[kIters];        // b | g << 8 | r << 16, bit 24 = pixel is inside the crop (else zero padding)
#pragma unroll
        for (int i = 0; i < kIters; ++i) {
            const int t = tid + 256 * i;
            const int r = t / kStemRW, j = t - r * kStemRW;
            const int yy = ry0 + r, xx = rx0 + j;
            unsigned int u = 0;
            if (t < kPix && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) {
                u = 1u << 24;
                const long long sy = oy + yy, sx = ox + xx;
                // outside the slide openslide pads with 0 (then normalised like any pixel)
                if (sy >= 0 && sy < rows && sx >= 0 && sx < pitch_px) {
                    const unsigned char* q = x + ((row0 + sy) * pitch_px + sx) * 3;
                    u |= (unsigned int)__ldg(q) | ((unsigned int)__ldg(q + 1) << 8) | ((unsigned int)__ldg(q + 2) << 16);
                }
            }
            bgr[i] = u;
        }
#pragma unroll
        for (int i = 0; i < kIters; ++i) {
            const int t = tid + 256 * i;
            if (t >= kPix) continue;
            const int r = t / kStemRW, j = t - r * kStemRW;
            const unsigned int u = bgr[i];
            const bool in_crop = (u >> 24) != 0;
#pragma unroll
            for (int c = 0; c < 3; ++c) put(c, r, j, in_crop ? lut[c * 256 + ((u >> (8 * c)) & 255u)] : 0.f);
        }
    }
    __syncthreads();

    const int lane = tid & 31, ty = tid >> 5;
    const int x2 = blockIdx.x * kStemTW + lane;
    const int y2a = blockIdx.y * kStemTH + ty;
    float acc[2][16];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[r][j] = 0.f;
    float pool[2][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int ky = t / 3, kx = t % 3;
            // region column 2 lane + 2 + kx: kx = 0, 2 even (index lane + 1, lane + 2), kx = 1 odd (index lane + 1)
            const float a0 = kx == 1 ? so[c][2 * ty + 2 + ky][lane + 1] : se[c][2 * ty + 2 + ky][lane + 1 + (kx >> 1)];
            const float a1 = kx == 1 ? so[c][2 * (ty + 8) + 2 + ky][lane + 1] : se[c][2 * (ty + 8) + 2 + ky][lane + 1 + (kx >> 1)];
            s0 += a0; s1 += a1;
            const float4* w4 = reinterpret_cast<const float4*>(sw + (c * 9 + t) * 16);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const float4 w = w4[j4];
                ffma2(acc[0][4 * j4 + 0], acc[0][4 * j4 + 1], a0, w.x, w.y);
                ffma2(acc[0][4 * j4 + 2], acc[0][4 * j4 + 3], a0, w.z, w.w);
                ffma2(acc[1][4 * j4 + 0], acc[1][4 * j4 + 1], a1, w.x, w.y);
                ffma2(acc[1][4 * j4 + 2], acc[1][4 * j4 + 3], a1, w.z, w.w);
            }
        }
        pool[0][c] = s0 / 9.f;                                            // count_include_pad: always / 9
        pool[1][c] = s1 / 9.f;
        sp1[c][1 + ty][1 + lane] = pool[0][c];
        sp1[c][1 + ty + 8][1 + lane] = pool[1][c];
    }
    // halo of the once-pooled tile: row y2 = y0 - 1 (33 values incl. the corner) and column x2 = x0 - 1 (16 values) per channel
    if (tid < 3 * 49) {
        const int c = tid / 49, hh = tid - 49 * c;
        const int hy = hh < 33 ? 0 : hh - 32, hx = hh < 33 ? hh : 0;
        const int y2 = (int)blockIdx.y * kStemTH - 1 + hy, xh = (int)blockIdx.x * kStemTW - 1 + hx;
        float s = 0.f;
        if (y2 >= 0 && xh >= 0) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int r = 2 * hy + t / 3, j = 2 * hx + t % 3;
                s += (j & 1) ? so[c][r][j >> 1] : se[c][r][j >> 1];
            }
            s = s / 9.f;
        }
        sp1[c][hy][hx] = s;
    }
    __syncthreads();
    // second pool + b2 BR: 3 channels x 8 x 16 quarter-resolution outputs per block; taps outside the half-resolution image are
    // skipped exactly like the zero padding of AvgPool2d (the divisor stays 9)
    if (tid < (kStemTH / 2) * (kStemTW / 2)) {
        const int H4 = H2 >> 1, W4 = W2 >> 1;
        const int yl = tid / (kStemTW / 2), xl = tid % (kStemTW / 2);
        const int y4 = (int)blockIdx.y * (kStemTH / 2) + yl, x4 = (int)blockIdx.x * (kStemTW / 2) + xl;
        if (y4 < H4 && x4 < W4) {
            bool ok[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int yy = 2 * y4 - 1 + t / 3, xx = 2 * x4 - 1 + t % 3;
                ok[t] = yy >= 0 && yy < H2 && xx >= 0 && xx < W2;
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float sum = 0.f;
#pragma unroll
                for (int t = 0; t < 9; ++t)
                    if (ok[t]) sum += sp1[c][2 * yl + t / 3][2 * xl + t % 3];
                sum = sum / 9.f;
                const int ch = 128 + c;
                p.out1cat[((size_t)b * 131 + ch) * H4 * W4 + (size_t)y4 * W4 + x4] = bn_prelu(sum, p.b2_s[ch], p.b2_t[ch], p.b2_a[ch]);
            }
        }
    }
    if (x2 >= W2) return;
    const size_t plane = (size_t)H2 * W2;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int y2 = y2a + 8 * r;
        if (y2 >= H2) continue;
        const size_t pix = (size_t)y2 * W2 + x2;
        float* o = p.out0cat + (size_t)b * 19 * plane + pix;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float y = bn_prelu(acc[r][j], sp[j], sp[16 + j], sp[32 + j]);    // level1.bn + level1.act
            y = bn_prelu(y, sp[48 + j], sp[67 + j], sp[86 + j]);              // b1 on channels 0..15
            o[(size_t)j * plane] = y;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float s = pool[r][c];
            o[(size_t)(16 + c) * plane] = bn_prelu(s, sp[48 + 16 + c], sp[67 + 16 + c], sp[86 + 16 + c]);
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Work decomposition shared by the reduce / branch kernels: item = (crop, 4-row strip, 32-column
// chunk); item i goes to CTA i % gridDim.x and, inside it, to warp (i / gridDim.x) % warps, so that
// every SM receives the same number of items (+-1) whatever the batch.
// ------------------------------------------------------------------------------------------------
struct TileIter {
    int tiles_x, tiles_y, total;
    __device__ TileIter(int B, int H, int W) {
        tiles_x = (W + 31) >> 5;
        tiles_y = (H + kRows - 1) / kRows;
        total = B * tiles_x * tiles_y;
    }
    __device__ void decode(int item, int& b, int& y0, int& x) const {
        const int tx = item % tiles_x;
        const int ty = (item / tiles_x) % tiles_y;
        b = item / (tiles_x * tiles_y);
        y0 = ty * kRows;
        x = tx * 32 + (threadIdx.x & 31);
    }
};

// ------------------------------------------------------------------------------------------------
// acc += 3x3 conv (stride S, dilation d, zero padding d) of the N-channel planar map `src` for the
// thread's 4 output pixels (rows yo0..yo0+3 at column xo).  Input channels are consumed in groups of G
// with a register double buffer: the global loads of group i+1 (possibly the first group of the next
// tap) are issued before the FFMAs of group i, so the L2 latency of the activation loads is covered
// by ~4*G*CO FFMAs per warp instead of being exposed at every step.  Taps that fall into the zero
// padding for the whole warp are dropped (warp-uniform mask).  wsm: [9][N][pad4(CO)] in shared memory
// (+ one row of slack when N % G != 0).
// ------------------------------------------------------------------------------------------------
template <int N, int CO, int G, int S>
struct Conv3x3Pipe {
    static constexpr int CP = pad4(CO);
    static constexpr int NG = (N + G - 1) / G;

    const float* __restrict__ src;
    size_t plane;          // input plane size in elements (Hi * pitch)
    int Hi, Wi, pitch, Ho, Wo, yo0, xo, d;

    __device__ __forceinline__ void geometry(int tap, int (&off)[kRows], bool (&ok)[kRows]) const {
        const int ky = tap / 3, kx = tap - 3 * ky;
        const int xi = S * xo + (kx - 1) * d;
        const bool xok = (xo < Wo) && (xi >= 0) && (xi < Wi);
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            const int yi = S * (yo0 + r) + (ky - 1) * d;
            ok[r] = xok && (yo0 + r < Ho) && (yi >= 0) && (yi < Hi);
            off[r] = ok[r] ? yi * pitch + xi : 0;
        }
    }
    __device__ __forceinline__ void load(float (&a)[G][kRows], const int (&off)[kRows], const bool (&ok)[kRows], int g) const {
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const int ci = g * G + j;
            const float* p = src + (size_t)ci * plane;
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                bool v = ok[r];
                if (N % G != 0) v = v && (ci < N);
                a[j][r] = v ? __ldg(p + off[r]) : 0.f;
            }
        }
    }
    __device__ __forceinline__ void fma_group(float (&acc)[kRows][CO], const float (&a)[G][kRows], const float* __restrict__ w) const {
#pragma unroll
        for (int j = 0; j < G; ++j) fma_tile<CO>(acc, a[j], w + j * CP);
    }

    // compiler-scheduled variant: plain tap / channel loops, loads hoisted by the unroller
    __device__ __forceinline__ void run_simple(float (&acc)[kRows][CO], const float* __restrict__ wsm) const {
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
            int off[kRows];
            bool ok[kRows];
            geometry(tap, off, ok);
            bool any = false;
#pragma unroll
            for (int r = 0; r < kRows; ++r) any |= ok[r];
            if (!__any_sync(0xffffffffu, any)) continue;
            const float* wt = wsm + (size_t)tap * N * CP;
#pragma unroll 4
            for (int ci = 0; ci < N; ++ci) {
                float a[kRows];
#pragma unroll
                for (int r = 0; r < kRows; ++r) a[r] = ok[r] ? __ldg(src + (size_t)ci * plane + off[r]) : 0.f;
                fma_tile<CO>(acc, a, wt + ci * CP);
            }
        }
    }

    __device__ __forceinline__ void run(float (&acc)[kRows][CO], const float* __restrict__ wsm) const {
#if ESPNET_PIPE == 0
        run_simple(acc, wsm);
        return;
#endif
        unsigned mask = 0;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            int off[kRows];
            bool ok[kRows];
            geometry(tap, off, ok);
            bool any = false;
#pragma unroll
            for (int r = 0; r < kRows; ++r) any |= ok[r];
            if (__any_sync(0xffffffffu, any)) mask |= 1u << tap;
        }
        if (mask == 0) return;
        int tap = __ffs(mask) - 1;
        mask &= mask - 1;
        int off[kRows];
        bool ok[kRows];
        geometry(tap, off, ok);
        float a0[G][kRows], a1[G][kRows];
        load(a0, off, ok, 0);
        int g = 0;
        // two groups per trip so that the double buffer is a static ping-pong (no register copies)
        while (true) {
            // ---- prefetch the group after (tap,g) into a1, compute (tap,g) from a0 ----
            int tap_n = tap, g_n = g + 1;
            bool more = true;
            if (g_n == NG) {
                g_n = 0;
                if (mask == 0) more = false;
                else { tap_n = __ffs(mask) - 1; mask &= mask - 1; geometry(tap_n, off, ok); }
            }
            const float* w_cur = wsm + ((size_t)tap * N + g * G) * CP;
            if (more) load(a1, off, ok, g_n);
            fma_group(acc, a0, w_cur);
            if (!more) break;
            tap = tap_n; g = g_n;
            // ---- same with the buffers swapped ----
            tap_n = tap; g_n = g + 1;
            if (g_n == NG) {
                g_n = 0;
                if (mask == 0) more = false;
                else { tap_n = __ffs(mask) - 1; mask &= mask - 1; geometry(tap_n, off, ok); }
            }
            w_cur = wsm + ((size_t)tap * N + g * G) * CP;
            if (more) load(a0, off, ok, g_n);
            fma_group(acc, a1, w_cur);
            if (!more) break;
            tap = tap_n; g = g_n;
        }
    }
};

// ------------------------------------------------------------------------------------------------
// ESP reduce: 1x1 conv CIN -> CO (Model.py:178,192), o1 planar [B,CO,H,W].
// ------------------------------------------------------------------------------------------------
template <int CIN, int CO>
__global__ void __launch_bounds__(256) reduce1x1_kernel(const float* __restrict__ in, const float* __restrict__ w /*[CIN][pad4(CO)]*/,
                                                        float* __restrict__ o1, int B, int HW, int W, int Wp) {
    constexpr int CP = pad4(CO);
    constexpr int G = 4;
    static_assert(CIN % (2 * G) == 0, "channel groups are processed in ping-pong pairs");
    extern __shared__ __align__(16) float smem[];
    copy_to_smem(smem, w, CIN * CP);
    __syncthreads();
    const int chunks = (HW + 127) / 128;           // a warp item = 128 consecutive pixels (4 per lane)
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
        // o1 rows are padded to Wp (16 B multiple, TMA requirement): pixel p -> (p / W) * Wp + p % W
        const size_t oplane = (size_t)(HW / W) * Wp;
        float* dst = o1 + (size_t)b * CO * oplane;
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            if (!ok[r]) continue;
            const int pp = p0 + 32 * r;
            const size_t o = (W == Wp) ? (size_t)pp : (size_t)(pp / W) * Wp + (pp % W);
#pragma unroll
            for (int j = 0; j < CO; ++j) dst[(size_t)j * oplane + o] = acc[r][j];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// DownSamplerB reduce: 3x3 stride-2 pad-1 conv CIN -> CO (Model.py:135,145), in [B,CIN,Hi,Wi]
// -> o1 [B,CO,Hi/2,Wi/2].  Weights [tap][CIN][pad4(CO)] resident in shared memory.
// ------------------------------------------------------------------------------------------------
template <int CIN, int CO>
__global__ void __launch_bounds__(kHeavyThreads, 1) reduce3x3s2_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                                       float* __restrict__ o1, int B, int Hi, int Wi, int Wop) {
    constexpr int CP = pad4(CO);
    constexpr int G = 4;
    extern __shared__ __align__(16) float smem[];
    copy_to_smem(smem, w, 9 * CIN * CP);
    for (int i = threadIdx.x; i < G * CP; i += blockDim.x) smem[9 * CIN * CP + i] = 0.f;   // slack rows read with a == 0
    __syncthreads();
    const int Ho = Hi >> 1, Wo = Wi >> 1;
    const size_t iplane = (size_t)Hi * Wi, oplane = (size_t)Ho * Wop;   // o1 rows are padded to a 16 B multiple (TMA)
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
            float* dst = o1 + (size_t)b * CO * oplane + (size_t)y0 * Wop + x;
#pragma unroll
            for (int j = 0; j < CO; ++j)
#pragma unroll
                for (int r = 0; r < kRows; ++r)
                    if (y0 + r < Ho) dst[(size_t)j * oplane + (size_t)r * Wop] = acc[r][j];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The ESP "split-transform-merge" stage shared by DownSamplerB (Model.py:146-159) and
// DilatedParllelResidualBlockB (Model.py:194-213): five dilated 3x3 branches on o1 (d=1,2,4,8,16),
// hierarchical-feature-fusion adds, concat (channel offsets), optional residual add BEFORE BN,
// folded BN + PReLU, and (tight fusion) an optional second BN + PReLU that writes the result
// straight into the following concat buffer (b2 / b3 of Model.py:263,269).
// The HFF running sum add_k = add_{k-1} + d_{2^k} is the accumulator itself: the chain branches
// keep accumulating into the same registers and each partial sum is emitted as one concat slice.
// ------------------------------------------------------------------------------------------------
struct BranchParams {
    const float* o1;        // [B,N,H,W]
    const float* w_d1;      // [9][N][pad4(CO1)]
    const float* w_chain;   // [4][9][N][pad4(CO)]   (d2,d4,d8,d16)
    const float* res;       // [B,C,H,W] residual input or nullptr
    const float *s, *t, *a; // own BN scale/shift + PReLU slope (C)
    float* out;             // [B,C,H,W] or nullptr
    const float *s2, *t2, *a2;  // second BR, indexed by the cat channel
    float* out2;            // [B,C2,H,W] or nullptr
    int C2, c2_off;
    int B, H, W;
    int Wp;                 // row pitch of o1 (W rounded up to 4 floats)
};

template <int N, int CO1, int CO>
__global__ void __launch_bounds__(kHeavyThreads, 1) esp_branch_kernel(const BranchParams p) {
    constexpr int C = CO1 + 4 * CO;
    constexpr int CP1 = pad4(CO1), CP = pad4(CO);
    constexpr int W1 = 9 * N * CP1, WC = 4 * 9 * N * CP;
    constexpr int G = (N == 25) ? ESPNET_G25 : ESPNET_G12;
    extern __shared__ __align__(16) float smem[];
    // layout: d1 weights | slack | chain weights | slack | epilogue params.  The slack rows are read (times
    // a == 0) by the last, partial channel group when N % G != 0.
    constexpr int SLACK = 8 * CP1;
    float* sw1 = smem;
    float* swc = smem + W1 + SLACK;
    float* sep = swc + WC + SLACK;   // 6*C epilogue params
    copy_to_smem(sw1, p.w_d1, W1);
    copy_to_smem(swc, p.w_chain, WC);
    for (int i = threadIdx.x; i < SLACK; i += blockDim.x) { sw1[W1 + i] = 0.f; swc[WC + i] = 0.f; }
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        sep[i] = p.s[i]; sep[C + i] = p.t[i]; sep[2 * C + i] = p.a[i];
        if (p.out2) {
            sep[3 * C + i] = p.s2[p.c2_off + i]; sep[4 * C + i] = p.t2[p.c2_off + i]; sep[5 * C + i] = p.a2[p.c2_off + i];
        }
    }
    __syncthreads();
    const int H = p.H, W = p.W;
    const size_t plane = (size_t)H * W;
    const TileIter it(p.B, H, W);
    const int warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const float* __restrict__ res = p.res;
    float* __restrict__ out = p.out;
    float* __restrict__ out2 = p.out2;

    for (int k = warp;; k += nwarp) {
        const int item = k * gridDim.x + blockIdx.x;
        if (item >= it.total) break;
        int b, y0, x;
        it.decode(item, b, y0, x);
        const size_t iplane = (size_t)H * p.Wp;
        const float* src = p.o1 + (size_t)b * N * iplane;
        const size_t pix0 = (size_t)y0 * W + x;

        // one concat slice [ch0, ch0+CNT): residual add before BN (Model.py:211-212), BN + PReLU, optional 2nd BR.
        // The residual loads of 8 channels x 4 rows are issued together before any of them is consumed, so the
        // epilogue pays one memory latency per 32 elements instead of one per element.
        auto emit = [&](auto& acc, int ch0, auto cnt_tag) {
            constexpr int CNT = decltype(cnt_tag)::value;
            constexpr int CH = 8;
            if (x >= W) return;
#pragma unroll
            for (int j0 = 0; j0 < CNT; j0 += CH) {
                float rv[CH][kRows];
#pragma unroll
                for (int jj = 0; jj < CH; ++jj)
#pragma unroll
                    for (int r = 0; r < kRows; ++r) {
                        rv[jj][r] = 0.f;
                        if (j0 + jj < CNT && res != nullptr && y0 + r < H)
                            rv[jj][r] = __ldg(res + ((size_t)b * C + ch0 + j0 + jj) * plane + pix0 + (size_t)r * W);
                    }
#pragma unroll
                for (int jj = 0; jj < CH; ++jj) {
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

        // ---- branch d1: CO1 outputs ---------------------------------------------------------------
        {
            float acc[kRows][CO1];
#pragma unroll
            for (int r = 0; r < kRows; ++r)
#pragma unroll
                for (int j = 0; j < CO1; ++j) acc[r][j] = 0.f;
            Conv3x3Pipe<N, CO1, G, 1> pipe{src, iplane, H, W, p.Wp, H, W, y0, x, 1};
            pipe.run(acc, sw1);
            emit(acc, 0, IntTag<CO1>());
        }
        // ---- chain d2 -> d4 -> d8 -> d16 with the HFF sum living in the accumulator -----------------
        {
            float acc[kRows][CO];
#pragma unroll
            for (int r = 0; r < kRows; ++r)
#pragma unroll
                for (int j = 0; j < CO; ++j) acc[r][j] = 0.f;
#pragma unroll 1
            for (int br = 0; br < 4; ++br) {
                Conv3x3Pipe<N, CO, G, 1> pipe{src, iplane, H, W, p.Wp, H, W, y0, x, 2 << br};
                pipe.run(acc, swc + br * 9 * N * CP);
                emit(acc, CO1 + br * CO, IntTag<CO>());
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Decoder / head kernels.  NC = classes.
// ------------------------------------------------------------------------------------------------
// S7: b3 output -> encoder.classifier 1x1 (Model.py:271,302); FULL net additionally applies `br`
// (BN, no activation, Model.py:331) and up_l3 ConvTranspose2d k2 s2 (Model.py:334,370):
// out[o][2y+a][2x+b] = sum_c v[c] * Wt[c][o][a][b].
template <int NC>
struct Head3Params {
    const float* in;     // out2cat [B,256,H8,W8]
    const float* w;      // [256][NC]
    const float *bn_s, *bn_t;   // br (NC)
    const float* wt;     // up_l3 [NC][NC][2][2]
    float* enc_out;      // [B,NC,H8,W8] or nullptr
    float* up_out;       // [B,NC,2*H8,2*W8] or nullptr
    int B, H8, W8;
};

template <int NC>
__global__ void __launch_bounds__(256) head3_kernel(const Head3Params<NC> p) {
    __shared__ float sw[256 * NC];
    __shared__ float swt[NC * NC * 4];
    __shared__ float sbn[2 * NC];
    for (int i = threadIdx.x; i < 256 * NC; i += 256) sw[i] = p.w[i];
    if (p.up_out) {
        for (int i = threadIdx.x; i < NC * NC * 4; i += 256) swt[i] = p.wt[i];
        for (int i = threadIdx.x; i < NC; i += 256) { sbn[i] = p.bn_s[i]; sbn[NC + i] = p.bn_t[i]; }
    }
    __syncthreads();
    const size_t plane = (size_t)p.H8 * p.W8;
    const size_t n = (size_t)p.B * plane;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const int b = (int)(i / plane);
        const size_t pix = i % plane;
        const float* src = p.in + (size_t)b * 256 * plane + pix;
        float acc[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) acc[j] = 0.f;
#pragma unroll 8
        for (int ci = 0; ci < 256; ++ci) {
            const float a = __ldg(src + (size_t)ci * plane);
#pragma unroll
            for (int j = 0; j < NC; ++j) acc[j] = fmaf(a, sw[ci * NC + j], acc[j]);
        }
        if (p.enc_out) {
#pragma unroll
            for (int j = 0; j < NC; ++j) p.enc_out[((size_t)b * NC + j) * plane + pix] = acc[j];
        }
        if (p.up_out) {
            const int y = (int)(pix / p.W8), x = (int)(pix % p.W8);
            const int W4 = 2 * p.W8;
            const size_t plane4 = 4 * plane;
#pragma unroll
            for (int j = 0; j < NC; ++j) acc[j] = fmaf(acc[j], sbn[j], sbn[NC + j]);
#pragma unroll
            for (int o = 0; o < NC; ++o) {
                float r00 = 0.f, r01 = 0.f, r10 = 0.f, r11 = 0.f;
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const float* wq = swt + (c * NC + o) * 4;
                    r00 = fmaf(acc[c], wq[0], r00); r01 = fmaf(acc[c], wq[1], r01);
                    r10 = fmaf(acc[c], wq[2], r10); r11 = fmaf(acc[c], wq[3], r11);
                }
                float* d = p.up_out + ((size_t)b * NC + o) * plane4 + (size_t)(2 * y) * W4 + 2 * x;
                *reinterpret_cast<float2*>(d) = make_float2(r00, r01);
                *reinterpret_cast<float2*>(d + W4) = make_float2(r10, r11);
            }
        }
    }
}

// S8 + first half of S9: level3_C 1x1 131->NC on output1_cat (Model.py:330,372), cat with up_l3's
// output (Model.py:373) and combine_l2_l3[0] BR(2NC) -> t [B,2NC,H4,W4].
template <int NC>
struct DecAParams {
    const float* out1cat;   // [B,131,H4,W4]
    const float* up3;       // [B,NC,H4,W4]
    const float* w;         // [131][NC]
    const float *s, *t, *a; // BR over 2NC
    float* tout;            // [B,2NC,H4,W4]
    int B, H4, W4;
};

template <int NC>
__global__ void __launch_bounds__(256) dec_a_kernel(const DecAParams<NC> p) {
    __shared__ float sw[131 * NC];
    __shared__ float sb[6 * NC];
    for (int i = threadIdx.x; i < 131 * NC; i += 256) sw[i] = p.w[i];
    for (int i = threadIdx.x; i < 2 * NC; i += 256) { sb[i] = p.s[i]; sb[2 * NC + i] = p.t[i]; sb[4 * NC + i] = p.a[i]; }
    __syncthreads();
    const size_t plane = (size_t)p.H4 * p.W4;
    const size_t n = (size_t)p.B * plane;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const int b = (int)(i / plane);
        const size_t pix = i % plane;
        const float* src = p.out1cat + (size_t)b * 131 * plane + pix;
        float acc[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) acc[j] = 0.f;
#pragma unroll 8
        for (int ci = 0; ci < 131; ++ci) {
            const float a = __ldg(src + (size_t)ci * plane);
#pragma unroll
            for (int j = 0; j < NC; ++j) acc[j] = fmaf(a, sw[ci * NC + j], acc[j]);
        }
        float* d = p.tout + (size_t)b * 2 * NC * plane + pix;
#pragma unroll
        for (int j = 0; j < NC; ++j) d[(size_t)j * plane] = bn_prelu(acc[j], sb[j], sb[2 * NC + j], sb[4 * NC + j]);
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            const float v = __ldg(p.up3 + ((size_t)b * NC + j) * plane + pix);
            d[(size_t)(NC + j) * plane] = bn_prelu(v, sb[NC + j], sb[3 * NC + j], sb[5 * NC + j]);
        }
    }
}

// Second half of S9: combine_l2_l3[1] CBR 3x3 2NC->NC (Model.py:335) + up_l2 = ConvT k2 s2 + BR(NC)
// (Model.py:337) -> comb [B,NC,H2,W2].
template <int NC>
struct DecBParams {
    const float* tin;       // [B,2NC,H4,W4]
    const float* w;         // [2NC][9][NC]
    const float *s, *t, *a; // CBR BN + PReLU (NC)
    const float* wt;        // up_l2.0 [NC][NC][2][2]
    const float *s2, *t2, *a2;  // up_l2.1 BR (NC)
    float* comb;            // [B,NC,H2,W2]
    int B, H4, W4;
};

template <int NC>
__global__ void __launch_bounds__(256) dec_b_kernel(const DecBParams<NC> p) {
    __shared__ float sw[2 * NC * 9 * NC];
    __shared__ float swt[NC * NC * 4];
    __shared__ float sb[6 * NC];
    for (int i = threadIdx.x; i < 2 * NC * 9 * NC; i += 256) sw[i] = p.w[i];
    for (int i = threadIdx.x; i < NC * NC * 4; i += 256) swt[i] = p.wt[i];
    for (int i = threadIdx.x; i < NC; i += 256) {
        sb[i] = p.s[i]; sb[NC + i] = p.t[i]; sb[2 * NC + i] = p.a[i];
        sb[3 * NC + i] = p.s2[i]; sb[4 * NC + i] = p.t2[i]; sb[5 * NC + i] = p.a2[i];
    }
    __syncthreads();
    const int H4 = p.H4, W4 = p.W4;
    const size_t plane = (size_t)H4 * W4;
    const size_t n = (size_t)p.B * plane;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const int b = (int)(i / plane);
        const int pix = (int)(i % plane);
        const int y = pix / W4, x = pix % W4;
        const float* src = p.tin + (size_t)b * 2 * NC * plane;
        float acc[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) acc[j] = 0.f;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
            if (yy < 0 || yy >= H4 || xx < 0 || xx >= W4) continue;
            const size_t o = (size_t)yy * W4 + xx;
#pragma unroll 2
            for (int ci = 0; ci < 2 * NC; ++ci) {
                const float a = __ldg(src + (size_t)ci * plane + o);
                const float* wr = sw + (ci * 9 + tap) * NC;
#pragma unroll
                for (int j = 0; j < NC; ++j) acc[j] = fmaf(a, wr[j], acc[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < NC; ++j) acc[j] = bn_prelu(acc[j], sb[j], sb[NC + j], sb[2 * NC + j]);
        const int W2 = 2 * W4;
        const size_t plane2 = 4 * plane;
#pragma unroll
        for (int o = 0; o < NC; ++o) {
            float r[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const float* wq = swt + (c * NC + o) * 4;
#pragma unroll
                for (int q = 0; q < 4; ++q) r[q] = fmaf(acc[c], wq[q], r[q]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) r[q] = bn_prelu(r[q], sb[3 * NC + o], sb[4 * NC + o], sb[5 * NC + o]);
            float* d = p.comb + ((size_t)b * NC + o) * plane2 + (size_t)(2 * y) * W2 + 2 * x;
            *reinterpret_cast<float2*>(d) = make_float2(r[0], r[1]);
            *reinterpret_cast<float2*>(d + W2) = make_float2(r[2], r[3]);
        }
    }
}

// S10: `conv` CBR 3x3 (NC+19)->NC on cat[comb, output0_cat] (Model.py:332,375) + classifier
// ConvT k2 s2 (Model.py:339,377) -> logits [B,NC,H,W]; fused epilogues: arg-max mask
// (VisualizeResults_iou.py:128, ties -> lowest index) and softmax accumulation (ensemble extension).
template <int NC>
struct DecCParams {
    const float* comb;      // [B,NC,H2,W2]
    const float* out0cat;   // [B,19,H2,W2]
    const float* w;         // [(NC+19)][9][NC]
    const float *s, *t, *a; // conv BN + PReLU
    const float* wt;        // classifier [NC][NC][2][2]
    float* logits;          // [B,NC,H,W] or nullptr
    unsigned char* mask;    // [B,H,W] or nullptr
    float* prob_acc;        // [B,NC,H,W] or nullptr
    int prob_init, mask_from_prob;
    int B, H2, W2;
};

template <int NC>
__device__ __forceinline__ int argmax_first(const float (&v)[NC]) {
    int best = 0;
    float bv = v[0];
#pragma unroll
    for (int j = 1; j < NC; ++j)
        if (v[j] > bv) { bv = v[j]; best = j; }
    return best;
}

template <int NC>
__global__ void __launch_bounds__(256) dec_c_kernel(const DecCParams<NC> p) {
    constexpr int CI = NC + 19;
    __shared__ float sw[CI * 9 * NC];
    __shared__ float swt[NC * NC * 4];
    __shared__ float sb[3 * NC];
    for (int i = threadIdx.x; i < CI * 9 * NC; i += 256) sw[i] = p.w[i];
    for (int i = threadIdx.x; i < NC * NC * 4; i += 256) swt[i] = p.wt[i];
    for (int i = threadIdx.x; i < NC; i += 256) { sb[i] = p.s[i]; sb[NC + i] = p.t[i]; sb[2 * NC + i] = p.a[i]; }
    __syncthreads();
    const int H2 = p.H2, W2 = p.W2;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int b = blockIdx.z;
    if (x >= W2 || y >= H2) return;
    const size_t plane = (size_t)H2 * W2;
    float acc[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) acc[j] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        if (yy < 0 || yy >= H2 || xx < 0 || xx >= W2) continue;
        const size_t o = (size_t)yy * W2 + xx;
        const float* s0 = p.comb + (size_t)b * NC * plane + o;
#pragma unroll
        for (int ci = 0; ci < NC; ++ci) {
            const float a = __ldg(s0 + (size_t)ci * plane);
            const float* wr = sw + (ci * 9 + tap) * NC;
#pragma unroll
            for (int j = 0; j < NC; ++j) acc[j] = fmaf(a, wr[j], acc[j]);
        }
        const float* s1 = p.out0cat + (size_t)b * 19 * plane + o;
#pragma unroll
        for (int ci = 0; ci < 19; ++ci) {
            const float a = __ldg(s1 + (size_t)ci * plane);
            const float* wr = sw + ((NC + ci) * 9 + tap) * NC;
#pragma unroll
            for (int j = 0; j < NC; ++j) acc[j] = fmaf(a, wr[j], acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < NC; ++j) acc[j] = bn_prelu(acc[j], sb[j], sb[NC + j], sb[2 * NC + j]);
    const int W = 2 * W2;
    const size_t fplane = 4 * plane;
    float lg[4][NC];   // the 2x2 output pixels x classes
#pragma unroll
    for (int o = 0; o < NC; ++o) {
#pragma unroll
        for (int q = 0; q < 4; ++q) lg[q][o] = 0.f;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const float* wq = swt + (c * NC + o) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) lg[q][o] = fmaf(acc[c], wq[q], lg[q][o]);
        }
    }
    const size_t base = (size_t)(2 * y) * W + 2 * x;
    if (p.logits) {
#pragma unroll
        for (int o = 0; o < NC; ++o) {
            float* d = p.logits + ((size_t)b * NC + o) * fplane + base;
            if ((reinterpret_cast<uintptr_t>(p.logits) & 7) == 0) {     // caller buffers may be 4-byte-offset views
                *reinterpret_cast<float2*>(d) = make_float2(lg[0][o], lg[1][o]);
                *reinterpret_cast<float2*>(d + W) = make_float2(lg[2][o], lg[3][o]);
            } else {
                d[0] = lg[0][o]; d[1] = lg[1][o]; d[W] = lg[2][o]; d[W + 1] = lg[3][o];
            }
        }
    }
    int am[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) am[q] = argmax_first<NC>(lg[q]);
    if (p.prob_acc) {
        float pr[4][NC];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float m = lg[q][0];
#pragma unroll
            for (int o = 1; o < NC; ++o) m = fmaxf(m, lg[q][o]);
            float sum = 0.f;
#pragma unroll
            for (int o = 0; o < NC; ++o) { pr[q][o] = expf(lg[q][o] - m); sum += pr[q][o]; }
            const float inv = 1.f / sum;
#pragma unroll
            for (int o = 0; o < NC; ++o) pr[q][o] *= inv;
        }
#pragma unroll
        for (int o = 0; o < NC; ++o) {
            float* d = p.prob_acc + ((size_t)b * NC + o) * fplane + base;
            const bool al8 = (reinterpret_cast<uintptr_t>(p.prob_acc) & 7) == 0;
            float2 t0 = make_float2(0.f, 0.f), t1 = make_float2(0.f, 0.f);
            if (!p.prob_init) {
                if (al8) { t0 = *reinterpret_cast<float2*>(d); t1 = *reinterpret_cast<float2*>(d + W); }
                else { t0 = make_float2(d[0], d[1]); t1 = make_float2(d[W], d[W + 1]); }
            }
            pr[0][o] += t0.x; pr[1][o] += t0.y; pr[2][o] += t1.x; pr[3][o] += t1.y;
            if (al8) {
                *reinterpret_cast<float2*>(d) = make_float2(pr[0][o], pr[1][o]);
                *reinterpret_cast<float2*>(d + W) = make_float2(pr[2][o], pr[3][o]);
            } else {
                d[0] = pr[0][o]; d[1] = pr[1][o]; d[W] = pr[2][o]; d[W + 1] = pr[3][o];
            }
        }
        if (p.mask_from_prob) {
#pragma unroll
            for (int q = 0; q < 4; ++q) am[q] = argmax_first<NC>(pr[q]);
        }
    }
    if (p.mask) {
        unsigned char* d = p.mask + (size_t)b * fplane + base;
        if ((reinterpret_cast<uintptr_t>(p.mask) & 1) == 0) {        // W and base are even; caller buffers may be odd-offset views
            *reinterpret_cast<uchar2*>(d) = make_uchar2((unsigned char)am[0], (unsigned char)am[1]);
            *reinterpret_cast<uchar2*>(d + W) = make_uchar2((unsigned char)am[2], (unsigned char)am[3]);
        } else {
            d[0] = (unsigned char)am[0]; d[1] = (unsigned char)am[1];
            d[W] = (unsigned char)am[2]; d[W + 1] = (unsigned char)am[3];
        }
    }
}

// ESPNet-C tail: nn.Upsample(scale_factor=8, mode='bilinear', align_corners=False)
// (VisualizeResults_iou.py:258-261,125-126) + arg-max (:128) of encoder logits [B,NC,H8,W8].
// One thread = 4 horizontally adjacent output pixels (a uchar4 / float4 store per row) x the 8 output rows that share the
// same pair of source rows (row group k: output rows [8k + 4, 8k + 12), k = -1 .. H8 - 1; the first and the last group hold the 4
// rows whose source index is clamped).  The 4 pixels read at most 3 source columns, loaded once per class; the two HORIZONTAL
// interpolations per (pixel, class) are computed once and reused by the rows, each of which only blends them vertically and
// updates its running arg-max (class-outer loop: registers do not grow with NC).  Same formula as torch's upsample_bilinear2d: h0 * (w0 * v00 + w1 * v01) + h1 * (w0 * v10 + w1 * v11).
// (The one-row-per-thread version of round 1 spent 130 instructions per pixel and 90 us per 64 crops of 512 x 512, issue-bound;
// this one 51 us.)
constexpr int kUpRows = 8;       // output rows per thread = all rows that share a source-row pair (4 rows per thread: 57 instead of 51 us)
template <int NC>
__global__ void __launch_bounds__(256, 2) upsample8_argmax_kernel(const float* __restrict__ enc, int B, int H8, int W8,
                                                               unsigned char* __restrict__ mask, float* __restrict__ up_logits) {
    pdl_trigger();
    pdl_wait();
    const int H = 8 * H8, W = 8 * W8;
    const int xg = 4 * (blockIdx.x * 32 + (threadIdx.x & 31));
    const int k = (int)blockIdx.y * 8 + (int)(threadIdx.x >> 5) - 1, ybase = 8 * k + 4;
    const int b = blockIdx.z;
    if (xg >= W || k >= H8) return;
    // area_pixel_compute_source_index(scale=1/8, align_corners=False): src = (dst+0.5)/8-0.5, clamped at 0
    const int y0 = k < 0 ? 0 : k;
    const int y1 = y0 + (y0 < H8 - 1 ? 1 : 0);
    int x0[4]; float lx1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float sx = 0.125f * ((float)(xg + j) + 0.5f) - 0.5f; sx = sx < 0.f ? 0.f : sx;
        x0[j] = (int)sx;
        lx1[j] = sx - (float)x0[j];
    }
    // source columns c0 = x0[0], c0 + 1, c0 + 2 (clamped): x0[j] is c0 or c0 + 1, its right neighbour min(x0[j] + 1, W8 - 1)
    const int c0 = x0[0], c1 = min(c0 + 1, W8 - 1), c2 = min(c0 + 2, W8 - 1);
    const size_t plane = (size_t)H8 * W8;
    // rows of this group and their vertical weights
    float ly1r[kUpRows];
#pragma unroll
    for (int r = 0; r < kUpRows; ++r) {
        float sy = 0.125f * ((float)(ybase + r) + 0.5f) - 0.5f; sy = sy < 0.f ? 0.f : sy;
        ly1r[r] = sy - (float)y0;
    }
    const bool lg_vec = (reinterpret_cast<uintptr_t>(up_logits) & 15) == 0;
    float bv[kUpRows][4];        // running arg-max per (row, pixel): the first maximum wins, like argmax_first
    int bi[kUpRows][4];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const float* s = enc + ((size_t)b * NC + c) * plane;
        const float* r0 = s + (size_t)y0 * W8;
        const float* r1 = s + (size_t)y1 * W8;
        const float a0 = __ldg(r0 + c0), a1 = __ldg(r0 + c1), a2 = __ldg(r0 + c2);
        const float b0 = __ldg(r1 + c0), b1 = __ldg(r1 + c1), b2 = __ldg(r1 + c2);
        float t0[4], t1[4];      // horizontal interpolation on source rows y0 / y1
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool sh = x0[j] != c0;
            const float v00 = sh ? a1 : a0, v01 = sh ? a2 : a1, v10 = sh ? b1 : b0, v11 = sh ? b2 : b1;
            const float lx0 = 1.f - lx1[j];
            t0[j] = lx0 * v00 + lx1[j] * v01;
            t1[j] = lx0 * v10 + lx1[j] * v11;
        }
#pragma unroll
        for (int r = 0; r < kUpRows; ++r) {
            const int y = ybase + r;
            const float ly1 = ly1r[r], ly0 = 1.f - ly1;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[j] = ly0 * t0[j] + ly1 * t1[j];
                if (c == 0) { bv[r][j] = v[j]; bi[r][j] = 0; }
                else if (v[j] > bv[r][j]) { bv[r][j] = v[j]; bi[r][j] = c; }
            }
            if (up_logits && y >= 0 && y < H) {
                float* d = up_logits + ((size_t)b * NC + c) * H * W + (size_t)y * W + xg;
                if (lg_vec) *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
                else { d[0] = v[0]; d[1] = v[1]; d[2] = v[2]; d[3] = v[3]; }
            }
        }
    }
    if (mask) {
        const bool mk_vec = (reinterpret_cast<uintptr_t>(mask) & 3) == 0;     // caller buffers may be unaligned views
#pragma unroll
        for (int r = 0; r < kUpRows; ++r) {
            const int y = ybase + r;
            if (y < 0 || y >= H) continue;
            unsigned char* d = mask + (size_t)b * H * W + (size_t)y * W + xg;
            const uchar4 m4 = make_uchar4((unsigned char)bi[r][0], (unsigned char)bi[r][1], (unsigned char)bi[r][2], (unsigned char)bi[r][3]);
            if (mk_vec) *reinterpret_cast<uchar4*>(d) = m4;
            else { d[0] = m4.x; d[1] = m4.y; d[2] = m4.z; d[3] = m4.w; }
        }
    }
}

}  // namespace espnet
