// Kernels for the steps immediately before and after the forward for crops that are NOT already the network's input
// size, and for the slide render ("next" rows 2-4):
//   * preprocess_resize_kernel   VisualizeResults_iou.py:107-119: u8 BGR -> fp32, per-channel (p - mean) / std,
//                                cv2.resize(..., INTER_LINEAR) to inWidth x inHeight, / 255, HWC -> CHW
//   * resize_nearest_u8_kernel   VisualizeResults_iou.py:129: class map back to the crop's size, cv2 INTER_NEAREST
//   * palette_overlay_kernel     VisualizeResults_iou.py:139-147: palette colour map + cv2.addWeighted(img, 0.4, map, 0.6, 0)
//   * render_ds8_kernel          eval_wsi_segmentation.py:225-240: /8 nearest of slide and label, palette, addWeighted,
//                                paste -- the *_pred.jpg pixels
//   * class_count_kernel         VisualizeResults_iou.py:151-155: per-class pixel counts of a class map
// Everything that involves double / Python-float index arithmetic (bilinear source coordinates, nearest LUTs) is computed on
// the host and handed over as tables, so indices are bit-exact; the float arithmetic below repeats OpenCV's generic
// (non-IPP) code path operation by operation with explicitly non-fused multiplies and adds.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace espnet {

// out[b][c][dy][dx] = ((r0 * (1-fy) + r1 * fy)) / 255 with r_k = S_k[x0] * (1-fx) + S_k[x1] * fx and
// S(y, x) = (float(u8) - mean_c) / std_c  -- horizontal pass first, then vertical, like cv2's HResizeLinear / VResizeLinear.
// xs/ys: source index (already clamped), xf/yf: weight of the +1 neighbour (0 where clamped).
__global__ void __launch_bounds__(256) preprocess_resize_kernel(const unsigned char* __restrict__ in, int B, int h, int w,
                                                                float m0, float m1, float m2, float s0, float s1, float s2,
                                                                const int32_t* __restrict__ xs, const float* __restrict__ xf,
                                                                const int32_t* __restrict__ ys, const float* __restrict__ yf,
                                                                float* __restrict__ out, int H, int W) {
    const int dx = blockIdx.x * 32 + (threadIdx.x & 31);
    const int dy = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int b = blockIdx.z;
    if (dx >= W || dy >= H) return;
    const int x0 = xs[dx], y0 = ys[dy];
    const int x1 = min(x0 + 1, w - 1), y1 = min(y0 + 1, h - 1);
    const float fx = xf[dx], fy = yf[dy];
    const float ax = __fsub_rn(1.f, fx), ay = __fsub_rn(1.f, fy);
    const unsigned char* base = in + (size_t)b * h * w * 3;
    const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        auto S = [&](int y, int x) { return __fdiv_rn(__fsub_rn((float)base[((size_t)y * w + x) * 3 + c], mean[c]), stdv[c]); };
        const float r0 = __fadd_rn(__fmul_rn(S(y0, x0), ax), __fmul_rn(S(y0, x1), fx));
        const float r1 = __fadd_rn(__fmul_rn(S(y1, x0), ax), __fmul_rn(S(y1, x1), fx));
        const float v = __fadd_rn(__fmul_rn(r0, ay), __fmul_rn(r1, fy));
        out[((size_t)(b * 3 + c) * H + dy) * W + dx] = __fdiv_rn(v, 255.f);
    }
}

// Box -> crop extraction fused with the same front-end (make_seg_data.py:347-361 output_org_files + VisualizeResults_iou.py:103-119):
// crop b is the level-0 region boxes[b] = (x0, y0, x1, y1) of the resident BGR slide (pixels outside the slide are 0, what
// OpenSlide's read_region pads and cv2.imread's alpha drop leave), resized to W x H.  Boxes differ in size, so every box has
// its own LUT row: xs/xf are [B][W], ys/yf are [B][H] (host-computed, OpenCV's index arithmetic).
__global__ void __launch_bounds__(256) preprocess_resize_boxes_kernel(const unsigned char* __restrict__ slide, int SH, int SW,
                                                                      const int32_t* __restrict__ boxes,
                                                                      float m0, float m1, float m2, float s0, float s1, float s2,
                                                                      const int32_t* __restrict__ xs, const float* __restrict__ xf,
                                                                      const int32_t* __restrict__ ys, const float* __restrict__ yf,
                                                                      float* __restrict__ out, int H, int W) {
    const int dx = blockIdx.x * 32 + (threadIdx.x & 31);
    const int dy = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int b = blockIdx.z;
    if (dx >= W || dy >= H) return;
    const int bx0 = boxes[4 * b], by0 = boxes[4 * b + 1];
    const int w = boxes[4 * b + 2] - bx0, h = boxes[4 * b + 3] - by0;
    const int x0 = xs[(size_t)b * W + dx], y0 = ys[(size_t)b * H + dy];
    const int x1 = min(x0 + 1, w - 1), y1 = min(y0 + 1, h - 1);
    const float fx = xf[(size_t)b * W + dx], fy = yf[(size_t)b * H + dy];
    const float ax = __fsub_rn(1.f, fx), ay = __fsub_rn(1.f, fy);
    const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        auto S = [&](int y, int x) {
            const int gy = by0 + y, gx = bx0 + x;
            const float p = (gy >= 0 && gy < SH && gx >= 0 && gx < SW) ? (float)slide[((size_t)gy * SW + gx) * 3 + c] : 0.f;
            return __fdiv_rn(__fsub_rn(p, mean[c]), stdv[c]);
        };
        const float r0 = __fadd_rn(__fmul_rn(S(y0, x0), ax), __fmul_rn(S(y0, x1), fx));
        const float r1 = __fadd_rn(__fmul_rn(S(y1, x0), ax), __fmul_rn(S(y1, x1), fx));
        const float v = __fadd_rn(__fmul_rn(r0, ay), __fmul_rn(r1, fy));
        out[((size_t)(b * 3 + c) * H + dy) * W + dx] = __fdiv_rn(v, 255.f);
    }
}

// dst[b][y][x] = src[b][ysrc[y]][xsrc[x]]  (cv2 INTER_NEAREST source indices from the host)
__global__ void __launch_bounds__(256) resize_nearest_u8_kernel(const unsigned char* __restrict__ src, int B, int sh, int sw,
                                                                unsigned char* __restrict__ dst, int dh, int dw,
                                                                const int32_t* __restrict__ ysrc, const int32_t* __restrict__ xsrc) {
    const int y = blockIdx.y, b = blockIdx.z;
    if (y >= dh) return;
    const unsigned char* s = src + ((size_t)b * sh + ysrc[y]) * sw;
    unsigned char* d = dst + ((size_t)b * dh + y) * dw;
    for (int x = blockIdx.x * 256 + threadIdx.x; x < dw; x += gridDim.x * 256) d[x] = s[xsrc[x]];
}

// cv2.addWeighted(a, 0.4, b, 0.6, 0) on u8: saturate(rint(a * 0.4f + b * 0.6f)), fp32 multiply and add NOT fused
// (exhaustively equal to OpenCV 4.13 over all 65536 (a, b) pairs, tests/test_oracle_frontend.py)
__device__ __forceinline__ unsigned char blend_04_06(unsigned char a, unsigned char b) {
    const float v = __fadd_rn(__fmul_rn((float)a, 0.4f), __fmul_rn((float)b, 0.6f));
    int r = __float2int_rn(v);
    r = r < 0 ? 0 : (r > 255 ? 255 : r);
    return (unsigned char)r;
}

// palette: 25 x 3 bytes (r, g, b) like the reference's PALLETE; the colour map is written as [b, g, r]
// (VisualizeResults_iou.py:141-143, eval_wsi_segmentation.py:231-233); labels >= n_pal stay black (np.zeros).
__global__ void __launch_bounds__(256) palette_overlay_kernel(const unsigned char* __restrict__ img, const unsigned char* __restrict__ label,
                                                              size_t npix, const unsigned char* __restrict__ palette, int n_pal,
                                                              unsigned char* __restrict__ color_out, unsigned char* __restrict__ overlay_out) {
    __shared__ unsigned char pal[256 * 3];
    for (int i = threadIdx.x; i < 256 * 3; i += 256) pal[i] = i < n_pal * 3 ? palette[i] : 0;
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < npix; i += (size_t)gridDim.x * 256) {
        const int l = label[i];
        const unsigned char cb = pal[3 * l + 2], cg = pal[3 * l + 1], cr = pal[3 * l + 0];
        if (color_out) { color_out[3 * i] = cb; color_out[3 * i + 1] = cg; color_out[3 * i + 2] = cr; }
        if (overlay_out) {
            overlay_out[3 * i] = blend_04_06(img[3 * i], cb);
            overlay_out[3 * i + 1] = blend_04_06(img[3 * i + 1], cg);
            overlay_out[3 * i + 2] = blend_04_06(img[3 * i + 2], cr);
        }
    }
}

// The /8 slide render: out[y][x][:] = addWeighted(slide[ysrc[y]][xsrc[x]][:], 0.4, bgr(palette[label[ysrc[y]][xsrc[x]]]), 0.6),
// 0 where a LUT entry is < 0 (windows the reference's loop skips).  The LUTs are the ones of espnet_ds8_lut.
__global__ void __launch_bounds__(256) render_ds8_kernel(const unsigned char* __restrict__ slide, const unsigned char* __restrict__ label,
                                                         int SW, const unsigned char* __restrict__ palette, int n_pal,
                                                         unsigned char* __restrict__ out, int dh, int dw,
                                                         const int32_t* __restrict__ ysrc, const int32_t* __restrict__ xsrc) {
    __shared__ unsigned char pal[256 * 3];
    for (int i = threadIdx.x; i < 256 * 3; i += 256) pal[i] = i < n_pal * 3 ? palette[i] : 0;
    __syncthreads();
    const int y = blockIdx.y;
    if (y >= dh) return;
    const int ys = ysrc[y];
    for (int x = blockIdx.x * 256 + threadIdx.x; x < dw; x += gridDim.x * 256) {
        const int xsx = xsrc[x];
        unsigned char* o = out + ((size_t)y * dw + x) * 3;
        if (ys < 0 || xsx < 0) { o[0] = o[1] = o[2] = 0; continue; }
        const size_t p = (size_t)ys * SW + xsx;
        const int l = label[p];
        o[0] = blend_04_06(slide[3 * p], pal[3 * l + 2]);
        o[1] = blend_04_06(slide[3 * p + 1], pal[3 * l + 1]);
        o[2] = blend_04_06(slide[3 * p + 2], pal[3 * l + 0]);
    }
}

// counts[b][k] += #pixels of class k in map b (k < n_classes <= 32); counts is int64 [B][n_classes], added to
__global__ void __launch_bounds__(256) class_count_kernel(const unsigned char* __restrict__ maps, size_t pix_per_map, int n_classes,
                                                          unsigned long long* __restrict__ counts) {
    __shared__ unsigned int sh[32];
    if (threadIdx.x < 32) sh[threadIdx.x] = 0;
    __syncthreads();
    const int b = blockIdx.y;
    const unsigned char* m = maps + (size_t)b * pix_per_map;
    const size_t per_cta = (pix_per_map + gridDim.x - 1) / gridDim.x;
    const size_t lo = (size_t)blockIdx.x * per_cta;
    const size_t hi = lo + per_cta < pix_per_map ? lo + per_cta : pix_per_map;
    for (size_t i = lo + threadIdx.x; i < hi; i += 256) {
        const int v = m[i];
        if (v < n_classes) atomicAdd(&sh[v], 1u);
    }
    __syncthreads();
    if (threadIdx.x < n_classes && sh[threadIdx.x]) atomicAdd(counts + (size_t)b * n_classes + threadIdx.x, (unsigned long long)sh[threadIdx.x]);
}

}  // namespace espnet
