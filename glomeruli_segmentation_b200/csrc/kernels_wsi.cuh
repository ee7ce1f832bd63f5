// Integer / byte kernels either side of the forward: tile -> slide stitching (T3), /8 nearest
// down-sample (T4) and the confusion-matrix histogram (IOUEval.py).  All index arithmetic that
// involves Python float semantics (T1 strides, cv2 nearest scale) is done on the host in double and
// handed to these kernels as integers / LUTs, so results are bit-exact by construction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace espnet {

// T3, scatter form, arbitrary boxes (eval_wsi_segmentation.py:259-316): one CTA column per box, one
// thread per aligned 32-bit word of the slide row segment the box covers; the four class bytes are
// max-merged with __vmaxu4 under an atomicCAS loop (max is commutative / idempotent, so the result
// does not depend on the order boxes land in).
__global__ void __launch_bounds__(256) stitch_boxes_kernel(unsigned char* __restrict__ slide, int SH, int SW, int y_limit,
                                                           const int32_t* __restrict__ boxes, const long long* __restrict__ offs,
                                                           const unsigned char* __restrict__ masks, int n_boxes) {
    const int bi = blockIdx.x;
    if (bi >= n_boxes) return;
    const int x0 = boxes[4 * bi], y0 = boxes[4 * bi + 1], x1 = boxes[4 * bi + 2], y1 = boxes[4 * bi + 3];
    const int bw = x1 - x0;
    const int cx0 = max(x0, 0), cx1 = min(x1, SW);
    const int cy0 = max(y0, 0), cy1 = min(min(y1, SH), y_limit);
    if (cx1 <= cx0 || cy1 <= cy0) return;
    const unsigned char* m = masks + offs[bi];
    unsigned int* words = reinterpret_cast<unsigned int*>(slide);
    const size_t total = (size_t)SH * SW;
    for (int y = cy0 + blockIdx.y; y < cy1; y += gridDim.y) {
        const size_t row = (size_t)y * SW;
        const size_t f0 = row + cx0, f1 = row + cx1;          // flat byte range [f0,f1)
        const size_t w0 = f0 >> 2, w1 = (f1 - 1) >> 2;
        const unsigned char* mrow = m + (size_t)(y - y0) * bw;        // mrow[x - x0] for slide column x
        for (size_t w = w0 + threadIdx.x; w <= w1; w += 256) {
            if (4 * w + 4 > total) continue;       // the partial last word of the buffer: stitch_boxes_tail_kernel (no access past the end)
            unsigned int cand = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const size_t f = 4 * w + k;
                if (f >= f0 && f < f1) cand |= (unsigned int)mrow[(int)(f - row) - x0] << (8 * k);
            }
            if (cand == 0) continue;
            unsigned int old = words[w];
            while (true) {
                const unsigned int nv = __vmaxu4(old, cand);
                if (nv == old) break;
                const unsigned int prev = atomicCAS(words + w, old, nv);
                if (prev == old) break;
                old = prev;
            }
        }
    }
}

// The last SH*SW % 4 bytes of the slide mask do not fill a 32-bit merge word: one thread per such byte walks all boxes
// (runs after stitch_boxes_kernel on the same stream, so nothing races) -- no word access ever crosses the buffer's end.
__global__ void stitch_boxes_tail_kernel(unsigned char* __restrict__ slide, int SH, int SW, int y_limit, const int32_t* __restrict__ boxes,
                                         const long long* __restrict__ offs, const unsigned char* __restrict__ masks, int n_boxes) {
    const size_t total = (size_t)SH * SW;
    const size_t f = (total & ~(size_t)3) + threadIdx.x;
    if (f >= total) return;
    const int y = (int)(f / SW), x = (int)(f % SW);
    if (y >= y_limit) return;
    int best = slide[f];
    for (int bi = 0; bi < n_boxes; ++bi) {
        const int x0 = boxes[4 * bi], y0 = boxes[4 * bi + 1], x1 = boxes[4 * bi + 2], y1 = boxes[4 * bi + 3];
        if (x >= x0 && x < x1 && y >= y0 && y < y1) best = max(best, (int)masks[offs[bi] + (size_t)(y - y0) * (x1 - x0) + (x - x0)]);
    }
    slide[f] = (unsigned char)best;
}

// T3, gather form for the regular T1 grid (detect_glomus_test.py:268-271): tile k = j*n_x + i sits at
// (i*stride_x, j*stride_y); each thread owns one slide pixel and takes the max over the (at most
// ceil(win/stride)^2) tiles that cover it.  Only tile rows [row0,row0+rows) are resident (multi-GPU
// band sharding); pixels no resident tile covers are left untouched.  The output buffer holds slide rows
// [out_y0, out_y0 + out_rows) (a rank's band, or the whole slide with out_y0 = 0); grid.y strides over the rows,
// so slides taller than 65535 px need no special casing.
//
// overwrite = 0: out = max(out, stitched) (the buffer is zero-initialised or already holds other tiles' pixels);
// overwrite = 1: every pixel of the covered rows is WRITTEN (0 where no tile covers it) and never read -- the form used when
// `out` is another GPU's memory mapped over NVLink (P2P band placement: a remote read-modify-write would cost a round trip).
__global__ void __launch_bounds__(256) stitch_grid_kernel(unsigned char* __restrict__ out, int out_y0, int out_rows, int SH, int SW, int y_limit,
                                                          const unsigned char* __restrict__ tiles, int n_x, int n_y,
                                                          int win_x, int win_y, int sx, int sy, int k0, int k1, int overwrite) {
    // resident tiles: indices [k0, k1) of the row-major grid (a rank's share; whole tile rows or a balanced range that starts /
    // ends inside a row); tile k sits at tiles + (k - k0) * win_x * win_y
    const int row0 = k0 / n_x, row1 = (k1 - 1) / n_x;
    const int ylo = max(row0 * sy, out_y0);
    const int yhi = min(min(min(row1 * sy + win_y, SH), y_limit), out_y0 + out_rows);
    const size_t tile_sz = (size_t)win_x * win_y;
    // max over the resident tiles that cover slide pixel (x, y); -1 if none does
    auto gather = [&](int x, int y, int j_lo, int j_hi) {
        const int i_hi = min(x / sx, n_x - 1);
        int i_lo = (x - win_x + sx) / sx;
        if (x - win_x + 1 <= 0) i_lo = 0;
        int best = -1;
        for (int j = j_lo; j <= j_hi; ++j) {
            const int ty = y - j * sy;
            if (ty < 0 || ty >= win_y) continue;
            for (int i = i_lo; i <= i_hi; ++i) {
                const int tx = x - i * sx;
                const int k = j * n_x + i;
                if (tx < 0 || tx >= win_x || k < k0 || k >= k1) continue;
                best = max(best, (int)tiles[(size_t)(k - k0) * tile_sz + (size_t)ty * win_x + tx]);
            }
        }
        return best;
    };
    // four pixels per thread and one 32-bit store when rows are 4-byte aligned (always for a multi-GPU band written over NVLink:
    // a warp then stores 128 contiguous bytes instead of 32); byte stores otherwise
    const bool vec4 = (SW % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 3) == 0);
    // ... and when stride, window and tile base are multiples of 4 as well (the reference's 512 px windows at overlap 0.1:
    // stride 460), an aligned group of four pixels is inside or outside a tile as a whole and its four bytes are ONE aligned
    // 32-bit load per covering tile: a quarter of the loads and of the index arithmetic
    const bool quad = vec4 && (sx % 4 == 0) && (win_x % 4 == 0) && ((reinterpret_cast<uintptr_t>(tiles) & 3) == 0);
    for (int y = ylo + blockIdx.y; y < yhi; y += gridDim.y) {
        const int j_hi = min(min(y / sy, n_y - 1), row1);
        int j_lo = (y - win_y + sy) / sy;                 // ceil((y - win_y + 1) / sy) for y-win_y+1 > 0
        if (y - win_y + 1 <= 0) j_lo = 0;
        j_lo = max(j_lo, row0);
        unsigned char* drow = out + (size_t)(y - out_y0) * SW;
        if (quad) {
            for (int x = 4 * (blockIdx.x * 256 + threadIdx.x); x < SW; x += 4 * gridDim.x * 256) {
                const int i_hi = min(x / sx, n_x - 1);
                int i_lo = (x - win_x + sx) / sx;
                if (x - win_x + 1 <= 0) i_lo = 0;
                unsigned int acc = 0;
                bool covered = false;
                for (int j = j_lo; j <= j_hi; ++j) {
                    const int ty = y - j * sy;
                    if (ty < 0 || ty >= win_y) continue;
                    for (int i = i_lo; i <= i_hi; ++i) {
                        const int tx = x - i * sx;
                        const int k = j * n_x + i;
                        if (tx < 0 || tx >= win_x || k < k0 || k >= k1) continue;
                        acc = __vmaxu4(acc, __ldg(reinterpret_cast<const unsigned int*>(tiles + (size_t)(k - k0) * tile_sz + (size_t)ty * win_x + tx)));
                        covered = true;
                    }
                }
                unsigned int* d = reinterpret_cast<unsigned int*>(drow + x);
                if (overwrite) *d = acc;
                else if (covered) *d = __vmaxu4(*d, acc);
            }
        } else if (vec4) {
            for (int x = 4 * (blockIdx.x * 256 + threadIdx.x); x < SW; x += 4 * gridDim.x * 256) {
                int b[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) b[q] = gather(x + q, y, j_lo, j_hi);
                unsigned int* d = reinterpret_cast<unsigned int*>(drow + x);
                if (overwrite) {
                    *d = (unsigned)max(b[0], 0) | ((unsigned)max(b[1], 0) << 8) | ((unsigned)max(b[2], 0) << 16) | ((unsigned)max(b[3], 0) << 24);
                } else if ((b[0] & b[1] & b[2] & b[3]) >= 0) {      // at least one of the four pixels is covered (b = -1 otherwise)
                    const unsigned int cand = (unsigned)max(b[0], 0) | ((unsigned)max(b[1], 0) << 8) | ((unsigned)max(b[2], 0) << 16) | ((unsigned)max(b[3], 0) << 24);
                    *d = __vmaxu4(*d, cand);
                }
            }
        } else {
            for (int x = blockIdx.x * 256 + threadIdx.x; x < SW; x += gridDim.x * 256) {
                const int best = gather(x, y, j_lo, j_hi);
                if (overwrite) drow[x] = (unsigned char)max(best, 0);
                else if (best >= 0) drow[x] = (unsigned char)max((int)drow[x], best);
            }
        }
    }
}

// dst[i] = max(dst[i], src[i]) over n bytes: the merge of the rows two ranks' bands share (tile overlap strip) when the
// band masks are gathered on rank 0.  16 B vectors where both pointers allow it, bytes otherwise.
__global__ void __launch_bounds__(256) max_merge_u8_kernel(unsigned char* __restrict__ dst, const unsigned char* __restrict__ src, size_t n) {
    const bool vec = (((uintptr_t)dst | (uintptr_t)src) & 15) == 0;
    const size_t nv = vec ? n / 16 : 0;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nv; i += (size_t)gridDim.x * 256) {
        uint4 a = reinterpret_cast<uint4*>(dst)[i];
        const uint4 b = reinterpret_cast<const uint4*>(src)[i];
        a.x = __vmaxu4(a.x, b.x); a.y = __vmaxu4(a.y, b.y); a.z = __vmaxu4(a.z, b.z); a.w = __vmaxu4(a.w, b.w);
        reinterpret_cast<uint4*>(dst)[i] = a;
    }
    for (size_t i = nv * 16 + (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
        dst[i] = dst[i] > src[i] ? dst[i] : src[i];
}

// T4 (eval_wsi_segmentation.py:228,236-240): ds[y][x] = level0[ysrc[y]][xsrc[x]], 0 where a LUT entry < 0.
__global__ void __launch_bounds__(256) downsample_lut_kernel(const unsigned char* __restrict__ level0, int SW,
                                                             unsigned char* __restrict__ ds, int dh, int dw,
                                                             const int32_t* __restrict__ ysrc, const int32_t* __restrict__ xsrc) {
    const int y = blockIdx.y;
    if (y >= dh) return;
    const int ys = ysrc[y];
    for (int x = blockIdx.x * 256 + threadIdx.x; x < dw; x += gridDim.x * 256) {
        const int xs = xsrc[x];
        ds[(size_t)y * dw + x] = (ys >= 0 && xs >= 0) ? level0[(size_t)ys * SW + xs] : (unsigned char)0;
    }
}

// IOUEval.py:19-21 fast_hist: hist[n*gt + pred] += 1 for 0 <= gt < n  (pred >= n is a caller bug in the
// reference -- np.bincount would grow the array and reshape would throw; here such pixels are dropped).
__global__ void __launch_bounds__(256) confusion_hist_kernel(const unsigned char* __restrict__ pred, const unsigned char* __restrict__ gt,
                                                             size_t count, int n, unsigned long long* __restrict__ hist) {
    __shared__ unsigned int sh[32 * 32];
    for (int i = threadIdx.x; i < n * n; i += 256) sh[i] = 0;
    __syncthreads();
    // each CTA handles a bounded slice so the 32-bit shared counters cannot overflow
    const size_t per_cta = (count + gridDim.x - 1) / gridDim.x;
    const size_t lo = (size_t)blockIdx.x * per_cta;
    const size_t hi = lo + per_cta < count ? lo + per_cta : count;
    for (size_t i = lo + threadIdx.x; i < hi; i += 256) {
        const int g = gt[i], p = pred[i];
        if (g < n && p < n) atomicAdd(&sh[g * n + p], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n * n; i += 256)
        if (sh[i]) atomicAdd(hist + i, (unsigned long long)sh[i]);
}

}  // namespace espnet
