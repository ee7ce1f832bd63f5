/*
 * espnet_b200.h -- C ABI of libespnet_b200.so: hand-written sm_100a CUDA kernels for the ESPNet
 * glomerular-segmentation inference hot path of jinseikenai/glomeruli_segmentation.
 *
 * Nothing like this ABI exists in the reference (it is 100 % Python on top of PyTorch); each entry
 * point names the reference code it replaces (paths relative to the reference root).
 * Plain pointers and sizes only -- no torch types.  All device pointers are raw CUDA device
 * pointers owned by the caller; the library owns only its packed-weight buffer (two exceptions, both opt-in and documented at
 * their entry points: espnet_segment_host keeps grow-only device staging buffers and a private stream inside the handle,
 * espnet_peer_alloc hands out device memory that espnet_peer_free releases).  Every function
 * returns 0 (ESPNET_OK) or a negative ESPNET_E* code; espnet_last_error() gives the message.
 * `stream` is a cudaStream_t passed as void* (NULL = default stream); all forward / stitch calls
 * are asynchronous on it.  There is no CPU and no cuDNN fallback anywhere behind this header.
 */
#ifndef ESPNET_B200_H_
#define ESPNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ESPNET_API __attribute__((visibility("default")))
#else
#define ESPNET_API
#endif

#define ESPNET_OK 0
#define ESPNET_EINVAL (-1)   /* bad argument / null pointer                        */
#define ESPNET_ESHAPE (-2)   /* unsupported shape (H,W not multiples of 8, ...)    */
#define ESPNET_ECUDA (-3)    /* CUDA runtime error (message has the cudaError)     */
#define ESPNET_ESTATE (-4)   /* weights not packed / workspace too small           */
#define ESPNET_EMISSING (-5) /* a state_dict tensor is missing or has a wrong shape */

typedef struct espnet_handle espnet_t;

/* One state_dict entry: name exactly as in the reference checkpoint,
 * host pointer to contiguous fp32 data, shape.  num_batches_tracked entries may be omitted. */
typedef struct {
    const char* name;
    const float* data;   /* HOST pointer, fp32, contiguous */
    int ndim;
    int64_t shape[4];
} espnet_tensor_desc;

/* input formats of espnet_forward */
#define ESPNET_IN_F32_NCHW 0    /* normalised fp32 [B,3,H,W]  (what Model.py:341 receives)                    */
#define ESPNET_IN_U8_BGR_HWC 1  /* raw u8 [B,H,W,3] BGR + mean/std: P0 fused (VisualizeResults_iou.py:107-119) */
#define ESPNET_IN_U8_SLIDE 2    /* tiles read straight out of a resident u8 [SH,SW,3] slide at origins[B][2]   */

/* compute modes */
#define ESPNET_MODE_FP32 0      /* fp32-equivalent arithmetic, logits within 1e-3 of the reference.  Option "fp32_impl":
                                   1 (default) = the contractions of the DownSampler / ESP blocks on tcgen05 with 3-term
                                   fp16 operand splits (22-bit mantissa products, fp32 accumulation), everything else
                                   fp32 FMA; 0 = fp32 CUDA-core FMA everywhere                                        */
#define ESPNET_MODE_F16TC 1     /* single fp16 operands on tcgen05 (fp32 accumulate, fp32 storage): arg-max parity bar */

/* what the network is: full ESPNet (Model.py:306) or ESPNet-C encoder only (Model.py:242) */
#define ESPNET_NET_FULL 0
#define ESPNET_NET_ENCODER 1

typedef struct {
    /* ---- input ---- */
    const void* x;        /* device pointer, format in_fmt                                                  */
    int in_fmt;
    int B, H, W;          /* crops, crop height/width (multiples of 8)                                      */
    float mean[3];        /* BGR mean / std for the u8 formats (README.md:243-249)                          */
    float std_[3];
    const int32_t* origins; /* device [B][2] (x0,y0) for ESPNET_IN_U8_SLIDE (T1, detect_glomus_test.py:270-271) */
    int slide_h, slide_w;   /* slide size for ESPNET_IN_U8_SLIDE; pixels outside are 0 (openslide padding)      */
    /* ---- outputs (any may be NULL) ---- */
    float* logits;        /* fp32 NCHW [B,classes,H,W] (FULL) or [B,classes,H/8,W/8] (ENCODER)              */
    uint8_t* mask;        /* u8 [B,H,W] arg-max, ties -> lowest class (VisualizeResults_iou.py:128); for the
                             ENCODER net the logits are first x8 bilinearly up-sampled (:125-126, :258-261)  */
    float* prob_acc;      /* fp32 NCHW [B,classes,H,W]: softmax(logits) is ADDED here (ensemble extension)   */
    int prob_init;        /* 1: overwrite prob_acc instead of adding (first fold)                            */
    int mask_from_prob;   /* 1: mask = argmax(prob_acc after this call's add) instead of argmax(logits)      */
    /* ---- scratch ---- */
    void* workspace;      /* device, >= espnet_workspace_bytes(...), 256-byte aligned (cudaMalloc / torch are) */
    size_t workspace_bytes;
    void* stream;         /* cudaStream_t                                                                   */
} espnet_forward_args;

/* Replaces Model.ESPNet(classes,p,q) / Model.ESPNet_Encoder(classes,p,q) construction (Model.py:246,311).  Any classes in
 * [1, 48] (the reference's default 20 and the shipped checkpoints' 5 run compile-time specialised tail kernels, other counts
 * the generic run-time ones); p, q >= 1. */
ESPNET_API int espnet_create(int classes, int p, int q, int net, int device, espnet_t** out);
ESPNET_API void espnet_destroy(espnet_t* h);
ESPNET_API const char* espnet_last_error(const espnet_t* h); /* h may be NULL: last error of failed create */

/* Replaces load_state_dict (VisualizeResults_iou.py:267,279): folds eval-mode BN (eps 1e-3) into
 * per-channel scale/shift in fp64, re-lays conv weights for the kernels, uploads.  For the ENCODER net
 * names carry no "encoder." prefix (Model.py:242 key set), for the FULL net they do. */
ESPNET_API int espnet_pack_weights(espnet_t* h, const espnet_tensor_desc* tensors, int n);

ESPNET_API int espnet_set_mode(espnet_t* h, int mode);            /* ESPNET_MODE_* ; default FP32 */
/* Options: "fp32_impl" (see ESPNET_MODE_FP32); "branch_impl" (CUDA-core fp32 branch kernel: 0 auto, 1 per-thread global
 * loads, 2 TMA-staged shared-memory halo tiles); "tc_reduce" (tensor-core modes: 1 = reduce convs on tensor cores, 0 = CUDA
 * cores); "l2_reverse" (1x1 reduce walks its tiles against the
 * producer's order to start on the L2-resident part, default 1); "dec_impl" (decoder tail: 1 = 4 pixels per thread); "down_impl" (tensor-core 3x3-s2 reduce: 1 = TMA-staged input regions (default), 0 = per-thread global loads; bit-identical);
 * "tail_impl" (1 = generic run-time-class-count tail kernels even for 5 / 20
 * classes; bit-identical to the scalar specialised ones); "pdl" (programmatic dependent launch: a kernel of the
 * forward sets up its barriers / tensor memory / weights while the one before it drains and waits for it before touching
 * activations; results are identical.  -1 (default) = automatic: on for forwards of at most 32 x 512 x 512 pixels, where launch
 * gaps matter (batch 1: -16 %), off above; 0 = off; otherwise a bit mask of kernel classes: 1 stem, 2 (unused), 4 3x3-s2
 * reduce, 8 branch, 16 1x1 reduce, 32 head / decoder, 64 last kernel.  The environment variable ESPNET_B200_PDL sets the
 * default of a new handle). */
ESPNET_API int espnet_set_option(espnet_t* h, const char* key, int value);
ESPNET_API size_t espnet_workspace_bytes(const espnet_t* h, int B, int H, int W);

/* Replaces `img_out = model(img_variable)` (+ normalise before, arg-max after):
 * VisualizeResults_iou.py:107-128 -> Model.py:341-378 (FULL) / :273-304 (ENCODER). */
ESPNET_API int espnet_forward(espnet_t* h, const espnet_forward_args* a);

/* The same forward recorded once into a CUDA graph with THESE buffers (x, outputs, workspace stay owned by the caller and must
 * stay valid and in place); espnet_graph_launch replays it on `stream` with one launch instead of 26 - 29.  For the per-crop loop of
 * VisualizeResults_iou.py:100-129 (batch 1), where the forward is launch-latency bound.  Repacking weights drops all graphs. */
ESPNET_API int espnet_graph_capture(espnet_t* h, const espnet_forward_args* a, int* graph_id);
ESPNET_API int espnet_graph_launch(espnet_t* h, int graph_id, void* stream);
ESPNET_API int espnet_graph_destroy(espnet_t* h, int graph_id);

/* Debug / parity taps: copies an internal stage of the LAST forward to `dst` (device, fp32 NCHW).
 * stage names follow the reference module names ("b1","level2_0","level2.1","b2","level3_0",
 * "level3.7","b3", "up_l3", "combine_l2_l3", "up_l2", "conv", ...).  *count receives elements. */
ESPNET_API int espnet_read_stage(espnet_t* h, const char* stage, float* dst, size_t dst_elems, size_t* count, void* stream);

/* Per-kernel timing for the roofline report: while on, every kernel launch of espnet_forward is bracketed
 * by CUDA events on the launching stream.  espnet_get_profile synchronises on them and returns, per kernel
 * name (64-byte slots in `names`), the summed device time and the launch count since profiling was switched on. */
ESPNET_API int espnet_set_profiling(espnet_t* h, int on);
ESPNET_API int espnet_get_profile(espnet_t* h, char* names, float* total_ms, int* launches, int max_entries, int* n_entries);

/* Host-buffer convenience (the call a reference user makes per crop, batched): pageable or pinned
 * HOST u8 BGR crops in, HOST u8 masks out; H2D, forward, D2H and the stream sync all inside.
 * The ONE forward-path entry point that allocates: the handle keeps a device input buffer (B*H*W*3 bytes), a device mask buffer
 * (B*H*W) and a workspace (espnet_workspace_bytes) of the largest shape seen so far (grow-only, freed by espnet_destroy) and runs
 * on a private non-blocking stream.  Callers that manage device memory themselves use espnet_forward. */
ESPNET_API int espnet_segment_host(espnet_t* h, const uint8_t* crops_host, int B, int H, int W,
                        const float mean[3], const float std_[3], uint8_t* masks_host);

/* ---------------- tile -> slide stitching (eval_wsi_segmentation.py:162-316) ---------------- */

/* T3 (eval_wsi_segmentation.py:259-316, annotation_handler.py:74-105): slide[y,x] = max(slide[y,x], mask_b[..])
 * for every box b = boxes[b] = (x0,y0,x1,y1) int32 level-0 px (may overhang the slide), masks packed
 * back to back (box b at mask_offsets[b], row pitch x1-x0).  Rows >= y_limit are left untouched
 * (the reference's `ymax > slide_width` window skip, :194).  slide must be zero-initialised by the caller and 4-byte aligned
 * (the merge works on 32-bit words; the last slide_h*slide_w % 4 bytes are merged byte-wise, nothing past the end is touched). */
ESPNET_API int espnet_stitch_boxes(uint8_t* slide_mask, int slide_h, int slide_w, int y_limit,
                        const int32_t* boxes, const int64_t* mask_offsets, const uint8_t* masks,
                        int n_boxes, void* stream);

/* Same merge for a regular tile grid (T1 order: tile k = j*n_x + i at (i*stride_x, j*stride_y)),
 * gather form, no atomics: every slide pixel takes the max over the tiles covering it.  No limit on the slide height. */
ESPNET_API int espnet_stitch_grid(uint8_t* slide_mask, int slide_h, int slide_w, int y_limit,
                       const uint8_t* tile_masks, int n_x, int n_y, int win_x, int win_y,
                       int stride_x, int stride_y, int tile_row0, int tile_rows, void* stream);
/* The same into a BAND buffer [band_rows][slide_w] that holds slide rows [band_y0, band_y0 + band_rows): what a rank of a
 * multi-GPU run stitches from its own tiles without allocating the whole slide mask.  The rank's tiles are the
 * row-major index range [tile_k0, tile_k1) of the T1 grid (whole tile rows, or a balanced share that starts / ends inside a row);
 * tile_masks holds exactly those, tile k at (k - tile_k0) * win_x * win_y.
 * overwrite = 0: band = max(band, tiles) (zero-initialised or partly filled buffer).  overwrite = 1: every pixel of the rows the
 * tiles' rows cover is written (0 where no resident tile covers it) and nothing is read back: `band_mask` may then be ANOTHER
 * GPU's memory mapped into this process (espnet_peer_open) -- the stitch kernel places the band straight into rank 0's slide mask. */
ESPNET_API int espnet_stitch_grid_band(uint8_t* band_mask, int band_y0, int band_rows, int slide_h, int slide_w, int y_limit,
                       const uint8_t* tile_masks, int n_x, int n_y, int win_x, int win_y,
                       int stride_x, int stride_y, int tile_k0, int tile_k1, int overwrite, void* stream);
/* Peer-visible buffers for the multi-GPU slide stitch ("the stitch scatter kernel may write directly into a
 * peer-mapped slide buffer"): espnet_peer_alloc allocates `bytes` of zeroed device memory on `device` and returns its 64-byte
 * CUDA IPC handle; ANOTHER process maps it with espnet_peer_open for kernels running on ITS `device` (NVLink peer access is
 * enabled by the driver) and passes the mapped pointer to espnet_stitch_grid_band(..., overwrite = 1).  The owner frees with
 * espnet_peer_free after every mapper called espnet_peer_close.  The only allocation this library makes besides its packed
 * weights, and only on request. */
ESPNET_API int espnet_peer_alloc(size_t bytes, int device, void** dptr, uint8_t handle64[64]);
ESPNET_API int espnet_peer_open(const uint8_t handle64[64], int device, void** dptr);
ESPNET_API int espnet_peer_close(void* dptr, int device);
ESPNET_API int espnet_peer_free(void* dptr, int device);
/* dst[i] = max(dst[i], src[i]), n bytes: merge of the rows that two adjacent bands share (the tile-overlap strip) after the
 * band gather; the element-wise max of eval_wsi_segmentation.py:311-312. */
ESPNET_API int espnet_max_merge_u8(uint8_t* dst, const uint8_t* src, size_t n, void* stream);

/* T4 (eval_wsi_segmentation.py:225-240): ds[y,x] = level0[ysrc[y], xsrc[x]] (or 0 where the LUT is <0).
 * The LUTs are computed on the host in double (espnet_ds8_lut) so that cv2's INTER_NEAREST index
 * is reproduced bit-exactly. */
ESPNET_API int espnet_ds8_lut(int slide_len, int ws, int limit, int32_t* lut_host, int lut_len);
ESPNET_API int espnet_downsample_lut(const uint8_t* level0, int slide_h, int slide_w, uint8_t* ds, int ds_h, int ds_w,
                          const int32_t* ysrc_dev, const int32_t* xsrc_dev, void* stream);

/* IOUEval.py:19-21 fast_hist: hist[n*gt + pred] += 1 for gt in [0,n) ; hist is int64[n*n] on device (added to). */
ESPNET_API int espnet_confusion_hist(const uint8_t* pred, const uint8_t* gt, size_t count, int n_classes,
                          unsigned long long* hist_dev, void* stream);

/* ---------------- crop front-end and render ---------------- */

/* Host LUTs reproducing OpenCV's index arithmetic bit-exactly: bilinear source index + weight of the +1 neighbour
 * (cv2.resize INTER_LINEAR, generic code path) and nearest source index (INTER_NEAREST). */
ESPNET_API int espnet_bilinear_lut(int src_len, int dst_len, int32_t* idx_host, float* weight_host);
ESPNET_API int espnet_nearest_lut(int src_len, int dst_len, int32_t* idx_host);

/* VisualizeResults_iou.py:107-119 for crops whose size differs from the network input: u8 BGR [B,h,w,3] -> fp32 NCHW [B,3,H,W]
 * = cv2.resize(((float)p - mean) / std, (W, H)) / 255.  LUTs (device) from espnet_bilinear_lut(w, W) / (h, H). */
ESPNET_API int espnet_preprocess_resize(const uint8_t* crops, int B, int h, int w, const float mean[3], const float std_[3],
                                        const int32_t* xs_dev, const float* xf_dev, const int32_t* ys_dev, const float* yf_dev,
                                        float* out, int H, int W, void* stream);
/* make_seg_data.py:347-361 (output_org_files: level-0 read_region of every detected box) fused with the front-end above: crop b is
 * boxes_dev[b] = (x0,y0,x1,y1) int32 of the resident BGR u8 slide [slide_h][slide_w][3] (outside pixels are 0 like OpenSlide pads),
 * resized to W x H.  Box sizes differ, so the LUTs are per box: xs/xf [B][W] from espnet_bilinear_lut(x1-x0, W), ys/yf [B][H]. */
ESPNET_API int espnet_preprocess_resize_boxes(const uint8_t* slide, int slide_h, int slide_w, const int32_t* boxes_dev, int B,
                                              const float mean[3], const float std_[3], const int32_t* xs_dev, const float* xf_dev,
                                              const int32_t* ys_dev, const float* yf_dev, float* out, int H, int W, void* stream);
/* VisualizeResults_iou.py:129: class maps [B,sh,sw] -> [B,dh,dw], cv2 INTER_NEAREST (LUTs from espnet_nearest_lut). */
ESPNET_API int espnet_resize_nearest_u8(const uint8_t* src, int B, int sh, int sw, uint8_t* dst, int dh, int dw,
                                        const int32_t* ysrc_dev, const int32_t* xsrc_dev, void* stream);
/* VisualizeResults_iou.py:139-147: colour map (palette rows are r,g,b; written b,g,r) and cv2.addWeighted(img, .4, map, .6, 0).
 * img / outputs are u8 [npix][3]; either output may be NULL. */
ESPNET_API int espnet_palette_overlay(const uint8_t* img, const uint8_t* label, size_t npix, const uint8_t* palette_dev, int n_pal,
                                      uint8_t* color_out, uint8_t* overlay_out, void* stream);
/* eval_wsi_segmentation.py:225-240 (generate_whole_img over all windows): the /8 rendered slide u8 [ds_h][ds_w][3] from the
 * level-0 slide [SH][SW][3] and the level-0 class map, LUTs as for espnet_downsample_lut. */
ESPNET_API int espnet_render_ds8(const uint8_t* slide, const uint8_t* label, int slide_h, int slide_w, const uint8_t* palette_dev, int n_pal,
                                 uint8_t* out, int ds_h, int ds_w, const int32_t* ysrc_dev, const int32_t* xsrc_dev, void* stream);
/* VisualizeResults_iou.py:151-155: per-map class pixel counts, int64 [B][n_classes] on device (added to). */
ESPNET_API int espnet_class_counts(const uint8_t* maps, int B, size_t pix_per_map, int n_classes, unsigned long long* counts_dev, void* stream);

/* Hardware self-test of the tensor-core operand convention used by ESPNET_MODE_F16TC (no reference counterpart):
 * one 128 x nout x (8*nkc) tcgen05.mma whose A window is shifted by (dy, dx) pixels inside a TMA-staged (use_tma = 1)
 * or plainly copied (use_tma = 0) 48 x 48 activation box; *max_abs_err is measured against a host fp64 reference. */
ESPNET_API int espnet_tc_selftest(int device, int nkc, int nout, int dy, int dx, int use_tma, float* max_abs_err);

/* counters for the bench: number of kernel launches issued by this library since process start */
ESPNET_API unsigned long long espnet_launch_count(void);
ESPNET_API int espnet_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ESPNET_B200_H_ */
