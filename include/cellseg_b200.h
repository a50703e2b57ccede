/*
 * cellseg_b200.h — C ABI of libcellseg_b200.so
 *
 * B200-native (sm_100a) kernels for the Stage-2 multiple-instance hot path of
 * Newiz430/CellSegmentation and the Stage-3 HSV mask refinement.  The reference
 * is pure Python and has no FFI of its own, so every entry point below names the
 * reference *function* it replaces (path:line under the reference tree); the
 * ctypes stubs that bind them live in cellsegmentation_b200/_capi.py and are
 * reproduced in INTEGRATION.md.
 *
 * Conventions (all entry points)
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the
 *     parameter name ends in `_host`
 *   - `stream` is a cudaStream_t passed as void*; work is asynchronous on it
 *   - no allocation inside the hot entry points: the caller owns all buffers;
 *     workspace sizes are queried first (cs_*_workspace_bytes)
 *   - return 0 (CS_OK) or a negative cs_status; cs_last_error() gives the text of
 *     the last failure on the calling thread
 *   - bags = images, instances = tiles; instance order is the reference's dataset
 *     order: bag-major, then get_tiles() row-major order inside a bag
 */
#ifndef CELLSEG_B200_H_
#define CELLSEG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum cs_status {
  CS_OK = 0,
  CS_ERR_INVALID_ARG = -1,
  CS_ERR_CUDA = -2,
  CS_ERR_WORKSPACE = -3,
  CS_ERR_UNSUPPORTED = -4,
  CS_ERR_DEVICE = -5
} cs_status;

/* Numeric mode of the tile-classifier forward. */
typedef enum cs_precision {
  CS_PREC_FP32 = 0, /* CUDA-core FFMA, fp32 everywhere: parity mode (<=1e-4 abs on probs) */
  CS_PREC_BF16 = 1  /* tcgen05 bf16 operands, fp32 accumulate in TMEM, hi/lo residual stream */
} cs_precision;

typedef enum cs_arch {
  CS_ARCH_RESNET18 = 0,        /* model/resnet.py:336-343   BasicBlock [2,2,2,2] */
  CS_ARCH_RESNET34 = 1,        /* model/resnet.py:346-352   BasicBlock [3,4,6,3] */
  CS_ARCH_RESNET50 = 2,        /* model/resnet.py:355-361   Bottleneck [3,4,6,3] */
  CS_ARCH_RESNEXT50_32X4D = 3, /* model/resnext.py:418-428  Bottleneck [3,4,6,3], groups 32 x 4 */
  CS_ARCH_RESNEXT101_32X8D = 4 /* model/resnext.py:431-443  Bottleneck [3,4,23,3], groups 32 x 8 */
} cs_arch;

/* Library version: major*10000 + minor*100 + patch. */
int cs_version(void);
/* Text of the last error raised on this thread ("" if none). Never NULL. */
const char* cs_last_error(void);
/* 0 if a compute-capability-10.x device is current, CS_ERR_DEVICE otherwise. */
int cs_check_device(void);

/* ---------------------------------------------------------------------------
 * Tile grid.  Replaces get_tiles(image, interval, size)
 * (dataset/dataset.py:718-742): upper-left coordinates 0, I, 2I, ... <= dim-S
 * plus one final row/col at dim-S when the stride does not land there.
 * cs_grid_count returns the number of grid positions along one axis
 * (tiles per bag = count(H) * count(W)); coordinate g is min(g*I, dim-S).
 * Host-only helpers, no GPU needed.
 * ------------------------------------------------------------------------- */
int cs_grid_count(int dim, int tile, int interval);
/* Fills xy_host[2*t] = row, xy_host[2*t+1] = col for all tiles of one HxW bag. */
int cs_grid_coords_host(int H, int W, int tile, int interval, int32_t* xy_host,
                        int64_t capacity_tiles);
/* Grid positions along one axis whose tile covers coordinate c (0 <= c < dim): the contiguous
 * range [*lo, *hi], empty (*lo > *hi) in the gap between tiles when interval > tile.  The
 * inverse of get_tiles used by cs_paint_heatmap_gather; host function, for tests. */
int cs_grid_cover_host(int c, int dim, int tile, int interval, int32_t* lo, int32_t* hi);

/* ---------------------------------------------------------------------------
 * K1  unfold + ToTensor + Normalize.
 * Replaces LystoTestset.__getitem__ mode "tile" / LystoDataset.__getitem__ mode 1
 * (dataset/dataset.py:409-416, 206-214) with the transform of :78-83 / :391-397:
 *   out[i][c][y][x] = ((float(u8)/255.f) - mean_c) / std_c        (fp32, NCHW)
 * img:  u8 [n_bags][H][W][3]  (dense, HWC)
 * instances inst_begin .. inst_begin+inst_count-1 of the bag range starting at
 * bag 0 of `img` (instance i -> bag i / T, tile i % T).
 * out:  f32 [inst_count][3][tile][tile]
 * ------------------------------------------------------------------------- */
int cs_unfold_normalize(const uint8_t* img, int n_bags, int H, int W, int tile,
                        int interval, int64_t inst_begin, int64_t inst_count,
                        float* out, void* stream);
/* Same transform for an explicit list of tiles (LystoDataset mode 3,
 * dataset/dataset.py:244-251): tile j is img[bag[j]][x[j]:x[j]+S, y[j]:y[j]+S]. */
int cs_gather_normalize(const uint8_t* img, int n_bags, int H, int W, int tile,
                        const int32_t* bag, const int32_t* x, const int32_t* y,
                        int64_t n_tiles, float* out, void* stream);

/* ---------------------------------------------------------------------------
 * K2  tile-classifier forward.  Replaces MILResNet.forward in tile mode under
 * model.eval() (model/resnet.py:250-269, 234-248, BasicBlock :28-43) followed by
 * softmax(dim=1)[:,1] (inference.py:24-27).
 *
 * The caller folds eval-mode BatchNorm into each conv in fp32
 * (W' = W*gamma/sqrt(var+eps), b' = beta - mean*gamma/sqrt(var+eps)) and passes
 * HOST pointers in network order:
 *   conv 0                 : stem 7x7/2            [64][3][7][7]
 *   per BasicBlock, in order: conv1 3x3, conv2 3x3, then (if the block has one)
 *                            the 1x1 downsample conv        [Cout][Cin][k][k]
 *   per Bottleneck, in order: conv1 1x1, conv2 3x3 (grouped: [W][W/groups][3][3], the torch
 *                            layout), conv3 1x1, then the 1x1 downsample conv if present
 *   fc_w [2][F], fc_b [2]  : fc_tile.1 (model/resnet.py:124-127), F = 512 (BasicBlock nets)
 *                            or 2048 (Bottleneck nets) = cs_model_feature_dim()
 * n_convs must equal the count implied by `arch` (resnet18: 20, resnet34: 36,
 * resnet50 / resnext50_32x4d: 53, resnext101_32x8d: 104).
 * The library packs GEMM-ready bf16 copies and keeps the fp32 originals.
 * ------------------------------------------------------------------------- */
typedef struct cs_model cs_model;

int cs_model_create(int arch, int n_convs, const float* const* conv_w_host,
                    const float* const* conv_b_host, const float* fc_w_host,
                    const float* fc_b_host, cs_model** out);
int cs_model_destroy(cs_model* m);
/* Length F of the pooled feature vector / fc_tile input (512 or 2048). */
int cs_model_feature_dim(const cs_model* m);
/* Replace fc_tile.1 (weight [2][F], bias [2], host pointers) after an optimizer step;
 * the encoder is frozen in tile mode (model/resnet.py:315-319) so nothing else changes. */
int cs_model_set_fc(cs_model* m, const float* fc_w_host, const float* fc_b_host);

/* Bytes of device workspace cs_model_forward_tiles needs for batches of up to
 * `max_batch` instances of `tile` x `tile` pixels at `precision`. */
int64_t cs_model_workspace_bytes(const cs_model* m, int tile, int64_t max_batch,
                                 int precision);

/* Fused unfold -> CNN -> prob for instances [inst_begin, inst_begin+inst_count)
 * of the bag array `img` (same instance numbering as cs_unfold_normalize).
 * prob_out[j] (fp32) receives softmax(logits)[1] of instance inst_begin + j.
 * feat_out (optional, may be NULL): fp32 [inst_count][F] pooled features
 * avgpool(x4)+maxpool(x4) (model/resnet.py:266), the input of fc_tile — used to
 * train fc_tile without a second encoder pass.
 * Internally processes the range in batches of <= max_batch the workspace was
 * sized for. */
int cs_model_forward_tiles(cs_model* m, const uint8_t* img, int n_bags, int H, int W,
                           int tile, int interval, int64_t inst_begin,
                           int64_t inst_count, int precision, float* prob_out,
                           float* feat_out, void* workspace, int64_t workspace_bytes,
                           int64_t max_batch, void* stream);

/* Same network on caller-materialised tiles (drop-in for model(x) with
 * x = f32 [n][3][tile][tile] already normalised): writes logits f32 [n][2]. */
int cs_model_forward_tensor(cs_model* m, const float* x, int64_t n, int tile,
                            int precision, float* logits_out, float* feat_out,
                            void* workspace, int64_t workspace_bytes, int64_t max_batch,
                            void* stream);

/* Number of kernels the last cs_model_forward_* call on this model launched. */
int64_t cs_model_last_launch_count(const cs_model* m);

/* ---------------------------------------------------------------------------
 * K3  segmented ordering and adaptive top-k selection.
 *
 * Segments: seg_offsets i64 [n_bags+1], non-decreasing, seg_offsets[0] = 0,
 * seg_offsets[n_bags] = N (bags may be empty — LystoDataset's bag 0 is,
 * dataset/dataset.py:142).  If seg_offsets is NULL every bag has `uniform_T`
 * instances.
 *
 * Order everywhere is numpy's  np.lexsort((probs, groups))  (inference.py:35,
 * test_tile.py:69, evaluate.py:13): ascending bag, then ascending prob, ties in
 * ascending instance index, NaN last, -0.0 == +0.0.
 * ------------------------------------------------------------------------- */

/* order_out i32 [N] = np.lexsort((probs, groups)). */
int cs_lexsort_segments(const float* prob, const int64_t* seg_offsets, int64_t uniform_T,
                        int n_bags, int32_t* order_out, void* stream);

/* Replaces sample() up to the call of make_train_data (inference.py:31-42) and
 * the pseudo-label rule of make_train_data (dataset/dataset.py:168-169):
 *   k_b = topk_neg if labels[b]==0 else labels[b]*tiles_per_pos
 *   keep sorted position i iff groups[i] != groups[(i + k_b) % N]     (literal,
 *   with wrap-around)
 * sel_idx_out   i32 [>= sum_b min(k_b, T_b)]: order[index], i.e. global instance
 *               indices ascending by (bag, prob, index)
 * sel_label_out u8, same length: 0 if labels[bag]==0 else 1
 * sel_offsets_out i64 [n_bags+1]: exclusive prefix of kept counts per bag;
 *               sel_offsets_out[n_bags] = M (total kept)
 * capacity: number of elements sel_idx_out / sel_label_out can hold; if M exceeds
 * it nothing beyond capacity is written and *M is still reported. */
int cs_select_topk(const float* prob, const int64_t* seg_offsets, int64_t uniform_T,
                   int n_bags, const int32_t* labels, int32_t tiles_per_pos,
                   int32_t topk_neg, int32_t* sel_idx_out, uint8_t* sel_label_out,
                   int64_t* sel_offsets_out, int64_t capacity, void* workspace,
                   int64_t workspace_bytes, void* stream);

/* Same for ONE SHARD of a larger bag set (bags partitioned across GPUs): the literal predicate
 * groups[i] != groups[(i + k) % N] of inference.py:37-40 is evaluated at GLOBAL positions, so the
 * shard passes the global index of its first tile and the global tile count N.  Inputs and the
 * returned indices are relative to the shard (index 0 = first tile of the shard).  With
 * global_total_tiles == 0 the shard is the whole set (= cs_select_topk). */
int cs_select_topk_shard(const float* prob, const int64_t* seg_offsets, int64_t uniform_T, int n_bags,
                         const int32_t* labels, int32_t tiles_per_pos, int32_t topk_neg,
                         int64_t global_tile_offset, int64_t global_total_tiles,
                         int32_t* sel_idx_out, uint8_t* sel_label_out, int64_t* sel_offsets_out,
                         int64_t capacity, void* workspace, int64_t workspace_bytes, void* stream);
/* Bytes of device workspace cs_select_topk needs (the list of bags the register-resident
 * fast path hands to the exact per-bag sort). */
int64_t cs_select_workspace_bytes(int n_bags);

/* Replaces rank() (test_tile.py:63-79, train_seg.py:234-247): keep prob > thr in
 * lexsort order.  Outputs as cs_select_topk; sel_prob_out f32 (optional). */
int cs_rank_threshold(const float* prob, const int64_t* seg_offsets, int64_t uniform_T,
                      int n_bags, float threshold, int32_t* sel_idx_out,
                      float* sel_prob_out, int64_t* sel_offsets_out, int64_t capacity,
                      void* stream);

/* ---------------------------------------------------------------------------
 * K4  mask painting and HSV refinement.
 * Tile j of a selection is instance sel_idx[j]; its bag and (row, col) follow
 * from the uniform grid (H, W, tile, interval) with T tiles per bag and
 * `bag_base` = index of the first bag that owns tiles (1 for LystoDataset, 0 for
 * LystoTestset): bag = bag_base + idx / T.
 * ------------------------------------------------------------------------- */

/* generate_masks() painting loop (utils/image_processing.py:91-98):
 * mask_out u8 [n_bags][H][W] must be zeroed by the caller; every kept tile sets
 * its tile x tile square to 1. */
int cs_paint_mask(const int32_t* sel_idx, int64_t n_sel, int H, int W, int tile,
                  int interval, int bag_base, int n_bags, uint8_t* mask_out,
                  void* stream);

/* heatmap() painting loop (utils/image_processing.py:153-158).  Tiles arrive in
 * ascending (bag, prob) order so the reference's last write wins = per-pixel max
 * of the kept covering probabilities.  heat_out f32 [n_bags][H][W], zeroed by the
 * caller; requires sel_prob >= 0. */
int cs_paint_heatmap(const int32_t* sel_idx, const float* sel_prob, int64_t n_sel, int H,
                     int W, int tile, int interval, int bag_base, int n_bags,
                     float* heat_out, void* stream);

/* The same map by a gather (no zero-fill of heat_out, no atomics on it, every pixel written
 * once): the kept tiles go into a dense table of T floats per bag in `workspace`
 * (cs_paint_heatmap_gather_workspace_bytes), then heat_out[b][y][x] = the maximum over the
 * kept tiles covering (y, x).  heat_out f32 [n_bags][H][W] is OVERWRITTEN for all n_bags maps
 * (0 where no kept tile covers a pixel).  CS_ERR_UNSUPPORTED when grid_rows x W floats exceed
 * 200 KB of shared memory (interval 1 on 299 x 299): use cs_paint_heatmap. */
int64_t cs_paint_heatmap_gather_workspace_bytes(int H, int W, int tile, int interval, int n_bags);
int cs_paint_heatmap_gather(const int32_t* sel_idx, const float* sel_prob, int64_t n_sel, int H,
                            int W, int tile, int interval, int bag_base, int n_bags,
                            float* heat_out, void* workspace, int64_t workspace_bytes,
                            void* stream);

/* Same two painting loops for explicit tile lists, the form the reference API passes
 * around (`tiles` [n][2] = (row, col) and `groups` [n] = bag): tiles outside the image
 * or bag range are ignored. */
int cs_paint_mask_xy(const int32_t* bag, const int32_t* x, const int32_t* y, int64_t n_sel,
                     int H, int W, int tile, int n_bags, uint8_t* mask_out, void* stream);
int cs_paint_heatmap_xy(const int32_t* bag, const int32_t* x, const int32_t* y,
                        const float* prob, int64_t n_sel, int H, int W, int tile, int n_bags,
                        float* heat_out, void* stream);

/* `255 - np.uint8(255 * masks[i])` (utils/image_processing.py:165) evaluated in
 * float64 exactly like numpy: gray_out u8 [n] = 255 - (uint8)(255.0 * (double)heat). */
int cs_heatmap_to_gray(const float* heat, int64_t n, uint8_t* gray_out, void* stream);

/* N3 -- heatmap() rendering tail, utils/image_processing.py:164-166, in one pass over the pixels:
 * gray = 255 - uint8(255*heat); cm = applyColorMap(gray, COLORMAP_JET) through the 256x3 LUT
 * `lut768` (device, entry g at lut768[3g..3g+2], channel order as cv2 returns it);
 * out = addWeighted(img, 0.5, cm, 0.5, 0) (round half to even).  heat f32 [n_px], img / out
 * u8 [n_px][3] (device).  heat 16-byte aligned, img / out 4-byte aligned. */
int cs_heatmap_blend(const float* heat, const uint8_t* img, const uint8_t* lut768, int64_t n_px,
                     uint8_t* out, void* stream);

/* preprocess_masks() lines 117-120 (utils/image_processing.py:114-120):
 *   V = max(R,G,B) (cv2.cvtColor(BGR2HSV) + split()[2]; channel-order free)
 *   out = (mask != 0) & !(V > v_thresh)          -> 0/1
 * img u8 [n_px][3], mask u8 [n_px], out u8 [n_px]; out may alias mask.
 * The reference constant is v_thresh = 170. */
int cs_hsv_refine(const uint8_t* img, const uint8_t* mask, int64_t n_px, int v_thresh,
                  uint8_t* out, void* stream);

/* Full 8-bit OpenCV RGB/BGR -> HSV (cv2.COLOR_BGR2HSV, H in [0,180)), bit-exact
 * integer restatement; hsv_out u8 [n_px][3].  Channel 0 of `img` is treated as B
 * (exactly what the reference does when it hands cv2 an RGB array). */
int cs_bgr2hsv_u8(const uint8_t* img, int64_t n_px, uint8_t* hsv_out, void* stream);

/* remove_small_regions() (utils/image_processing.py:14-17): skimage 0.19
 * remove_small_objects(min_size) then remove_small_holes(area_threshold), both 4-connected:
 * components with size < threshold are flipped.  mask u8 [n_bags][H][W] (non-zero =
 * foreground) is rewritten in place as 0/1.  A threshold of 0 skips that step. */
int64_t cs_cc_workspace_bytes(int n_bags, int H, int W);
int cs_remove_small_regions(uint8_t* mask, int n_bags, int H, int W, int min_object_size,
                            int hole_area_threshold, void* workspace, int64_t workspace_bytes,
                            void* stream);

/* ---------------------------------------------------------------------------
 * N4 — whole-image encoder and Stage-3 decoder building blocks (fp32 CUDA-core
 * path; standard-shape convs outside the tile hot path).
 * ------------------------------------------------------------------------- */

/* The encoder on whole square images (reference: MILResNet.forward modes "image" and
 * "segment", model/resnet.py:250-278; callers test_tile.py:88-105 --reg_limit,
 * train_seg.py:255-269, inference.py:46-95).  x: f32 [n][3][size][size] normalised (device).
 * feat_out (nullable): f32 [n][F] = avgpool_image(x4) + maxpool_image(x4) at map size 1.
 * x1_out..x4_out (nullable): NHWC f32 outputs of layer1..layer4 (resnet_forward(x, True)),
 * e.g. 75x75x64, 38x38x128, 19x19x256, 10x10x512 for ResNet-34 at 299.
 * Workspace: cs_model_workspace_bytes(m, size, max_batch, CS_PREC_FP32). */
int cs_model_forward_image(cs_model* m, const float* x, int64_t n, int size, float* feat_out,
                           float* x1_out, float* x2_out, float* x3_out, float* x4_out,
                           void* workspace, int64_t workspace_bytes, int64_t max_batch,
                           void* stream);

/* conv2d on NHWC f32 maps with BatchNorm folded by the caller (decoder layers upconv1..8 and
 * seg_out_conv, model/resnet.py:194-199, 280-303).  w_dev: [k*k*Cin][Cout], row index
 * (dy*k + dx)*Cin + ci; bias_dev [Cout]; Cout % 4 == 0.  out: [n][Ho][Wo][Cout]. */
int cs_conv2d_nhwc_f32(const float* in, int64_t n, int H, int W, int Cin, const float* w_dev,
                       const float* bias_dev, int Cout, int k, int stride, int pad, int relu,
                       float* out, void* stream);

/* F.interpolate(x, size=(Ho, Wo), mode="bilinear", align_corners=True) on NHWC f32 maps
 * (model/resnet.py:282, 287, 292, 297, 300). */
int cs_resize_bilinear_nhwc_f32(const float* in, int64_t n, int Hi, int Wi, int C, int Ho, int Wo,
                                float* out, void* stream);

/* ---------------------------------------------------------------------------
 * Diagnostics — used by tests/ to exercise the tcgen05 GEMM kernel and the
 * production conv planner in isolation.  Not part of the reference surface.
 * ------------------------------------------------------------------------- */

/* out_f32[M][N] = A[M][K] . B[N][K]^T + bias[N]; A, B bf16 row-major (device).
 * K % 64 == 0, K <= 2560, N % bn == 0, bn in {64, 128, 256}. */
int cs_debug_gemm_bf16(const void* a_bf16, const void* b_bf16, int64_t M, int N, int K,
                       const float* bias, int bn, float* out_f32, void* stream);

/* One k x k convolution (k = 3: pad 1, k = 1: pad 0; stride 1|2; `groups` groups, torch weight
 * layout [Cout][Cin/groups][k][k]) through the production planner:
 * in_hi bf16 [n][Hi*Wi][Cin] (device), w_host / bias_host fp32 (host)
 * -> out_f32 [n][Ho*Wo][Cout] (device), no ReLU.  n % 128 == 0.  reverse != 0 walks the work
 * items from the last to the first, as odd layers of the forward do.  Synchronises. */
int cs_debug_conv_bf16(const void* in_hi, int64_t n, int Hi, int Wi, int Cin, int Cout, int k,
                       int stride, int groups, const float* w_host, const float* bias_host,
                       int reverse, float* out_f32, void* stream);

/* The tile-32 tensor-core stem in isolation: img u8 [n_bags][H][W][3] (device), folded conv1
 * weights [64][3][7][7] and bias [64] fp32 (host) -> out_bf16 [inst_count][64 px][64 ch] (device)
 * = maxpool3x3/2(relu(conv7x7/2(normalise(tile)) + bias)) of instances inst_begin.. of the
 * uniform grid (model/resnet.py:236-239 on dataset/dataset.py:409-416 tiles).  Synchronises. */
int cs_debug_stem_bf16(const uint8_t* img, int n_bags, int H, int W, int interval, int64_t inst_begin,
                       int64_t inst_count, const float* w_host, const float* bias_host, void* out_bf16,
                       void* stream);

/* One layer-1 BasicBlock (8x8 maps, 64 channels, no downsample) through the production planner:
 * y = relu(conv2(relu(conv1(x) + b1)) + b2 + x) (model/resnet.py:28-43, eval-mode BN folded).
 * in_hi / out_bf16: bf16 [n][8*8][64] (device); w*_host [64][64][3][3], b*_host [64] fp32 (host);
 * n % 128 == 0.  *launches_out (nullable) receives the number of kernel launches used: 1 = the
 * fused block kernel, 2 = two y-sum convolutions.  Synchronises. */
int cs_debug_basic_block_bf16(const void* in_hi, int64_t n, const float* w1_host, const float* b1_host,
                              const float* w2_host, const float* b2_host, int reverse, void* out_bf16,
                              int* launches_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CELLSEG_B200_H_ */
