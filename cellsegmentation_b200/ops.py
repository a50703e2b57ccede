"""Torch-tensor front end of the C ABI: one function per kernel family.

PyTorch is used for device memory and streams only; every computation below is a call
into libcellseg_b200.so on the current CUDA stream.
"""
import ctypes

import numpy as np
import torch

from . import _capi
from ._capi import CS_PREC_BF16, CS_PREC_FP32, check, cur_stream, lib, ptr

PRECISIONS = {"fp32": CS_PREC_FP32, "bf16": CS_PREC_BF16}


def _req_cuda(t, name, dtype=None):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise _capi.CellSegError("%s must be a CUDA tensor (no CPU fallback exists)" % name)
    if dtype is not None and t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t


def grid_count(dim, tile, interval):
    """Grid positions along one axis (dataset/dataset.py:728-740)."""
    return lib().cs_grid_count(int(dim), int(tile), int(interval))


def grid_coords(H, W, tile, interval):
    """int32 [T,2] (row, col) upper-left corners in get_tiles order (dataset/dataset.py:718-742)."""
    T = grid_count(H, tile, interval) * grid_count(W, tile, interval)
    if T <= 0:
        raise ValueError("bad tile geometry H=%d W=%d tile=%d interval=%d" % (H, W, tile, interval))
    out = np.empty((T, 2), np.int32)
    check(lib().cs_grid_coords_host(H, W, tile, interval,
                                    out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), T),
          "cs_grid_coords_host")
    return out


def unfold_normalize(img, tile, interval, inst_begin=0, inst_count=None):
    """img u8 [Nb,H,W,3] (cuda) -> f32 [n,3,tile,tile]: crop + ToTensor + Normalize."""
    _req_cuda(img, "img", torch.uint8)
    Nb, H, W, _ = img.shape
    T = grid_count(H, tile, interval) * grid_count(W, tile, interval)
    if inst_count is None:
        inst_count = Nb * T - inst_begin
    out = torch.empty((inst_count, 3, tile, tile), dtype=torch.float32, device=img.device)
    check(lib().cs_unfold_normalize(ptr(img), Nb, H, W, tile, interval, inst_begin, inst_count,
                                    ptr(out), cur_stream()), "cs_unfold_normalize")
    return out


def gather_normalize(img, tile, bag, x, y):
    """Tiles img[bag[j], x[j]:x[j]+S, y[j]:y[j]+S] -> f32 [n,3,S,S] (LystoDataset mode 3)."""
    _req_cuda(img, "img", torch.uint8)
    for t, nm in ((bag, "bag"), (x, "x"), (y, "y")):
        _req_cuda(t, nm, torch.int32)
    Nb, H, W, _ = img.shape
    n = bag.numel()
    out = torch.empty((n, 3, tile, tile), dtype=torch.float32, device=img.device)
    check(lib().cs_gather_normalize(ptr(img), Nb, H, W, tile, ptr(bag), ptr(x), ptr(y), n, ptr(out),
                                    cur_stream()), "cs_gather_normalize")
    return out


def _segs(seg_offsets, uniform_T, n_bags):
    if seg_offsets is not None:
        _req_cuda(seg_offsets, "seg_offsets", torch.int64)
        if seg_offsets.numel() != n_bags + 1:
            raise ValueError("seg_offsets must have n_bags+1 entries")
    return ptr(seg_offsets), int(uniform_T)


def lexsort_segments(prob, n_bags, uniform_T, seg_offsets=None):
    """np.lexsort((probs, groups)) for non-decreasing groups; int32 [N]."""
    _req_cuda(prob, "prob", torch.float32)
    so, T = _segs(seg_offsets, uniform_T, n_bags)
    order = torch.empty(prob.numel(), dtype=torch.int32, device=prob.device)
    check(lib().cs_lexsort_segments(ptr(prob), so, T, n_bags, ptr(order), cur_stream()),
          "cs_lexsort_segments")
    return order


def select_buffers(n_bags, capacity, device):
    """Pre-allocated outputs + workspace of select_topk: (idx i32 [capacity], pseudo-label u8
    [capacity], offsets i64 [n_bags+1], workspace u8)."""
    return (torch.empty(capacity, dtype=torch.int32, device=device),
            torch.empty(capacity, dtype=torch.uint8, device=device),
            torch.empty(n_bags + 1, dtype=torch.int64, device=device),
            torch.empty(max(int(lib().cs_select_workspace_bytes(n_bags)), 1), dtype=torch.uint8, device=device))


def select_topk(prob, labels, n_bags, uniform_T, tiles_per_pos, topk_neg, seg_offsets=None,
                capacity=None, global_offset=0, global_total=0, sync=True, out=None):
    """Adaptive top-k (inference.py:31-42). Returns (idx i32 [M], pseudo-label u8 [M], offsets i64).
    For one shard of a larger set pass the global index of its first tile and the global tile count
    (the predicate wraps around the GLOBAL array); indices stay relative to the shard.

    sync=False: no host round trip at all -- the full-capacity buffers are returned and the kept
    count M stays on the device as offsets[-1] (entries at and beyond M are unspecified; the
    kernels never write past `capacity`).  out: buffers from select_buffers() to write into."""
    _req_cuda(prob, "prob", torch.float32)
    _req_cuda(labels, "labels", torch.int32)
    so, T = _segs(seg_offsets, uniform_T, n_bags)
    if out is not None:
        idx, lab, off, ws = out
        capacity = idx.numel()
        if lab.numel() < capacity or off.numel() != n_bags + 1:
            raise ValueError("select_topk: out buffers do not match capacity / n_bags")
    else:
        if capacity is None:
            capacity = prob.numel()
        idx, lab, off, ws = select_buffers(n_bags, capacity, prob.device)
    check(lib().cs_select_topk_shard(ptr(prob), so, T, n_bags, ptr(labels), int(tiles_per_pos),
                                     int(topk_neg), int(global_offset), int(global_total), ptr(idx),
                                     ptr(lab), ptr(off), capacity, ptr(ws), ws.numel(), cur_stream()),
          "cs_select_topk_shard")
    if not sync:
        return idx, lab, off
    M = int(off[-1].item())
    if M > capacity:
        raise _capi.CellSegError("select_topk: %d kept instances exceed capacity %d" % (M, capacity))
    return idx[:M], lab[:M], off


def rank_threshold(prob, n_bags, uniform_T, threshold, seg_offsets=None, capacity=None):
    """rank() (test_tile.py:63-79): kept idx i32 [M], prob f32 [M], offsets i64 [n_bags+1]."""
    _req_cuda(prob, "prob", torch.float32)
    so, T = _segs(seg_offsets, uniform_T, n_bags)
    if capacity is None:
        capacity = prob.numel()
    idx = torch.empty(capacity, dtype=torch.int32, device=prob.device)
    sp = torch.empty(capacity, dtype=torch.float32, device=prob.device)
    off = torch.empty(n_bags + 1, dtype=torch.int64, device=prob.device)
    check(lib().cs_rank_threshold(ptr(prob), so, T, n_bags, float(threshold), ptr(idx), ptr(sp),
                                  ptr(off), capacity, cur_stream()), "cs_rank_threshold")
    M = int(off[-1].item())
    if M > capacity:
        raise _capi.CellSegError("rank_threshold: %d kept instances exceed capacity %d" % (M, capacity))
    return idx[:M], sp[:M], off


def paint_mask(sel_idx, n_bags, H, W, tile, interval, bag_base=0, out=None):
    """generate_masks painting loop: u8 [n_bags,H,W] with 1 inside every kept tile."""
    _req_cuda(sel_idx, "sel_idx", torch.int32)
    if out is None:
        out = torch.zeros((n_bags, H, W), dtype=torch.uint8, device=sel_idx.device)
    check(lib().cs_paint_mask(ptr(sel_idx), sel_idx.numel(), H, W, tile, interval, bag_base, n_bags,
                              ptr(out), cur_stream()), "cs_paint_mask")
    return out


def paint_heatmap(sel_idx, sel_prob, n_bags, H, W, tile, interval, bag_base=0, out=None):
    """heatmap painting loop: f32 [n_bags,H,W], per-pixel max of kept covering probs."""
    _req_cuda(sel_idx, "sel_idx", torch.int32)
    _req_cuda(sel_prob, "sel_prob", torch.float32)
    if out is None:
        out = torch.zeros((n_bags, H, W), dtype=torch.float32, device=sel_idx.device)
    check(lib().cs_paint_heatmap(ptr(sel_idx), ptr(sel_prob), sel_idx.numel(), H, W, tile, interval,
                                 bag_base, n_bags, ptr(out), cur_stream()), "cs_paint_heatmap")
    return out


def paint_heatmap_gather(sel_idx, sel_prob, n_bags, H, W, tile, interval, bag_base=0, out=None):
    """paint_heatmap as a gather: every pixel written once as the max over the kept covering tiles
    (no zero-fill, no atomics on the maps).  Raises CellSegError (unsupported) when the geometry
    does not fit shared memory -- callers fall back to paint_heatmap."""
    _req_cuda(sel_idx, "sel_idx", torch.int32)
    _req_cuda(sel_prob, "sel_prob", torch.float32)
    if out is None:
        out = torch.empty((n_bags, H, W), dtype=torch.float32, device=sel_idx.device)
    nbytes = int(lib().cs_paint_heatmap_gather_workspace_bytes(H, W, tile, interval, n_bags))
    ws = torch.empty(max(nbytes, 4), dtype=torch.uint8, device=sel_idx.device)
    check(lib().cs_paint_heatmap_gather(ptr(sel_idx), ptr(sel_prob), sel_idx.numel(), H, W, tile, interval,
                                        bag_base, n_bags, ptr(out), ptr(ws), nbytes, cur_stream()),
          "cs_paint_heatmap_gather")
    return out


def paint_mask_xy(bag, x, y, n_bags, H, W, tile, out=None):
    """paint_mask for explicit (bag, row, col) int32 arrays."""
    for t, nm in ((bag, "bag"), (x, "x"), (y, "y")):
        _req_cuda(t, nm, torch.int32)
    if out is None:
        out = torch.zeros((n_bags, H, W), dtype=torch.uint8, device=bag.device)
    check(lib().cs_paint_mask_xy(ptr(bag), ptr(x), ptr(y), bag.numel(), H, W, tile, n_bags, ptr(out),
                                 cur_stream()), "cs_paint_mask_xy")
    return out


def paint_heatmap_xy(bag, x, y, prob, n_bags, H, W, tile, out=None):
    """paint_heatmap for explicit (bag, row, col) int32 arrays and f32 probs."""
    for t, nm in ((bag, "bag"), (x, "x"), (y, "y")):
        _req_cuda(t, nm, torch.int32)
    _req_cuda(prob, "prob", torch.float32)
    if out is None:
        out = torch.zeros((n_bags, H, W), dtype=torch.float32, device=bag.device)
    check(lib().cs_paint_heatmap_xy(ptr(bag), ptr(x), ptr(y), ptr(prob), bag.numel(), H, W, tile,
                                    n_bags, ptr(out), cur_stream()), "cs_paint_heatmap_xy")
    return out


def heatmap_to_gray(heat):
    """255 - np.uint8(255 * heat) with the float64 product numpy performs."""
    _req_cuda(heat, "heat", torch.float32)
    out = torch.empty(heat.shape, dtype=torch.uint8, device=heat.device)
    check(lib().cs_heatmap_to_gray(ptr(heat), heat.numel(), ptr(out), cur_stream()),
          "cs_heatmap_to_gray")
    return out


def heatmap_blend(heat, img, lut):
    """heatmap() lines 164-166 fused: applyColorMap(255 - uint8(255*heat), JET) blended 0.5/0.5 with
    img.  heat f32 [n,H,W], img u8 [n,H,W,3], lut u8 [256,3] (cv2's colormap table) -> u8 [n,H,W,3]."""
    _req_cuda(heat, "heat", torch.float32)
    _req_cuda(img, "img", torch.uint8)
    _req_cuda(lut, "lut", torch.uint8)
    if tuple(img.shape) != tuple(heat.shape) + (3,) or lut.numel() != 768:
        raise ValueError("heatmap_blend: img %s / lut %s do not match heat %s"
                         % (tuple(img.shape), tuple(lut.shape), tuple(heat.shape)))
    out = torch.empty_like(img)
    check(lib().cs_heatmap_blend(ptr(heat), ptr(img), ptr(lut), heat.numel(), ptr(out), cur_stream()),
          "cs_heatmap_blend")
    return out


def hsv_refine(img, mask, v_thresh=170, out=None):
    """preprocess_masks lines 117-120: (mask != 0) & (max(R,G,B) <= v_thresh) as u8 0/1."""
    _req_cuda(img, "img", torch.uint8)
    _req_cuda(mask, "mask", torch.uint8)
    n_px = mask.numel()
    if img.numel() != 3 * n_px:
        raise ValueError("img must hold 3 bytes per mask pixel")
    if out is None:
        out = torch.empty_like(mask)
    check(lib().cs_hsv_refine(ptr(img), ptr(mask), n_px, int(v_thresh), ptr(out), cur_stream()),
          "cs_hsv_refine")
    return out


def remove_small_regions(mask, min_object_size, hole_area_threshold):
    """remove_small_regions (utils/image_processing.py:14-17) on u8 [n,H,W] masks, in place (0/1)."""
    _req_cuda(mask, "mask", torch.uint8)
    if mask.dim() == 2:
        mask = mask.unsqueeze(0)
    n, H, W = mask.shape
    nbytes = lib().cs_cc_workspace_bytes(n, H, W)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=mask.device)
    check(lib().cs_remove_small_regions(ptr(mask), n, H, W, int(min_object_size),
                                        int(hole_area_threshold), ptr(ws), nbytes, cur_stream()),
          "cs_remove_small_regions")
    return mask


def bgr2hsv(img):
    """cv2.cvtColor(img, cv2.COLOR_BGR2HSV) for u8 [...,3], bit-exact."""
    _req_cuda(img, "img", torch.uint8)
    out = torch.empty_like(img)
    check(lib().cs_bgr2hsv_u8(ptr(img), img.numel() // 3, ptr(out), cur_stream()), "cs_bgr2hsv_u8")
    return out


class TileClassifier:
    """Device-side tile classifier built from BN-folded fp32 conv weights.

    convs: list of (weight [Cout,Cin/groups,k,k], bias [Cout]) CPU float32 tensors in network order
    (stem; per BasicBlock conv1, conv2, [downsample]; per Bottleneck conv1, conv2, conv3,
    [downsample]); fc_w [2,F], fc_b [2] with F = 512 or 2048.
    """

    def __init__(self, arch, convs, fc_w, fc_b, device=None):
        if arch not in _capi.CS_ARCH:
            raise _capi.CellSegError("arch %r has no sm_100a kernels (%s)" % (arch, sorted(_capi.CS_ARCH)))
        self.arch = arch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        ws = [np.ascontiguousarray(w.detach().cpu().numpy(), dtype=np.float32) for w, _ in convs]
        bs = [np.ascontiguousarray(b.detach().cpu().numpy(), dtype=np.float32) for _, b in convs]
        fw = np.ascontiguousarray(fc_w.detach().cpu().numpy(), dtype=np.float32)
        fb = np.ascontiguousarray(fc_b.detach().cpu().numpy(), dtype=np.float32)
        n = len(ws)
        wp = (ctypes.c_void_p * n)(*[w.ctypes.data for w in ws])
        bp = (ctypes.c_void_p * n)(*[b.ctypes.data for b in bs])
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().cs_model_create(_capi.CS_ARCH[arch], n, wp, bp, fw.ctypes.data, fb.ctypes.data,
                                        ctypes.byref(handle)), "cs_model_create")
        self._h = handle
        self.feature_dim = int(lib().cs_model_feature_dim(handle))
        self._ws = None
        self._ws_key = None

    def set_fc(self, fc_w, fc_b):
        """Updates fc_tile.1 on the device (the only weights MIL tile training changes)."""
        fw = np.ascontiguousarray(fc_w.detach().cpu().numpy(), dtype=np.float32)
        fb = np.ascontiguousarray(fc_b.detach().cpu().numpy(), dtype=np.float32)
        with torch.cuda.device(self.device):
            check(lib().cs_model_set_fc(self._h, fw.ctypes.data, fb.ctypes.data), "cs_model_set_fc")

    def close(self):
        if getattr(self, "_h", None):
            lib().cs_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _workspace(self, tile, max_batch, precision):
        key = (tile, max_batch, precision)
        if self._ws_key != key:
            nbytes = lib().cs_model_workspace_bytes(self._h, tile, max_batch, precision)
            if nbytes <= 0:
                raise _capi.CellSegError("unsupported tile size %d" % tile)
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._ws_key = key
        return self._ws

    def forward_tiles(self, img, tile, interval, inst_begin=0, inst_count=None, precision="bf16",
                      max_batch=75776, want_features=False, prob_out=None):
        """Fused unfold -> CNN -> softmax[:,1] over instances of the resident u8 bag array."""
        _req_cuda(img, "img", torch.uint8)
        Nb, H, W, _ = img.shape
        T = grid_count(H, tile, interval) * grid_count(W, tile, interval)
        if inst_count is None:
            inst_count = Nb * T - inst_begin
        prec = PRECISIONS[precision]
        ws = self._workspace(tile, max_batch, prec)
        if prob_out is None:
            prob_out = torch.empty(inst_count, dtype=torch.float32, device=img.device)
        feat = torch.empty((inst_count, self.feature_dim), dtype=torch.float32, device=img.device) if want_features else None
        check(lib().cs_model_forward_tiles(self._h, ptr(img), Nb, H, W, tile, interval, inst_begin,
                                           inst_count, prec, ptr(prob_out), ptr(feat), ptr(ws),
                                           ws.numel(), max_batch, cur_stream()),
              "cs_model_forward_tiles")
        return (prob_out, feat) if want_features else prob_out

    def forward_tensor(self, x, precision="bf16", max_batch=75776, want_features=False, want_logits=True):
        """Drop-in for model(x): x f32 [n,3,S,S] normalised tiles -> logits f32 [n,2] (and / or the
        pooled features f32 [n,F]; want_logits=False skips the device fc, e.g. when fc_tile runs
        under autograd in torch)."""
        _req_cuda(x, "x", torch.float32)
        n, _, S, _ = x.shape
        prec = PRECISIONS[precision]
        ws = self._workspace(S, max_batch, prec)
        logits = torch.empty((n, 2), dtype=torch.float32, device=x.device) if want_logits else None
        feat = torch.empty((n, self.feature_dim), dtype=torch.float32, device=x.device) if want_features else None
        check(lib().cs_model_forward_tensor(self._h, ptr(x), n, S, prec, ptr(logits), ptr(feat),
                                            ptr(ws), ws.numel(), max_batch, cur_stream()),
              "cs_model_forward_tensor")
        if want_features and want_logits:
            return logits, feat
        return feat if want_features else logits

    def map_shapes(self, size):
        """[(H, C)] of x1..x4 for square inputs of `size` pixels (75/38/19/10 at 299)."""
        h = ((size - 1) // 2 + 1 - 1) // 2 + 1                   # conv 7x7/2 pad 3, max pool 3x3/2 pad 1
        exp = self.feature_dim // 512
        out = []
        for li in range(4):
            if li > 0:
                h = (h - 1) // 2 + 1
            out.append((h, 64 * exp << li))
        return out

    def forward_image(self, x, max_batch=4, want_features=True, want_maps=False):
        """N4: the encoder on whole normalised images x f32 [n,3,S,S] (fp32 CUDA-core path).
        Returns the pooled feature avgpool(x4)+maxpool(x4) f32 [n,F] and / or the NHWC maps
        [x1, x2, x3, x4] (model/resnet.py:234-248 with return_intermediate=True)."""
        _req_cuda(x, "x", torch.float32)
        n, _, S, S2 = x.shape
        if S != S2:
            raise ValueError("forward_image needs square images")
        ws = self._workspace(S, max_batch, CS_PREC_FP32)
        feat = torch.empty((n, self.feature_dim), dtype=torch.float32, device=x.device) if want_features else None
        maps = [torch.empty((n, h, h, c), dtype=torch.float32, device=x.device) for h, c in self.map_shapes(S)] \
            if want_maps else [None] * 4
        check(lib().cs_model_forward_image(self._h, ptr(x), n, S, ptr(feat), ptr(maps[0]), ptr(maps[1]),
                                           ptr(maps[2]), ptr(maps[3]), ptr(ws), ws.numel(), max_batch,
                                           cur_stream()), "cs_model_forward_image")
        if want_features and want_maps:
            return feat, maps
        return maps if want_maps else feat

    @property
    def last_launch_count(self):
        return int(lib().cs_model_last_launch_count(self._h))


def conv2d_nhwc(x, w_packed, bias, k, stride=1, pad=0, relu=False):
    """conv2d on NHWC f32 maps: x [n,H,W,Cin], w_packed [k*k*Cin, Cout] (row (dy*k+dx)*Cin+ci),
    bias [Cout] (BatchNorm folded by the caller) -> [n,Ho,Wo,Cout].  Decoder layers of N4."""
    _req_cuda(x, "x", torch.float32)
    _req_cuda(w_packed, "w_packed", torch.float32)
    _req_cuda(bias, "bias", torch.float32)
    n, H, W, Cin = x.shape
    Cout = w_packed.shape[1]
    if w_packed.shape[0] != k * k * Cin or bias.numel() != Cout or Cout % 4:
        raise ValueError("conv2d_nhwc: weight %s / bias %s do not match Cin %d, k %d (Cout %% 4 == 0)"
                         % (tuple(w_packed.shape), tuple(bias.shape), Cin, k))
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    out = torch.empty((n, Ho, Wo, Cout), dtype=torch.float32, device=x.device)
    check(lib().cs_conv2d_nhwc_f32(ptr(x), n, H, W, Cin, ptr(w_packed), ptr(bias), Cout, k, stride, pad,
                                   int(bool(relu)), ptr(out), cur_stream()), "cs_conv2d_nhwc_f32")
    return out


def resize_bilinear_nhwc(x, size):
    """F.interpolate(x, size=size, mode="bilinear", align_corners=True) on an NHWC f32 map."""
    _req_cuda(x, "x", torch.float32)
    n, H, W, C = x.shape
    out = torch.empty((n, size, size, C), dtype=torch.float32, device=x.device)
    check(lib().cs_resize_bilinear_nhwc_f32(ptr(x), n, H, W, C, size, size, ptr(out), cur_stream()),
          "cs_resize_bilinear_nhwc_f32")
    return out


def debug_gemm_bf16(a, b, bias, bn):
    """A [M,K] bf16, B [N,K] bf16, bias f32 [N] -> f32 [M,N] through the tcgen05 kernel."""
    _req_cuda(a, "a", torch.bfloat16)
    _req_cuda(b, "b", torch.bfloat16)
    _req_cuda(bias, "bias", torch.float32)
    M, K = a.shape
    N = b.shape[0]
    out = torch.zeros((M, N), dtype=torch.float32, device=a.device)
    check(lib().cs_debug_gemm_bf16(ptr(a), ptr(b), M, N, K, ptr(bias), bn, ptr(out), cur_stream()),
          "cs_debug_gemm_bf16")
    return out


def debug_stem_bf16(images, w, bias, interval, inst_begin=0, inst_count=None):
    """images u8 [B,H,W,3] (cuda), folded conv1 w f32 [64,3,7,7] / bias f32 [64] (cpu) -> bf16
    [count,8,8,64]: the tile-32 tensor-core stem (normalise, conv 7x7/2, bias, ReLU, maxpool 3x3/2)."""
    _req_cuda(images, "images", torch.uint8)
    B, H, W, _ = images.shape
    T = grid_count(H, 32, interval) * grid_count(W, 32, interval)
    if inst_count is None:
        inst_count = B * T - inst_begin
    wn = np.ascontiguousarray(w.detach().cpu().numpy(), dtype=np.float32)
    bn_ = np.ascontiguousarray(bias.detach().cpu().numpy(), dtype=np.float32)
    out = torch.zeros((inst_count, 8, 8, 64), dtype=torch.bfloat16, device=images.device)
    check(lib().cs_debug_stem_bf16(ptr(images), B, H, W, interval, inst_begin, inst_count, wn.ctypes.data,
                                   bn_.ctypes.data, ptr(out), cur_stream()), "cs_debug_stem_bf16")
    return out


def debug_basic_block_bf16(x_hi, w1, b1, w2, b2, reverse=False):
    """x_hi bf16 [n,8,8,64] (cuda), w1/w2 f32 [64,64,3,3], b1/b2 f32 [64] (cpu) -> (bf16 [n,8,8,64],
    launches): relu(conv2(relu(conv1(x)+b1))+b2+x) through the production planner (layer-1 BasicBlock)."""
    import ctypes
    _req_cuda(x_hi, "x_hi", torch.bfloat16)
    n = x_hi.shape[0]
    arrs = [np.ascontiguousarray(t.detach().cpu().numpy(), dtype=np.float32) for t in (w1, b1, w2, b2)]
    out = torch.zeros_like(x_hi)
    launches = ctypes.c_int(0)
    check(lib().cs_debug_basic_block_bf16(ptr(x_hi), n, arrs[0].ctypes.data, arrs[1].ctypes.data,
                                          arrs[2].ctypes.data, arrs[3].ctypes.data, int(bool(reverse)), ptr(out),
                                          ctypes.byref(launches), cur_stream()), "cs_debug_basic_block_bf16")
    return out, launches.value


def debug_conv_bf16(x_hi, w, bias, stride, groups=1, reverse=False):
    """x_hi bf16 [n,H,W,Cin] (cuda), w f32 [Cout,Cin/groups,k,k] (cpu), bias f32 [Cout] (cpu)
    -> f32 [n,Ho,Wo,Cout] through the production planner + tcgen05 kernels (k = 1 or 3)."""
    _req_cuda(x_hi, "x_hi", torch.bfloat16)
    n, H, W, Cin = x_hi.shape
    Cout, k = w.shape[0], w.shape[2]
    pad = 1 if k == 3 else 0
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    wn = np.ascontiguousarray(w.detach().cpu().numpy(), dtype=np.float32)
    bn_ = np.ascontiguousarray(bias.detach().cpu().numpy(), dtype=np.float32)
    out = torch.zeros((n, Ho, Wo, Cout), dtype=torch.float32, device=x_hi.device)
    check(lib().cs_debug_conv_bf16(ptr(x_hi), n, H, W, Cin, Cout, k, stride, groups, wn.ctypes.data,
                                   bn_.ctypes.data, int(bool(reverse)), ptr(out), cur_stream()),
          "cs_debug_conv_bf16")
    return out
