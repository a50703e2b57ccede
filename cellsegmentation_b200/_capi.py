"""ctypes binding of libcellseg_b200.so (include/cellseg_b200.h).

The library is the product: there is no Python/PyTorch fallback for any kernel.  If the
shared object is missing or a call fails, a CellSegError is raised.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_uint8, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libcellseg_b200.so")

CS_OK = 0
CS_PREC_FP32 = 0
CS_PREC_BF16 = 1
CS_ARCH = {"resnet18": 0, "resnet34": 1, "resnet50": 2, "resnext50_32x4d": 3, "resnext101_32x8d": 4}


class CellSegError(RuntimeError):
    """Raised when libcellseg_b200 is missing or reports an error."""


_lib = None

# name -> (restype, argtypes); mirrors include/cellseg_b200.h one to one.
_PROTOS = {
    "cs_version": (c_int, []),
    "cs_last_error": (c_char_p, []),
    "cs_check_device": (c_int, []),
    "cs_grid_count": (c_int, [c_int, c_int, c_int]),
    "cs_grid_coords_host": (c_int, [c_int, c_int, c_int, c_int, POINTER(c_int32), c_int64]),
    "cs_grid_cover_host": (c_int, [c_int, c_int, c_int, c_int, POINTER(c_int32), POINTER(c_int32)]),
    "cs_unfold_normalize": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int64, c_int64,
                                    c_void_p, c_void_p]),
    "cs_gather_normalize": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                    c_void_p, c_int64, c_void_p, c_void_p]),
    "cs_model_create": (c_int, [c_int, c_int, POINTER(c_void_p), POINTER(c_void_p), c_void_p,
                                c_void_p, POINTER(c_void_p)]),
    "cs_model_destroy": (c_int, [c_void_p]),
    "cs_model_feature_dim": (c_int, [c_void_p]),
    "cs_model_set_fc": (c_int, [c_void_p, c_void_p, c_void_p]),
    "cs_model_workspace_bytes": (c_int64, [c_void_p, c_int, c_int64, c_int]),
    "cs_model_forward_tiles": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                       c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p,
                                       c_int64, c_int64, c_void_p]),
    "cs_model_forward_tensor": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p,
                                        c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "cs_model_last_launch_count": (c_int64, [c_void_p]),
    "cs_model_forward_image": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "cs_conv2d_nhwc_f32": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int,
                                   c_int, c_int, c_int, c_void_p, c_void_p]),
    "cs_resize_bilinear_nhwc_f32": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                            c_void_p]),
    "cs_lexsort_segments": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "cs_select_topk": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int32, c_int32,
                               c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "cs_select_topk_shard": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int32, c_int32, c_int64,
                                     c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                                     c_void_p]),
    "cs_select_workspace_bytes": (c_int64, [c_int]),
    "cs_rank_threshold": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p, c_void_p,
                                  c_void_p, c_int64, c_void_p]),
    "cs_paint_mask": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_int,
                              c_void_p, c_void_p]),
    "cs_paint_heatmap": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int,
                                 c_int, c_void_p, c_void_p]),
    "cs_paint_mask_xy": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int,
                                 c_void_p, c_void_p]),
    "cs_paint_heatmap_gather_workspace_bytes": (c_int64, [c_int, c_int, c_int, c_int, c_int]),
    "cs_paint_heatmap_gather": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int,
                                        c_int, c_void_p, c_void_p, c_int64, c_void_p]),
    "cs_paint_heatmap_xy": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int,
                                    c_int, c_int, c_void_p, c_void_p]),
    "cs_heatmap_blend": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "cs_heatmap_to_gray": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "cs_hsv_refine": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "cs_bgr2hsv_u8": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "cs_cc_workspace_bytes": (c_int64, [c_int, c_int, c_int]),
    "cs_remove_small_regions": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int64,
                                        c_void_p]),
    "cs_debug_gemm_bf16": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_int,
                                   c_void_p, c_void_p]),
    "cs_debug_stem_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int64, c_int64, c_void_p, c_void_p,
                                   c_void_p, c_void_p]),
    "cs_debug_basic_block_bf16": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                          c_void_p, POINTER(c_int), c_void_p]),
    "cs_debug_conv_bf16": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                   c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
}

EXPORTED = tuple(_PROTOS)


def lib():
    """Loads (once) and returns the shared library; fails loudly if it was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CellSegError(
                "libcellseg_b200.so not found at %s — run `python -c 'import __graft_entry__ as g; "
                "g.build()'` (there is no CPU or PyTorch fallback)" % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc, what):
    if rc != CS_OK:
        msg = lib().cs_last_error()
        raise CellSegError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else ""))


def ptr(t):
    """Device (or host) data pointer of a torch tensor / numpy array, or None."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return c_void_p(t.data_ptr())
    return c_void_p(t.ctypes.data)


def cur_stream():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)
