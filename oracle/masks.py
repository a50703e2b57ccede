"""Mask painting, HSV refinement, connected-component clean-up (oracle; test infrastructure only)."""
import numpy as np


def paint_masks(n_images, image_size, tile_size, tiles, groups):
    """generate_masks painting loop (utils/image_processing.py:89-98)."""
    pm = np.zeros((n_images, *image_size)).astype(np.uint8)
    for i in range(len(groups)):
        g0, g1 = int(tiles[i][0]), int(tiles[i][1])
        pm[groups[i]][g0:g0 + tile_size, g1:g1 + tile_size] = 1
    return pm


def paint_heatmaps(n_images, image_size, tile_size, tiles, probs, groups):
    """heatmap painting loop (utils/image_processing.py:153-158): float64 maps, last write wins."""
    masks = np.zeros((n_images, *image_size))
    for i, g in enumerate(groups):
        g0, g1 = int(tiles[i][0]), int(tiles[i][1])
        masks[g][g0:g0 + tile_size, g1:g1 + tile_size] = np.full((tile_size, tile_size), probs[i])
    return masks


def heat_to_gray(mask_f64):
    """255 - np.uint8(255 * masks[i]) (utils/image_processing.py:165)."""
    return 255 - np.uint8(255 * mask_f64)


def value_channel(img):
    """V of cv2.cvtColor(img, COLOR_BGR2HSV) for u8 input == max over channels."""
    return img.max(axis=-1)


def hsv_refine(img, mask, v_thresh=170):
    """preprocess_masks lines 117-120 (utils/image_processing.py:114-120), before the CC step:
    mask AND NOT(V > 170), restated in integers."""
    return np.logical_and(mask != 0, value_channel(img) <= v_thresh)


def hsv_refine_cv2(img, mask, v_thresh=170):
    """Same lines through the OpenCV calls the reference makes."""
    import cv2
    img_split = cv2.split(cv2.cvtColor(img, cv2.COLOR_BGR2HSV))
    _, mask_hsv = cv2.threshold(img_split[2], thresh=v_thresh, maxval=255, type=cv2.THRESH_BINARY)
    return np.logical_and(mask, (1 - mask_hsv / 255).astype(bool))


def bgr2hsv_u8(img):
    """Integer restatement of OpenCV's 8-bit BGR2HSV (hsv_shift 12, H range 180), SURVEY 8c."""
    a = img.reshape(-1, 3).astype(np.int64)
    b, g, r = a[:, 0], a[:, 1], a[:, 2]
    v = np.maximum(np.maximum(b, g), r)
    vmin = np.minimum(np.minimum(b, g), r)
    diff = v - vmin
    idx = np.arange(256, dtype=np.float64)
    with np.errstate(divide="ignore"):
        sdiv = np.rint((255 << 12) / idx)
        hdiv = np.rint((180 << 12) / (6.0 * idx))
    sdiv[0] = hdiv[0] = 0
    sdiv = sdiv.astype(np.int64)
    hdiv = hdiv.astype(np.int64)
    s = (diff * sdiv[v] + (1 << 11)) >> 12
    h = np.where(v == r, g - b, np.where(v == g, b - r + 2 * diff, r - g + 4 * diff))
    h = (h * hdiv[diff] + (1 << 11)) >> 12
    h = h + np.where(h < 0, 180, 0)
    return np.stack([h, s, v], axis=1).astype(np.uint8).reshape(img.shape)


def remove_small_objects(ar, min_size=64, connectivity=1, **_):
    """skimage 0.19.0 morphology.remove_small_objects for bool input (scipy restatement)."""
    from scipy import ndimage as ndi
    out = np.array(ar, dtype=bool, copy=True)
    if min_size == 0:
        return out
    ccs, _ = ndi.label(out, ndi.generate_binary_structure(out.ndim, connectivity))
    sizes = np.bincount(ccs.ravel())
    too_small = sizes < min_size
    out[too_small[ccs]] = False
    return out


def remove_small_holes(ar, area_threshold=64, connectivity=1, **_):
    """skimage 0.19.0 morphology.remove_small_holes (scipy restatement)."""
    out = np.logical_not(np.array(ar, dtype=bool))
    out = remove_small_objects(out, area_threshold, connectivity)
    return np.logical_not(out)


def remove_small_regions(img_bin, min_object_size, hole_area_threshold):
    """utils/image_processing.py:14-17."""
    img_bin = remove_small_objects(img_bin, min_size=min_object_size)
    return remove_small_holes(img_bin, area_threshold=hole_area_threshold)


def preprocess_masks(img, mask):
    """utils/image_processing.py:114-124."""
    m = hsv_refine(img, mask, 170)
    return remove_small_regions(m, min_object_size=400, hole_area_threshold=120)
