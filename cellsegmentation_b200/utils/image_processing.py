"""Heatmaps, pseudo-masks and HSV refinement (mirror of utils/image_processing.py).

heatmap / generate_masks / preprocess_masks / remove_small_regions keep the reference
signatures (utils/image_processing.py:146, 79, 114, 14).  Painting, the gray map + JET colour
map + 0.5/0.5 blend, the HSV-threshold AND and the connected-component clean-up of
remove_small_regions all run on the GPU (paint.cu, hsv_refine.cu, cc.cu); only the PNG encode
(a host thread pool) and the CSV writing stay on the host.
"""
import csv
import os

import numpy as np
import torch

from .. import _capi, ops


def _cuda_dev():
    return torch.device("cuda", torch.cuda.current_device())


def _imsave(path, rgb):
    import cv2
    cv2.imwrite(path, np.ascontiguousarray(rgb[..., ::-1]) if rgb.ndim == 3 else rgb)


def remove_small_regions(img_bin, min_object_size, hole_area_threshold):
    """utils/image_processing.py:14-17: remove_small_objects then remove_small_holes (bool [H,W])."""
    d = torch.from_numpy(np.ascontiguousarray(np.asarray(img_bin) != 0).astype(np.uint8)).to(_cuda_dev())
    return ops.remove_small_regions(d, min_object_size, hole_area_threshold)[0].cpu().numpy().astype(bool)


def _xy_tensors(tiles, groups, dev):
    tiles = np.asarray(tiles).reshape(-1, 2).astype(np.int32)
    g = torch.from_numpy(np.ascontiguousarray(np.asarray(groups).astype(np.int32))).to(dev)
    x = torch.from_numpy(np.ascontiguousarray(tiles[:, 0])).to(dev)
    y = torch.from_numpy(np.ascontiguousarray(tiles[:, 1])).to(dev)
    return g, x, y


def _grid_instances(dataset, tiles, groups):
    """Instance indices (bag * T + grid position) of an explicit tile list when every tile lies on
    the dataset's regular grid (what rank() / sample() hand out), else None."""
    if not getattr(dataset, "has_tiles", False):
        return None
    grid = dataset._ensure_grid()
    tiles = np.asarray(tiles).reshape(-1, 2).astype(np.int64)
    groups = np.asarray(groups).astype(np.int64)
    n, T = len(dataset.images), len(grid)
    if len(tiles) == 0 or T == 0 or n * T >= 2 ** 31:
        return None
    H, W = int(dataset.image_size[0]), int(dataset.image_size[1])
    S, I = int(dataset.tile_size), int(dataset.interval)

    def inv(c, dim):
        last = dim - S
        ok = (c >= 0) & (c <= last) & ((c == last) | (c % I == 0))
        cnt = (last // I) + 1 + (1 if last % I else 0)
        return np.where(c == last, cnt - 1, c // I), ok, cnt

    gy, ok_y, _ = inv(tiles[:, 0], H)
    gx, ok_x, gw = inv(tiles[:, 1], W)
    if not (ok_y.all() and ok_x.all() and (groups >= 0).all() and (groups < n).all()):
        return None
    t = gy * gw + gx
    if not np.array_equal(np.asarray(grid)[t].astype(np.int64), tiles):
        return None
    return (groups * T + t).astype(np.int32)


def hsv_refine_batch(images_dev, masks_dev, v_thresh=170):
    """Lines 117-120 of preprocess_masks for a whole bag array on the device (u8 0/1)."""
    return ops.hsv_refine(images_dev, masks_dev, v_thresh)


def preprocess_masks(img, mask):
    """utils/image_processing.py:114-124 for one image: HSV-threshold AND and connected-component clean-up, both on the GPU."""
    dev = _cuda_dev()
    d_img = torch.from_numpy(np.ascontiguousarray(img, dtype=np.uint8)).to(dev)
    d_mask = torch.from_numpy(np.ascontiguousarray(np.asarray(mask) != 0).astype(np.uint8)).to(dev)
    refined = ops.hsv_refine(d_img, d_mask, 170)
    return ops.remove_small_regions(refined, 400, 120)[0].cpu().numpy().astype(bool)


def generate_masks(dataset, tiles, groups, preprocess, save_masks=True, output_path="./data/pseudomask"):
    """Transform predicted pos cell regions into binary masks (utils/image_processing.py:79-111)."""
    os.makedirs(os.path.join(output_path, "rgb"), exist_ok=True)
    os.makedirs(os.path.join(output_path, "mask"), exist_ok=True)
    dev = _cuda_dev()
    n = len(dataset.images)
    H, W = int(dataset.image_size[0]), int(dataset.image_size[1])
    g, x, y = _xy_tensors(tiles, groups, dev)
    masks_dev = ops.paint_mask_xy(g, x, y, n, H, W, dataset.tile_size)
    if preprocess:      # preprocess_masks for every bag at once (:100-102, :114-124)
        masks_dev = ops.hsv_refine(dataset.device_images(dev), masks_dev, 170)
        masks_dev = ops.remove_small_regions(masks_dev, 400, 120)
    pseudo_masks = masks_dev.cpu().numpy()
    for i, img in enumerate(dataset.images):
        if save_masks:
            _imsave(os.path.join(output_path, "rgb/{:05}.png".format(i + 1)), np.uint8(img))
            _imsave(os.path.join(output_path, "mask/{:05}.png".format(i + 1)), np.uint8(pseudo_masks[i] * 255))
    if save_masks:
        print("Original images & masks saved in \'{}\'.".format(output_path))
    return pseudo_masks


_JET_LUT = None


def _jet_lut(dev):
    """cv2's COLORMAP_JET table (a constant of the third-party library, read once on the host)."""
    global _JET_LUT
    if _JET_LUT is None:
        import cv2
        _JET_LUT = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(256, 1), cv2.COLORMAP_JET)
        _JET_LUT = np.ascontiguousarray(_JET_LUT.reshape(256, 3))
    return torch.from_numpy(_JET_LUT).to(dev)


def heatmap_arrays(testset, tiles, probs, groups, chunk=2048):
    """u8 [n,H,W,3] blended heatmaps of heatmap() without file output: paint (per-pixel max),
    gray, JET colour map and the 0.5/0.5 blend all run on the GPU (utils/image_processing.py:152-166)."""
    dev = _cuda_dev()
    n = len(testset.images)
    H, W = int(testset.image_size[0]), int(testset.image_size[1])
    p = torch.from_numpy(np.ascontiguousarray(np.asarray(probs, dtype=np.float32))).to(dev)
    heat = None
    inst = _grid_instances(testset, tiles, groups)
    if inst is not None:
        # tiles of the regular grid (rank() output): gather form, every pixel written once
        try:
            heat = ops.paint_heatmap_gather(torch.from_numpy(inst).to(dev), p, n, H, W, testset.tile_size,
                                            testset.interval)
        except _capi.CellSegError as e:
            if "shared memory" not in str(e):
                raise
    if heat is None:
        g, x, y = _xy_tensors(tiles, groups, dev)
        heat = ops.paint_heatmap_xy(g, x, y, p, n, H, W, testset.tile_size)
    lut = _jet_lut(dev)
    out = np.empty((n, H, W, 3), np.uint8)
    for b0 in range(0, n, chunk):
        b1 = min(n, b0 + chunk)
        imgs = torch.from_numpy(np.ascontiguousarray(np.stack([np.asarray(im) for im in testset.images[b0:b1]]))).to(dev)
        out[b0:b1] = ops.heatmap_blend(heat[b0:b1].contiguous(), imgs, lut).cpu().numpy()
    return out


def heatmap(testset, tiles, probs, groups, csv_file, output_path):
    """utils/image_processing.py:146-167: CSV row per kept tile + one blended PNG per image."""
    w = csv.writer(csv_file)
    for i, g in enumerate(groups):
        grid = list(map(int, tiles[i]))
        w.writerow([g, '{}'.format(grid), probs[i]])
    imgs = heatmap_arrays(testset, tiles, probs, groups)
    # PNG encoding releases the GIL inside cv2: write the files from a small thread pool
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as pool:
        list(pool.map(lambda i: _imsave(os.path.join(output_path, "test_{:05}.png".format(i + 1)), imgs[i]),
                      range(len(testset.images))))
