"""Tile grid and per-tile transform (oracle; test infrastructure only)."""
import numpy as np

MEAN = np.array([0.485, 0.456, 0.406], np.float32)
STD = np.array([0.229, 0.224, 0.225], np.float32)


def get_tiles(shape, interval, size):
    """Literal restatement of get_tiles (dataset/dataset.py:718-742) on an image shape."""
    H, W = shape[0], shape[1]
    tiles = []
    for x in np.arange(0, H - size + 1, interval):
        for y in np.arange(0, W - size + 1, interval):
            tiles.append((int(x), int(y)))
        if tiles[-1][1] + size != W:
            tiles.append((int(x), W - size))
    if tiles[-1][0] + size != H:
        for y in np.arange(0, W - size + 1, interval):
            tiles.append((H - size, int(y)))
        if tiles[-1][1] + size != W:
            tiles.append((H - size, W - size))
    return tiles


def normalize_tile(tile_u8):
    """ToTensor + Normalize (dataset/dataset.py:78-83, 391-397) in fp32: HWC u8 -> CHW f32."""
    t = tile_u8.astype(np.float32) / np.float32(255.0)      # ToTensor: float32 div 255
    t = (t - MEAN) / STD                                    # Normalize: sub_(mean).div_(std), fp32
    return np.ascontiguousarray(t.transpose(2, 0, 1))


def unfold(images, interval, size, inst_begin=0, inst_count=None):
    """Dataset-order tile batch (LystoTestset.__getitem__ 'tile', dataset/dataset.py:409-416)."""
    grid = get_tiles(images[0].shape, interval, size)
    T = len(grid)
    total = len(images) * T
    if inst_count is None:
        inst_count = total - inst_begin
    out = np.empty((inst_count, 3, size, size), np.float32)
    for j in range(inst_count):
        i = inst_begin + j
        x, y = grid[i % T]
        out[j] = normalize_tile(images[i // T][x:x + size, y:y + size])
    return out
