"""Drop-in datasets for the Stage-2 tile path (mirror of dataset/dataset.py of the reference).

Same class names, constructor arguments, attributes and methods as the reference's
`LystoDataset` (dataset/dataset.py:29-289) and `LystoTestset` (:346-435) as far as the tile
modes use them; the images additionally live in HBM as one u8 [Nb,299,299,3] tensor so the
CUDA path can unfold tiles on the device.  The tile grid is a closed-form function of
(interval, tile_size), so `tileIDX` / `tiles_grid` are lazy sequences instead of the
reference's Python lists of up to 60 M tuples ("host-side object model").

Behaviour kept on purpose: bag 0 of a LystoDataset owns no tiles
(`add_data(..., tileidx=0)` is falsy, dataset/dataset.py:142); pseudo-labels and the
shuffle/prune of make_train_data consume the global np.random state like the reference.
"""
import os

import numpy as np
import torch
from torch.utils.data import Dataset

from . import ops

patch_size = np.array([299, 299])


def get_tiles(image, interval, size):
    """Upper-left (row, col) of every sliding window (dataset/dataset.py:718-742)."""
    return [(int(x), int(y)) for x, y in ops.grid_coords(image.shape[0], image.shape[1], size, interval)]


def categorize(x):
    """7-class LYSTO count category (dataset/dataset.py:745-761)."""
    for cls, hi in enumerate((0, 5, 10, 20, 50, 200)):
        if x <= hi:
            return cls
    return 6


def de_categorize(label):
    """Count range of a category (dataset/dataset.py:764-780)."""
    return [(0, 0), (1, 5), (6, 10), (11, 20), (21, 50), (51, 200), (201, 100000)][min(int(label), 6)]


class _LazyTileIDX:
    """tileIDX: bag index of every tile; bags listed in `bags`, T tiles each."""

    def __init__(self, bags, T):
        self._bags, self._T = bags, T

    def __len__(self):
        return len(self._bags) * self._T

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        return self._bags[i // self._T]

    def __array__(self, dtype=None, copy=None):
        a = np.repeat(np.asarray(self._bags, dtype=np.int64), self._T)
        return a if dtype is None else a.astype(dtype)

    def __iter__(self):
        return iter(np.asarray(self))


class _LazyGrid:
    """tiles_grid: (row, col) of every tile; the same grid repeats for every bag."""

    def __init__(self, n_bags, grid):
        self._n, self._grid = n_bags, grid

    def __len__(self):
        return self._n * len(self._grid)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        x, y = self._grid[i % len(self._grid)]
        return (int(x), int(y))

    def __array__(self, dtype=None, copy=None):
        a = np.tile(np.asarray(self._grid), (self._n, 1))
        return a if dtype is None else a.astype(dtype)


class _TileSetBase(Dataset):
    """Shared storage: host images list + lazily uploaded device copy + uniform tile grid."""

    def _init_common(self, tile_size, interval):
        self.images = []
        self.organs = []
        self.interval = interval
        self.tile_size = tile_size
        self.image_size = patch_size
        self.mode = None
        self._tile_bags = []           # bag indices that own tiles, ascending
        self._dev = None               # u8 [Nb,H,W,3] on the GPU
        self._grid = None

    # ---- grid ---------------------------------------------------------------------------
    @property
    def has_tiles(self):
        return self.interval is not None and self.tile_size is not None

    def _ensure_grid(self):
        if self._grid is None and self.has_tiles:
            H, W = int(self.image_size[0]), int(self.image_size[1])
            self._grid = ops.grid_coords(H, W, self.tile_size, self.interval)
        return self._grid

    @property
    def tiles_per_bag(self):
        g = self._ensure_grid()
        return 0 if g is None else len(g)

    @property
    def tileIDX(self):
        return _LazyTileIDX(self._tile_bags, self.tiles_per_bag)

    @property
    def tiles_grid(self):
        g = self._ensure_grid()
        return _LazyGrid(len(self._tile_bags), g if g is not None else np.zeros((0, 2), np.int32))

    def num_tiles(self):
        return len(self._tile_bags) * self.tiles_per_bag

    @property
    def first_tile_bag(self):
        """Index of the first bag that owns tiles; tile-owning bags must be contiguous."""
        if not self._tile_bags:
            return 0
        b0 = self._tile_bags[0]
        if self._tile_bags != list(range(b0, b0 + len(self._tile_bags))):
            raise ValueError("tile-owning bags are not contiguous")
        return b0

    def seg_offsets(self):
        """i64 [n_images+1]: first tile index of every bag (bags without tiles are empty)."""
        n, T = len(self.images), self.tiles_per_bag
        cnt = np.zeros(n, np.int64)
        cnt[self._tile_bags] = T
        return np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)

    # ---- device residency ---------------------------------------------------------------
    def device_images(self, device=None):
        """u8 [Nb,H,W,3] tensor resident in HBM (uploaded once from pinned host memory)."""
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        device = torch.device(device)
        if self._dev is None or self._dev.device != device or self._dev.shape[0] != len(self.images):
            if isinstance(self.images, torch.Tensor) and self.images.is_cuda:
                self._dev = self.images
            else:
                host = torch.from_numpy(np.ascontiguousarray(np.stack([np.asarray(i) for i in self.images])))
                if host.dtype != torch.uint8:
                    raise TypeError("images must be uint8")
                self._dev = host.pin_memory().to(device, non_blocking=True)
        return self._dev

    def image_tensor(self, idx_begin, count, device=None):
        """Normalised whole images [count,3,H,W] (the per-image transform of the "image" modes,
        dataset/dataset.py:424-428): the bag is its own single tile."""
        img = self.device_images(device)
        H, W = int(img.shape[1]), int(img.shape[2])
        if H != W:
            raise ValueError("image modes need square bags")
        return ops.unfold_normalize(img, H, H, idx_begin, count)

    def tile_tensor(self, idx_begin, count, device=None):
        """Normalised tiles [count,3,S,S] of dataset indices idx_begin.. (mode 1 / 'tile')."""
        img = self.device_images(device)
        b0 = self.first_tile_bag
        return ops.unfold_normalize(img[b0:], self.tile_size, self.interval, idx_begin, count)


class LystoDataset(_TileSetBase):
    """Training / validation set; tile modes 1 (instance inference) and 3 (selected tiles)."""

    def __init__(self, filepath=None, tile_size=None, interval=None, train=True, organ=None,
                 augment=False, kfold=10, shuffle=False, num_of_imgs=0, _ensemble_init=False):
        super().__init__()
        self._init_common(tile_size, interval)
        if augment:
            raise NotImplementedError("augment=True belongs to Stage-1 image training (out of scope); "
                                      "train_tile.py builds its sets with augment=False")
        self.train = train
        self.organ = organ
        self.labels = []
        self.cls_labels = []
        self.transformIDX = []
        self.augment = augment
        self.train_index = None
        self.train_labels = None
        self._train_data = None
        if _ensemble_init:
            return
        if filepath is None or not os.path.exists(filepath):
            raise FileNotFoundError("Invalid data directory.")
        if kfold is not None and kfold <= 0:
            raise Exception("Invalid k-fold cross-validation argument.")
        self.kfold = kfold
        import h5py  # only needed for real LYSTO files
        f = h5py.File(filepath, "r")
        tileIDX = -1
        for i, (org, img, label) in enumerate(zip(f["organ"], f["x"], f["y"])):
            org = org.decode("utf-8")
            if num_of_imgs != 0 and i == num_of_imgs:
                break
            if self.kfold is not None:
                if (self.train and (i + 1) % self.kfold == 0) or (not self.train and (i + 1) % self.kfold != 0):
                    continue
            if self.organ is None or self.organ == org.partition("_")[0]:
                tileIDX += 1
                self.add_data(org, img, label, tileidx=tileIDX)
        assert len(self.labels) == len(self.images), "Mismatched number of labels and images."
        if shuffle:
            raise NotImplementedError("shuffle=True is only used by Stage-1 scripts (out of scope)")

    def add_data(self, organ, img, label, transidx=0, tileidx=None):
        """dataset/dataset.py:131-147 — note `and tileidx:` makes tileidx == 0 own no tiles."""
        self.organs.append(organ)
        self.images.append(img)
        self.labels.append(label)
        cls_label = categorize(label)
        self.cls_labels.append(cls_label)
        self.transformIDX.append(transidx)
        if self.interval is not None and self.tile_size is not None and tileidx:
            self._tile_bags.append(int(tileidx))
        self._dev = None
        return cls_label

    def setmode(self, mode):
        self.mode = mode

    def make_train_data(self, idxs, pos_neg_ratio=None, pseudo_labels=None):
        """dataset/dataset.py:166-201.  train_data is a structured (bag, x, y, label) array in
        the reference's row order: shuffled with the global np.random state, then the first n
        rows of the over-represented class are deleted."""
        idxs = np.asarray(idxs, np.int64)
        T = self.tiles_per_bag
        grid = self._ensure_grid()
        bags_arr = np.asarray(self._tile_bags, np.int64)
        if pseudo_labels is None:
            lab = (np.asarray(self.labels)[bags_arr[idxs // T]] != 0).astype(np.int64) if len(idxs) \
                else np.zeros(0, np.int64)
        else:
            lab = np.asarray(pseudo_labels, np.int64)
        pos = int(lab.sum())
        neg = int(len(lab) - pos)
        perm = np.arange(len(idxs))
        np.random.shuffle(perm)            # same draws as shuffling the reference's object rows
        keep = perm
        if pos_neg_ratio is not None:
            flag = n = None
            if pos > int(neg * pos_neg_ratio):
                flag, n, pos = 1, pos - int(neg * pos_neg_ratio), int(neg * pos_neg_ratio)
                print('Note: Positive superpixels are pruned to meet the pos_neg_ratio. ')
            elif neg > int(pos / pos_neg_ratio):
                flag, n, neg = 0, neg - int(pos / pos_neg_ratio), int(pos / pos_neg_ratio)
                print('Note: Negative superpixels are pruned to meet the pos_neg_ratio. ')
            if flag is not None:
                rows = np.nonzero(lab[perm] == flag)[0][:n]
                keep = np.delete(perm, rows)
        # the (bag, x, y, label) rows are materialised lazily, for the kept selection only: the
        # feature-cache epoch never looks at them (it trains on train_index / train_labels)
        self._train_data = None
        self.train_index = idxs[keep]      # dataset (tile) index of every train_data row
        self.train_labels = lab[keep]      # pseudo-label of every train_data row (int64)
        return pos, neg

    @property
    def train_data(self):
        """Structured (bag, x, y, label) rows in the reference's order; None before make_train_data."""
        if self._train_data is None and getattr(self, "train_index", None) is not None:
            self._train_data = self.rows_of(np.arange(len(self.train_index)))
        return self._train_data

    @train_data.setter
    def train_data(self, value):
        self._train_data = value
        if value is None:
            self.train_index = None
            self.train_labels = None

    def rows_of(self, rows):
        """train_data[rows] without materialising the whole table."""
        sel = self.train_index[rows]
        q, r = np.divmod(sel, max(self.tiles_per_bag, 1))
        td = np.empty(len(sel), dtype=[("bag", np.int32), ("x", np.int32), ("y", np.int32), ("label", np.int64)])
        if len(sel):
            g = self._ensure_grid()[r]
            td["bag"], td["x"], td["y"] = np.asarray(self._tile_bags, np.int64)[q], g[:, 0], g[:, 1]
            td["label"] = self.train_labels[rows]
        return td

    def train_tensor(self, begin, count, device=None):
        """Normalised tiles of train_data rows begin.. (mode 3), on the device."""
        td = self.train_data[begin:begin + count]
        img = self.device_images(device)
        dev = img.device
        t = ops.gather_normalize(img, self.tile_size, torch.from_numpy(td["bag"].copy()).to(dev),
                                 torch.from_numpy(td["x"].copy()).to(dev),
                                 torch.from_numpy(td["y"].copy()).to(dev))
        return t, torch.from_numpy(td["label"].copy()).to(dev)

    def __getitem__(self, idx):
        if self.mode == 1:
            assert self.num_tiles() > 0, "Dataset tile size and interval have to be settled for tile inference. "
            tile = self.tile_tensor(idx, 1)[0].cpu()
            return tile, self.labels[self.tileIDX[idx]]
        elif self.mode == 3:
            assert self.num_tiles() > 0, "Dataset tile size and interval have to be settled for tile-mode training. "
            t, lab = self.train_tensor(idx, 1)
            return t[0].cpu(), int(lab[0])
        elif self.mode in (2, 4, 5):
            raise NotImplementedError("image modes belong to Stage 1 / alternative training (out of scope)")
        raise Exception("Something wrong in setmode.")

    def __len__(self):
        assert self.mode is not None, "Something wrong in setmode."
        if self.mode == 1:
            assert self.num_tiles() > 0, "Dataset tile size and interval have to be settled for tile mode. "
            return self.num_tiles()
        elif self.mode == 2:
            return len(self.images)
        elif self.mode == 3:
            return len(self.train_data) if self._train_data is not None else len(self.train_index)
        return len(self.labels)

    @classmethod
    def from_arrays(cls, images, labels, tile_size, interval, organs=None):
        """Synthetic / in-memory construction through the same add_data path the reference's
        __init__ uses (tileidx counts from 0, so bag 0 owns no tiles)."""
        ds = cls(tile_size=tile_size, interval=interval, kfold=None, _ensemble_init=True)
        for i, (img, lab) in enumerate(zip(images, labels)):
            ds.add_data(organs[i] if organs else "synthetic_%d" % i, img, int(lab), tileidx=i)
        return ds


    @classmethod
    def from_device_tensor(cls, images, labels, tile_size, interval):
        """Bags that already live in HBM as one u8 [Nb,H,W,3] CUDA tensor (synthetic benchmarks):
        same bag / tile bookkeeping as from_arrays without a host copy of the images."""
        if not (getattr(images, "is_cuda", False) and images.dtype == torch.uint8):
            raise TypeError("images must be a CUDA uint8 tensor [Nb,H,W,3] (or a sliceable view of one)")
        ds = cls(tile_size=tile_size, interval=interval, kfold=None, _ensemble_init=True)
        n = int(images.shape[0])
        ds.images = images
        ds.organs = ["synthetic"] * n
        ds.labels = [int(v) for v in labels]
        ds.cls_labels = [categorize(v) for v in ds.labels]
        ds.transformIDX = [0] * n
        ds._tile_bags = list(range(1, n))        # add_data(tileidx=0) owns no tiles
        return ds


class LystoTestset(_TileSetBase):
    """Test set; mode "tile" (dataset/dataset.py:346-435).  Every bag owns tiles."""

    def __init__(self, filepath=None, tile_size=None, interval=None, organ=None, num_of_imgs=0,
                 _images=None, _organs=None):
        super().__init__()
        self._init_common(tile_size, interval)
        self.organ = organ
        self.id = []
        if _images is None:
            if filepath is None or not os.path.exists(filepath):
                raise FileNotFoundError("Invalid data directory.")
            import h5py
            f = h5py.File(filepath, "r")
            src = ((org.decode("utf-8"), img) for org, img in zip(f["organ"], f["x"]))
        else:
            src = zip(_organs or ["synthetic"] * len(_images), _images)
        tileIDX = -1
        for i, (org, img) in enumerate(src):
            if num_of_imgs != 0 and i == num_of_imgs:
                break
            if self.organ is None or self.organ == org.partition("_")[0]:
                tileIDX += 1
                self.id.append(i)
                self.organs.append(org)
                self.images.append(img)
                if self.has_tiles:
                    self._tile_bags.append(tileIDX)

    @classmethod
    def from_arrays(cls, images, tile_size, interval, organs=None):
        return cls(tile_size=tile_size, interval=interval, _images=images, _organs=organs)

    def setmode(self, mode):
        self.mode = mode

    def __getitem__(self, idx):
        if self.mode == "tile":
            assert self.num_tiles() > 0, "Dataset tile size and interval have to be settled for tile mode. "
            return self.tile_tensor(idx, 1)[0].cpu()
        elif self.mode == "image":                     # dataset/dataset.py:424-428
            return self.id[idx], self.image_tensor(idx, 1)[0].cpu()
        raise Exception("Something wrong in setmode.")

    def __len__(self):
        if self.mode == "tile":
            assert self.num_tiles() > 0, "Dataset tile size and interval have to be settled for tile mode. "
            return self.num_tiles()
        elif self.mode == "image":
            return len(self.images)
        raise Exception("Something wrong in setmode.")
