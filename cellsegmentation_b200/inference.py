"""Forward-only epoch loops of the Stage-2 path (mirror of inference.py of the reference).

inference_tiles / sample keep the reference signatures (inference.py:9-43).  When the loader's
dataset is one of this package's tile sets the per-tile __getitem__ / DataLoader-worker path
is bypassed: tiles are unfolded from the HBM-resident u8 bags inside the fused CUDA forward and
probabilities are produced in dataset order.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import ops
from .dataset import _TileSetBase


def _device_of(device):
    device = torch.device(device)
    if device.type != "cuda":
        raise ops._capi.CellSegError("inference_tiles needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())


def inference_tiles_device(dataset, model, device, want_features=False):
    """f32 CUDA tensor [N] of softmax(logits)[:,1] for every tile of `dataset`, dataset order;
    with want_features also the pooled encoder features f32 [N, F] (the input of fc_tile)."""
    device = _device_of(device)
    img = dataset.device_images(device)
    b0 = dataset.first_tile_bag
    n = dataset.num_tiles()
    clf = model.classifier(device)
    return clf.forward_tiles(img[b0:b0 + len(dataset._tile_bags)], dataset.tile_size, dataset.interval, 0, n,
                             precision=model.precision, max_batch=model.max_batch,
                             want_features=want_features)


def inference_tiles(loader, model, device, epoch=None, total_epochs=None, mode='train'):
    """Forward inference to obtain instance classification probs (inference.py:9-28)."""
    model.eval()
    ds = loader.dataset
    if isinstance(ds, _TileSetBase) and ds.mode in (1, "tile"):
        with torch.cuda.device(_device_of(device)):
            probs_dev = inference_tiles_device(ds, model, device)
            ds._last_probs = (probs_dev, probs_dev.cpu().numpy())
        return ds._last_probs[1]
    # foreign dataset: keep the reference loop, the model forward is still the CUDA path
    probs = torch.empty(len(ds), dtype=torch.float32)
    with torch.no_grad():
        for i, input in enumerate(loader):
            if mode == 'train':
                input = input[0]
            output = F.softmax(model(input.to(device)), dim=1)
            probs[i * loader.batch_size:i * loader.batch_size + input.size(0)] = output.detach()[:, 1].cpu()
    return probs.numpy()


def inference_image(loader, model, device, epoch=None, total_epochs=None, mode='train', cls_limit=False,
                    return_id=False):
    """Forward inference to obtain image-level categories and cell counts (inference.py:46-95).
    With this package's LystoTestset in mode "image" the whole images are normalised on the
    device; the model must be in mode "image" (fp32 encoder at 299 x 299 + fc_image heads)."""
    from .dataset import categorize, de_categorize
    model.eval()
    ds = loader.dataset
    ids, categories, counts = np.array(()), np.array(()), np.array(())

    def consume(output, batch_ids):
        nonlocal ids, categories, counts
        if batch_ids is not None:
            ids = np.concatenate((ids, batch_ids))
        output_cls = F.softmax(output[0], dim=1).detach().clone().cpu()
        output_reg = output[1].detach()[:, 0].clone().cpu()
        output_reg = np.round(output_reg.numpy()).astype(int)
        cat_labels = np.argmax(output_cls, axis=1)
        if cls_limit:
            for i, x in enumerate(output_reg):
                if categorize(x) > cat_labels[i]:
                    output_reg[i] = de_categorize(cat_labels[i])[1]
                elif categorize(x) < cat_labels[i]:
                    output_reg[i] = de_categorize(cat_labels[i])[0]
        categories = np.concatenate((categories, cat_labels))
        counts = np.concatenate((counts, output_reg))

    with torch.no_grad():
        if isinstance(ds, _TileSetBase) and ds.mode == "image":
            dev = _device_of(device)
            bs = loader.batch_size or 1
            with torch.cuda.device(dev):
                for b in range(0, len(ds.images), bs):
                    cnt = min(bs, len(ds.images) - b)
                    consume(model(ds.image_tensor(b, cnt, dev)), None if mode == 'train' else np.asarray(ds.id[b:b + cnt]))
        else:
            for data in loader:
                if mode == 'train':
                    consume(model(data[0].to(device)), None)
                else:
                    consume(model(data[1].to(device)), data[0])
    if return_id:
        return ids, categories, counts
    return categories, counts


def _probs_on_device(dataset, probs, device):
    last = getattr(dataset, "_last_probs", None)
    if last is not None and probs is last[1]:
        return last[0]                      # still resident from inference_tiles: no re-upload
    return torch.as_tensor(np.ascontiguousarray(probs, dtype=np.float32)).to(device)


def sample_indices(trainset, probs, tiles_per_pos, topk_neg, device=None):
    """order[index] and pseudo-labels of sample() (inference.py:34-40) computed on the GPU."""
    device = _device_of(device if device is not None else "cuda")
    with torch.cuda.device(device):
        p = _probs_on_device(trainset, probs, device)
        labels = torch.as_tensor(np.asarray(trainset.labels, dtype=np.int32)).to(device)
        off = torch.from_numpy(trainset.seg_offsets()).to(device)
        idx, pl, _ = ops.select_topk(p, labels, len(trainset.images), max(trainset.tiles_per_bag, 1),
                                     tiles_per_pos, topk_neg, seg_offsets=off)
        return idx.cpu().numpy().astype(np.int64), pl.cpu().numpy()


def sample(trainset, probs, tiles_per_pos, topk_neg, pos_neg_ratio):
    """Select top-k superpixels to create a instance training set (inference.py:31-43)."""
    idxs, pl = sample_indices(trainset, probs, tiles_per_pos, topk_neg)
    p, n = trainset.make_train_data(idxs, pos_neg_ratio, pseudo_labels=pl)
    print("Training data is sampled. (Pos samples: {} | Neg samples: {})".format(p, n))


def rank(testset, probs, threshold, device=None):
    """rank() of test_tile.py:63-79 / train_seg.py:234-247: (tiles, probs, groups) of the tiles
    with prob > threshold in np.lexsort((probs, groups)) order."""
    device = _device_of(device if device is not None else "cuda")
    with torch.cuda.device(device):
        p = _probs_on_device(testset, probs, device)
        off = torch.from_numpy(testset.seg_offsets()).to(device)
        idx, kp, _ = ops.rank_threshold(p, len(testset.images), max(testset.tiles_per_bag, 1), threshold,
                                        seg_offsets=off)
        idx = idx.cpu().numpy().astype(np.int64)
    T = testset.tiles_per_bag
    grid = testset._ensure_grid()
    groups = np.asarray(testset._tile_bags, np.int64)[idx // T] if len(idx) else np.zeros(0, np.int64)
    tiles = grid[idx % T].astype(np.int64) if len(idx) else np.zeros((0, 2), np.int64)
    return tiles, kp.cpu().numpy(), groups
