"""One MIL epoch of Stage 2 across the GPUs of one box (train_tile.py:116-122 made data parallel).

    trainset.setmode(1); probs = inference_tiles(...)          -> every rank scores its own bags
    sample(trainset, probs, tiles_per_pos, topk_neg, ratio)    -> local top-k, ONE all-gather of
                                                                  the selected indices, identical
                                                                  make_train_data on every rank
    trainset.setmode(3); train_tile(...)                       -> every rank takes the rows of each
                                                                  global batch that fall into its
                                                                  own bag shard; fc_tile gradients
                                                                  are all-reduced (sum of 1/B-scaled
                                                                  partial gradients = global mean)
Feature cache: the encoder is frozen and runs with running BN statistics in both
passes (model/resnet.py:254-258, 315-319), and every kernel computes an instance independently
of its batch neighbours, so the pooled features of the scoring pass ARE the fc_tile inputs of the
training pass, bit for bit.  With cache_features the scoring pass keeps them (2 KB per instance
for ResNet-34) and training is fc_tile alone on gathered rows: no second encoder pass.

Rows are split by bag ownership in both modes (the gradient sum over the batch is the same
however the rows are split), so a rank only ever touches its own shard of the bag array: the
shard is uploaded once and cached (distributed.shard_dataset), features never move.

The reference's own --distributed switch cannot start; this is the working
equivalent.  Collectives per epoch: one all-gather of `1 + capacity` int64 words per rank
(selected indices + pseudo-labels, capacity closed-form from the count labels), one broadcast
of the shuffle seed when none is given, one 1 027-float all-reduce per training step.
No host synchronisation inside the step loop: the loss is accumulated on the device.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import ops
from .distributed import (allgather_selection, allreduce_flat, broadcast_seed, selection_capacity,
                          shard_dataset)
from .inference import inference_tiles_device


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def select_global(trainset, model, device, tiles_per_pos, topk_neg, cache_features=False, timing=None):
    """Scores this rank's shard and returns the GLOBAL selection (idx int64, pseudo-labels u8),
    identical on every rank and equal to the single-process sample() indices.  With
    cache_features a third value is returned: {"feat": f32 [n_local, F] on the device,
    "begin": first dataset index of this rank's shard, "end": one past the last}."""
    rank, world = _world()
    shard, tile_off = shard_dataset(trainset, rank, world)
    T = max(trainset.tiles_per_bag, 1)
    all_labels = np.asarray(trainset.labels, dtype=np.int64)[np.asarray(trainset._tile_bags, dtype=np.int64)] \
        if len(trainset._tile_bags) else np.zeros(0, np.int64)
    cap = selection_capacity(all_labels, T, tiles_per_pos, topk_neg, world)
    model.eval()
    ev = None
    with torch.cuda.device(device):
        if timing is not None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
        n_local = shard.num_tiles()
        feat = None
        if n_local > 0:
            probs = inference_tiles_device(shard, model, device, want_features=cache_features)
            if cache_features:
                probs, feat = probs
            if ev:
                ev[1].record()
            labels = torch.as_tensor(np.asarray(shard.labels, dtype=np.int32)).to(device)
            off = torch.from_numpy(shard.seg_offsets()).to(device)
            # the predicate of sample() wraps around the GLOBAL tile array (inference.py:37-40): a shard
            # evaluates it at its global position, or a one-bag shard would compare a bag with itself
            idx, pl, soff = ops.select_topk(probs, labels, len(shard.images), T, tiles_per_pos, topk_neg,
                                            seg_offsets=off, capacity=max(cap, 1), global_offset=tile_off,
                                            global_total=trainset.num_tiles(), sync=False)
            count = soff[-1:]
        else:
            if ev:
                ev[1].record()
            idx = torch.zeros(max(cap, 1), dtype=torch.int32, device=device)
            pl = torch.zeros(max(cap, 1), dtype=torch.uint8, device=device)
            count = torch.zeros(1, dtype=torch.int64, device=device)
        if cache_features and feat is None:
            feat = torch.zeros((0, model.fc_tile[1].in_features), dtype=torch.float32, device=device)
        if ev:
            ev[2].record()
        gidx, glab = allgather_selection(idx, pl, count, tile_off, cap, timing=timing)
        if ev:
            timing["score_events"] = (ev[0], ev[1])
            timing["select_events"] = (ev[1], ev[2])
    if cache_features:
        return gidx, glab, {"feat": feat, "begin": int(tile_off), "end": int(tile_off) + int(n_local)}
    return gidx, glab


def _is_mean_ce(criterion):
    return isinstance(criterion, torch.nn.CrossEntropyLoss) and criterion.reduction == "mean"


def _local_loss(criterion, out, label, n_global_rows, denom):
    """Loss of this rank's rows scaled so that the SUM over ranks is the criterion's value on the
    whole global batch.  CrossEntropyLoss 'mean' (what train_tile.py uses; class weights, label
    smoothing and ignore_index passed through): local sum / global normaliser.  Any other
    criterion is taken as a per-row mean and re-weighted by rows_here / rows_global."""
    if _is_mean_ce(criterion):
        s = torch.nn.functional.cross_entropy(out, label, weight=criterion.weight,
                                              ignore_index=criterion.ignore_index, reduction="sum",
                                              label_smoothing=criterion.label_smoothing)
        return s / denom
    return criterion(out, label) * (out.shape[0] / float(n_global_rows))


def train_selected(trainset, model, device, criterion, optimizer, batch_size, gamma=1.0, shuffle_seed=None,
                   feature_cache=None, scheduler=None, timing=None):
    """train_tile (train/train.py:12-48) over trainset.train_data with every global batch split
    across the ranks by bag ownership.  Returns the mean loss over the epoch (one host read at
    the end).  scheduler: stepped per batch when it is a CyclicLR / OneCycleLR, else once at the
    end, as the reference does (train/train.py:41-46)."""
    from torch.optim.lr_scheduler import CyclicLR, OneCycleLR
    rank, world = _world()
    model.train()
    n = len(trainset.train_index)
    labels_all = trainset.train_labels
    order = np.arange(n)
    if shuffle_seed is not None:
        order = np.random.RandomState(shuffle_seed).permutation(n)      # same on every rank
    shard, tile_off = shard_dataset(trainset, rank, world)
    if feature_cache is not None:
        begin, end = feature_cache["begin"], feature_cache["end"]
    else:
        begin, end = int(tile_off), int(tile_off) + shard.num_tiles()
    first_img = getattr(shard, "_first_img", 0)
    # tile mode trains fc_tile alone (model/resnet.py:315-319); the decoder's upconv5..8 also keep
    # requires_grad there but never receive a gradient -- they must not enter the all-reduce, or
    # zero gradients + weight decay would move them where the reference's optimizer skips them
    params = [p for p in model.fc_tile.parameters() if p.requires_grad] if hasattr(model, "fc_tile") else \
        [p for p in model.parameters() if p.requires_grad]
    dev = params[0].device
    loss_acc = torch.zeros((), dtype=torch.float64, device=dev)
    tile_num = 0
    class_w = None
    if _is_mean_ce(criterion) and criterion.weight is not None:
        class_w = criterion.weight.detach().cpu().numpy()
    for b in range(0, n, batch_size):
        rows = order[b:b + batch_size]
        ti = trainset.train_index[rows]
        mine = rows[(ti >= begin) & (ti < end)]
        # normaliser of the global batch: row count, or the weight sum of its valid labels when the
        # criterion is class-weighted (what reduction='mean' divides by); known on every rank
        denom = float(len(rows))
        if _is_mean_ce(criterion):
            lab_rows = labels_all[rows]
            valid = lab_rows[lab_rows != criterion.ignore_index]
            denom = float(class_w[valid].sum()) if class_w is not None else float(len(valid))
        denom = max(denom, 1e-30)
        optimizer.zero_grad()
        if len(mine):
            label = torch.from_numpy(labels_all[mine]).to(dev, non_blocking=True)
            if feature_cache is not None:
                sel = torch.from_numpy(trainset.train_index[mine] - begin).to(dev, non_blocking=True)
                out = model.fc_tile(feature_cache["feat"].index_select(0, sel))
            else:
                sub = trainset.rows_of(mine)
                img = shard.device_images(device)
                data = ops.gather_normalize(img, trainset.tile_size,
                                            torch.from_numpy((sub["bag"] - first_img).astype(np.int32)).to(dev),
                                            torch.from_numpy(sub["x"].copy()).to(dev),
                                            torch.from_numpy(sub["y"].copy()).to(dev))
                out = model(data, freeze_bn=True)
            loss = _local_loss(criterion, out, label, len(rows), denom) * gamma
            loss.backward()
            local = loss.detach().reshape(1).to(torch.float32)
        else:
            local = torch.zeros(1, dtype=torch.float32, device=dev)
        if world > 1:
            for p in params:
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
            flat = allreduce_flat([p.grad for p in params] + [local], timing=timing)
            o = 0
            for p in params:
                p.grad.copy_(flat[o:o + p.numel()].view_as(p))
                o += p.numel()
            local = flat[-1:]
        optimizer.step()
        if isinstance(scheduler, (CyclicLR, OneCycleLR)):
            scheduler.step()
        tile_num += len(rows)
        loss_acc += local[0].double() * len(rows)
    if not (scheduler is None or isinstance(scheduler, (CyclicLR, OneCycleLR))):
        scheduler.step()
    return float(loss_acc.item()) / max(tile_num, 1)


def mil_epoch(trainset, model, device, criterion, optimizer, tiles_per_pos, topk_neg, pos_neg_ratio,
              batch_size, gamma=1.0, seed=None, cache_features=True, scheduler=None, timing=None):
    """inference_tiles -> sample -> train_tile for one epoch; returns (mean loss, pos, neg).
    seed: np.random seed of make_train_data's shuffle / pruning.  Under world > 1 every rank must
    draw the same permutation, so rank 0's seed (given, or drawn from its np.random state when
    None) is broadcast."""
    trainset.setmode(1)
    cache = None
    if cache_features:
        gidx, glab, cache = select_global(trainset, model, device, tiles_per_pos, topk_neg, True, timing=timing)
    else:
        gidx, glab = select_global(trainset, model, device, tiles_per_pos, topk_neg, timing=timing)
    seed = broadcast_seed(seed, device)
    if seed is not None:
        np.random.seed(seed)          # make_train_data's shuffle / pruning must agree on all ranks
    pos, neg = trainset.make_train_data(gidx, pos_neg_ratio, pseudo_labels=glab)
    trainset.setmode(3)
    loss = train_selected(trainset, model, device, criterion, optimizer, batch_size, gamma,
                          feature_cache=cache, scheduler=scheduler, timing=timing)
    return loss, pos, neg
