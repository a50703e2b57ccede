"""One MIL epoch of Stage 2 across the GPUs of one box (train_tile.py:116-122 made data parallel).

    trainset.setmode(1); probs = inference_tiles(...)          -> every rank scores its own bags
    sample(trainset, probs, tiles_per_pos, topk_neg, ratio)    -> local top-k, all-gather of the
                                                                  selected indices, identical
                                                                  make_train_data on every rank
    trainset.setmode(3); train_tile(...)                       -> every rank takes a slice of each
                                                                  global batch; fc_tile gradients
                                                                  are all-reduced (mean)
Feature cache: the encoder is frozen and runs with running BN statistics in both
passes (model/resnet.py:254-258, 315-319), and every kernel computes an instance independently
of its batch neighbours, so the pooled features of the scoring pass ARE the fc_tile inputs of the
training pass, bit for bit.  With cache_features the scoring pass keeps them (2 KB per instance
for ResNet-34) and training is fc_tile alone on gathered rows: no second encoder pass.  Each
rank then trains on the rows of a global batch that fall into its own bag shard (the gradient
sum over the batch is the same however the rows are split).

The reference's own --distributed switch cannot start; this is the working
equivalent: bags are sharded in contiguous blocks, the only collectives are the all-gather(v) of
selected indices / pseudo-labels and one 1 026-float gradient all-reduce per step.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import ops
from .distributed import allgather_selection, allreduce_mean_grads, shard_dataset
from .inference import inference_tiles_device


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def select_global(trainset, model, device, tiles_per_pos, topk_neg, cache_features=False):
    """Scores this rank's shard and returns the GLOBAL selection (idx int64, pseudo-labels u8),
    identical on every rank and equal to the single-process sample() indices.  With
    cache_features a third value is returned: {"feat": f32 [n_local, F] on the device,
    "begin": first dataset index of this rank's shard, "end": one past the last}."""
    rank, world = _world()
    shard, tile_off = shard_dataset(trainset, rank, world)
    model.eval()
    with torch.cuda.device(device):
        if shard.num_tiles() > 0:
            probs = inference_tiles_device(shard, model, device, want_features=cache_features)
            if cache_features:
                probs, feat = probs
            labels = torch.as_tensor(np.asarray(shard.labels, dtype=np.int32)).to(device)
            off = torch.from_numpy(shard.seg_offsets()).to(device)
            # the predicate of sample() wraps around the GLOBAL tile array (inference.py:37-40): a shard
            # evaluates it at its global position, or a one-bag shard would compare a bag with itself
            idx, pl, _ = ops.select_topk(probs, labels, len(shard.images), max(shard.tiles_per_bag, 1),
                                         tiles_per_pos, topk_neg, seg_offsets=off,
                                         global_offset=tile_off, global_total=trainset.num_tiles())
        else:
            idx = torch.zeros(0, dtype=torch.int32, device=device)
            pl = torch.zeros(0, dtype=torch.uint8, device=device)
            feat = torch.zeros((0, model.fc_tile[1].in_features), dtype=torch.float32, device=device)
        gidx, glab = allgather_selection(idx, pl, tile_off)
    gidx, glab = gidx.cpu().numpy().astype(np.int64), glab.cpu().numpy()
    if cache_features:
        return gidx, glab, {"feat": feat, "begin": int(tile_off), "end": int(tile_off) + int(feat.shape[0])}
    return gidx, glab


def train_selected(trainset, model, device, criterion, optimizer, batch_size, gamma=1.0, shuffle_seed=None,
                   feature_cache=None):
    """train_tile over trainset.train_data with every global batch split across the ranks
    (round robin, or by bag ownership when the rows' features come from feature_cache)."""
    rank, world = _world()
    model.train()
    td = trainset.train_data
    n = len(td)
    order = np.arange(n)
    if shuffle_seed is not None:
        order = np.random.RandomState(shuffle_seed).permutation(n)      # same on every rank
    tile_num, loss_sum = 0, 0.0
    params = [p for p in model.parameters() if p.requires_grad]
    for b in range(0, n, batch_size):
        rows = order[b:b + batch_size]
        if feature_cache is not None:
            ti = trainset.train_index[rows]
            mine = rows[(ti >= feature_cache["begin"]) & (ti < feature_cache["end"])]
        else:
            mine = rows[rank::world]
        optimizer.zero_grad()
        if len(mine):
            if feature_cache is not None:
                sel = torch.from_numpy(trainset.train_index[mine] - feature_cache["begin"]).to(device)
                label = torch.from_numpy(td["label"][mine].copy()).to(device)
                out = model.fc_tile(feature_cache["feat"].index_select(0, sel))
            else:
                saved = trainset.train_data
                trainset.train_data = saved[mine]
                data, label = trainset.train_tensor(0, len(mine), device)
                trainset.train_data = saved
                out = model(data, freeze_bn=True)
            # sum over this rank's rows / global batch size == mean over the global batch
            loss = torch.nn.functional.cross_entropy(out, label, reduction="sum") / len(rows) * gamma \
                if isinstance(criterion, torch.nn.CrossEntropyLoss) else criterion(out, label) * gamma
            loss.backward()
            local = float(loss.item())
        else:
            local = 0.0
        if world > 1:
            # gradients were scaled by 1/len(rows) already: sum them (mean-of-means would be wrong)
            for p in params:
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
            flat = torch.cat([p.grad.reshape(-1) for p in params] +
                             [torch.tensor([local], device=params[0].device)])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            o = 0
            for p in params:
                p.grad.copy_(flat[o:o + p.numel()].view_as(p))
                o += p.numel()
            local = float(flat[-1].item())
        optimizer.step()
        tile_num += len(rows)
        loss_sum += local * len(rows)
    return loss_sum / max(tile_num, 1)


def mil_epoch(trainset, model, device, criterion, optimizer, tiles_per_pos, topk_neg, pos_neg_ratio,
              batch_size, gamma=1.0, seed=None, cache_features=True):
    """inference_tiles -> sample -> train_tile for one epoch; returns (mean loss, pos, neg)."""
    trainset.setmode(1)
    cache = None
    if cache_features:
        gidx, glab, cache = select_global(trainset, model, device, tiles_per_pos, topk_neg, True)
    else:
        gidx, glab = select_global(trainset, model, device, tiles_per_pos, topk_neg)
    if seed is not None:
        np.random.seed(seed)          # make_train_data's shuffle / pruning must agree on all ranks
    pos, neg = trainset.make_train_data(gidx, pos_neg_ratio, pseudo_labels=glab)
    trainset.setmode(3)
    loss = train_selected(trainset, model, device, criterion, optimizer, batch_size, gamma,
                          feature_cache=cache)
    return loss, pos, neg
