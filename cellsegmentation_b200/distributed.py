"""Multi-GPU plumbing: one process per GPU, bags sharded contiguously, NCCL only where the
path has a real exchange.

  * inference / selection / masks are per-bag independent -> no data-path collective
  * MIL training: one all-gather(v) of the selected global tile indices + pseudo-labels so every
    rank rebuilds the identical train_data (make_train_data's shuffle/prune depends on GLOBAL
    pos/neg counts, dataset/dataset.py:171-199), and an all-reduce (mean) of the fc_tile
    gradients each step (1 026 floats, one bucket)

Works with the "nccl" backend on CUDA tensors and with "gloo" on CPU tensors (tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_bags, rank, world):
    """Contiguous block of ceil(n/world) bags per rank: keeps tileIDX monotone, segments whole."""
    per = -(-n_bags // world)
    lo = min(rank * per, n_bags)
    return lo, min(lo + per, n_bags)


def _world(group=None):
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def allgather_varlen(t, group=None):
    """Concatenation over ranks (rank order) of 1-D tensors of different lengths."""
    rank, world = _world(group)
    if world == 1:
        return t
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    mx = max(counts) if counts else 0
    pad = torch.zeros(mx, dtype=t.dtype, device=t.device)
    pad[:t.numel()] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)])


def allgather_selection(idx_local, labels_local, tile_offset, group=None):
    """Local selection (indices relative to this rank's first tile) -> global, identical on all
    ranks, ascending by (bag, prob, index) because shards are contiguous blocks in rank order."""
    idx = allgather_varlen(idx_local.to(torch.int64) + int(tile_offset), group)
    lab = allgather_varlen(labels_local, group)
    return idx, lab


def allreduce_mean_grads(params, group=None):
    """One bucket all-reduce (sum) / world of the gradients of `params` (fc_tile in tile mode)."""
    rank, world = _world(group)
    grads = [p.grad for p in params if p.grad is not None]
    if world == 1 or not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= world
    o = 0
    for g in grads:
        g.copy_(flat[o:o + g.numel()].view_as(g))
        o += g.numel()


def shard_dataset(ds, rank, world):
    """A view of `ds` (this package's LystoDataset / LystoTestset) holding only this rank's
    tile-owning bags; returns (shard, first_global_tile_index)."""
    import copy
    bags = ds._tile_bags
    lo, hi = shard_range(len(bags), rank, world)
    sh = copy.copy(ds)
    sh._dev = None
    keep = bags[lo:hi]
    b0 = keep[0] if keep else 0
    # keep bag numbering dense: shard bag j <-> global bag b0 + j (images before b0 are dropped,
    # except that a LystoDataset shard starting at bag 0 keeps the tile-less first bag)
    first_img = 0 if (lo == 0) else b0
    last_img = (keep[-1] + 1) if keep else first_img
    sh.images = list(ds.images[first_img:last_img])
    sh.organs = list(ds.organs[first_img:last_img])
    if hasattr(ds, "labels"):
        sh.labels = list(ds.labels[first_img:last_img])
    sh._tile_bags = [b - first_img for b in keep]
    return sh, lo * ds.tiles_per_bag
