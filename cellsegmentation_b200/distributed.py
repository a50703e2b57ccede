"""Multi-GPU plumbing: one process per GPU, bags sharded contiguously, NCCL only where the
path has a real exchange.

  * inference / selection / masks are per-bag independent -> no data-path collective
  * MIL training: ONE fixed-capacity all-gather of the selected global tile indices +
    pseudo-labels so every rank rebuilds the identical train_data (make_train_data's
    shuffle/prune depends on GLOBAL pos/neg counts, dataset/dataset.py:171-199), and one
    all-reduce (sum) of the fc_tile gradients + loss each step (1 027 floats, one bucket)

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


def kept_upper_bound(labels, tiles_per_bag, tiles_per_pos, topk_neg):
    """Per-bag upper bound min(k_b, T) of what sample() keeps (inference.py:37-40): the literal
    predicate keeps exactly that many sorted positions of a bag, fewer only when the (i + k) % N
    wrap-around lands back inside the bag.  Known on every rank from the count labels alone, so
    the selection exchange needs no size negotiation."""
    lab = np.asarray(labels, dtype=np.int64)
    k = np.where(lab == 0, int(topk_neg), lab * int(tiles_per_pos))
    return np.minimum(k, int(tiles_per_bag))


def selection_capacity(labels_of_tile_bags, tiles_per_bag, tiles_per_pos, topk_neg, world):
    """Slots per rank of the packed all-gather buffer: the largest per-shard sum of min(k_b, T)."""
    ub = kept_upper_bound(labels_of_tile_bags, tiles_per_bag, tiles_per_pos, topk_neg)
    cap = 0
    for r in range(world):
        lo, hi = shard_range(len(ub), r, world)
        cap = max(cap, int(ub[lo:hi].sum()))
    return cap


def allgather_selection(idx_local, labels_local, count_local, tile_offset, capacity, group=None,
                        timing=None):
    """Local selection -> global, identical on all ranks, ascending by (bag, prob, index)
    because shards are contiguous blocks in rank order.

    idx_local i32 [>= capacity] (indices relative to this rank's first tile), labels_local u8,
    count_local: 0-d / 1-element integer tensor ON THE DEVICE (the kept count; never read on the
    host before the exchange).  One all_gather_into_tensor of `1 + capacity` int64 words per
    rank: word 0 = count, word 1+j = (global index << 1) | pseudo-label.  Returns host arrays
    (idx int64 [M], labels u8 [M]) after ONE device->host copy of the gathered buffer."""
    rank, world = _world(group)
    dev = idx_local.device
    buf = torch.empty(1 + capacity, dtype=torch.int64, device=dev)
    buf[0:1] = count_local.reshape(-1)[:1].to(torch.int64)
    if capacity:
        buf[1:] = ((idx_local[:capacity].to(torch.int64) + int(tile_offset)) << 1) | \
            labels_local[:capacity].to(torch.int64)
    if world == 1:
        out = buf
    else:
        out = torch.empty(world * (1 + capacity), dtype=torch.int64, device=dev)
        ev = None
        if timing is not None and dev.type == "cuda":
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        dist.all_gather_into_tensor(out, buf, group=group)
        if ev is not None:
            ev[1].record()
            timing.setdefault("allgather_events", []).append(ev)
            timing["allgather_bytes"] = int(out.numel() * 8)
    host = out.cpu().numpy().reshape(world, 1 + capacity)
    parts = []
    for r in range(world):
        m = int(host[r, 0])
        if m > capacity:
            raise RuntimeError("rank %d kept %d instances, above the closed-form bound %d" % (r, m, capacity))
        parts.append(host[r, 1:1 + m])
    packed = np.concatenate(parts) if parts else np.zeros(0, np.int64)
    return packed >> 1, (packed & 1).astype(np.uint8)


def allreduce_flat(tensors, group=None, timing=None):
    """One-bucket all-reduce (sum) of a list of tensors; returns the reduced flat buffer (the
    caller scatters it back).  fc_tile's gradients and the batch loss travel together: one NCCL
    call per training step."""
    flat = torch.cat([t.reshape(-1) for t in tensors])
    rank, world = _world(group)
    if world > 1:
        ev = None
        if timing is not None and flat.is_cuda:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if ev is not None:
            ev[1].record()
            timing.setdefault("allreduce_events", []).append(ev)
            timing["allreduce_bytes"] = int(flat.numel() * flat.element_size())
    return flat


def broadcast_seed(seed, device, group=None):
    """The seed every rank must use for make_train_data's shuffle / pruning.  With seed=None rank 0
    draws one from its own np.random state (so a seeded single-process run and rank 0 of a
    distributed run consume the global state alike) and broadcasts it."""
    rank, world = _world(group)
    if world == 1:
        return seed
    s = torch.zeros(1, dtype=torch.int64, device=device)
    if rank == 0:
        s[0] = int(seed) if seed is not None else int(np.random.randint(0, 2 ** 31 - 1))
    dist.broadcast(s, src=0, group=group)
    return int(s.item())


def shard_dataset(ds, rank, world):
    """A view of `ds` (this package's LystoDataset / LystoTestset) holding only this rank's
    tile-owning bags; returns (shard, first_global_tile_index).  The shard (and with it the
    device copy of its bags) is cached on the parent, so an epoch loop uploads a shard once."""
    import copy
    bags = ds._tile_bags
    key = (rank, world, len(ds.images), len(bags), ds.tile_size, ds.interval)
    cached = getattr(ds, "_shard_cache", None)
    if cached is not None and cached[0] == key:
        return cached[1], cached[2]
    lo, hi = shard_range(len(bags), rank, world)
    sh = copy.copy(ds)
    sh._shard_cache = None
    keep = bags[lo:hi]
    b0 = keep[0] if keep else 0
    # keep bag numbering dense: shard bag j <-> global bag b0 + j (images before b0 are dropped,
    # except that a LystoDataset shard starting at bag 0 keeps the tile-less first bag)
    first_img = 0 if (lo == 0) else b0
    last_img = (keep[-1] + 1) if keep else first_img
    if world == 1 and first_img == 0 and last_img == len(ds.images):
        sh = ds                                   # the whole set: share the resident tensor
    else:
        sh._dev = None
        if getattr(ds.images, "is_cuda", False):          # bags already in HBM: a view, no copy
            sh.images = ds.images[first_img:last_img]
        else:
            sh.images = list(ds.images[first_img:last_img])
        sh.organs = list(ds.organs[first_img:last_img])
        if hasattr(ds, "labels"):
            sh.labels = list(ds.labels[first_img:last_img])
        sh._tile_bags = [b - first_img for b in keep]
        sh._first_img = first_img
    ds._shard_cache = (key, sh, lo * ds.tiles_per_bag)
    return sh, lo * ds.tiles_per_bag
