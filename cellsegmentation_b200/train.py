"""One-epoch MIL tile training (mirror of train/train.py:12-48 of the reference).

The encoder is frozen and in eval mode during tile training (model/resnet.py:254-258,
315-319), so every step is: CUDA encoder features (no grad) -> fc_tile -> CE loss -> backward
-> optimizer step, exactly the parameters the reference updates.  With this package's
LystoDataset in mode 3 the tiles are gathered on the device instead of through DataLoader
workers; batch composition follows the loader's batch_size / shuffle setting.
"""
import torch
from torch.optim.lr_scheduler import CyclicLR, OneCycleLR
from torch.utils.data import RandomSampler

from .dataset import LystoDataset


def _batches(loader, n):
    bs = loader.batch_size
    if isinstance(getattr(loader, "sampler", None), RandomSampler):
        perm = torch.randperm(n).numpy()
    else:
        perm = None
    for b in range(0, n, bs):
        yield (b, min(bs, n - b), None) if perm is None else (b, min(bs, n - b), perm[b:b + bs])


def train_tile(loader, epoch, total_epochs, model, device, criterion, optimizer, scheduler, gamma,
               grad_sync=None):
    """Tile training for one epoch.  grad_sync (optional) is called after backward() and before
    optimizer.step() — the hook the multi-GPU driver uses for the NCCL gradient all-reduce."""
    model.train()
    tile_num, train_loss = 0, 0.
    ds = loader.dataset
    fast = isinstance(ds, LystoDataset) and ds.mode == 3

    def run(data, label):
        nonlocal tile_num, train_loss
        optimizer.zero_grad()
        output = model(data, freeze_bn=True)
        loss = criterion(output, label) * gamma
        loss.backward()
        if grad_sync is not None:
            grad_sync(model)
        optimizer.step()
        if isinstance(scheduler, (CyclicLR, OneCycleLR)):
            scheduler.step()
        tile_num += data.size(0)
        train_loss += loss.item() * data.size(0)

    if fast:
        saved = ds.train_data
        for b, cnt, rows in _batches(loader, len(saved)):
            if rows is not None:
                ds.train_data = saved[rows]
                data, label = ds.train_tensor(0, cnt, device)
                ds.train_data = saved
            else:
                data, label = ds.train_tensor(b, cnt, device)
            run(data, label)
    else:
        for data, label in loader:
            run(data.to(device), label.to(device))

    if not (scheduler is None or isinstance(scheduler, (CyclicLR, OneCycleLR))):
        scheduler.step()
    return train_loss / max(tile_num, 1)
