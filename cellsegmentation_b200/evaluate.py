"""Tile-mode validation (mirror of evaluate.py:8-27 + metrics/metrics.py:7-16)."""
import numpy as np
import torch

from . import ops
from .inference import _device_of, _probs_on_device


def calc_err(pred, real):
    """Error rate, FPR, FNR for tile mode (metrics/metrics.py:7-16)."""
    pred = np.asarray(pred)
    real = np.asarray(real)
    neq = np.not_equal(pred, real)
    err = float(neq.sum()) / pred.shape[0]
    fpr = float(np.logical_and(pred == 1, neq).sum()) / (real == 0).sum()
    fnr = float(np.logical_and(pred == 0, neq).sum()) / (real == 1).sum()
    return err, fpr, fnr


def evaluate_tile(valset, probs, tiles_per_pos, threshold, device=None):
    """pred = prob > threshold in lexsort order; label = 1 for the last count*tiles_per_pos sorted
    positions before each bag end (literal slice semantics of evaluate.py:20-23, including the
    spill into earlier bags when count*tiles_per_pos exceeds the bag size)."""
    device = _device_of(device if device is not None else "cuda")
    with torch.cuda.device(device):
        p = _probs_on_device(valset, probs, device)
        off_h = valset.seg_offsets()
        order = ops.lexsort_segments(p, len(valset.images), max(valset.tiles_per_bag, 1),
                                     seg_offsets=torch.from_numpy(off_h).to(device))
        pred = (p[order.long()] > threshold)
        n = p.numel()
        ends = off_h[1:]
        m = np.asarray(valset.labels, np.int64) * int(tiles_per_pos)
        has = np.diff(off_h) > 0
        starts = ends - m
        if np.any(starts[has & (m > 0)] < 0):
            raise ValueError("count*tiles_per_pos exceeds the tiles before a bag end "
                             "(the reference's slice assignment raises here too)")
        delta = torch.zeros(n + 1, dtype=torch.int32, device=device)
        sel = has & (m > 0)
        delta.index_add_(0, torch.from_numpy(starts[sel]).to(device), torch.ones(int(sel.sum()), dtype=torch.int32, device=device))
        delta.index_add_(0, torch.from_numpy(ends[sel]).to(device), -torch.ones(int(sel.sum()), dtype=torch.int32, device=device))
        real = torch.cumsum(delta[:n], 0) > 0
        neq = pred != real
        n_neq = int(neq.sum())
        fp = int((pred & neq).sum())
        fn = int((~pred & neq).sum())
        n0 = int((~real).sum())
        n1 = int(real.sum())
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(n_neq) / n, np.float64(fp) / np.int64(n0), np.float64(fn) / np.int64(n1)
