"""Adaptive top-k selection, ranking and tile evaluation (oracle; test infrastructure only)."""
import numpy as np


def sample_indices_loop(tileIDX, labels, probs, tiles_per_pos, topk_neg):
    """Literal sample() up to make_train_data (inference.py:31-42): returns order[index]."""
    groups = np.array(tileIDX)
    order = np.lexsort((probs, groups))
    n = len(groups)
    index = np.empty(n, 'bool')
    for i in range(n):
        topk = topk_neg if labels[groups[i]] == 0 else labels[groups[i]] * tiles_per_pos
        index[i] = groups[i] != groups[(i + topk) % n]
    return order[index]


def sample_indices(tileIDX, labels, probs, tiles_per_pos, topk_neg):
    """Vectorised identity of the loop above."""
    groups = np.asarray(tileIDX, np.int64)
    n = len(groups)
    if n == 0:
        return np.zeros(0, np.int64)
    order = np.lexsort((probs, groups))
    lab = np.asarray(labels, np.int64)[groups]
    k = np.where(lab == 0, np.int64(topk_neg), lab * np.int64(tiles_per_pos))
    index = groups != groups[(np.arange(n, dtype=np.int64) + k) % n]
    return order[index]


def pseudo_labels(tileIDX, labels, idxs):
    """make_train_data label rule (dataset/dataset.py:168-169)."""
    g = np.asarray(tileIDX)[idxs]
    return (np.asarray(labels)[g] != 0).astype(np.uint8)


def make_train_data(tileIDX, tiles_grid, labels, idxs, pos_neg_ratio):
    """make_train_data (dataset/dataset.py:166-201) with an explicit object array (the
    reference's ragged np.array(...) raises on numpy >= 1.24).  Uses the global
    np.random state exactly like the reference.  Returns (train_data object[M',3], pos, neg)."""
    td = np.empty((len(idxs), 3), dtype=object)
    for r, i in enumerate(idxs):
        td[r, 0] = tileIDX[i]
        td[r, 1] = tiles_grid[i]
        td[r, 2] = 0 if labels[tileIDX[i]] == 0 else 1
    pos = 0
    for _, _, label in td:
        pos += label
    neg = len(td) - pos
    np.random.shuffle(td)
    if pos_neg_ratio is not None:
        if pos > int(neg * pos_neg_ratio):
            flag = 1
            n = pos - int(neg * pos_neg_ratio)
            pos = int(neg * pos_neg_ratio)
        elif neg > int(pos / pos_neg_ratio):
            flag = 0
            n = neg - int(pos / pos_neg_ratio)
            neg = int(pos / pos_neg_ratio)
        else:
            return td, pos, neg
        excess = []
        for i, (_, _, label) in enumerate(td):
            if label == flag:
                excess.append(i)
            if len(excess) == n:
                break
        td = np.delete(td, excess, 0)
    return td, pos, neg


def rank(tileIDX, tiles_grid, probs, threshold):
    """rank() (test_tile.py:63-79): tiles, probs, groups of prob > threshold in lexsort order."""
    groups = np.array(tileIDX)
    tiles = np.array(tiles_grid)
    order = np.lexsort((probs, groups))
    groups = groups[order]
    probs = probs[order]
    tiles = tiles[order]
    index = np.array([prob > threshold for prob in probs], dtype=bool)
    return tiles[index], probs[index], groups[index], order[index]


def evaluate_tile(tileIDX, labels, probs, tiles_per_pos, threshold):
    """evaluate_tile + calc_err (evaluate.py:8-27, metrics/metrics.py:7-16)."""
    val_groups = np.array(tileIDX)
    order = np.lexsort((probs, val_groups))
    val_groups = val_groups[order]
    val_probs = probs[order]
    val_index = np.array([prob > threshold for prob in val_probs])
    lab = np.zeros(len(val_probs))
    for i in range(1, len(val_probs) + 1):
        if i == len(val_probs) or val_groups[i] != val_groups[i - 1]:
            m = labels[val_groups[i - 1]] * tiles_per_pos
            lab[i - m: i] = [1] * m
    pred = np.asarray(val_index)
    real = np.asarray(lab)
    neq = np.not_equal(pred, real)
    err = float(neq.sum()) / pred.shape[0]
    fpr = float(np.logical_and(pred == 1, neq).sum()) / (real == 0).sum()
    fnr = float(np.logical_and(pred == 0, neq).sum()) / (real == 1).sum()
    return err, fpr, fnr
