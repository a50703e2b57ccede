"""Seeded synthetic LYSTO-shaped inputs (oracle side; test infrastructure only)."""
import numpy as np


def make_bags(n_bags, H=299, W=299, seed=0):
    """u8 [n,H,W,3]: smooth blobs + noise whose brightness straddles V = 170."""
    out = np.empty((n_bags, H, W, 3), np.uint8)
    for b in range(n_bags):
        rng = np.random.default_rng([seed, b])
        coarse = rng.uniform(60, 250, (H // 23 + 2, W // 23 + 2, 3))
        up = np.kron(coarse, np.ones((23, 23, 1)))[:H, :W]
        up = up + rng.normal(0, 12, (H, W, 3))
        out[b] = np.clip(up, 0, 255).astype(np.uint8)
    return out


def make_labels(n_bags, seed=0):
    """LYSTO-like counts: ~30 % zeros, rest geometric (mean ~8), capped at 300."""
    rng = np.random.default_rng([seed, 7919])
    lab = rng.geometric(1.0 / 8.0, n_bags)
    lab[rng.uniform(size=n_bags) < 0.3] = 0
    return np.minimum(lab, 300).astype(np.int32)


def make_probs(n, seed=0, ties=False):
    rng = np.random.default_rng([seed, 104729])
    p = rng.uniform(0, 1, n).astype(np.float32)
    if ties:
        p = np.round(p * 8) / np.float32(8)
    return p.astype(np.float32)
