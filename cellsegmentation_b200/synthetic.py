"""Synthetic LYSTO-shaped inputs generated on the device (there is no network for the dataset).

Shapes follow dataset/dataset.py:26,59-65: 299x299x3 u8 bags with one count label each.
"""
import numpy as np
import torch

IMAGE_SIZE = 299


def make_bags_device(n_bags, device, seed=0, chunk=512, kind="blobs"):
    """u8 [n_bags,299,299,3] resident in HBM, generated on the device.
    kind="blobs" (default): smooth 23-pixel blobs of brightness 60..250 plus N(0, 12) pixel noise,
    the recipe of the parity tests' bags (oracle/synth.py; SURVEY 8d: "smooth blobs + noise so
    that V straddles 170") -- tiles of one bag differ from each other as LYSTO tiles do, so a
    calibrated head separates them and the bf16-vs-fp32 check of bench.py means something.
    kind="noise": iid uniform bytes (every tile statistically the same)."""
    g = torch.Generator(device=device)
    g.manual_seed(1234567 + seed)
    out = torch.empty((n_bags, IMAGE_SIZE, IMAGE_SIZE, 3), dtype=torch.uint8, device=device)
    cells = IMAGE_SIZE // 23 + 2
    for b in range(0, n_bags, chunk):
        e = min(n_bags, b + chunk)
        if kind == "noise":
            out[b:e] = torch.randint(0, 256, (e - b, IMAGE_SIZE, IMAGE_SIZE, 3), dtype=torch.uint8,
                                     device=device, generator=g)
            continue
        coarse = torch.rand((e - b, cells, cells, 3), device=device, generator=g) * 190.0 + 60.0
        up = coarse.repeat_interleave(23, 1).repeat_interleave(23, 2)[:, :IMAGE_SIZE, :IMAGE_SIZE]
        up = up + torch.randn(up.shape, device=device, generator=g) * 12.0
        out[b:e] = up.clamp_(0, 255).to(torch.uint8)
    return out


def make_labels(n_bags, seed=0):
    """LYSTO-like counts: ~30 % zeros, the rest geometric with mean ~8, capped at 300 (int32)."""
    rng = np.random.default_rng([seed, 7919])
    lab = rng.geometric(1.0 / 8.0, n_bags)
    lab[rng.uniform(size=n_bags) < 0.3] = 0
    return np.minimum(lab, 300).astype(np.int32)


def make_resnet_weights(arch="resnet34", seed=0):
    """Random-init BN-folded conv list + fc_tile for benchmarking: kaiming-normal convs
    (model/resnet.py:171-175), BN at init (gamma 1, beta 0, stats 0/1 -> identity fold up to eps)."""
    layers = {"resnet18": [2, 2, 2, 2], "resnet34": [3, 4, 6, 3], "resnet50": [3, 4, 6, 3],
              "resnext50_32x4d": [3, 4, 6, 3], "resnext101_32x8d": [3, 4, 23, 3]}[arch]
    bottleneck = {"resnet50": (1, 64), "resnext50_32x4d": (32, 4), "resnext101_32x8d": (32, 8)}.get(arch)   # (groups, width/group)
    rng = np.random.default_rng(seed)
    s = np.float32(1.0 / np.sqrt(1.0 + 1e-5))

    def conv(cout, cin, k):
        w = rng.standard_normal((cout, cin, k, k)).astype(np.float32) * np.float32(np.sqrt(2.0 / (cin * k * k)))
        return torch.from_numpy(w * s), torch.zeros(cout)

    convs = [conv(64, 3, 7)]
    inpl = 64
    for L, nb in enumerate(layers):
        pl = 64 << L
        for b in range(nb):
            stride = 2 if (b == 0 and L > 0) else 1
            if bottleneck:
                groups, wpg = bottleneck
                width = pl * wpg // 64 * groups
                convs.append(conv(width, inpl, 1))
                convs.append(conv(width, width // groups, 3))
                convs.append(conv(pl * 4, width, 1))
                if stride != 1 or inpl != pl * 4:
                    convs.append(conv(pl * 4, inpl, 1))
                inpl = pl * 4
                continue
            convs.append(conv(pl, inpl, 3))
            convs.append(conv(pl, pl, 3))
            if stride != 1 or inpl != pl:
                convs.append(conv(pl, inpl, 1))
            inpl = pl
    fd = 2048 if bottleneck else 512
    bound = 1.0 / np.sqrt(fd)
    fc_w = torch.from_numpy(rng.uniform(-bound, bound, (2, fd)).astype(np.float32)) * 0.05
    fc_b = torch.zeros(2)
    return convs, fc_w, fc_b


def calibrate_head(clf, bags, tile, interval, n_calib=4096, sigma=2.0, max_batch=512):
    """fc_tile for a random-init encoder whose probabilities spread over (0, 1) instead of
    saturating (SURVEY 3.5-12 / 7 "Precision gates" recipe, the one the parity tests use): the
    head is aligned with the first principal component of the pooled features of `n_calib`
    instances and scaled so the logit difference has standard deviation `sigma`.  Features come
    from the classifier's own fp32 CUDA path.  Returns (fc_w [2,F], fc_b [2]) and installs them."""
    n_calib = min(n_calib, bags.shape[0] * 3025 if interval == 5 else n_calib)
    _, feat = clf.forward_tiles(bags, tile, interval, inst_begin=0, inst_count=n_calib, precision="fp32",
                                max_batch=max_batch, want_features=True)
    f = feat.double().cpu().numpy()
    mu = f.mean(0)
    _, _, vt = np.linalg.svd(f - mu, full_matrices=False)
    w = vt[0]
    s = sigma / ((f - mu) @ w).std()
    fc_w = torch.from_numpy(np.stack([-w * s / 2, w * s / 2]).astype(np.float32))
    fc_b = torch.from_numpy(np.array([(mu @ w) * s / 2, -(mu @ w) * s / 2], np.float32))
    clf.set_fc(fc_w, fc_b)
    return fc_w, fc_b
