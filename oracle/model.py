"""Tile-classifier forward (oracle; test infrastructure only).

Restates MILResNet tile mode on CPU fp32 with the same torch ops the reference calls:
  resnet_forward   model/resnet.py:234-248
  BasicBlock       model/resnet.py:28-43
  Bottleneck       model/resnet.py:60-78 (resnet50), model/resnext.py:93-113 (grouped 3x3)
  tile head        model/resnet.py:264-269  (avgpool_tile + maxpool_tile -> Flatten -> Linear)
  prob             inference.py:24-27       (softmax(dim=1)[:, 1])
State-dict keys are the reference's (torchvision-style + fc_tile.1.*), SURVEY Appendix B.
"""
import numpy as np
import torch
import torch.nn.functional as F

LAYERS = {"resnet18": [2, 2, 2, 2], "resnet34": [3, 4, 6, 3], "resnet50": [3, 4, 6, 3],
          "resnext50_32x4d": [3, 4, 6, 3], "resnext101_32x8d": [3, 4, 23, 3]}
# (groups, width_per_group) of the Bottleneck nets; model/resnet.py:355-361, model/resnext.py:418-428
BOTTLENECK = {"resnet50": (1, 64), "resnext50_32x4d": (32, 4), "resnext101_32x8d": (32, 8)}
PLANES = [64, 128, 256, 512]


def feature_dim(arch):
    return 2048 if arch in BOTTLENECK else 512
BN_EPS = 1e-5


def _bn(sd, x, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                        sd[p + ".bias"], training=False, eps=BN_EPS)


def forward_features(sd, x, arch="resnet34", return_intermediate=False):
    """x f32 [n,3,S,S] -> x4 (and x3,x2,x1); model/resnet.py:234-248 under model.eval()."""
    x = F.conv2d(x, sd["conv1.weight"], stride=2, padding=3)
    x = F.relu(_bn(sd, x, "bn1"))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    outs = []
    for L, nb in enumerate(LAYERS[arch], start=1):
        for b in range(nb):
            p = "layer%d.%d" % (L, b)
            stride = 2 if (b == 0 and L > 1) else 1
            residual = x
            if arch in BOTTLENECK:
                out = F.relu(_bn(sd, F.conv2d(x, sd[p + ".conv1.weight"]), p + ".bn1"))
                out = F.conv2d(out, sd[p + ".conv2.weight"], stride=stride, padding=1,
                               groups=BOTTLENECK[arch][0])
                out = F.relu(_bn(sd, out, p + ".bn2"))
                out = _bn(sd, F.conv2d(out, sd[p + ".conv3.weight"]), p + ".bn3")
            else:
                out = F.conv2d(x, sd[p + ".conv1.weight"], stride=stride, padding=1)
                out = F.relu(_bn(sd, out, p + ".bn1"))
                out = F.conv2d(out, sd[p + ".conv2.weight"], stride=1, padding=1)
                out = _bn(sd, out, p + ".bn2")
            if (p + ".downsample.0.weight") in sd:
                residual = _bn(sd, F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride),
                               p + ".downsample.1")
            x = F.relu(out + residual)
        outs.append(x)
    return tuple(reversed(outs)) if return_intermediate else x


def pooled(x4):
    """avgpool_tile(x4) + maxpool_tile(x4), flattened (model/resnet.py:266-267)."""
    return (F.adaptive_avg_pool2d(x4, 1) + F.adaptive_max_pool2d(x4, 1)).flatten(1)


def forward_logits(sd, x, arch="resnet34"):
    with torch.no_grad():
        f = pooled(forward_features(sd, x, arch))
        return F.linear(f, sd["fc_tile.1.weight"], sd["fc_tile.1.bias"])


def forward_probs(sd, x, arch="resnet34", batch=1024):
    """inference_tiles body (inference.py:19-28) on a materialised tile tensor."""
    out = []
    with torch.no_grad():
        for i in range(0, x.shape[0], batch):
            out.append(F.softmax(forward_logits(sd, x[i:i + batch], arch), dim=1)[:, 1].clone())
    return torch.cat(out).numpy()


# ---------------------------------------------------------------------------
# Deterministic synthetic weights (numpy PCG64 streams are version-stable)
# ---------------------------------------------------------------------------
def make_state_dict(arch="resnet34", seed=0, random_bn=True):
    """kaiming-normal convs (model/resnet.py:171-178, fan_in / leaky_relu(0) gain sqrt(2)),
    BN gamma/beta/running stats randomised mildly so folding is exercised."""
    rng = np.random.default_rng(seed)
    sd = {}

    def conv(name, cout, cin, k):
        std = np.sqrt(2.0 / (cin * k * k))
        sd[name] = torch.from_numpy((rng.standard_normal((cout, cin, k, k)) * std).astype(np.float32))

    def bn(name, c):
        if random_bn:
            sd[name + ".weight"] = torch.from_numpy(rng.uniform(0.8, 1.2, c).astype(np.float32))
            sd[name + ".bias"] = torch.from_numpy((rng.standard_normal(c) * 0.1).astype(np.float32))
            sd[name + ".running_mean"] = torch.from_numpy((rng.standard_normal(c) * 0.1).astype(np.float32))
            sd[name + ".running_var"] = torch.from_numpy(rng.uniform(0.8, 1.2, c).astype(np.float32))
        else:
            sd[name + ".weight"] = torch.ones(c)
            sd[name + ".bias"] = torch.zeros(c)
            sd[name + ".running_mean"] = torch.zeros(c)
            sd[name + ".running_var"] = torch.ones(c)
        sd[name + ".num_batches_tracked"] = torch.tensor(0)

    conv("conv1.weight", 64, 3, 7)
    bn("bn1", 64)
    inplanes = 64
    for L, nb in enumerate(LAYERS[arch], start=1):
        planes = PLANES[L - 1]
        for b in range(nb):
            p = "layer%d.%d" % (L, b)
            stride = 2 if (b == 0 and L > 1) else 1
            if arch in BOTTLENECK:
                groups, wpg = BOTTLENECK[arch]
                width = int(planes * (wpg / 64.)) * groups
                conv(p + ".conv1.weight", width, inplanes, 1)
                bn(p + ".bn1", width)
                conv(p + ".conv2.weight", width, width // groups, 3)
                bn(p + ".bn2", width)
                conv(p + ".conv3.weight", planes * 4, width, 1)
                bn(p + ".bn3", planes * 4)
                if stride != 1 or inplanes != planes * 4:
                    conv(p + ".downsample.0.weight", planes * 4, inplanes, 1)
                    bn(p + ".downsample.1", planes * 4)
                inplanes = planes * 4
                continue
            conv(p + ".conv1.weight", planes, inplanes, 3)
            bn(p + ".bn1", planes)
            conv(p + ".conv2.weight", planes, planes, 3)
            bn(p + ".bn2", planes)
            if stride != 1 or inplanes != planes:
                conv(p + ".downsample.0.weight", planes, inplanes, 1)
                bn(p + ".downsample.1", planes)
            inplanes = planes
    fd = feature_dim(arch)
    bound = 1.0 / np.sqrt(fd)
    sd["fc_tile.1.weight"] = torch.from_numpy(rng.uniform(-bound, bound, (2, fd)).astype(np.float32))
    sd["fc_tile.1.bias"] = torch.from_numpy(rng.uniform(-bound, bound, 2).astype(np.float32))
    return sd


def make_image_seg_state(arch="resnet34", seed=0):
    """Deterministic weights for the Stage-1 heads (fc_image_cls / fc_image_reg, model/resnet.py:129-153)
    and the Stage-3 decoder (upconv1..8, seg_out_conv, :155-165): same keys and shapes as the
    reference modules, loaded with load_state_dict(strict=False) into the reference (golden
    generation) and into the mirror (tests)."""
    rng = np.random.default_rng([seed, 4242])
    e = 4 if arch in BOTTLENECK else 1
    feat = 512 * e
    sd = {}

    def t(a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))

    def bn(name, c):
        sd[name + ".weight"] = t(rng.uniform(0.8, 1.2, c))
        sd[name + ".bias"] = t(rng.standard_normal(c) * 0.1)
        sd[name + ".running_mean"] = t(rng.standard_normal(c) * 0.1)
        sd[name + ".running_var"] = t(rng.uniform(0.8, 1.2, c))

    for head, n_out in (("fc_image_cls", 7), ("fc_image_reg", 1)):
        bn(head + ".1", feat)
        sd[head + ".4.weight"] = t(rng.standard_normal((64, feat)) / np.sqrt(feat))
        sd[head + ".4.bias"] = t(rng.standard_normal(64) * 0.1)
        bn(head + ".5", 64)
        sd[head + ".7.weight"] = t(rng.standard_normal((n_out, 64)) / 8.0)
        sd[head + ".7.bias"] = t(rng.standard_normal(n_out) * 0.1 + (1.0 if n_out == 1 else 0.0))
    chans = [(512 * e, 256 * e), (512 * e, 256 * e), (256 * e, 128 * e), (256 * e, 128 * e), (128 * e, 64 * e),
             (128 * e, 64 * e), (64 * e, 64 if e == 1 else 32 * e), (64 if e == 1 else 32 * e, 64)]
    for i, (ci, co) in enumerate(chans, start=1):
        sd["upconv%d.0.weight" % i] = t(rng.standard_normal((co, ci, 3, 3)) * np.sqrt(2.0 / (ci * 9)))
        sd["upconv%d.0.bias" % i] = t(rng.standard_normal(co) * 0.05)
        bn("upconv%d.1" % i, co)
    sd["seg_out_conv.weight"] = t(rng.standard_normal((2, 64, 1, 1)) * np.sqrt(2.0 / 64))
    sd["seg_out_conv.bias"] = t(rng.standard_normal(2) * 0.05)
    return sd


def calibrate_head(sd, calib_x, arch="resnet34", sigma=2.0):
    """PC-1-aligned, zero-centred fc_tile so probabilities spread over (0,1) instead of
    saturating (SURVEY 3.5-12 / 7 'Precision gates' recipe)."""
    with torch.no_grad():
        f = pooled(forward_features(sd, calib_x, arch)).double().numpy()
    mu = f.mean(0)
    _, _, vt = np.linalg.svd(f - mu, full_matrices=False)
    w = vt[0]
    proj = (f - mu) @ w
    s = sigma / proj.std()
    W = np.stack([-w * s / 2, w * s / 2]).astype(np.float32)
    b = np.array([(mu @ w) * s / 2, -(mu @ w) * s / 2], np.float32)
    sd = dict(sd)
    sd["fc_tile.1.weight"] = torch.from_numpy(W)
    sd["fc_tile.1.bias"] = torch.from_numpy(b)
    return sd


def fold_bn(sd, arch="resnet34"):
    """Eval-mode BN folded into the preceding conv, fp32:
    s = gamma / sqrt(var + eps); W' = W * s; b' = beta - mean * s.
    Returns [(W', b')] in network order: stem; per block conv1, conv2, [conv3], [downsample]."""
    def fold(wname, bnname):
        w = sd[wname].float()
        s = sd[bnname + ".weight"].float() / torch.sqrt(sd[bnname + ".running_var"].float() + BN_EPS)
        return (w * s[:, None, None, None]).contiguous(), \
               (sd[bnname + ".bias"].float() - sd[bnname + ".running_mean"].float() * s).contiguous()

    convs = [fold("conv1.weight", "bn1")]
    for L, nb in enumerate(LAYERS[arch], start=1):
        for b in range(nb):
            p = "layer%d.%d" % (L, b)
            convs.append(fold(p + ".conv1.weight", p + ".bn1"))
            convs.append(fold(p + ".conv2.weight", p + ".bn2"))
            if (p + ".conv3.weight") in sd:
                convs.append(fold(p + ".conv3.weight", p + ".bn3"))
            if (p + ".downsample.0.weight") in sd:
                convs.append(fold(p + ".downsample.0.weight", p + ".downsample.1"))
    return convs
