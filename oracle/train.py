"""MIL tile-training epoch (oracle; test infrastructure only).

Restates train_tile (train/train.py:12-48) with the oracle forward: the encoder runs in eval
mode without gradients (model/resnet.py:254-258 with freeze_bn=True, requires_grad False from
setmode("tile"), :315-319), only fc_tile.1 is trained with CrossEntropyLoss * gamma."""
import torch
import torch.nn.functional as F

from . import model as omodel


def train_tile_epoch(sd, tiles, labels, batch_size, lr, arch="resnet34", gamma=1.0, weight_decay=0.0):
    """tiles f32 [M,3,S,S] and int64 labels [M] in loader order (shuffle=False).
    Plain SGD like torch.optim.SGD(lr, weight_decay).  Returns (mean loss, fc weight, fc bias)."""
    w = sd["fc_tile.1.weight"].clone().requires_grad_(True)
    b = sd["fc_tile.1.bias"].clone().requires_grad_(True)
    opt = torch.optim.SGD([w, b], lr=lr, weight_decay=weight_decay)
    tile_num, train_loss = 0, 0.0
    for i in range(0, tiles.shape[0], batch_size):
        x, y = tiles[i:i + batch_size], labels[i:i + batch_size]
        opt.zero_grad()
        with torch.no_grad():
            f = omodel.pooled(omodel.forward_features(sd, x, arch))
        loss = F.cross_entropy(F.linear(f, w, b), y) * gamma
        loss.backward()
        opt.step()
        tile_num += x.shape[0]
        train_loss += loss.item() * x.shape[0]
    return train_loss / tile_num, w.detach(), b.detach()
