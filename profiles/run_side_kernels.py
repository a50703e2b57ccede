"""One call of every non-CNN kernel at a realistic size (for an ncu metrics pass)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cellsegmentation_b200 import ops, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
H, T, nb = 299, 3025, 4000
g = torch.Generator(device=dev); g.manual_seed(7)
bags = synthetic.make_bags_device(nb, dev, seed=1)
p = torch.rand(nb * T, device=dev, generator=g)
lab = torch.from_numpy(synthetic.make_labels(nb, seed=3)).to(dev)
for _ in range(2):
    idx, pl, off = ops.select_topk(p, lab, nb, T, 1, 30, capacity=nb * 330)
    ridx, rp, roff = ops.rank_threshold(p, nb, T, 0.95, capacity=nb * T)
    mask = ops.paint_mask(idx, nb, H, H, 32, 5)
    heat = ops.paint_heatmap(ridx, rp, nb, H, H, 32, 5)
    ref = ops.hsv_refine(bags, mask, 170)
    cc = ops.remove_small_regions(ref.clone(), 400, 120)
    lut = torch.arange(768, dtype=torch.int32, device=dev).remainder(256).to(torch.uint8).reshape(256, 3)
    blend = ops.heatmap_blend(heat, bags, lut)
    tiles = ops.unfold_normalize(bags[:8], 32, 5)
torch.cuda.synchronize()
print("ok", idx.numel(), ridx.numel(), int(ref.sum()), int(cc.sum()))
