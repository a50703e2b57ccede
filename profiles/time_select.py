"""K3 at the full config size (20 000 bags x 3025): run a few selections (for ncu launch lists)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cellsegmentation_b200 import ops, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
nb, T = 20000, 3025
g = torch.Generator(device=dev)
g.manual_seed(7)
p = torch.rand(nb * T, device=dev, generator=g)
lab = torch.from_numpy(synthetic.make_labels(nb, seed=3)).to(dev)
for _ in range(3):
    idx, pl, off = ops.select_topk(p, lab, nb, T, 1, 30, capacity=nb * 330)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
idx, pl, off = ops.select_topk(p, lab, nb, T, 1, 30, capacity=nb * 330)
b.record()
torch.cuda.synchronize()
print("select 20k bags: %.3f ms, kept %d" % (a.elapsed_time(b), idx.numel()))
