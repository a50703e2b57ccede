"""One encoder at one (interval, max_batch): two forward passes over a few bags (for ncu launch lists).
  python profiles/run_arch.py resnext50_32x4d 3 37888"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cellsegmentation_b200 import ops, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
arch = sys.argv[1] if len(sys.argv) > 1 else "resnext50_32x4d"
interval = int(sys.argv[2]) if len(sys.argv) > 2 else 3
max_batch = int(sys.argv[3]) if len(sys.argv) > 3 else 37888
T = ((299 - 32 + interval - 1) // interval + 1) ** 2
n_bags = -(-max_batch // T)
bags = synthetic.make_bags_device(n_bags, dev, seed=0)
c, fw, fb = synthetic.make_resnet_weights(arch, seed=0)
clf = ops.TileClassifier(arch, c, fw, fb, device=dev)
for _ in range(2):
    p = clf.forward_tiles(bags, 32, interval, inst_count=max_batch, precision="bf16", max_batch=max_batch)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    p = clf.forward_tiles(bags, 32, interval, inst_count=max_batch, precision="bf16", max_batch=max_batch)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print(arch, "interval", interval, "instances", max_batch, "launches per batch", clf.last_launch_count,
      "| %.3f ms per batch, %.4g instances/s" % (ms, max_batch / ms * 1e3))
