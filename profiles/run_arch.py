import sys, torch
sys.path.insert(0, "/root/repo")
from cellsegmentation_b200 import ops, synthetic
dev = torch.device("cuda", 0)
arch = sys.argv[1] if len(sys.argv) > 1 else "resnext50_32x4d"
bags = synthetic.make_bags_device(8, dev, seed=0)
c, fw, fb = synthetic.make_resnet_weights(arch, seed=0)
clf = ops.TileClassifier(arch, c, fw, fb, device=dev)
for _ in range(2):
    p = clf.forward_tiles(bags, 32, 5, precision="bf16", max_batch=18944)
torch.cuda.synchronize()
print(p.shape, clf.last_launch_count)
