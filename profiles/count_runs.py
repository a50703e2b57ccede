import sys, torch
sys.path.insert(0, "/root/repo")
from cellsegmentation_b200 import ops, synthetic
dev = torch.device("cuda", 0)
H = 299
nb = 64
imgs = synthetic.make_bags_device(nb, dev, seed=99)
g = torch.Generator(device=dev); g.manual_seed(7)
blob = torch.rand((nb, H // 13 + 1, H // 13 + 1), device=dev, generator=g) < 0.45
masks = blob.repeat_interleave(13, 1).repeat_interleave(13, 2)[:, :H, :H].contiguous().to(torch.uint8)
out = ops.hsv_refine(imgs, masks, 170)
x = out.to(torch.int16)
pad = torch.zeros((nb, H, 1), dtype=torch.int16, device=dev)
d = torch.diff(torch.cat([pad, x, pad], 2), dim=2)
fg = (d == 1).sum((1, 2)); 
xi = 1 - x
d2 = torch.diff(torch.cat([pad, xi, pad], 2), dim=2)
bg = (d2 == 1).sum((1, 2))
print("fg runs", fg.float().mean().item(), fg.max().item(), "bg runs", bg.float().mean().item(), bg.max().item(), "fg px", out.float().sum((1,2)).mean().item())
work = out.clone()
for _ in range(3):
    work.copy_(out)
    ops.remove_small_regions(work, 400, 120)
torch.cuda.synchronize()
