"""Sweep of the HSV-refine launch shape (run on a B200): GB/s at several sizes / grid caps."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
from cellsegmentation_b200 import ops
dev = torch.device("cuda", 0)
for bags in (1792, 8000, 20000):
    n = bags * 299 * 299
    img = torch.randint(0, 256, (n, 3), dtype=torch.uint8, device=dev)
    mask = (torch.rand(n, device=dev) < 0.3).to(torch.uint8)
    out = torch.empty_like(mask)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    best = 1e9
    for _ in range(6):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.hsv_refine(img, mask, 170, out=out); b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    src = torch.empty(5 * n // 2, dtype=torch.uint8, device=dev); dst = torch.empty_like(src)
    cb = 1e9
    for _ in range(4):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); dst.copy_(src); b.record()
        torch.cuda.synchronize()
        cb = min(cb, a.elapsed_time(b))
    print("bags %%5d  hsv %%.3f ms  %%.0f GB/s | copy of the same bytes %%.3f ms %%.0f GB/s" %% (
        bags, best, 5 * n / best / 1e6, cb, 5 * n / cb / 1e6), flush=True)
    del img, mask, out, src, dst
''' % ROOT

for cps in (4, 8, 16, 32, 64):
    print("CELLSEG_HSV_CTAS_PER_SM =", cps, flush=True)
    env = dict(os.environ, CELLSEG_HSV_CTAS_PER_SM=str(cps))
    subprocess.run([sys.executable, "-c", CHILD], env=env, check=False)
