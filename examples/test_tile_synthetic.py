#!/usr/bin/env python
"""The body of the reference's test_tile.py (:82-111) and of test_seg.py --draw_masks, with the
drop-in modules and synthetic LYSTO-shaped bags (no .h5 file, random-init weights):

    inference_tiles -> rank -> heatmap (CSV + PNGs) -> generate_masks(preprocess=True)

    python examples/test_tile_synthetic.py --bags 64 --out /tmp/cellseg_demo
"""
import argparse
import csv
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cellsegmentation_b200 import synthetic  # noqa: E402
from cellsegmentation_b200.dataset import LystoTestset  # noqa: E402
from cellsegmentation_b200.inference import inference_tiles, rank  # noqa: E402
from cellsegmentation_b200.model import nets  # noqa: E402
from cellsegmentation_b200.utils import generate_masks, heatmap  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bags", type=int, default=64)
    ap.add_argument("--encoder", default="resnet34")
    ap.add_argument("--interval", type=int, default=5)
    ap.add_argument("--threshold", type=float, default=None,
                    help="rank() threshold (default: the 98th percentile of the probabilities, since the\n"
                         "random-init head does not reach the reference's 0.95)")
    ap.add_argument("--out", default="/tmp/cellseg_demo")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    os.makedirs(os.path.join(args.out, "heatmap"), exist_ok=True)

    images = list(synthetic.make_bags_device(args.bags, dev, seed=0).cpu().numpy())
    testset = LystoTestset.from_arrays(images, tile_size=32, interval=args.interval)
    loader = torch.utils.data.DataLoader(testset, batch_size=40960, shuffle=False)
    model = nets[args.encoder]
    model.setmode("tile")
    model.to(dev)
    model.eval()

    testset.setmode("tile")
    t0 = time.perf_counter()
    probs = inference_tiles(loader, model, dev, mode="test")                 # test_tile.py:84
    t1 = time.perf_counter()
    thr = args.threshold if args.threshold is not None else float(np.quantile(probs, 0.98))
    tiles, kept_probs, groups = rank(testset, probs, thr)                    # test_tile.py:86
    t2 = time.perf_counter()
    with open(os.path.join(args.out, "pred.csv"), "w", newline="") as f:
        csv.writer(f).writerow(["id", "grid", "prob"])
        heatmap(testset, tiles, kept_probs, groups, f, os.path.join(args.out, "heatmap"))   # :110
    t3 = time.perf_counter()
    masks = generate_masks(testset, tiles, groups, preprocess=True, save_masks=True,
                           output_path=os.path.join(args.out, "pseudomask"))  # test_seg.py --draw_masks
    t4 = time.perf_counter()
    print("%d bags, %d instances: inference %.3f s (%.2f M instances/s), rank %.3f s (%d kept), "
          "heatmaps + PNG %.3f s, masks + clean-up + PNG %.3f s, mask pixels %d"
          % (args.bags, len(probs), t1 - t0, len(probs) / (t1 - t0) / 1e6, t2 - t1, len(kept_probs), t3 - t2,
             t4 - t3, int(np.asarray(masks).sum())))


if __name__ == "__main__":
    main()
