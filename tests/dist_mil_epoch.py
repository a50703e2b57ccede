"""Multi-GPU check of the MIL epoch (run with torchrun, one rank per GPU):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_mil_epoch.py
Every rank must end with the selection the single-process path produces and with identical
fc_tile weights; rank 0 prints PASS."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cellsegmentation_b200.dataset import LystoDataset  # noqa: E402
from cellsegmentation_b200.inference import inference_tiles, sample_indices  # noqa: E402
from cellsegmentation_b200.mil import mil_epoch, select_global  # noqa: E402
from cellsegmentation_b200.model.resnet import MILresnet34  # noqa: E402
from oracle import model as omodel, synth, tiles as otiles  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    bags = synth.make_bags(7, seed=5)
    labels = [2, 0, 5, 1, 0, 9, 3]
    ds = LystoDataset.from_arrays(list(bags), labels, 32, 20)
    x = torch.from_numpy(otiles.unfold(list(bags[1:3]), 20, 32))
    sd = omodel.calibrate_head(omodel.make_state_dict("resnet34", seed=3), x[::3], "resnet34")
    net = MILresnet34()
    net.load_state_dict(sd, strict=False)
    net.setmode("tile")
    net.max_batch = 512
    net.to(dev)
    ds.setmode(1)
    loader = torch.utils.data.DataLoader(ds, batch_size=64, shuffle=False)
    probs = inference_tiles(loader, net, dev)                 # whole set on every rank: the oracle
    want_idx, want_pl = sample_indices(ds, probs, 1, 30, device=dev)
    gidx, glab = select_global(ds, net, dev, 1, 30)           # sharded + all-gather
    ok = np.array_equal(gidx, want_idx) and np.array_equal(glab, want_pl)
    opt = torch.optim.SGD(filter(lambda p: p.requires_grad, net.parameters()), lr=1e-6)
    loss, pos, neg = mil_epoch(ds, net, dev, torch.nn.CrossEntropyLoss(), opt, 1, 30, 0.5, 16, seed=3)
    w = net.fc_tile[1].weight.detach().reshape(-1)
    ws = [torch.empty_like(w) for _ in range(world)]
    dist.all_gather(ws, w)
    same = all(torch.equal(ws[0], t) for t in ws)
    flag = torch.tensor([int(ok and same and np.isfinite(loss))], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("PASS" if int(flag.item()) == 1 else "FAIL", "world", world, "selected", len(gidx), "loss", loss,
              "pos/neg", pos, neg)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
