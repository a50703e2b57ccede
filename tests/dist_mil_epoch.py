"""Multi-GPU check of the MIL epoch (run with torchrun, one rank per GPU):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_mil_epoch.py
Every rank must end with (a) the selection the single-process path produces, (b) the same
train_data as every other rank also when no seed is given (rank 0's draw is broadcast), and
(c) the fc_tile weights and epoch loss of single-process train_tile on that train_data, with
and without the feature cache.  Rank 0 prints PASS."""
import copy
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
from cellsegmentation_b200.train import train_tile  # noqa: E402
from oracle import model as omodel, synth, tiles as otiles  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    bags = synth.make_bags(9, seed=5)
    labels = [2, 0, 5, 1, 0, 9, 3, 0, 40]
    ds = LystoDataset.from_arrays(list(bags), labels, 32, 20)
    x = torch.from_numpy(otiles.unfold(list(bags[1:3]), 20, 32))
    sd = omodel.calibrate_head(omodel.make_state_dict("resnet34", seed=3), x[::3], "resnet34")

    def make_net():
        m = MILresnet34()
        m.load_state_dict(sd, strict=False)
        m.setmode("tile")
        m.max_batch = 512
        return m.to(dev)

    net = make_net()
    ds.setmode(1)
    loader = torch.utils.data.DataLoader(ds, batch_size=16, shuffle=False)
    probs = inference_tiles(loader, net, dev)                 # whole set on every rank: the oracle
    want_idx, want_pl = sample_indices(ds, probs, 1, 30, device=dev)
    gidx, glab = select_global(ds, net, dev, 1, 30)           # sharded + one packed all-gather
    ok = np.array_equal(gidx, want_idx) and np.array_equal(glab, want_pl)
    notes = []
    if not ok:
        notes.append("selection differs")

    def single_process(seed):
        """The reference sequence on ONE process: make_train_data under `seed`, then train_tile."""
        ref_ds = copy.copy(ds)
        ref_ds._shard_cache = None
        ref_net = make_net()
        opt = torch.optim.SGD(filter(lambda p: p.requires_grad, ref_net.parameters()), lr=2e-5)
        np.random.seed(seed)
        ref_ds.make_train_data(want_idx, 0.5, pseudo_labels=want_pl)
        ref_ds.setmode(3)
        ld = torch.utils.data.DataLoader(ref_ds, batch_size=16, shuffle=False)
        loss = train_tile(ld, 1, 1, ref_net, dev, torch.nn.CrossEntropyLoss(), opt, None, 1.0)
        return loss, ref_net.fc_tile[1].weight.detach().clone(), ref_ds.train_data.copy()

    for cache in (True, False):
        for seed in (3, None):
            if seed is None and world == 1:
                continue                                       # nothing to agree on
            run_net = make_net()
            opt = torch.optim.SGD(filter(lambda p: p.requires_grad, run_net.parameters()), lr=2e-5)
            np.random.seed(1000 + rank)                        # ranks disagree unless the seed is shared
            if seed is None and rank == 0:
                st = np.random.get_state()
                eff = int(np.random.randint(0, 2 ** 31 - 1))   # what broadcast_seed will draw on rank 0
                np.random.set_state(st)
            else:
                eff = seed
            eff_t = torch.tensor([eff if eff is not None else 0], device=dev)
            dist.broadcast(eff_t, src=0)
            loss, pos, neg = mil_epoch(ds, run_net, dev, torch.nn.CrossEntropyLoss(), opt, 1, 30, 0.5, 16,
                                       seed=seed, cache_features=cache)
            w_loss, w_w, w_td = single_process(int(eff_t.item()))
            w = run_net.fc_tile[1].weight.detach()
            same_td = np.array_equal(ds.train_data, w_td)
            dw = float((w - w_w).abs().max())
            dl = abs(loss - w_loss)
            ws = [torch.empty_like(w) for _ in range(world)]
            dist.all_gather(ws, w.contiguous())
            same = all(torch.equal(ws[0], t) for t in ws)
            good = same_td and same and dw < 5e-6 and dl < 2e-5 * max(1.0, abs(w_loss)) and np.isfinite(loss)
            if not good:
                notes.append("cache=%s seed=%s: same_td %s ranks_equal %s dw %.3g dl %.3g" % (cache, seed, same_td, same, dw, dl))
            ok = ok and good
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if notes:
        print("rank", rank, "; ".join(notes))
    if rank == 0:
        print("PASS" if int(flag.item()) == 1 else "FAIL", "world", world, "selected", len(gidx), "pos/neg", pos, neg)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
