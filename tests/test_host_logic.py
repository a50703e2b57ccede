"""Host-side logic that needs no GPU: lazy tile sets, make_train_data, sharding, gloo collectives."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import select as oselect, tiles as otiles


def _ds(n=4, labels=(9, 3, 0, 7), S=32, I=20):
    from cellsegmentation_b200.dataset import LystoDataset
    imgs = [np.full((299, 299, 3), i, np.uint8) for i in range(n)]
    return LystoDataset.from_arrays(imgs, labels, S, I)


def test_lazy_tile_sequences_match_reference_lists():
    ds = _ds()
    ds.setmode(1)
    grid = otiles.get_tiles((299, 299, 3), 20, 32)
    assert len(ds) == 3 * 225                               # bag 0 owns no tiles
    assert list(np.array(ds.tileIDX)) == [1] * 225 + [2] * 225 + [3] * 225
    assert ds.tileIDX[0] == 1 and ds.tileIDX[-1] == 3 and ds.tileIDX[225] == 2
    assert [tuple(t) for t in np.array(ds.tiles_grid)] == grid * 3
    assert ds.tiles_grid[226] == grid[1]
    assert list(ds.seg_offsets()) == [0, 0, 225, 450, 675]
    from cellsegmentation_b200.dataset import get_tiles
    assert get_tiles(np.zeros((299, 299, 3)), 5, 32) == otiles.get_tiles((299, 299, 3), 5, 32)


def test_testset_every_bag_owns_tiles():
    from cellsegmentation_b200.dataset import LystoTestset
    ts = LystoTestset.from_arrays([np.zeros((299, 299, 3), np.uint8)] * 2, 32, 5)
    ts.setmode("tile")
    assert len(ts) == 2 * 3025 and ts.tileIDX[0] == 0 and ts.first_tile_bag == 0


@pytest.mark.parametrize("ratio", [0.5, 2.0, None])
def test_make_train_data_matches_oracle_restatement(ratio):
    ds = _ds(5, (9, 3, 0, 7, 0))
    grid = otiles.get_tiles((299, 299, 3), 20, 32)
    tid = np.array(ds.tileIDX)
    tiles_grid = [grid[i % 225] for i in range(len(tid))]
    idxs = np.arange(3, 900, 5)
    np.random.seed(7)
    want, wp, wn = oselect.make_train_data(tid, tiles_grid, ds.labels, idxs, ratio)
    np.random.seed(7)
    pos, neg = ds.make_train_data(idxs, ratio)
    assert (pos, neg) == (wp, wn)
    got = ds.train_data
    assert len(got) == len(want)
    assert [int(r[0]) for r in want] == list(got["bag"])
    assert [tuple(r[1]) for r in want] == list(zip(got["x"].tolist(), got["y"].tolist()))
    assert [int(r[2]) for r in want] == list(got["label"])


def test_shard_range_and_dataset_shards():
    from cellsegmentation_b200.distributed import shard_dataset, shard_range
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert [shard_range(2, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
    ds = _ds(7, (1, 2, 3, 4, 5, 6, 7))
    tot, seen = 0, []
    for r in range(3):
        sh, off = shard_dataset(ds, r, 3)
        assert off == tot
        tot += sh.num_tiles()
        seen += [sh.labels[b] for b in sh._tile_bags]
    assert tot == ds.num_tiles() and seen == [2, 3, 4, 5, 6, 7]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cellsegmentation_b200.distributed import allgather_selection, allreduce_mean_grads
    # each rank "selected" a different number of tiles of its own shard
    n = 3 + 2 * rank
    idx = torch.arange(n, dtype=torch.int32) * 2
    lab = torch.full((n,), rank, dtype=torch.uint8)
    gi, gl = allgather_selection(idx, lab, tile_offset=1000 * rank)
    lin = torch.nn.Linear(4, 2)
    with torch.no_grad():
        lin.weight.fill_(0.5); lin.bias.zero_()
    lin(torch.full((1, 4), float(rank + 1))).sum().backward()
    allreduce_mean_grads(list(lin.parameters()))
    q.put((rank, gi.tolist(), gl.tolist(), lin.weight.grad[0].tolist()))
    dist.destroy_process_group()


def test_gloo_world2_allgather_and_grad_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    want_idx = [0, 2, 4] + [1000, 1002, 1004, 1006, 1008]
    for rank, gi, gl, g in res:
        assert gi == want_idx
        assert gl == [0, 0, 0, 1, 1, 1, 1, 1]
        assert g == [1.5] * 4                      # mean of grads 1.0 and 2.0
