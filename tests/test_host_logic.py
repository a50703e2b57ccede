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


def _spawn(target, world, *args):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + _spawn.calls * 7
    _spawn.calls += 1
    procs = [ctx.Process(target=target, args=(r, world, port, q) + args) for r in range(world)]
    [p.start() for p in procs]
    res = sorted((q.get(timeout=180) for _ in range(world)), key=lambda t: t[0])
    [p.join(60) for p in procs]
    return res


_spawn.calls = 0


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cellsegmentation_b200.distributed import allgather_selection, allreduce_flat, broadcast_seed
    # each rank "selected" a different number of tiles of its own shard; the buffers are
    # capacity-sized and the count lives in a tensor (on the GPU path it never visits the host)
    n, cap = 3 + 2 * rank, 6
    idx = torch.full((cap,), -7, dtype=torch.int32)
    idx[:n] = torch.arange(n, dtype=torch.int32) * 2
    lab = torch.full((cap,), rank, dtype=torch.uint8)
    gi, gl = allgather_selection(idx, lab, torch.tensor([n]), tile_offset=1000 * rank, capacity=cap)
    g1, g2 = torch.full((2, 4), float(rank + 1)), torch.tensor([10.0 * (rank + 1)])
    flat = allreduce_flat([g1, g2])
    np.random.seed(100 + rank)                     # different states: only rank 0's draw may count
    seed_none = broadcast_seed(None, torch.device("cpu"))
    seed_given = broadcast_seed(41 + rank, torch.device("cpu"))
    q.put((rank, gi.tolist(), gl.tolist(), flat.tolist(), seed_none, seed_given))
    dist.destroy_process_group()


def test_gloo_world2_allgather_and_grad_allreduce():
    res = _spawn(_worker, 2)
    want_idx = [0, 2, 4] + [1000, 1002, 1004, 1006, 1008]
    np.random.seed(100)
    want_seed = int(np.random.randint(0, 2 ** 31 - 1))
    for rank, gi, gl, flat, seed_none, seed_given in res:
        assert gi == want_idx
        assert gl == [0, 0, 0, 1, 1, 1, 1, 1]
        assert flat == [3.0] * 8 + [30.0]          # one bucket: sum over ranks
        assert seed_none == want_seed and seed_given == 41


def test_selection_capacity_is_the_closed_form_bound():
    from cellsegmentation_b200.distributed import kept_upper_bound, selection_capacity
    labels = [3, 0, 7, 1, 300, 0]
    assert kept_upper_bound(labels, 225, 1, 30).tolist() == [3, 30, 7, 1, 225, 30]
    assert selection_capacity(labels, 225, 1, 30, 1) == 296
    assert selection_capacity(labels, 225, 1, 30, 2) == 256      # shards [3,0,7] and [1,300,0]
    assert selection_capacity(labels, 225, 2, 30, 4) == 255      # shards of 2 bags; [300, 0] -> 225 + 30
    # the bound holds for the literal predicate on a toy set, wrap-around included
    tid = np.repeat(np.arange(6), 225)
    rng = np.random.default_rng(0)
    kept = oselect.sample_indices(tid, np.array(labels), rng.random(len(tid)).astype(np.float32), 1, 30)
    per_bag = np.bincount(tid[kept], minlength=6)
    assert (per_bag <= kept_upper_bound(labels, 225, 1, 30)).all()


class _FakeModel(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.fc_tile = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(8, 2))


def _train_worker(rank, world, port, q, crit_kind):
    """train_selected with a feature cache is torch-only: run it on CPU tensors under gloo."""
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from cellsegmentation_b200 import mil
    from cellsegmentation_b200.distributed import shard_dataset
    ds = _ds(7, (2, 0, 5, 1, 0, 9, 3))
    T = ds.tiles_per_bag
    g = torch.Generator().manual_seed(5)
    feat_all = torch.randn(ds.num_tiles(), 8, generator=g)
    idxs = np.sort(np.random.default_rng(1).choice(ds.num_tiles(), 300, replace=False))
    np.random.seed(11)
    ds.make_train_data(idxs, 0.5)
    shard, off = shard_dataset(ds, rank, world)
    cache = {"feat": feat_all[off:off + shard.num_tiles()], "begin": off, "end": off + shard.num_tiles()}
    torch.manual_seed(3)
    net = _FakeModel()
    crit = {"ce": torch.nn.CrossEntropyLoss(),
            "ce_weighted": torch.nn.CrossEntropyLoss(weight=torch.tensor([0.3, 1.7]), label_smoothing=0.1),
            "nll_like": lambda o, l: torch.nn.functional.nll_loss(torch.log_softmax(o, 1), l)}[crit_kind]
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=0.2, total_steps=50)
    loss = mil.train_selected(ds, net, torch.device("cpu"), crit, opt, 64, gamma=1.0, shuffle_seed=2,
                              feature_cache=cache, scheduler=sched)
    q.put((rank, loss, net.fc_tile[1].weight.detach().numpy().copy(), net.fc_tile[1].bias.detach().numpy().copy(),
           sched.last_epoch))
    if world > 1:
        dist.destroy_process_group()


@pytest.mark.parametrize("crit_kind", ["ce", "ce_weighted", "nll_like"])
def test_gloo_world2_train_selected_equals_single_process(crit_kind):
    """the data-parallel step must reproduce single-process training (loss and weights),
    including class-weighted / label-smoothed CE and criteria that are plain per-row means."""
    (_, loss1, w1, b1, steps1), = _spawn(_train_worker, 1, crit_kind)
    res = _spawn(_train_worker, 2, crit_kind)
    for rank, loss, w, b, steps in res:
        assert steps == steps1 > 0
        assert abs(loss - loss1) < 1e-5, (loss, loss1)
        np.testing.assert_allclose(w, w1, rtol=0, atol=2e-6)
        np.testing.assert_allclose(b, b1, rtol=0, atol=2e-6)
    assert np.array_equal(res[0][2], res[1][2])


@pytest.mark.parametrize("arch", ["resnet18", "resnet50"])
def test_scratch_autograd_path_matches_reference_modules(arch):
    """train_tile.py --scratch (:272-273) trains the encoder: the mirror then evaluates its module
    tree with torch autograd.  Same logits and gradients as the reference class on CPU (only
    where /root/reference exists: the build container)."""
    from oracle import model as omodel, ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    from cellsegmentation_b200.model import resnet as mirror
    sd = omodel.make_state_dict(arch, seed=3)
    nets = []
    for mod in (mirror, ref_shim.load_reference_resnet()):
        m = getattr(mod, "MIL" + arch)()
        m.load_state_dict(sd, strict=False)
        m.setmode("tile")
        m.set_encoder_grads(True)
        m.train()
        nets.append(m)
    x = torch.randn(5, 3, 32, 32, generator=torch.Generator().manual_seed(1))
    for freeze in (True, False):
        outs = [m(x, freeze_bn=freeze) for m in nets]
        assert torch.equal(outs[0], outs[1])
        assert nets[0].training and nets[1].training
        for m, o in zip(nets, outs):
            m.zero_grad()
            o.square().sum().backward()
        assert torch.equal(nets[0].layer2[0].conv1.weight.grad, nets[1].layer2[0].conv1.weight.grad)
        assert torch.equal(nets[0].bn1.running_mean, nets[1].bn1.running_mean)


def test_grid_instances_inverts_the_tile_grid():
    """utils.image_processing._grid_instances: (row, col, bag) of grid tiles -> bag * T + position,
    None for a tile off the grid (heatmap() then keeps the explicit-coordinate painter)."""
    from cellsegmentation_b200.utils import image_processing as ip

    class DS:
        has_tiles = True
        image_size = (100, 131)
        tile_size, interval = 32, 7

        def __init__(self):
            self.images = [None] * 4
            self._g = np.array(otiles.get_tiles((100, 131, 3), 7, 32), np.int32)

        def _ensure_grid(self):
            return self._g

    ds = DS()
    T = len(ds._g)
    rng = np.random.default_rng(0)
    t = rng.integers(0, T, 200)
    g = rng.integers(0, 4, 200)
    inst = ip._grid_instances(ds, ds._g[t], g)
    assert inst is not None and np.array_equal(inst, g * T + t)
    # the final (clamped) position of either axis and position 0
    edge = np.array([0, T - 1, len(set(ds._g[:, 1])) - 1])
    assert np.array_equal(ip._grid_instances(ds, ds._g[edge], np.zeros(3, int)), edge)
    off = ds._g[t].copy()
    off[3, 1] += 1
    assert ip._grid_instances(ds, off, g) is None
    assert ip._grid_instances(ds, ds._g[t], np.where(np.arange(200) == 7, 4, g)) is None
