"""GPU tests of the reference-facing Python surfaces (drop-in API) against the oracle / goldens."""
import contextlib
import io
import os

import numpy as np
import pytest
import torch

from conftest import golden
from oracle import masks as omasks, model as omodel, select as oselect, synth, tiles as otiles

pytestmark = pytest.mark.gpu


def _model(arch, sd, cuda, precision):
    from cellsegmentation_b200.model import nets
    from cellsegmentation_b200.model.resnet import MILresnet18, MILresnet34, MILresnet50
    from cellsegmentation_b200.model.resnext import MILresnext50_32x4d
    from cellsegmentation_b200.model.resnext import MILresnext101_32x8d
    net = {"resnet18": MILresnet18, "resnet34": MILresnet34, "resnet50": MILresnet50,
           "resnext50_32x4d": MILresnext50_32x4d, "resnext101_32x8d": MILresnext101_32x8d}[arch]()
    missing, unexpected = net.load_state_dict(sd, strict=False)
    # the oracle state holds the encoder + fc_tile; Stage-1 heads / decoder keep their init
    assert not unexpected and all(k.startswith(net.image_module_prefix + net.seg_module_prefix) for k in missing), \
        (missing, unexpected)
    net.setmode("tile")
    net.precision = precision
    net.max_batch = 512
    assert arch in nets and type(nets[arch]) is type(net)
    return net.to(cuda)


def _trainset(n_bags=3, interval=20):
    from cellsegmentation_b200.dataset import LystoDataset
    bags = synth.make_bags(n_bags, seed=11)
    return bags, LystoDataset.from_arrays(list(bags), [4, 0, 9][:n_bags], 32, interval)


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_inference_tiles_dataset_path_matches_reference_golden(cuda, precision, tol):
    from cellsegmentation_b200.inference import inference_tiles
    g = golden("model_resnet34.npz")
    bags, ds = _trainset()
    ds.setmode(1)
    x = torch.from_numpy(otiles.unfold(list(bags[1:]), 20, 32))
    sd = omodel.calibrate_head(omodel.make_state_dict("resnet34", seed=3), x[::3], "resnet34")
    net = _model("resnet34", sd, cuda, precision)
    loader = torch.utils.data.DataLoader(ds, batch_size=64, shuffle=False)
    probs = inference_tiles(loader, net, cuda, mode="train")
    assert probs.dtype == np.float32 and probs.shape == (450,)
    assert np.abs(probs - g["probs"]).max() < tol          # reference's own inference_tiles output
    # drop-in module call on a materialised batch == reference logits
    net.eval()
    logits = net(x[:16].to(cuda)).detach().cpu().numpy()
    assert np.abs(logits - g["logits16"]).max() < (1e-3 if precision == "fp32" else 0.15)
    # __getitem__ compatibility (mode 1): same tensor the reference's dataset returns
    t, lab = ds[14]
    gt = golden("transform.npz")
    assert np.array_equal(t.numpy().view(np.uint32), gt["tiles"][list(gt["pick"]).index(14)].view(np.uint32))
    assert lab == 0 or lab == ds.labels[ds.tileIDX[14]]


@pytest.mark.parametrize("arch", ["resnet50", "resnext50_32x4d", "resnext101_32x8d"])
def test_bottleneck_nets_match_reference_golden(cuda, arch):
    """BASELINE config 4 (encoder swap): reference state_dict keys load, module call == golden."""
    from cellsegmentation_b200.inference import inference_tiles
    g = golden("model_%s.npz" % arch)
    bags, ds = _trainset()
    ds.setmode(1)
    x = torch.from_numpy(otiles.unfold(list(bags[1:]), 20, 32))
    sd = omodel.calibrate_head(omodel.make_state_dict(arch, seed=3), x[::3], arch)
    for precision, tol, ltol in (("fp32", 1e-4, 2e-3), ("bf16", 2e-2, 0.2)):
        net = _model(arch, sd, cuda, precision)
        net.eval()
        loader = torch.utils.data.DataLoader(ds, batch_size=64, shuffle=False)
        probs = inference_tiles(loader, net, cuda, mode="train")
        assert np.abs(probs - g["probs"]).max() < tol
        logits = net(x[:16].to(cuda)).detach().cpu().numpy()
        assert np.abs(logits - g["logits16"]).max() < ltol
        feat = net.encode(x[:16].to(cuda))
        assert feat.shape == (16, 2048)


def test_sample_rebuilds_reference_selection(cuda, capsys):
    from cellsegmentation_b200.dataset import LystoDataset
    from cellsegmentation_b200.inference import sample, sample_indices
    g = golden("select.npz")
    bag = synth.make_bags(1, seed=1)[0]
    for case in ("toy", "ties", "nan", "wrapbig", "k0"):
        lab = g[case + "_labels"]
        ds = LystoDataset.from_arrays([bag] * len(lab), lab, 32, 20)
        ds.setmode(1)
        tpp, tk = (int(v) for v in g[case + "_params"])
        idx, pl = sample_indices(ds, g[case + "_probs"], tpp, tk)
        assert np.array_equal(idx, g[case + "_idx"])
        assert np.array_equal(pl, oselect.pseudo_labels(g[case + "_tileIDX"], lab, g[case + "_idx"]))
    # full sample(): same train_data as the oracle restatement of make_train_data
    lab = g["toy_labels"]
    ds = LystoDataset.from_arrays([bag] * len(lab), lab, 32, 20)
    ds.setmode(1)
    grid = otiles.get_tiles((299, 299, 3), 20, 32)
    tiles_grid = [grid[i % 225] for i in range(len(g["toy_tileIDX"]))]
    np.random.seed(5)
    want, wp, wn = oselect.make_train_data(g["toy_tileIDX"], tiles_grid, lab, g["toy_idx"], 0.5)
    np.random.seed(5)
    sample(ds, g["toy_probs"], 1, 30, 0.5)
    assert "Training data is sampled. (Pos samples: %d | Neg samples: %d)" % (wp, wn) in capsys.readouterr().out
    assert [int(r[0]) for r in want] == list(ds.train_data["bag"])
    assert [tuple(r[1]) for r in want] == list(zip(ds.train_data["x"].tolist(), ds.train_data["y"].tolist()))
    ds.setmode(3)
    assert len(ds) == len(want)
    t, l = ds[0]
    b, (x, y), wl = want[0]
    assert np.array_equal(t.numpy(), otiles.normalize_tile(bag[x:x + 32, y:y + 32])) and l == wl


def test_rank_and_evaluate_tile_match_reference(cuda):
    from cellsegmentation_b200.dataset import LystoDataset
    from cellsegmentation_b200.evaluate import evaluate_tile
    from cellsegmentation_b200.inference import rank
    bag = synth.make_bags(1, seed=2)[0]
    g = golden("rank.npz")
    ds = LystoDataset.from_arrays([bag] * 4, [1, 2, 3, 4], 32, 20)
    tiles, probs, groups = rank(ds, g["probs"], float(g["threshold"]))
    assert np.array_equal(tiles.astype(np.int32), g["tiles"])
    assert np.array_equal(probs.view(np.uint32), g["kept_probs"].view(np.uint32))
    assert np.array_equal(groups.astype(np.int32), g["groups"])
    e = golden("evaluate.npz")
    dse = LystoDataset.from_arrays([bag] * 5, e["labels"], 32, 20)
    out = evaluate_tile(dse, e["probs"], int(e["params"][0]), float(e["params"][1]))
    assert np.allclose(out, e["out"], rtol=0, atol=1e-15)


def test_heatmap_and_generate_masks_match_reference(cuda, tmp_path):
    import types
    from cellsegmentation_b200 import utils
    g = golden("masks.npz")
    small = synth.make_bags(3, H=96, W=96, seed=31)
    grid = np.array(otiles.get_tiles((96, 96, 3), 5, 16), np.int32)
    T = len(grid)
    keep = g["kept"]
    k_tiles, k_probs, k_groups = grid[keep % T], g["probs"][keep], keep // T

    class Fake:
        images = list(small)
        image_size = np.array([96, 96])
        tile_size = 16

        def device_images(self, dev):
            return torch.from_numpy(small).to(dev)

    fake = Fake()
    csvf = io.StringIO(newline="")
    utils.heatmap(fake, k_tiles, k_probs, k_groups, csvf, str(tmp_path))
    # the golden text was read back with universal newlines; the writer emits \r\n like the reference
    assert csvf.getvalue().replace("\r\n", "\n").encode() == g["csv"].tobytes()
    import cv2
    for i in range(3):
        img = cv2.imread(os.path.join(str(tmp_path), "test_%05d.png" % (i + 1)))[..., ::-1]
        assert np.array_equal(img, g["heat_imgs"][i])
    raw = utils.generate_masks(fake, k_tiles, k_groups, preprocess=False, save_masks=False,
                               output_path=str(tmp_path))
    assert np.array_equal(raw, g["raw"])
    full = utils.generate_masks(fake, k_tiles, k_groups, preprocess=True, save_masks=True,
                                output_path=str(tmp_path))
    assert np.array_equal(full.astype(np.uint8), g["full"])
    m = cv2.imread(os.path.join(str(tmp_path), "mask", "00002.png"), cv2.IMREAD_GRAYSCALE)
    assert np.array_equal(m, g["full"][1] * 255)
    one = utils.preprocess_masks(small[0], g["raw"][0])
    assert np.array_equal(one.astype(np.uint8), g["full"][0])


def test_train_tile_updates_only_fc_tile(cuda):
    from cellsegmentation_b200.inference import inference_tiles, sample
    from cellsegmentation_b200.train import train_tile
    bags, ds = _trainset()
    ds.setmode(1)
    x = torch.from_numpy(otiles.unfold(list(bags[1:]), 20, 32))
    sd = omodel.calibrate_head(omodel.make_state_dict("resnet34", seed=3), x[::3], "resnet34")
    net = _model("resnet34", sd, cuda, "bf16")
    loader = torch.utils.data.DataLoader(ds, batch_size=64, shuffle=True)
    probs = inference_tiles(loader, net, cuda)
    np.random.seed(0)
    sample(ds, probs, 1, 30, 0.5)
    ds.setmode(3)
    enc_before = net.layer3[2].conv1.weight.detach().clone()
    fc_before = net.fc_tile[1].weight.detach().clone()
    opt = torch.optim.Adam(filter(lambda p: p.requires_grad, net.parameters()), lr=1e-5, weight_decay=1e-4)
    crit = torch.nn.CrossEntropyLoss()
    losses = [train_tile(loader, e, 6, net, cuda, crit, opt, None, 1.) for e in range(1, 7)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]
    assert torch.equal(enc_before, net.layer3[2].conv1.weight.detach())
    assert not torch.equal(fc_before, net.fc_tile[1].weight.detach())
    # the device-side fc follows the optimizer: probabilities change, and match the module's own head
    ds.setmode(1)
    probs2 = inference_tiles(loader, net, cuda)
    assert np.abs(probs2 - probs).max() > 1e-6
    net.eval()
    with torch.no_grad():
        p_mod = torch.softmax(net(x[:64].to(cuda)), 1)[:, 1].cpu().numpy()
    assert np.abs(p_mod - probs2[:64]).max() < 1e-4


def test_train_tile_matches_reference_loop_fp32(cuda):
    """One epoch of train_tile (SGD, loader order) against the oracle restatement of
    train/train.py:12-48: mean loss and the updated fc_tile within fp32 tolerance."""
    from oracle import train as otrain
    from cellsegmentation_b200.inference import inference_tiles, sample
    from cellsegmentation_b200.train import train_tile
    bags, ds = _trainset()
    ds.setmode(1)
    x = torch.from_numpy(otiles.unfold(list(bags[1:]), 20, 32))
    sd = omodel.calibrate_head(omodel.make_state_dict("resnet34", seed=3), x[::3], "resnet34")
    net = _model("resnet34", sd, cuda, "fp32")
    loader = torch.utils.data.DataLoader(ds, batch_size=16, shuffle=False)
    probs = inference_tiles(loader, net, cuda)
    np.random.seed(1)
    sample(ds, probs, 1, 30, 0.5)
    ds.setmode(3)
    td = ds.train_data
    tiles = torch.stack([torch.from_numpy(otiles.normalize_tile(bags[r["bag"]][r["x"]:r["x"] + 32, r["y"]:r["y"] + 32]))
                         for r in td])
    labels = torch.from_numpy(td["label"].copy())
    want_loss, want_w, want_b = otrain.train_tile_epoch(sd, tiles, labels, 16, lr=1e-6)
    opt = torch.optim.SGD(filter(lambda p: p.requires_grad, net.parameters()), lr=1e-6)
    got_loss = train_tile(loader, 1, 1, net, cuda, torch.nn.CrossEntropyLoss(), opt, None, 1.)
    assert abs(got_loss - want_loss) < 1e-4 * max(1.0, abs(want_loss))
    assert (net.fc_tile[1].weight.detach().cpu() - want_w).abs().max() < 1e-6
    assert (net.fc_tile[1].bias.detach().cpu() - want_b).abs().max() < 1e-6
    assert (net.fc_tile[1].weight.detach().cpu() - sd["fc_tile.1.weight"]).abs().max() > 0


def test_mil_epoch_single_process_equals_sample_plus_train_tile(cuda):
    from cellsegmentation_b200.inference import sample_indices, inference_tiles
    from cellsegmentation_b200.mil import mil_epoch, select_global
    bags, ds = _trainset()
    x = torch.from_numpy(otiles.unfold(list(bags[1:]), 20, 32))
    sd = omodel.calibrate_head(omodel.make_state_dict("resnet34", seed=3), x[::3], "resnet34")
    net = _model("resnet34", sd, cuda, "bf16")
    ds.setmode(1)
    loader = torch.utils.data.DataLoader(ds, batch_size=64, shuffle=False)
    probs = inference_tiles(loader, net, cuda)
    want_idx, want_pl = sample_indices(ds, probs, 1, 30)
    gidx, glab = select_global(ds, net, cuda, 1, 30)
    assert np.array_equal(gidx, want_idx) and np.array_equal(glab, want_pl)
    opt = torch.optim.SGD(filter(lambda p: p.requires_grad, net.parameters()), lr=1e-6)
    w0 = net.fc_tile[1].weight.detach().clone()
    b0 = net.fc_tile[1].bias.detach().clone()
    loss, pos, neg = mil_epoch(ds, net, cuda, torch.nn.CrossEntropyLoss(), opt, 1, 30, 0.5, 16, seed=3,
                               cache_features=False)
    assert np.isfinite(loss) and pos + neg == len(ds.train_data)
    w1 = net.fc_tile[1].weight.detach().clone()
    # N2: the cached-feature epoch is the same computation without the second encoder pass
    with torch.no_grad():
        net.fc_tile[1].weight.copy_(w0); net.fc_tile[1].bias.copy_(b0)
    loss_c, pos_c, neg_c = mil_epoch(ds, net, cuda, torch.nn.CrossEntropyLoss(), opt, 1, 30, 0.5, 16, seed=3,
                                     cache_features=True)
    assert (pos_c, neg_c) == (pos, neg)
    assert abs(loss_c - loss) <= 1e-6 * max(1.0, abs(loss))
    assert torch.allclose(net.fc_tile[1].weight.detach(), w1, rtol=0, atol=1e-9)


def test_distributed_mil_epoch_torchrun(cuda):
    """tests/dist_mil_epoch.py under torchrun on every visible GPU (needs >= 2): rank-identical global
    selection (sharded scoring + shard-aware top-k + NCCL all-gather) and identical fc_tile weights."""
    import subprocess
    import sys
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                        "--master-addr", "127.0.0.1", "--master-port", "29541",
                        os.path.join(here, "dist_mil_epoch.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_train_tile_matches_reference_golden_fp32(cuda):
    """The product's train_tile + LystoDataset (mode 3) against train/train.py:12-48 executed
    unmodified on the reference's own MILresnet34 / DataLoader (tests/golden/train.npz): two
    epochs of SGD, mean losses and the updated fc_tile."""
    from cellsegmentation_b200.dataset import LystoDataset
    from cellsegmentation_b200.train import train_tile
    g = golden("train.npz")
    bags = synth.make_bags(5, seed=41)
    ds = LystoDataset.from_arrays(list(bags), [9, 3, 0, 7, 0], 32, 20)
    x = torch.from_numpy(otiles.unfold(list(bags[1:]), 20, 32))
    sd = omodel.calibrate_head(omodel.make_state_dict("resnet34", seed=3), x[::3], "resnet34")
    net = _model("resnet34", sd, cuda, "fp32")
    np.random.seed(7)
    with contextlib.redirect_stdout(io.StringIO()):
        ds.make_train_data(g["idxs"], 0.5)
    td = ds.train_data
    assert np.array_equal(np.stack([td["bag"], td["x"], td["y"], td["label"]], 1).astype(np.int64), g["td_r05"])
    ds.setmode(3)
    bs, lr, wd, gamma = g["train_hparams"]
    loader = torch.utils.data.DataLoader(ds, batch_size=int(bs), shuffle=False)
    opt = torch.optim.SGD(filter(lambda p: p.requires_grad, net.parameters()), lr=float(lr), weight_decay=float(wd))
    losses = [train_tile(loader, e, 2, net, cuda, torch.nn.CrossEntropyLoss(), opt, None, float(gamma))
              for e in (1, 2)]
    assert np.allclose(losses, g["train_losses"], rtol=2e-5, atol=2e-5), (losses, g["train_losses"])
    assert np.abs(net.fc_tile[1].weight.detach().cpu().numpy() - g["train_fc_w"]).max() < 2e-5
    assert np.abs(net.fc_tile[1].bias.detach().cpu().numpy() - g["train_fc_b"]).max() < 2e-5


@pytest.mark.parametrize("arch", ["resnet34", "resnet18"])
def test_image_and_segment_modes_match_reference(cuda, arch):
    """N4: whole-image encoder + Stage-1 heads (mode "image") and the Stage-3 decoder (mode
    "segment") against MILResNet.forward of the reference (model/resnet.py:271-303) executed on
    the same 299 x 299 inputs and weights (tests/golden/image_seg.npz), fp32 CUDA-core path."""
    from cellsegmentation_b200.model.resnet import MILresnet18, MILresnet34
    g = golden("image_seg.npz")
    bags = synth.make_bags(3, seed=51)
    x = torch.from_numpy(np.stack([otiles.normalize_tile(b) for b in bags])).to(cuda)
    net = {"resnet34": MILresnet34, "resnet18": MILresnet18}[arch]()
    sd = omodel.make_state_dict(arch, seed=3)
    sd.update(omodel.make_image_seg_state(arch, seed=5))
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected and all("num_batches_tracked" in k for k in missing), (missing, unexpected)
    net.to(cuda).eval()
    net.setmode("image")
    cls, reg = net(x)
    tol = 1e-4 * float(np.abs(g[arch + "_cls"]).max())
    assert np.abs(cls.cpu().numpy() - g[arch + "_cls"]).max() < tol
    assert np.abs(reg.cpu().numpy() - g[arch + "_reg"]).max() < 1e-4 * max(1.0, float(np.abs(g[arch + "_reg"]).max()))
    net.setmode("segment")
    seg = net(x[:2])
    assert tuple(seg.shape) == (2, 2, 299, 299)
    amax = float(g[arch + "_seg_absmax"])
    assert np.abs(seg.cpu().numpy()[:, :, ::5, ::5] - g[arch + "_seg_sample"]).max() < 1e-4 * amax
    assert np.allclose(seg.double().sum(dim=(2, 3)).cpu().numpy(), g[arch + "_seg_sum"], rtol=1e-5)
    net.train()
    with pytest.raises(NotImplementedError):
        net(x[:1])


def test_inference_image_matches_reference(cuda):
    """inference_image (inference.py:46-95) over a LystoTestset in mode "image": ids, categories and
    counts (with and without cls_limit) equal the reference's on the same weights and images."""
    from cellsegmentation_b200.dataset import LystoTestset
    from cellsegmentation_b200.inference import inference_image
    from cellsegmentation_b200.model.resnet import MILresnet34
    g = golden("image_seg.npz")
    bags = synth.make_bags(3, seed=51)
    ts = LystoTestset.from_arrays(list(bags), 32, 5)
    ts.id = [7, 8, 9]
    ts.setmode("image")
    net = MILresnet34()
    sd = omodel.make_state_dict("resnet34", seed=3)
    sd.update(omodel.make_image_seg_state("resnet34", seed=5))
    net.load_state_dict(sd, strict=False)
    net.to(cuda)
    net.setmode("image")
    loader = torch.utils.data.DataLoader(ts, batch_size=2, shuffle=False)
    for lim in (False, True):
        ids, cats, counts = inference_image(loader, net, cuda, mode="test", cls_limit=lim, return_id=True)
        assert np.array_equal(ids, g["inf_ids"])
        assert np.array_equal(cats, g["inf_cats_%d" % lim])
        assert np.array_equal(counts, g["inf_counts_%d" % lim])
    # the per-item path of the dataset (what a foreign DataLoader would use) is the same transform
    i, img = ts[1]
    assert i == 8 and np.array_equal(img.numpy().view(np.uint32), otiles.normalize_tile(bags[1]).view(np.uint32))
