"""Pins the CPU oracle to vectors produced by the reference's own code
(tests/golden/make_golden.py executed in the build container)."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import masks as omasks, model as omodel, select as oselect, synth, tiles as otiles


def test_get_tiles_matches_reference():
    g = golden("tiles.npz")
    for (I, S) in [(5, 32), (20, 32), (10, 32), (3, 32), (2, 32), (5, 16), (7, 16)]:
        got = np.array(otiles.get_tiles((299, 299, 3), I, S), np.int32)
        assert np.array_equal(got, g["g_%d_%d" % (I, S)])
    assert np.array_equal(np.array(otiles.get_tiles((64, 48, 3), 9, 16), np.int32), g["g_9_16_64x48"])
    # known-answer counts from SURVEY 8c
    for (I, S), T in {(5, 32): 3025, (20, 32): 225, (10, 32): 784, (3, 32): 8100, (2, 32): 18225,
                      (5, 16): 3364}.items():
        assert len(otiles.get_tiles((299, 299, 3), I, S)) == T


def _transform_dataset():
    bags = synth.make_bags(3, seed=11)
    x = otiles.unfold(list(bags[1:]), 20, 32)     # bag 0 owns no tiles in LystoDataset
    return bags, x


def test_transform_matches_reference_bit_exact():
    g = golden("transform.npz")
    _, x = _transform_dataset()
    assert x.shape[0] == int(g["n"])
    for j, i in enumerate(g["pick"]):
        assert np.array_equal(x[i].view(np.uint32), g["tiles"][j].view(np.uint32))
    assert g["tileIDX"].min() == 1            # first bag has no tiles


@pytest.mark.parametrize("arch", ["resnet34", "resnet18", "resnet50", "resnext50_32x4d", "resnext101_32x8d"])
def test_model_forward_matches_reference(arch):
    g = golden("model_%s.npz" % arch)
    _, x = _transform_dataset()
    xt = torch.from_numpy(x)
    sd = omodel.make_state_dict(arch, seed=3)
    sd = omodel.calibrate_head(sd, xt[::3], arch)
    logits = omodel.forward_logits(sd, xt[:16], arch).numpy()
    assert np.abs(logits - g["logits16"]).max() < 2e-4
    probs = omodel.forward_probs(sd, xt, arch, batch=64)
    assert np.abs(probs - g["probs"]).max() < 1e-5
    assert 0.05 < probs.std() < 0.45          # calibrated head: not saturated
    with torch.no_grad():
        x4, x3, x2, x1 = omodel.forward_features(sd, xt[:16], arch, return_intermediate=True)
    assert np.allclose(x4.numpy().reshape(16, -1), g["x4"], rtol=1e-4, atol=1e-4)
    for t, k in ((x1, "x1_sum"), (x2, "x2_sum"), (x3, "x3_sum")):
        assert np.allclose(t.double().sum(dim=(1, 2, 3)).numpy(), g[k], rtol=1e-5)


def test_bn_folding_is_fp32_exact_enough():
    _, x = _transform_dataset()
    xt = torch.from_numpy(x[:32])
    sd = omodel.make_state_dict("resnet34", seed=3)
    convs = omodel.fold_bn(sd, "resnet34")
    assert len(convs) == 36
    w0, b0 = convs[0]
    import torch.nn.functional as F
    ref = F.relu(omodel._bn(sd, F.conv2d(xt, sd["conv1.weight"], stride=2, padding=3), "bn1"))
    got = F.relu(F.conv2d(xt, w0, b0, stride=2, padding=3))
    assert (ref - got).abs().max() < 1e-4


SELECT_CASES = ["toy", "ties", "k0", "tpp3", "nan", "single", "wrapbig", "dense"]


@pytest.mark.parametrize("case", SELECT_CASES)
def test_sample_matches_reference(case):
    g = golden("select.npz")
    tid, lab, p = g[case + "_tileIDX"], g[case + "_labels"], g[case + "_probs"]
    tpp, tk = (int(v) for v in g[case + "_params"])
    want = g[case + "_idx"]
    assert np.array_equal(oselect.sample_indices(tid, lab, p, tpp, tk), want)
    if len(tid) <= 1200:
        assert np.array_equal(oselect.sample_indices_loop(tid, lab, p, tpp, tk), want)


def test_sample_known_answer_counts():
    g = golden("select.npz")
    kept = np.bincount(g["toy_tileIDX"][g["toy_idx"]], minlength=6)
    assert list(kept) == [0, 3, 30, 7, 1, 225]
    assert len(g["single_idx"]) == 0 and len(g["k0_idx"]) == 0


def test_rank_matches_reference():
    g = golden("rank.npz")
    grid = otiles.get_tiles((299, 299, 3), 20, 32)
    tid = g["tileIDX"]
    tiles_grid = [grid[i % 225] for i in range(len(tid))]
    t, p, gr, _ = oselect.rank(tid, tiles_grid, g["probs"], float(g["threshold"]))
    assert np.array_equal(np.array(t, np.int32), g["tiles"])
    assert np.array_equal(p.view(np.uint32), g["kept_probs"].view(np.uint32))
    assert np.array_equal(gr, g["groups"])


def test_evaluate_tile_matches_reference():
    g = golden("evaluate.npz")
    out = oselect.evaluate_tile(g["tileIDX"], g["labels"], g["probs"], int(g["params"][0]),
                                float(g["params"][1]))
    assert np.allclose(out, g["out"], rtol=0, atol=0)


def _mask_inputs():
    g = golden("masks.npz")
    small = synth.make_bags(3, H=96, W=96, seed=31)
    grid = np.array(otiles.get_tiles((96, 96, 3), 5, 16), np.int32)
    T = len(grid)
    keep = g["kept"]
    return g, small, grid[keep % T], g["probs"][keep], (keep // T)


def test_mask_painting_and_refinement_match_reference():
    g, small, k_tiles, k_probs, k_groups = _mask_inputs()
    raw = omasks.paint_masks(3, (96, 96), 16, k_tiles, k_groups)
    assert np.array_equal(raw, g["raw"])
    pre = np.stack([omasks.hsv_refine(small[i], raw[i]) for i in range(3)]).astype(np.uint8)
    assert np.array_equal(pre, g["pre_cc"])
    pre_cv = np.stack([omasks.hsv_refine_cv2(small[i], raw[i]) for i in range(3)]).astype(np.uint8)
    assert np.array_equal(pre_cv, g["pre_cc"])
    full = np.stack([omasks.preprocess_masks(small[i], raw[i]) for i in range(3)]).astype(np.uint8)
    assert np.array_equal(full, g["full"])
    assert 0 < pre.sum() < raw.sum()


def test_heatmap_matches_reference():
    import cv2
    g, small, k_tiles, k_probs, k_groups = _mask_inputs()
    heat = omasks.paint_heatmaps(3, (96, 96), 16, k_tiles, k_probs, k_groups)
    for i in range(3):
        cm = cv2.applyColorMap(omasks.heat_to_gray(heat[i]), cv2.COLORMAP_JET)
        img = cv2.addWeighted(small[i], 0.5, cm, 0.5, 0)
        assert np.array_equal(np.uint8(img), g["heat_imgs"][i])
    # painting in ascending-prob order == per-pixel max over kept covering tiles
    mx = np.zeros_like(heat)
    for t, p, gi in zip(k_tiles, k_probs, k_groups):
        sl = mx[gi][t[0]:t[0] + 16, t[1]:t[1] + 16]
        np.maximum(sl, p, out=sl)
    assert np.array_equal(mx, heat)


def test_bgr2hsv_restatement_matches_opencv():
    import cv2
    rng = np.random.default_rng(0)
    cols = rng.integers(0, 256, (1 << 18, 1, 3), dtype=np.uint8)
    edge = np.array([[[v, v, v]] for v in range(256)] + [[[255, 0, 0]], [[0, 255, 0]], [[0, 0, 255]],
                    [[170, 171, 169]], [[0, 0, 0]]], np.uint8)
    cols = np.concatenate([cols, edge])
    want = cv2.cvtColor(cols, cv2.COLOR_BGR2HSV)
    assert np.array_equal(omasks.bgr2hsv_u8(cols), want)
    assert np.array_equal(want[..., 2], cols.max(-1))            # V == max(B,G,R)


def test_make_train_data_shuffle_is_index_shuffle():
    """np.random.shuffle of the object rows draws the same permutation as shuffling indices,
    which is what the product does on the host."""
    grid = otiles.get_tiles((299, 299, 3), 20, 32)
    tid = np.repeat(np.arange(1, 5), 225)
    tiles_grid = [grid[i % 225] for i in range(len(tid))]
    labels = [9, 3, 0, 7, 0]
    idxs = np.arange(0, 900, 7)
    np.random.seed(123)
    td, pos, neg = oselect.make_train_data(tid, tiles_grid, labels, idxs, 0.5)
    np.random.seed(123)
    perm = np.arange(len(idxs))
    np.random.shuffle(perm)
    lab = (np.asarray(labels)[tid[idxs]] != 0).astype(int)
    # before pruning the shuffled order must be idxs[perm]
    np.random.seed(123)
    td2, _, _ = oselect.make_train_data(tid, tiles_grid, labels, idxs, None)
    assert [r[0] for r in td2] == list(tid[idxs][perm])
    assert [tuple(r[1]) for r in td2] == [tuple(tiles_grid[i]) for i in idxs[perm]]
    assert pos + neg == len(td) and pos == int(neg * 0.5) or neg == int(pos / 0.5)
    assert lab.sum() >= pos


# ---- make_train_data / train_tile: pinned to the reference's own methods (train.npz) -------------
def _train_golden_dataset():
    bags = synth.make_bags(5, seed=41)
    labels = [9, 3, 0, 7, 0]
    grid = otiles.get_tiles((299, 299, 3), 20, 32)
    tid = np.repeat(np.arange(1, 5), 225)            # bag 0 owns no tiles
    tiles_grid = [grid[i % 225] for i in range(len(tid))]
    return bags, labels, tid, tiles_grid


@pytest.mark.parametrize("case", ["r05", "r20", "rnone", "r01", "posonly"])
def test_make_train_data_matches_reference_method(case):
    """dataset/dataset.py:166-201 executed unmodified (under the numpy it was written for) vs the
    oracle restatement AND the product's LystoDataset.make_train_data, same np.random seed."""
    g = golden("train.npz")
    _, labels, tid, tiles_grid = _train_golden_dataset()
    idxs = g["idxs_posonly"] if case == "posonly" else g["idxs"]
    pos_w, neg_w, seed = (int(v) for v in g["pn_" + case])
    ratio = 0.5 if case == "posonly" else (None if np.isnan(g["ratio_" + case]) else float(g["ratio_" + case]))
    np.random.seed(seed)
    td, pos, neg = oselect.make_train_data(tid, tiles_grid, labels, idxs, ratio)
    rows = np.array([[int(r[0]), int(r[1][0]), int(r[1][1]), int(r[2])] for r in td], np.int64).reshape(-1, 4)
    assert (pos, neg) == (pos_w, neg_w)
    assert np.array_equal(rows, g["td_" + case])
    from cellsegmentation_b200.dataset import LystoDataset
    ds = LystoDataset.from_arrays([np.zeros((299, 299, 3), np.uint8)] * 5, labels, 32, 20)
    np.random.seed(seed)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        pos_p, neg_p = ds.make_train_data(idxs, ratio)
    got = ds.train_data
    assert (pos_p, neg_p) == (pos_w, neg_w)
    assert np.array_equal(np.stack([got["bag"], got["x"], got["y"], got["label"]], 1).astype(np.int64).reshape(-1, 4),
                          g["td_" + case])


def test_train_tile_oracle_matches_reference_loop():
    """train/train.py:12-48 executed unmodified (reference MILresnet34, DataLoader, SGD) vs the
    oracle restatement: two epochs' mean losses and the updated fc_tile."""
    from oracle import train as otrain
    g = golden("train.npz")
    bags, labels, tid, tiles_grid = _train_golden_dataset()
    x = torch.from_numpy(otiles.unfold(list(bags[1:]), 20, 32))
    sd = omodel.calibrate_head(omodel.make_state_dict("resnet34", seed=3), x[::3], "resnet34")
    td = g["td_r05"]
    tiles = torch.stack([torch.from_numpy(otiles.normalize_tile(bags[b][r:r + 32, c:c + 32])) for b, r, c, _ in td])
    y = torch.from_numpy(td[:, 3].copy())
    bs, lr, wd, gamma = g["train_hparams"]
    losses = []
    for _ in range(2):
        loss, w, b = otrain.train_tile_epoch(sd, tiles, y, int(bs), lr=float(lr), gamma=float(gamma),
                                             weight_decay=float(wd))
        sd = dict(sd)
        sd["fc_tile.1.weight"], sd["fc_tile.1.bias"] = w, b
        losses.append(loss)
    assert np.allclose(losses, g["train_losses"], rtol=1e-5, atol=2e-6), (losses, g["train_losses"])
    assert np.abs(w.numpy() - g["train_fc_w"]).max() < 2e-6
    assert np.abs(b.numpy() - g["train_fc_b"]).max() < 2e-6


# ---- remove_small_regions: known-answer masks (skimage itself cannot be run here) ----------------
def _kat_masks():
    """Masks whose clean-up is derivable by hand from skimage 0.19's definition: objects with
    size < min_size are removed, holes with area < area_threshold are filled, connectivity 1
    (4-neighbours; diagonal contact does not join), background touching the border is one
    big component (never a small hole unless it is small)."""
    H = W = 96
    cases = []
    # objects of 399 / 400 / 401 pixels: only the first disappears (min_size 400)
    m = np.zeros((H, W), bool)
    m[2:21, 2:23] = True           # 19 x 21 = 399
    m[30:50, 2:22] = True          # 20 x 20 = 400
    m[60:80, 2:22] = True; m[80, 2] = True     # 401
    w = m.copy(); w[2:21, 2:23] = False
    cases.append(("sizes_399_400_401", m, w))
    # holes of 119 / 120 / 121 pixels inside one big object: the 119-hole is filled (threshold 120)
    m = np.ones((H, W), bool)
    m[5:12, 5:22] = False          # 7 x 17 = 119
    m[20:30, 5:17] = False         # 10 x 12 = 120
    m[40:51, 5:16] = False         # 11 x 11 = 121
    w = m.copy(); w[5:12, 5:22] = True
    cases.append(("holes_119_120_121", m, w))
    # two 15 x 15 blobs touching only diagonally: 225 + 225 pixels, NOT joined under connectivity 1
    m = np.zeros((H, W), bool)
    m[10:25, 10:25] = True; m[25:40, 25:40] = True
    cases.append(("diagonal_blobs_removed", m, np.zeros((H, W), bool)))
    # the same blobs joined by one edge-adjacent pixel: one 451-pixel object, kept
    m2 = m.copy(); m2[24, 25] = True
    cases.append(("edge_joined_blobs_kept", m2, m2.copy()))
    # a small background pocket on the border is a hole like any other (area 30 < 120): filled;
    # a diagonal-only leak does not connect it to the outside background
    m = np.ones((H, W), bool)
    m[0:5, 0:6] = False            # 30-pixel pocket in the corner
    m[40:96, 60:96] = False        # large background region (2016): stays
    w = m.copy(); w[0:5, 0:6] = True
    cases.append(("border_pocket_filled", m, w))
    # object removal happens BEFORE hole filling: a 380-pixel ring (20x20 minus a 20-px hole) is removed
    # as an object (380 < 400) even though filling its hole first would have made it 400
    m = np.zeros((H, W), bool)
    m[10:30, 10:30] = True; m[15:19, 15:20] = False     # 400 - 20 = 380
    cases.append(("order_objects_then_holes", m, np.zeros((H, W), bool)))
    return cases


@pytest.mark.parametrize("idx", range(6))
def test_remove_small_regions_known_answers_oracle(idx):
    name, m, want = _kat_masks()[idx]
    got = omasks.remove_small_regions(m, 400, 120)
    assert np.array_equal(got, want), name
