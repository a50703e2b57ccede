"""Generates tests/golden/*.npz by executing the UNMODIFIED reference (/root/reference).

Run once in the build container:  python tests/golden/make_golden.py
The reference has no tests or golden vectors of its own; these fixtures pin the
oracle (and, through it, the CUDA path) to the reference's own code on seeded inputs.  Inputs
are regenerated in the tests from oracle/synth.py (numpy PCG64 streams are version-stable),
only outputs are stored.
"""
import ast
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import model as omodel, ref_shim, synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def save(name, **arrs):
    np.savez_compressed(os.path.join(OUT, name), **arrs)
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in arrs.items()})


def ref_dataset(ref, bags, labels, tile, interval):
    ds = ref.dataset.LystoDataset(tile_size=tile, interval=interval, kfold=None, _ensemble_init=True)
    for i, (img, lab) in enumerate(zip(bags, labels)):
        ds.add_data("colon_%d" % i, img, int(lab), tileidx=i)   # tileidx=0 adds no tiles (:142)
    return ds


def extract_nested_rank():
    """test_tile.rank is a closure inside test_tile(); compile its source verbatim."""
    src = open(os.path.join(ref_shim.REF_ROOT, "test_tile.py")).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "rank":
            seg = ast.get_source_segment(src, node)
            import textwrap
            return textwrap.dedent(seg)
    raise RuntimeError("rank() not found in test_tile.py")


def gen_model(ref, ds, arch):
    """Model forward through the reference classes + inference_tiles."""
    if arch.startswith("resnext"):
        net = getattr(ref_shim.load_reference_module("resnext"), "MIL" + arch)(num_classes=2)
    else:
        net = getattr(ref_shim.load_reference_module("resnet"), "MIL" + arch)()
    sd = omodel.make_state_dict(arch, seed=3)
    ds.setmode(1)
    calib = torch.stack([ds[i][0] for i in range(0, len(ds), 3)])
    sd = omodel.calibrate_head(sd, calib, arch)
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected and all(not k.startswith(net.encoder_prefix + net.tile_module_prefix)
                                  for k in missing), (missing, unexpected)
    net.setmode("tile")
    net.eval()
    loader = torch.utils.data.DataLoader(ds, batch_size=64, shuffle=False, num_workers=0)
    import contextlib
    with contextlib.redirect_stderr(io.StringIO()):
        probs = ref.inference.inference_tiles(loader, net, torch.device("cpu"), mode="train")
    fwd = net.resnext_forward if arch.startswith("resnext") else net.resnet_forward
    with torch.no_grad():
        x16 = torch.stack([ds[i][0] for i in range(16)])
        logits = net(x16).numpy()
        x4, x3, x2, x1 = fwd(x16, True)
    save("model_%s.npz" % arch, probs=probs.astype(np.float32), logits16=logits,
         x1_sum=x1.double().sum(dim=(1, 2, 3)).numpy(), x2_sum=x2.double().sum(dim=(1, 2, 3)).numpy(),
         x3_sum=x3.double().sum(dim=(1, 2, 3)).numpy(), x4=x4.numpy().reshape(16, -1))


def td_rows(train_data):
    """object[M,3] rows (bag, (x, y), label) -> int64 [M,4]."""
    return np.array([[int(r[0]), int(r[1][0]), int(r[1][1]), int(r[2])] for r in train_data], np.int64).reshape(-1, 4)


def gen_train(ref):
    """make_train_data (dataset/dataset.py:166-201) and train_tile (train/train.py:12-48) executed
    unmodified: the reference's own LystoDataset, MILresnet34, DataLoader and training loop on CPU."""
    import contextlib
    out = {}
    bags = synth.make_bags(5, seed=41)
    labels = [9, 3, 0, 7, 0]
    ds = ref_dataset(ref, bags, labels, 32, 20)
    n = len(ds.tileIDX)
    idxs = np.arange(3, n, 5)
    out["idxs"] = idxs.astype(np.int64)
    for name, ratio, seed in (("r05", 0.5, 7), ("r20", 2.0, 8), ("rnone", None, 9), ("r01", 0.1, 10)):
        np.random.seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            pos, neg = ds.make_train_data(list(idxs), ratio)
        out["td_" + name] = td_rows(ds.train_data)
        out["pn_" + name] = np.array([pos, neg, seed], np.int64)
        out["ratio_" + name] = np.array(np.nan if ratio is None else ratio)
    # all-positive / all-negative selections (one class is pruned to nothing or kept whole)
    pos_only = np.array([i for i in range(n) if labels[ds.tileIDX[i]] != 0][::7], np.int64)
    np.random.seed(11)
    with contextlib.redirect_stdout(io.StringIO()):
        pos, neg = ds.make_train_data(list(pos_only), 0.5)
    out["idxs_posonly"], out["td_posonly"], out["pn_posonly"] = pos_only, td_rows(ds.train_data), np.array([pos, neg, 11])

    # train_tile on the r05 train_data, shuffle=False, plain SGD: loss + updated fc_tile
    np.random.seed(7)
    with contextlib.redirect_stdout(io.StringIO()):
        ds.make_train_data(list(idxs), 0.5)
    net = ref_shim.load_reference_module("resnet").MILresnet34()
    sd = omodel.make_state_dict("resnet34", seed=3)
    ds.setmode(1)
    calib = torch.stack([ds[i][0] for i in range(0, len(ds), 3)])
    sd = omodel.calibrate_head(sd, calib, "resnet34")
    net.load_state_dict(sd, strict=False)
    net.setmode("tile")
    ds.setmode(3)
    loader = torch.utils.data.DataLoader(ds, batch_size=16, shuffle=False, num_workers=0)
    opt = torch.optim.SGD(filter(lambda p: p.requires_grad, net.parameters()), lr=2e-5, weight_decay=1e-4)
    crit = torch.nn.CrossEntropyLoss()
    losses = []
    with contextlib.redirect_stderr(io.StringIO()):
        for epoch in (1, 2):
            losses.append(ref.train.train_tile(loader, epoch, 2, net, torch.device("cpu"), crit, opt, None, 1.0))
    trained = [n_ for n_, p in net.named_parameters() if p.requires_grad]
    # (the decoder's upconv* also stay requires_grad in tile mode, model/resnet.py:315-319, but are
    # not part of the tile graph: no gradient ever reaches them)
    assert trained[:2] == ["fc_tile.1.weight", "fc_tile.1.bias"] and \
        all(t.startswith(("upconv", "seg_out_conv")) for t in trained[2:]), trained
    assert all(p.grad is None for n_, p in net.named_parameters() if n_.startswith(("upconv", "seg_out_conv")))
    out["train_losses"] = np.array(losses, np.float64)
    out["train_fc_w"] = net.fc_tile[1].weight.detach().numpy().copy()
    out["train_fc_b"] = net.fc_tile[1].bias.detach().numpy().copy()
    out["train_hparams"] = np.array([16, 2e-5, 1e-4, 1.0])          # batch, lr, weight decay, gamma
    enc_same = all(torch.equal(net.state_dict()[k], sd[k]) for k in sd if not k.startswith("fc_tile") and
                   k in net.state_dict() and not k.endswith("num_batches_tracked"))
    assert enc_same, "train_tile changed encoder weights or BN statistics"
    save("train.npz", **out)


def gen_image_seg(ref):
    """N4: MILResNet.forward in modes "image" and "segment" (model/resnet.py:271-303) on whole
    299 x 299 images, through the reference class + inference_image (inference.py:46-95)."""
    import contextlib
    from oracle import tiles as otiles
    out = {}
    bags = synth.make_bags(3, seed=51)
    x = torch.from_numpy(np.stack([otiles.normalize_tile(b) for b in bags]))
    for arch in ("resnet34", "resnet18"):
        net = getattr(ref_shim.load_reference_module("resnet"), "MIL" + arch)()
        sd = omodel.make_state_dict(arch, seed=3)
        sd.update(omodel.make_image_seg_state(arch, seed=5))
        missing, unexpected = net.load_state_dict(sd, strict=False)
        assert not unexpected and all("num_batches_tracked" in k for k in missing), (missing, unexpected)
        net.eval()
        with torch.no_grad():
            net.setmode("image")
            cls, reg = net(x)
            net.setmode("segment")
            seg = net(x[:2])
        out[arch + "_cls"], out[arch + "_reg"] = cls.numpy(), reg.numpy()
        out[arch + "_seg_sample"] = seg.numpy()[:, :, ::5, ::5].copy()
        out[arch + "_seg_sum"] = seg.double().sum(dim=(2, 3)).numpy()
        out[arch + "_seg_absmax"] = np.array(float(seg.abs().max()))
        if arch == "resnet34":      # inference_image over a (ids, image) loader, with and without cls_limit
            net.setmode("image")
            loader = [(np.array([7, 8]), x[:2]), (np.array([9]), x[2:])]
            for lim in (False, True):
                with contextlib.redirect_stderr(io.StringIO()):
                    ids, cats, counts = ref.inference.inference_image(loader, net, torch.device("cpu"), mode="test",
                                                                      cls_limit=lim, return_id=True)
                out["inf_ids"], out["inf_cats_%d" % lim], out["inf_counts_%d" % lim] = ids, cats, counts
    save("image_seg.npz", **out)


def main():
    ref = ref_shim.import_reference()
    if sys.argv[1:] == ["train"]:
        return gen_train(ref)
    if sys.argv[1:] == ["image_seg"]:
        return gen_image_seg(ref)
    if len(sys.argv) > 1:          # python make_golden.py resnet50 ...: only those model fixtures
        ds = ref_dataset(ref, synth.make_bags(3, seed=11), [4, 0, 9], 32, 20)
        for arch in sys.argv[1:]:
            gen_model(ref, ds, arch)
        return

    # 1. tile grids -----------------------------------------------------------------
    grids = {}
    for (I, S) in [(5, 32), (20, 32), (10, 32), (3, 32), (2, 32), (5, 16), (7, 16)]:
        t = ref.dataset.get_tiles(np.zeros((299, 299, 3), np.uint8), I, S)
        grids["g_%d_%d" % (I, S)] = np.array(t, np.int32)
    t = ref.dataset.get_tiles(np.zeros((64, 48, 3), np.uint8), 9, 16)
    grids["g_9_16_64x48"] = np.array(t, np.int32)
    save("tiles.npz", **grids)

    # 2. per-tile transform through LystoDataset.__getitem__ (mode 1) -----------------
    bags = synth.make_bags(3, seed=11)
    labels = [4, 0, 9]
    ds = ref_dataset(ref, bags, labels, 32, 20)
    ds.setmode(1)
    pick = [0, 1, 14, 15, 224, 225, 300, 449]       # dataset indices (bag 0 owns no tiles)
    tiles_t = np.stack([ds[i][0].numpy() for i in pick])
    save("transform.npz", pick=np.array(pick), tiles=tiles_t, tileIDX=np.array(ds.tileIDX, np.int32),
         n=np.array(len(ds)))

    # 3. model forward through the reference classes + inference_tiles ----------------
    for arch in ("resnet34", "resnet18", "resnet50", "resnext50_32x4d", "resnext101_32x8d"):
        gen_model(ref, ds, arch)

    # 4. sample(): capture the idxs handed to make_train_data -------------------------
    cases = {}

    def run_sample(name, n_bags, labels, probs, tiles_per_pos, topk_neg, interval=20, first_has_tiles=False):
        b = synth.make_bags(1, seed=1)
        dsx = ref.dataset.LystoDataset(tile_size=32, interval=interval, kfold=None, _ensemble_init=True)
        for i in range(n_bags):
            # tileidx must be truthy to own tiles; emulate LystoTestset-like sets with tileidx=i+1
            if first_has_tiles:
                dsx.add_data("x_%d" % i, b[0], int(labels[i]), tileidx=i + 1)
            else:
                dsx.add_data("x_%d" % i, b[0], int(labels[i]), tileidx=i)
        if first_has_tiles:
            dsx.labels = [0] + dsx.labels        # tileIDX i+1 indexes labels[i+1]
        dsx.setmode(1)
        got = {}
        dsx.make_train_data = lambda idxs, r: (got.setdefault("idxs", list(idxs)), (0, 0))[1]
        p = probs(len(dsx))
        with contextlib.redirect_stdout(io.StringIO()):
            ref.inference.sample(dsx, p, tiles_per_pos, topk_neg, 0.5)
        cases[name + "_idx"] = np.array(got["idxs"], np.int64)
        cases[name + "_probs"] = p
        cases[name + "_labels"] = np.array(dsx.labels, np.int32)
        cases[name + "_tileIDX"] = np.array(dsx.tileIDX, np.int32)
        cases[name + "_params"] = np.array([tiles_per_pos, topk_neg], np.int32)

    import contextlib
    toy = [9, 3, 0, 7, 1, 300]                 # bag 0 owns no tiles; kept (3,30,7,1,225)
    run_sample("toy", 6, toy, lambda n: synth.make_probs(n, seed=5), 1, 30)
    run_sample("ties", 6, toy, lambda n: synth.make_probs(n, seed=6, ties=True), 1, 30)
    run_sample("k0", 4, [0, 2, 0, 5], lambda n: synth.make_probs(n, seed=7), 0, 0)
    run_sample("tpp3", 5, [1, 2, 0, 80, 4], lambda n: synth.make_probs(n, seed=8), 3, 10)

    def with_nan(n):
        p = synth.make_probs(n, seed=9, ties=True)
        p[::37] = np.nan
        p[5::41] = 1.0
        p[3::43] = 0.0
        return p
    run_sample("nan", 4, [3, 5, 0, 2], with_nan, 2, 7)
    run_sample("single", 2, [5, 4], lambda n: synth.make_probs(n, seed=10), 1, 30)   # one bag with tiles
    run_sample("wrapbig", 3, [0, 2, 600], lambda n: synth.make_probs(n, seed=12), 1, 400)
    run_sample("dense", 3, [0, 12, 0], lambda n: synth.make_probs(n, seed=13), 1, 30, interval=5)
    save("select.npz", **cases)

    # 5. rank() — nested in test_tile.test_tile, compiled verbatim --------------------
    import types
    ns = {"np": np, "args": types.SimpleNamespace(threshold=0.95)}
    exec(extract_nested_rank(), ns)
    bags5 = synth.make_bags(1, seed=2)
    dsr = ref_dataset(ref, [bags5[0]] * 4, [1, 2, 3, 4], 32, 20)
    pr = synth.make_probs(len(dsr.tileIDX), seed=21)
    pr[::9] = np.float32(0.97)
    pr[4::53] = np.nan
    tiles_k, probs_k, groups_k = ns["rank"](dsr, pr)
    save("rank.npz", probs=pr, tileIDX=np.array(dsr.tileIDX, np.int32), tiles=np.array(tiles_k, np.int32),
         kept_probs=np.array(probs_k, np.float32), groups=np.array(groups_k, np.int32),
         threshold=np.array(0.95, np.float32))

    # 6. evaluate_tile ----------------------------------------------------------------
    dse = ref_dataset(ref, [bags5[0]] * 5, [2, 3, 0, 6, 1], 32, 20)
    pe = synth.make_probs(len(dse.tileIDX), seed=22)
    err = ref.evaluate.evaluate_tile(dse, pe, 2, 0.9)
    save("evaluate.npz", probs=pe, tileIDX=np.array(dse.tileIDX, np.int32),
         labels=np.array(dse.labels, np.int32), out=np.array(err, np.float64),
         params=np.array([2, 0.9], np.float64))

    # 7. heatmap / generate_masks / preprocess_masks on small images --------------------
    small = synth.make_bags(3, H=96, W=96, seed=31)
    fake = types.SimpleNamespace(images=list(small), image_size=np.array([96, 96]), tile_size=16)
    grid = np.array(ref.dataset.get_tiles(small[0], 5, 16), np.int32)
    T = len(grid)
    rngp = np.random.default_rng(77)
    ph = rngp.uniform(0.9, 1.0, 3 * T).astype(np.float32)
    groups_all = np.repeat(np.arange(3), T)
    order = np.lexsort((ph, groups_all))
    keep = order[ph[order] > 0.985]
    k_tiles, k_probs, k_groups = grid[keep % T], ph[keep], groups_all[keep]
    ref.captured["imsave"].clear()
    import tempfile
    with tempfile.TemporaryDirectory() as td, contextlib.redirect_stderr(io.StringIO()):
        csvf = open(os.path.join(td, "h.csv"), "w", newline="")
        ref.utils.heatmap(fake, k_tiles, k_probs, k_groups, csvf, td)
        csvf.close()
        csv_text = open(os.path.join(td, "h.csv")).read()
        heat_imgs = np.stack([a for _, a in ref.captured["imsave"]])
        ref.captured["imsave"].clear()
        with contextlib.redirect_stdout(io.StringIO()):
            raw = ref.utils.generate_masks(fake, k_tiles, k_groups, preprocess=False, save_masks=False,
                                           output_path=td).copy()
            full = ref.utils.generate_masks(fake, k_tiles, k_groups, preprocess=True, save_masks=False,
                                            output_path=td).copy()
            import utils.image_processing as ip
            keep_fn = ip.remove_small_regions
            ip.remove_small_regions = lambda m, **k: m          # capture lines 117-120 only
            pre_cc = ref.utils.generate_masks(fake, k_tiles, k_groups, preprocess=True, save_masks=False,
                                              output_path=td).copy()
            ip.remove_small_regions = keep_fn
    save("masks.npz", kept=keep.astype(np.int64), probs=ph, heat_imgs=heat_imgs,
         csv=np.frombuffer(csv_text.encode(), np.uint8), raw=raw, pre_cc=pre_cc, full=full)

    # 8. make_train_data + train_tile, executed unmodified ---------------------------------
    gen_train(ref)

    # 9. N4: image / segment modes on whole images -----------------------------------------
    gen_image_seg(ref)


if __name__ == "__main__":
    main()
