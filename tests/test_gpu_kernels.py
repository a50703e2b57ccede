"""GPU parity tests: every kernel family through the C ABI against the CPU oracle."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import masks as omasks, select as oselect, synth, tiles as otiles

pytestmark = pytest.mark.gpu


def _ops():
    from cellsegmentation_b200 import ops
    return ops


# ----------------------------------------------------------------------------- K4b
@pytest.mark.parametrize("n_px", [1, 31, 32, 33, 1000, 96 * 96 * 3, 299 * 299 * 5 + 7])
def test_hsv_refine_bit_exact(cuda, n_px):
    ops = _ops()
    rng = np.random.default_rng(n_px)
    img = rng.integers(0, 256, (n_px, 3), dtype=np.uint8)
    img[rng.uniform(size=n_px) < 0.2] = rng.integers(168, 173, 3, dtype=np.uint8)   # straddle 170
    mask = (rng.uniform(size=n_px) < 0.5).astype(np.uint8) * rng.integers(1, 256, n_px).astype(np.uint8)
    want = omasks.hsv_refine(img, mask).astype(np.uint8)
    d_img, d_mask = torch.from_numpy(img).to(cuda), torch.from_numpy(mask).to(cuda)
    got = ops.hsv_refine(d_img, d_mask).cpu().numpy()
    assert int((got != want).sum()) == 0
    for thr in (0, 169, 255):
        got = ops.hsv_refine(d_img, d_mask, thr).cpu().numpy()
        assert np.array_equal(got, omasks.hsv_refine(img, mask, thr).astype(np.uint8))
    # in place (out aliases mask) and idempotence
    m2 = d_mask.clone()
    ops.hsv_refine(d_img, m2, out=m2)
    assert np.array_equal(m2.cpu().numpy(), want)
    ops.hsv_refine(d_img, m2, out=m2)
    assert np.array_equal(m2.cpu().numpy(), want)


def test_hsv_refine_unaligned_pointers(cuda):
    ops = _ops()
    rng = np.random.default_rng(5)
    n = 5000
    img = rng.integers(0, 256, (n + 1, 3), dtype=np.uint8)
    mask = rng.integers(0, 2, n + 3, dtype=np.uint8)
    d_img, d_mask = torch.from_numpy(img).to(cuda), torch.from_numpy(mask).to(cuda)
    got = ops.hsv_refine(d_img[1:], d_mask[3:]).cpu().numpy()
    assert np.array_equal(got, omasks.hsv_refine(img[1:], mask[3:]).astype(np.uint8))


def test_hsv_refine_golden_pre_cc(cuda):
    ops = _ops()
    g = golden("masks.npz")
    small = synth.make_bags(3, H=96, W=96, seed=31)
    got = ops.hsv_refine(torch.from_numpy(small).to(cuda), torch.from_numpy(g["raw"]).to(cuda))
    assert int((got.cpu().numpy() != g["pre_cc"]).sum()) == 0


def test_bgr2hsv_matches_opencv(cuda):
    import cv2
    ops = _ops()
    rng = np.random.default_rng(1)
    cols = rng.integers(0, 256, (1 << 20, 1, 3), dtype=np.uint8)
    grey = np.repeat(np.arange(256, dtype=np.uint8), 3).reshape(256, 1, 3)
    cols = np.concatenate([cols, grey])
    got = ops.bgr2hsv(torch.from_numpy(cols).to(cuda)).cpu().numpy()
    assert np.array_equal(got, cv2.cvtColor(cols, cv2.COLOR_BGR2HSV))


# ----------------------------------------------------------------------------- K1
@pytest.mark.parametrize("S,I", [(32, 20), (32, 5), (16, 5)])
def test_unfold_normalize_bit_exact(cuda, S, I):
    ops = _ops()
    bags = synth.make_bags(2, seed=3)
    T = len(otiles.get_tiles((299, 299, 3), I, S))
    begin, count = T - 40, 90                           # crosses the bag boundary
    want = otiles.unfold(list(bags), I, S, begin, count)
    got = ops.unfold_normalize(torch.from_numpy(bags).to(cuda), S, I, begin, count).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_unfold_golden_transform(cuda):
    ops = _ops()
    g = golden("transform.npz")
    bags = synth.make_bags(3, seed=11)
    got = ops.unfold_normalize(torch.from_numpy(bags[1:]).to(cuda), 32, 20).cpu().numpy()
    for j, i in enumerate(g["pick"]):
        assert np.array_equal(got[i].view(np.uint32), g["tiles"][j].view(np.uint32))


def test_gather_normalize(cuda):
    ops = _ops()
    bags = synth.make_bags(3, seed=4)
    rng = np.random.default_rng(0)
    n = 200
    bag = rng.integers(0, 3, n).astype(np.int32)
    x = rng.integers(0, 299 - 32 + 1, n).astype(np.int32)
    y = rng.integers(0, 299 - 32 + 1, n).astype(np.int32)
    got = ops.gather_normalize(torch.from_numpy(bags).to(cuda), 32, torch.from_numpy(bag).to(cuda),
                               torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda)).cpu().numpy()
    for j in range(n):
        want = otiles.normalize_tile(bags[bag[j]][x[j]:x[j] + 32, y[j]:y[j] + 32])
        assert np.array_equal(got[j].view(np.uint32), want.view(np.uint32))


# ----------------------------------------------------------------------------- K3
def _offsets(tileIDX, n_bags):
    return np.concatenate([[0], np.cumsum(np.bincount(tileIDX, minlength=n_bags))]).astype(np.int64)


@pytest.mark.parametrize("case", ["toy", "ties", "k0", "tpp3", "nan", "single", "wrapbig", "dense"])
def test_select_topk_golden(cuda, case):
    ops = _ops()
    g = golden("select.npz")
    tid, lab, p = g[case + "_tileIDX"], g[case + "_labels"], g[case + "_probs"]
    tpp, tk = (int(v) for v in g[case + "_params"])
    want = g[case + "_idx"]
    off = _offsets(tid, len(lab))
    maxT = int(np.diff(off).max())
    idx, pl, sel_off = ops.select_topk(torch.from_numpy(p).to(cuda), torch.from_numpy(lab).to(cuda),
                                       len(lab), maxT, tpp, tk,
                                       seg_offsets=torch.from_numpy(off).to(cuda))
    assert np.array_equal(idx.cpu().numpy().astype(np.int64), want)
    assert np.array_equal(pl.cpu().numpy(), oselect.pseudo_labels(tid, lab, want))
    kept = np.bincount(tid[want], minlength=len(lab)) if len(want) else np.zeros(len(lab), int)
    assert np.array_equal(np.diff(sel_off.cpu().numpy()), kept)
    # lexsort order itself
    order = ops.lexsort_segments(torch.from_numpy(p).to(cuda), len(lab), maxT,
                                 seg_offsets=torch.from_numpy(off).to(cuda)).cpu().numpy()
    assert np.array_equal(order, np.lexsort((p, tid)))


@pytest.mark.parametrize("T,n_bags", [(1, 5), (2, 3), (33, 7), (225, 16), (1021, 5), (3025, 9), (3069, 4), (3364, 6),
                                       (4093, 3), (4094, 3), (4097, 3), (8100, 2)])
def test_select_topk_random_uniform(cuda, T, n_bags):
    ops = _ops()
    rng = np.random.default_rng(T * 31 + n_bags)
    p = rng.uniform(0, 1, T * n_bags).astype(np.float32)
    p[rng.uniform(size=p.size) < 0.05] = np.float32(1.0)          # saturated ties
    p[rng.uniform(size=p.size) < 0.05] = np.float32(0.0)
    p[rng.uniform(size=p.size) < 0.01] = np.float32(-0.0)
    lab = rng.integers(0, max(2, T // 2), n_bags).astype(np.int32)
    lab[rng.uniform(size=n_bags) < 0.3] = 0
    lab[-1] = T + 5                                               # k >= T on the wrapping last bag
    tid = np.repeat(np.arange(n_bags), T)
    for tpp, tk in [(1, 30), (2, 3), (1, 0)]:
        want = oselect.sample_indices(tid, lab, p, tpp, tk)
        idx, pl, _ = ops.select_topk(torch.from_numpy(p).to(cuda), torch.from_numpy(lab).to(cuda),
                                     n_bags, T, tpp, tk)
        assert np.array_equal(idx.cpu().numpy().astype(np.int64), want)
        assert np.array_equal(pl.cpu().numpy(), oselect.pseudo_labels(tid, lab, want))


@pytest.mark.parametrize("T,n_bags,misalign", [(3025, 2500, 0), (225, 6000, 0), (3025, 700, 3), (784, 1900, 1)])
def test_select_topk_many_bags_ring_wraps(cuda, T, n_bags, misalign):
    """Enough bags per SM for the streaming ring to wrap several times; LYSTO-like labels (30 % zeros,
    geometric counts up to 300: n <= 32, 32 < n <= 128 and declined n > 128 all occur), ties, and a
    probability array whose start / end are not 16-byte aligned (clipped copy spans)."""
    ops = _ops()
    rng = np.random.default_rng(T + n_bags)
    p = rng.uniform(0, 1, T * n_bags).astype(np.float32)
    p[rng.uniform(size=p.size) < 0.02] = np.float32(1.0)
    p[:T] = np.float32(0.5)                                        # one all-tied bag -> heavy-tie decline
    lab = np.minimum(rng.geometric(1.0 / 8.0, n_bags), 300).astype(np.int32)
    lab[rng.uniform(size=n_bags) < 0.3] = 0
    lab[5] = 200
    tid = np.repeat(np.arange(n_bags), T)
    want = oselect.sample_indices(tid, lab, p, 1, 30)
    buf = torch.zeros(p.size + 8, dtype=torch.float32, device=cuda)
    view = buf[misalign:misalign + p.size]
    view.copy_(torch.from_numpy(p))
    idx, pl, _ = ops.select_topk(view, torch.from_numpy(lab).to(cuda), n_bags, T, 1, 30)
    assert np.array_equal(idx.cpu().numpy().astype(np.int64), want)
    assert np.array_equal(pl.cpu().numpy(), oselect.pseudo_labels(tid, lab, want))


def test_select_topk_more_bags_than_the_lookback_scan_takes(cuda):
    """40 000 small bags: 40 offset blocks > the 31 the look-back scan takes, so the last-block
    ("ticket") scan runs, in two shared-memory chunks (24 576 counts each), in front of the two-vector
    register kernel; checked against the oracle incl. the offsets themselves."""
    ops = _ops()
    T, n_bags = 21, 40000
    rng = np.random.default_rng(5)
    p = rng.uniform(0, 1, T * n_bags).astype(np.float32)
    lab = np.minimum(rng.geometric(1.0 / 3.0, n_bags), 40).astype(np.int32)
    lab[rng.uniform(size=n_bags) < 0.3] = 0
    tid = np.repeat(np.arange(n_bags), T)
    want = oselect.sample_indices(tid, lab, p, 1, 4)
    idx, pl, off = ops.select_topk(torch.from_numpy(p).to(cuda), torch.from_numpy(lab).to(cuda), n_bags, T, 1, 4)
    assert np.array_equal(idx.cpu().numpy().astype(np.int64), want)
    assert np.array_equal(pl.cpu().numpy(), oselect.pseudo_labels(tid, lab, want))
    kept = np.bincount(tid[want], minlength=n_bags)
    assert np.array_equal(np.diff(off.cpu().numpy()), kept)


@pytest.mark.parametrize("T,sizes", [(225, [2, 1, 1, 3, 1]), (64, [1, 1, 1, 1]), (300, [3, 0, 2, 1])])
def test_select_topk_shards_equal_global(cuda, T, sizes):
    """Bags partitioned over ranks: every shard evaluates the wrap-around predicate at its GLOBAL
    position (cs_select_topk_shard), so one-bag shards, empty shards and counts beyond the shard
    size (k >= shard tiles) reproduce the single-process selection."""
    ops = _ops()
    n_bags = sum(sizes)
    rng = np.random.default_rng(T + n_bags)
    p = rng.uniform(0, 1, T * n_bags).astype(np.float32)
    lab = rng.integers(0, 12, n_bags).astype(np.int32)
    lab[0] = 0
    lab[-1] = 2 * T + 7                      # k wraps around the global array
    lab[1 % n_bags] = T + 3                  # k larger than a one-bag shard
    tid = np.repeat(np.arange(n_bags), T)
    for tpp, tk in [(1, 30), (3, 5)]:
        want = oselect.sample_indices(tid, lab, p, tpp, tk)
        got_idx, got_pl, b0 = [], [], 0
        for nb in sizes:
            if nb > 0:
                sl = slice(b0 * T, (b0 + nb) * T)
                idx, pl, _ = ops.select_topk(torch.from_numpy(p[sl]).to(cuda),
                                             torch.from_numpy(lab[b0:b0 + nb]).to(cuda), nb, T, tpp, tk,
                                             global_offset=b0 * T, global_total=n_bags * T)
                got_idx.append(idx.cpu().numpy().astype(np.int64) + b0 * T)
                got_pl.append(pl.cpu().numpy())
            b0 += nb
        assert np.array_equal(np.concatenate(got_idx), want)
        assert np.array_equal(np.concatenate(got_pl), oselect.pseudo_labels(tid, lab, want))


@pytest.mark.parametrize("mode", ["staged", "0", "persist", "warp", "ticket", "occ16", "cta64", "sortplain", "recount", "col64"])
def test_select_topk_other_paths_subprocess(cuda, mode):
    """CELLSEG_SELECT_FAST is read when the library loads: =staged routes every bag through the
    shared-memory fast path of round 1 (still used for bags longer than 4093 instances), =0 through
    the exact bitonic kernel alone; CELLSEG_SELECT_PERSIST=1 takes the persistent form of the
    CTA-per-bag register kernel, CELLSEG_SELECT_WARP=1 the warp-per-bag kernel,
    CELLSEG_SELECT_OFFSETS=ticket | recount the last-block scan / the recount kernel instead of the
    look-back offsets scan,
    CELLSEG_SELECT_OCC=12 | 16 the register kernel held to 40 / 32 registers, CELLSEG_SELECT_CTA=64
    its two-warp form (CELLSEG_SELECT_OCC64: resident CTAs per SM), CELLSEG_SELECT_SORT_PDL=0 a plain
    launch of the exact clean-up pass, CELLSEG_SELECT_COL32=0 the 64-column threshold for every kept count."""
    import os
    import subprocess
    import sys
    here = os.path.abspath(__file__)
    r = subprocess.run([sys.executable, "-m", "pytest", here, "-x", "-q", "-p", "no:cacheprovider", "-k",
                        "select_topk and not subprocess"],
                       env=dict(os.environ, **({"persist": {"CELLSEG_SELECT_PERSIST": "1"},
                                                "warp": {"CELLSEG_SELECT_WARP": "1"},
                                                "occ16": {"CELLSEG_SELECT_OCC": "16"},
                                                "sortplain": {"CELLSEG_SELECT_SORT_PDL": "0"},
                                                "cta64": {"CELLSEG_SELECT_CTA": "64"},

                                                "ticket": {"CELLSEG_SELECT_OFFSETS": "ticket"},
                                                "recount": {"CELLSEG_SELECT_OFFSETS": "recount"},
                                                "col64": {"CELLSEG_SELECT_COL32": "0"}}.get(
                                                    mode, {"CELLSEG_SELECT_FAST": mode}))),
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_select_topk_ragged_with_empty_bags(cuda):
    ops = _ops()
    rng = np.random.default_rng(9)
    sizes = np.array([0, 17, 0, 300, 1, 64, 0])
    tid = np.repeat(np.arange(len(sizes)), sizes)
    p = rng.uniform(0, 1, tid.size).astype(np.float32)
    lab = np.array([3, 2, 5, 0, 1, 100, 2], np.int32)
    want = oselect.sample_indices(tid, lab, p, 1, 30)
    off = _offsets(tid, len(sizes))
    idx, _, _ = ops.select_topk(torch.from_numpy(p).to(cuda), torch.from_numpy(lab).to(cuda),
                                len(sizes), int(sizes.max()), 1, 30,
                                seg_offsets=torch.from_numpy(off).to(cuda))
    assert np.array_equal(idx.cpu().numpy().astype(np.int64), want)


def test_rank_threshold_golden(cuda):
    ops = _ops()
    g = golden("rank.npz")
    tid, p = g["tileIDX"], g["probs"]
    off = _offsets(tid, 5)
    idx, kp, _ = ops.rank_threshold(torch.from_numpy(p).to(cuda), 5, 225, float(g["threshold"]),
                                    seg_offsets=torch.from_numpy(off).to(cuda))
    idx = idx.cpu().numpy()
    grid = np.array(otiles.get_tiles((299, 299, 3), 20, 32), np.int32)
    assert np.array_equal(grid[idx % 225], g["tiles"])
    assert np.array_equal(kp.cpu().numpy().view(np.uint32), g["kept_probs"].view(np.uint32))
    assert np.array_equal(tid[idx], g["groups"])


@pytest.mark.parametrize("thr", [0.95, 0.5, 0.999, -1.0])
def test_rank_threshold_random_fast_and_exact_paths(cuda, thr):
    """rank(): bags keeping <= 512 entries take the compact + rank-by-counting kernel, the others the
    exact per-bag sort; both must give lexsort order restricted to prob > thr (ties, NaN, -0.0)."""
    ops = _ops()
    rng = np.random.default_rng(17)
    n_bags, T = 300, 3025
    p = rng.uniform(0, 1, n_bags * T).astype(np.float32)
    p[rng.uniform(size=p.size) < 0.01] = np.float32(0.97)          # ties above the usual threshold
    p[rng.uniform(size=p.size) < 0.001] = np.float32(np.nan)
    p[5 * T:6 * T] = np.float32(0.99)                              # one bag keeps everything
    p[7 * T:8 * T] = np.float32(0.1)                               # one keeps nothing (thr > 0.1)
    tid = np.repeat(np.arange(n_bags), T)
    order = np.lexsort((p, tid))
    want = order[p[order] > np.float32(thr)]
    idx, kp, off = ops.rank_threshold(torch.from_numpy(p).to(cuda), n_bags, T, thr)
    assert np.array_equal(idx.cpu().numpy().astype(np.int64), want)
    assert np.array_equal(kp.cpu().numpy().view(np.uint32), p[want].view(np.uint32))
    assert np.array_equal(np.diff(off.cpu().numpy()), np.bincount(tid[want], minlength=n_bags))


# ----------------------------------------------------------------------------- K4a
def test_paint_mask_and_heatmap_golden(cuda):
    import cv2
    ops = _ops()
    g = golden("masks.npz")
    small = synth.make_bags(3, H=96, W=96, seed=31)
    keep = torch.from_numpy(g["kept"].astype(np.int32)).to(cuda)
    kp = torch.from_numpy(g["probs"][g["kept"]]).to(cuda)
    raw = ops.paint_mask(keep, 3, 96, 96, 16, 5)
    assert np.array_equal(raw.cpu().numpy(), g["raw"])
    heat = ops.paint_heatmap(keep, kp, 3, 96, 96, 16, 5)
    gray = ops.heatmap_to_gray(heat).cpu().numpy()
    for i in range(3):
        cm = cv2.applyColorMap(gray[i], cv2.COLORMAP_JET)
        img = cv2.addWeighted(small[i], 0.5, cm, 0.5, 0)
        assert np.array_equal(np.uint8(img), g["heat_imgs"][i])


@pytest.mark.parametrize("H,W,S,I,n_bags,bag_base", [(96, 96, 16, 5, 3, 0), (299, 299, 32, 5, 5, 0), (100, 131, 32, 7, 4, 1),
                                                       (64, 64, 16, 8, 2, 0), (299, 299, 32, 20, 6, 2), (40, 33, 32, 3, 2, 0)])
def test_paint_heatmap_gather_equals_scatter(cuda, H, W, S, I, n_bags, bag_base):
    """cs_paint_heatmap_gather (every pixel written once: max over the kept covering tiles) against
    the scatter form cs_paint_heatmap (atomicMax on zeroed maps), bit-exact: stride not dividing the
    span, S a multiple of I, single-position axes, a bag offset, duplicates, NaN / negative
    probabilities, bags without kept tiles, and an output buffer full of garbage."""
    ops = _ops()
    rng = np.random.default_rng(H * 7 + W + I)
    T = len(otiles.get_tiles((H, W, 3), I, S))
    own = n_bags - bag_base                                   # bags the instance indices can address
    n_sel = max(1, (own * T) // 3)
    sel = rng.integers(0, own * T, n_sel).astype(np.int32)
    if own > 1:
        sel = sel[sel // T != own - 1]                        # one bag keeps nothing
    sel = np.concatenate([sel, sel[:5]])                      # duplicates with other probabilities
    p = rng.uniform(0, 1, sel.size).astype(np.float32)
    p[::17] = np.float32(-0.25)
    p[5::23] = np.float32("nan")
    d_sel, d_p = torch.from_numpy(sel).to(cuda), torch.from_numpy(p).to(cuda)
    want = ops.paint_heatmap(d_sel, d_p, n_bags, H, W, S, I, bag_base=bag_base)
    out = torch.full((n_bags, H, W), 7.5, dtype=torch.float32, device=cuda)
    got = ops.paint_heatmap_gather(d_sel, d_p, n_bags, H, W, S, I, bag_base=bag_base, out=out)
    assert torch.equal(got, want)
    assert float(want.max()) > 0


def test_heatmap_to_gray_all_thresholded_probs(cuda):
    """uint8(255*p) in float64 for every float32 p in [0.9, 1] plus a sweep of [0,1]."""
    ops = _ops()
    lo = np.float32(0.9).view(np.uint32)
    hi = np.float32(1.0).view(np.uint32)
    p = np.arange(lo, hi + 1, dtype=np.uint32).view(np.float32)
    p = np.concatenate([p, np.linspace(0, 1, 100001, dtype=np.float32)])
    got = ops.heatmap_to_gray(torch.from_numpy(p).to(cuda)).cpu().numpy()
    assert np.array_equal(got, omasks.heat_to_gray(p.astype(np.float64)))


# ----------------------------------------------------------------------------- N1 connected components
@pytest.mark.parametrize("H,W,density,seed", [(299, 299, 0.5, 0), (299, 299, 0.62, 1), (96, 96, 0.55, 2),
                                               (64, 48, 0.4, 3), (299, 299, 0.9, 4), (17, 301, 0.5, 5)])
def test_remove_small_regions_bit_exact(cuda, H, W, density, seed):
    ops = _ops()
    rng = np.random.default_rng(seed)
    n = 5
    # blobs + salt noise: components of all sizes around the 400 / 120 thresholds
    coarse = rng.uniform(size=(n, H // 7 + 2, W // 7 + 2)) < density
    m = np.kron(coarse, np.ones((7, 7), bool))[:, :H, :W]
    m ^= rng.uniform(size=m.shape) < 0.03
    m[0] = 0
    m[1] = 1
    for mo, ha in [(400, 120), (30, 9), (0, 120), (400, 0), (1, 1)]:
        want = np.stack([omasks.remove_small_regions(x, mo, ha) for x in m]).astype(np.uint8)
        got = ops.remove_small_regions(torch.from_numpy(m.astype(np.uint8) * 3).to(cuda), mo, ha)
        assert int((got.cpu().numpy() != want).sum()) == 0, (mo, ha)


def test_remove_small_regions_noise_batch(cuda):
    """Per-pixel noise at every density (thousands of runs per image, one giant component on one side),
    more images than resident CTAs, and shapes that take the pixel-level fallback (wider than the
    bit-image limit / more runs than the run table)."""
    ops = _ops()
    rng = np.random.default_rng(11)
    dens = np.linspace(0.05, 0.95, 160)
    m = rng.uniform(size=(160, 299, 299)) < dens[:, None, None]
    want = np.stack([omasks.remove_small_regions(x, 400, 120) for x in m]).astype(np.uint8)
    got = ops.remove_small_regions(torch.from_numpy(m.astype(np.uint8)).to(cuda), 400, 120)
    assert int((got.cpu().numpy() != want).sum()) == 0
    for shape in [(40, 400), (330, 331)]:
        m2 = rng.uniform(size=(3,) + shape) < 0.5
        want2 = np.stack([omasks.remove_small_regions(x, 30, 9) for x in m2]).astype(np.uint8)
        got2 = ops.remove_small_regions(torch.from_numpy(m2.astype(np.uint8)).to(cuda), 30, 9)
        assert int((got2.cpu().numpy() != want2).sum()) == 0, shape


def test_heatmap_blend_matches_cv2(cuda):
    """N3: applyColorMap(JET) + addWeighted(.5,.5) fused on the GPU == cv2, every byte pair covered."""
    import cv2
    ops = _ops()
    lut = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(256, 1), cv2.COLORMAP_JET).reshape(256, 3)
    rng = np.random.default_rng(5)
    # (a) all 256 x 256 (image byte, gray level) pairs: heat chosen so that 255 - uint8(255 h) = level
    lv = np.repeat(np.arange(256), 256).reshape(1, 256, 256)
    heat = ((255 - lv) / 255.0).astype(np.float32)
    heat = np.where((255 - (255.0 * heat.astype(np.float64)).astype(np.uint8)) == lv, heat,
                    np.nextafter(heat, np.float32(2))).astype(np.float32)
    assert np.array_equal(255 - (255.0 * heat.astype(np.float64)).astype(np.uint8), lv)
    img = np.tile(np.arange(256, dtype=np.uint8), 256).reshape(1, 256, 256, 1).repeat(3, axis=3)
    # (b) random maps with an odd pixel count (tail path) and unaligned-looking shapes
    heat2 = rng.random((3, 37, 41), dtype=np.float32)
    heat2[rng.random(heat2.shape) < 0.5] = 0.0
    img2 = rng.integers(0, 256, (3, 37, 41, 3), dtype=np.uint8)
    for h, im in ((heat, img), (heat2, img2)):
        got = ops.heatmap_blend(torch.from_numpy(h).to(cuda), torch.from_numpy(np.ascontiguousarray(im)).to(cuda),
                                torch.from_numpy(np.ascontiguousarray(lut)).to(cuda)).cpu().numpy()
        for i in range(h.shape[0]):
            gray = (255 - np.uint8(255 * h[i].astype(np.float64)))
            want = cv2.addWeighted(np.ascontiguousarray(im[i]), 0.5, cv2.applyColorMap(gray, cv2.COLORMAP_JET), 0.5, 0)
            assert np.array_equal(got[i], want)


def test_remove_small_regions_known_answers(cuda):
    """Hand-derivable masks (sizes 399/400/401, holes 119/120/121, diagonal contact, border pocket,
    objects-before-holes order): the GPU clean-up against answers that do not come from the scipy
    restatement (skimage itself cannot be executed in this environment)."""
    from test_oracle_golden import _kat_masks
    ops = _ops()
    cases = _kat_masks()
    m = np.stack([c[1] for c in cases]).astype(np.uint8)
    want = np.stack([c[2] for c in cases]).astype(np.uint8)
    got = ops.remove_small_regions(torch.from_numpy(m).to(cuda), 400, 120).cpu().numpy()
    for i, c in enumerate(cases):
        assert np.array_equal(got[i], want[i]), c[0]
    # the same masks embedded in 299 x 299 images (the production shape)
    big = np.zeros((len(cases), 299, 299), np.uint8)
    big[:, 100:196, 150:246] = m
    wbig = np.zeros_like(big)
    for i, c in enumerate(cases):
        wbig[i] = omasks.remove_small_regions(big[i] != 0, 400, 120)
    got = ops.remove_small_regions(torch.from_numpy(big).to(cuda), 400, 120).cpu().numpy()
    assert np.array_equal(got, wbig)
