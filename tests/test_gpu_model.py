"""GPU parity of the tile-classifier forward (fp32 and bf16/tcgen05) against the oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import golden
from oracle import model as omodel, synth, tiles as otiles

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4      # north_star: end-to-end tile probabilities within 1e-4 abs in fp32
BF16_TOL = 2e-2      # ... and 2e-2 abs in bf16


def _ops():
    from cellsegmentation_b200 import ops
    return ops


def _bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


@pytest.mark.parametrize("M,N,K,bn", [(128, 64, 64, 64), (256, 256, 128, 128), (384, 512, 1024, 256),
                                       (200, 128, 192, 64), (128 * 5 + 3, 256, 576, 256)])
def test_tcgen05_gemm_matches_fp32_matmul(cuda, M, N, K, bn):
    ops = _ops()
    g = torch.Generator().manual_seed(M + N + K)
    a = _bf16_round(torch.randn(M, K, generator=g))
    b = _bf16_round(torch.randn(N, K, generator=g) / K ** 0.5)
    bias = torch.randn(N, generator=g)
    want = a.double() @ b.double().t() + bias.double()
    got = ops.debug_gemm_bf16(a.to(torch.bfloat16).to(cuda), b.to(torch.bfloat16).to(cuda),
                              bias.to(cuda), bn).cpu().double()
    err = (got - want).abs().max().item()
    assert err < 1e-3, err


@pytest.mark.parametrize("H,Cin,Cout,k,stride,groups", [
    (8, 64, 64, 3, 1, 1), (4, 128, 128, 3, 1, 1), (8, 64, 128, 3, 2, 1), (2, 256, 256, 3, 1, 1),
    (4, 128, 256, 3, 2, 1), (2, 256, 512, 3, 2, 1), (1, 512, 512, 3, 1, 1),
    # Bottleneck / ResNeXt shapes: pointwise, strided pointwise, grouped 3x3 in every form
    (8, 64, 256, 1, 1, 1), (4, 512, 128, 1, 1, 1), (8, 256, 512, 1, 2, 1), (2, 1024, 2048, 1, 2, 1),
    (1, 2048, 512, 1, 1, 1), (8, 128, 128, 3, 1, 32), (8, 256, 256, 3, 2, 32), (4, 256, 256, 3, 1, 32),
    (4, 512, 512, 3, 2, 32), (2, 512, 512, 3, 1, 32), (2, 1024, 1024, 3, 2, 32), (1, 1024, 1024, 3, 1, 32),
    # ResNeXt-101 32x8d widths: more step-list variants than fit the kernel parameters
    (8, 256, 256, 3, 1, 32), (4, 1024, 1024, 3, 2, 32), (2, 2048, 2048, 3, 2, 32)])
def test_tcgen05_conv_matches_conv2d(cuda, H, Cin, Cout, k, stride, groups):
    ops = _ops()
    g = torch.Generator().manual_seed(H * 1000 + Cin + Cout + stride + 7 * groups + k)
    n = 256
    x = _bf16_round(torch.randn(n, H, H, Cin, generator=g))
    w = _bf16_round(torch.randn(Cout, Cin // groups, k, k, generator=g) / (k * (Cin // groups) ** 0.5))
    b = torch.randn(Cout, generator=g)
    want = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), stride=stride,
                    padding=1 if k == 3 else 0, groups=groups)
    want = want.permute(0, 2, 3, 1)
    got = ops.debug_conv_bf16(x.to(torch.bfloat16).to(cuda), w, b, stride, groups).cpu().double()
    err = (got - want).abs().max().item()
    assert err < 2e-3, err


# ---- the benchmarked regime: every persistent kernel loops many times per CTA ------------------
# bench.py runs forward batches of 75 776 instances (~256 work items per CTA in layer 1): ring
# and accumulator phases wrap dozens of times and odd layers walk their tiles backwards.  The
# small-n cases above are checked against F.conv2d; here the same inputs go through ONE large
# launch (>= 8 iterations per CTA, forwards and backwards) and must reproduce, bit for bit, the
# result of 128..256-instance launches of the same kernel.
@pytest.mark.parametrize("H,Cin,Cout,k,stride,groups,n", [
    (8, 64, 64, 3, 1, 1, 8192),          # layer 1: y-sum kernel, 27 M tiles per CTA
    (8, 64, 128, 3, 2, 1, 16384),        # layer-2 entry: shifted boxes over four parity maps
    (4, 128, 128, 3, 1, 1, 16384),       # layer 2: halo kernel, CTA pairs
    (4, 128, 256, 3, 2, 1, 32768),       # layer-3 entry: dense form
    (2, 256, 256, 3, 1, 1, 65536),       # layer 3: dense form, 4 N tiles
    (1, 512, 512, 3, 1, 1, 81920),       # layer 4: centre tap only
    (8, 64, 256, 1, 1, 1, 8192),         # Bottleneck pointwise
    (8, 256, 256, 3, 2, 32, 8192),       # ResNeXt grouped 3x3, per-tile step lists
])
@pytest.mark.parametrize("reverse", [False, True], ids=["fwd", "rev"])
def test_tcgen05_conv_many_iterations_per_cta(cuda, H, Cin, Cout, k, stride, groups, n, reverse):
    ops = _ops()
    g = torch.Generator().manual_seed(H * 1000 + Cin + Cout + stride + 7 * groups + k + 99)
    x = torch.randn(n, H, H, Cin, generator=g).to(torch.bfloat16).to(cuda)
    w = _bf16_round(torch.randn(Cout, Cin // groups, k, k, generator=g) / (k * (Cin // groups) ** 0.5))
    b = torch.randn(Cout, generator=g)
    got = ops.debug_conv_bf16(x, w, b, stride, groups, reverse=reverse)
    chunk = 256
    # spot-check chunks at the beginning, middle and end (first / last work items of many CTAs)
    for c0 in (0, chunk, n // 2 - chunk, n // 2 + 3 * chunk, n - 2 * chunk, n - chunk):
        ref = ops.debug_conv_bf16(x[c0:c0 + chunk].contiguous(), w, b, stride, groups)
        assert torch.equal(got[c0:c0 + chunk], ref), (c0, (got[c0:c0 + chunk] - ref).abs().max().item())
    # ... and the whole tensor against a second pass in the opposite direction
    other = ops.debug_conv_bf16(x, w, b, stride, groups, reverse=not reverse)
    assert torch.equal(got, other)
    # numerics of a sample against fp64 conv2d
    xs = x[n - 64:].float().cpu()
    want = F.conv2d(xs.permute(0, 3, 1, 2).double(), w.double(), b.double(), stride=stride,
                    padding=1 if k == 3 else 0, groups=groups).permute(0, 2, 3, 1)
    assert (got[n - 64:].cpu().double() - want).abs().max().item() < 2e-3


@pytest.mark.parametrize("n,reverse", [(128, False), (384, True), (16384 + 128, False), (16384 + 128, True)])
def test_tcgen05_basic_block_matches_two_convs(cuda, n, reverse):
    """A layer-1 BasicBlock through the planner (one fused launch by default: conv_block.cu) against
    the two convolutions run separately through the y-sum kernel with the glue in torch -- the same
    arithmetic in the same order, so bit-exact -- and a sample against fp64 (model/resnet.py:28-43)."""
    import os
    ops = _ops()
    g = torch.Generator().manual_seed(n + 17 * reverse)
    x = torch.randn(n, 8, 8, 64, generator=g).to(torch.bfloat16).to(cuda)
    w1 = _bf16_round(torch.randn(64, 64, 3, 3, generator=g) / 24.0)
    w2 = _bf16_round(torch.randn(64, 64, 3, 3, generator=g) / 24.0)
    b1, b2 = torch.randn(64, generator=g) * 0.3, torch.randn(64, generator=g) * 0.3
    got, launches = ops.debug_basic_block_bf16(x, w1, b1, w2, b2, reverse=reverse)
    # the fused kernel needs CTA pairs and the y-sum form of layer 1
    unfused = (os.environ.get("CELLSEG_BLOCK_FUSE") == "0" or os.environ.get("CELLSEG_YSUM_PAIRS") == "0" or
               os.environ.get("CELLSEG_YSUM") == "0" or os.environ.get("CELLSEG_CLUSTER") == "1" or
               os.environ.get("CELLSEG_DENSE_PO") == "64" or os.environ.get("CELLSEG_L1_SUB"))
    assert launches == (2 if unfused else 1)
    mid = F.relu(ops.debug_conv_bf16(x, w1, b1, 1)).to(torch.bfloat16)
    want = F.relu(ops.debug_conv_bf16(mid, w2, b2, 1) + x.float()).to(torch.bfloat16)
    assert torch.equal(got, want), (got.float() - want.float()).abs().max().item()
    xs = x[n - 32:].float().cpu().permute(0, 3, 1, 2).double()
    m64 = _bf16_round(F.relu(F.conv2d(xs, w1.double(), b1.double(), padding=1)).float()).double()
    y64 = F.relu(F.conv2d(m64, w2.double(), b2.double(), padding=1) + xs).permute(0, 2, 3, 1)
    err = (got[n - 32:].float().cpu().double() - y64).abs()
    assert (err <= 2.0 ** -7 * y64.abs() + 2e-3).all(), err.max().item()


@pytest.mark.parametrize("interval,n_bags,begin", [(20, 2, 0), (10, 3, 0), (7, 2, 37)])
def test_tcgen05_stem_matches_conv_pool(cuda, interval, n_bags, begin):
    """The tile-32 stem kernel alone (normalise, conv 7x7/2 + bias, ReLU, maxpool 3x3/2) against
    fp64 conv2d on the same bf16-rounded inputs and weights (model/resnet.py:236-239)."""
    ops = _ops()
    bags = synth.make_bags(n_bags + 1, seed=5)[1:]
    x = torch.from_numpy(otiles.unfold(list(bags), interval, 32))[begin:]
    g = torch.Generator().manual_seed(interval)
    w = _bf16_round(torch.randn(64, 3, 7, 7, generator=g) / 147 ** 0.5)
    b = torch.randn(64, generator=g) * 0.5
    images = torch.from_numpy(np.stack(bags)).to(cuda)
    got = ops.debug_stem_bf16(images, w, b, interval, inst_begin=begin).float().cpu()      # [n, 8, 8, 64]
    assert got.shape[0] == x.shape[0]
    conv = F.conv2d(_bf16_round(x).double(), w.double(), b.double(), stride=2, padding=3)
    want = F.max_pool2d(F.relu(_bf16_round(conv.float())), 3, 2, 1).permute(0, 2, 3, 1)
    diff = (got - want).abs()
    # fp32 accumulation order can move a sum across a bf16 rounding boundary: one ulp at most
    assert (diff <= 2.0 ** -7 * want.abs() + 1e-6).all(), diff.max().item()
    assert (diff == 0).float().mean().item() > 0.99


def _setup(arch, n_bags=2, interval=20, tile=32, seed=3):
    bags = synth.make_bags(n_bags + 1, seed=11)[1:]
    x = torch.from_numpy(otiles.unfold(list(bags), interval, tile))
    sd = omodel.make_state_dict(arch, seed=seed)
    sd = omodel.calibrate_head(sd, x[::3], arch)
    return bags, x, sd


ARCHS = ["resnet34", "resnet18", "resnet50", "resnext50_32x4d", "resnext101_32x8d"]


@pytest.mark.parametrize("arch", ARCHS)
def test_forward_fp32_within_1e4(cuda, arch):
    ops = _ops()
    bags, x, sd = _setup(arch)
    want = omodel.forward_probs(sd, x, arch)
    clf = ops.TileClassifier(arch, omodel.fold_bn(sd, arch), sd["fc_tile.1.weight"], sd["fc_tile.1.bias"])
    d_img = torch.from_numpy(bags).to(cuda)
    got = clf.forward_tiles(d_img, 32, 20, precision="fp32", max_batch=300).cpu().numpy()
    err = np.abs(got - want).max()
    assert err < FP32_TOL, err
    assert clf.last_launch_count > 0
    g = golden("model_%s.npz" % arch)          # the reference's own output on the same inputs
    assert np.abs(got - g["probs"]).max() < FP32_TOL
    # drop-in tensor entry: logits + pooled features
    logits, feat = clf.forward_tensor(x[:100].to(cuda), precision="fp32", max_batch=64, want_features=True)
    wl = omodel.forward_logits(sd, x[:100], arch)
    assert (logits.cpu() - wl).abs().max() < 5e-4
    with torch.no_grad():
        wf = omodel.pooled(omodel.forward_features(sd, x[:100], arch))
    assert torch.allclose(feat.cpu(), wf, rtol=1e-4, atol=1e-5 * float(wf.abs().max()) + 1e-3)
    clf.close()


@pytest.mark.parametrize("arch", ARCHS)
def test_forward_bf16_within_2e2(cuda, arch):
    ops = _ops()
    bags, x, sd = _setup(arch)
    want = omodel.forward_probs(sd, x, arch)
    clf = ops.TileClassifier(arch, omodel.fold_bn(sd, arch), sd["fc_tile.1.weight"], sd["fc_tile.1.bias"])
    d_img = torch.from_numpy(bags).to(cuda)
    got = clf.forward_tiles(d_img, 32, 20, precision="bf16", max_batch=256).cpu().numpy()
    diff = np.abs(got - want)
    print("bf16 %s: max|dp| %.4g mean %.4g frac>tol %.4g" % (arch, diff.max(), diff.mean(),
                                                          (diff > BF16_TOL).mean()))
    assert diff.max() < BF16_TOL, diff.max()
    # batches that are not a multiple of 128 and a non-zero instance offset
    got2 = clf.forward_tiles(d_img, 32, 20, inst_begin=37, inst_count=301, precision="bf16",
                             max_batch=128).cpu().numpy()
    assert np.abs(got2 - got[37:338]).max() < 1e-6
    logits = clf.forward_tensor(x[:130].to(cuda), precision="bf16", max_batch=256)
    wl = omodel.forward_logits(sd, x[:130], arch)
    p_got = torch.softmax(logits.cpu(), 1)[:, 1]
    assert (p_got - torch.softmax(wl, 1)[:, 1]).abs().max() < BF16_TOL
    clf.close()


@pytest.mark.parametrize("arch,interval,n_bags", [("resnet34", 5, 26), ("resnext50_32x4d", 3, 10)])
def test_forward_bf16_at_bench_scale(cuda, arch, interval, n_bags):
    """The benchmarked configuration: max_batch 75 776 (the default of bench.py and of the model), test-time interval
    (3025 / 8100 instances per bag), more instances than one batch with a ragged last batch that
    is not a multiple of 128.  (i) every probability within 2e-2 of the CPU oracle, (ii) the same
    call in 256-instance batches (the regime of the small tests) agrees to 1e-6."""
    ops = _ops()
    bags = synth.make_bags(n_bags, seed=31)
    x = torch.from_numpy(otiles.unfold(list(bags), interval, 32))
    n = x.shape[0]
    assert n > 75776 and (n - 75776) % 128 != 0
    sd = omodel.calibrate_head(omodel.make_state_dict(arch, seed=3), x[::97], arch)
    want = omodel.forward_probs(sd, x, arch, batch=4096)
    clf = ops.TileClassifier(arch, omodel.fold_bn(sd, arch), sd["fc_tile.1.weight"], sd["fc_tile.1.bias"])
    d_img = torch.from_numpy(bags).to(cuda)
    got = clf.forward_tiles(d_img, 32, interval, precision="bf16", max_batch=75776)
    launches_big = clf.last_launch_count
    diff = np.abs(got.cpu().numpy() - want)
    print("bf16 %s at scale: n %d max|dp| %.4g mean %.4g" % (arch, n, diff.max(), diff.mean()))
    assert diff.max() < BF16_TOL, diff.max()
    small = clf.forward_tiles(d_img, 32, interval, precision="bf16", max_batch=256)
    assert launches_big < clf.last_launch_count
    assert (got - small).abs().max().item() <= 1e-6
    # pooled features of the big batches (the MIL feature cache) against the small ones too
    _, f_big = clf.forward_tiles(d_img, 32, interval, inst_begin=5, inst_count=40001, precision="bf16",
                                 max_batch=37888, want_features=True)     # the round-1 batch size, two batches
    _, f_small = clf.forward_tiles(d_img, 32, interval, inst_begin=5, inst_count=40001, precision="bf16",
                                   max_batch=512, want_features=True)
    assert torch.equal(f_big, f_small)
    clf.close()


@pytest.mark.parametrize("arch", ["resnet34", "resnext50_32x4d"])
def test_forward_tile16(cuda, arch):
    ops = _ops()
    bags = synth.make_bags(2, seed=21)
    x = torch.from_numpy(otiles.unfold(list(bags), 20, 16))
    sd = omodel.calibrate_head(omodel.make_state_dict(arch, seed=5), x[::3], arch)
    want = omodel.forward_probs(sd, x, arch)
    clf = ops.TileClassifier(arch, omodel.fold_bn(sd, arch), sd["fc_tile.1.weight"], sd["fc_tile.1.bias"])
    d_img = torch.from_numpy(bags).to(cuda)
    got32 = clf.forward_tiles(d_img, 16, 20, precision="fp32", max_batch=512).cpu().numpy()
    assert np.abs(got32 - want).max() < FP32_TOL
    got16 = clf.forward_tiles(d_img, 16, 20, precision="bf16", max_batch=512).cpu().numpy()
    assert np.abs(got16 - want).max() < BF16_TOL
    clf.close()


ALT_PATHS = [
    # (environment, also run the many-iteration and at-scale cases)
    ({"CELLSEG_STEM": "win"}, True), ({"CELLSEG_BLOCK_FUSE": "0"}, True), ({"CELLSEG_EPI_RING": "1"}, True), ({"CELLSEG_YSUM": "0"}, False), ({"CELLSEG_YSUM_PAIRS": "0"}, True), ({"CELLSEG_CLUSTER": "1"}, True),
    ({"CELLSEG_YSUM_EPI": "8"}, False), ({"CELLSEG_YSUM_BOX": "0"}, True), ({"CELLSEG_DENSE_PO": "4"}, True),
    ({"CELLSEG_DENSE_PO": "4", "CELLSEG_HALO_DS": "0"}, False), ({"CELLSEG_DENSE_BN": "128"}, False),
    ({"CELLSEG_DENSE_PO": "64"}, False), ({"CELLSEG_L1_SUB": "4736"}, True),
    ({"CELLSEG_RESIDUAL": "hilo", "CELLSEG_YSUM": "0"}, False), ({"CELLSEG_RESIDUAL": "hilo"}, False),
]
# Bottleneck / ResNeXt switches: grouped convs back in the generic shifted-box kernel
ALT_PATHS_RX = [{"CELLSEG_GROUP_YSUM": "0"}, {"CELLSEG_DENSE_GROUP_PO": "4"}]


def test_alternative_kernel_paths_subprocess(cuda):
    """The switches are read when the library loads: run the conv-form and ResNet-34 parity tests in
    child processes for every alternative kernel path (window-form stem, unfused layer-1 blocks, four-set epilogue ring, halo layer 1, single-CTA y-sum MMAs, single-CTA
    MMAs everywhere, 8-warp y-sum epilogue, three-box y-sum, halo kernel for layer 2 with and without
    the fused shortcut, 128-wide dense tiles, dense 8x8 stage, L2-resident layer-1 sub-batches, hi/lo
    residual stream); the paths that change how a forward batch is walked also run the
    many-iteration and at-scale cases.  The children run concurrently (they share the GPU)."""
    import os
    import subprocess
    import sys
    here = os.path.abspath(__file__)
    light = "conv_matches or stem_matches or basic_block or (within_2e2 and resnet34) or tile16"
    heavy = light + " or many_iterations or (bench_scale and resnet34)"
    procs = []
    for env in ALT_PATHS_RX:
        procs.append((env, subprocess.Popen(
            [sys.executable, "-m", "pytest", here, "-x", "-q", "-p", "no:cacheprovider", "-k",
             "conv_matches or (within_2e2 and resnext)"],
            env=dict(os.environ, OMP_NUM_THREADS="4", **env), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for env, is_heavy in ALT_PATHS:
        procs.append((env, subprocess.Popen(
            [sys.executable, "-m", "pytest", here, "-x", "-q", "-p", "no:cacheprovider", "-k", heavy if is_heavy else light],
            env=dict(os.environ, OMP_NUM_THREADS="4", **env), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = []
    for env, pr in procs:
        try:
            out, _ = pr.communicate(timeout=1500)
        except subprocess.TimeoutExpired:
            pr.kill()
            out, _ = pr.communicate()
            out += "\n[timed out]"
        if pr.returncode != 0:
            failed.append("%s:\n%s" % (env, out[-2500:]))
    assert not failed, "\n\n".join(failed)
