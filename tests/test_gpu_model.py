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


def test_hilo_residual_mode_subprocess(cuda):
    """CELLSEG_RESIDUAL=hilo (read when the library loads) keeps the hi/lo residual stream: same gate."""
    import os
    import subprocess
    import sys
    code = (
        "import sys, numpy as np, torch; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from test_gpu_model import _setup, _ops, BF16_TOL\n"
        "from oracle import model as omodel\n"
        "ops = _ops(); bags, x, sd = _setup('resnet34')\n"
        "want = omodel.forward_probs(sd, x, 'resnet34')\n"
        "clf = ops.TileClassifier('resnet34', omodel.fold_bn(sd, 'resnet34'), sd['fc_tile.1.weight'], sd['fc_tile.1.bias'])\n"
        "got = clf.forward_tiles(torch.from_numpy(bags).cuda(), 32, 20, precision='bf16', max_batch=256).cpu().numpy()\n"
        "d = float(np.abs(got - want).max()); print('hilo max|dp|', d); assert d < BF16_TOL\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CELLSEG_RESIDUAL="hilo")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "hilo max|dp|" in r.stdout


@pytest.mark.parametrize("env", [{"CELLSEG_YSUM": "0"}, {"CELLSEG_YSUM_PAIRS": "1"}, {"CELLSEG_CLUSTER": "1"},
                                 {"CELLSEG_STEM": "im2col"}, {"CELLSEG_RESIDUAL": "hilo", "CELLSEG_YSUM": "0"},
                                 {"CELLSEG_RESIDUAL": "hilo"}],
                         ids=lambda e: ",".join("%s=%s" % kv for kv in e.items()))
def test_alternative_kernel_paths_subprocess(cuda, env):
    """The switches are read when the library loads: run the conv-form and ResNet-34 parity tests in a
    child process for every alternative kernel path (halo layer 1, y-sum CTA pairs, single-CTA MMAs,
    im2col stem, hi/lo residual stream)."""
    import os
    import subprocess
    import sys
    here = os.path.abspath(__file__)
    r = subprocess.run([sys.executable, "-m", "pytest", here, "-x", "-q", "-p", "no:cacheprovider", "-k",
                        "conv_matches or (within_2e2 and resnet34) or tile16"],
                       env=dict(os.environ, **env), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
