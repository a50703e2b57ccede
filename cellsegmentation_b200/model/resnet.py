"""MILResNet drop-in (mirror of model/resnet.py of the reference) whose tile-mode forward runs
on the sm_100a kernels.

Same module tree and state_dict keys as the reference for the encoder and the tile head
(conv1, bn1, layer{1..4}.{b}.{conv1,bn1,conv2,bn2,[conv3,bn3,]downsample.{0,1}}, fc_tile.1.*), the same
prefix tuples and setmode() contract (model/resnet.py:83-106, 308-333), so reference
checkpoints load with the same `load_state_dict(..., strict=False)` calls the scripts make.

Tile mode: the encoder is frozen (setmode("tile") sets requires_grad False, :315-319) and is
evaluated with running BN statistics both under model.eval() (inference.py:12) and under
freeze_bn=True (train/train.py:33, model/resnet.py:254-258), so the forward is
    features = CUDA encoder(x)  [no grad]      ->  logits = fc_tile(features)  [autograd]
Image and segment modes (Stage 1 / Stage 3 networks) are out of scope and raise.
"""
import torch
import torch.nn as nn

from .. import ops

__all__ = ["MILresnet18", "MILresnet34", "MILresnet50", "MILResNet", "BasicBlock", "Bottleneck"]

BN_EPS = 1e-5


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, groups=1, base_width=64):
        super().__init__()
        if groups != 1 or base_width != 64:
            raise ValueError("BasicBlock only supports groups=1 and base_width=64")   # model/resnext.py:35-36
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride


class Bottleneck(nn.Module):
    """model/resnet.py:46-78 (groups=1, base_width=64) and model/resnext.py:67-113 (grouped 3x3,
    stride on conv2)."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, groups=1, base_width=64):
        super().__init__()
        width = int(planes * (base_width / 64.)) * groups
        self.conv1 = nn.Conv2d(inplanes, width, kernel_size=1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        self.conv2 = nn.Conv2d(width, width, kernel_size=3, stride=stride, padding=1, groups=groups, bias=False)
        self.bn2 = nn.BatchNorm2d(width)
        self.conv3 = nn.Conv2d(width, planes * self.expansion, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * self.expansion)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride


def _fold(conv, bn):
    """Eval-mode BN folded into the conv in fp32: W' = W*s, b' = beta - mean*s, s = gamma/sqrt(var+eps)."""
    w = conv.weight.detach().float().cpu()
    s = bn.weight.detach().float().cpu() / torch.sqrt(bn.running_var.detach().float().cpu() + bn.eps)
    b = bn.bias.detach().float().cpu() - bn.running_mean.detach().float().cpu() * s
    return (w * s[:, None, None, None]).contiguous(), b.contiguous()


class MILResNet(nn.Module):

    def __init__(self, encoder, block, layers, num_classes=1000, expansion=1, groups=1, width_per_group=64):
        super().__init__()
        if encoder not in ops._capi.CS_ARCH:
            raise NotImplementedError("encoder %r has no sm_100a kernels (available: %s)"
                                      % (encoder, sorted(ops._capi.CS_ARCH)))
        self.encoder_name = encoder
        self.groups = groups
        self.base_width = width_per_group
        self.mode = None
        self.encoder_prefix = ("conv1", "bn1", "relu", "layer1", "layer2", "layer3", "layer4")
        self.image_module_prefix = ("fc_image_cls", "fc_image_reg")
        self.tile_module_prefix = ("fc_tile",)
        self.seg_module_prefix = ("upconv", "seg_out_conv")
        self.precision = "bf16"        # "fp32" selects the CUDA-core parity path
        self.max_batch = 37888

        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=2)
        self.avgpool_tile = nn.AdaptiveAvgPool2d((1, 1))
        self.maxpool_tile = nn.AdaptiveMaxPool2d((1, 1))
        self.fc_tile = nn.Sequential(nn.Flatten(), nn.Linear(512 * block.expansion, num_classes))
        for m in self.modules():                       # model/resnet.py:171-178
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        self._clf = None
        self._clf_key = None
        self._fc_key = None

    def _make_layer(self, block, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                nn.Conv2d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False),
                nn.BatchNorm2d(planes * block.expansion))
        layers = [block(self.inplanes, planes, stride, downsample, self.groups, self.base_width)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes, groups=self.groups, base_width=self.base_width))
        return nn.Sequential(*layers)

    # ---- requires_grad groups (model/resnet.py:196-232, 308-333) ----------------------------
    def set_encoder_grads(self, requires_grad):
        for m in (self.conv1, self.bn1, self.layer1, self.layer2, self.layer3, self.layer4):
            m.requires_grad_(requires_grad)

    def set_tile_module_grads(self, requires_grad):
        self.fc_tile.requires_grad_(requires_grad)

    def setmode(self, mode):
        if mode == "tile":
            self.set_encoder_grads(False)
            self.set_tile_module_grads(True)
        elif mode in ("image", "segment"):
            self.set_encoder_grads(mode == "image")
            self.set_tile_module_grads(False)
        else:
            raise Exception("Invalid mode: {}.".format(mode))
        self.mode = mode

    # ---- device classifier ------------------------------------------------------------------
    def folded_convs(self):
        """[(W', b')] in the order cs_model_create expects: stem; per block conv1, conv2, [conv3],
        [downsample]."""
        convs = [_fold(self.conv1, self.bn1)]
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            for blk in layer:
                convs.append(_fold(blk.conv1, blk.bn1))
                convs.append(_fold(blk.conv2, blk.bn2))
                if hasattr(blk, "conv3"):
                    convs.append(_fold(blk.conv3, blk.bn3))
                if blk.downsample is not None:
                    convs.append(_fold(blk.downsample[0], blk.downsample[1]))
        return convs

    def _encoder_version(self):
        enc = [t for n, t in list(self.named_parameters()) + list(self.named_buffers())
               if not n.startswith("fc_tile")]
        return tuple((t._version, t.data_ptr()) for t in enc)

    def _fc_version(self):
        return tuple((t._version, t.data_ptr()) for t in self.fc_tile.parameters())

    def classifier(self, device=None, need_fc=True):
        """The libcellseg_b200 model for the current weights (rebuilt when they change).
        need_fc=False leaves a stale device copy of fc_tile alone: callers that only want the
        pooled features (training, where fc_tile runs under autograd) then pay no host sync
        after every optimizer step."""
        if device is None:
            device = self.conv1.weight.device
        if device.type != "cuda":
            raise ops._capi.CellSegError("the tile classifier runs on sm_100a only; move the model to a "
                                         "CUDA device (there is no CPU fallback)")
        key = (str(device), self._encoder_version())
        fc = self.fc_tile[1]
        if fc.out_features != 2:
            raise ops._capi.CellSegError("fc_tile must have 2 outputs (the MILres* constructors set this; pass "
                                         "num_classes=2 to a bare MILResNet / MILResNeXt)")
        if self._clf is None or self._clf_key != key:
            if self._clf is not None:
                self._clf.close()
            self._clf = ops.TileClassifier(self.encoder_name, self.folded_convs(), fc.weight, fc.bias,
                                           device=device)
            self._clf_key = key
            self._fc_key = self._fc_version()
        elif need_fc and self._fc_key != self._fc_version():
            self._clf.set_fc(fc.weight, fc.bias)
            self._fc_key = self._fc_version()
        return self._clf

    def encode(self, x):
        """Pooled 512*expansion-d features avgpool(x4)+maxpool(x4) (model/resnet.py:266), no grad."""
        with torch.no_grad():
            feat = self.classifier(x.device, need_fc=False).forward_tensor(
                x.contiguous().float(), precision=self.precision, max_batch=self.max_batch,
                want_features=True, want_logits=False)
        return feat

    def forward(self, x, freeze_bn=False):
        if self.mode != "tile":
            if self.mode in ("image", "segment"):
                raise NotImplementedError("mode %r (Stage 1 / Stage 3 heads) is out of scope of the "
                                          "B200 hot path" % self.mode)
            raise Exception("Something wrong in setmode.")
        if self.training and not freeze_bn:
            raise NotImplementedError("tile mode with batch-statistics BN: the hot path always runs the "
                                      "encoder with running statistics (model.eval() or freeze_bn=True)")
        if any(p.requires_grad for p in self.conv1.parameters()):
            raise NotImplementedError("training the encoder in tile mode (--scratch) is out of scope")
        feat = self.encode(x)
        return self.fc_tile(feat)


def MILresnet18(pretrained=False, **kwargs):
    if pretrained:
        raise RuntimeError("no network access: load weights with load_state_dict() instead")
    model = MILResNet("resnet18", BasicBlock, [2, 2, 2, 2], **kwargs)
    model.fc_tile[1] = nn.Linear(model.fc_tile[1].in_features, 2)   # model/resnet.py:342
    return model


def MILresnet34(pretrained=False, **kwargs):
    if pretrained:
        raise RuntimeError("no network access: load weights with load_state_dict() instead")
    model = MILResNet("resnet34", BasicBlock, [3, 4, 6, 3], **kwargs)
    model.fc_tile[1] = nn.Linear(model.fc_tile[1].in_features, 2)   # model/resnet.py:351
    return model


def MILresnet50(pretrained=False, **kwargs):
    if pretrained:
        raise RuntimeError("no network access: load weights with load_state_dict() instead")
    model = MILResNet("resnet50", Bottleneck, [3, 4, 6, 3], expansion=4, **kwargs)
    model.fc_tile[1] = nn.Linear(model.fc_tile[1].in_features, 2)   # model/resnet.py:360
    return model
