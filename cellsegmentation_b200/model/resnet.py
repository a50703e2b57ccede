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
Image and segment modes run forward-only (model.eval()): the encoder at
299 x 299 and the Stage-3 decoder on the fp32 CUDA-core kernels (cs_model_forward_image,
cs_conv2d_nhwc_f32, cs_resize_bilinear_nhwc_f32), the small fc_image heads in torch.  That is
what test_tile.py --reg_limit (:88-105), train_seg.py's mask generation (:255-269) and
test_seg.py call; TRAINING the Stage-1 heads or the decoder stays out of scope and raises.
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

    def forward(self, x):
        """Autograd path (encoder training, --scratch): the module tree evaluated by torch."""
        idt = x if self.downsample is None else self.downsample(x)
        out = self.relu(self.bn1(self.conv1(x)))
        return self.relu(self.bn2(self.conv2(out)) + idt)


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

    def forward(self, x):
        """Autograd path (encoder training, --scratch): the module tree evaluated by torch."""
        idt = x if self.downsample is None else self.downsample(x)
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.relu(self.bn2(self.conv2(out)))
        return self.relu(self.bn3(self.conv3(out)) + idt)


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
        self.max_batch = 75776

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
        # Stage-1 heads (model/resnet.py:129-153, map_size = 1) and Stage-3 decoder (:155-165)
        feat = 512 * block.expansion
        self.avgpool_image = nn.AdaptiveAvgPool2d((1, 1))
        self.maxpool_image = nn.AdaptiveMaxPool2d((1, 1))
        self.fc_image_cls = nn.Sequential(nn.Flatten(), nn.BatchNorm1d(feat), nn.Dropout(p=0.25), nn.ReLU(inplace=True),
                                          nn.Linear(feat, 64), nn.BatchNorm1d(64), nn.Dropout(), nn.Linear(64, 7))
        self.fc_image_reg = nn.Sequential(nn.Flatten(), nn.BatchNorm1d(feat), nn.Dropout(p=0.25), nn.ReLU(inplace=True),
                                          nn.Linear(feat, 64), nn.BatchNorm1d(64), nn.Dropout(), nn.Linear(64, 1),
                                          nn.ReLU(inplace=True))
        e = expansion
        self.expansion = expansion
        self.upconv1 = self.upsample_conv(512 * e, 256 * e)
        self.upconv2 = self.upsample_conv(512 * e, 256 * e)
        self.upconv3 = self.upsample_conv(256 * e, 128 * e)
        self.upconv4 = self.upsample_conv(256 * e, 128 * e)
        self.upconv5 = self.upsample_conv(128 * e, 64 * e)
        self.upconv6 = self.upsample_conv(128 * e, 64 * e)
        self.upconv7 = self.upsample_conv(64 * e, 64 if e == 1 else 32 * e)
        self.upconv8 = self.upsample_conv(64 if e == 1 else 32 * e, 64)
        self.seg_out_conv = nn.Conv2d(64, 2, kernel_size=1)
        self.image_max_batch = 4       # images per fp32 encoder chunk (299 x 299: ~15 MB of maps each)
        self._dec_key = None
        self._dec = None
        for m in self.modules():                       # model/resnet.py:171-178
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
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

    @staticmethod
    def upsample_conv(in_channels, out_channels):      # model/resnet.py:194-199
        return nn.Sequential(nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1),
                             nn.BatchNorm2d(out_channels), nn.ReLU(inplace=True))

    # ---- requires_grad groups (model/resnet.py:196-232, 308-333) ----------------------------
    def set_encoder_grads(self, requires_grad):
        for m in (self.conv1, self.bn1, self.layer1, self.layer2, self.layer3, self.layer4):
            m.requires_grad_(requires_grad)

    def set_tile_module_grads(self, requires_grad):
        self.fc_tile.requires_grad_(requires_grad)

    def set_image_module_grads(self, requires_grad):
        self.fc_image_cls.requires_grad_(requires_grad)
        self.fc_image_reg.requires_grad_(requires_grad)

    def set_seg_module_grads(self, requires_grad):     # upconv5..8 keep their state, as in the reference (:228-232)
        for m in (self.upconv1, self.upconv2, self.upconv3, self.upconv4, self.seg_out_conv):
            m.requires_grad_(requires_grad)

    def setmode(self, mode):
        if mode not in ("tile", "image", "segment"):
            raise Exception("Invalid mode: {}.".format(mode))
        self.set_encoder_grads(mode == "image")
        self.set_tile_module_grads(mode == "tile")
        self.set_image_module_grads(mode == "image")
        self.set_seg_module_grads(mode == "segment")
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
               if n.startswith(("conv1", "bn1", "layer"))]
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

    # ---- N4: whole-image encoder, Stage-1 heads, Stage-3 decoder (forward only) --------------
    def _decoder(self, device):
        """BN-folded decoder weights on the device, packed [k*k*Cin, Cout] for cs_conv2d_nhwc_f32."""
        mods = [self.upconv1, self.upconv2, self.upconv3, self.upconv4, self.upconv5, self.upconv6,
                self.upconv7, self.upconv8]
        tensors = [t for m in mods + [self.seg_out_conv] for t in list(m.parameters()) + list(m.buffers())]
        key = (str(device), tuple((t._version, t.data_ptr()) for t in tensors))
        if self._dec_key != key:
            dec = []
            with torch.no_grad():
                for m in mods:
                    conv, bn = m[0], m[1]
                    s = bn.weight.float() / torch.sqrt(bn.running_var.float() + bn.eps)
                    w = conv.weight.float() * s[:, None, None, None]
                    b = (conv.bias.float() - bn.running_mean.float()) * s + bn.bias.float()
                    dec.append((w.permute(2, 3, 1, 0).reshape(-1, w.shape[0]).contiguous().to(device),
                                b.contiguous().to(device)))
                w = torch.zeros((4, 64, 1, 1), dtype=torch.float32, device=self.seg_out_conv.weight.device)
                b = torch.zeros(4, dtype=torch.float32, device=w.device)
                w[:2] = self.seg_out_conv.weight.float()       # Cout padded 2 -> 4 for the float4 epilogue
                b[:2] = self.seg_out_conv.bias.float()
                dec.append((w.permute(2, 3, 1, 0).reshape(-1, 4).contiguous().to(device), b.to(device)))
            self._dec, self._dec_key = dec, key
        return self._dec

    def _forward_whole_image(self, x):
        if self.training:
            raise NotImplementedError("mode %r runs forward-only on the B200 stack (call model.eval()): training "
                                      "the Stage-1 heads / Stage-3 decoder is out of scope" % self.mode)
        if x.device.type != "cuda":
            raise ops._capi.CellSegError("the encoder runs on sm_100a only; move the input to a CUDA device")
        clf = self.classifier(x.device, need_fc=False)
        x = x.contiguous().float()
        with torch.no_grad():
            if self.mode == "image":                           # model/resnet.py:271-278
                feat = clf.forward_image(x, max_batch=self.image_max_batch)
                out = feat.view(feat.shape[0], -1, 1, 1)       # avgpool_image(x4) + maxpool_image(x4), map size 1
                return self.fc_image_cls(out), self.fc_image_reg(out)
            if self.encoder_name.startswith("resnext"):
                raise NotImplementedError("segment mode of MILResNeXt: the reference's decoder (512-channel upconv1, "
                                          "model/resnext.py:209) does not fit its own 2048-channel x4 either")
            x1, x2, x3, x4 = clf.forward_image(x, max_batch=self.image_max_batch, want_features=False, want_maps=True)
            d = self._decoder(x.device)
            up = lambda t, i: ops.conv2d_nhwc(t, d[i][0], d[i][1], 3, 1, 1, relu=True)     # noqa: E731
            o = up(ops.resize_bilinear_nhwc(x4, x3.shape[1]), 0)                            # :282-283
            o = up(torch.cat([o, x3], dim=3), 1)                                            # :284-285
            o = up(ops.resize_bilinear_nhwc(o, x2.shape[1]), 2)                             # :287-288
            o = up(torch.cat([o, x2], dim=3), 3)                                            # :289-290
            o = up(ops.resize_bilinear_nhwc(o, x1.shape[1]), 4)                             # :292-293
            o = up(torch.cat([o, x1], dim=3), 5)                                            # :294-295
            o = up(ops.resize_bilinear_nhwc(o, (x.shape[2] - 1) // 2 + 1), 6)               # :297-298 (150)
            o = up(o, 7)                                                                    # :299
            o = ops.resize_bilinear_nhwc(o, x.shape[2])                                     # :300
            o = ops.conv2d_nhwc(o, d[8][0], d[8][1], 1, 1, 0, relu=False)[..., :2]          # :301
            return o.permute(0, 3, 1, 2).contiguous()

    def forward(self, x, freeze_bn=False):
        if self.mode != "tile":
            if self.mode in ("image", "segment"):
                return self._forward_whole_image(x)
            raise Exception("Something wrong in setmode.")
        if self.conv1.weight.requires_grad or (self.training and not freeze_bn):
            # encoder training (train_tile.py --scratch, :272-273) or batch-statistics BN: gradients /
            # statistics have to flow through the encoder, which the forward-only sm_100a kernels do
            # not provide.  Evaluate the module tree with torch autograd instead (functional, not the
            # accelerated path; the frozen-encoder case below is the hot path).
            return self._forward_autograd(x, freeze_bn)
        feat = self.encode(x)
        return self.fc_tile(feat)

    def _forward_autograd(self, x, freeze_bn):
        was_training = self.training
        if freeze_bn:                                  # model/resnet.py:254-258
            self.eval()
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x4 = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        if freeze_bn and was_training:
            self.train()
        return self.fc_tile(self.avgpool_tile(x4) + self.maxpool_tile(x4))


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
