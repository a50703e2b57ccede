"""MILResNeXt drop-in (mirror of model/resnext.py of the reference): the same module tree and
state_dict keys for the encoder and the tile head, tile-mode forward on the sm_100a kernels.

The reference class differs from MILResNet only in the Bottleneck it builds (grouped 3x3 of
width planes * width_per_group / 64 * groups, model/resnext.py:76-91) and in keeping
fc_tile at `num_classes` outputs unless pretrained weights are loaded (model/resnext.py:407-415);
pass num_classes=2, as the pretrained path ends up with.
"""
import torch.nn as nn

from .resnet import MILResNet, Bottleneck

__all__ = ["MILResNeXt", "MILresnext50_32x4d", "MILresnext101_32x8d"]


class MILResNeXt(MILResNet):

    def __init__(self, encoder, block, layers, num_classes=1000, groups=1, width_per_group=64):
        super().__init__(encoder, block, layers, num_classes=num_classes, expansion=block.expansion,
                         groups=groups, width_per_group=width_per_group)


def MILresnext50_32x4d(pretrained=False, progress=True, **kwargs):
    if pretrained:
        raise RuntimeError("no network access: load weights with load_state_dict() instead")
    kwargs["groups"] = 32
    kwargs["width_per_group"] = 4
    kwargs.setdefault("num_classes", 2)
    model = MILResNeXt("resnext50_32x4d", Bottleneck, [3, 4, 6, 3], **kwargs)
    if model.fc_tile[1].out_features != 2:
        model.fc_tile[1] = nn.Linear(model.fc_tile[1].in_features, 2)   # model/resnext.py:414
    return model


def MILresnext101_32x8d(pretrained=False, progress=True, **kwargs):
    if pretrained:
        raise RuntimeError("no network access: load weights with load_state_dict() instead")
    kwargs["groups"] = 32
    kwargs["width_per_group"] = 8
    kwargs.setdefault("num_classes", 2)
    model = MILResNeXt("resnext101_32x8d", Bottleneck, [3, 4, 23, 3], **kwargs)
    if model.fc_tile[1].out_features != 2:
        model.fc_tile[1] = nn.Linear(model.fc_tile[1].in_features, 2)   # model/resnext.py:414
    return model
