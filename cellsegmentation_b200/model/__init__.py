"""`nets` registry (mirror of model/__init__.py:5-13).  The reference instantiates every
backbone with pretrained=True at import time (needs the network); here entries are built
lazily with random init and weights come from load_state_dict()."""
from .resnet import MILResNet, MILresnet18, MILresnet34, MILresnet50
from .resnext import MILResNeXt, MILresnext50_32x4d, MILresnext101_32x8d

_CTORS = {"resnet18": MILresnet18, "resnet34": MILresnet34, "resnet50": MILresnet50,
          "resnext50_32x4d": MILresnext50_32x4d, "resnext101_32x8d": MILresnext101_32x8d}


class _Nets(dict):
    def __missing__(self, key):
        if key not in _CTORS:
            raise KeyError("%r has no B200 kernels yet (available: %s)" % (key, sorted(_CTORS)))
        self[key] = _CTORS[key]()
        return self[key]

    def __contains__(self, key):
        return key in _CTORS


nets = _Nets()
