"""Import the UNMODIFIED reference from /root/reference in this container (no network, no
h5py / skimage / matplotlib / openslide).  Used only to generate tests/golden/*.npz and to
cross-check the oracle locally; /root/reference does not exist on the GPU box.

Recipe from SURVEY Appendix A.
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("CELLSEG_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "model"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _LegacyNumpy:
    """numpy as dataset/dataset.py:168 expects it (requirements.txt pins numpy 1.22): np.array()
    of ragged rows `(bag, (x, y), label)` yields an object[M,3] array (with a deprecation warning)
    instead of raising as numpy >= 1.24 does.  Everything else is the installed numpy."""

    def __init__(self, real):
        self._real = real

    def __getattr__(self, name):
        return getattr(self._real, name)

    def array(self, obj, *args, **kwargs):
        try:
            return self._real.array(obj, *args, **kwargs)
        except ValueError:
            rows = list(obj)
            out = self._real.empty((len(rows), 3), dtype=object)
            for r, row in enumerate(rows):
                out[r, 0], out[r, 1], out[r, 2] = row
            return out


_imported = None


def import_reference():
    """Returns a namespace with the reference's dataset, utils, inference, evaluate modules."""
    global _imported
    if _imported is not None:
        return _imported
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    import torch
    import torch.hub
    import torch.utils.model_zoo as mz
    from . import masks as omasks

    _stub("h5py")
    mp = _stub("matplotlib")
    mp.pyplot = _stub("matplotlib.pyplot")
    sk = _stub("skimage")
    captured = {"imsave": []}

    def _imsave(path, arr, *a, **k):
        captured["imsave"].append((path, arr.copy()))

    sk.io = _stub("skimage.io", imsave=_imsave, imread=None)
    # skimage is not installed: the two morphology calls get the scipy restatement
    sk.morphology = _stub("skimage.morphology",
                          remove_small_objects=omasks.remove_small_objects,
                          remove_small_holes=omasks.remove_small_holes)
    _stub("openslide", OpenSlide=object)
    _stub("torch._six", string_classes=(str, bytes))
    torch.hub.load_state_dict_from_url = mz.load_url = lambda *a, **k: {}
    sys.path.insert(0, REF_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import dataset as ref_dataset
        import utils as ref_utils
        import inference as ref_inference
        import evaluate as ref_evaluate
        import train as ref_train
    import numpy
    # make_train_data (dataset/dataset.py:166-201) runs unmodified once its module sees the numpy
    # it was written for (ragged rows -> object array)
    sys.modules[ref_dataset.LystoDataset.__module__].np = _LegacyNumpy(numpy)
    ns = types.SimpleNamespace(dataset=ref_dataset, utils=ref_utils, inference=ref_inference,
                               evaluate=ref_evaluate, train=ref_train, captured=captured)
    _imported = ns
    return ns


def load_reference_module(name):
    """model/<name>.py alone (resnet.py / resnext.py need only torch)."""
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF_ROOT, "model", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_resnet():
    return load_reference_module("resnet")
