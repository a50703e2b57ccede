"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares."""
import ctypes
import os
import re

from conftest import ROOT
from cellsegmentation_b200 import _capi


def _declared():
    text = open(os.path.join(ROOT, "include", "cellseg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cs_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    names = _declared()
    assert len(names) >= 20
    l = ctypes.CDLL(_capi.LIB_PATH)
    for n in names:
        assert hasattr(l, n), "missing export: " + n
    assert sorted(_capi.EXPORTED) == names, "ctypes prototypes out of sync with the header"


def test_host_only_entry_points():
    l = _capi.lib()
    assert l.cs_version() >= 100
    assert l.cs_last_error() is not None
    for (dim, S, I), want in {(299, 32, 5): 55, (299, 32, 20): 15, (299, 16, 5): 58, (299, 32, 2): 135,
                              (31, 32, 5): 0, (32, 32, 5): 1}.items():
        assert l.cs_grid_count(dim, S, I) == want


def test_grid_coords_match_oracle():
    from cellsegmentation_b200 import ops
    from oracle import tiles as otiles
    import numpy as np
    for (H, W, S, I) in [(299, 299, 32, 5), (299, 299, 32, 20), (299, 299, 16, 5), (64, 48, 16, 9),
                         (299, 299, 32, 3)]:
        got = ops.grid_coords(H, W, S, I)
        assert np.array_equal(got, np.array(otiles.get_tiles((H, W, 3), I, S), np.int32))
