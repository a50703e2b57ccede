"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares."""
import ctypes
import os
import re

from conftest import ROOT
from cellsegmentation_b200 import _capi
from oracle import tiles as otiles


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


def test_grid_cover_matches_brute_force():
    """cs_grid_cover_host (the range of grid positions covering a pixel, used by the heat-map
    gather) against the definition: positions g with coord(g) <= c < coord(g) + S, coord from
    get_tiles (dataset/dataset.py:718-742), incl. strides that do not divide the span, strides
    larger than the tile (gaps: empty range) and single-position axes."""
    import ctypes
    from cellsegmentation_b200 import _capi
    lib = _capi.lib()
    lo, hi = ctypes.c_int32(), ctypes.c_int32()
    for dim in (32, 33, 40, 64, 100, 131, 299):
        for S in (8, 16, 32):
            if S > dim:
                continue
            for I in (1, 2, 3, 5, 7, 8, 16, 20, 33, 40):
                coords = sorted({r for r, _ in otiles.get_tiles((dim, S, 3), I, S)})
                assert len(coords) == lib.cs_grid_count(dim, S, I)
                for c in range(dim):
                    assert lib.cs_grid_cover_host(c, dim, S, I, ctypes.byref(lo), ctypes.byref(hi)) == 0
                    want = [g for g, r in enumerate(coords) if r <= c < r + S]
                    assert list(range(lo.value, hi.value + 1)) == want, (dim, S, I, c)
