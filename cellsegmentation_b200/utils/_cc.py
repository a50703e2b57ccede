"""Host connected-component clean-up used by remove_small_regions.

The reference calls skimage.morphology.remove_small_objects / remove_small_holes
(utils/image_processing.py:15-16, scikit-image 0.19.0, connectivity 1); scikit-image is not in
this image, so the same semantics are expressed with scipy.ndimage (label with a 4-neighbour
structure, drop components with size < threshold).  GPU labelling is the next row (N1)."""
import numpy as np
from scipy import ndimage as ndi


def remove_small_objects(ar, min_size=64, connectivity=1):
    out = np.array(ar, dtype=bool, copy=True)
    if min_size == 0:
        return out
    ccs, _ = ndi.label(out, ndi.generate_binary_structure(out.ndim, connectivity))
    sizes = np.bincount(ccs.ravel())
    out[(sizes < min_size)[ccs]] = False
    return out


def remove_small_holes(ar, area_threshold=64, connectivity=1):
    out = np.logical_not(np.array(ar, dtype=bool))
    out = remove_small_objects(out, area_threshold, connectivity)
    return np.logical_not(out)
