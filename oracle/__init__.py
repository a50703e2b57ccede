"""CPU oracle for the Stage-2 MIL hot path + Stage-3 HSV refinement.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it,
and there only as the checker (or as the timed CPU baseline), never as the thing shipped.

Each function restates one reference function and cites the file:line it follows under the
reference tree (Newiz430/CellSegmentation).  The reference is pure Python on top of
torch / torchvision / numpy / OpenCV / scikit-image; those third-party kernels are used
here exactly where the reference calls them (torch CPU conv/BN/pool/linear/softmax,
np.lexsort, cv2.cvtColor/threshold).  scikit-image is absent from this image, so
remove_small_regions is restated with scipy.ndimage (skimage 0.19.0 semantics).

Pinning: the reference ships no tests or golden vectors, so the oracle is pinned
against outputs of the reference's own code imported in the build container
(oracle/ref_shim.py + tests/golden/make_golden.py); the generated vectors are committed
under tests/golden/ and checked by tests/test_oracle_golden.py.
"""
