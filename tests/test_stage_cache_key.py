"""The staging cache of the host layer (namaste_b200/namaste.py: stage_lightcurve) finds a staged light
curve again by identity AND full content: an array changed in place, or a different array that lands on
the same address with the same times and errors, must not hit the old entry."""
import numpy as np

from namaste_b200.namaste import _fingerprint


def test_fingerprint_sees_any_single_element():
    rng = np.random.default_rng(0)
    lc = np.column_stack((2000 + 0.02 * np.arange(226), 1 + 1e-4 * rng.standard_normal(226), np.full(226, 5e-5)))
    base = _fingerprint(lc)
    assert _fingerprint(lc) == base
    for i, j in [(0, 0), (17, 1), (100, 1), (225, 2), (113, 0)]:
        old = lc[i, j]
        lc[i, j] = np.nextafter(old, np.inf)      # one ulp, in place: same address, shape and strides
        assert _fingerprint(lc) != base, (i, j)
        lc[i, j] = old
    assert _fingerprint(lc) == base
    # swapping two flux values keeps the sum of the values; the key still has to change
    lc[[3, 4], 1] = lc[[4, 3], 1]
    assert _fingerprint(lc)[3:] != base[3:] or lc[3, 1] == lc[4, 1]


def test_fingerprint_layouts():
    lc = np.arange(30, dtype=np.float64).reshape(10, 3)
    assert _fingerprint(lc) != _fingerprint(lc.copy())            # different address
    assert _fingerprint(lc[::2]) != _fingerprint(lc[:5])          # different strides / content
    assert _fingerprint(lc.astype(np.float32))[1] == (10, 3)      # other dtypes are keyed on their float64 image


def test_readonly_array_is_keyed_by_identity():
    lc = np.arange(30, dtype=np.float64).reshape(10, 3).copy()    # owns its data
    full = _fingerprint(lc)
    lc.setflags(write=False)
    assert _fingerprint(lc) == full[:4]                            # no content pass for an immutable array
    view = lc[:5]                                                  # read-only view of a read-only base
    assert len(_fingerprint(view)) == 4
    rw = np.arange(30, dtype=np.float64).reshape(10, 3)
    ro_view = rw[:]
    ro_view.setflags(write=False)                                  # the base can still change under it
    assert len(_fingerprint(ro_view)) == 6
