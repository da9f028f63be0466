"""The brute-force sweep's one-fma classification (csrc/lnprob_kernels.cuh: make_sweep_consts) must never declare
a fine sample out of transit that the reference puts inside the limb.  This test restates the device constants
with exactly rounded arithmetic (fractions -> float is correctly rounded, so fma is emulated exactly) and checks the
superset property against the reference's own rounding sequence (namaste/namaste.py:1157-1161: ft = t + off,
z = sqrt(b**2 + (vel * (ft - tcen))**2), in transit iff z <= 1 + p as Crossfield_transit.py:800-803) on samples
placed at and around the edge of the window, where a wrong bound would show.  CPU only."""
import math
from fractions import Fraction as Fr

import numpy as np
import pytest

K_EPS = 2.220446049250313e-16


def fma(a, b, c):
    return float(Fr(a) * Fr(b) + Fr(c))


def sweep_consts(tcen, b, vel, rprs, span):
    """make_sweep_consts for a walker with a regular limb-darkening normalisation (fbar == 1)."""
    p = abs(rprs)
    onep = 1.0 + p
    b2 = b * b
    qthr = -1.0 if p == 0.0 else (onep * onep) * (1.0 + 16 * K_EPS)
    vtc = vel * tcen
    hx2 = qthr - b2
    if not (hx2 >= 0.0):
        return None                                   # never in transit
    xq = math.sqrt(hx2) * (1.0 + 4 * K_EPS)
    ev = K_EPS * (4 * abs(vtc) + 2 * abs(vel) * abs(span))
    X = (xq + ev) * (1.0 + 64 * K_EPS)
    if not (X >= 1e-100):
        return "always"
    s = (2.0 / X) * (1.0 - 4 * K_EPS)
    return dict(vel=vel, nvtc=-vtc, s=s, vs=vel * s)


def device_in_window(c, t, off):
    ck = fma(c["vel"], off, c["nvtc"]) * c["s"]
    y = fma(c["vs"], t, ck)
    return abs(y) < 2.0                               # bit 30 of the high word clear


def reference_in_transit(tcen, b, vel, rprs, t, off):
    ft = t + off
    z = np.sqrt(np.float64(b) ** 2 + (np.float64(vel) * (np.float64(ft) - np.float64(tcen))) ** 2)
    return bool(z <= 1.0 + abs(rprs))


@pytest.mark.parametrize("seed", range(4))
def test_window_contains_every_in_transit_sample(seed):
    rng = np.random.default_rng(seed)
    checked = inside = 0
    for _ in range(300):
        t0 = float(rng.choice([0.0, 131.5, 2061.0, 2457000.0, -5000.0]))      # BJD offsets of the BASELINE light curves
        cad = float(rng.choice([0.0204318, 2.0 / 1440, 0.5]))
        S = int(rng.choice([1, 10, 11, 15]))
        tcen = t0 + float(rng.uniform(0, 80))
        b = float(rng.choice([0.0, rng.uniform(0, 1.3), 1.0 + 0.06 - 1e-9]))
        vel = float(rng.choice([rng.uniform(0.05, 30.0), -rng.uniform(0.5, 5.0), 1e-6, 3e4]))
        rprs = float(rng.choice([rng.uniform(0.01, 0.3), -0.06, 0.7]))
        offs = cad * np.arange(0, 1.0, 1.0 / S)
        c = sweep_consts(tcen, b, vel, rprs, float(offs[-1]))
        if c is None:
            # |b| beyond the limb: the reference can have no in-transit sample at all
            assert not reference_in_transit(tcen, b, vel, rprs, tcen, 0.0)
            continue
        if c == "always":
            continue
        # the exact contact times, then timestamps a few ulps / a few 1e-9 either side of them
        h2 = (1.0 + abs(rprs)) ** 2 - b * b
        if h2 < 0:
            continue
        half = math.sqrt(h2) / abs(vel)
        for edge in (tcen - half, tcen + half):
            for k in rng.integers(0, S, size=2):
                off = float(offs[k])
                base = edge - off
                cands = [base]
                x = base
                for _ in range(6):
                    x = np.nextafter(x, np.inf); cands.append(float(x))
                x = base
                for _ in range(6):
                    x = np.nextafter(x, -np.inf); cands.append(float(x))
                cands += [base + d for d in (1e-12, -1e-12, 1e-9, -1e-9, 1e-7, -1e-7)]
                for t in cands:
                    checked += 1
                    if reference_in_transit(tcen, b, vel, rprs, t, off):
                        inside += 1
                        assert device_in_window(c, t, off), (tcen, b, vel, rprs, t, off)
        # mid-transit is always inside
        assert device_in_window(c, tcen, 0.0)
    assert checked > 1000 and inside > 200


def test_window_is_tight():
    """The window may be a superset only by rounding-sized margins: a sample 1e-6 stellar radii beyond the limb
    along the chord is already outside it (otherwise the sweep would hand stage 2 needless work)."""
    tcen, b, vel, rprs, cad = 2061.37, 0.5, 3.0, 0.06, 0.0204318
    c = sweep_consts(tcen, b, vel, rprs, cad * 10 / 11)
    half = math.sqrt((1.0 + rprs) ** 2 - b * b) / vel
    assert device_in_window(c, tcen + half - 1e-9, 0.0)
    assert not device_in_window(c, tcen + half + 1e-6 / vel, 0.0)
    assert not device_in_window(c, tcen - half - 1e-6 / vel, 0.0)
