"""Bitwise comparison of the oracle with the reference's own modules, imported by path.
Runs only where /root/reference is mounted (this container); skipped on the GPU box."""
import numpy as np
import pytest

from oracle import _refimport, namaste_oracle as o

pytestmark = pytest.mark.skipif(not _refimport.available(), reason="/root/reference not mounted")


@pytest.fixture(scope="module")
def ref():
    return _refimport.crossfield()


def test_elliptic(ref):
    rng = np.random.default_rng(7)
    k = rng.uniform(0, 1, 20000)
    n = rng.uniform(0, 500, 20000)
    for a, b in zip(ref.ellke(k), o.ellke(k)):
        assert np.array_equal(a, b)
    assert np.array_equal(ref.ellpic_bulirsch(n, k), o.ellpic_bulirsch(n, k))


@pytest.mark.parametrize("p", [0.01, 0.0585, 0.1, 0.3, 0.5, 0.6, 1.0, 1.5, -0.1])
def test_occult(ref, p):
    rng = np.random.default_rng(11)
    ap = abs(p)
    z = np.abs(np.concatenate([rng.uniform(0, 2.6, 4000), [0.0, ap, 1 - ap, 1 + ap, abs(1 - ap), 0.5]]))
    assert np.array_equal(ref.occultuniform(z, p), o.occultuniform(z, p), equal_nan=True)
    for g in [(0.3, 0.2), (0, 0), (0.6, -0.1), (1.2, 0.4)]:
        A = ref.occultquad(z, p, np.array(g), retall=True)
        B = o.occultquad(z, p, np.array(g), retall=True)
        for x, y in zip(A, B):
            assert np.array_equal(x, y, equal_nan=True)
