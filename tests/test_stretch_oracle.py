"""CPU checks of the sampler oracle: Philox4x32-10 known answers (Random123 kat_vectors) and
detailed balance sanity on a Gaussian target."""
import numpy as np

from oracle import stretch_oracle as so


def test_philox_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = so.philox4x32_10([ctr[0]], [ctr[1]], [ctr[2]], [ctr[3]], key[0], key[1])
        assert tuple(int(x[0]) for x in got) == want


def test_uniforms_are_uniform_and_reproducible():
    uz, uj, ua = so.draws(1234, 7, 1, np.arange(20000))
    for u in (uz, uj, ua):
        assert 0.0 <= u.min() and u.max() < 1.0
        assert abs(u.mean() - 0.5) < 0.01 and abs(u.var() - 1 / 12) < 0.005
    again = so.draws(1234, 7, 1, np.arange(20000))
    assert all(np.array_equal(a, b) for a, b in zip((uz, uj, ua), again))
    other = so.draws(1234, 8, 1, np.arange(20000))
    assert not np.array_equal(uz, other[0])


def test_stretch_move_samples_a_gaussian():
    rng = np.random.default_rng(0)
    sig = np.array([1.0, 0.1, 3.0])
    lnp = lambda p: -0.5 * np.sum((p / sig) ** 2, axis=1)
    chain, lnc, nacc = so.run(lnp, rng.standard_normal((40, 3)) * sig, 1500, seed=99)
    flat = chain[:, 500:, :].reshape(-1, 3)
    assert np.allclose(flat.std(axis=0), sig, rtol=0.1)
    assert np.all(np.abs(flat.mean(axis=0)) < 0.15 * sig)
    assert 0.2 < nacc.mean() / 1500 < 0.8
