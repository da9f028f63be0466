"""GP objective oracle: the celerite recurrence restated from the paper must
agree with dense linear algebra on the same kernel (the solver itself is an absent third-party
dependency: parity unpinned, see oracle/gp_oracle.py)."""
import numpy as np
import pytest

from conftest import load_golden


@pytest.mark.parametrize("kind,kpar", [
    ("Real", [-14.0, -1.0986122886681098, -9.8]),
    ("Real", [-9.0, 1.5, -11.0]),
    ("quasi", [-12.842600381780318, 0.488322328246102, 0.9162907318741551, 0.0, -8.996145323242303]),  # logged GP vector
    ("quasi", [-10.0, 2.0, 1.2, 0.7, -9.5]),
])
def test_recurrence_matches_dense_cholesky(kind, kpar):
    from oracle import gp_oracle as gp
    g = load_golden("kat_epic203311200")
    lc = g["lc"]
    rng = np.random.default_rng(4)
    resid = lc[:, 1] - 1.0 + 2e-4 * np.sin(lc[:, 0] * 3.0) + 1e-5 * rng.standard_normal(len(lc))
    co = gp.coefficients(kind, kpar)
    a = gp.celerite_loglike(lc[:, 0], resid, lc[:, 2], co)
    b = gp.dense_loglike(lc[:, 0], resid, lc[:, 2], co)
    assert abs(a - b) < 1e-9 * abs(b)


def test_mono_logprob_reduces_to_white_noise_when_the_kernel_vanishes():
    """With a -> 0 the GP objective is the Gaussian log-likelihood with the -1/2 convention and the
    normalisation the no-GP objective drops (quirk Q1): ll_gp = ll_mono/4 - sum(log sigma) - N/2 log 2pi."""
    from oracle import gp_oracle as gp, namaste_oracle as o
    g = load_golden("kat_epic203311200")
    lc, th = g["lc"], g["params"][0]
    params = np.concatenate([[-80.0, 0.0], th])
    got = gp.mono_logprob(params, lc, {}, "Real", [-80.0, 0.0, -80.0], [True, True, False])
    want = o.log_like(th, lc) / 4 - np.sum(np.log(lc[:, 2])) - 0.5 * len(lc) * np.log(2 * np.pi)
    assert abs(got - want) < 1e-10 * abs(want)
