"""The oracle must reproduce the reference's only recorded numerical outputs for this
path: three (params -> negative log-prob) triples logged by Optimize_mono
(Example.ipynb cell 47; namaste/namaste_runlog.log tag ``1048:namaste.py``)."""
import numpy as np

from conftest import load_golden, priors_from_table
from oracle import namaste_oracle as o


def test_three_logged_known_answers():
    g = load_golden("kat_epic203311200")
    lc, pri = g["lc"], priors_from_table(g["priors"])
    assert lc.shape == (226, 3)
    for th, nll in zip(g["params"], g["nll_logged"]):
        got = -o.mono_only_logprob(th, lc, pri)
        # 1.14e-11 residual = LAPACK polyfit noise in the (out-of-scope) flattening step
        assert abs(got - nll) / nll < 5e-11


def test_kat_components_bitwise_vs_reference_run():
    g = load_golden("kat_epic203311200")
    lc, pri = g["lc"], priors_from_table(g["priors"])
    for th, ll, lp, mdl in zip(g["params"], g["ll"], g["lp"], g["model"]):
        assert o.log_like(th, lc) == ll
        assert o.mono_ln_prior(th, pri) == lp
        assert np.array_equal(o.get_value(th, lc[:, 0]), mdl)


def test_prior_spot_values():
    # SURVEY.md App. C.1 (KAT #1): LD1 pdf, LD2 pdf, normlim term
    g = load_golden("kat_epic203311200")
    pri = priors_from_table(g["priors"])
    th = g["params"][0]
    assert abs(o.mono_ln_prior(th, pri) - 24.7726729934) < 1e-9
