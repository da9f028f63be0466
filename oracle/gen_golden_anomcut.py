"""Golden vectors for the single-point outlier cut that precedes the detrending row (SURVEY.md 8f
rank 3): the reference's own ``AnomCutDiff`` (namaste/namaste.py:1459-1467, run from its source text,
see oracle/_refimport.namaste_function; build container only) on (a) the example light curve before
the cut, at the threshold Namaste uses (3.75, namaste.py:64 / :1109) and at the default 4.2, and (b) a
synthetic series with planted single- and double-point outliers.
    python oracle/gen_golden_anomcut.py  ->  tests/golden/anomcut.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import _refimport, gen_golden as gg  # noqa: E402


def example_flux_before_cut():
    lc = np.load(os.path.join(gg.REF, "ExampleData", "EPIC203311200_bestlc_ev.npy"))
    if len(lc) == 2:
        lc = lc[0]
    lc = lc[np.logical_not(np.isnan(np.sum(lc, axis=1)))]
    lc[:, 1:] /= np.median(lc[:, 1])
    return lc[:, 1].copy()


def main():
    cut = _refimport.namaste_function("AnomCutDiff")
    epic = example_flux_before_cut()
    rng = np.random.default_rng(1459)
    syn = 1.0 + 1e-4 * rng.standard_normal(5000)
    spikes = rng.choice(np.arange(5, 4995), 60, replace=False)
    syn[spikes] += rng.choice([-1, 1], 60) * rng.uniform(3e-4, 3e-3, 60)      # single points
    pairs = rng.choice(np.arange(5, 4990), 20, replace=False)
    syn[pairs] += 2e-3
    syn[pairs + 1] += 2e-3                                                       # two in a row: kept
    out = os.path.join(gg.OUT, "anomcut.npz")
    np.savez_compressed(out, epic=epic, epic_375=cut(epic.copy(), 3.75), epic_420=cut(epic.copy()),
                        syn=syn, syn_375=cut(syn.copy(), 3.75), syn_420=cut(syn.copy()))
    print("wrote", out, "cut from the example:", int((~cut(epic.copy(), 3.75)).sum()), "points;",
          "from the synthetic series:", int((~cut(syn.copy())).sum()))


if __name__ == "__main__":
    main()
