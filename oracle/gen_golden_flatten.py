"""Golden vectors for the detrending row: the reference's own
namaste/k2flatten.py (imported unmodified from the reference checkout, build container only) applied
to its example light curve the way Namaste does (namaste.py:442 and Example.ipynb cell 45:
ReduceNoise(lc, winsize=4.5, stepsize=0.125), twice), plus one run with the function's defaults on a
light curve with an artificial gap (exercises the window-shifting branches of formwindow).
    python oracle/gen_golden_flatten.py  ->  tests/golden/flatten_epic203311200.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import _refimport, gen_golden as gg  # noqa: E402


def main():
    kf = _refimport.k2flatten()
    lc0 = gg.load_example_lc()
    flat1 = kf.ReduceNoise(lc0, winsize=4.5, stepsize=0.125)
    flat2 = kf.ReduceNoise(flat1, winsize=4.5, stepsize=0.125)
    gap = lc0[(lc0[:, 0] < lc0[0, 0] + 12) | (lc0[:, 0] > lc0[0, 0] + 14.5)][:1500]
    gap_out = kf.ReduceNoise(gap)
    out = os.path.join(gg.OUT, "flatten_epic203311200.npz")
    np.savez_compressed(out, lc=lc0, flat1=flat1[:, 1], flat2=flat2[:, 1], err1=flat1[:, 2], gap_lc=gap, gap_flat=gap_out[:, 1])
    print("wrote", out, os.path.getsize(out), "bytes;", "zeros left in flat1:", int((flat1[:, 1] == 0).sum()))


if __name__ == "__main__":
    main()
