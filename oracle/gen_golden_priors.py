"""Golden fixture for the prior-construction row of the path.

Inputs come from the reference checkout (run in the build container only): the example light
curve and the parameter PDFs the reference saved for EPIC 203311200
(InputFiles/203311200.1_inputsamples.npy: columns tcen, b, RpRs, vel, teff, srad, smass, logg,
sdens, LD1, LD2).  The expected table is NOT computed here: it is the priors dict the reference
itself logged (Example.ipynb cell 47 / namaste_runlog.log tag 1048, see gen_golden.KAT_PRIORS), so
the test pins calcminp / calcmaxvel / monoPriors / initLD against the reference's own output.
    python oracle/gen_golden_priors.py  ->  tests/golden/priors_epic203311200.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import gen_golden as gg  # noqa: E402


def main():
    lc = gg.load_example_lc()
    pdfs = np.load(os.path.join(gg.REF, "InputFiles", "203311200.1_inputsamples.npy"))
    out = os.path.join(gg.OUT, "priors_epic203311200.npz")
    np.savez_compressed(out, t=lc[:, 0], sdenss=pdfs[:, 8], LD1s=pdfs[:, 9], LD2s=pdfs[:, 10],
                        tcen=2121.02, tdur=0.5, expected=gg.KAT_PRIORS, pdfs=pdfs[:, [0, 1, 3, 2, 9, 10]])
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
