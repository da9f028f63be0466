"""Repeat the same evaluations many times and report any run whose bits differ from the first
(the pipeline promises bit-reproducible results).  usage: python tools/gpu_stress_determinism.py [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import namaste_b200 as nb
from conftest import load_golden, priors_from_table

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
g = load_golden("kat_epic203311200")
pri = priors_from_table(g["priors"])
cases = []
lc1 = nb.LightCurve(g["lc"], oversamp=10)
cases.append(("kat W=1", lc1, g["params"][:1], g["priors"]))
cases.append(("kat W=3", lc1, g["params"], g["priors"]))
for name in ("config1_epic203311200", "synth_k2_s11"):
    f = load_golden(name)
    S = int(f["oversamp"]) if "oversamp" in f.files else 10
    cases.append((name, nb.LightCurve(f["lc"], oversamp=S), f["walkers"], f["priors"]))
    big = np.tile(f["walkers"], (40, 1))[:1000]
    cases.append((name + " x1000", cases[-1][1], big, f["priors"]))
bad = 0
for label, lc, pos, tab in cases:
    for mode in ("brute", "windowed"):
        for dt in ("f64", "strict"):
            first = lc.lnprob(pos, tab, mode=mode, dtype=dt)
            nmis, worst = 0, 0.0
            for r in range(reps):
                # alternate with the other mode in between, like the test-suite does
                if r % 3 == 0:
                    lc.lnprob(pos, tab, mode="windowed" if mode == "brute" else "brute", dtype=dt)
                got = lc.lnprob(pos, tab, mode=mode, dtype=dt)
                if not np.array_equal(got, first, equal_nan=True):
                    nmis += 1
                    fin = np.isfinite(first)
                    worst = max(worst, float(np.max(np.abs(got[fin] - first[fin]) / np.abs(first[fin]))))
            print(f"{label:32s} {mode:9s} {dt:7s} W={len(pos):5d} mismatching runs {nmis}/{reps} worst rel {worst:.3g}", flush=True)
            bad += nmis
print("TOTAL mismatches", bad)
