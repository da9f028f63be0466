import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import namaste_b200 as nb
g = np.load(os.path.join(ROOT, "tests/golden/occultquad_cases.npz"))
z, p, g1, g2 = g["z"], g["p"], g["g1"], g["g2"]
key = np.column_stack((p, g1, g2))
start = 0
np.set_printoptions(precision=17, linewidth=200)
tot_bad = 0
while start < len(z):
    stop = start
    while stop < len(z) and np.array_equal(key[stop], key[start]):
        stop += 1
    out = np.column_stack(nb.occultquad(z[start:stop], p[start], (g1[start], g2[start]), retall=True))
    ref = g["each"][start:stop]; refb = g["batch"][start:stop]
    nanmis = np.isnan(out) != np.isnan(ref)
    ok = ~np.isnan(ref) & ~np.isnan(out)
    d = np.where(ok, np.abs(out - ref), 0)
    db = np.where(~np.isnan(refb) & ~np.isnan(out), np.abs(out - refb), 0)
    rows = np.where(nanmis.any(axis=1) | (d.max(axis=1) > 2e-15))[0]
    print("p=%g g=(%g,%g): n=%d nanmis=%d maxd_each=%.3e maxd_batch=%.3e exact_rows=%d" % (
        p[start], g1[start], g2[start], stop - start, nanmis.sum(), d.max(), db.max(), np.sum((out == ref).all(axis=1))))
    for r in rows[:6]:
        zz = z[start + r]
        print("   z=%r (z-(1+p)=%.3e, z-(1-p)=%.3e, z-p=%.3e)\n      got %s\n      ref %s" % (
            zz, zz - (1 + abs(p[start])), zz - (1 - abs(p[start])), zz - abs(p[start]), out[r], ref[r]))
    tot_bad += len(rows)
    start = stop
print("total flagged rows", tot_bad)
