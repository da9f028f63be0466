"""bench.py's FLOP accounting (SURVEY.md 8d convention) on a small workload: the interior / limb / in-transit counts
against a direct classification of every fine sample, and the split of the convention's FLOPs between the two
kernels.  CPU only (no kernels run)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_flop_accounting_matches_a_direct_count():
    import bench
    bench.WORKLOADS["tiny"] = (2, 600, 0.0204318, 2061.0, 5.4e-5, 5, 8)
    bench.WORKLOAD_DESC["tiny"] = "test"
    wl = bench.make_workload("tiny")
    t, S, cad = wl["lc"][:, 0], wl["S"], wl["cad"]
    off = cad * np.arange(0, 1.0, 1 / S)
    m_int = m_edge = t_in = 0
    for tcen, b, vel, rp, _, _ in wl["pos"]:
        z = np.sqrt(b * b + (vel * ((t[:, None] + off[None, :]) - tcen)) ** 2)
        inner = z < abs(1 - abs(rp))
        limb = (z < 1 + abs(rp)) & ~inner
        m_int += int(inner.sum()); m_edge += int(limb.sum()); t_in += int(np.any(inner | limb, axis=1).sum())
    flops, fl_sweep, fl_eval, counts = bench.algorithmic_flops(wl)
    m, c_int, c_edge, c_in = counts
    assert (c_int, c_edge, c_in) == (m_int, m_edge, t_in) and m_int > 0 and m_edge > 0
    wn = wl["W"] * wl["N"]
    assert m == wn * S
    assert fl_sweep == 7 * m + 5 * (wn - t_in)
    assert fl_eval == 143 * m_int + 180 * m_edge + 5 * t_in + 40 * wl["W"]
    assert flops == fl_sweep + fl_eval
    ex_sweep, ex_eval = bench.executed_fp64(wl, counts)
    assert ex_sweep == m + wn // 16    # one DFMA per fine sample + one DADD per 16-timestamp trip
    assert ex_eval == 210 * (m_int + m_edge) + 40 * t_in


def test_reference_and_gpu_arms_print_the_same_config():
    import bench
    wl = bench.make_workload("k2")
    a = bench.workload_config("k2", wl, 1)
    b = bench.workload_config("k2", bench.make_workload("k2"), 1)
    assert a == b and a["workload"].startswith("BASELINE configs[1]")
    assert bench.pick_workload(type("A", (), {"workload": "auto", "gpus": 1})()) == "kepler"
    assert bench.pick_workload(type("A", (), {"workload": "auto", "gpus": 8})()) == "batch"
