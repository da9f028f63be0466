"""Turn the files a `tools/gpu_round.sh` visit left in gpurun_out/ into the tracked summaries under
profiles/ (per-kernel ncu summaries, launch list, bench lines, measured DRAM traffic).
usage: python tools/make_profiles.py <tag, e.g. r01d>"""
import csv, io, json, os, shutil, subprocess, sys
tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, pr = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
os.makedirs(os.path.join(pr, "bench"), exist_ok=True)


def run(*cmd):
    return subprocess.run(cmd, capture_output=True, text=True).stdout


def traffic(rep):
    rows = list(csv.reader(io.StringIO(run("ncu", "-i", rep, "--page", "raw", "--csv"))))
    h, u = rows[0], rows[1]
    ir, iw = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    vals = [float(r[ir].replace(",", "")) * mult[u[ir]] + float(r[iw].replace(",", "")) * mult[u[iw]] for r in rows[2:]]
    return sum(vals) / len(vals)


tr = {"k2": {}, "kepler": {}}
for k in ("sweep_walkers", "eval_mixed", "eval_slots", "window_ranges", "window_direct", "sweep_walkers_kepler"):
    rep = os.path.join(go, "prof_%s.ncu-rep" % k)
    if os.path.exists(rep + ".gz"):
        subprocess.run(["gunzip", "-f", rep + ".gz"], check=True)
    if not os.path.exists(rep):
        continue
    sass = os.path.join(go, k + "_sass.csv")
    open(sass, "w").write(run("ncu", "-i", rep, "--page", "source", "--csv"))
    txt = run(sys.executable, os.path.join(root, "tools", "ncu_raw_summary.py"), rep, "1")
    txt += "\n".join(run(sys.executable, os.path.join(root, "tools", "ncu_sass_summary.py"), sass).splitlines()[:30]) + "\n"
    open(os.path.join(pr, "%s_%s_summary.txt" % (tag, k)), "w").write(txt)
    wl, kern = ("kepler", "sweep_walkers_kernel") if k.endswith("_kepler") else ("k2", k + "_kernel")
    tr[wl][kern] = traffic(rep)
tr["_source"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch from the %s ncu --set full captures (profiles/%s_*_summary.txt)" % (tag, tag)
json.dump(tr, open(os.path.join(pr, "traffic.json"), "w"), indent=1)
open(os.path.join(pr, tag + "_launches.csv"), "w").writelines(l for l in open(os.path.join(go, "launches.csv")) if l.startswith('"'))
for w in ("k2", "kepler", "tess", "epic", "ref"):
    src = os.path.join(go, "bench_%s.json" % w)
    if os.path.exists(src):
        shutil.copy(src, os.path.join(pr, "bench", "%s_bench_%s.json" % (tag, w)))
print(json.dumps(tr))
