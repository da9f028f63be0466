"""Top CUDA source lines by warp-stall samples from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`.
usage: ncu_cuda_lines.py <csv> [top N]"""
import csv, sys, os
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = os.path.basename(r[1]); hdr = None; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) != len(hdr) or r[0] == "":
        continue
    d = dict(zip(hdr, r))
    # the header has two "Source" columns; dict keeps the last (SASS); first is r[1]
    stalls = {k: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}
    out.append((int(d["# Samples"]), int(d["Instructions Executed"]), cur_file, r[0], r[1].strip()[:90], stalls))
tot = sum(o[0] for o in out); toti = sum(o[1] for o in out)
print("total samples %d, warp instructions %d" % (tot, toti))
for smp, inst, f, ln, src, st in sorted(out, key=lambda x: -x[0])[:top]:
    s3 = ", ".join("%s %d" % (k[6:], v) for k, v in sorted(st.items(), key=lambda x: -x[1])[:3])
    print("%5.1f%% smp %5.1f%% inst  %s:%s  %s\n        [%s]" % (100.0 * smp / max(tot, 1), 100.0 * inst / max(toti, 1), f, ln, src, s3))
