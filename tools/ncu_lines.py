"""Attribute ncu SASS-level stall samples / executed instructions to CUDA source lines.

usage: ncu_lines.py <sass.csv from `ncu --page source --csv`> <demangled substring> <mangled substring> [launch index]
Joins the i-th SASS instruction of the kernel in the ncu dump with the i-th instruction of
`nvdisasm -g` on the cubin extracted from the library (same code, same order)."""
import csv, collections, os, re, subprocess, sys, tempfile
sass_csv, dname, kname = sys.argv[1], sys.argv[2], sys.argv[3]
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "namaste_b200", "libnamaste_b200.so")
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rows = list(csv.reader(open(sass_csv)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}; blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
blocks = [b for b in blocks if dname in b["name"]]
b = blocks[which]
idx = {h: i for i, h in enumerate(b["hdr"])}
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-sf", cubin], capture_output=True, text=True).stdout.splitlines()
# find the .text section of the kernel
start = None
for i, l in enumerate(dis):
    if l.lstrip().startswith(".section") and ".text." in l and kname in l:
        start = i
        break
if start is None:
    cands = [l for l in dis if l.startswith(".section") and ".text." in l]
    raise SystemExit("kernel section not found; have: " + "\n".join(c[:120] for c in cands[:40]))
line_of = []
curline = ("?", 0)
for l in dis[start + 1:]:
    if l.lstrip().startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        curline = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        line_of.append(curline)
data = b["data"]
print("ncu sass rows", len(data), "nvdisasm instr", len(line_of))
n = min(len(data), len(line_of))
samp = collections.Counter(); inst = collections.Counter()
for i in range(n):
    samp[line_of[i]] += int(data[i][idx["# Samples"]]); inst[line_of[i]] += int(data[i][idx["Instructions Executed"]])
ts, ti = sum(samp.values()), sum(inst.values())
src_cache = {}
def src(f, ln):
    for d in ("namaste_b200/csrc",):
        pth = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, f)
        if os.path.exists(pth):
            if pth not in src_cache: src_cache[pth] = open(pth).read().splitlines()
            L = src_cache[pth]
            return L[ln - 1].strip()[:90] if 0 < ln <= len(L) else ""
    return ""
print("%-28s %7s %7s   source" % ("file:line", "stall%", "inst%"))
for (f, ln), v in samp.most_common(28):
    print("%-28s %6.1f%% %6.1f%%   %s" % ("%s:%d" % (f, ln), 100 * v / max(ts, 1), 100 * inst[(f, ln)] / max(ti, 1), src(f, ln)))
