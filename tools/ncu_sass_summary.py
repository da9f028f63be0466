"""Summarise an `ncu --page source --csv` dump: stall reasons + opcode mix per kernel."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
# file holds one block per kernel launch: a "Kernel Name" row, a header row, then SASS rows
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}; blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
b = blocks[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
hdr, data = b["hdr"], b["data"]
idx = {h: i for i, h in enumerate(hdr)}
print(b["name"][:100], "launch blocks:", len(blocks))
tot_samples = sum(int(r[idx['# Samples']]) for r in data)
tot_inst = sum(int(r[idx['Instructions Executed']]) for r in data)
print("total samples", tot_samples, "total warp inst", tot_inst, "n sass", len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(int(r[idx[s]]) for r in data) for s in stalls}
for s, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
    print("  %-24s %8d %5.1f%%" % (s, v, 100 * v / max(tot_samples, 1)))
op = collections.Counter(); ops = collections.Counter()
for r in data:
    parts = r[idx['Source']].split()
    o = parts[1] if parts[0].startswith('@') else parts[0]
    o = o.split('.')[0]
    op[o] += int(r[idx['Instructions Executed']]); ops[o] += int(r[idx['# Samples']])
print("opcode: inst_exec share, stall-sample share")
for o, v in op.most_common(24):
    print("  %-10s %10d %5.1f%%  %5.1f%%" % (o, v, 100 * v / tot_inst, 100 * ops[o] / max(tot_samples, 1)))
