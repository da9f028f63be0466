"""SASS opcode histogram per kernel of the built library (cuobjdump -sass; runs without a GPU).
Writes the evidence of Blackwell-native data movement and of the FP64 instruction mix that DESIGN.md cites:
UBLKCP / SYNCS (TMA bulk copy + mbarrier), MUFU.RCP64H / MUFU.RSQ64H seeds, DFMA / DMUL / DADD counts.
usage: python tools/sass_histogram.py [pattern ...] > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "namaste_b200", "libnamaste_b200.so")
DEFAULT = ["sweep_walkers_kernelILi15ELi128", "sweep_walkers_kernelILi11ELi128", "eval_slots_kernelILi15ELi128ELi5ELi3",
           "eval_mixed_kernelILi11ELi128ELi5ELi3", "window_direct_kernelILi11ELi128ELi4ELi3", "window_ranges_kernel",
           "split_exchange_kernel", "ensemble_persistent"]
KEY = ["UBLKCP", "SYNCS", "UTMALDG", "MUFU.RCP64H", "MUFU.RSQ64H", "DFMA", "DMUL", "DADD", "DSETP", "VIMNMX", "LDS", "LDG",
       "STG", "ATOMG", "RED", "UMOV", "IMAD.MOV", "ACQBULK", "BAR", "NANOSLEEP", "CCTL", "MEMBAR", "ERRBAR"]


def main():
    pats = sys.argv[1:] or DEFAULT
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, name = {}, None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and name:
            funcs[name][m.group(1)] += 1
    for pat in pats:
        for fn, cnt in sorted(funcs.items()):
            if pat not in fn:
                continue
            total = sum(cnt.values())
            base = collections.Counter()
            for op, c in cnt.items():
                base[op.split(".")[0]] += c
            print("== %s\n   %d instructions (static)" % (fn, total))
            keys = []
            for k in KEY:
                c = sum(v for op, v in cnt.items() if op == k or op.startswith(k + ".") or (("." not in k) and op.split(".")[0] == k))
                if c:
                    keys.append("%s %d" % (k, c))
            print("   key opcodes: " + ", ".join(keys))
            print("   top: " + ", ".join("%s %d" % (op, c) for op, c in base.most_common(14)))


if __name__ == "__main__":
    main()
