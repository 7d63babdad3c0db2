#!/usr/bin/env python3
"""Instruction mix of one kernel launch from `ncu -i rep --page source --csv` output: share of issued warp instructions,
share of stall samples and mean active lanes per opcode.  usage: sass_mix.py source.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # n-th kernel section of the file
starts = [i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r]
hdr = rows[starts[which]]
end = starts[which + 1] - 1 if which + 1 < len(starts) else len(rows)
print(rows[starts[which] - 1][1][:60] if starts[which] > 0 else "")
data = [r for r in rows[starts[which] + 1:end] if len(r) == len(hdr)]
iS, iE, iT, iP = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[iE]) for r in data); totS = sum(int(r[iP]) for r in data)
print("sass lines", len(data), "warp-instr", tot, "samples", totS)
ops, opsS, opsT = collections.Counter(), collections.Counter(), collections.Counter()
for r in data:
    p = r[iS].strip().split()
    op = (p[1] if p[0].startswith("@") else p[0]).split(".")[0]
    ops[op] += int(r[iE]); opsS[op] += int(r[iP]); opsT[op] += int(r[iT])
for op, c in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"{op:10s} {c / tot * 100:6.2f}% instr  {opsS[op] / totS * 100:6.2f}% samples  avg lanes {opsT[op] / max(c, 1):5.1f}")
