#!/usr/bin/env python3
"""Attribute the per-SASS-instruction counters of an ncu report to CUDA source lines.
usage: sass_by_line.py <source.csv from `ncu -i rep --page source --csv`> <nvdisasm -g dump> <mangled kernel name> [top] [which]
The n-th instruction of the kernel in both listings is the same instruction (same cubin), so the join is positional."""
import collections
import csv
import re
import sys

src_csv, dump, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
which = int(sys.argv[5]) if len(sys.argv) > 5 else 0
rows = list(csv.reader(open(src_csv)))
starts = [i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r]
hdr = rows[starts[which]]
end = starts[which + 1] - 1 if which + 1 < len(starts) else len(rows)
data = [r for r in rows[starts[which] + 1:end] if len(r) == len(hdr)]
iE, iT, iP = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
lines, cur, inside = [], ("?", 0), False
for l in open(dump, errors="replace"):
    if l.startswith(".text."):
        inside = l.strip().rstrip(":") == ".text." + kernel
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
assert len(lines) == len(data), (len(lines), len(data))
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for loc, r in zip(lines, data):
    a = agg[loc]
    a[0] += int(r[iE]); a[1] += int(r[iT]); a[2] += int(r[iP]); a[3] += 1
tot, totp = sum(a[0] for a in agg.values()), sum(a[2] for a in agg.values())
print(f"{len(data)} sass instructions, {tot} warp instructions, {totp} samples")
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{a[0] / tot * 100:5.1f}% instr {a[2] / totp * 100:5.1f}% samples  lanes {a[1] / max(a[0], 1):4.1f}  sass {a[3]:4d}  {loc[0]}:{loc[1]}")
