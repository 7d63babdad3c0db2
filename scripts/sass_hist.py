#!/usr/bin/env python3
"""Opcode histogram of one kernel of a library: sass_hist.py lib.so <substring of the mangled name> [top]"""
import collections, re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
cur, fn = None, {}
for line in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m: cur = m.group(1); fn[cur] = []; continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if m and cur: fn[cur].append(m.group(1))
for name, ins in fn.items():
    if sys.argv[2] not in name: continue
    c = collections.Counter()
    for i in ins:
        p = i.split()
        op = p[1] if p[0].startswith("@") and len(p) > 1 else p[0]
        c[op.split(".")[0]] += 1
    print(name, len(ins), "instructions")
    print("  " + "  ".join(f"{k}:{v}" for k, v in c.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 24)))
