"""Opcode histogram of the hot loop of a kernel from `cuobjdump -sass` text.

    cuobjdump -sass -fun <mangled name> lib.so | python tools/sass_loop_hist.py [edges_per_iteration]

The hot loop is taken to be the smallest backward branch span holding >= 80 % of the MUFU.EX2/LG2 of the function.  With an edge count the
histogram is also printed per (edge) -- divide by 2 yourself for a kernel that handles two frames per thread.
"""
import re
import sys
from collections import Counter

ins = []
for line in sys.stdin:
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
loops = []
for addr, text in ins:
    m = re.search(r"\bBRA(?:\.\w+)*\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", text)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < addr:
            loops.append((tgt, addr))
if not loops:
    sys.exit("no backward branch found")
# the hot loop = the SMALLEST backward span that still holds at least 80 % of the function's MUFU.LG2/EX2
total_mufu = sum(1 for a, t in ins if "MUFU.LG2" in t or "MUFU.EX2" in t)
cands = [(hi - lo, lo, hi) for lo, hi in loops
         if sum(1 for a, t in ins if lo <= a <= hi and ("MUFU.LG2" in t or "MUFU.EX2" in t)) >= 0.8 * total_mufu]
_, lo, hi = min(cands) if cands else max((hi - lo, lo, hi) for lo, hi in loops)
body = [t for a, t in ins if lo <= a <= hi]
hist = Counter()
for t in body:
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    op = t.split()[0]
    parts = op.split(".")
    key = parts[0]
    if key in ("MUFU", "LDS", "STS", "SYNCS", "IMAD", "LDG", "STG", "BAR", "ATOMS"):
        key = ".".join(parts[:2])
    hist[key] += 1
edges = float(sys.argv[1]) if len(sys.argv) > 1 else None
print(f"loop 0x{lo:x}..0x{hi:x}: {len(body)} instructions, {len(body) * 16 / 1024:.1f} KB")
for k, v in hist.most_common():
    print(f"  {k:20s} {v:6d}" + (f"   {v / edges:6.2f} per edge" if edges else ""))
if edges:
    print(f"  {'total':20s} {len(body):6d}   {len(body) / edges:6.2f} per edge")
