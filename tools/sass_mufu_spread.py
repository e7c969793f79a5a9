"""How evenly are the MUFU instructions spread over the hot loop of a kernel?

    cuobjdump -sass -fun <mangled name> lib.so | python tools/sass_mufu_spread.py [window]

Prints the histogram of the number of MUFU instructions per window of W consecutive instructions of the hot loop (the
smallest backward-branch span that holds >= 80 % of the function's MUFU), and the longest stretches without / made of MUFU.
A warp issues in order: a window with more MUFU than the pipe takes (one per 8 cycles and scheduler) blocks it.
"""
import re
import sys
from collections import Counter

W = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ins = []
for line in sys.stdin:
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
loops = []
for addr, text in ins:
    m = re.search(r"\bBRA(?:\.\w+)*\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", text)
    if m and int(m.group(1), 16) < addr:
        loops.append((int(m.group(1), 16), addr))
total = sum(1 for a, t in ins if "MUFU." in t)
cands = [(hi - lo, lo, hi) for lo, hi in loops if sum(1 for a, t in ins if lo <= a <= hi and "MUFU." in t) >= 0.8 * total]
_, lo, hi = min(cands)
body = [t for a, t in ins if lo <= a <= hi]
is_m = [1 if "MUFU." in t else 0 for t in body]
n = len(body)
hist = Counter(sum(is_m[i:i + W]) for i in range(0, n - W + 1))
print(f"loop: {n} instructions, {sum(is_m)} MUFU ({sum(is_m) / n * W:.2f} per {W}-instruction window if evenly spread)")
for k in sorted(hist):
    print(f"  {k:2d} MUFU in window: {hist[k] / (n - W + 1) * 100:5.1f} %")
# demand profile: cycles the MUFU pipe needs (8 per MUFU) minus issue slots elapsed, as a running backlog
backlog, peak, blocked = 0.0, 0.0, 0.0
for m_ in is_m:
    backlog = max(0.0, backlog - 1.0)          # one issue slot passes
    if m_:
        backlog += 8.0
    peak = max(peak, backlog)
    blocked += max(0.0, backlog - 8.0 * 6) / 1  # beyond a 6-entry queue the warp would wait
print(f"single-warp MUFU backlog (cycles of pipe work queued, one instruction per cycle otherwise): peak {peak:.0f}")
runs, cur = [], 0
for m_ in is_m:
    if m_:
        if cur:
            runs.append(cur)
        cur = 0
    else:
        cur += 1
print(f"MUFU-free stretches: longest {max(runs)}, mean {sum(runs) / len(runs):.1f} instructions")
