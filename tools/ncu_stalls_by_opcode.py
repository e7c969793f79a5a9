"""Aggregate the per-instruction warp-stall samples of an `ncu --page source --csv --print-source sass` dump by opcode.

    ncu -i report.ncu-rep --page source --csv --print-source sass > src.csv
    python tools/ncu_stalls_by_opcode.py src.csv [top_n]

A sample taken at an instruction means: a warp was waiting TO ISSUE that instruction (for the reason given).
"""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
reasons = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" not in h]
by_op = defaultdict(lambda: defaultdict(int))
count = defaultdict(int)
total = 0
lines = []
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    text = re.sub(r"^@!?U?P\d+\s+", "", r[col["Source"]].strip())
    op = text.split()[0] if text else "?"
    parts = op.split(".")
    key = parts[0] if parts[0] not in ("MUFU", "LDS", "STS", "SYNCS", "LDTM", "STTM") else ".".join(parts[:2])
    n = int(r[col["# Samples"]] or 0)
    count[key] += int(r[col["Instructions Executed"]] or 0)
    total += n
    for h in reasons:
        by_op[key][h] += int(r[col[h]] or 0)
    lines.append((n, r[col["Address"]], text, {h: int(r[col[h]] or 0) for h in reasons}))
print(f"total samples {total}")
print(f"{'opcode':16s} {'executed':>12s} {'samples':>9s} {'share':>7s}  top reasons")
tot_reason = defaultdict(int)
for key, d in sorted(by_op.items(), key=lambda kv: -sum(kv[1].values())):
    s = sum(d.values())
    if s == 0:
        continue
    for h, v in d.items():
        tot_reason[h] += v
    top = ", ".join(f"{h[6:]} {v / s:.0%}" for h, v in sorted(d.items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"{key:16s} {count[key]:12d} {s:9d} {s / max(total, 1):7.1%}  {top}")
print("by reason: " + ", ".join(f"{h[6:]} {v / max(total, 1):.1%}" for h, v in sorted(tot_reason.items(), key=lambda kv: -kv[1]) if v))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for n, addr, text, d in sorted(lines, key=lambda t: -t[0])[:top_n]:
    top = ", ".join(f"{h[6:]} {v}" for h, v in sorted(d.items(), key=lambda kv: -kv[1])[:2] if v)
    print(f"  {n:6d}  {addr[-6:]}  {text[:70]:70s} {top}")
