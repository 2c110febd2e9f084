"""Summarise an `ncu --page source --csv` export: executed-instruction histogram by opcode
and the hottest SASS lines per kernel section.

    python tools/ncu_source_top.py gpurun_out/x_source.csv [section_index] [top_n]
"""
import csv
import sys
from collections import Counter

path = sys.argv[1]
only = int(sys.argv[2]) if len(sys.argv) > 2 else None
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
csv.field_size_limit(1 << 30)
sections = []
with open(path, newline="") as fh:
    cur = None
    for row in csv.reader(fh):
        if not row:
            continue
        if row[0] == "Kernel Name":
            cur = {"name": row[1], "hdr": None, "rows": []}
            sections.append(cur)
        elif row[0] == "Address":
            cur["hdr"] = row
        elif cur is not None and cur["hdr"] is not None:
            cur["rows"].append(row)
for si, sec in enumerate(sections):
    if only is not None and si != only:
        continue
    h = {n: i for i, n in enumerate(sec["hdr"])}
    ie = h["Instructions Executed"]
    total = 0
    ops = Counter()
    lines = []
    for r in sec["rows"]:
        try:
            n = int(float(r[ie]))
        except (ValueError, IndexError):
            continue
        src = r[h["Source"]].strip()
        op = src.split()[0] if src else "?"
        if op.startswith("@"):
            op = src.split()[1] if len(src.split()) > 1 else op
        op = op.split(".")[0]
        ops[op] += n
        total += n
        lines.append((n, r[h["Address"]], src, r[h["# Samples"]] if "# Samples" in h else ""))
    print(f"== section {si}: {sec['name'][:80]}  total warp-instructions {total:.4g}, {len(sec['rows'])} SASS lines")
    for op, n in ops.most_common(22):
        print(f"   {op:10s} {n:14d} {100.0 * n / max(total, 1):6.2f}%")
    if only is not None:
        print("   hottest lines:")
        for n, addr, src, samples in sorted(lines, reverse=True)[:top_n]:
            print(f"   {n:12d} smp={samples:>6s} {addr} {src[:110]}")
