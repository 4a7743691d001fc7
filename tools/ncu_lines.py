"""Per-CUDA-source-line totals from `ncu --page source --csv --print-source cuda,sass`.
usage: python tools/ncu_lines.py <csv> [min_pct]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
h = rows[hi]
ci = h.index("Instructions Executed")
si = h.index("# Samples")
lines = []
for r in rows[hi + 1:]:
    if len(r) != len(h) or not r[0].strip().isdigit():
        continue
    try:
        lines.append((int(r[0]), r[1], int(r[ci]), int(r[si] or 0)))
    except ValueError:
        pass
tot = sum(x[2] for x in lines)
samp = sum(x[3] for x in lines)
print(f"total warp-instr {tot}  samples {samp}")
for ln, src, n, s in lines:
    if n >= minpct / 100 * tot or s >= minpct / 100 * samp:
        print(f"L{ln:<4d} instr {n / tot * 100:5.1f}%  stall-samples {s / max(samp,1) * 100:5.1f}%  {src.strip()[:105]}")
