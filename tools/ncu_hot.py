"""Summarise `ncu --page source --csv` output: hottest SASS instructions by executed count.
usage: python tools/ncu_hot.py <src.csv> [min_pct]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
hdr = next(r for r in rows if "Instructions Executed" in r)
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows:
    if len(r) != len(hdr) or r is hdr:
        continue
    try:
        n = int(r[ix["Instructions Executed"]])
    except ValueError:
        continue
    data.append((r, n))
tot = sum(n for _, n in data)
samp = sum(int(r[ix["# Samples"]] or 0) for r, _ in data)
print("total warp-instr", tot, "SASS lines", len(data), "samples", samp)
for r, n in data:
    if n >= minpct / 100 * tot or int(r[ix["# Samples"]] or 0) >= minpct / 100 * samp:
        print(f"{r[ix['Address']][-5:]} {n / tot * 100:5.2f}% thr={r[ix['Avg. Threads Executed']]:>5s} "
              f"samp={int(r[ix['# Samples']] or 0) / max(samp, 1) * 100:5.2f}%  {r[ix['Source']][:100]}")
