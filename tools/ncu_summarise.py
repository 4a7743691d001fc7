"""Turn ncu outputs into the small text/JSON summaries committed under profiles/.
  python tools/ncu_summarise.py launches <launches.csv>        -> per-kernel total time and share
  python tools/ncu_summarise.py full <report.ncu-rep>          -> per-kernel key metrics (time, DRAM bytes, pipes)"""
import csv
import io
import json
import re
import subprocess
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"void ", "", name)
    return name[:70]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    ui = h.index("Metric Unit")
    tot = defaultdict(lambda: [0.0, 0])
    for r in rows[hdr + 1:]:
        if len(r) != len(h):
            continue
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui].replace("second", "s").replace("usecond", "us").strip(), None)
        if scale is None:
            scale = {"nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(r[ui], 1e-6)
        tot[short(r[ki])][0] += v * scale
        tot[short(r[ki])][1] += 1
    total = sum(v[0] for v in tot.values())
    print(f"# {path}: {sum(v[1] for v in tot.values())} launches, {total:.3f} ms of kernel time (cold-cache, serialised: compare SHARES)")
    for k, (ms, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
        print(f"{ms / total * 100:6.2f}%  {ms:9.3f} ms  {n:5d}x  {k}")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__block_size",
        "sm__inst_executed_pipe_uniform.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    idx = {n: i for i, n in enumerate(h)}
    res = []
    for r in rows[2:]:
        d = {"kernel": short(r[idx["Kernel Name"]]), "grid": r[idx.get("Grid Size", 0)]}
        for w in WANT:
            if w in idx:
                d[w] = f"{r[idx[w]]} {units[idx[w]]}".strip()
        # warp-state statistics: cycles a warp spends stalled per issued instruction, by reason (top reasons only)
        stalls = {}
        for n, i in idx.items():
            m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", n)
            if m:
                try:
                    stalls[m.group(1)] = round(float(r[i].replace(",", "")), 3)
                except ValueError:
                    pass
        if stalls:
            d["stall_cycles_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
        for n in ("smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
                  "smsp__issue_inst0.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum",
                  "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum"):
            if n in idx:
                d[n] = f"{r[idx[n]]} {units[idx[n]]}".strip()
        res.append(d)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
