#!/usr/bin/env bash
# Run under gpurun on ONE GPU:   gpurun --timeout 1200 -- 'bash tools/profile_on_box.sh r02'
# Captures the launch lists and compact `--set full` summaries for the two bench legs and leaves ONLY small text/JSON files
# in gpurun_out/ (the .ncu-rep files stay in /tmp: a GPU job may copy back at most 64 MiB, and two full captures were lost
# to that cap in round 1).  Numbers printed by anything run under ncu are never bench values.
set -u
TAG="${1:-rXX}"
OUT=gpurun_out
mkdir -p "$OUT"
NCU="ncu --clock-control none"

# ---- rasterizer step: launch list + full set of the six kernels ----
timeout 300 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file /tmp/launches_raster.csv \
    python tools/prof_step.py > /dev/null 2>&1
python tools/ncu_summarise.py launches /tmp/launches_raster.csv > "$OUT/${TAG}_launches_raster.txt" 2>&1
timeout 300 $NCU --set full -k regex:"project_fwd|blend_fwd|blend_bwd|project_bwd" -c 4 -f -o /tmp/raster_full \
    python tools/prof_step.py > /dev/null 2>&1
python tools/ncu_summarise.py full /tmp/raster_full.ncu-rep > "$OUT/${TAG}_ncu_full_raster.json" 2>&1

# ---- MASt3R forward at the bench batch (B=4): launch list of ONE forward + full set of the first encoder-layer kernels ----
export ADB_B=4 ADB_REPS=2 ADB_GRAPH=0 ADB_CUPROF=1
timeout 400 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file /tmp/launches_mast3r.csv \
    python tools/prof_mast3r.py > /dev/null 2>&1
python tools/ncu_summarise.py launches /tmp/launches_mast3r.csv > "$OUT/${TAG}_launches_mast3r_b4.txt" 2>&1
timeout 300 $NCU --profile-from-start off --set full -k regex:"gemm_tc_kernel|attn_fused" -c 12 -f -o /tmp/mast3r_full \
    python tools/prof_mast3r.py > /dev/null 2>&1
python tools/ncu_summarise.py full /tmp/mast3r_full.ncu-rep > "$OUT/${TAG}_ncu_full_mast3r_b4.json" 2>&1
unset ADB_B ADB_REPS ADB_GRAPH ADB_CUPROF

# ---- attention alone (both kernel variants when ADB_ATTN_KERNEL=2 is being validated) ----
timeout 120 python tools/bench_attn.py > "$OUT/${TAG}_bench_attn.json" 2>/dev/null
ADB_SWEEP=1 timeout 120 python tools/bench_attn.py > "$OUT/${TAG}_bench_attn_sweep.json" 2>/dev/null

rm -f /tmp/*.ncu-rep
du -sh "$OUT"; ls -la "$OUT"
