#!/usr/bin/env bash
# Run under gpurun on ONE GPU:   gpurun --timeout 1500 -- 'bash tools/profile_on_box.sh r02'
# Captures the launch lists and compact `--set full` summaries for the two bench legs and leaves ONLY small text/JSON files
# in gpurun_out/ (the .ncu-rep files stay in /tmp: a GPU job may copy back at most 64 MiB).
# Numbers printed by anything run under ncu are never bench values.
set -u
TAG="${1:-rXX}"
OUT=gpurun_out
mkdir -p "$OUT"
NCU="ncu --clock-control none"

# ---- rasterizer: launch list of ONE 8-view optimiser step (the bench's step) + full set of every kernel on it ----
export ADB_STEPS=1 ADB_CUPROF=1
timeout 400 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file /tmp/launches_raster.csv \
    python tools/prof_multiview.py > /dev/null 2>&1
python tools/ncu_summarise.py launches /tmp/launches_raster.csv > "$OUT/${TAG}_launches_raster.txt" 2>&1
export ADB_VIEWS=1
timeout 400 $NCU --profile-from-start off --set full --import-source on -f -o /tmp/raster_full \
    python tools/prof_multiview.py > /dev/null 2>&1
python tools/ncu_summarise.py full /tmp/raster_full.ncu-rep > "$OUT/${TAG}_ncu_full_raster.json" 2>&1
for k in blend_fwd blend_bwd; do
    ncu -i /tmp/raster_full.ncu-rep --page source --csv --print-source sass -k regex:$k -c 1 > /tmp/${k}_src.csv 2>/dev/null
    python tools/ncu_hot.py /tmp/${k}_src.csv 0.6 > "$OUT/${TAG}_ncu_hot_${k}.txt" 2>&1
done
unset ADB_VIEWS ADB_STEPS ADB_CUPROF

# ---- MASt3R forward at the bench batch (B=4): launch list of ONE forward + full set of the first encoder-layer kernels ----
export ADB_B=4 ADB_REPS=2 ADB_GRAPH=0 ADB_CUPROF=1
timeout 400 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file /tmp/launches_mast3r.csv \
    python tools/prof_mast3r.py > /dev/null 2>&1
python tools/ncu_summarise.py launches /tmp/launches_mast3r.csv > "$OUT/${TAG}_launches_mast3r_b4.txt" 2>&1
timeout 300 $NCU --profile-from-start off --set full -k regex:"gemm_tc_kernel|attn_fused" -c 14 -f -o /tmp/mast3r_full \
    python tools/prof_mast3r.py > /dev/null 2>&1
python tools/ncu_summarise.py full /tmp/mast3r_full.ncu-rep > "$OUT/${TAG}_ncu_full_mast3r_b4.json" 2>&1
unset ADB_B ADB_REPS ADB_GRAPH ADB_CUPROF

# ---- attention alone: both kernel variants and the automatic choice ----
for v in 1 2 0; do
    ADB_ATTN_KERNEL=$v timeout 120 python tools/bench_attn.py > "$OUT/${TAG}_bench_attn_variant$v.json" 2>/dev/null
done

rm -f /tmp/*.ncu-rep
du -sh "$OUT"; ls -la "$OUT" | tail -15
