#!/usr/bin/env bash
# Run under gpurun on ONE GPU:   gpurun --timeout 600 -- 'bash tools/profile_raster_final.sh r02'
# Rasterizer half of tools/profile_on_box.sh (launch list of one 8-view step, `--set full` of one view's kernels, hot SASS lines
# of the two blend kernels) plus the COMPLETE per-SASS-line execution histogram of both blend kernels as compact text.
set -u
TAG="${1:-rXX}"
OUT=gpurun_out
mkdir -p "$OUT"
NCU="ncu --clock-control none"
export ADB_STEPS=1 ADB_CUPROF=1
timeout 300 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file /tmp/launches_raster.csv \
    python tools/prof_multiview.py > /dev/null 2>&1
python tools/ncu_summarise.py launches /tmp/launches_raster.csv > "$OUT/${TAG}_launches_raster.txt" 2>&1
export ADB_VIEWS=1
timeout 300 $NCU --profile-from-start off --set full --import-source on -f -o /tmp/raster_full \
    python tools/prof_multiview.py > /dev/null 2>&1
python tools/ncu_summarise.py full /tmp/raster_full.ncu-rep > "$OUT/${TAG}_ncu_full_raster.json" 2>&1
for k in blend_fwd blend_bwd; do
    ncu -i /tmp/raster_full.ncu-rep --page source --csv --print-source sass -k regex:$k -c 1 > /tmp/${k}_src.csv 2>/dev/null
    python tools/ncu_hot.py /tmp/${k}_src.csv 0.6 > "$OUT/${TAG}_ncu_hot_${k}.txt" 2>&1
    python tools/ncu_hot.py /tmp/${k}_src.csv 0.0 > "$OUT/${TAG}_ncu_sass_hist_${k}.txt" 2>&1
done
rm -f /tmp/*.ncu-rep
ls -la "$OUT" | tail -8
