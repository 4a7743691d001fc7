mkdir -p gpurun_out
(python -m pytest tests/test_raster.py tests/test_legacy.py -m gpu -q -x 2>&1 | tail -8) > gpurun_out/r02_tests3.log 2>&1
python tools/time_raster.py >> gpurun_out/r02_time_raster2.jsonl 2>>gpurun_out/r02_time_raster2.err
ADB_STEPS=2 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"blend_fwd|blend_bwd|tile_sort|tile_scatter|tile_count|tile_scan" -c 12 -f -o /tmp/raster_full python tools/prof_step.py > /dev/null 2>&1
python tools/ncu_summarise.py full /tmp/raster_full.ncu-rep > gpurun_out/r02_ncu_full_raster.json 2>&1
ncu -i /tmp/raster_full.ncu-rep --page source --csv --print-source sass -k regex:blend_bwd -c 1 > /tmp/bwd_src.csv 2>/dev/null
python tools/ncu_hot.py /tmp/bwd_src.csv 0.7 > gpurun_out/r02_ncu_hot_blend_bwd.txt 2>&1
ncu -i /tmp/raster_full.ncu-rep --page source --csv --print-source sass -k regex:blend_fwd -c 1 > /tmp/fwd_src.csv 2>/dev/null
python tools/ncu_hot.py /tmp/fwd_src.csv 0.7 > gpurun_out/r02_ncu_hot_blend_fwd.txt 2>&1
ncu -i /tmp/raster_full.ncu-rep --page source --csv --print-source cuda -k regex:blend_bwd -c 1 > /tmp/bwd_lines.csv 2>/dev/null
python tools/ncu_lines.py /tmp/bwd_lines.csv 1.0 > gpurun_out/r02_ncu_lines_blend_bwd.txt 2>&1
ncu -i /tmp/raster_full.ncu-rep --page source --csv --print-source cuda -k regex:blend_fwd -c 1 > /tmp/fwd_lines.csv 2>/dev/null
python tools/ncu_lines.py /tmp/fwd_lines.csv 1.0 > gpurun_out/r02_ncu_lines_blend_fwd.txt 2>&1
cat gpurun_out/r02_tests3.log; cat gpurun_out/r02_time_raster2.jsonl; ls -la gpurun_out | tail -12
