#!/bin/bash
# Round-2 baseline diagnostics (run through gpurun): per-tile phase split of C3 / C5, C5 timing, one full ncu capture of the C5 geometry kernel.
OUT=gpurun_out
mkdir -p $OUT
python tools/tile_stats.py c3 64 > $OUT/r02a_tilestats_c3_64.txt 2>&1
python tools/tile_stats.py c3 32 > $OUT/r02a_tilestats_c3_32.txt 2>&1
python tools/tile_stats.py c5 64 > $OUT/r02a_tilestats_c5_64.txt 2>&1
python tools/gpu_time.py c5 c2 c4l c4p --tile 64 > $OUT/r02a_time.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:geometryKernel -s 1 -c 1 -o $OUT/r02a_geom_c5 -f \
    python tools/prof_run.py c5 64 > $OUT/r02a_ncu_geom_c5.log 2>&1
tail -3 $OUT/r02a_time.txt
