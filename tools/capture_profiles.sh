#!/bin/bash
# Captures the ncu evidence of profiles/ on a GPU box (run through gpurun from the repo root):
#   launch list of the bench command, one `--set full` capture of each kernel, raw pages as csv.
# usage: bash tools/capture_profiles.sh [tag]      (outputs under gpurun_out/)
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tileKernel -s 1 -c 1 -o $OUT/prof_tile_c3_$TAG -f \
    python tools/prof_run.py c3 64 > $OUT/ncu_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:geometryKernel -s 1 -c 1 -o $OUT/prof_geom_c3_$TAG -f \
    python tools/prof_run.py c3 64 >> $OUT/ncu_$TAG.log 2>&1
for k in tile geom; do
    ncu -i $OUT/prof_${k}_c3_$TAG.ncu-rep --page raw --csv > $OUT/prof_${k}_c3_${TAG}_raw.csv 2>/dev/null
done
tail -2 $OUT/ncu_$TAG.log
