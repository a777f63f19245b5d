#!/bin/bash
# Round-2 ncu evidence (run through gpurun from the repo root): the launch list of the bench command and one
# `--set full` capture of the tile and the geometry kernel per workload, at the tile size bench.py uses.  The reports
# are condensed on the box (raw page as csv + per-source-line summary); only the C3 reports travel back whole.
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/r02_launches_bench_c3.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > /dev/null 2>&1
for spec in c3:64 c2:32 c5:64 fill:32 c0_4k:32; do
  W=${spec%%:*}; T=${spec##*:}
  for K in tile geometry; do
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:${K}Kernel -s 1 -c 1 -o $OUT/r02_${K}_${W} -f \
        python tools/prof_run.py $W $T > $OUT/r02_ncu_${K}_${W}.log 2>&1
    tail -1 $OUT/r02_ncu_${K}_${W}.log
    ncu -i $OUT/r02_${K}_${W}.ncu-rep --page raw --csv > $OUT/r02_${K}_${W}_raw.csv 2>/dev/null
    python tools/ncu_lines.py $OUT/r02_${K}_${W}.ncu-rep 50 > $OUT/r02_${K}_${W}_lines.txt 2>/dev/null
    if [ "$W" != "c3" ]; then rm -f $OUT/r02_${K}_${W}.ncu-rep; fi
    rm -f $OUT/r02_ncu_${K}_${W}.log
  done
done
cuobjdump -sass softwarerenderer_b200/libswr_b200.so 2>/dev/null | grep -c "REDG\|RED\." > $OUT/r02_sass_red_count.txt
du -sh $OUT
