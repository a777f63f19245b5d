#!/bin/bash
# Final round-2 evidence on one B200: parity suite, the contract's bench line, ncu launch list of the bench command,
# one `--set full` capture of the tile and the geometry kernel on C3 (64-pixel tiles) and C2 (32).
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/r02h_pytest.txt 2>&1; tail -2 $OUT/r02h_pytest.txt
python bench.py > $OUT/r02_bench_c3_n1.json 2> $OUT/r02_bench_c3_n1.err; python -c "
import json; d=json.loads(open('$OUT/r02_bench_c3_n1.json').read().strip().splitlines()[-1]); print('bench c3: ms/step %.3f value %.4g e2e %.3f ms frac %.4f geom %.3f tile %.3f cpu %.3g' % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['geometry_kernel_ms'], d['roofline']['kernel_ms'], d['cpu_baseline']['value']))"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/r02_launches_bench_c3.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > /dev/null 2>&1
for spec in c3:64 c2:32; do
  W=${spec%%:*}; T=${spec##*:}
  for K in tile geometry; do
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:${K}Kernel -s 1 -c 1 -o $OUT/r02_${K}_${W} -f \
        python tools/prof_run.py $W $T > $OUT/r02_ncu_${K}_${W}.log 2>&1
    tail -1 $OUT/r02_ncu_${K}_${W}.log
    ncu -i $OUT/r02_${K}_${W}.ncu-rep --page raw --csv > $OUT/r02_${K}_${W}_raw.csv 2>/dev/null
    python tools/ncu_lines.py $OUT/r02_${K}_${W}.ncu-rep 50 > $OUT/r02_${K}_${W}_lines.txt 2>/dev/null
    if [ "$W" != "c3" ]; then rm -f $OUT/r02_${K}_${W}.ncu-rep; fi
    rm -f $OUT/r02_ncu_${K}_${W}.log
  done
done
python bench.py --workload c2 --no-cpu > $OUT/r02_bench_c2_n1.json 2> /dev/null
