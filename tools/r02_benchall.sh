#!/bin/bash
# bench.py lines of every workload at N = 1 (gpurun_out/r02_bench_<workload>_n1.json) + the reference arm of c3.
OUT=gpurun_out
mkdir -p $OUT
python bench.py > $OUT/r02_bench_c3_n1.json 2> $OUT/r02_bench_c3_n1.err; tail -c 600 $OUT/r02_bench_c3_n1.json
for w in c2 c5 fill c0_4k c0; do
  python bench.py --workload $w --steps 10 --no-cpu > $OUT/r02_bench_${w}_n1.json 2> $OUT/r02_bench_${w}_n1.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/r02_bench_${w}_n1.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("$w", "ms/step %.3f"%d["ms_per_step"], "value %.3g"%d["value"], "tile", d["config"]["tile_size"], "tile_ms %.3f"%r["kernel_ms"], "geom_ms %.3f"%r["geometry_kernel_ms"], "frac %.4f"%r["frac"], "e2e %.3f"%d["e2e"]["ms_per_step"])
except Exception as e:
    print("$w failed", e); print(open("$OUT/r02_bench_${w}_n1.err").read()[-1200:])
PY
done
python bench.py --impl reference --steps 1 --warmup 0 > $OUT/r02_bench_c3_reference_arm.json 2>/dev/null; tail -c 300 $OUT/r02_bench_c3_reference_arm.json
