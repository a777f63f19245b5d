#!/bin/bash
OUT=gpurun_out
for t in 6 9 12 16; do
  SWR_PACK_THREADS=$t python bench.py --no-cpu --steps 20 > $OUT/r02k_bench_t$t.json 2> /dev/null
  python -c "
import json; d=json.loads(open('$OUT/r02k_bench_t$t.json').read().strip().splitlines()[-1]); e=d['e2e']
print('pack threads $t: e2e %.3f ms  host enqueue %.3f ms  h2d %.1f MB' % (e['ms_per_step'], e['host_enqueue_ms_per_step'], e['h2d_bytes_per_step']/1e6))"
done
