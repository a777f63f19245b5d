#!/bin/bash
# Narrowed index upload: parity of the streamed paths, then the end-to-end leg off / forced / automatic.
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q -k "streamed or host_attribute or error_codes" > $OUT/r02i_pytest.txt 2>&1; tail -3 $OUT/r02i_pytest.txt
for mode in 0 1 auto; do
  if [ $mode = auto ]; then unset SWR_INDEX_NARROWING; else export SWR_INDEX_NARROWING=$mode; fi
  python bench.py --no-cpu --steps 20 > $OUT/r02i_bench_$mode.json 2> $OUT/r02i_bench_$mode.err
  python -c "
import json; d=json.loads(open('$OUT/r02i_bench_$mode.json').read().strip().splitlines()[-1]); e=d['e2e']
print('narrowing $mode: ms/step %.3f  e2e %.3f ms  host enqueue %.3f ms  h2d %.1f MB of %.1f MB' % (d['ms_per_step'], e['ms_per_step'], e['host_enqueue_ms_per_step'], e['h2d_bytes_per_step']/1e6, e['input_bytes_per_step']/1e6))"
done
nproc; grep -m1 "model name" /proc/cpuinfo
