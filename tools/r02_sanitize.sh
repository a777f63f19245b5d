#!/bin/bash
# compute-sanitizer over small draws of every kind (tools/sanitize_run.py) -> gpurun_out/r02_compute_sanitizer.txt
OUT=gpurun_out/r02_compute_sanitizer.txt
: > $OUT
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool python tools/sanitize_run.py" >> $OUT
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_run.py 2>&1 | grep -v "^$" | tail -40 >> $OUT
done
tail -5 $OUT
