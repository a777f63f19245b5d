#!/bin/bash
# Group compaction (single-destination geometry) + one-thread-per-group record binning: parity, then A/B timing.
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/r02d_pytest.txt 2>&1; tail -3 $OUT/r02d_pytest.txt
for v in "" _a _b; do
  SWR_LIB_VARIANT=$v python tools/gpu_time.py c3 c2 fill > $OUT/r02d_time$v.txt 2>&1
  echo "variant '$v'"; grep TIME $OUT/r02d_time$v.txt
  SWR_LIB_VARIANT=$v python tools/gpu_time.py c5 --tile 64 >> $OUT/r02d_time$v.txt 2>&1; grep "TIME c5" $OUT/r02d_time$v.txt
done
