#!/bin/bash
# Vertex prefetch in the geometry kernel (off / L2 / L1): A/B timing.
OUT=gpurun_out
mkdir -p $OUT
for v in "" _p1 _p2; do
  SWR_LIB_VARIANT=$v python tools/gpu_time.py c3 c2 --tile 64 > $OUT/r02e_time$v.txt 2>&1
  SWR_LIB_VARIANT=$v python tools/gpu_time.py c5 --tile 64 >> $OUT/r02e_time$v.txt 2>&1
  echo "variant '$v'"; grep TIME $OUT/r02e_time$v.txt
done
python -m pytest tests -m gpu -x -q -k "not cpp_" > $OUT/r02e_pytest.txt 2>&1; tail -3 $OUT/r02e_pytest.txt
