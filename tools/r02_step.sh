#!/bin/bash
# One build -> measure step on the GPU box: parity suite, timings of the configs, per-tile phase split.
TAG=${1:-r02x}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.txt 2>&1; tail -3 $OUT/${TAG}_pytest.txt
python tools/gpu_time.py c3 c2 c0 c0_4k > $OUT/${TAG}_time.txt 2>&1
python tools/gpu_time.py c5 c4l c4p --tile 64 >> $OUT/${TAG}_time.txt 2>&1
grep TIME $OUT/${TAG}_time.txt
python tools/tile_stats.py c3 64 > $OUT/${TAG}_tilestats_c3_64.txt 2>&1
head -8 $OUT/${TAG}_tilestats_c3_64.txt
