#!/bin/bash
# Heavy-tile split: parity, then the tile phase of every rank of a 1/2/4/8 partition emulated on one GPU; F3 prefetch A/B.
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q -k "heavy_tile_split or partition or in_kernel or tile_size_independence" > $OUT/r02c_pytest.txt 2>&1; tail -3 $OUT/r02c_pytest.txt
python tools/rank_emul.py c3 > $OUT/r02c_emul_c3.txt 2>&1; grep EMUL $OUT/r02c_emul_c3.txt
SWR_LIB_VARIANT=_pf python tools/rank_emul.py c3 --worlds 1,4 --configs 32:0,64:0 > $OUT/r02c_emul_c3_pf.txt 2>&1; grep EMUL $OUT/r02c_emul_c3_pf.txt
