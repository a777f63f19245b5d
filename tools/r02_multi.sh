#!/bin/bash
# Multi-GPU step: correctness of sharded geometry + mirrors over IPC, then bench at N (sharded / replicated geometry).
N=${1:-2}
TAG=${2:-r02m}
OUT=gpurun_out
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
timeout 300 $RUN tools/mirror_check.py --shards > $OUT/${TAG}_mirror_shards_n$N.txt 2>&1; grep -E "MIRROR|rank 0 frame|Error|error" $OUT/${TAG}_mirror_shards_n$N.txt | head -8
for geo in sharded replicated; do
  timeout 600 $RUN bench.py --gpus $N --steps 20 --warmup 3 --geometry $geo --no-cpu > $OUT/${TAG}_bench_c3_n${N}_$geo.json 2> $OUT/${TAG}_bench_c3_n${N}_$geo.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_bench_c3_n${N}_$geo.json").read().strip().splitlines()[-1])
    print("$geo N=$N", "ms/step %.3f"%d["ms_per_step"], "value %.3g"%d["value"], "e2e ms %.3f"%d["e2e"]["ms_per_step"], "tile", d["config"]["tile_size"], "kernel_ms", d["roofline"]["kernel_ms"], "geom_ms", d["roofline"]["geometry_kernel_ms"])
except Exception as e:
    print("$geo failed", e); print(open("$OUT/${TAG}_bench_c3_n${N}_$geo.err").read()[-1500:])
PY
done
