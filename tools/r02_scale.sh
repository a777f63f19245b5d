#!/bin/bash
# bench.py at several N / tile sizes on one box.  usage: bash tools/r02_scale.sh TAG "N:tile N:tile ..." [extra bench args]
TAG=$1; shift
SPECS=$1; shift
OUT=gpurun_out
mkdir -p $OUT
for spec in $SPECS; do
  N=${spec%%:*}; T=${spec##*:}
  if [ "$N" = "1" ]; then RUN="python"; else RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"; fi
  F=$OUT/${TAG}_n${N}_t${T}
  timeout 900 $RUN bench.py --gpus $N --steps 20 --warmup 3 --tile $T --no-cpu "$@" > $F.json 2> $F.err
  python - <<PY
import json
try:
    d=json.loads(open("$F.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("N=$N tile=$T", "ms/step %.3f"%d["ms_per_step"], "value %.3g"%d["value"], "e2e ms %.3f"%d["e2e"]["ms_per_step"], "tile_used", d["config"]["tile_size"], "tile_ms %.3f"%r["kernel_ms"], "geom_ms %.3f"%r["geometry_kernel_ms"], "launches", d["gpu_launches"])
except Exception as e:
    print("N=$N tile=$T failed", e); print(open("$F.err").read()[-1500:])
PY
done
