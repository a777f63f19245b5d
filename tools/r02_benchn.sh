#!/bin/bash
# The contract's bench command at N GPUs (gpurun_out/r02_bench_n<N>.json).  usage: bash tools/r02_benchn.sh N [extra args]
N=$1; shift
OUT=gpurun_out
mkdir -p $OUT
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 "$@" > $OUT/r02_bench_n$N.json 2> $OUT/r02_bench_n$N.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/r02_bench_n$N.json").read().strip().splitlines()[-1])
    print("c3 N=$N: ms/step %.3f value %.3g e2e %.3f tile %d tile_ms %.3f geom_ms %.3f" % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['config']['tile_size'], d['roofline']['kernel_ms'], d['roofline']['geometry_kernel_ms']))
    c5=d['also']['c5']; print("c5 N=$N: ms/step %.3f value %.3g tile %d passes %d" % (c5['ms_per_step'], c5['value'], c5['config']['tile_size'], c5['config']['passes_per_step']))
except Exception as e:
    print("N=$N failed", e); print(open("$OUT/r02_bench_n$N.err").read()[-1500:])
PY
