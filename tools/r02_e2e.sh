#!/bin/bash
# N-GPU end-to-end leg: plain form vs. index upload under the all-gather + banded read-back.  usage: bash tools/r02_e2e.sh N
N=$1; shift
OUT=gpurun_out
mkdir -p $OUT
for ov in 0 1; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$ov bench.py --gpus $N --steps 20 --warmup 3 --also none --e2e-overlap $ov "$@" > $OUT/r02f_e2e_n${N}_ov$ov.json 2> $OUT/r02f_e2e_n${N}_ov$ov.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/r02f_e2e_n${N}_ov$ov.json").read().strip().splitlines()[-1])
    e=d['e2e']
    print("N=$N overlap=$ov: ms/step %.3f  e2e %.3f ms (host enqueue %.3f ms)  tile %d  path: %s" % (d['ms_per_step'], e['ms_per_step'], e.get('host_enqueue_ms_per_step', -1), d['config']['tile_size'], e['path'][-90:]))
except Exception as ex:
    print("N=$N overlap=$ov failed", ex); print(open("$OUT/r02f_e2e_n${N}_ov$ov.err").read()[-2500:])
PY
done
