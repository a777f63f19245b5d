#!/bin/bash
N=${1:-4}
OUT=gpurun_out
mkdir -p $OUT
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus $N --steps 20 --warmup 3 --also none > $OUT/r02g_e2e_n${N}.json 2> $OUT/r02g_e2e_n${N}.err
grep "\[bench\]" $OUT/r02g_e2e_n${N}.err | head -8
python - <<PY
import json
try:
    d=json.loads(open("$OUT/r02g_e2e_n${N}.json").read().strip().splitlines()[-1])
    e=d['e2e']
    print("N=$N: ms/step %.3f  e2e %.3f ms (host enqueue %.3f ms)  tile %d" % (d['ms_per_step'], e['ms_per_step'], e.get('host_enqueue_ms_per_step', -1), d['config']['tile_size']))
except Exception as ex:
    print("N=$N failed", ex); print(open("$OUT/r02g_e2e_n${N}.err").read()[-2500:])
PY
nvidia-smi topo -m > $OUT/r02g_topo.txt 2>&1
