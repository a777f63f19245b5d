"""Multi-GPU check of the fused composite (swr_set_tile_mirrors over CUDA IPC): every rank renders its own tiles,
stores them into all peers' surfaces, and must end up with the frame the CPU oracle renders.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/mirror_check.py [scene]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from softwarerenderer_b200 import api, scenes as S  # noqa: E402
from softwarerenderer_b200.api import SceneRenderer  # noqa: E402
from softwarerenderer_b200.dist import GeometryShards, TileMirror  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

argv = [a for a in sys.argv[1:] if not a.startswith("--")]
scene = S.config_c2(nx=300, ny=200, width=1280, height=720) if not argv else getattr(S, argv[0])()
sr = SceneRenderer(scene.width, scene.height, device=local)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
sr.r.setStream(stream.cuda_stream)
sr.r.setTilePartition(rank, world)
shards = None
if "--shards" in sys.argv:                # sharded geometry: records pushed to the tile owners, library barrier
    sr.draw(scene)
    shards = GeometryShards(sr.r, rank, world, dev, 2 << 30)
mirror = TileMirror(sr.r, api.RT_COLOR, sr.targets.ptr(api.RT_COLOR), rank, world, dev,
                    peer_barrier=shards.barrier if shards else None)
ok = True
for frame in range(3):
    sr.targets.clear()
    mirror.barrier()                     # nobody draws into a surface that is still being cleared
    sr.draw(scene, wait=False)
    mirror.barrier()                     # every rank's tiles have landed
    out = np.empty((scene.height, scene.width), dtype=np.uint32)
    sr.r.download(sr.targets.ptr(api.RT_COLOR), out)
    sr.r.finish()
    if frame == 0:
        from oracle import pyoracle as O
        O.build()
        want = O.run(scene, "oracle")["color"].reshape(out.shape)
    good = bool(np.array_equal(out, want))
    ok = ok and good
    print(f"rank {rank} frame {frame}: {'full frame matches the oracle' if good else 'MISMATCH ' + str(int((out != want).sum()))}", flush=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
torch.cuda.synchronize()
mirror.close()
if shards:
    shards.close()
sr.close()
if rank == 0:
    print("MIRROR OK" if int(flag.item()) == 1 else "MIRROR FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
