"""CPU: the sort-first partition and the tile composite layout, world_size 2 over gloo.
Each rank renders (with the oracle, standing in for its GPU) only the tiles it owns, the packed
tiles are all-gathered, and every rank must end up with the full reference image."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from softwarerenderer_b200 import dist as D
from softwarerenderer_b200 import scenes as S


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tile, result_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle as O
        scene = S.config_c2(60, 40, 200, 136, ps=S.PS_GOURAUD_DEPTH)       # 200x136: partial edge tiles
        full = O.run(scene, "oracle")["color"].reshape(scene.height, scene.width)
        # this rank's GPU would have rendered only its own tiles: blank the others
        mine = np.zeros_like(full)
        tiles_x, _ = D.tile_grid(scene.width, scene.height, tile)
        for t in D.owned_tiles(scene.width, scene.height, tile, rank, world):
            ty, tx = divmod(int(t), tiles_x)
            sl = (slice(ty * tile, (ty + 1) * tile), slice(tx * tile, (tx + 1) * tile))
            mine[sl] = full[sl]

        def all_gather(send):
            t = torch.from_numpy(send.view(np.int32).copy())
            outs = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(outs, t)
            return [o.numpy().view(send.dtype) for o in outs]

        D.composite_host(mine, tile, rank, world, all_gather)
        result_q.put((rank, bool(np.array_equal(mine, full))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("tile", [32, 64])
def test_two_rank_composite_gloo(oracle, tile):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, tile, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]


def _upload_worker(rank, world, port, result_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)                                      # same arrays on every rank
        verts = rng.random((1001, 7), dtype=np.float32)
        idx = rng.integers(0, 1001, size=2999, dtype=np.int32)
        up = D.ReplicatedUpload([verts, idx], rank, world, torch.device("cpu"))
        up.run()
        full = up.full.numpy()
        ok = bool(np.array_equal(full[up.offsets[0]:up.offsets[0] + verts.nbytes].view(np.float32).reshape(verts.shape), verts) and
                  np.array_equal(full[up.offsets[1]:up.offsets[1] + idx.nbytes].view(np.int32), idx))
        result_q.put((rank, ok and up.h2d_bytes * world >= verts.nbytes + idx.nbytes and up.offsets[1] % 256 == 0))
    finally:
        dist.destroy_process_group()


def test_replicated_upload_gloo():
    """Each rank contributes 1/world of the packed geometry; after the all-gather everyone holds all of it."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_upload_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]


def test_owned_tiles_partition_and_match_abi():
    from softwarerenderer_b200 import _lib
    lib = _lib.load()
    for (w, h, tile) in ((200, 136, 32), (3840, 2160, 64), (1920, 1080, 32)):
        for world in (1, 2, 4, 8):
            seen = np.concatenate([D.owned_tiles(w, h, tile, r, world) for r in range(world)])
            tx, ty = D.tile_grid(w, h, tile)
            assert sorted(seen.tolist()) == list(range(tx * ty))
            for r in range(world):
                assert len(D.owned_tiles(w, h, tile, r, world)) == lib.swr_owned_tile_count(w, h, tile, r, world)


def test_pack_unpack_roundtrip_host():
    rng = np.random.default_rng(0)
    surf = rng.integers(0, 1 << 32, size=(136, 200), dtype=np.uint32)
    for world in (2, 4):
        out = np.zeros_like(surf)
        for r in range(world):
            D.unpack_tiles_host(out, D.pack_tiles_host(surf, 32, r, world, D.max_owned(200, 136, 32, world)), 32, r, world)
        assert np.array_equal(out, surf)


def _shard_worker(rank, world, port, nbatches, result_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = D.shard_batches(nbatches, rank, world)
        flags = torch.zeros(nbatches, dtype=torch.int32)
        flags[torch.from_numpy(mine)] = 1
        dist.all_reduce(flags)                       # every batch must be run by exactly one rank
        ranges = D.shard_index_ranges(nbatches * 1024 * 3 - 5 * 3, 3, rank, world)
        covered = torch.zeros(1, dtype=torch.int64)
        covered[0] = sum(n for _, n in ranges)
        dist.all_reduce(covered)
        result_q.put((rank, bool((flags == 1).all()), int(covered.item())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nbatches", [1, 17, 9766])
def test_geometry_shards_partition_the_batches_gloo(nbatches):
    """Sharded geometry: runs of 16 batches round-robin over the ranks -- every batch exactly once, and the index
    ranges the ranks upload add up to the whole index array (world_size 2 over gloo)."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, nbatches, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, once, covered in results:
        assert once, f"rank {rank}: a batch is run twice or never"
        assert covered == nbatches * 1024 * 3 - 15


def test_shard_batches_match_the_kernel_mapping():
    """dist.shard_batches is the host-side statement of geometry.cuh's batchOfBlock: CTA `block` of rank r runs batch
    ((block / 16) * world + r) * 16 + block % 16, CTAs past the end of the pass exit."""
    for world in (2, 3, 4, 8):
        for nb in (1, 16, 33, 1000):
            seen = []
            for r in range(world):
                runs = -(-nb // 16)
                mine = (runs - r + world - 1) // world if runs > r else 0
                got = [((blk // 16) * world + r) * 16 + blk % 16 for blk in range(mine * 16)]
                got = [b for b in got if b < nb]
                assert got == D.shard_batches(nb, r, world).tolist()
                seen += got
            assert sorted(seen) == list(range(nb))
