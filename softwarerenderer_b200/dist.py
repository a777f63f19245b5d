"""Sort-first multi-GPU partition and framebuffer composite (K7 of SURVEY.md 2.3 / 8(e)).

Screen tiles are interleaved across ranks with
    owner(tx, ty) = (tx + 3 * ty) % world           (include/swr/detail/common.h: tileOwned)
and each rank rasterizes only its tiles.  The vertex stage is either replicated or sharded (GeometryShards: every rank
runs 1/world of the batches and its geometry kernel stores the records into the tile owners' scratch over NVLink).
The composite is a pure copy of disjoint tiles (no reduction, so bit-exactness is preserved), in two forms:
TileMirror -- the tile kernel's final store of a tile also goes to every peer's surface (peer mappings over NVLink,
one barrier per frame) -- and TileComposite -- every rank packs its tiles into a dense tile-major buffer, one NCCL
all-gather moves them, every rank unpacks its peers' tiles -- for set-ups without peer access.  Either way every rank
ends up with the full image.

The layout helpers are plain numpy (host logic, covered by world_size-2 gloo tests on CPU); on
the GPU the pack / unpack are CUDA kernels behind swr_pack_tiles / swr_unpack_tiles.
"""
from __future__ import annotations

import numpy as np


def tile_grid(width: int, height: int, tile: int):
    return (width + tile - 1) // tile, (height + tile - 1) // tile


def tile_owner(tx, ty, world: int):
    return (tx + 3 * ty) % world if world > 1 else 0 * (tx + ty)


def owned_tiles(width: int, height: int, tile: int, rank: int, world: int) -> np.ndarray:
    """Tile ids (ty * tiles_x + tx) owned by `rank`, ascending -- the order of the exchange buffer."""
    tiles_x, tiles_y = tile_grid(width, height, tile)
    ty, tx = np.divmod(np.arange(tiles_x * tiles_y), tiles_x)
    return np.nonzero(tile_owner(tx, ty, world) == rank)[0]


def max_owned(width: int, height: int, tile: int, world: int) -> int:
    return max(len(owned_tiles(width, height, tile, r, world)) for r in range(world))


def pack_tiles_host(surface: np.ndarray, tile: int, rank: int, world: int, slots: int = 0) -> np.ndarray:
    """surface [H, W] 32-bit -> [n_owned (or slots), tile, tile]; pixels outside the surface are 0."""
    h, w = surface.shape
    tiles_x, _ = tile_grid(w, h, tile)
    ids = owned_tiles(w, h, tile, rank, world)
    out = np.zeros((max(slots, len(ids)), tile, tile), dtype=surface.dtype)
    for k, t in enumerate(ids):
        ty, tx = divmod(int(t), tiles_x)
        blk = surface[ty * tile:(ty + 1) * tile, tx * tile:(tx + 1) * tile]
        out[k, :blk.shape[0], :blk.shape[1]] = blk
    return out


def unpack_tiles_host(surface: np.ndarray, buf: np.ndarray, tile: int, rank: int, world: int) -> None:
    """Inverse of pack_tiles_host: writes rank's tiles from `buf` into `surface` in place."""
    h, w = surface.shape
    tiles_x, _ = tile_grid(w, h, tile)
    for k, t in enumerate(owned_tiles(w, h, tile, rank, world)):
        ty, tx = divmod(int(t), tiles_x)
        blk = surface[ty * tile:(ty + 1) * tile, tx * tile:(tx + 1) * tile]
        blk[...] = buf[k, :blk.shape[0], :blk.shape[1]]


def composite_host(surface: np.ndarray, tile: int, rank: int, world: int, all_gather) -> None:
    """CPU restatement of TileComposite.run for the gloo tests: all_gather(send) -> list of buffers."""
    slots = max_owned(surface.shape[1], surface.shape[0], tile, world)
    send = pack_tiles_host(surface, tile, rank, world, slots)
    for r, buf in enumerate(all_gather(send)):
        if r != rank:
            unpack_tiles_host(surface, buf, tile, r, world)


class TileComposite:
    """GPU composite of one registered render-target slot with torch.distributed (NCCL)."""

    def __init__(self, rasterizer, width: int, height: int, tile: int, rank: int, world: int, device):
        import torch
        self.r, self.tile, self.rank, self.world = rasterizer, tile, rank, world
        self.slots = max_owned(width, height, tile, world)
        n = self.slots * tile * tile
        self.send = torch.zeros(n, dtype=torch.int32, device=device)
        self.recv = torch.zeros(world * n, dtype=torch.int32, device=device)
        self.bytes_per_rank = n * 4

    def run(self, slot: int) -> None:
        """pack (CUDA kernel) -> NCCL all_gather over NVLink -> unpack peers' tiles (CUDA kernels).
        The rasterizer must enqueue on torch's current stream (Rasterizer.setStream)."""
        import torch.distributed as dist
        from . import _lib
        lib = _lib.load()
        r, w = self.rank, self.world
        _lib.check(lib.swr_pack_tiles(self.r.ctx, slot, r, w, self.tile, self.send.data_ptr()), "pack_tiles")
        dist.all_gather_into_tensor(self.recv, self.send)
        n = self.send.numel()
        for peer in range(w):
            if peer != r:
                _lib.check(lib.swr_unpack_tiles(self.r.ctx, slot, peer, w, self.tile, self.recv.data_ptr() + peer * n * 4),
                           "unpack_tiles")


class TileMirror:
    """Composite fused into the tile kernel: every rank maps the peers' surface of one render-target slot
    (CUDA IPC over NVLink) and its tile kernel stores each finished tile to all of them, so after the draws
    plus one barrier every rank holds the whole frame -- no pack / all-gather / unpack pass.  torch.distributed
    only carries the 64-byte handles at set-up and the barrier.

    Contract (see swr_set_tile_mirrors): all ranks clear the surface the same way, and call barrier() between
    that clear and the first draw of a frame as well as after the last draw, before the surface is read."""

    def __init__(self, rasterizer, slot: int, surface_ptr: int, rank: int, world: int, device, peer_barrier=None):
        import torch
        import torch.distributed as dist
        from . import api
        self.r, self.slot, self.rank, self.world = rasterizer, slot, rank, world
        self._peer_barrier = peer_barrier     # e.g. GeometryShards.barrier: the library's own flag barrier instead of NCCL
        # every step that can fail is followed by a collective agreement, so that all ranks raise together
        # (and the caller can fall back to TileComposite on all of them) instead of one rank leaving a collective
        try:
            mine = api.ipc_handle(surface_ptr)
        except Exception as e:                       # e.g. memory that cannot be exported (cuMemMap-based allocators)
            mine = ("error", str(e))
        handles = [None] * world
        dist.all_gather_object(handles, mine)
        bad = [h for h in handles if h[0] == "error"]
        if bad:
            raise RuntimeError(f"TileMirror: a surface cannot be exported over CUDA IPC: {bad[0][1]}")
        self.peers = []
        err = None
        try:
            for i, (h, off) in enumerate(handles):
                if i != rank:
                    self.peers.append(rasterizer.ipcOpen(h, off))
        except Exception as e:
            err = e
        ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            for p in self.peers:
                rasterizer.ipcClose(p)
            self.peers = []
            raise RuntimeError(f"TileMirror: a peer surface cannot be mapped: {err}")
        rasterizer.setTileMirrors(slot, self.peers)
        self._token = torch.zeros(1, dtype=torch.int32, device=device)

    def barrier(self) -> None:
        """Stream-ordered cross-rank barrier: the library's flag barrier when the partition has one (sharded geometry),
        else a one-word NCCL all-reduce on torch's current stream."""
        if self._peer_barrier is not None:
            self._peer_barrier()
            return
        import torch.distributed as dist
        dist.all_reduce(self._token)

    def run(self, slot: int) -> None:       # same call shape as TileComposite.run
        self.barrier()

    def close(self) -> None:
        self.r.setTileMirrors(self.slot, [])
        for p in self.peers:
            self.r.ipcClose(p)
        self.peers = []


SHARD_BATCHES = 16      # include/swr/detail/common.h: kShardBatches
BATCH_PRIMS = 1024      # VertexProcessor.cpp:110


def shard_batches(num_batches: int, rank: int, world: int) -> np.ndarray:
    """Batches (of 1024 input primitives) the geometry kernel of `rank` runs under swr_set_geometry_shards: runs of 16
    consecutive batches go round-robin to the ranks (geometry.cuh: batchOfBlock / launchGeometry)."""
    runs = -(-num_batches // SHARD_BATCHES)
    mine = np.arange(rank, runs, world)
    b = (mine[:, None] * SHARD_BATCHES + np.arange(SHARD_BATCHES)[None, :]).reshape(-1)
    return b[b < num_batches]


def shard_index_ranges(index_count: int, per: int, rank: int, world: int):
    """[(first index, index count)] of the index array a rank reads with sharded geometry -- what it has to upload."""
    nprims = index_count // per
    out = []
    for b in shard_batches(-(-nprims // BATCH_PRIMS), rank, world):
        first = int(b) * BATCH_PRIMS
        n = min(BATCH_PRIMS, nprims - first)
        if out and out[-1][0] + out[-1][1] == first * per:
            out[-1] = (out[-1][0], out[-1][1] + n * per)
        else:
            out.append((first * per, n * per))
    return out


class GeometryShards:
    """Sharded vertex stage of a sort-first partition (swr_set_geometry_shards): every rank runs 1/world of the batches
    and its geometry kernel stores each surviving record straight into the scratch of the ranks whose tiles it touches
    (peer mappings over NVLink); the kernels of the ranks are ordered by the library's own flag barrier.
    torch.distributed only carries the 64-byte IPC handles at set-up.  All ranks must pass the same `scratch_bytes`
    and then issue the same sequence of draws and barrier() calls."""

    def __init__(self, rasterizer, rank: int, world: int, device, scratch_bytes: int):
        import torch
        import torch.distributed as dist
        from . import api
        self.r, self.rank, self.world = rasterizer, rank, world
        self.peers = []
        base = rasterizer.createSharedScratch(scratch_bytes)
        try:
            mine = api.ipc_handle(base)
        except Exception as e:
            mine = ("error", str(e))
        handles = [None] * world
        dist.all_gather_object(handles, mine)
        bad = [h for h in handles if h[0] == "error"]
        if bad:
            raise RuntimeError(f"GeometryShards: a scratch arena cannot be exported over CUDA IPC: {bad[0][1]}")
        arenas, err = [0] * world, None
        try:
            for i, (h, off) in enumerate(handles):
                if i != rank:
                    arenas[i] = rasterizer.ipcOpen(h, off)
                    self.peers.append(arenas[i])
        except Exception as e:
            err = e
        ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)          # also the barrier "every arena exists and is zeroed"
        if int(ok.item()) == 0:
            for p in self.peers:
                rasterizer.ipcClose(p)
            self.peers = []
            raise RuntimeError(f"GeometryShards: a peer scratch arena cannot be mapped: {err}")
        rasterizer.setGeometryShards(rank, world, arenas)

    def barrier(self) -> None:
        self.r.peerBarrier()

    def run(self, slot: int) -> None:       # same call shape as TileComposite.run / TileMirror.run
        self.r.peerBarrier()

    def close(self) -> None:
        self.r.setGeometryShards(0, 1, [])
        for p in self.peers:
            self.r.ipcClose(p)
        self.peers = []


class ReplicatedUpload:
    """Host -> every GPU, for buffers that all ranks need in full (sort-first replicates the geometry): each rank
    copies only 1/world of the bytes over its own PCIe link, then one NCCL all-gather over NVLink gives every
    rank the whole buffer.  Several arrays are packed into one byte buffer (256-byte aligned sections)."""

    def __init__(self, arrays, rank: int, world: int, device, group=None):
        import numpy as np
        import torch
        self.group = group                              # a process group of its own lets the gather overlap other collectives
        self.offsets = []
        total = 0
        for a in arrays:
            self.offsets.append(total)
            total += (a.nbytes + 255) // 256 * 256
        self.per = ((total + world - 1) // world + 255) // 256 * 256
        self.total = total
        host = np.zeros(self.per * world, dtype=np.uint8)
        for a, off in zip(arrays, self.offsets):
            host[off:off + a.nbytes] = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
        self.slice_host = torch.from_numpy(host[rank * self.per:(rank + 1) * self.per].copy())
        if torch.cuda.is_available():
            self.slice_host = self.slice_host.pin_memory()
        self.slice_dev = torch.empty(self.per, dtype=torch.uint8, device=device)
        self.full = torch.empty(self.per * world, dtype=torch.uint8, device=device)
        self.h2d_bytes = self.per                       # per rank and step

    def run(self, h2d_done=None):
        """Enqueue on torch's current stream; returns the device address of each array.  `h2d_done` (a CUDA event) is
        recorded between the PCIe copy and the all-gather: other host -> device copies of the step can wait for it on
        a stream of their own and then run under the all-gather (PCIe and NVLink are independent)."""
        import torch.distributed as dist
        self.slice_dev.copy_(self.slice_host, non_blocking=True)
        if h2d_done is not None:
            h2d_done.record()
        dist.all_gather_into_tensor(self.full, self.slice_dev, group=self.group)
        base = self.full.data_ptr()
        return [base + off for off in self.offsets]
