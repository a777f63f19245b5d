#!/usr/bin/env python
"""bench.py -- the headline benchmark of the draw path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c3|c2|c0|c0_4k|c5]

Workload (config.workload): BASELINE.json configs[2], the configuration the metric is quoted on --
a synthetic 10M tiny-triangle mesh at 3840x2160 (heavy near-plane clipping, ~45 % clockwise
triangles removed by CullMode::CW), RasterMode::Block, Gouraud pixel shader, one draw per step.

  value   shaded fragments/s (drawPixel invocations / s), whole job, inputs resident in HBM,
          CUDA-event timed on the stream the kernels are launched on, max over ranks.
  e2e     the same metric through the reference-facing call with HOST buffers: every step copies the
          vertex and index buffers host->device (pinned memory), draws, and reads the colour buffer
          back device->host, all inside the timed region.
  N > 1   sort-first: screen tiles interleaved across ranks, the vertex stage sharded by batches with the
          records pushed to the tile owners over NVLink, the composite fused into the tile kernel's
          store (peer surfaces) -- "scaling": "strong": the frame is fixed.
  --impl reference   the reference's own CPU renderer (oracle/_ref when it was built from
          /root/reference, else the oracle port) on the host cores, same workload and metric.

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from softwarerenderer_b200 import scenes as S  # noqa: E402

METRIC = "shaded_fragments_per_s"
UNIT = "fragments/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload: str, tile: int, world: int):
    """(bytes, note): dram__bytes_read.sum + dram__bytes_write.sum of the dominant (tile) kernel, per launch, from the
    committed `ncu --set full` capture of the SAME workload, tile size and GPU count (profiles/r02_traffic.json, written
    by tools/profile_summary.py); (None, why) when no such capture exists -- a number from another configuration
    would not describe this line."""
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    except Exception:
        return None, "profiles/r02_traffic.json not found"
    for e in table.get("captures", []):
        if e["workload"] == workload and e["tile"] == tile and e["world"] == world and e["kernel"] == "tile":
            return float(e["dram_bytes"]), f"ncu --set full, {e['source']}"
    return None, f"no ncu capture of workload {workload} with {tile}-pixel tiles on {world} GPU(s) is committed"


def make_scene(name: str):
    if name == "c3":
        return S.config_c3(), 4, "BASELINE.json configs[2]: 10M tiny-triangle mesh, 3840x2160, Block, Gouraud, CullMode::CW"
    if name == "c2":
        return S.config_c2(), 12, "BASELINE.json configs[1]: 1M-triangle grid, 1920x1080, Block, Gouraud + depth test"
    if name == "c0":
        return S.config_c0(), 4, "Benchmark.cpp: 40 960 random triangles, 640x480, Span, flat shader"
    if name == "c0_4k":
        return S.config_c0(3840, 2160), 4, "Benchmark.cpp's triangles at 3840x2160, Span, flat shader"
    if name == "c5":
        return S.config_c5(), 4, "BASELINE.json configs[4]: 50M perspective-textured triangles (5 layers), 7680x4320, Block, CullMode::CW"
    if name == "fill":
        return S.config_fill(), 4, "fill-bound: 2 layers of 96 x 54 x 2 large triangles (40 pixels across) over 3840x2160, overdraw 2, Block, flat 4-byte store"
    raise SystemExit(f"unknown workload {name}")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def algorithmic_bytes(scene, fragments: int, b_frag: int):
    """B_alg = 4*I + S*V_ref + F*b_frag."""
    v_ref = int(np.unique(scene.indices).size)
    geom = 4 * int(scene.indices.size) + scene.stride * v_ref
    return geom, fragments * b_frag, v_ref


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_rate(scene, budget_s: float = 20.0, steps: int = 1):
    """Times the reference CPU renderer (or the oracle port) on this box's host cores.
    Returns (fragments/s, triangles/s, description dict)."""
    from oracle import pyoracle as O
    O.build()
    impl, kind = ("ref", "reference") if O.have_ref() else ("oracle", "port")
    nprim = scene.num_primitives
    per = scene.draw_mode + 1
    # probe on 64 batches to size a bounded sample of the same workload
    probe_n = min(nprim, 64 * 1024)
    sub = scene.replace(indices=scene.indices[:probe_n * per])
    t0 = time.perf_counter()
    O.run(sub, impl)
    probe = max(time.perf_counter() - t0, 1e-4)
    n = int(min(nprim, max(probe_n, (budget_s / steps) / probe * probe_n)))
    n = max(1024, n // 1024 * 1024) if n < nprim else nprim
    # spread the sample over the whole mesh (perspective makes fragment density non-uniform):
    # every k-th batch of 1024 primitives
    nb_all = (nprim + 1023) // 1024
    nb = nb_all if n >= nprim else max(1, min(nb_all, n // 1024))
    pick = np.unique(np.linspace(0, nb_all - 1, nb).astype(np.int64))
    idx = scene.indices.reshape(-1, per)
    sel = np.concatenate([idx[b * 1024:(b + 1) * 1024] for b in pick]).reshape(-1) if nb < nb_all else scene.indices
    sample = scene.replace(indices=np.ascontiguousarray(sel))
    times, frags = [], 0
    for _ in range(steps):
        out = O.run(sample, impl)
        times.append(out["seconds"])
        frags = out["fragments"]
    t = float(np.median(times))
    desc = {"kind": kind, "cores": 1,
            "sample": f"{sample.num_primitives} of {nprim} primitives ({len(pick)} of {nb_all} batches of 1024, evenly spaced), "
                      f"{frags} fragments, median of {steps} run(s) of {t:.3f} s; the reference draw is single-threaded "
                      f"(OpenMP only inside one triangle, Rasterizer.h:257,369,396)"}
    return frags / t, sample.num_primitives / t, desc, t


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene, b_frag, wl = make_scene(args.workload)
    budget = args.ref_budget
    per_step = budget / max(1, args.steps + args.warmup)
    fps, tps, desc, t = cpu_reference_rate(scene, budget_s=per_step * max(1, args.steps), steps=max(1, args.steps))
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "triangles_per_s": tps,
        "config": {"workload": wl, "note": "CPU reference renderer on host cores; each step is a bounded sample of the workload"},
        "cpu_baseline": dict(desc, value=fps, unit=UNIT),
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
DEPTH_TESTED = ("c2",)          # workloads whose pixel shader reads what earlier fragments wrote (depth test)


def bench_workload(args, workload, steps, warmup, rank, local_rank, world, dev, stream, with_cpu, with_clocks):
    """One workload on this process's GPU (all ranks call it together).  Returns the result dictionary on rank 0."""
    import torch
    import torch.distributed as dist
    from softwarerenderer_b200 import api
    from softwarerenderer_b200.dist import GeometryShards, ReplicatedUpload, TileComposite, TileMirror

    scene, b_frag, wl = make_scene(workload)
    W, H = scene.width, scene.height
    geometry = args.geometry
    composite = args.composite
    r = api.Rasterizer(local_rank)
    v = api.VertexProcessor(r)
    r.setStream(stream.cuda_stream)
    if args.tile:
        r.setTileSize(args.tile)

    # device surfaces (torch owns the memory; the library gets raw pointers)
    nrt = api._lib.MAX_RENDER_TARGETS if workload != "c5" else 4          # (8K: 12 surfaces would be 1.6 GB for nothing)
    targets = torch.zeros((nrt, H, W), dtype=torch.int32, device=dev)
    for s in range(nrt):
        r.setRenderTarget(s, targets[s].data_ptr(), W * 4, W, H)
    depth_tested = workload in DEPTH_TESTED

    def clear():
        targets.zero_()
        targets[api.RT_DEPTH].fill_(0x3F800000)

    def clear_in_step():
        # depth-tested workloads: every timed step starts from a cleared colour and depth buffer, with the library's
        # own fill kernel on the same stream (otherwise every step after the first would fail the depth test everywhere)
        r.fill32(targets[api.RT_COLOR].data_ptr(), 0, W * H)
        r.fill32(targets[api.RT_DEPTH].data_ptr(), 0x3F800000, W * H)

    # geometry: pinned host copies (e2e) and resident device copies (value)
    h_vert = torch.from_numpy(scene.vertices).pin_memory()
    h_idx = torch.from_numpy(scene.indices).pin_memory()
    h_color = torch.empty((H, W), dtype=torch.int32).pin_memory()
    d_vert = h_vert.to(dev)
    d_idx = h_idx.to(dev)

    # state, exactly as a user of the reference sets it up
    r.setRasterMode(scene.raster_mode)
    r.setScissorRect(*scene.scissor)
    r.setPixelShader(scene.ps)
    v.setViewport(*scene.viewport)
    v.setDepthRange(*scene.depth_range)
    v.setCullMode(scene.cull_mode)
    v.setVertexShader(scene.vs)
    u = api.StockUniforms()
    u.mvp = (C.c_float * 16)(*[float(x) for x in scene.mvp.reshape(-1)])
    tex = None
    if scene.texture is not None:
        tex = torch.from_numpy(np.ascontiguousarray(scene.texture).view(np.int32)).to(dev)
        u.texture = tex.data_ptr()
        u.tex_h, u.tex_w = scene.texture.shape
    r.setUniforms(u)
    r.setTilePartition(rank, world)
    shards = None
    if world > 1 and geometry == "sharded":
        # sharded vertex stage: every rank runs 1/world of the batches and pushes the records to the tile owners
        scratch_gb = args.scratch_gb if args.scratch_gb > 0 else (40.0 if workload == "c5" else 14.0)
        try:
            shards = GeometryShards(r, rank, world, dev, int(scratch_gb * (1 << 30)))
        except RuntimeError as e:                     # raised on every rank together
            print(f"[bench] {e}; geometry stays replicated", file=sys.stderr)
            geometry = "replicated"

    count = int(scene.indices.size)
    count_all = count
    comp = None

    def step_resident():
        if depth_tested:
            clear_in_step()
            if comp is not None:
                comp.barrier()            # no peer stores a tile into a surface that is still being cleared
        v.setVertexAttribPointer(0, scene.stride, d_vert)
        v.drawElements(scene.draw_mode, count, d_idx, wait=False)
        if comp is not None:
            comp.run(api.RT_COLOR)

    # N > 1: every rank uploads 1/N of the geometry over its own PCIe link and an NCCL all-gather replicates it.
    # Two buffers alternate and the upload of step k+1 (on a second stream and its own process group) runs
    # under the draw of step k; every timed step still contains exactly one upload, one draw and one read-back.
    repl = None
    own_runs = None
    idx_stream = None
    if world > 1 and with_cpu is not None:
        up_group = dist.new_group(ranks=list(range(world)))
        up_stream = torch.cuda.Stream(device=dev)
        if shards is not None:
            # the vertices are needed in full by every rank (replicated: 1/world over PCIe each + one all-gather over
            # NVLink); of the indices a rank only reads the runs of 16 batches it processes: they go straight from its
            # own host memory to their place in the index buffer
            run_len = 16 * 1024 * (scene.draw_mode + 1)
            n_runs = -(-count_all // run_len)
            pad_runs = -(-n_runs // world) * world
            h_pad = np.zeros(pad_runs * run_len, dtype=np.int32)
            h_pad[:count_all] = scene.indices
            own_runs = torch.from_numpy(h_pad.reshape(pad_runs // world, world, run_len)[:, rank].copy()).pin_memory()
            d_idx_sh = [torch.zeros(pad_runs * run_len, dtype=torch.int32, device=dev) for _ in range(2)]
            repl = [ReplicatedUpload([scene.vertices], rank, world, dev, group=up_group) for _ in range(2)]
        else:
            repl = [ReplicatedUpload([scene.vertices, scene.indices], rank, world, dev, group=up_group) for _ in range(2)]
        up_ready = [torch.cuda.Event(), torch.cuda.Event()]
        up_free = [torch.cuda.Event(), torch.cuda.Event()]
        up_state = {"k": 0, "ptrs": [None, None], "drawn": [False, False]}
        # the index runs go over PCIe right behind the vertex slice, on a stream of their own, i.e. under the all-gather
        # of the vertices (NVLink) instead of behind it
        idx_stream = torch.cuda.Stream(device=dev) if shards is not None and args.e2e_overlap else None
        h2d_done = [torch.cuda.Event(), torch.cuda.Event()]
        idx_ready = [torch.cuda.Event(), torch.cuda.Event()]

        def enqueue_upload(slot):
            with torch.cuda.stream(up_stream):
                if up_state["drawn"][slot]:
                    up_stream.wait_event(up_free[slot])          # the draw that last read this buffer is done
                ptrs = repl[slot].run(h2d_done=h2d_done[slot] if idx_stream is not None else None)
                if shards is not None:
                    if idx_stream is None:
                        d_idx_sh[slot].view(-1, world, run_len)[:, rank].copy_(own_runs, non_blocking=True)
                    ptrs = [ptrs[0], d_idx_sh[slot].data_ptr()]
                up_state["ptrs"][slot] = ptrs
                up_ready[slot].record(up_stream)
            if idx_stream is not None:
                with torch.cuda.stream(idx_stream):
                    if up_state["drawn"][slot]:
                        idx_stream.wait_event(up_free[slot])
                    idx_stream.wait_event(h2d_done[slot])
                    d_idx_sh[slot].view(-1, world, run_len)[:, rank].copy_(own_runs, non_blocking=True)
                    idx_ready[slot].record(idx_stream)

    # N > 1 read-back: after the composite every rank holds the whole frame, so every rank copies one band of rows over
    # its own PCIe link into ONE host buffer (POSIX shared memory, page-locked by every process): the frame arrives
    # in 1/N of the time it takes rank 0 alone.  Falls back to rank 0 reading everything when the mapping fails.
    frame_shared, band = None, None
    if world > 1 and with_cpu is not None and args.e2e_overlap:
        name = [f"/dev/shm/swr_frame_{os.getpid()}_{workload}"] if rank == 0 else [None]
        dist.broadcast_object_list(name, src=0)
        ok = 1
        for creator in (True, False):                # rank 0 creates and sizes the file, then the others map it
            if creator == (rank == 0):
                try:
                    frame_shared = torch.from_file(name[0], shared=True, size=W * H, dtype=torch.int32)
                    rc = torch.cuda.cudart().cudaHostRegister(frame_shared.data_ptr(), W * H * 4, 0)
                    if int(rc) != 0 or not frame_shared.is_pinned():
                        ok = 0
                except Exception as e:               # noqa: BLE001
                    print(f"[bench] shared frame buffer: {e}", file=sys.stderr)
                    ok = 0
            dist.barrier()
        okt = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        if int(okt.item()) == 0:
            frame_shared = None
        else:
            frame_shared = frame_shared.view(H, W)
            band = (rank * H // world, (rank + 1) * H // world)
        if rank == 0:
            try:
                os.unlink(name[0])                    # the mappings stay valid; nothing is left behind in /dev/shm
            except OSError:
                pass

    def step_e2e():
        if depth_tested:
            clear_in_step()
            if comp is not None:
                comp.barrier()
        if world == 1:
            v.setVertexAttribPointer(0, scene.stride, h_vert)      # host pointers: staged H2D by the library
            v.drawElements(scene.draw_mode, count, h_idx, wait=False)
        else:
            slot = up_state["k"] & 1
            if up_state["ptrs"][slot] is None:
                enqueue_upload(slot)                              # first step only: nothing was started ahead
            if args.pipeline_upload:
                enqueue_upload(slot ^ 1)                          # the next step's geometry, under this step's draw
            stream.wait_event(up_ready[slot])
            if idx_stream is not None:
                stream.wait_event(idx_ready[slot])
            pv, pi = up_state["ptrs"][slot]
            v.setVertexAttribPointer(0, scene.stride, pv, nbytes=scene.vertices.nbytes)
            v.drawElements(scene.draw_mode, count, pi, wait=False)
            comp.run(api.RT_COLOR)
            up_free[slot].record(stream)
            up_state["drawn"][slot] = True
            if not args.pipeline_upload:
                up_state["ptrs"][slot] = None                     # the next step uploads for itself
            up_state["k"] += 1
        # D2H of the step's result (the composed frame)
        if frame_shared is not None:
            frame_shared[band[0]:band[1]].copy_(targets[api.RT_COLOR][band[0]:band[1]], non_blocking=True)
        elif rank == 0:
            h_color.copy_(targets[api.RT_COLOR], non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = {"last": 0.0}

    def timed(fn, nsteps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        for _ in range(nsteps):
            fn()
        host_ms["last"] = (time.perf_counter() - t0) * 1e3 / max(1, nsteps)      # CPU time to enqueue one step
        e1.record(stream)
        torch.cuda.synchronize()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # one untimed draw sizes the scratch and picks the tile size; the composite needs it
    clear()
    step_resident()
    r.finish()
    tile = r.stats().last_tile_size
    if world > 1:
        if composite == "mirror":
            # fused composite: the tile kernel stores finished tiles into the peers' surfaces (CUDA IPC / NVLink);
            # per step only a barrier remains (the library's flag barrier with geometry shards, else one NCCL word)
            try:
                comp = TileMirror(r, api.RT_COLOR, targets[api.RT_COLOR].data_ptr(), rank, world, dev,
                                  peer_barrier=shards.barrier if shards is not None else None)
                clear()
                comp.barrier()
            except RuntimeError as e:                 # raised on every rank together (see TileMirror)
                print(f"[bench] {e}; using the NCCL all-gather composite", file=sys.stderr)
                composite = "nccl"
        if comp is None:
            comp = TileComposite(r, W, H, tile, rank, world, dev)

    # ---- fragments per draw (deterministic): summed over ranks
    clear()
    if isinstance(comp, TileMirror):
        comp.barrier()                    # no peer stores into a surface that is still being cleared
    r.resetStats()
    step_resident()
    r.finish()
    frag_t = torch.tensor([int(r.stats().fragments)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(frag_t)
        # after the composite every rank must hold the same, complete frame
        sums = torch.stack([targets[api.RT_COLOR].to(torch.int64).sum(), (targets[api.RT_COLOR] != 0).sum()])
        lo, hi = sums.clone(), sums.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if not torch.equal(lo, hi) and not os.environ.get("SWR_DEBUG_FAKE_PEERS"):
            raise SystemExit(f"composite differs between ranks: {lo.tolist()} vs {hi.tolist()}")
    fragments = int(frag_t.item())

    for _ in range(warmup):
        step_resident()
    torch.cuda.synchronize()

    # ---- value: K steps, inputs resident
    r.resetStats()
    sampler = ClockSampler(local_rank)
    if rank == 0 and with_clocks:
        sampler.start()
    ms_total = timed(step_resident, steps)
    if with_clocks:
        # keep the GPU under the same load a little longer so nvidia-smi (100 ms period) gets samples
        for _ in range(int(min(2000, max(0.0, 700.0 - ms_total) / max(ms_total / steps, 1e-3)))):
            step_resident()
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 and with_clocks else None

    # ---- per-kernel device times (CUDA events on the launching stream), outside the K-step timing
    r.resetStats()
    geom_ms, tile_ms = [], []
    nk = min(steps, 10)
    for _ in range(nk):
        step_resident()
        st = r.stats()
        geom_ms.append(st.last_geometry_ms)
        tile_ms.append(st.last_tile_ms)
    launches_per_step = int(r.stats().kernel_launches) // max(1, nk)
    passes_per_step = int(r.stats().passes) // max(1, nk)

    # ---- e2e: host buffers in, colour buffer out, every step
    ms_e2e = None
    if with_cpu is not None:
        for _ in range(2):
            step_e2e()
        r.resetStats()
        ms_e2e = timed(step_e2e, steps)
        host_e2e = host_ms["last"]
        h2d_lib = int(r.stats().h2d_bytes) // max(1, steps)       # what the library's staging really copied per step
        if frame_shared is not None:
            # the bands of all ranks make up the frame rank 0 holds on its device
            same = torch.tensor([1], dtype=torch.int32, device=dev)
            if rank == 0 and not torch.equal(frame_shared, targets[api.RT_COLOR].cpu()):
                same[0] = 0
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            if int(same.item()) == 0:
                raise SystemExit("e2e: the frame assembled in host memory from the ranks' bands differs from the device frame")

    res = None
    if rank == 0:
        peak, peak_src = load_peaks()
        ms_step = ms_total / steps
        geom_b, frag_b, v_ref = algorithmic_bytes(scene, fragments, b_frag)
        t_tile = float(np.mean(tile_ms)) * 1e-3
        t_geom = float(np.mean(geom_ms)) * 1e-3
        # dominant kernel = the tile kernel; its algorithmic traffic is the fragment traffic F*b_frag.
        # Per launch on this rank: the rank's share of the fragments (and of the passes of a multi-pass draw).
        ach_tile = (frag_b / world) / t_tile / 1e9
        ach_draw = (geom_b / (world if shards is not None else 1) + frag_b / world) / (ms_step * 1e-3) / 1e9
        traffic, traffic_note = ncu_traffic(workload, tile, world)
        in_bytes = scene.indices.nbytes + scene.vertices.nbytes
        if depth_tested:
            clear_note = "colour and depth are cleared inside every timed step (library fill kernels on the same stream; included in the step time)"
        else:
            clear_note = "render targets cleared once before timing: the pixel shader overwrites without reading (no depth test), every step shades the same fragments"
        par = "single GPU"
        if world > 1:
            par = (f"sort-first tiles x{world}; geometry " +
                   ("sharded by batches, records pushed to the tile owners over NVLink, flag barrier between the GPUs" if shards is not None else "replicated") +
                   "; composite: " + ("tile kernel stores to peer surfaces over NVLink + " + ("flag barrier" if shards is not None else "1-word NCCL barrier")
                                      if composite == "mirror" else "pack + NCCL all-gather + unpack"))
        res = {
            "metric": METRIC, "value": fragments / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "triangles_per_s": scene.num_primitives / (ms_step * 1e-3),
            "fragments_per_step": fragments, "triangles_per_step": scene.num_primitives,
            "config": {"workload": wl, "tile_size": tile, "parallelism": par,
                       "l2": (f"inputs larger than L2: {in_bytes / 1e6:.0f} MB of indices + vertices are re-read every step (126 MB L2)"
                              if in_bytes > 126e6 else f"inputs ({in_bytes / 1e6:.0f} MB) fit in L2 and no flush is done between steps: the figures are L2-warm"),
                       "clear": clear_note, "passes_per_step": passes_per_step},
            "roofline": {"bound": "hbm", "kernel": "tileKernel (record binning + coverage + shading of one screen tile per CTA)",
                         "achieved": ach_tile, "peak": peak, "unit": "GB/s", "frac": ach_tile / peak,
                         "traffic": traffic, "traffic_source": traffic_note,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": frag_b / world / max(1, passes_per_step),
                         "kernel_ms": t_tile * 1e3, "geometry_kernel_ms": t_geom * 1e3,
                         "kernel_ms_note": "device time from the first tile-phase launch to the end of the draw / of the geometry launches (CUDA events on the launching stream); in a multi-pass draw the two intervals interleave",
                         "draw": {"algorithmic_bytes": geom_b / (world if shards is not None else 1) + frag_b / world, "achieved": ach_draw,
                                  "frac": ach_draw / peak, "distinct_vertices": v_ref}},
            "gpu_launches": launches_per_step * steps,
            "clocks": clocks,
        }
        if ms_e2e is not None:
            res["e2e"] = {"value": fragments / (ms_e2e / steps * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / steps,
                          "host_enqueue_ms_per_step": host_e2e,       # host time inside the calls of a step: enqueueing, packing the
                                                                      # index slices, and waiting for a staging set of two steps ago
                          "host_binding": (f"each rank runs on the {len(args.numa_cpus)} CPUs next to its GPU (NVML affinity), its staging buffers are local to that socket"
                                           if getattr(args, "numa_cpus", None) else "none"),
                          "h2d_bytes_per_step": int(h2d_lib) if world == 1 else int((repl[0].h2d_bytes + (own_runs.numel() * 4 if own_runs is not None else 0)) * world),
                          "d2h_bytes_per_step": int(W * H * 4),
                          "input_bytes_per_step": int(in_bytes),
                          "path": ("host buffers -> swr_draw_elements (staged by the library, indices streamed pass by pass"
                                   + (", slices of local indices packed to 16 bits + block bases by host threads and widened on the device" if h2d_lib < in_bytes else "")
                                   + ") -> frame to host") if world == 1 else
                                  ((f"each rank uploads 1/{world} of the vertices (NCCL all-gather replicates them) and the index runs of its own batches"
                                    + (" (under the all-gather)" if idx_stream is not None else "") if shards is not None else
                                    f"each rank uploads 1/{world} of the geometry, NCCL all-gather replicates it") + ", draw + composite, "
                                   + (f"every rank copies 1/{world} of the rows of the composed frame into one page-locked shared host buffer" if frame_shared is not None
                                      else "rank 0 reads the frame"))}
        if with_cpu:
            cpu_fps, cpu_tps, cpu_desc, _ = cpu_reference_rate(scene, budget_s=20.0, steps=1)
            res["cpu_baseline"] = dict(cpu_desc, value=cpu_fps, unit=UNIT, triangles_per_s=cpu_tps)

    # tear down: the next workload gets a fresh context
    torch.cuda.synchronize()
    if frame_shared is not None:
        torch.cuda.cudart().cudaHostUnregister(frame_shared.data_ptr())
        frame_shared = None
    if isinstance(comp, TileMirror):
        comp.close()
    if shards is not None:
        shards.close()
    if world > 1:
        dist.barrier()                   # every peer has closed its mappings of this rank's surfaces / scratch
    r.close()
    del targets, d_vert, d_idx, h_vert, h_idx
    torch.cuda.empty_cache()
    return res


def bind_to_gpu_numa_node(torch, local_rank):
    """One process per GPU: run on the CPUs next to this process's GPU, so that the page-locked staging buffers it
    allocates (first touch) sit in the memory of that socket and its host <-> device copies do not cross the socket
    interconnect.  With 8 ranks uploading at once that link, not PCIe, is what saturates.  Best effort: returns the
    CPU set it bound to, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        p = torch.cuda.get_device_properties(local_rank)
        bus = f"{p.pci_domain_id:08x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        pynvml.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        print(f"[bench] rank on GPU {local_rank} ({bus}) bound to {len(cpus)} CPUs {cpus[0]}..{cpus[-1]}", file=sys.stderr)
        return cpus
    except Exception as e:                            # noqa: BLE001
        print(f"[bench] NUMA binding skipped: {e}", file=sys.stderr)
        return None


def run_gpu_arm(args):
    # NCCL and friends write banners straight to fd 1; the contract is ONE JSON line on stdout, so
    # everything else is sent to stderr and the line goes to the saved descriptor at the end.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    args.numa_cpus = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # everything (torch copies, NCCL, and the library's kernels) is enqueued on ONE non-default stream
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    line = bench_workload(args, args.workload, args.steps, args.warmup, rank, local_rank, world, dev, stream,
                          with_cpu=(world == 1 and not args.no_cpu), with_clocks=True)
    # BASELINE.json names configs[4] (50M textured triangles at 8K) as the 2/4/8-GPU configuration: with more than one
    # GPU the line also carries that workload, measured the same way with a few steps (it is 10x the frame of configs[2])
    extra = [w for w in args.also.split(",") if w] if args.also else (["c5"] if world > 1 and args.workload == "c3" else [])
    for w in extra:
        sub = bench_workload(args, w, max(3, min(args.steps, 5)), 3, rank, local_rank, world, dev, stream, with_cpu=None, with_clocks=False)
        if rank == 0:
            line.setdefault("also", {})[w] = sub
    if rank == 0:
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--pipeline-upload", type=int, default=1, help="N>1 end-to-end leg: upload step k+1 under the draw of step k")
    ap.add_argument("--e2e-overlap", type=int, default=1,
                    help="N>1 end-to-end leg: index upload under the vertex all-gather and the frame read back in bands by all ranks (0: the plain form)")
    ap.add_argument("--composite", default="mirror", choices=["mirror", "nccl"],
                    help="N>1: fused peer stores from the tile kernel (default) or pack + NCCL all-gather + unpack")
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--geometry", default="sharded", choices=["sharded", "replicated"],
                    help="N>1: vertex stage sharded by batches with records pushed to the tile owners (default) or run in full by every rank")
    ap.add_argument("--scratch-gb", type=float, default=0.0, help="N>1, sharded geometry: size of the shared scratch arena per rank (0 = by workload)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="--impl reference: seconds of CPU work for the whole run (bounded sample of the workload)")
    ap.add_argument("--also", default="", help="comma-separated further workloads measured with a few steps into line['also'] (default: c5 when N > 1); 'none' for none")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.also == "none":
        args.also = ","
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
