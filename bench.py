#!/usr/bin/env python3
"""bench.py — the hot path of realtrace_b200 measured on B200 (contract: see DESIGN.md §7).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload synth1m|bob1080|blub4k|analytic]
  python bench.py --impl reference ...      # the reference's own CPU renderer, same metric/config

One step = one frame of the workload.  Default workload: BASELINE.json configs[3], the synthetic
1 003 522-triangle sphere grid at 3840x2160, primary + shadow rays — the configuration the north
star quotes its throughput and scaling targets on; it fits one GPU.  configs[1] (bob 1080p) is
measured in the same run and reported under "other_workloads".

Metric: Mrays/s, one ray = one World::firstIntersection call of the reference (nearest-hit or
any-hit query): primary + shadow + secondary.
  value   frames rendered into a device buffer (scene + BVH resident in HBM, rt_render_device)
  e2e     the same through rt_render with HOST buffers: camera in, RGB8 frame copied to pinned host
          memory inside the timed region
N > 1 (torchrun, one process per GPU): the frame is cut into interleaved 64x32 screen tiles, every
rank renders its share with the scene replicated, NCCL gathers the packed tiles on rank 0, which
scatters them into the frame; "scaling": "strong" (fixed total work).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Mrays/s (primary+shadow+secondary)"
FMA_LANES_PER_CLK = 148 * 128          # FP32 lanes x SMs
L2_FLUSH_BYTES = 512 << 20


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler(threading.Thread):
    """Samples SM clock, power and throttle reasons through NVML (nvidia_ml_py) every few ms while the
    bench runs; the summary uses the samples that fall inside the timed region.  Falls back to polling
    nvidia-smi when NVML is not importable."""
    SMI_Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.t_begin = self.t_end = None
        self.source = None

    def _nvml_loop(self):
        import pynvml as nv
        nv.nvmlInit()
        idx = self.index
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                idx = int(vis.split(",")[self.index])
            except Exception:
                pass
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        sm_max = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        self.source = "nvml"
        while not self.stop_flag:
            r = get_reasons(h)
            self.rows.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), float(sm_max),
                              nv.nvmlDeviceGetPowerUsage(h) / 1000.0, [n for n, b in bits.items() if r & b]))
            time.sleep(0.01)

    def _smi_loop(self):
        self.source = "nvidia-smi"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.SMI_Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.rows.append((time.perf_counter(), float(f[0]), float(f[1]), float(f[2]),
                                      [n for k, n in enumerate(names) if f[3 + k].lower().startswith("active")]))
            except Exception:
                pass
            time.sleep(0.02)

    def run(self):
        try:
            self._nvml_loop()
        except Exception:
            try:
                self._smi_loop()
            except Exception:
                pass

    def stop(self):
        self.stop_flag = True

    def summary(self):
        rows = [r for r in self.rows if self.t_begin is not None and self.t_begin <= r[0] <= self.t_end]
        where = "timed region"
        if len(rows) < 2:
            rows, where = list(self.rows), "whole run (timed region shorter than the sampling period)"
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples (NVML and nvidia-smi unavailable)"]}
        sm = sorted(r[1] for r in rows)
        reasons = sorted({n for r in rows for n in r[4]})
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": rows[0][2], "reasons": reasons,
                "samples": len(rows), "window": where, "source": self.source, "power_w_max": max(r[3] for r in rows)}


def total_rays(st):
    return st["rays_primary"] + st["rays_shadow"] + st["rays_secondary"]


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own Serial renderer (oracle/_ref) or, where it was not built, the port.
def cpu_render_sample(scene, cam, depth, budget_s, threads):
    from oracle import binding as ob
    import __graft_entry__ as entry
    entry.build_oracle()
    orc = ob.best_available()
    W = cam.width
    # calibrate on a thin column sample, then size the real sample for ~budget_s of work
    step0 = max(1, W // 8)
    t0 = time.time()
    _, _, _, info = orc.render(scene, cam, depth, ob.MODE_AS_SHIPPED, col_begin=step0 // 2, col_step=step0,
                               nthreads=threads, aux=False)
    per_col = max(info["render_seconds"], 1e-4) / max(info["columns_rendered"], 1)
    build_s = info["build_seconds"]
    cols = int(min(W, max(8, budget_s / per_col)))
    col_step = max(1, W // cols)
    return orc, col_step, build_s


def cpu_step(orc, scene, cam, depth, col_step, threads):
    from oracle import binding as ob
    _, _, _, info = orc.render(scene, cam, depth, ob.MODE_AS_SHIPPED, col_begin=col_step // 2, col_step=col_step,
                               nthreads=threads, aux=False)
    return info


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from realtrace_b200 import scenes
    scene, cam, depth, desc = scenes.workload(args.workload)
    threads = os.cpu_count() or 1
    per_step_budget = max(2.0, 60.0 / max(args.steps + args.warmup, 1))
    orc, col_step, build_s = cpu_render_sample(scene, cam, depth, per_step_budget, threads)
    for _ in range(args.warmup):
        cpu_step(orc, scene, cam, depth, col_step, threads)
    rays, secs = 0, 0.0
    for _ in range(args.steps):
        info = cpu_step(orc, scene, cam, depth, col_step, threads)
        rays += info["rays_total"]
        secs += info["render_seconds"]
    value = rays / secs / 1e6
    cols = info["columns_rendered"]
    sample = (f"every {col_step}th column of the {cam.width}x{cam.height} frame ({cols} columns, "
              f"{info['rays_total']} rays per step), grid as shipped; scene+grid build {build_s:.2f}s excluded")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": base_config(args.workload, desc, cam, depth, scene),
            "rays_per_frame": {"total": float(info["rays_total"]) * col_step, "note": "sample x column step (estimate)"},
            "run": {"parallelism": f"{threads} host threads over columns", "l2": "n/a (CPU)"},
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": orc.kind, "sample": sample},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def base_config(workload, desc, cam, depth, scene):
    """The workload as both arms (ours / --impl reference) describe it: identical keys AND values."""
    return {"workload": workload, "description": desc, "width": cam.width, "height": cam.height, "max_depth": depth,
            "triangles": int(len(scene.tri_v)), "lights": int(len(scene.lights))}


def ncu_counts(workload):
    """Per-launch counters of the dominant kernel from the committed ncu capture (profiles/ncu_counts.json, written by
    profiles/summarize_ncu.py from `ncu --set full`): executed warp instructions, DRAM and L2 bytes."""
    path = os.path.join(ROOT, "profiles", "ncu_counts.json")
    if not os.path.exists(path):
        return None
    return json.load(open(path)).get(workload)


def moving_cameras(workload, cam, n):
    """n cameras on a circle around the workload's target, same height and distance: every timed frame sees the scene
    from a new direction (so the heavy-tiles-first order of the previous frame is only a prediction)."""
    import math
    from realtrace_b200.scene import Camera
    tx, ty, tz = cam.target
    dx, dz = cam.pos[0] - tx, cam.pos[2] - tz
    r, a0 = math.hypot(dx, dz), math.atan2(dx, dz)
    out = []
    for k in range(n):
        a = a0 + math.radians(3.0) * (k + 1)
        out.append(Camera(pos=(tx + r * math.sin(a), cam.pos[1], tz + r * math.cos(a)), target=cam.target, up=cam.up,
                          fovy=cam.fovy, width=cam.width, height=cam.height))
    return out


def measure_gpu(args, workload, steps, warmup, dist_ctx, main_leg):
    import torch
    from realtrace_b200 import api, multigpu, scenes

    rank, world = dist_ctx["rank"], dist_ctx["world"]
    dev = torch.device("cuda", dist_ctx["local_rank"])
    torch.cuda.set_device(dev)
    scene, cam, depth, desc = scenes.workload(workload)
    W, H = cam.width, cam.height

    def host_barrier():
        """All ranks line up WITHOUT device work (gloo): used around the legs in which rank 0 alone drives every GPU."""
        if world > 1:
            torch.distributed.barrier(group=dist_ctx["gloo"])

    ctx = api.Context(dist_ctx["local_rank"])
    stream = torch.cuda.current_stream()
    # run the library on torch's current stream so that torch CUDA events and NCCL collectives order
    # with its kernels; handle 0 is the legacy default stream, which the ABI names cudaStreamLegacy (0x1)
    ctx.set_stream(stream.cuda_stream or 1)
    t0 = time.time()
    ctx.set_scene(scene)
    bstats = ctx.commit()
    commit_s = time.time() - t0
    # a second, warm build: allocations and module loading are behind us
    bstats_warm = ctx.commit()

    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    # tile size of the interleaved ownership: the library default (64x32) up to 2 GPUs; from 4 GPUs on 32x16, whose
    # finer heavy-tiles-first order shortens the tail of the primary kernel (profiles/r1_tuning.md section 13)
    tile = (0, 0)
    if world > 1:
        tile = tuple(args.tile) if args.tile[0] > 0 else ((32, 16) if world >= 4 else (64, 32))
    _, owned, tile_bytes = api.tile_layout(W, H, tile[0], tile[1], rank, world)
    max_owned = max(api.tile_layout(W, H, tile[0], tile[1], r, world)[1] for r in range(world))
    frame = torch.zeros(H * W * 3, dtype=torch.uint8, device=dev) if world == 1 else None
    frame_ptr = frame.data_ptr() if world == 1 else None
    tick = torch.zeros(1, dtype=torch.int32, device=dev)

    # ---- N > 1: frame assembly.  Preferred: rank 0 owns the frame in a CUDA-IPC buffer that every
    # rank maps over NVLink; each rank stores its tiles straight into it and signals a flag in peer memory
    # (no collective).  Fallback (--assemble gather, or IPC unavailable): packed tiles + NCCL gather + scatter
    # kernel on rank 0.
    assemble = "single"
    packed = gathered = None
    cursor_ptr = sync_ptr = None
    frame_no = [0]
    if world > 1:
        assemble = args.assemble
        if assemble == "p2p":
            ok = 1
            try:
                hbuf = torch.zeros(64, dtype=torch.uint8, device=dev)
                if rank == 0:
                    frame_ptr, handle = ctx.shared_buffer_create(H * W * 3)
                    hbuf.copy_(torch.frombuffer(bytearray(handle), dtype=torch.uint8))
                torch.distributed.broadcast(hbuf, 0)
                if rank != 0:
                    frame_ptr = ctx.shared_buffer_open(hbuf.cpu().numpy().tobytes())
                # handshake buffer (arrival counters + consumed counter) on rank 0
                if rank == 0:
                    sync_ptr, handle = ctx.shared_buffer_create(1024)
                    hbuf.copy_(torch.frombuffer(bytearray(handle), dtype=torch.uint8))
                torch.distributed.broadcast(hbuf, 0)
                if rank != 0:
                    sync_ptr = ctx.shared_buffer_open(hbuf.cpu().numpy().tobytes())
                packed = torch.zeros(max_owned * tile_bytes, dtype=torch.uint8, device=dev)
                if args.steal_div > 0:                  # the shared tile-stealing cursor (64 x uint32) lives on rank 0
                    if rank == 0:
                        cursor_ptr, handle = ctx.shared_buffer_create(256)
                        hbuf.copy_(torch.frombuffer(bytearray(handle), dtype=torch.uint8))
                    torch.distributed.broadcast(hbuf, 0)
                    if rank != 0:
                        cursor_ptr = ctx.shared_buffer_open(hbuf.cpu().numpy().tobytes())
            except Exception as e:   # noqa: BLE001 — any failure means "use the gather path"
                print(f"[bench rank {rank}] CUDA IPC unavailable ({e}); falling back to gather", file=sys.stderr)
                ok = 0
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
            if int(flag[0]) == 0:
                assemble = "gather"
        if assemble == "gather":
            packed = torch.zeros(max_owned * tile_bytes, dtype=torch.uint8, device=dev)
            if rank == 0:
                gathered = [torch.zeros(max_owned * tile_bytes, dtype=torch.uint8, device=dev) for _ in range(world)]
                frame = torch.zeros(H * W * 3, dtype=torch.uint8, device=dev)
                frame_ptr = frame.data_ptr()

    push_params = api.Context._params(depth, tile=tile, rank=rank, world=world, flags=api.FLAG_PACKED_TILES)
    align_epoch = [0]

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def align_ranks():
        """Untimed line-up of the ranks on the device (peer-memory barrier kernel, or NCCL without IPC)."""
        if sync_ptr is not None:
            ctx.peer_barrier(sync_ptr, world, align_epoch[0])
            align_epoch[0] += 1
        else:
            torch.distributed.all_reduce(tick)

    orbit = workload == "orbit"
    orbit_k = [0]
    cam_py = cam
    cam = api.camera_struct(cam_py)      # flattened once: the per-frame host path is then a few ctypes calls
    orbit_cams = [api.camera_struct(scenes.orbit_camera(k, width=W, height=H)) for k in range(120)] if orbit else None

    def step_device(want_stats=False):
        """One frame with everything resident: render (+ assembly at N > 1).  Without stats the call only
        enqueues work on the stream.  Workload "orbit" (config 5): every frame first refits the LBVH
        (rt_scene_commit REFIT) and moves the camera one step along the 120-frame orbit."""
        nonlocal cam
        if orbit:
            ctx.commit(api.COMMIT_REFIT, want_stats=False)
            cam = orbit_cams[orbit_k[0] % 120]
            orbit_k[0] += 1
        if world == 1:
            return ctx.render_device(cam, depth, frame_ptr, want_stats=want_stats)
        if assemble == "p2p":
            k = frame_no[0]
            frame_no[0] += 1
            if cursor_ptr:      # stealing: any rank may render any pool tile -> resolve straight into the shared frame
                ctx.peer_sync(sync_ptr, rank, world, k, 0)
                st = ctx.render_device(cam, depth, frame_ptr, tile=tile, rank=rank, world=world, want_stats=want_stats,
                                       steal=(args.steal_div, k, cursor_ptr))
                ctx.peer_sync(sync_ptr, rank, world, k, 1)   # completion handshake: after it, rank 0 holds the frame
                return st
            # render + push into rank 0's frame + handshake: one library call, only enqueued
            ctx.render_push(cam, push_params, packed.data_ptr(), frame_ptr, sync_ptr, k)
            return None
        st = ctx.render_device(cam, depth, packed.data_ptr(), tile=tile, rank=rank, world=world, flags=api.FLAG_PACKED_TILES,
                               want_stats=want_stats)
        multigpu.gather_packed(packed, rank, world, gathered)
        if rank == 0:
            for r in range(world):
                ctx.assemble_tiles(gathered[r].data_ptr(), r, world, W, H, frame_ptr, tile=tile)
        return st

    def share_stats(flags=0):
        """Statistics of this rank's share of the frame (ray counts, per-kernel times): rendered into the local packed
        buffer, nothing pushed."""
        if world == 1:
            return ctx.render_device(cam, depth, frame_ptr, flags=flags)
        return ctx.render_device(cam, depth, packed.data_ptr(), tile=tile, rank=rank, world=world,
                                 flags=api.FLAG_PACKED_TILES | flags)

    # ---- value: device-resident frames, CUDA events around every step, L2 flushed between steps.
    # Warm-up frames run through the timed path; statistics come from separate share renders.
    sampler = ClockSampler(dist_ctx["local_rank"])
    sampler.start()
    stats = [share_stats() for _ in range(3)]
    for _ in range(max(warmup, 3) + 2):
        step_device(False)
    barrier()
    ctx.synchronize()

    def timed_loop(n_steps, cams=None):
        nonlocal cam
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        t_host0 = time.perf_counter()
        launches0 = ctx.frame_launches()
        for k in range(n_steps):
            flush.fill_(k & 0xff)
            if cams is not None:
                cam = cams[k]
            if world > 1:
                # untimed alignment: without it the L2 flush of a lagging rank (outside its own event pair) would be
                # counted inside the event pair of every rank that waits for it
                align_ranks()
            evs[k][0].record(stream)
            step_device(False)
            evs[k][1].record(stream)
        host_ms = (time.perf_counter() - t_host0) * 1e3 / n_steps
        barrier()
        ctx.synchronize()                                   # raises if any timed frame flagged an error
        per = np.array([a.elapsed_time(b) for a, b in evs])
        return per, host_ms, ctx.frame_launches() - launches0

    sampler.t_begin = time.perf_counter()
    per_step, host_enqueue_ms, launches = timed_loop(steps)
    sampler.t_end = time.perf_counter()
    ms = float(per_step.sum())
    rays_local = total_rays(stats[-1]) * steps          # static scene + camera: every frame casts the same rays
    if orbit:                                           # moving camera: count the rays of the timed frames exactly
        orbit_k[0] -= steps
        rays_local = 0
        for _ in range(steps):
            ctx.commit(api.COMMIT_REFIT, want_stats=False)
            cam = orbit_cams[orbit_k[0] % 120]
            orbit_k[0] += 1
            rays_local += total_rays(share_stats())
    ray_classes = np.array([stats[-1]["rays_primary"], stats[-1]["rays_shadow"], stats[-1]["rays_secondary"]], np.float64)
    if world > 1:
        t = torch.tensor([ms, float(rays_local), float(launches)], dtype=torch.float64, device=dev)
        tmax = t.clone()
        torch.distributed.all_reduce(tmax, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)
        ms, rays_total, launches = float(tmax[0]), float(t[1]), int(t[2])
        rc = torch.tensor(ray_classes, dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(rc, op=torch.distributed.ReduceOp.SUM)
        ray_classes = rc.cpu().numpy()
    else:
        rays_total = float(rays_local)
    value = rays_total / (ms * 1e-3) / 1e6

    # ---- the same with a moving camera (main leg, static workloads): no frame repeats the previous one
    moving = None
    if main_leg and not orbit:
        mcams_py = moving_cameras(workload, cam_py, steps)
        mcams = [api.camera_struct(c) for c in mcams_py]
        keep = cam
        m_rays = 0.0
        for c in mcams:
            cam = c
            m_rays += total_rays(share_stats())
        cam = keep
        barrier()
        m_per, _, _ = timed_loop(steps, mcams)
        cam = keep
        m_ms = float(m_per.sum())
        if world > 1:
            t = torch.tensor([m_ms, m_rays], dtype=torch.float64, device=dev)
            tmax = t.clone()
            torch.distributed.all_reduce(tmax, op=torch.distributed.ReduceOp.MAX)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)
            m_ms, m_rays = float(tmax[0]), float(t[1])
        moving = {"value": m_rays / (m_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": m_ms / steps,
                  "what": "the camera moves 3 degrees around the scene before every timed frame (no frame repeats, "
                          "the heavy-tiles-first order comes from a different view)"}

    # ---- N > 1: is the frame the ranks assembled in rank 0's memory the frame one GPU renders?  One more frame
    # through the timed path (untimed), then rank 0 renders the whole frame by itself and compares byte for byte.
    # Every collective sits outside the rank-0 block, so a failure in the check cannot hang the other ranks.
    frame_check = None
    if world > 1 or os.environ.get("RT_BENCH_FRAME_CHECK"):
        barrier()
        step_device(False)
        barrier()
        ctx.synchronize()
        if rank == 0:
            try:
                got = np.empty((H, W, 3), np.uint8)
                ctx.download(frame_ptr, got)
                alone = torch.zeros(H * W * 3, dtype=torch.uint8, device=dev)
                ctx.render_device(cam, depth, alone.data_ptr())
                want = alone.cpu().numpy().reshape(H, W, 3)
                bad = int((got != want).sum())
                frame_check = {"equal_to_the_1_gpu_frame": bad == 0, "mismatching_bytes": bad, "bytes": int(got.size)}
                if bad:
                    print(f"[bench] WARNING: the assembled frame differs from rank 0's own render in {bad} bytes", file=sys.stderr)
            except Exception as e:   # noqa: BLE001 — report, never hang or lose the measurement
                frame_check = {"error": repr(e)}
        barrier()

    # ---- N > 1 diagnostics (RT_BENCH_PHASES=1, with RT_PUSH_INLINE=0): per-rank phase times of k_frame inside a frame
    if world > 1 and assemble == "p2p" and not cursor_ptr and os.environ.get("RT_BENCH_PHASES"):
        pp = api.Context._params(depth, tile=tile, rank=rank, world=world, flags=api.FLAG_PACKED_TILES | api.FLAG_WARP_TIMES)
        rows = []
        for f_ in range(6):
            k = frame_no[0]
            frame_no[0] += 1
            flush.fill_(f_ & 0xff)
            align_ranks()
            ctx.render_push(cam, pp, packed.data_ptr(), frame_ptr, sync_ptr, k)
            barrier()
            ctx.synchronize()
            raw = ctx.warp_times(1 << 17).astype(np.int64)
            ph = raw.reshape(-1)[: (raw.size // 8) * 8].reshape(-1, 8)
            ph = ph[ph[:, 0] > 0]
            if len(ph):
                t0_ = ph[:, 0].min()
                rows.append([float((ph[:, kk].max() - t0_) / 1e3) if (ph[:, kk] > 0).any() else 0.0 for kk in range(7)] +
                            [float(np.median(ph[:, 1] - t0_) / 1e3)])
        if rows:
            r_ = np.median(np.array(rows[1:]), axis=0)
            print(f"[phases rank {rank}] traced p50 {r_[7]:.1f} last {r_[1]:.1f} | barrier1 (inline push: frame open) {r_[2]:.1f} | shaded {r_[3]:.1f} | barrier2 {r_[4]:.1f} | "
                  f"pushed {r_[5]:.1f} | exit {r_[6]:.1f} us", file=sys.stderr, flush=True)

    # ---- work counts + per-kernel times of this rank's share (roofline inputs)
    cnt = share_stats(api.FLAG_COUNT_WORK)
    share = [share_stats() for _ in range(5)]
    kernel_ms = {"primary wave: trace + fused shadow rays (+ shading when the frame is one k_frame launch)":
                 float(np.median([s_["ms_trace"] for s_ in share])),
                 "k_traverse<shadow> (any hit)": float(np.median([s_["ms_shadow"] for s_ in share])),
                 "k_shade<primary>": float(np.median([s_["ms_shade"] for s_ in share])),
                 "k_paths (all bounce generations)": float(np.median([s_["ms_secondary"] for s_ in share])),
                 "k_resolve": float(np.median([s_["ms_resolve"] for s_ in share]))}
    n_rays_share = total_rays(cnt)
    work = {"rays": n_rays_share, "node_visits": cnt["node_visits"] + cnt["shadow_node_visits"],
            "tri_tests": cnt["tri_tests"] + cnt["shadow_tri_tests"],
            "nearest": {"rays": cnt["rays_primary"] + cnt["rays_secondary"], "node_visits": cnt["node_visits"], "tri_tests": cnt["tri_tests"]},
            "shadow": {"rays": cnt["rays_shadow"], "node_visits": cnt["shadow_node_visits"], "tri_tests": cnt["shadow_tri_tests"]},
            "share_ms_device": float(np.median([s_["ms_device"] for s_ in share]))}

    micro = None
    if main_leg:
        micro = ctx.microbench()

    out = {"value": value, "ms_per_step": ms / steps,
           "ms_per_step_p50": float(np.percentile(per_step, 50)), "ms_per_step_p99": float(np.percentile(per_step, 99)),
           "launches": launches, "host_enqueue_ms_per_step": host_enqueue_ms, "desc": desc, "assemble": assemble,
           "tile": (tile[0] or 64, tile[1] or 32), "frame_check": frame_check, "cam": cam_py, "depth": depth,
           "scene": scene, "build": bstats, "build_warm": bstats_warm, "commit_s": commit_s, "rays_per_frame": rays_total / steps,
           "ray_classes": [float(x) for x in ray_classes], "kernel_ms": kernel_ms, "work": work, "micro": micro, "moving": moving}

    # ---- e2e: the public host-buffer API.  Frames go through rt_render_enqueue / rt_render_wait into two page-locked
    # host buffers (camera in, RGB8 frame out, two frames in flight: the copy of frame k overlaps the rendering of
    # frame k+1); `latency` is one synchronous rt_render.  At N > 1 rank 0 drives ALL N GPUs through ONE multi-device
    # context (rt_create_multi — what RenderEngine::render() does); the other ranks wait on the host meanwhile.
    barrier()
    ctx.synchronize()
    e2e = None
    host_barrier()
    if rank == 0:
        ectx = ctx
        if world > 1:
            ectx = api.Context(devices=list(range(world)))
            ectx.set_scene(scene)
            ectx.commit()
        bufs = [api.host_alloc(H * W * 3), api.host_alloc(H * W * 3)]
        ecams = [cam] if not orbit else orbit_cams

        def e2e_frame(k, slot):
            if orbit:
                ectx.commit(api.COMMIT_REFIT, want_stats=False)
            ectx.render_enqueue(ecams[k % len(ecams)], depth, bufs[slot], slot)

        for _ in range(max(2, warmup)):                 # warm: frames, allocations, pinned registrations
            st_e = ectx.render(ecams[0], depth, out=bufs[0].reshape(H, W, 3))[3]
        e_rays_frame = total_rays(st_e)
        lat = []
        for _ in range(5):
            t0 = time.perf_counter()
            ectx.render(ecams[0], depth, out=bufs[0].reshape(H, W, 3))
            lat.append((time.perf_counter() - t0) * 1e3)
        e2e_steps = max(steps, 10)
        for slot in (0, 1, 0, 1):                       # both frame slots warm
            e2e_frame(0, slot)
            ectx.render_wait(slot)
        t0 = time.perf_counter()
        e2e_frame(0, 0)
        for k in range(1, e2e_steps + 1):
            if k < e2e_steps:
                e2e_frame(k, k % 2)
            ectx.render_wait((k - 1) % 2)
        e_s = time.perf_counter() - t0
        e_rays = e_rays_frame * e2e_steps
        if orbit:
            e_rays = rays_total / steps * e2e_steps      # (per-frame counts differ slightly along the orbit)
        same = None
        if world > 1 and frame_check is not None and "error" not in frame_check:
            one = np.empty((H, W, 3), np.uint8)
            ctx.render(cam, depth, out=one)
            ectx.render(ecams[0], depth, out=bufs[1].reshape(H, W, 3))
            same = bool(np.array_equal(one, bufs[1].reshape(H, W, 3)))
        e2e = {"value": e_rays / e_s / 1e6, "ms_per_step": e_s / e2e_steps * 1e3, "latency_ms": float(np.median(lat)),
               "steps": e2e_steps, "devices": ectx.device_count(), "equal_to_the_1_gpu_frame": same}
        if world == 1 and main_leg:
            # cold e2e: scene upload + LBVH build + one frame to the host
            t0 = time.perf_counter()
            ctx.set_scene(scene)
            ctx.commit()
            ctx.render(cam, depth, out=bufs[0].reshape(H, W, 3))
            e2e["cold_ms_incl_scene_upload_and_bvh_build"] = (time.perf_counter() - t0) * 1e3
        if ectx is not ctx:
            ectx.close()
        for b in bufs:
            api.host_free(b)
    host_barrier()
    out["e2e"] = e2e
    sampler.stop()
    out["clocks"] = sampler.summary()
    out["ms_refit"] = ctx.build_stats().get("ms_refit")
    ctx.close()
    del flush
    return out


def roofline_of(main, workload, hbm_gbs, peak_src, sm_max_mhz):
    """max(T_issue, T_fma, T_l2, T_hbm) / T_kernel for this rank's share."""
    work, kms, micro = main["work"], main["kernel_ms"], main["micro"]
    dom = max(kms, key=kms.get)
    t_kernel = kms[dom] * 1e-3
    n_rays = work["rays"]
    nodes_per_ray = work["node_visits"] / max(n_rays, 1)
    tris_per_ray = work["tri_tests"] / max(n_rays, 1)
    frame_ms = work["share_ms_device"]
    # algorithmic work of ALL rays of the share against the time of ALL its traversal kernels
    t_all = sum(kms.values()) * 1e-3
    bytes_per_ray = 64.0 * nodes_per_ray + 48.0 * tris_per_ray + 84.0
    fma_per_ray = 12.0 * nodes_per_ray + 30.0 * tris_per_ray + 60.0
    alg_bytes, alg_fma = bytes_per_ray * n_rays, fma_per_ray * n_rays
    fma_peak = micro["fma_lane_instr_per_s"] if micro else FMA_LANES_PER_CLK * sm_max_mhz * 1e6
    issue_peak = micro["issue_warp_instr_per_s"] if micro else fma_peak / 32.0
    l2_peak = micro["l2_read_gbs"] * 1e9 if micro else None
    terms = {"T_kernels_ms": t_all * 1e3, "T_fma_ms": alg_fma / fma_peak * 1e3}
    nc = ncu_counts(workload)
    traffic = None
    if nc and nc.get("rays"):
        # the capture is of a whole 1-GPU frame: a rank's share executes its part of it
        k = n_rays / float(nc["rays"])
        nc = {**nc, **{key: nc[key] * k for key in ("inst_executed", "lts_bytes", "dram_bytes") if nc.get(key)}, "scaled_by": k}
    if nc:
        traffic = nc.get("dram_bytes")
        if nc.get("inst_executed"):
            terms["T_issue_ms"] = nc["inst_executed"] / issue_peak * 1e3
        if nc.get("lts_bytes") and l2_peak:
            terms["T_l2_ms"] = nc["lts_bytes"] / l2_peak * 1e3
        if traffic:
            terms["T_hbm_ms"] = traffic / (hbm_gbs * 1e9) * 1e3
    cand = {k: v for k, v in terms.items() if k != "T_kernels_ms"}
    bound_key = max(cand, key=cand.get)
    bound = {"T_issue_ms": "issue", "T_fma_ms": "fma", "T_l2_ms": "l2", "T_hbm_ms": "hbm"}[bound_key]
    frac = cand[bound_key] / terms["T_kernels_ms"]
    if bound == "issue":
        achieved, peak, unit = nc["inst_executed"] / t_all / 1e9, issue_peak / 1e9, "Gwarp-instr/s"
    elif bound == "fma":
        achieved, peak, unit = alg_fma / t_all / 1e12, fma_peak / 1e12, "T lane-instr/s"
    elif bound == "l2":
        achieved, peak, unit = nc["lts_bytes"] / t_all / 1e9, l2_peak / 1e9, "GB/s"
    else:
        achieved, peak, unit = traffic / t_all / 1e9, hbm_gbs, "GB/s"
    return {"bound": bound, "achieved": achieved, "peak": peak, "unit": unit, "frac": frac, "traffic": traffic,
            "peak_source": ("rt_microbench on this GPU in this run" if micro else "nominal") + "; HBM: " + peak_src,
            "terms": terms, "dominant_kernel": dom, "kernel_ms": kms, "share_ms_device": frame_ms,
            "algorithmic": {"rays": n_rays, "node_visits_per_ray": nodes_per_ray, "tri_tests_per_ray": tris_per_ray,
                            "bytes_per_ray": bytes_per_ray, "fma_lane_instr_per_ray": fma_per_ray,
                            "fma_frac_of_measured_peak": alg_fma / t_all / fma_peak,
                            "hbm_yardstick": {"achieved_gbs": alg_bytes / t_all / 1e9, "peak_gbs": hbm_gbs,
                                              "frac": alg_bytes / t_all / 1e9 / hbm_gbs,
                                              "note": "SURVEY 8(d)'s algorithmic bytes over the HBM copy peak: NOT a physical "
                                                      "fraction here — neighbouring rays share nodes through L1/L2, DRAM moves "
                                                      "only `traffic` bytes per launch"}},
            "measured_peaks": micro,
            "ncu": nc,
            "note": "frac = max(T_issue, T_fma, T_l2, T_hbm) / T_kernels over this rank's share of the frame: T_issue = warp "
                    "instructions executed (ncu capture in profiles/) / measured issue rate, T_fma = SURVEY 8(d)'s algorithmic "
                    "FMA lane-instructions / measured FMA rate, T_l2 and T_hbm = bytes the capture saw at L2 / DRAM over the "
                    "measured bandwidths.  Traversal is bound by instruction issue and dependent-fetch latency, not by HBM."}


def run_gpu_arm(args):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU arm)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    gloo = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        gloo = torch.distributed.new_group(backend="gloo")
    dist_ctx = {"rank": rank, "world": world, "local_rank": local_rank, "gloo": gloo}
    hbm_gbs, peak_src, sm_max_mhz = peaks()

    main = measure_gpu(args, args.workload, args.steps, args.warmup, dist_ctx, True)
    cam, scene, depth = main["cam"], main["scene"], main["depth"]
    rp, rs, rsec = main["ray_classes"]

    line = {"metric": METRIC, "value": main["value"], "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "ms_per_step_p50": main["ms_per_step_p50"],
            "ms_per_step_p99": main["ms_per_step_p99"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": base_config(args.workload, main["desc"], cam, depth, scene),
            "rays_per_frame": {"total": main["rays_per_frame"], "R_primary": rp, "R_shadow": rs, "R_secondary": rsec},
            "run": {"parallelism": (f"{world} GPUs, interleaved {main['tile'][0]}x{main['tile'][1]} tiles, scene replicated, frame assembly: "
                                    + {"p2p": "every rank copies its tiles into rank 0's frame over NVLink (CUDA IPC peer stores) from inside "
                                              "its one frame kernel on bounce-free scenes (else a push kernel per rank) and signals a flag in "
                                              "peer memory; no collective"
                                              + (f", dynamic tile stealing (1/{args.steal_div} of the tile groups pooled)" if args.steal_div > 0 else ""),
                                       "gather": "packed tiles + NCCL gather + scatter kernel"}[main["assemble"]])
                    if world > 1 else "1 GPU",
                    "l2": f"L2 flushed between timed steps ({L2_FLUSH_BYTES >> 20} MiB write)"},
            "gpu_launches": main["launches"], "clocks": main["clocks"],
            **({"frame_check": main["frame_check"]} if main.get("frame_check") is not None else {}),
            "host_enqueue_ms_per_step": main["host_enqueue_ms_per_step"],
            "build": {"commit_s": main["commit_s"], **main["build"], "ms_build_warm": main["build_warm"]["ms_build"]}}
    if main["moving"]:
        line["moving_camera"] = main["moving"]
    if rank == 0:
        e = main["e2e"]
        line["e2e"] = {"value": e["value"], "unit": "Mrays/s", "ms_per_step": e["ms_per_step"], "latency_ms": e["latency_ms"],
                       "h2d_bytes_per_step": 64 + 40, "d2h_bytes_per_step": cam.width * cam.height * 3,
                       "devices": e["devices"],
                       **({"cold_ms_incl_scene_upload_and_bvh_build": e["cold_ms_incl_scene_upload_and_bvh_build"]}
                          if "cold_ms_incl_scene_upload_and_bvh_build" in e else {}),
                       **({"equal_to_the_1_gpu_frame": e["equal_to_the_1_gpu_frame"]} if e.get("equal_to_the_1_gpu_frame") is not None else {}),
                       "what": "rt_render_enqueue / rt_render_wait: camera + params in, RGB8 frame into page-locked host memory, two "
                               "frames in flight (the copy of frame k overlaps the rendering of frame k+1); latency_ms = one "
                               "synchronous rt_render"
                               + ("" if world == 1 else f"; ONE multi-device context (rt_create_multi) on rank 0 drives all {world} GPUs: "
                                  "frame striped over the GPUs' memories, every GPU copies its granules over its own PCIe link")}
        line["roofline"] = roofline_of(main, args.workload, hbm_gbs, peak_src, sm_max_mhz)

    if world == 1 and not args.no_cpu:
        # the reference's CPU renderer on a bounded sample of the same workload, 1 core
        orc, col_step, build_s = cpu_render_sample(scene, cam, depth, 12.0, 1)
        info = cpu_step(orc, scene, cam, depth, col_step, 1)
        line["cpu_baseline"] = {"value": info["rays_total"] / info["render_seconds"] / 1e6, "unit": "Mrays/s",
                                "cores": 1, "kind": orc.kind,
                                "sample": f"every {col_step}th column of the {cam.width}x{cam.height} frame "
                                          f"({info['columns_rendered']} columns, {info['rays_total']} rays, "
                                          f"{info['render_seconds']:.1f}s), grid as shipped; build {build_s:.2f}s excluded"}
    if not args.no_others and args.workload == "synth1m":
        others = {}
        for name in ("bob1080", "blub4k", "orbit"):
            o = measure_gpu(args, name, max(5, min(args.steps, 10)), args.warmup, dist_ctx, False)
            ent = {"value": o["value"], "unit": "Mrays/s", "ms_per_step": o["ms_per_step"], "rays_per_frame": o["rays_per_frame"],
                   "description": o["desc"], "kernel_ms": o["kernel_ms"], "frame_check": o["frame_check"]}
            if rank == 0:
                ent["e2e_value"] = o["e2e"]["value"]
                ent["e2e_ms_per_step"] = o["e2e"]["ms_per_step"]
            if name == "orbit":
                ent["ms_refit"] = o["ms_refit"]
            if world == 1 and not args.no_cpu and name == "bob1080":
                sc, cm, dp = o["scene"], o["cam"], o["depth"]
                orc, col_step, build_s = cpu_render_sample(sc, cm, dp, 6.0, 1)
                info = cpu_step(orc, sc, cm, dp, col_step, 1)
                ent["cpu_baseline"] = {"value": info["rays_total"] / info["render_seconds"] / 1e6, "unit": "Mrays/s", "cores": 1,
                                       "kind": orc.kind, "sample": f"every {col_step}th column ({info['rays_total']} rays, "
                                                                    f"{info['render_seconds']:.1f}s), grid as shipped"}
            others[name] = ent
        line["other_workloads"] = others
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="synth1m")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-others", action="store_true", help="skip the secondary workloads")
    ap.add_argument("--assemble", default="p2p", choices=["p2p", "gather"], help="N > 1 frame assembly")
    ap.add_argument("--tile", type=int, nargs=2, default=[0, 0], metavar=("W", "H"),
                    help="N > 1: screen tile size of the interleaved ownership (0 0 = 64x32 at N=2, 32x16 from N=4)")
    ap.add_argument("--steal-div", type=int, default=0,
                    help="N > 1 with p2p assembly: every k-th tile group forms the shared pool ranks steal from (0 = off)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        import __graft_entry__ as entry
        if not os.path.exists(os.path.join(ROOT, "realtrace_b200", "librealtrace_b200.so")):
            entry.build()
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
