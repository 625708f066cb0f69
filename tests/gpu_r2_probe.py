"""Development probe (round 2): one process per knob setting; full frame and a 1/8 share on ONE GPU,
cold (L2 flushed before every frame) and warm, with the per-warp finish-time tail of the share.
usage: gpu_r2_probe.py            -> spawns the variants
       gpu_r2_probe.py --child    -> measures under the current environment"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# PROBE_LIBS: comma list of A/B builds ("" = the shipped library, "s1" = librt_variant_s1.so ...; build.build_variant),
# PROBE_WORKLOADS: comma list of workloads; every pair runs in its own process
LIBS = os.environ.get("PROBE_LIBS", "").split(",")
WORKLOADS = os.environ.get("PROBE_WORKLOADS", "synth1m").split(",")
VARIANTS = []
for _w in WORKLOADS:
    for _l in LIBS:
        _v = {"PROBE_WORKLOAD": _w}
        if _l:
            _v["RT_LIB_PATH"] = os.path.join(ROOT, "realtrace_b200", f"librt_variant_{_l}.so")
        VARIANTS.append(_v)


def child():
    import numpy as np
    import torch
    from realtrace_b200 import api, scenes
    name = os.environ.get("PROBE_WORKLOAD", "synth1m")
    scene, cam, depth, _ = scenes.workload(name)
    ctx = api.Context(0)
    ctx.set_scene(scene)
    ctx.commit()
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream or 1)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    W, H = cam.width, cam.height
    out = {"env": {k: os.path.basename(v) for k, v in os.environ.items() if k.startswith("RT_")}, "workload": name}
    for world, tile in ((1, (0, 0)), (8, (32, 16))):
        _, owned, tb = api.tile_layout(W, H, tile[0], tile[1], 0, world)
        buf = torch.zeros(max(owned * tb, W * H * 3), dtype=torch.uint8, device="cuda")
        fl = api.FLAG_PACKED_TILES if world > 1 else 0
        for _ in range(12):
            ctx.render_device(cam, depth, buf.data_ptr(), tile=tile, rank=0, world=world, flags=fl)
        for cold in (1, 0):
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(16)]
            for k, (a, b) in enumerate(evs):
                if cold:
                    flush.fill_(k)
                a.record(stream)
                ctx.render_device(cam, depth, buf.data_ptr(), tile=tile, rank=0, world=world, flags=fl, want_stats=False)
                b.record(stream)
            torch.cuda.synchronize()
            ctx.synchronize()
            t = np.array([a.elapsed_time(b) for a, b in evs])
            out[f"w{world}_{'cold' if cold else 'warm'}_ms"] = [round(float(np.median(t)), 4), round(float(t.min()), 4)]
        # warp finish times of the share (cold)
        flush.fill_(1)
        st = ctx.render_device(cam, depth, buf.data_ptr(), tile=tile, rank=0, world=world, flags=fl | api.FLAG_WARP_TIMES)
        raw = ctx.warp_times(1 << 17).astype(np.int64)
        if os.environ.get("RT_FRAME_KERNEL", "2") == "2":
            ph = raw.reshape(-1)[: (raw.size // 8) * 8].reshape(-1, 8)
            ph = ph[ph[:, 0] > 0]
            t0 = ph[:, 0].min()
            out[f"w{world}_phase_us"] = {name: {q: round(float(np.percentile((ph[:, k] - t0) / 1e3, q)), 1) for q in (0, 50, 99, 100)}
                                         for k, name in enumerate(["start", "traced", "barrier1", "shaded", "barrier2", "pushed", "exit"])
                                         if (ph[:, k] > 0).any()}
        else:
            t = raw[raw[:, 1] > 0]
            if len(t):
                end = (t[:, 1] - t[:, 0].min()) / 1e3
                out[f"w{world}_warp_end_us"] = {q: round(float(np.percentile(end, q)), 1) for q in (50, 90, 99, 100)}
        out[f"w{world}_stats"] = {k: round(st[k], 4) for k in ("ms_device", "ms_trace", "ms_shade")}
    # the multi-GPU frame step of a rank > 0 (render + push + its half of the handshake), frame and flags on this GPU
    world, tile, rank = 8, (32, 16), 1
    _, owned, tb = api.tile_layout(W, H, tile[0], tile[1], rank, world)
    packed = torch.zeros(owned * tb, dtype=torch.uint8, device="cuda")
    frame = torch.zeros(W * H * 3, dtype=torch.uint8, device="cuda")
    sync_ptr, _ = ctx.shared_buffer_create(1024)
    cs = api.camera_struct(cam)
    params = api.Context._params(depth, tile=tile, rank=rank, world=world, flags=api.FLAG_PACKED_TILES)
    k = 0
    for _ in range(12):
        ctx.peer_sync(sync_ptr, 0, world, k, 0)
        ctx.render_push(cs, params, packed.data_ptr(), frame.data_ptr(), sync_ptr, k)
        k += 1
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(16)]
    for i, (a, b) in enumerate(evs):
        flush.fill_(i)
        ctx.peer_sync(sync_ptr, 0, world, k, 0)
        a.record(stream)
        ctx.render_push(cs, params, packed.data_ptr(), frame.data_ptr(), sync_ptr, k)
        b.record(stream)
        k += 1
    torch.cuda.synchronize()
    ctx.synchronize()
    t = np.array([a.elapsed_time(b) for a, b in evs])
    out["w8_push_cold_ms"] = [round(float(np.median(t)), 4), round(float(t.min()), 4)]
    if os.environ.get("RT_FRAME_KERNEL", "2") == "2":
        flush.fill_(3)
        ctx.peer_sync(sync_ptr, 0, world, k, 0)
        params.flags |= api.FLAG_WARP_TIMES
        ctx.render_push(cs, params, packed.data_ptr(), frame.data_ptr(), sync_ptr, k)
        ctx.synchronize()
        raw = ctx.warp_times(1 << 17).astype(np.int64)
        ph = raw.reshape(-1)[: (raw.size // 8) * 8].reshape(-1, 8)
        ph = ph[ph[:, 0] > 0]
        t0 = ph[:, 0].min() if len(ph) else 0
        out["w8_push_phase_us"] = {name: {q: round(float(np.percentile((ph[:, kk] - t0) / 1e3, q)), 1) for q in (0, 50, 99, 100)}
                                   for kk, name in enumerate(["start", "traced", "barrier1", "shaded", "barrier2", "pushed", "exit"])
                                   if len(ph) and (ph[:, kk] > 0).any()}
    print(json.dumps(out), flush=True)
    ctx.close()


if __name__ == "__main__":
    if "--child" in sys.argv:
        child()
    else:
        for v in VARIANTS:
            env = dict(os.environ)
            env.update(v)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True,
                               timeout=300)
            print(r.stdout.strip() or f"FAILED {v}: {r.stderr[-800:]}", flush=True)
