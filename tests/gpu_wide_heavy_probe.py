"""Development probe: 1/8 shares of the 4K synth1m frame on ONE GPU with the heaviest tiles on the 4-wide view
(RT_WIDE_HEAVY / RT_WIDE_HEAVY_DIV) against the binary walk, in the 7- and the 9-CTAs-per-SM build (RT_DENSE_MIN_PIXELS).
One process per setting; every setting must produce the same bytes (sha1 of the share).
usage: gpu_wide_heavy_probe.py          -> spawns the settings
       gpu_wide_heavy_probe.py --child  -> measures under the current environment"""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SETTINGS = [
    {"RT_WIDE_HEAVY": "0"},
    {"RT_WIDE_HEAVY": "1", "RT_WIDE_AFTER_BURSTS": "0"},
    {"RT_WIDE_HEAVY": "1", "RT_WIDE_AFTER_BURSTS": "1"},
    {"RT_WIDE_HEAVY": "1", "RT_WIDE_AFTER_BURSTS": "2"},
    {"RT_WIDE_HEAVY": "1", "RT_WIDE_AFTER_BURSTS": "1", "RT_WIDE_HEAVY_DIV": "65536"},
    {"RT_WIDE_HEAVY": "1", "RT_WIDE_AFTER_BURSTS": "1", "RT_LOOP_PRIMARY": "32"},
    {"RT_WIDE_HEAVY": "1", "RT_WIDE_AFTER_BURSTS": "1", "RT_DENSE_MIN_PIXELS": "0"},
]
if os.environ.get("PROBE_SETTINGS"):
    SETTINGS = [json.loads(a) for a in os.environ["PROBE_SETTINGS"].split(";")]
if os.environ.get("PROBE_HEAD"):
    SETTINGS.insert(0, {"RT_LIB_PATH": os.path.join(ROOT, "realtrace_b200", "librt_variant_head.so")})
RANKS = [int(x) for x in os.environ.get("PROBE_RANKS", "0,7").split(",")]


def phases(ctx, np):
    raw = ctx.warp_times(1 << 17).astype(np.int64)
    ph = raw.reshape(-1)[: (raw.size // 8) * 8].reshape(-1, 8)
    ph = ph[ph[:, 0] > 0]
    if not len(ph):
        return {}
    t0 = ph[:, 0].min()
    return {name: [round(float(np.percentile((ph[:, k] - t0) / 1e3, q)), 1) for q in (50, 99, 100)]
            for k, name in enumerate(["start", "traced", "barrier1", "shaded", "barrier2", "pushed", "exit"])
            if k in (1, 3, 6) and (ph[:, k] > 0).any()}


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
    out = {"env": {k: os.path.basename(v) for k, v in os.environ.items() if k.startswith("RT_")}}
    world, tile = 8, (32, 16)
    cs = api.camera_struct(cam)
    for rank in RANKS:
        _, owned, tb = api.tile_layout(W, H, tile[0], tile[1], rank, world)
        res = {}
        # (a) k_frame into a packed local buffer
        buf = torch.zeros(owned * tb, dtype=torch.uint8, device="cuda")
        for _ in range(12):
            ctx.render_device(cam, depth, buf.data_ptr(), tile=tile, rank=rank, world=world, flags=api.FLAG_PACKED_TILES)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(24)]
        for k, (a, b) in enumerate(evs):
            flush.fill_(k)
            a.record(stream)
            ctx.render_device(cam, depth, buf.data_ptr(), tile=tile, rank=rank, world=world, flags=api.FLAG_PACKED_TILES,
                              want_stats=False)
            b.record(stream)
        torch.cuda.synchronize()
        ctx.synchronize()
        t = np.array([a.elapsed_time(b) for a, b in evs])
        res["frame_ms"] = [round(float(np.median(t)), 4), round(float(t.min()), 4)]
        res["sha1"] = hashlib.sha1(buf.cpu().numpy().tobytes()).hexdigest()[:12]
        flush.fill_(1)
        ctx.render_device(cam, depth, buf.data_ptr(), tile=tile, rank=rank, world=world,
                          flags=api.FLAG_PACKED_TILES | api.FLAG_WARP_TIMES)
        res["frame_phase_us"] = phases(ctx, np)
        # (b) k_frame_push: the rank's step of the 8-GPU frame (frame and flags on this GPU); rank 0 would wait for the others
        if rank != 0:
            packed = torch.zeros(owned * tb, dtype=torch.uint8, device="cuda")
            frame = torch.zeros(W * H * 3, dtype=torch.uint8, device="cuda")
            sync_ptr, _ = ctx.shared_buffer_create(1024)
            params = api.Context._params(depth, tile=tile, rank=rank, world=world, flags=api.FLAG_PACKED_TILES)
            fi = 0
            for _ in range(12):
                ctx.peer_sync(sync_ptr, 0, world, fi, 0)
                ctx.render_push(cs, params, packed.data_ptr(), frame.data_ptr(), sync_ptr, fi)
                fi += 1
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(24)]
            for i, (a, b) in enumerate(evs):
                flush.fill_(i)
                ctx.peer_sync(sync_ptr, 0, world, fi, 0)
                a.record(stream)
                ctx.render_push(cs, params, packed.data_ptr(), frame.data_ptr(), sync_ptr, fi)
                b.record(stream)
                fi += 1
            torch.cuda.synchronize()
            ctx.synchronize()
            t = np.array([a.elapsed_time(b) for a, b in evs])
            res["push_ms"] = [round(float(np.median(t)), 4), round(float(t.min()), 4)]
            res["push_sha1"] = hashlib.sha1(frame.cpu().numpy().tobytes()).hexdigest()[:12]
            flush.fill_(3)
            ctx.peer_sync(sync_ptr, 0, world, fi, 0)
            params.flags |= api.FLAG_WARP_TIMES
            ctx.render_push(cs, params, packed.data_ptr(), frame.data_ptr(), sync_ptr, fi)
            ctx.synchronize()
            res["push_phase_us"] = phases(ctx, np)
        out[f"rank{rank}"] = res
    print("PROBE " + json.dumps(out), flush=True)
    ctx.close()


if __name__ == "__main__":
    if "--child" in sys.argv:
        child()
    else:
        for v in SETTINGS:
            env = dict(os.environ)
            env.update(v)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True,
                               timeout=300)
            print(r.stdout.strip() or f"FAILED {v}: {r.stderr[-1500:]}", flush=True)
