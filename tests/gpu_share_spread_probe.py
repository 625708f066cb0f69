"""Development probe: every rank's share of an N-way split rendered alone on ONE GPU (L2 flushed before each frame), for
N = 1, 2, 4, 8 — how evenly the interleaved tiles load the ranks, and which build of k_frame (7 or 9 CTAs per SM,
RT_DENSE_MIN_PIXELS) suits which share size.  usage: gpu_share_spread_probe.py [workload]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


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
    out = {"env": {k: v for k, v in os.environ.items() if k.startswith("RT_")}, "workload": name}
    for world in (1, 2, 4, 8):
        tile = (0, 0) if world == 1 else ((32, 16) if world >= 4 else (64, 32))
        per_rank = []
        traced = []
        for rank in range(world):
            _, owned, tb = api.tile_layout(W, H, tile[0], tile[1], rank, world)
            buf = torch.zeros(max(owned * tb, W * H * 3), dtype=torch.uint8, device="cuda")
            fl = api.FLAG_PACKED_TILES if world > 1 else 0
            for _ in range(8):
                ctx.render_device(cam, depth, buf.data_ptr(), tile=tile, rank=rank, world=world, flags=fl)
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(12)]
            for k, (a, b) in enumerate(evs):
                flush.fill_(k)
                a.record(stream)
                ctx.render_device(cam, depth, buf.data_ptr(), tile=tile, rank=rank, world=world, flags=fl, want_stats=False)
                b.record(stream)
            torch.cuda.synchronize()
            ctx.synchronize()
            t = np.array([a.elapsed_time(b) for a, b in evs])
            per_rank.append(round(float(np.median(t)), 4))
            if world == 8:      # where the tracing phase of this share ends: median warp, 99th percentile, last warp (us)
                flush.fill_(1)
                ctx.render_device(cam, depth, buf.data_ptr(), tile=tile, rank=rank, world=world, flags=fl | api.FLAG_WARP_TIMES)
                raw = ctx.warp_times(1 << 17).astype(np.int64)
                ph = raw.reshape(-1)[: (raw.size // 8) * 8].reshape(-1, 8)
                ph = ph[ph[:, 0] > 0]
                if len(ph):
                    t0 = ph[:, 0].min()
                    traced.append([round(float(np.percentile((ph[:, 1] - t0) / 1e3, q)), 1) for q in (50, 99, 100)])
            del buf
        out[f"w{world}"] = per_rank
        if world == 8:
            out["w8_traced_us_p50_p99_last"] = traced
    print(json.dumps(out), flush=True)
    ctx.close()


if __name__ == "__main__":
    if "--child" in sys.argv:
        child()
    else:
        for v in [json.loads(a) for a in os.environ.get("PROBE_ENVS", "{}").split(";")]:
            env = dict(os.environ)
            env.update(v)
            if len(sys.argv) > 1:
                env["PROBE_WORKLOAD"] = sys.argv[1]
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True,
                               timeout=600)
            print(r.stdout.strip() or f"FAILED {v}: {r.stderr[-800:]}", flush=True)
