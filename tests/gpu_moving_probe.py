"""Development probe: a rank's 1/8 share of the 4K frame on ONE GPU with a camera that moves 3 degrees per frame
(bench.py's moving_cameras), against the same share with a static camera — what the heavy-tiles-first order is worth
when it is a prediction.  usage: gpu_moving_probe.py   (PROBE_ENVS: ';'-separated JSON dicts of RT_* settings)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import numpy as np
    import torch
    import bench
    from realtrace_b200 import api, scenes
    scene, cam, depth, _ = scenes.workload("synth1m")
    ctx = api.Context(0)
    ctx.set_scene(scene)
    ctx.commit()
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream or 1)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    W, H = cam.width, cam.height
    out = {"env": {k: v for k, v in os.environ.items() if k.startswith("RT_")}}
    n = 24
    cams = {"static": [api.camera_struct(cam)] * n, "moving": [api.camera_struct(c) for c in bench.moving_cameras("synth1m", cam, n)]}
    for world, tile, ranks in ((8, (32, 16), (0, 7)), (1, (0, 0), (0,))):
        for rank in ranks:
            _, owned, tb = api.tile_layout(W, H, tile[0], tile[1], rank, world)
            buf = torch.zeros(max(owned * tb, W * H * 3), dtype=torch.uint8, device="cuda")
            fl = api.FLAG_PACKED_TILES if world > 1 else 0
            for kind in ("static", "moving"):
                for _ in range(4):
                    ctx.render_device(cams["static"][0], depth, buf.data_ptr(), tile=tile, rank=rank, world=world, flags=fl, want_stats=False)
                evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
                for k, (a, b) in enumerate(evs):
                    flush.fill_(k)
                    a.record(stream)
                    ctx.render_device(cams[kind][k], depth, buf.data_ptr(), tile=tile, rank=rank, world=world, flags=fl, want_stats=False)
                    b.record(stream)
                torch.cuda.synchronize()
                ctx.synchronize()
                t = np.array([a.elapsed_time(b) for a, b in evs])[4:]
                out[f"w{world}_r{rank}_{kind}"] = [round(float(t.mean()), 4), round(float(np.median(t)), 4), round(float(t.max()), 4)]
            del buf
    print(json.dumps(out), flush=True)
    ctx.close()


if __name__ == "__main__":
    if "--child" in sys.argv:
        child()
    else:
        for v in [json.loads(a) for a in os.environ.get("PROBE_ENVS", "{}").split(";")]:
            env = dict(os.environ)
            env.update(v)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True,
                               timeout=600)
            print(r.stdout.strip() or f"FAILED {v}: {r.stderr[-800:]}", flush=True)
