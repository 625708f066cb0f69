"""Development probe: one rank's share of the frame on ONE GPU, the way bench.py runs it at N > 1 (packed
tiles, asynchronous frames, L2 flushed before every frame), for several tile sizes.
usage: gpu_share_probe.py [workload] [world]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from realtrace_b200 import api, scenes
name = sys.argv[1] if len(sys.argv) > 1 else "synth1m"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
scene, cam, depth, desc = scenes.workload(name)
ctx = api.Context(0); ctx.set_scene(scene); ctx.commit()
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream or 1)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
W, H = cam.width, cam.height
for tile in ((64, 32), (32, 32), (32, 16), (16, 16), (128, 32)):
    for rank in (0, world // 2 + 1 if world > 2 else world - 1):
        _, owned, tb = api.tile_layout(W, H, tile[0], tile[1], rank, world)
        packed = torch.zeros(owned * tb, dtype=torch.uint8, device="cuda")
        for _ in range(10):
            st = ctx.render_device(cam, depth, packed.data_ptr(), tile=tile, rank=rank, world=world, flags=api.FLAG_PACKED_TILES)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(16)]
        for k, (a, b) in enumerate(evs):
            flush.fill_(k)
            a.record(stream)
            ctx.render_device(cam, depth, packed.data_ptr(), tile=tile, rank=rank, world=world, flags=api.FLAG_PACKED_TILES, want_stats=False)
            b.record(stream)
        torch.cuda.synchronize()
        t = np.array([a.elapsed_time(b) for a, b in evs])
        print(json.dumps({"tile": tile, "world": world, "rank": rank, "tiles": owned, "cold_ms_p50": round(float(np.median(t)), 4),
                          "cold_ms_max": round(float(t.max()), 4), "stats_frame": {k: round(st[k], 4) for k in ("ms_device", "ms_trace", "ms_shade", "ms_resolve")}}), flush=True)
ctx.close()
