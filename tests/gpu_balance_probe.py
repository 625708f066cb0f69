"""Development probe: cold frame time of EVERY rank's share on one GPU (how well does the tile interleaving
balance the ranks?).  usage: gpu_balance_probe.py [workload] [world] [tile_w tile_h]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from realtrace_b200 import api, scenes
name = sys.argv[1] if len(sys.argv) > 1 else "synth1m"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
tile = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (32, 16)
scene, cam, depth, desc = scenes.workload(name)
ctx = api.Context(0); ctx.set_scene(scene); ctx.commit()
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream or 1)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
W, H = cam.width, cam.height
res = []
for rank in range(world):
    _, owned, tb = api.tile_layout(W, H, tile[0], tile[1], rank, world)
    packed = torch.zeros(owned * tb, dtype=torch.uint8, device="cuda")
    for _ in range(10):
        st = ctx.render_device(cam, depth, packed.data_ptr(), tile=tile, rank=rank, world=world, flags=api.FLAG_PACKED_TILES)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(12)]
    for k, (a, b) in enumerate(evs):
        flush.fill_(k)
        a.record(stream)
        ctx.render_device(cam, depth, packed.data_ptr(), tile=tile, rank=rank, world=world, flags=api.FLAG_PACKED_TILES, want_stats=False)
        b.record(stream)
    torch.cuda.synchronize()
    t = float(np.median([a.elapsed_time(b) for a, b in evs]))
    rays = st["rays_primary"] + st["rays_shadow"]
    res.append((t, rays))
ts = np.array([r[0] for r in res]); rays = np.array([r[1] for r in res], dtype=np.float64)
print(json.dumps({"tile": tile, "world": world, "cold_ms": [round(x, 4) for x in ts], "max/mean": round(float(ts.max() / ts.mean()), 4),
                  "rays max/mean": round(float(rays.max() / rays.mean()), 4)}))
ctx.close()
