"""Development probe: per-rank frame time for world sizes 1..8 on ONE GPU (each rank's share is
independent, so this predicts the strong-scaling curve without 8 GPUs)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from realtrace_b200 import api, scenes
name = sys.argv[1] if len(sys.argv) > 1 else "synth1m"
scene, cam, depth, desc = scenes.workload(name)
ctx = api.Context(0); ctx.set_scene(scene); ctx.commit()
buf = torch.zeros(cam.width * cam.height * 3, dtype=torch.uint8, device="cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
ctx.set_stream(torch.cuda.current_stream().cuda_stream or 1)
for world in (1, 2, 4, 8):
    for rank in sorted({0, world - 1}):
        best = None
        for rep in range(6):
            flush.fill_(rep)
            st = ctx.render_device(cam, depth, buf.data_ptr(), rank=rank, world=world)
            if best is None or st["ms_device"] < best["ms_device"]:
                best = st
        # async timing of 10 frames with events
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ctx.render_device(cam, depth, buf.data_ptr(), rank=rank, world=world, want_stats=False)
        e1.record(); torch.cuda.synchronize()
        print(json.dumps({"world": world, "rank": rank, "ms_device": round(best["ms_device"], 4), "trace": round(best["ms_trace"], 4),
                          "shadow": round(best["ms_shadow"], 4), "shade": round(best["ms_shade"], 4), "resolve": round(best["ms_resolve"], 4),
                          "async_ms_per_frame": round(e0.elapsed_time(e1) / 10, 4), "ideal": round(0, 3)}), flush=True)
ctx.close()
