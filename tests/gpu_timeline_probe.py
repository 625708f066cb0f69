"""Development probe: when do the warps of the primary traversal kernel finish?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from realtrace_b200 import api, scenes
scene, cam, depth, desc = scenes.workload("synth1m")
ctx = api.Context(0); ctx.set_scene(scene); ctx.commit()
buf = torch.zeros(cam.width * cam.height * 3, dtype=torch.uint8, device="cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
ctx.set_stream(torch.cuda.current_stream().cuda_stream or 1)
for world, cold in ((1, 0), (8, 0), (8, 1), (16, 0), (16, 1)):
    for rep in range(4):
        if cold:
            flush.fill_(rep)
        st = ctx.render_device(cam, depth, buf.data_ptr(), rank=0, world=world, flags=api.FLAG_WARP_TIMES)
    print(f"--- world {world} cold {cold}: ms_device {st['ms_device']:.4f} trace {st['ms_trace']:.4f} shade {st['ms_shade']:.4f} resolve {st['ms_resolve']:.4f}")
    t = ctx.warp_times().astype(np.int64)
    t = t[t[:, 1] > 0]
    t0 = t[:, 0].min()
    start, end = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3
    dur = end.max()
    print(f"world {world}: warps {len(t)}  kernel span {dur:.1f} us (ms_trace {st['ms_trace']*1e3:.1f})  start p50 {np.median(start):.1f} p99 {np.percentile(start,99):.1f} max {start.max():.1f}")
    qs = [1, 5, 10, 25, 50, 75, 90, 95, 99, 100]
    print("   warp finish time percentiles (us):", {q: round(float(np.percentile(end, q)), 1) for q in qs})
    busy = (end - start).sum() / (len(t) * dur)
    print(f"   mean warp residency / span = {busy:.3f}")
    # resident-warp count over time
    grid = np.linspace(0, dur, 21)
    alive = [(int(((start <= g) & (end > g)).sum())) for g in grid]
    print("   alive warps at 5% steps:", alive)
ctx.close()
