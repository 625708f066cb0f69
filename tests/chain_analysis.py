"""Development analysis (CPU emulation, no GPU): how long are the per-lane dependency chains of the fused primary
kernel on a workload, and what would shorten the longest 32-pixel batch?  For every 8x4 block of the full-size
frame: steps = node visits + triangle tests of a lane's primary ray (+ its shadow rays, which the fused kernel
traces in the same lane, one after the other).  usage: chain_analysis.py [workload] [scale]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from realtrace_b200 import scenes
import emul_binding

name = sys.argv[1] if len(sys.argv) > 1 else "synth1m"
scale = int(sys.argv[2]) if len(sys.argv) > 2 else 1
scene, cam, depth, desc = scenes.workload(name)
cam.width //= scale; cam.height //= scale
W, H = cam.width, cam.height

def blocks(a):                      # (H, W) -> (H/4, W/8, 32)
    h4, w8 = H // 4 * 4, W // 8 * 8
    return a[:h4, :w8].reshape(h4 // 4, 4, w8 // 8, 8).transpose(0, 2, 1, 3).reshape(h4 // 4, w8 // 8, 32)

for tag, env in (("binary tree", {}), ("4-wide view", {"EMUL_WIDE": "1"})):
    os.environ.pop("EMUL_WIDE", None); os.environ.update(env)
    t0 = time.time()
    nodes, tris, sh = emul_binding.Emulation().primary_cost(scene, cam, 1, shadows=True)
    p = blocks(nodes.astype(np.int64) + tris); s = blocks(sh.astype(np.int64))
    fused = (p + s).max(axis=2)                 # a lane walks its primary ray, then its shadow ray(s)
    work = (p + s).sum(axis=2)
    print(f"{name} {W}x{H} {tag}: {p.size} lanes, mean steps/lane {float((p + s).mean()):.1f}  ({time.time() - t0:.0f}s)")
    for label, c in (("primary + shadow chain", fused),):
        q = np.percentile(c, [50, 90, 99, 99.9, 99.99])
        print(f"   {label:24s} per-batch chain: p50 {q[0]:.0f} p90 {q[1]:.0f} p99 {q[2]:.0f} p99.9 {q[3]:.0f} p99.99 {q[4]:.0f} max {c.max()}")
    eff = work.sum() / (32.0 * fused.sum())
    print(f"   lane utilisation if every batch ran alone (sum of work / 32 x chain): {eff:.3f}")
    # batches longer than 2x / 4x the mean chain, and where they are
    m = fused.mean()
    print(f"   mean chain {m:.1f}; batches > 4x mean: {(fused > 4 * m).sum()} of {fused.size}; longest at block {np.unravel_index(fused.argmax(), fused.shape)}")
