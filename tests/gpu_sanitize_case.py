"""Small end-to-end run for compute-sanitizer (one tool per gpurun call): build, render with bounces,
per-ray queries, tile sharding with stealing, refit."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from cases import build_case
from realtrace_b200 import api
import kat

for name in ("blubmixed_d5", "analytic_close_d5", "synth_small_d1", "bobtex_d3"):
    scene, cam, depth, _ = build_case(name)
    cam.width, cam.height = 104, 60
    ctx = api.Context(0)
    ctx.set_scene(scene)
    ctx.commit()
    a = ctx.render(cam, depth, aux=True)
    b = ctx.render(cam, depth, aux=True, flags=api.FLAG_COUNT_WORK)
    assert np.array_equal(a[0], b[0])
    ctx.commit(api.COMMIT_REFIT)
    c = ctx.render(cam, depth)
    assert np.array_equal(a[0], c[0])
    print(name, "ok", a[3]["rays_shadow"], a[3]["rays_secondary"], flush=True)
    ctx.close()
names, rays = kat.kat_rays()
ctx = api.Context(0); ctx.set_scene(kat.kat_scene()); ctx.commit()
print("kat", ctx.trace_rays(rays)[0][:8], ctx.shade_rays(rays, 3)[:2].tolist(), flush=True)
ctx.close()
# logical ranks with stealing into one frame (device buffer from the library itself)
scene, cam, depth, _ = build_case("blubmixed_d5")
cam.width, cam.height = 200, 100
ctx = api.Context(0); ctx.set_scene(scene); ctx.commit()
full = ctx.render(cam, depth)[0]
ptr, _ = ctx.shared_buffer_create(cam.width * cam.height * 3)
for r in range(3):
    ctx.render_device(cam, depth, ptr, rank=r, world=3, steal=(2, 0, None))
out = np.zeros_like(full); ctx.download(ptr, out)
assert np.array_equal(out, full)
print("stealing ok", flush=True)
ctx.close()
