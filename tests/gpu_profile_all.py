"""One pass over every kernel of the library (for the ncu captures under profiles/): LBVH build + refit of the
1 M-triangle scene, a bounce-free 4K frame, a dielectric 4K frame (accumulators, k_paths, k_resolve), packed
tiles + assemble, tile feedback sort."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from realtrace_b200 import api, scenes

scene, cam, depth, _ = scenes.workload("synth1m")
ctx = api.Context(0)
ctx.set_scene(scene)
print("build", ctx.commit())
print("refit", ctx.commit(api.COMMIT_REFIT))
for _ in range(2):
    st = ctx.render(cam, depth)[3]
print("synth1m", st["ms_device"])
ptr, _ = ctx.shared_buffer_create(cam.width * cam.height * 3)
_, owned, tb = api.tile_layout(cam.width, cam.height, 0, 0, 0, 2)
packed, _ = ctx.shared_buffer_create(owned * tb)
ctx.render_device(cam, depth, packed, rank=0, world=2, flags=api.FLAG_PACKED_TILES)
ctx.assemble_tiles(packed, 0, 2, cam.width, cam.height, ptr)
ctx.synchronize()
ctx.close()
scene, cam, depth, _ = scenes.workload("blub4k")
ctx = api.Context(0)
ctx.set_scene(scene)
ctx.commit()
for _ in range(2):
    st = ctx.render(cam, depth)[3]
print("blub4k", st["ms_device"], st["rays_secondary"])
ctx.close()
