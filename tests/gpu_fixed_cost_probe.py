"""Development probe: what part of a frame is fixed cost?  Tiny shares of the frame and an all-sky camera."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from realtrace_b200 import api, scenes
from realtrace_b200.scene import Camera
scene, cam, depth, desc = scenes.workload("synth1m")
ctx = api.Context(0); ctx.set_scene(scene); ctx.commit()
buf = torch.zeros(cam.width * cam.height * 3, dtype=torch.uint8, device="cuda")
def run(tag, cam, world, rank=0):
    best = None
    for rep in range(5):
        st = ctx.render_device(cam, depth, buf.data_ptr(), rank=rank, world=world)
        if best is None or st["ms_device"] < best["ms_device"]:
            best = st
    print(json.dumps({"case": tag, "world": world, "tiles": best["tiles"], "hits": best["max_queue"], "ms_device": round(best["ms_device"], 4), "trace": round(best["ms_trace"], 4),
                      "shadow": round(best["ms_shadow"], 4), "shade": round(best["ms_shade"], 4), "resolve": round(best["ms_resolve"], 4)}), flush=True)
for world in (8, 16, 64, 256, 1024, 4080):
    run("scene", cam, world)
sky = Camera(pos=cam.pos, target=(cam.pos[0], cam.pos[1] + 50, cam.pos[2] + 10), up=(0, 0, 1), fovy=45.0, width=cam.width, height=cam.height)
for world in (1, 8, 64):
    run("sky", sky, world)
# lower half of the image only (dense hits): a 3840x1080 crop is not expressible; use a camera looking straight down
down = Camera(pos=(50.0, 60.0, 50.0), target=(50.0, 10.0, 50.0), up=(0, 0, -1), fovy=45.0, width=cam.width, height=cam.height)
for world in (1, 2, 4, 8):
    run("down", down, world)
ctx.close()
