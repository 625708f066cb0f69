"""Quick GPU probe used during development: timings of the named workloads."""
import sys, time, json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from realtrace_b200 import api, scenes

def run(name, reps=5, flags=0, **kw):
    t0 = time.time()
    scene, cam, depth, desc = scenes.workload(name, **kw)
    t1 = time.time()
    ctx = api.Context(0)
    ctx.set_scene(scene)
    bs = ctx.commit()
    t2 = time.time()
    best = None
    for r in range(reps):
        tt = time.time()
        rgb, prim, t, st = ctx.render(cam, depth, aux=False, flags=flags)
        st["ms_wall"] = (time.time() - tt) * 1e3
        if best is None or st["ms_device"] < best["ms_device"]:
            best = st
    rays = best["rays_primary"] + best["rays_shadow"] + best["rays_secondary"]
    best["mrays_s_device"] = rays / best["ms_device"] / 1e3
    if best["node_visits"]:
        best["nodes_per_ray"] = best["node_visits"] / rays
        best["tris_per_ray"] = best["tri_tests"] / rays
    print(json.dumps({"workload": name, "scene_s": round(t1 - t0, 2), "commit_s": round(t2 - t1, 3), "build": bs, "best": best}), flush=True)
    ctx.close()
    return rgb

if __name__ == "__main__":
    names = sys.argv[1:] or ["bob1080", "synth1m", "blub4k"]
    for n in names:
        rgb = run(n)
        run(n, reps=1, flags=api.FLAG_COUNT_WORK)
        try:
            from PIL import Image
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            Image.fromarray(rgb[::-1]).resize((rgb.shape[1] // 4, rgb.shape[0] // 4)).save(os.path.join(ROOT, "gpurun_out", f"{n}.png"))
        except Exception as e:
            print("png failed", e)
