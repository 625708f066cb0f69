"""Development probe: loop style / leaf size / refill sweeps, full frame and a 1/8 share."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
configs = [dict(RT_WIDE_BVH=0), dict(RT_WIDE_BVH=1), dict(RT_WIDE_BVH=2)]
code = f"""
import sys, os; sys.path.insert(0, {ROOT!r})
from realtrace_b200 import api, scenes
scene, cam, depth, desc = scenes.workload(os.environ.get('WL', 'synth1m'))
ctx = api.Context(0); ctx.set_scene(scene); ctx.commit()
out = []
for world in (1, 8):
    best = None
    for r in range(5):
        st = ctx.render(cam, depth, world=world, rank=0)[3]
        if best is None or st['ms_device'] < best['ms_device']: best = st
    out.append('w%d: %.3f (tr %.3f shade %.3f)' % (world, best['ms_device'], best['ms_trace'], best['ms_shade']))
print(os.environ.get('TAG'), ' | '.join(out), flush=True)
"""
for cfg in configs:
    env = dict(os.environ, TAG=json.dumps(cfg), **{k: str(v) for k, v in cfg.items()})
    subprocess.run([sys.executable, "-c", code], env=env)
