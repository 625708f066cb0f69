"""Development probe: sweeps the runtime tuning knobs on the GPU (not part of the test suite)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
names = sys.argv[1:] or ["synth1m", "bob1080"]
for leaf in (4,):
    for refill in (1, 4, 8, 16, 24, 32):
        env = dict(os.environ, RT_LEAF_SIZE=str(leaf), RT_REFILL_PRIMARY=str(refill), RT_REFILL_QUEUE=str(refill), RT_REFILL_SHADOW=str(refill))
        code = f"""
import sys; sys.path.insert(0, {ROOT!r})
from realtrace_b200 import api, scenes
for name in {names!r}:
    scene, cam, depth, desc = scenes.workload(name)
    ctx = api.Context(0); ctx.set_scene(scene); ctx.commit()
    best = None
    for r in range(6):
        st = ctx.render(cam, depth)[3]
        if best is None or st['ms_device'] < best['ms_device']: best = st
    rays = best['rays_primary'] + best['rays_shadow'] + best['rays_secondary']
    print(name, 'leaf', {leaf}, 'refill', {refill}, 'ms %.3f trace %.3f shadow %.3f shade %.3f sec %.3f res %.3f  Mrays/s %.0f' % (best['ms_device'], best['ms_trace'], best['ms_shadow'], best['ms_shade'], best['ms_secondary'], best['ms_resolve'], rays / best['ms_device'] / 1e3), flush=True)
    ctx.close()
"""
        subprocess.run([sys.executable, "-c", code], env=env)
