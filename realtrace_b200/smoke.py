"""__graft_entry__.smoke(): one small frame of the hot path on cuda:0, checked against the CPU oracle."""
import os
import sys

import numpy as np


def run():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from realtrace_b200 import api, scenes
    from oracle import binding as ob   # the checker (allowed here, see oracle/oracle_abi.h)
    import parity

    scene = scenes.bob_textured(max_faces=4000)
    cam = scenes.stock_camera(320, 240)
    ctx = api.Context(0)
    ctx.set_scene(scene)
    bs = ctx.commit()
    rgb, prim, t, st = ctx.render(cam, 3, aux=True)
    oracle = ob.best_available()
    ref_rgb, ref_prim, ref_t, info = oracle.render(scene, cam, 3, ob.MODE_TRUE_NEAREST)
    m = parity.compare(rgb, prim, t, ref_rgb, ref_prim, ref_t)
    parity.assert_parity(m, "smoke")
    print(f"smoke ok: oracle={oracle.name} build={bs} stats={st} parity={m}")
    ctx.close()
