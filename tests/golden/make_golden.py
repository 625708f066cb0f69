#!/usr/bin/env python3
"""Generate the committed golden fixtures FROM THE REFERENCE BUILD (oracle/_ref).

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
Every case renders with the reference's own Serial sources (oracle/build_ref.py);
the port and the GPU path are then tested against these files, which travel to
machines where /root/reference does not exist.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import binding as ob  # noqa: E402
from oracle import build_ref  # noqa: E402
from realtrace_b200 import scenes  # noqa: E402
from cases import GOLDEN_CASES, build_case  # noqa: E402
import kat  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    if not build_ref.build(verbose=True):
        raise SystemExit("reference build unavailable")
    ref = ob.load_reference()
    pins = {"generator": ref.name, "cases": {}, "frames_640x480": {}}
    for name in GOLDEN_CASES:
        scene, cam, depth, mode = build_case(name)
        rgb, prim, t, info = ref.render(scene, cam, depth, mode)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), rgb=rgb, prim_id=prim, t=t)
        pins["cases"][name] = {"fnv1a64": ob.fnv1a64(rgb), "rays_total": int(info["rays_total"]),
                               "hit_pixels": int((prim >= 0).sum()), "depth": depth, "mode": mode,
                               "width": cam.width, "height": cam.height}
        print(name, pins["cases"][name])
    # full-size frame hashes (SURVEY Appendix B pins; lumina defaults, depth 10, as shipped)
    cam = scenes.stock_camera(640, 480)
    for key, (obj, cap) in {"bob2000": ("bob_tri.obj", 2000), "tetrahedron": ("tetrahedron.obj", None),
                            "bob_full": ("bob_tri.obj", None)}.items():
        rgb, prim, t, info = ref.render(scenes.obj_scene(obj, cap), cam, 10, ob.MODE_AS_SHIPPED, aux=False)
        pins["frames_640x480"][key] = {"fnv1a64": ob.fnv1a64(rgb), "rays_total": int(info["rays_total"])}
        print(key, pins["frames_640x480"][key])
    # known-answer rays
    names, rays = kat.kat_rays()
    s = kat.kat_scene()
    out = {}
    for mode_name, mode in (("as_shipped", ob.MODE_AS_SHIPPED), ("true_nearest", ob.MODE_TRUE_NEAREST)):
        prim, t = ref.trace_rays(s, rays, mode)
        out[mode_name] = {"prim": prim, "t": t}
    shade = ref.shade_rays(s, rays, 3, ob.MODE_TRUE_NEAREST)
    np.savez_compressed(os.path.join(HERE, "kat_rays.npz"), rays=rays, names=np.asarray(names),
                        prim_as_shipped=out["as_shipped"]["prim"], t_as_shipped=out["as_shipped"]["t"],
                        prim_true_nearest=out["true_nearest"]["prim"], t_true_nearest=out["true_nearest"]["t"],
                        shade_true_nearest=shade)
    for n, p, tt in zip(names, out["true_nearest"]["prim"], out["true_nearest"]["t"]):
        print(f"  {n:28s} prim {p:3d} t {tt:.6g}")
    with open(os.path.join(HERE, "pins.json"), "w") as f:
        json.dump(pins, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
