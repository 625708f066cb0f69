"""Development probe: time of a full BVH build (commit BUILD) and of a REFIT commit, cold (first build of the
context: allocations included) and warm, for the shipped library and for A/B builds.
usage: gpu_build_probe.py          -> one child per library in PROBE_LIBS ("" = shipped, "head" = librt_variant_head.so)
       gpu_build_probe.py --child  -> measures under the current environment
Every figure is the library's own CUDA-event pair around the build (rt_build_stats.ms_build / ms_refit) plus the host
wall clock of the commit call."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIBS = os.environ.get("PROBE_LIBS", ",head").split(",")
WORKLOADS = os.environ.get("PROBE_WORKLOADS", "synth1m,bob1080").split(",")


def child():
    import numpy as np
    import torch
    from realtrace_b200 import api, scenes
    out = {"lib": os.path.basename(os.environ.get("RT_LIB_PATH", "shipped"))}
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for name in WORKLOADS:
        scene, cam, depth, _ = scenes.workload(name)
        ctx = api.Context(0)
        ctx.set_scene(scene)
        t0 = time.perf_counter()
        cold = ctx.commit()
        wall_cold = (time.perf_counter() - t0) * 1e3
        nodes0, order0, keys0 = ctx.bvh_download()
        warm, wall, refit = [], [], []
        for k in range(8):
            flush.fill_(k)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            st = ctx.commit()
            wall.append((time.perf_counter() - t0) * 1e3)
            warm.append(st["ms_build"])
        nodes1, order1, keys1 = ctx.bvh_download()
        same = bool(np.array_equal(nodes0.view(np.uint32), nodes1.view(np.uint32)) and np.array_equal(order0, order1))
        for k in range(8):
            flush.fill_(k)
            torch.cuda.synchronize()
            st = ctx.commit(api.COMMIT_REFIT)
            refit.append(st["ms_refit"])
        nodes2, _, _ = ctx.bvh_download()
        same_refit = bool(np.array_equal(nodes0.view(np.uint32), nodes2.view(np.uint32)))
        import hashlib
        out[name] = {"n_tri": int(cold["n_triangles"]), "sort_passes": int(cold["sort_passes"]),
                     "ms_build_cold": round(cold["ms_build"], 3), "wall_cold_ms": round(wall_cold, 2),
                     "ms_build_warm_median": round(float(np.median(warm)), 3), "ms_build_warm_min": round(min(warm), 3),
                     "wall_warm_ms_median": round(float(np.median(wall)), 3),
                     "ms_refit_median": round(float(np.median(refit)), 3), "ms_refit_min": round(min(refit), 3),
                     "rebuild_identical": same, "refit_identical": same_refit,
                     "nodes_sha1": hashlib.sha1(nodes0.tobytes()).hexdigest()[:16],
                     "order_sha1": hashlib.sha1(order0.tobytes()).hexdigest()[:16]}
        ctx.close()
    print("PROBE " + json.dumps(out), flush=True)


if __name__ == "__main__":
    if "--child" in sys.argv:
        child()
    else:
        for lib in LIBS:
            env = dict(os.environ)
            if lib:
                env["RT_LIB_PATH"] = os.path.join(ROOT, "realtrace_b200", f"librt_variant_{lib}.so")
            subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, check=False)
