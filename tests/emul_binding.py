"""Builds and binds tests/emul/libemul.so — the TEST-ONLY host emulation of the GPU pipeline."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle import binding as ob
from realtrace_b200 import api

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "emul.cu")
LIB = os.path.join(HERE, "emul", "libemul.so")
ROOT = os.path.dirname(HERE)


def build():
    deps = [SRC] + [os.path.join(ROOT, "realtrace_b200", "csrc", h) for h in
                    ("rt_hd.h", "rt_scene.h", "rt_intersect.h", "rt_traverse.h", "rt_shade.h", "rt_bvh.h")]
    if os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    nvcc = "/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else "nvcc"
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-diag-suppress", "20011",
                           "-Xcompiler", "-fPIC,-mfma,-ffp-contract=fast", "-shared", "-o", LIB, SRC])
    return LIB


class Emulation:
    def __init__(self):
        self.lib = C.CDLL(build())
        vp = C.c_void_p
        self.lib.emul_render.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]
        self.lib.emul_trace_rays.argtypes = [vp, C.c_int, C.c_int, vp, C.c_uint32, vp, vp]
        self.lib.emul_bvh.argtypes = [vp, C.c_int, vp, vp, vp, vp]

    def render(self, scene, cam, depth, leaf=4, brute=0):
        s, keep = ob._scene_struct(scene)
        W, H = cam.width, cam.height
        rgb = np.zeros((H, W, 3), np.uint8)
        prim = np.zeros((H, W), np.int32)
        t = np.zeros((H, W), np.float32)
        counts = np.zeros(5, np.uint64)
        c = api.camera_struct(cam)
        rc = self.lib.emul_render(C.byref(s), C.byref(c), depth, leaf, brute, rgb.ctypes.data, prim.ctypes.data,
                                  t.ctypes.data, counts.ctypes.data)
        assert rc == 0, rc
        return rgb, prim, t, counts

    def trace_rays(self, scene, rays, leaf=4, brute=0):
        s, keep = ob._scene_struct(scene)
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        prim = np.zeros(len(rays), np.int32)
        t = np.zeros(len(rays), np.float32)
        rc = self.lib.emul_trace_rays(C.byref(s), leaf, brute, rays.ctypes.data, len(rays), prim.ctypes.data, t.ctypes.data)
        assert rc == 0, rc
        return prim, t

    def primary_cost(self, scene, cam, leaf=1, shadows=False):
        """Per-pixel (node visits, triangle tests) of the primary rays [and the steps of the hit's shadow rays]."""
        s, keep = ob._scene_struct(scene)
        W, H = cam.width, cam.height
        nodes = np.zeros((H, W), np.uint32)
        tris = np.zeros((H, W), np.uint32)
        shadow = np.zeros((H, W), np.uint32) if shadows else None
        c = api.camera_struct(cam)
        self.lib.emul_primary_cost.argtypes = [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 3
        rc = self.lib.emul_primary_cost(C.byref(s), C.byref(c), leaf, nodes.ctypes.data, tris.ctypes.data,
                                        shadow.ctypes.data if shadows else None)
        assert rc == 0, rc
        return (nodes, tris, shadow) if shadows else (nodes, tris)

    def hybrid_walk(self, scene, rays, switch_after):
        """Nearest hits of a walk that switches from the binary tree to its 4-wide view after `switch_after` steps
        (same stack).  -> (prim code, t, dependent steps) per ray."""
        s, keep = ob._scene_struct(scene)
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        prim = np.zeros(len(rays), np.int32)
        t = np.zeros(len(rays), np.float32)
        steps = np.zeros(len(rays), np.uint32)
        self.lib.emul_hybrid_walk.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint32] + [C.c_void_p] * 3
        rc = self.lib.emul_hybrid_walk(C.byref(s), switch_after, rays.ctypes.data, len(rays), prim.ctypes.data,
                                       t.ctypes.data, steps.ctypes.data)
        assert rc == 0, rc
        return prim, t, steps

    def bvh(self, scene, leaf=4):
        s, keep = ob._scene_struct(scene)
        n = len(scene.tri_v)
        nodes = np.zeros((max(n, 1), 16), np.float32)
        order = np.zeros(max(n, 1), np.uint32)
        keys = np.zeros(max(n, 1), np.uint64)
        nb = C.c_uint32()
        nn = self.lib.emul_bvh(C.byref(s), leaf, nodes.ctypes.data, order.ctypes.data, keys.ctypes.data, C.byref(nb))
        return nodes[:nn], order[:nb.value], keys[:nb.value]
