"""ctypes access to the CPU checkers (oracle_abi.h).  TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs — never by the product package realtrace_b200.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libserial_ref.so")
PORT_LIB = os.path.join(HERE, "libserial_port.so")

MODE_AS_SHIPPED, MODE_BBOX_FIXED, MODE_TRUE_NEAREST = 0, 1, 2


class _Material(C.Structure):
    _fields_ = [("color", C.c_float * 3), ("ka", C.c_float), ("kd", C.c_float), ("ks", C.c_float),
                ("kr", C.c_float), ("kt", C.c_float), ("eta", C.c_float), ("barycentric", C.c_uint32)]


_FP = C.POINTER(C.c_float)
_UP = C.POINTER(C.c_uint32)


class _Scene(C.Structure):
    _fields_ = [("tri_v", _FP), ("tri_material", _UP), ("tri_rgb", _FP), ("tri_object_id", _UP), ("n_tri", C.c_uint32),
                ("sph", _FP), ("sph_material", _UP), ("sph_object_id", _UP), ("n_sph", C.c_uint32),
                ("pln", _FP), ("pln_material", _UP), ("pln_object_id", _UP), ("n_pln", C.c_uint32),
                ("cyl", _FP), ("cyl_material", _UP), ("cyl_object_id", _UP), ("n_cyl", C.c_uint32),
                ("materials", C.POINTER(_Material)), ("n_materials", C.c_uint32),
                ("lights", _FP), ("n_lights", C.c_uint32),
                ("ambient", C.c_float * 3), ("background", C.c_float * 3)]


class _Camera(C.Structure):
    _fields_ = [("pos", C.c_double * 3), ("target", C.c_double * 3), ("up", C.c_double * 3),
                ("fovy", C.c_float), ("width", C.c_int32), ("height", C.c_int32)]


class _Result(C.Structure):
    _fields_ = [("rays_total", C.c_uint64), ("rays_primary", C.c_uint64), ("rays_shadow", C.c_uint64),
                ("rays_secondary", C.c_uint64), ("build_seconds", C.c_double), ("render_seconds", C.c_double),
                ("columns_rendered", C.c_uint32), ("threads_used", C.c_uint32)]


def _fp(a):
    return a.ctypes.data_as(_FP) if a is not None and a.size else None


def _up(a):
    return a.ctypes.data_as(_UP) if a is not None and a.size else None


def _scene_struct(scene):
    """Returns (struct, keepalive)."""
    s = _Scene()
    keep = [scene]
    s.tri_v, s.tri_material, s.tri_rgb = _fp(scene.tri_v), _up(scene.tri_material), _fp(scene.tri_rgb)
    s.tri_object_id, s.n_tri = _up(scene.tri_object_id), len(scene.tri_v)
    s.sph, s.sph_material, s.sph_object_id, s.n_sph = _fp(scene.sph), _up(scene.sph_material), _up(scene.sph_object_id), len(scene.sph)
    s.pln, s.pln_material, s.pln_object_id, s.n_pln = _fp(scene.pln), _up(scene.pln_material), _up(scene.pln_object_id), len(scene.pln)
    s.cyl, s.cyl_material, s.cyl_object_id, s.n_cyl = _fp(scene.cyl), _up(scene.cyl_material), _up(scene.cyl_object_id), len(scene.cyl)
    mats = np.ascontiguousarray(scene.materials)
    keep.append(mats)
    s.materials = mats.ctypes.data_as(C.POINTER(_Material))
    s.n_materials = len(mats)
    s.lights, s.n_lights = _fp(scene.lights), len(scene.lights)
    s.ambient = (C.c_float * 3)(*scene.ambient)
    s.background = (C.c_float * 3)(*scene.background)
    return s, keep


class Oracle:
    def __init__(self, path):
        self.path = path
        self.lib = C.CDLL(path)
        self.lib.oracle_name.restype = C.c_char_p
        self.lib.oracle_render.restype = C.c_int
        self.lib.oracle_render.argtypes = [C.POINTER(_Scene), C.POINTER(_Camera), C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_Result)]
        self.lib.oracle_trace_rays.restype = C.c_int
        self.lib.oracle_trace_rays.argtypes = [C.POINTER(_Scene), C.c_int, _FP, C.c_uint32, C.c_void_p, C.c_void_p]
        self.lib.oracle_shade_rays.restype = C.c_int
        self.lib.oracle_shade_rays.argtypes = [C.POINTER(_Scene), C.c_int, C.c_int, _FP, C.c_uint32, C.c_void_p]
        self.name = self.lib.oracle_name().decode()
        self.kind = "reference" if self.name.startswith("reference") else "port"

    def render(self, scene, cam, max_depth, mode=MODE_AS_SHIPPED, col_begin=0, col_step=1, nthreads=1, aux=True):
        s, keep = _scene_struct(scene)
        c = _Camera()
        c.pos = (C.c_double * 3)(*cam.pos)
        c.target = (C.c_double * 3)(*cam.target)
        c.up = (C.c_double * 3)(*cam.up)
        c.fovy, c.width, c.height = cam.fovy, cam.width, cam.height
        W, H = cam.width, cam.height
        rgb = np.zeros((H, W, 3), np.uint8)
        prim = np.full((H, W), -1, np.int32) if aux else None
        t = np.full((H, W), np.finfo(np.float32).max, np.float32) if aux else None
        res = _Result()
        rc = self.lib.oracle_render(C.byref(s), C.byref(c), max_depth, mode, col_begin, col_step, nthreads,
                                    rgb.ctypes.data, prim.ctypes.data if aux else None,
                                    t.ctypes.data if aux else None, C.byref(res))
        if rc != 0:
            raise RuntimeError(f"oracle_render failed: {rc}")
        info = {k: getattr(res, k) for k, _ in _Result._fields_}
        return rgb, prim, t, info

    def trace_rays(self, scene, rays, mode=MODE_TRUE_NEAREST):
        s, keep = _scene_struct(scene)
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        prim = np.zeros(len(rays), np.int32)
        t = np.zeros(len(rays), np.float32)
        rc = self.lib.oracle_trace_rays(C.byref(s), mode, _fp(rays), len(rays), prim.ctypes.data, t.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"oracle_trace_rays failed: {rc}")
        return prim, t

    def shade_rays(self, scene, rays, max_depth, mode=MODE_TRUE_NEAREST):
        s, keep = _scene_struct(scene)
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        out = np.zeros((len(rays), 3), np.float64)
        rc = self.lib.oracle_shade_rays(C.byref(s), max_depth, mode, _fp(rays), len(rays), out.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"oracle_shade_rays failed: {rc}")
        return out


def load_reference():
    """The reference's own Serial core, if it was built here (oracle/build_ref.py); else None."""
    return Oracle(REF_LIB) if os.path.exists(REF_LIB) else None


def load_port():
    """The from-scratch restatement (oracle/serial_port.cpp); raises if it has not been built."""
    if not os.path.exists(PORT_LIB):
        raise FileNotFoundError(f"{PORT_LIB} missing — run `python -c 'import __graft_entry__ as g; g.build()'`")
    return Oracle(PORT_LIB)


def best_available():
    return load_reference() or load_port()


def fnv1a64(buf) -> str:
    """FNV-1a-64 over a byte buffer (the hash SURVEY Appendix B pins frames with)."""
    h = 1469598103934665603
    data = np.ascontiguousarray(buf).reshape(-1).view(np.uint8)
    # vectorising FNV is not possible (sequential); frames are <= a few MB in tests
    for b in data.tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"
