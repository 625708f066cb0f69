"""Flat scene and camera records used by the harness (tests, bench, smoke).

The product boundary is the C ABI in include/realtrace_b200.h; these classes
only hold the float32 arrays that are handed to it (and, in tests, to the CPU
oracle), so both sides see bit-identical inputs.

Reference types being flattened (file:line in /root/reference/Serial):
  Triangle triangle.h:14-37, Sphere sphere.h:10-25, Plane plane.h:10-30,
  Cylinder cylinder.h:10-25, Material material.h:11-32,
  BarycentricMaterial material.h:35-50, PointLightSource pointlightsource.h:6-14,
  World world.h:12-45, Camera camera.h:7-34 / camera.cpp:4-25.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# Layout of rt_material / oracle_material: float color[3]; float ka,kd,ks,kr,kt,eta; uint32 flags
MATERIAL_DTYPE = np.dtype([("color", "<f4", (3,)), ("ka", "<f4"), ("kd", "<f4"), ("ks", "<f4"),
                           ("kr", "<f4"), ("kt", "<f4"), ("eta", "<f4"), ("flags", "<u4")])
MATERIAL_BARYCENTRIC = 1


def make_materials(rows):
    """rows: iterable of dict(color=(r,g,b), ka=, kd=, ks=, kr=, kt=, eta=, barycentric=bool).

    Defaults are Material's constructor defaults (material.h:27-29)."""
    out = np.zeros(len(rows), dtype=MATERIAL_DTYPE)
    for i, r in enumerate(rows):
        out[i]["color"] = r.get("color", (0.0, 0.0, 0.0))
        out[i]["ka"] = r.get("ka", 0.2)
        out[i]["kd"] = r.get("kd", 1.0)
        out[i]["ks"] = r.get("ks", 0.4)
        out[i]["kr"] = r.get("kr", 0.0)
        out[i]["kt"] = r.get("kt", 0.0)
        out[i]["eta"] = r.get("eta", 128.0)
        out[i]["flags"] = MATERIAL_BARYCENTRIC if r.get("barycentric", False) else 0
    return out


def _f32(a, cols):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1, cols))
    return a


def _u32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint32).reshape(-1))


@dataclass
class Scene:
    tri_v: np.ndarray = field(default_factory=lambda: np.zeros((0, 9), np.float32))
    tri_material: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))
    tri_rgb: np.ndarray | None = None
    tri_object_id: np.ndarray | None = None
    sph: np.ndarray = field(default_factory=lambda: np.zeros((0, 4), np.float32))
    sph_material: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))
    sph_object_id: np.ndarray | None = None
    pln: np.ndarray = field(default_factory=lambda: np.zeros((0, 12), np.float32))
    pln_material: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))
    pln_object_id: np.ndarray | None = None
    cyl: np.ndarray = field(default_factory=lambda: np.zeros((0, 7), np.float32))
    cyl_material: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))
    cyl_object_id: np.ndarray | None = None
    materials: np.ndarray = field(default_factory=lambda: make_materials([{}]))
    lights: np.ndarray = field(default_factory=lambda: np.zeros((0, 6), np.float32))
    ambient: tuple = (0.0, 0.0, 0.0)
    background: tuple = (0.0, 0.0, 0.0)
    name: str = "scene"

    def normalise(self) -> "Scene":
        self.tri_v = _f32(self.tri_v, 9)
        self.tri_material = _u32(self.tri_material)
        if self.tri_rgb is not None:
            self.tri_rgb = _f32(self.tri_rgb, 9)
        if self.tri_object_id is not None:
            self.tri_object_id = _u32(self.tri_object_id)
        self.sph = _f32(self.sph, 4)
        self.sph_material = _u32(self.sph_material)
        self.pln = _f32(self.pln, 12)
        self.pln_material = _u32(self.pln_material)
        self.cyl = _f32(self.cyl, 7)
        self.cyl_material = _u32(self.cyl_material)
        for k in ("sph_object_id", "pln_object_id", "cyl_object_id"):
            if getattr(self, k) is not None:
                setattr(self, k, _u32(getattr(self, k)))
        self.lights = _f32(self.lights, 6)
        self.materials = np.ascontiguousarray(self.materials, dtype=MATERIAL_DTYPE)
        assert len(self.tri_material) == len(self.tri_v)
        assert len(self.sph_material) == len(self.sph)
        assert len(self.pln_material) == len(self.pln)
        assert len(self.cyl_material) == len(self.cyl)
        return self

    @property
    def n_objects(self) -> int:
        return len(self.tri_v) + len(self.sph) + len(self.pln) + len(self.cyl)


@dataclass
class Camera:
    """Pin-hole camera with the reference's constructor arguments (camera.h:25)."""
    pos: tuple = (60.0, 60.0, 0.0)
    target: tuple = (0.0, 0.0, 0.0)
    up: tuple = (0.0, 1.0, 0.0)
    fovy: float = 45.0
    width: int = 640
    height: int = 480

    def basis(self):
        """Camera basis exactly as camera.cpp:4-25 computes it (FP64, then the
        float members focalDistance / aspect)."""
        pos = np.asarray(self.pos, np.float64)
        tgt = np.asarray(self.target, np.float64)
        up = np.asarray(self.up, np.float64)
        up = up / np.sqrt((up * up).sum())
        w = -(tgt - pos)
        w = w / np.sqrt((w * w).sum())
        u = np.cross(up, w)
        u = u / np.sqrt((u * u).sum())
        v = np.cross(w, u)
        v = v / np.sqrt((v * v).sum())
        aspect = np.float32(np.float32(self.width) / np.float32(self.height))
        fovy32 = np.float32(self.fovy)
        focal = np.float32(1.0 / (2.0 * math.tan(float(fovy32) * math.pi / (180.0 * 2.0))))
        return u, v, w, focal, aspect


def orbit_camera(k: int, n: int = 120, radius: float = 84.85281374238570, pitch: float = 0.3,
                 width: int = 1920, height: int = 1080, fovy: float = 45.0) -> Camera:
    """Config-5 orbit: eye position from InteractiveCamera::buildRenderCamera
    (/root/reference/Parellel/interactive_camera.cu:64-72), which computes the
    three direction components in float."""
    yaw = np.float32(2.0 * math.pi * k / n)
    p = np.float32(pitch)
    x = np.float32(np.sin(yaw) * np.cos(p))
    y = np.float32(np.sin(p))
    z = np.float32(np.cos(yaw) * np.cos(p))
    r = np.float32(radius)
    eye = (float(x * r), float(y * r), float(z * r))
    return Camera(pos=eye, target=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), fovy=fovy, width=width, height=height)
