"""The named workloads of BASELINE.json `configs`, as flat Scene/Camera records.

Every scene is a deterministic file or closed form; the only RNG is
the synthetic scene's LCG (seed 12345).  Stock camera, light, ambient,
background and material come from /root/reference/Serial/lumina.cpp:302-310,
:360-362 and :163-172; the analytic objects are the scene literal the reference
keeps commented out at :312-356.
"""
from __future__ import annotations

import math
import os

import numpy as np

from .objio import ASSETS, load_texture, triangles_from_obj, vertex_colours
from .scene import Camera, Scene, make_materials, orbit_camera

STOCK_LIGHT = (0.0, 30.0, 30.0, 0.5, 1.0, 1.0)        # lumina.cpp:360
SECOND_LIGHT = (0.0, 10.0, 0.0, 1.0, 1.0, 1.0)        # lumina.cpp:361 (commented out there)
STOCK_AMBIENT = (1.0, 1.0, 1.0)                       # lumina.cpp:309
STOCK_BACKGROUND = (0.1, 0.3, 0.6)                    # lumina.cpp:310
OBJ_MATERIAL = dict(color=(0.8, 0.1, 0.0), ka=0.2, kd=0.9, ks=0.4, kr=0.4, kt=0.0, eta=3.0)  # :163-172


def stock_camera(width=640, height=480):
    return Camera(pos=(60.0, 60.0, 0.0), target=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), fovy=45.0,
                  width=width, height=height)                                   # lumina.cpp:302-306


def obj_scene(name, max_faces=None, material=None, lights=(STOCK_LIGHT,)):
    """An OBJ file with the loader's default material on every face (lumina.cpp:256-258)."""
    tri, _, _, _ = triangles_from_obj(os.path.join(ASSETS, name), max_faces)
    s = Scene(tri_v=tri, tri_material=np.zeros(len(tri), np.uint32),
              materials=make_materials([material or OBJ_MATERIAL]),
              lights=np.asarray(lights, np.float32), ambient=STOCK_AMBIENT, background=STOCK_BACKGROUND,
              name=name)
    return s.normalise()


# ---- config 1: analytic objects + tetrahedron.obj ------------------------------------------
def analytic_scene(with_mirror_sphere=True, with_tetrahedron=True):
    mats = make_materials([
        dict(color=(0.1, 0.7, 0.0), ka=0.2, kd=0.9, ks=0.4, kr=1.0, kt=0.0, eta=1.0),   # m   :312-320
        dict(color=(0.8, 0.1, 0.0), ka=0.2, kd=0.9, ks=0.4, kr=0.0, kt=0.0, eta=1.0),   # m1  :323-331
        dict(color=(1.0, 1.0, 1.0), ka=0.4, kd=0.9, ks=0.4, kr=0.1, kt=0.8, eta=2.0),   # m2  :333-341
        dict(color=(0.5, 0.5, 0.5), ka=0.1, kd=0.9, ks=0.2, kr=0.5, kt=0.0, eta=1.0),   # floorMat :344-351
        OBJ_MATERIAL,
    ])
    sph, sph_m = [], []
    if with_mirror_sphere:
        sph.append((0.0, 0.0, 0.0, 3.0)); sph_m.append(0)                 # sphere   :322
    sph.append((4.0, 0.0, 4.0, 3.0)); sph_m.append(1)                     # sphere2  :343
    pln = [(10, -3, 10, -10, -3, 10, -10, -3, -10, 10, -3, -10)]          # plane    :352
    cyl = [(-7, 0, -3, 1, 0, 0, 1)]                                       # cylinder :342
    # objectList order of the literal: sphere, sphere2, plane, cylinder (:353-356), then the OBJ faces (:366)
    n_s = len(sph)
    sph_id = list(range(n_s))
    pln_id = [n_s]
    cyl_id = [n_s + 1]
    if with_tetrahedron:
        tri, _, _, _ = triangles_from_obj(os.path.join(ASSETS, "tetrahedron.obj"))
    else:
        tri = np.zeros((0, 9), np.float32)
    tri_id = np.arange(len(tri), dtype=np.uint32) + (n_s + 2)
    s = Scene(tri_v=tri, tri_material=np.full(len(tri), 4, np.uint32), tri_object_id=tri_id,
              sph=sph, sph_material=sph_m, sph_object_id=sph_id,
              pln=pln, pln_material=[3], pln_object_id=pln_id,
              cyl=cyl, cyl_material=[2], cyl_object_id=cyl_id,
              materials=mats, lights=np.asarray([STOCK_LIGHT], np.float32),
              ambient=STOCK_AMBIENT, background=STOCK_BACKGROUND, name="analytic")
    return s.normalise()


def close_camera(width=640, height=480):
    """Second camera for config 1: the stock camera sees the analytic objects only as a few pixels."""
    return Camera(pos=(0.0, 10.0, 30.0), target=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), fovy=45.0,
                  width=width, height=height)


# ---- config 2 / 5: textured bob ------------------------------------------------------------
def bob_textured(texture_mode="normalised", max_faces=None, lights=(STOCK_LIGHT, SECOND_LIGHT)):
    """bob_tri.obj with one BarycentricMaterial per face (lumina.cpp:248-249); the material
    coefficients are the loader's (:163-172) so that depth 3 has mirror bounces to follow."""
    tri, _, tex, faces = triangles_from_obj(os.path.join(ASSETS, "bob_tri.obj"), max_faces)
    rgb = vertex_colours(faces, tex, load_texture(os.path.join(ASSETS, "bob_diffuse.png")), texture_mode)
    m = dict(OBJ_MATERIAL)
    m["barycentric"] = True
    s = Scene(tri_v=tri, tri_material=np.zeros(len(tri), np.uint32), tri_rgb=rgb,
              materials=make_materials([m]), lights=np.asarray(lights, np.float32),
              ambient=STOCK_AMBIENT, background=STOCK_BACKGROUND, name=f"bob_textured[{texture_mode}]")
    return s.normalise()


# ---- config 3: blub with a deterministic material rule ---------------------------------------
def blub_mixed(max_faces=None):
    tri, _, _, _ = triangles_from_obj(os.path.join(ASSETS, "blub_triangulated.obj"), max_faces)
    mats = make_materials([
        dict(color=(1.0, 1.0, 1.0), ka=0.2, kd=0.9, ks=0.4, kr=0.1, kt=0.8, eta=1.5),   # id % 3 == 0 dielectric
        dict(color=(0.8, 0.1, 0.0), ka=0.2, kd=0.9, ks=0.4, kr=0.4, kt=0.0, eta=3.0),   # == 1 mirror
        dict(color=(0.1, 0.7, 0.0), ka=0.2, kd=0.9, ks=0.4, kr=0.0, kt=0.0, eta=1.0),   # == 2 diffuse
    ])
    s = Scene(tri_v=tri, tri_material=(np.arange(len(tri)) % 3).astype(np.uint32), materials=mats,
              lights=np.asarray([STOCK_LIGHT], np.float32), ambient=STOCK_AMBIENT,
              background=STOCK_BACKGROUND, name="blub_mixed")
    return s.normalise()


# ---- config 4: synthetic tessellated-sphere grid ---------------------------------------------
def icosphere(level):
    """Unit icosphere: 20*4**level faces.  Returns (V,3) f64 vertices, (F,3) int faces."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = np.array([(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
                  (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)], np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2),
                  (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5),
                  (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)], np.int64)
    for _ in range(level):
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
        e.sort(axis=1)
        uniq, inv = np.unique(e, axis=0, return_inverse=True)
        inv = inv.reshape(-1)
        mid = v[uniq[:, 0]] + v[uniq[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = len(v)
        v = np.concatenate([v, mid], axis=0)
        n = len(f)
        ab, bc, ca = base + inv[:n], base + inv[n:2 * n], base + inv[2 * n:]
        a, b, c = f[:, 0], f[:, 1], f[:, 2]
        f = np.concatenate([np.stack([a, ab, ca], 1), np.stack([b, bc, ab], 1),
                            np.stack([c, ca, bc], 1), np.stack([ab, bc, ca], 1)], axis=0)
    return v, f


SYNTH_ORIGIN = (50.0, 10.0, 50.0)   # grid centre; the whole scene sits in the positive octant


def synthetic_sphere_grid(grid=14, level=4, pitch=2.5, seed=12345):
    """grid x grid icospheres of radius U[0.8,1.0] on the plane y = origin.y plus a 2-triangle floor.

    14 x 14 x 5120 = 1 003 520 triangles + 2.  Radius jitter: 31-bit LCG
    x <- (1103515245 x + 12345) mod 2^31, u = x / 2^31, seed 12345."""
    v, f = icosphere(level)
    unit = v[f].reshape(-1, 9)                      # (F, 9)
    ox, oy, oz = SYNTH_ORIGIN
    x = seed
    tris, mats = [], []
    half = (grid - 1) / 2.0
    for gz in range(grid):
        for gx in range(grid):
            x = (1103515245 * x + 12345) % (1 << 31)
            r = 0.8 + 0.2 * (x / float(1 << 31))
            c = np.array([ox + (gx - half) * pitch, oy, oz + (gz - half) * pitch] * 3, np.float64)
            tris.append(unit * r + c)
            mats.append(np.full(len(unit), (gx + gz * grid) % 4, np.uint32))
    ext = grid * pitch / 2.0 + 2.0
    fy = oy - 1.2
    floor = np.array([[ox - ext, fy, oz - ext, ox - ext, fy, oz + ext, ox + ext, fy, oz + ext],
                      [ox - ext, fy, oz - ext, ox + ext, fy, oz + ext, ox + ext, fy, oz - ext]], np.float64)
    tris.append(floor)
    mats.append(np.full(2, 4, np.uint32))
    tri = np.concatenate(tris, axis=0).astype(np.float32)
    mat = np.concatenate(mats, axis=0)
    materials = make_materials([
        dict(color=(0.8, 0.1, 0.0), ka=0.2, kd=0.9, ks=0.4),
        dict(color=(0.1, 0.7, 0.0), ka=0.2, kd=0.9, ks=0.4),
        dict(color=(0.1, 0.2, 0.8), ka=0.2, kd=0.9, ks=0.4),
        dict(color=(0.8, 0.7, 0.1), ka=0.2, kd=0.9, ks=0.4),
        dict(color=(0.5, 0.5, 0.5), ka=0.1, kd=0.9, ks=0.2),
    ])
    light = (ox, oy + 40.0, oz + 20.0, 1.0, 1.0, 1.0)
    s = Scene(tri_v=tri, tri_material=mat, materials=materials, lights=np.asarray([light], np.float32),
              ambient=STOCK_AMBIENT, background=STOCK_BACKGROUND,
              name=f"synthetic_sphere_grid[{grid}x{grid},level{level}]")
    return s.normalise()


def synthetic_camera(width=3840, height=2160, grid=14, pitch=2.5):
    """Looks down at 35 degrees on the grid centre from a distance that frames the whole grid."""
    ox, oy, oz = SYNTH_ORIGIN
    dist = 1.22 * grid * pitch
    a = math.radians(35.0)
    return Camera(pos=(ox, oy + dist * math.sin(a), oz + dist * math.cos(a)), target=(ox, oy, oz),
                  up=(0.0, 1.0, 0.0), fovy=45.0, width=width, height=height)


# ---- registry used by bench.py and the tests -------------------------------------------------
def workload(name, width=None, height=None):
    """Returns (scene, camera, max_depth, description)."""
    if name == "analytic":          # config 1
        w, h = width or 640, height or 480
        return analytic_scene(), stock_camera(w, h), 1, "config1: sphere/plane/cylinder + tetrahedron.obj, depth 1"
    if name == "analytic_close":
        w, h = width or 640, height or 480
        return analytic_scene(), close_camera(w, h), 1, "config1 (close camera)"
    if name == "bob1080":           # config 2
        w, h = width or 1920, height or 1080
        return bob_textured(), stock_camera(w, h), 3, "config2: bob_tri.obj + bob_diffuse.png, 2 lights, depth 3"
    if name == "blub4k":            # config 3
        w, h = width or 3840, height or 2160
        return blub_mixed(), stock_camera(w, h), 5, "config3: blub dielectric/mirror/diffuse by face id, depth 5"
    if name == "synth1m":           # config 4
        w, h = width or 3840, height or 2160
        return (synthetic_sphere_grid(), synthetic_camera(w, h), 1,
                "config4: 14x14 level-4 icospheres (1 003 522 triangles), primary+shadow")
    if name == "orbit":             # config 5, frame 0; bench iterates orbit_camera(k)
        w, h = width or 1920, height or 1080
        return bob_textured(), orbit_camera(0, width=w, height=h), 3, "config5: 120-frame orbit over bob, refit per frame"
    raise KeyError(name)
