"""Named parity cases shared by the golden generator and the CPU/GPU tests."""
from oracle import binding as ob
from realtrace_b200 import scenes

SMALL = (160, 120)

GOLDEN_CASES = [
    "tetra_d10", "bob2000_d10", "analytic_close_d5", "analytic_notetra_d5", "analytic_stock_d1",
    "bobtex_d3", "blubmixed_d5", "synth_small_d1", "bob_full_bboxfixed_d10", "glass_ball_d0",
]


def build_case(name):
    """-> (scene, camera, max_depth, oracle mode)"""
    w, h = SMALL
    if name == "tetra_d10":
        return scenes.obj_scene("tetrahedron.obj"), scenes.stock_camera(w, h), 10, ob.MODE_AS_SHIPPED
    if name == "bob2000_d10":
        return scenes.obj_scene("bob_tri.obj", 2000), scenes.stock_camera(w, h), 10, ob.MODE_AS_SHIPPED
    if name == "analytic_close_d5":
        return scenes.analytic_scene(), scenes.close_camera(w, h), 5, ob.MODE_AS_SHIPPED
    if name == "analytic_notetra_d5":
        return scenes.analytic_scene(with_tetrahedron=False), scenes.close_camera(w, h), 5, ob.MODE_AS_SHIPPED
    if name == "analytic_stock_d1":
        return scenes.analytic_scene(), scenes.stock_camera(w, h), 1, ob.MODE_AS_SHIPPED
    if name == "bobtex_d3":
        return scenes.bob_textured(), scenes.stock_camera(w, h), 3, ob.MODE_AS_SHIPPED
    if name == "blubmixed_d5":
        return scenes.blub_mixed(), scenes.stock_camera(w, h), 5, ob.MODE_AS_SHIPPED
    if name == "synth_small_d1":
        return (scenes.synthetic_sphere_grid(grid=3, level=2), scenes.synthetic_camera(w, h, grid=3), 1,
                ob.MODE_AS_SHIPPED)
    if name == "glass_ball_d0":              # depth 0: the refracted child of a dielectric hit (level 0 * 2 = 0) is alive
        return glass_ball_scene(), scenes.close_camera(320, 240), 0, ob.MODE_TRUE_NEAREST
    if name == "bob_full_bboxfixed_d10":
        return scenes.obj_scene("bob_tri.obj"), scenes.stock_camera(w, h), 10, ob.MODE_BBOX_FIXED
    raise KeyError(name)


def glass_ball_scene():
    """A dielectric sphere (eta 1.5: refraction succeeds over the middle of the disc) in front of a diffuse sphere,
    over a diffuse floor."""
    import numpy as np
    from realtrace_b200.scene import Scene, make_materials
    mats = make_materials([dict(color=(1.0, 1.0, 1.0), ka=0.4, kd=0.9, ks=0.4, kr=0.1, kt=0.8, eta=1.5),
                           dict(color=(0.8, 0.1, 0.0), ka=0.2, kd=0.9, ks=0.4),
                           dict(color=(0.5, 0.5, 0.5), ka=0.1, kd=0.9, ks=0.2)])
    return Scene(sph=[(0.0, 0.0, 6.0, 4.0), (3.0, 1.0, -6.0, 4.0)], sph_material=[0, 1], sph_object_id=[0, 1],
                 pln=[(20, -5, 20, -20, -5, 20, -20, -5, -20, 20, -5, -20)], pln_material=[2], pln_object_id=[2],
                 materials=mats, lights=np.asarray([scenes.STOCK_LIGHT], np.float32), ambient=scenes.STOCK_AMBIENT,
                 background=scenes.STOCK_BACKGROUND, name="glass_ball").normalise()


# ---- edge cases (no golden fixtures: checked against the oracle on the spot) --------------------------------------
EDGE_CASES = ["empty_world", "no_lights", "one_triangle", "two_triangles", "degenerate_and_duplicate_triangles",
              "coincident_centroids", "tiny_frame_1x1", "tiny_frame_7x5", "tiny_frame_33x1"]


def build_edge_case(name):
    """-> (scene, camera, max_depth).  Inputs the domain offers at its borders: nothing to hit, nothing to light,
    hierarchies of one and two leaves, zero-area and repeated triangles (|A| < 1e-7 rejection, duplicate Morton
    keys), frames smaller than a tile and than a warp batch."""
    import numpy as np
    from realtrace_b200.scene import Scene, make_materials
    w, h = 96, 64
    cam = scenes.close_camera(w, h)
    base = dict(materials=make_materials([scenes.OBJ_MATERIAL, dict(color=(0.7, 0.7, 0.2), kr=0.4)]),
                lights=np.asarray([scenes.STOCK_LIGHT], np.float32), ambient=scenes.STOCK_AMBIENT,
                background=scenes.STOCK_BACKGROUND)
    tri = np.asarray([[-6, -4, 0, 6, -4, 0, 0, 6, 0]], np.float32)                     # faces the close camera
    if name == "empty_world":
        return Scene(name=name, **base).normalise(), cam, 3
    if name == "no_lights":
        s = scenes.obj_scene("bob_tri.obj", 2000, lights=())
        s.lights = np.zeros((0, 6), np.float32)
        return s.normalise(), scenes.stock_camera(w, h), 3
    if name == "one_triangle":
        return Scene(tri_v=tri, tri_material=np.zeros(1, np.uint32), name=name, **base).normalise(), cam, 3
    if name == "two_triangles":
        t2 = np.concatenate([tri, tri + np.asarray([3, 1, -5] * 3, np.float32)])
        return Scene(tri_v=t2, tri_material=np.asarray([0, 1], np.uint32), name=name, **base).normalise(), cam, 3
    if name == "degenerate_and_duplicate_triangles":
        s = scenes.obj_scene("tetrahedron.obj")
        v = np.asarray(s.tri_v, np.float32).reshape(-1, 9)
        zero_area = np.repeat(v[:4, :3], 3, axis=1).reshape(4, 9)                   # three equal vertices
        collinear = np.asarray([[0, 0, 0, 1, 1, 1, 2, 2, 2]], np.float32)
        allv = np.concatenate([v, v[:5], zero_area, collinear, v[5:8]])                # repeats: equal t, equal keys
        s.tri_v = allv
        s.tri_material = np.zeros(len(allv), np.uint32)
        return s.normalise(), scenes.stock_camera(w, h), 3
    if name == "coincident_centroids":
        # 40 triangles of growing size around ONE centroid: 40 identical Morton keys
        k = np.arange(1, 41, dtype=np.float32)[:, None]
        unit = np.asarray([[-1, -1, 0, 1, -1, 0, 0, 2, 0]], np.float32)
        allv = unit * (0.2 * k) + np.asarray([0, 0, 1] * 3, np.float32) * (0.1 * k)   # (z offset keeps t distinct)
        allv[:, 2::3] -= allv[:, 2::3].mean(axis=1, keepdims=True) - 0.0                # same centroid in z too
        allv[:, 2::3] += 0.05 * k                                                       # ... up to a small stagger
        return Scene(tri_v=allv, tri_material=(np.arange(40) % 2).astype(np.uint32), name=name, **base).normalise(), cam, 3
    if name.startswith("tiny_frame_"):
        fw, fh = (int(x) for x in name[len("tiny_frame_"):].split("x"))
        return scenes.obj_scene("tetrahedron.obj"), scenes.stock_camera(fw, fh), 3
    raise KeyError(name)


def random_scene(seed, n_tri=300, w=96, h=64):
    """Seeded random world for fuzzing: a triangle soup around the origin, a few spheres / quads / cylinders, mixed
    diffuse / mirror / dielectric / vertex-coloured materials, one to three lights, a camera looking at the cloud."""
    import numpy as np
    from realtrace_b200.scene import Camera, Scene, make_materials
    rng = np.random.default_rng(seed)
    c = rng.uniform(-12, 12, (n_tri, 1, 3))
    tri = (c + rng.uniform(-2.5, 2.5, (n_tri, 3, 3))).reshape(n_tri, 9).astype(np.float32)
    mats = make_materials([
        dict(color=tuple(rng.uniform(0.1, 0.9, 3)), ka=0.2, kd=0.9, ks=0.4),
        dict(color=tuple(rng.uniform(0.1, 0.9, 3)), ka=0.1, kd=0.7, ks=0.3, kr=float(rng.uniform(0.2, 0.8))),
        dict(color=(1.0, 1.0, 1.0), ka=0.3, kd=0.9, ks=0.4, kr=0.1, kt=0.8, eta=float(rng.uniform(1.2, 2.0))),
        dict(color=(0.5, 0.5, 0.5), ka=0.2, kd=1.0, ks=0.4, barycentric=True),
    ])
    tri_mat = rng.integers(0, 4, n_tri).astype(np.uint32)
    n_s, n_p, n_c = (int(x) for x in rng.integers(0, 4, 3))
    sph = np.concatenate([rng.uniform(-10, 10, (n_s, 3)), rng.uniform(0.8, 3.0, (n_s, 1))], axis=1).astype(np.float32)
    y0 = float(rng.uniform(-14, -9))
    pln = np.asarray([[-25, y0 - k, -25, 25, y0 - k, -25, 25, y0 - k, 25, -25, y0 - k, 25] for k in range(n_p)], np.float32).reshape(n_p, 12)
    cyl = np.concatenate([rng.uniform(-10, 10, (n_c, 3)), rng.uniform(0.4, 1.2, (n_c, 1)),
                          rng.normal(size=(n_c, 3))], axis=1).astype(np.float32)      # axis not normalised (cylinder.h:17-21)
    nl = int(rng.integers(1, 4))
    lights = np.concatenate([rng.uniform(-30, 30, (nl, 3)) + np.asarray([0, 35, 0]), rng.uniform(0.3, 1.0, (nl, 3))], axis=1).astype(np.float32)
    scene = Scene(tri_v=tri, tri_material=tri_mat, tri_rgb=rng.uniform(0, 1, (n_tri, 9)).astype(np.float32),
                  sph=sph, sph_material=rng.integers(0, 3, n_s).astype(np.uint32),
                  pln=pln, pln_material=rng.integers(0, 2, n_p).astype(np.uint32),
                  cyl=cyl, cyl_material=rng.integers(0, 3, n_c).astype(np.uint32),
                  materials=mats, lights=lights, ambient=(1.0, 1.0, 1.0), background=(0.1, 0.3, 0.6),
                  name=f"random{seed}").normalise()
    eye = rng.normal(size=3)
    eye = eye / np.linalg.norm(eye) * float(rng.uniform(28, 45))
    cam = Camera(pos=tuple(float(x) for x in eye), target=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), fovy=45.0, width=w, height=h)
    return scene, cam, int(rng.integers(1, 6))
