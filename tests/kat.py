"""Hand-computable known-answer rays for each primitive kind.

Used by tests/golden/make_golden.py (answers taken from the reference build)
and by the CPU and GPU parity tests.
"""
import numpy as np

from realtrace_b200.scene import Scene, make_materials


def kat_scene():
    """Sphere, quad, cylinder and one triangle, far enough apart that every ray below
    concerns one primitive.  Object ids: 0 sphere, 1 plane, 2 cylinder, 3 triangle."""
    mats = make_materials([dict(color=(0.8, 0.1, 0.0), ka=0.2, kd=0.9, ks=0.4)])
    s = Scene(
        tri_v=[(100, 0, 0, 104, 0, 0, 100, 4, 0)], tri_material=[0], tri_object_id=[3],
        sph=[(0, 0, 0, 3)], sph_material=[0], sph_object_id=[0],
        pln=[(10 + 200, -3, 10, -10 + 200, -3, 10, -10 + 200, -3, -10, 10 + 200, -3, -10)], pln_material=[0],
        pln_object_id=[1],
        cyl=[(-100, 0, 0, 1, 0, 0, 1)], cyl_material=[0], cyl_object_id=[2],
        materials=mats, lights=[(0, 30, 30, 0.5, 1, 1)], ambient=(1, 1, 1), background=(0.1, 0.3, 0.6), name="kat")
    return s.normalise()


def kat_rays():
    """(name, origin, direction) triples."""
    R = []
    # --- sphere (sphere.cpp:5-39), centre 0, r 3
    R.append(("sph_front", (0, 0, 10), (0, 0, -1)))            # t = 7
    R.append(("sph_tangent_exact", (3, 0, 10), (0, 0, -1)))    # disc == 0 branch, t = 10
    R.append(("sph_inside", (0, 0, 0), (1, 0, 0)))             # far root 3
    R.append(("sph_near_root_rejected", (0, 0, 3.00005), (0, 0, -1)))   # near root 5e-5 < 1e-4 -> far root
    R.append(("sph_behind", (0, 0, 10), (0, 0, 1)))            # both roots negative
    R.append(("sph_miss", (0, 5, 10), (0, 0, -1)))
    R.append(("sph_oblique", (5, 4, 9), (-0.5, -0.4, -1)))
    # --- plane quad (plane.cpp:12-27), y = -3, x in 190..210, z in -10..10
    R.append(("pln_tri1", (195, 5, 5), (0, -1, 0)))            # inside (p1,p2,p3)
    R.append(("pln_tri2", (205, 5, -5), (0, -1, 0)))           # inside (p1,p3,p4)
    R.append(("pln_diagonal", (200, 5, 0), (0, -1, 0)))        # on the shared diagonal: strict test misses both
    R.append(("pln_parallel", (200, -3, 0), (1, 0, 0)))        # |A| < 1e-7
    R.append(("pln_from_below", (195, -9, 5), (0, 1, 0)))      # back face still hits
    R.append(("pln_outside", (215, 5, 0), (0, -1, 0)))
    R.append(("pln_oblique", (190, 7, -20), (0.4, -0.5, 1)))
    # --- cylinder (cylinder.cpp:4-32), infinite about z through (-100,0,0), r 1
    R.append(("cyl_perp", (-100, 10, 0), (0, -1, 0)))          # t = 9
    R.append(("cyl_inside", (-100, 0, 5), (1, 0, 0)))          # t1 < 0 -> t2 = 1
    R.append(("cyl_parallel_axis", (-100, 0.5, 0), (0, 0, 1))) # A == 0 -> NaN roots -> miss
    R.append(("cyl_t1_tiny_never_t2", (-100, 1.00005, 0), (0, -1, 0)))  # 0 < t1 <= 1e-4: t2 never tried
    R.append(("cyl_miss", (-100, 10, 0), (1, 0, 0)))
    R.append(("cyl_oblique", (-95, 3, -40), (-0.5, -0.3, 1)))
    # --- triangle (triangle.cpp:10-24), (100,0,0) (104,0,0) (100,4,0)
    R.append(("tri_centre", (101, 1, 5), (0, 0, -1)))          # t = 5
    R.append(("tri_back", (101, 1, -5), (0, 0, 1)))            # no culling
    R.append(("tri_edge_beta0", (100, 2, 5), (0, 0, -1)))      # on edge a-c: strict > fails
    R.append(("tri_vertex", (100, 0, 5), (0, 0, -1)))
    R.append(("tri_hypotenuse", (102, 2, 5), (0, 0, -1)))      # beta + gamma == 1: strict < fails
    R.append(("tri_parallel", (99, 1, 0), (1, 0, 0)))          # |A| < 1e-7
    R.append(("tri_t_below_eps", (101, 1, 0.00005), (0, 0, -1)))   # t = 5e-5 rejected
    R.append(("tri_t_above_eps", (101, 1, 0.0002), (0, 0, -1)))    # t = 2e-4 accepted
    R.append(("tri_oblique", (103, 3, 7), (-0.3, -0.35, -1)))
    R.append(("all_miss", (0, 50, 0), (0, 1, 0)))
    names = [r[0] for r in R]
    rays = np.asarray([tuple(r[1]) + tuple(r[2]) for r in R], np.float32)
    return names, rays
