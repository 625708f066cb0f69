"""Pins the CPU checkers: the port must be bit-identical to the fixtures that were generated
from the reference's own Serial sources (tests/golden/make_golden.py), and — where the
reference build exists on this machine — to the reference itself on fresh inputs."""
import json
import os

import numpy as np
import pytest

import kat
from cases import GOLDEN_CASES, build_case
from conftest import GOLDEN
from oracle import binding as ob
from realtrace_b200 import scenes


def _golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


PINS = json.load(open(os.path.join(GOLDEN, "pins.json")))


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_port_matches_golden_frames(port_oracle, name):
    scene, cam, depth, mode = build_case(name)
    rgb, prim, t, info = port_oracle.render(scene, cam, depth, mode)
    g = _golden(name)
    assert np.array_equal(rgb, g["rgb"])
    assert np.array_equal(prim, g["prim_id"])
    assert np.array_equal(t, g["t"])
    assert info["rays_total"] == PINS["cases"][name]["rays_total"]
    assert ob.fnv1a64(rgb) == PINS["cases"][name]["fnv1a64"]


def test_port_matches_kat_rays(port_oracle):
    g = _golden("kat_rays")
    names, rays = kat.kat_rays()
    assert list(g["names"]) == names
    assert np.array_equal(g["rays"], rays)
    s = kat.kat_scene()
    for mode_name, mode in (("as_shipped", ob.MODE_AS_SHIPPED), ("true_nearest", ob.MODE_TRUE_NEAREST)):
        prim, t = port_oracle.trace_rays(s, rays, mode)
        assert np.array_equal(prim, g["prim_" + mode_name]), mode_name
        assert np.array_equal(t, g["t_" + mode_name]), mode_name
    shade = port_oracle.shade_rays(s, rays, 3, ob.MODE_TRUE_NEAREST)
    assert np.array_equal(shade, g["shade_true_nearest"])


def test_kat_answers_are_the_hand_computed_ones():
    g = _golden("kat_rays")
    ans = dict(zip(g["names"], zip(g["prim_true_nearest"], g["t_true_nearest"])))
    exp = {"sph_front": (0, 7.0), "sph_tangent_exact": (0, 10.0), "sph_inside": (0, 3.0), "pln_tri1": (1, 8.0),
           "pln_tri2": (1, 8.0), "pln_from_below": (1, 6.0), "cyl_perp": (2, 9.0), "cyl_inside": (2, 1.0),
           "tri_centre": (3, 5.0), "tri_back": (3, 5.0)}
    for k, (p, t) in exp.items():
        assert ans[k][0] == p and abs(ans[k][1] - t) < 1e-5, k
    for k in ("sph_behind", "sph_miss", "pln_diagonal", "pln_parallel", "pln_outside", "cyl_parallel_axis",
              "cyl_t1_tiny_never_t2", "cyl_miss", "tri_edge_beta0", "tri_vertex", "tri_hypotenuse", "tri_parallel",
              "tri_t_below_eps", "all_miss"):
        assert ans[k][0] == -1, k
    assert ans["tri_t_above_eps"][0] == 3
    assert abs(ans["sph_near_root_rejected"][1] - 6.00005) < 1e-4


def test_port_reproduces_full_frame_pins(port_oracle):
    """SURVEY Appendix B hashes (reference build, lumina defaults, 640x480, depth 10)."""
    cam = scenes.stock_camera(640, 480)
    for key, (obj, cap) in {"bob2000": ("bob_tri.obj", 2000), "tetrahedron": ("tetrahedron.obj", None)}.items():
        rgb, _, _, info = port_oracle.render(scenes.obj_scene(obj, cap), cam, 10, ob.MODE_AS_SHIPPED, aux=False)
        assert ob.fnv1a64(rgb) == PINS["frames_640x480"][key]["fnv1a64"], key
        assert info["rays_total"] == PINS["frames_640x480"][key]["rays_total"], key
    assert PINS["frames_640x480"]["bob2000"]["fnv1a64"] == "cada7080ae5b837a"
    assert PINS["frames_640x480"]["tetrahedron"]["fnv1a64"] == "505bc5c802b4cb83"
    assert PINS["frames_640x480"]["bob_full"]["fnv1a64"] == "dff4eeec81e9e9cc"


def test_threaded_and_column_subset_agree(port_oracle):
    scene, cam, depth, mode = build_case("bobtex_d3")
    full, _, _, _ = port_oracle.render(scene, cam, depth, mode, aux=False)
    part, _, _, info = port_oracle.render(scene, cam, depth, mode, col_begin=3, col_step=4, nthreads=3, aux=False)
    assert info["columns_rendered"] == len(range(3, cam.width, 4))
    assert np.array_equal(part[:, 3::4], full[:, 3::4])
    mask = np.ones(cam.width, bool)
    mask[3::4] = False
    assert not part[:, mask].any()


def test_reference_build_matches_golden_and_port(ref_oracle, port_oracle):
    if ref_oracle is None:
        pytest.skip("reference build not present on this machine (oracle/_ref)")
    for name in ("tetra_d10", "analytic_notetra_d5", "blubmixed_d5"):
        scene, cam, depth, mode = build_case(name)
        rgb, prim, t, info = ref_oracle.render(scene, cam, depth, mode)
        g = _golden(name)
        assert np.array_equal(rgb, g["rgb"]) and np.array_equal(prim, g["prim_id"]) and np.array_equal(t, g["t"])
    # fresh inputs the fixtures do not cover: random small triangle soups, all three modes
    rng = np.random.default_rng(7)
    for trial in range(3):
        n = 200
        c = rng.uniform(-20, 20, (n, 1, 3))
        tri = (c + rng.uniform(-3, 3, (n, 3, 3))).reshape(n, 9).astype(np.float32)
        s = scenes.obj_scene("tetrahedron.obj")
        s.tri_v = tri
        s.tri_material = np.zeros(n, np.uint32)
        s.normalise()
        cam = scenes.stock_camera(96, 64)
        for mode in (ob.MODE_AS_SHIPPED, ob.MODE_BBOX_FIXED, ob.MODE_TRUE_NEAREST):
            a = ref_oracle.render(s, cam, 4, mode)
            b = port_oracle.render(s, cam, 4, mode)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
            assert a[3]["rays_total"] == b[3]["rays_total"]
