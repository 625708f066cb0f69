"""Logic tests that need no GPU: the RT_HD functions the kernels are made of (traversal, primitive
tests, shading, LBVH construction) run on the CPU through tests/emul and are checked against the
golden fixtures of the reference.  This validates the algorithm, NOT the CUDA path — the -m gpu
tests do that through the C ABI."""
import os

import numpy as np
import pytest

import kat
import parity
from bvh_checks import check_bvh
from cases import EDGE_CASES, GOLDEN_CASES, build_case, build_edge_case, random_scene
from conftest import GOLDEN
from oracle import binding as ob


@pytest.fixture(scope="module")
def emul():
    import emul_binding
    return emul_binding.Emulation()


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_emulated_frames_match_the_reference(emul, port_oracle, name):
    scene, cam, depth, mode = build_case(name)
    rgb, prim, t, counts = emul.render(scene, cam, depth)
    # strict check against the true-nearest reference (the linear loop of world.cpp:7-14)
    tr = port_oracle.render(scene, cam, depth, ob.MODE_TRUE_NEAREST)
    parity.assert_parity(parity.compare(rgb, prim, t, tr[0], tr[1], tr[2]), name + " vs true-nearest")
    # and against the as-shipped frame generated from the reference build
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    m = parity.compare(rgb, prim, t, g["rgb"], g["prim_id"], g["t"])
    if name == "analytic_close_d5":
        # camera INSIDE the triangle grid: the shipped grid walk starts behind the origin and stops at the
        # first voxel with a hit (uniform-grid.cpp:190-251) -> 22 px of non-nearest reflections
        assert m["id_match"] == 1.0 and m["colour_within_1"] >= 0.998, m
    else:
        parity.assert_parity(m, name + " vs as-shipped golden")


@pytest.mark.parametrize("name", EDGE_CASES)
def test_emulated_edge_cases_match_the_reference(emul, port_oracle, name):
    """Empty world, no lights, one- and two-leaf hierarchies, zero-area / repeated triangles, duplicate Morton
    keys, frames smaller than a tile: the kernels' logic against the reference's linear loop."""
    scene, cam, depth = build_edge_case(name)
    rgb, prim, t, counts = emul.render(scene, cam, depth, leaf=1)
    tr = port_oracle.render(scene, cam, depth, ob.MODE_TRUE_NEAREST)
    m = parity.compare(rgb, prim, t, tr[0], tr[1], tr[2])
    assert m["id_match"] == 1.0 or m["id_mismatches"] <= 1, m          # (equal-t repeats may pick either copy)
    assert m["colour_within_1"] >= parity.COLOUR_MATCH_MIN and m["t_max_rel"] <= parity.T_REL_TOL, m
    if name == "empty_world":
        assert (prim == -1).all() and counts[1] == 0


@pytest.mark.parametrize("seed", range(8))
def test_emulated_random_worlds_match_the_reference(emul, port_oracle, seed):
    """Fuzz: seeded random worlds (triangle soup + spheres / quads / cylinders with un-normalised axes, diffuse /
    mirror / dielectric / vertex-coloured materials, 1-3 lights, depth 1-5) — the kernels' logic against the
    reference's linear loop at the full parity bar."""
    scene, cam, depth = random_scene(seed)
    rgb, prim, t, counts = emul.render(scene, cam, depth, leaf=1)
    tr = port_oracle.render(scene, cam, depth, ob.MODE_TRUE_NEAREST)
    m = parity.compare(rgb, prim, t, tr[0], tr[1], tr[2])
    assert m["hit_pixels"] > 500, m
    parity.assert_parity(m, f"random world {seed}")


@pytest.mark.parametrize("switch_after", [0, 7, 40])
def test_binary_and_wide_views_are_interchangeable_mid_ray(emul, switch_after):
    """The 4-wide view (k_collapse4) indexes the same nodes as the binary tree, so a ray may change views at any
    step and keep its stack: hits are bit-identical, the number of dependent steps drops."""
    scene, _, _, _ = build_case("blubmixed_d5")
    rng = np.random.default_rng(11)
    o = rng.uniform(-40, 40, (4000, 3))
    rays = np.concatenate([o, rng.uniform(-8, 8, (4000, 3)) - o], axis=1).astype(np.float32)
    p0, t0, s0 = emul.hybrid_walk(scene, rays, 1 << 30)          # binary all the way
    p1, t1, s1 = emul.hybrid_walk(scene, rays, switch_after)
    assert np.array_equal(p0, p1) and np.array_equal(t0, t1)
    assert (p0 >= 0).sum() > 300
    assert s1.sum() < s0.sum() and s1.max() <= s0.max()


def test_emulated_bvh_equals_brute_force(emul):
    scene, cam, depth, _ = build_case("bobtex_d3")
    a = emul.render(scene, cam, depth)
    b = emul.render(scene, cam, depth, brute=1)
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[0], b[0])
    assert a[3][4] < b[3][4] / 50     # the hierarchy prunes > 98 % of the triangle tests


@pytest.mark.parametrize("leaf", [1, 2, 4, 8])
def test_emulated_leaf_sizes_agree(emul, leaf):
    scene, cam, depth, _ = build_case("blubmixed_d5")
    ref = emul.render(scene, cam, depth, leaf=4)
    out = emul.render(scene, cam, depth, leaf=leaf)
    assert np.array_equal(ref[0], out[0]) and np.array_equal(ref[1], out[1])


def test_emulated_kat_rays(emul):
    g = np.load(os.path.join(GOLDEN, "kat_rays.npz"))
    names, rays = kat.kat_rays()
    prim, t = emul.trace_rays(kat.kat_scene(), rays)
    assert np.array_equal(prim, g["prim_true_nearest"]), list(zip(names, prim, g["prim_true_nearest"]))
    hit = prim >= 0
    assert np.allclose(t[hit], g["t_true_nearest"][hit], rtol=1e-5)


@pytest.mark.parametrize("name,leaf", [("bobtex_d3", 4), ("blubmixed_d5", 1), ("synth_small_d1", 4), ("tetra_d10", 2)])
def test_emulated_bvh_structure(emul, name, leaf):
    scene, _, _, _ = build_case(name)
    nodes, order, keys = emul.bvh(scene, leaf)
    info = check_bvh(nodes, order, keys, scene.tri_v, leaf)
    assert info["depth"] <= 64
