"""The parity tests proper: the CUDA path, called through the C ABI, against the golden fixtures
generated from the reference's own Serial sources and against the CPU oracle on fresh inputs.

Bars (tests/parity.py, from BASELINE.json): first-hit primitive id equal on >= 99.9 % of pixels,
8-bit colour within +-1 LSB on >= 99.9 % of pixels, FP32 hit distance within 1e-4 relative."""
import os

import numpy as np
import pytest

import kat
import parity
from bvh_checks import check_bvh
from cases import EDGE_CASES, GOLDEN_CASES, build_case, build_edge_case
from conftest import GOLDEN
from oracle import binding as ob
from realtrace_b200 import api, scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oracle():
    import __graft_entry__ as entry
    entry.build_oracle()
    return ob.best_available()        # the reference's own Serial build where oracle/_ref travelled, else the port


NTHREADS = os.cpu_count() or 1


def make_ctx(scene):
    ctx = api.Context(0)
    ctx.set_scene(scene)
    ctx.commit()
    return ctx


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_frames_match_golden_and_true_nearest(oracle, name):
    scene, cam, depth, mode = build_case(name)
    ctx = make_ctx(scene)
    rgb, prim, t, st = ctx.render(cam, depth, aux=True)
    ctx.close()
    tr = oracle.render(scene, cam, depth, ob.MODE_TRUE_NEAREST)
    parity.assert_parity(parity.compare(rgb, prim, t, tr[0], tr[1], tr[2]), name + " vs true-nearest oracle")
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    m = parity.compare(rgb, prim, t, g["rgb"], g["prim_id"], g["t"])
    if name == "analytic_close_d5":
        # camera inside the triangle grid: the shipped grid walk returns non-nearest hits for a patch of
        # reflections (uniform-grid.cpp:190-251); ids still agree everywhere
        assert m["id_match"] == 1.0 and m["colour_within_1"] >= 0.998, m
    else:
        parity.assert_parity(m, name + " vs golden (reference build, as shipped)")
    assert st["rays_primary"] == cam.width * cam.height


@pytest.mark.parametrize("name", EDGE_CASES)
def test_edge_cases_through_the_abi(oracle, name):
    """Empty world, no lights, one- and two-leaf hierarchies, zero-area / repeated triangles, duplicate Morton
    keys, frames smaller than a tile and than a warp batch — against the reference's linear loop."""
    scene, cam, depth = build_edge_case(name)
    ctx = make_ctx(scene)
    rgb, prim, t, st = ctx.render(cam, depth, aux=True)
    again = ctx.render(cam, depth)[0]
    ctx.close()
    tr = oracle.render(scene, cam, depth, ob.MODE_TRUE_NEAREST)
    m = parity.compare(rgb, prim, t, tr[0], tr[1], tr[2])
    assert m["id_match"] == 1.0 or m["id_mismatches"] <= 1, m          # (equal-t repeats may pick either copy)
    assert m["colour_within_1"] >= parity.COLOUR_MATCH_MIN and m["t_max_rel"] <= parity.T_REL_TOL, m
    assert st["rays_primary"] == cam.width * cam.height
    assert np.array_equal(rgb, again)
    if name == "empty_world":
        assert (prim == -1).all() and st["rays_shadow"] == 0


def test_kat_rays_through_the_abi():
    g = np.load(os.path.join(GOLDEN, "kat_rays.npz"))
    names, rays = kat.kat_rays()
    ctx = make_ctx(kat.kat_scene())
    prim, t = ctx.trace_rays(rays)
    assert np.array_equal(prim, g["prim_true_nearest"]), list(zip(names, prim, g["prim_true_nearest"]))
    hit = prim >= 0
    assert np.allclose(t[hit], g["t_true_nearest"][hit], rtol=1e-5)
    assert np.all(t[~hit] == np.finfo(np.float32).max)
    shade = ctx.shade_rays(rays, 3)
    assert np.allclose(shade, g["shade_true_nearest"], rtol=2e-4, atol=2e-4)
    ctx.close()


def test_random_rays_bvh_vs_brute_force_vs_oracle(oracle):
    scene = scenes.obj_scene("blub_triangulated.obj")
    ctx = make_ctx(scene)
    rng = np.random.default_rng(3)
    n = 20000
    o = rng.uniform(-40, 40, (n, 3))
    target = rng.uniform(-8, 8, (n, 3))
    rays = np.concatenate([o, target - o], axis=1).astype(np.float32)
    prim, t = ctx.trace_rays(rays)
    prim_b, t_b = ctx.trace_rays(rays, flags=api.FLAG_BRUTE_FORCE)
    ctx.close()
    assert np.array_equal(prim, prim_b) and np.array_equal(t, t_b)      # same FP32 tests, different walk
    ref_prim, ref_t = oracle.trace_rays(scene, rays[:4000], ob.MODE_TRUE_NEAREST)
    same = prim[:4000] == ref_prim
    assert same.mean() >= 0.999
    both = same & (ref_prim >= 0)
    assert both.sum() > 500
    assert np.max(np.abs(t[:4000][both] - ref_t[both]) / ref_t[both]) <= 1e-4


@pytest.mark.parametrize("name,leaf", [("bobtex_d3", 4), ("blubmixed_d5", 4), ("synth_small_d1", 4), ("tetra_d10", 4)])
def test_bvh_structure(name, leaf):
    scene, _, _, _ = build_case(name)
    ctx = make_ctx(scene)
    nodes, order, keys = ctx.bvh_download()
    bs = ctx.build_stats()
    ctx.close()
    assert bs["n_triangles"] == len(order)
    info = check_bvh(nodes, order, keys, scene.tri_v, bs["leaf_size"])
    assert info["depth"] <= 64


@pytest.mark.parametrize("name", ["synth_small_d1", "bobtex_d3"])
def test_work_counters_match_a_cpu_walk_of_the_same_tree(name, monkeypatch):
    """N_node / N_tri of the roofline come from counters in the traversal kernels
    (RT_FLAG_COUNT_WORK); a CPU walk of the same tree over the same rays (tests/emul, the kernels' own RT_HD
    functions run serially) must agree within 1 %.  The bounce paths walk the binary tree here like the CPU does
    (by default k_paths uses the 4-wide view, where one visit covers up to three binary nodes)."""
    import emul_binding
    monkeypatch.setenv("RT_WIDE_BVH", "0")
    scene, cam, depth, _ = build_case(name)
    ctx = make_ctx(scene)
    leaf = ctx.build_stats()["leaf_size"]
    st = ctx.render(cam, depth, flags=api.FLAG_COUNT_WORK)[3]
    ctx.close()
    _, _, _, counts = emul_binding.Emulation().render(scene, cam, depth, leaf=leaf)
    rays_gpu = st["rays_primary"] + st["rays_shadow"] + st["rays_secondary"]
    rays_cpu = int(counts[0] + counts[1] + counts[2])
    nodes_gpu = st["node_visits"] + st["shadow_node_visits"]
    tris_gpu = st["tri_tests"] + st["shadow_tri_tests"]
    assert nodes_gpu > 0 and tris_gpu > 0
    assert abs(rays_gpu - rays_cpu) <= 0.001 * rays_cpu, (rays_gpu, rays_cpu)
    assert abs(nodes_gpu - int(counts[3])) <= 0.01 * int(counts[3]), (nodes_gpu, int(counts[3]))
    assert abs(tris_gpu - int(counts[4])) <= 0.01 * int(counts[4]), (tris_gpu, int(counts[4]))


def test_bvh_equals_the_emulated_build():
    """Device radix sort + Karras + refit produce exactly the tree the serial emulation builds."""
    import emul_binding
    scene, _, _, _ = build_case("blubmixed_d5")
    ctx = make_ctx(scene)
    nodes, order, keys = ctx.bvh_download()
    leaf = ctx.build_stats()["leaf_size"]
    ctx.close()
    e_nodes, e_order, e_keys = emul_binding.Emulation().bvh(scene, leaf)
    assert np.array_equal(order, e_order)
    assert np.array_equal(keys, e_keys)
    assert np.array_equal(nodes.view(np.uint32), e_nodes.view(np.uint32))


@pytest.mark.parametrize("n", [1, 2, 31, 4096, 4097, 100000, 1 << 20])
def test_device_radix_sort(n):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 1 << 63, n, dtype=np.uint64)
    if n > 100:
        keys[::7] = keys[3]          # duplicates: stability matters
        keys[n // 2:] &= np.uint64(0xFFFFFFFF)
    vals = np.arange(n, dtype=np.uint32)
    ctx = api.Context(0)
    k, v = ctx.sort_pairs(keys, vals)
    ctx.close()
    ref = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[ref])
    assert np.array_equal(v, vals[ref])


@pytest.mark.parametrize("pattern", ["low32", "byte2_only", "all_equal", "top_byte_only"])
@pytest.mark.parametrize("n", [2, 4097, 300000])
def test_device_radix_sort_passes_decided_on_the_device(n, pattern):
    """Passes in which every key has the same digit are recognised by the kernels themselves (no host read-back): their
    scatter is a copy and the result still ends in buffer 0.  Keys that leave 4, 7, 8 and 7 of the 8 passes trivial."""
    rng = np.random.default_rng(n)
    if pattern == "low32":
        keys = rng.integers(0, 1 << 32, n, dtype=np.uint64)
    elif pattern == "byte2_only":
        keys = rng.integers(0, 256, n, dtype=np.uint64) << np.uint64(16) | np.uint64(0x5A0000000000005A)
    elif pattern == "all_equal":
        keys = np.full(n, 0x0123456789ABCDEF, np.uint64)
    else:
        keys = rng.integers(0, 128, n, dtype=np.uint64) << np.uint64(56)
    vals = rng.permutation(n).astype(np.uint32)
    ctx = api.Context(0)
    k, v = ctx.sort_pairs(keys, vals)
    ctx.close()
    ref = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[ref])
    assert np.array_equal(v, vals[ref])


def test_refit_is_idempotent_and_tracks_vertices(oracle):
    scene, cam, depth, _ = build_case("bobtex_d3")
    ctx = make_ctx(scene)
    base = ctx.render(cam, depth, aux=True)
    nodes0, order0, _ = ctx.bvh_download()
    ctx.commit(api.COMMIT_REFIT)
    nodes1, order1, _ = ctx.bvh_download()
    assert np.array_equal(nodes0.view(np.uint32), nodes1.view(np.uint32)) and np.array_equal(order0, order1)
    again = ctx.render(cam, depth, aux=True)
    assert np.array_equal(base[0], again[0]) and np.array_equal(base[1], again[1])
    # move the model: refit must follow (same topology), and match a fresh build of the moved scene
    moved = scene.tri_v.copy()
    moved[:, 1::3] += 1.5
    ctx.update_vertices(moved)
    ctx.commit(api.COMMIT_REFIT)
    r = ctx.render(cam, depth, aux=True)
    ctx.close()
    scene2, _, _, _ = build_case("bobtex_d3")
    scene2.tri_v = moved
    tr = oracle.render(scene2, cam, depth, ob.MODE_TRUE_NEAREST)
    parity.assert_parity(parity.compare(r[0], r[1], r[2], tr[0], tr[1], tr[2]), "refit after translation")


def test_vertices_deformed_on_the_device_refit_and_rebuild():
    """an animation step that never leaves the GPU.  The vertices are deformed by a device-side
    op (torch stands in for the caller's kernel, on the context's stream), handed over with
    rt_scene_update_vertices_device, and the LBVH is refitted; frames must equal those of a context that was built
    from the same deformed vertices, and a later BUILD commit must start from the device-side vertices."""
    import torch
    scene, cam, depth, _ = build_case("bobtex_d3")
    ctx = make_ctx(scene)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream or 1)
    v = torch.from_numpy(np.ascontiguousarray(scene.tri_v, np.float32)).cuda().reshape(-1, 3, 3)
    for step in (1, 2):
        phase = 0.7 * step
        d = v.clone()
        d[:, :, 1] += 0.8 * torch.sin(0.35 * v[:, :, 0] + phase)         # a travelling wave along x
        d[:, :, 2] += 0.5 * torch.cos(0.25 * v[:, :, 1] - phase)
        ctx.update_vertices_device(d.data_ptr(), d.shape[0])
        ctx.commit(api.COMMIT_REFIT, want_stats=False)
        got = ctx.render(cam, depth, aux=True)
        scene2, _, _, _ = build_case("bobtex_d3")
        scene2.tri_v = d.reshape(-1, 9).cpu().numpy()
        fresh = make_ctx(scene2)
        want = fresh.render(cam, depth, aux=True)
        fresh.close()
        m = parity.compare(got[0], got[1], got[2], want[0], want[1], want[2])
        parity.assert_parity(m, f"device-side deformation step {step}: refit vs fresh build")
    # a rebuild without new rt_scene_set_triangles picks the device-side vertices up
    ctx.commit(api.COMMIT_BUILD)
    again = ctx.render(cam, depth, aux=True)
    ctx.close()
    assert np.array_equal(again[0], want[0]) and np.array_equal(again[1], want[1])


@pytest.mark.parametrize("world", [2, 3, 8])
def test_tile_sharding_is_bit_identical(world):
    """N logical ranks on one GPU: every rank renders its interleaved tiles; assembling the packed
    tiles reproduces the single-rank frame bit for bit."""
    import torch
    scene, cam, depth, _ = build_case("blubmixed_d5")
    cam.width, cam.height = 200, 150          # not a multiple of the tile: ragged edge tiles
    ctx = make_ctx(scene)
    full, _, _, st_full = ctx.render(cam, depth)
    frame = torch.zeros(cam.height * cam.width * 3, dtype=torch.uint8, device="cuda")
    rays = 0
    for r in range(world):
        _, owned, tb = api.tile_layout(cam.width, cam.height, 0, 0, r, world)
        packed = torch.zeros(max(owned * tb, 1), dtype=torch.uint8, device="cuda")
        st = ctx.render_device(cam, depth, packed.data_ptr(), rank=r, world=world, flags=api.FLAG_PACKED_TILES)
        rays += st["rays_primary"]
        ctx.assemble_tiles(packed.data_ptr(), r, world, cam.width, cam.height, frame.data_ptr())
    torch.cuda.synchronize()
    ctx.close()
    assert rays == cam.width * cam.height
    out = frame.cpu().numpy().reshape(cam.height, cam.width, 3)
    assert np.array_equal(out, full)


@pytest.mark.parametrize("name", ["blubmixed_d5", "synth_small_d1"])
def test_packed_tiles_keep_their_slots_when_the_tile_order_changes(name):
    """The heavy-tiles-first feedback re-sorts a rank's tile list after the first frames; the packed slot of a
    tile (t / world) must not move with it.  Frames 1..5 of every logical rank assemble to the single-rank
    frame (bounce scene: accumulator + k_resolve path; bounce-free scene: direct RGB8 stores)."""
    import torch
    scene, cam, depth, _ = build_case(name)
    cam.width, cam.height = 712, 404          # 12 x 13 tiles, ragged right and bottom edges
    world = 3
    ctx = make_ctx(scene)
    full, _, _, _ = ctx.render(cam, depth)
    for frame_no in range(5):
        frame = torch.zeros(cam.height * cam.width * 3, dtype=torch.uint8, device="cuda")
        for r in range(world):
            _, owned, tb = api.tile_layout(cam.width, cam.height, 0, 0, r, world)
            packed = torch.zeros(max(owned * tb, 1), dtype=torch.uint8, device="cuda")
            # the layout (and with it the learned order) is per context: give every logical rank its own frames
            for _ in range(frame_no + 1):
                ctx.render_device(cam, depth, packed.data_ptr(), rank=r, world=world, flags=api.FLAG_PACKED_TILES)
            ctx.assemble_tiles(packed.data_ptr(), r, world, cam.width, cam.height, frame.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(frame.cpu().numpy().reshape(cam.height, cam.width, 3), full), frame_no
    ctx.close()


def test_full_size_property_checks():
    """Config-2 size (1920x1080, 2 lights, depth 3): size-independent properties — every ray kind is
    counted, rendering twice is bit-identical, brute force over the same rays gives the same frame on
    the hit region, background pixels carry exactly the background colour."""
    scene, cam, depth, _ = scenes.workload("bob1080")
    ctx = make_ctx(scene)
    a = ctx.render(cam, depth, aux=True)
    b = ctx.render(cam, depth, aux=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    st = a[3]
    hits = int((a[1] >= 0).sum())
    assert st["rays_primary"] == 1920 * 1080
    assert st["rays_shadow"] >= 2 * hits              # two lights per shaded hit, bounces add more
    assert 0 < st["rays_secondary"] <= 3 * hits
    miss = a[1] < 0
    assert np.all(a[0][miss] == np.array([25, 76, 153], np.uint8))
    small = scenes.stock_camera(480, 270)
    c = ctx.render(small, depth, aux=True)
    d = ctx.render(small, depth, aux=True, flags=api.FLAG_BRUTE_FORCE)
    ctx.close()
    assert np.array_equal(c[0], d[0]) and np.array_equal(c[1], d[1])


def test_errors_are_reported_not_swallowed():
    ctx = api.Context(0)
    with pytest.raises(api.RtError) as e:
        ctx.render(scenes.stock_camera(64, 48), 1)
    assert e.value.code == -4
    scene, cam, depth, _ = build_case("tetra_d10")
    ctx.set_scene(scene)
    ctx.commit()
    with pytest.raises(api.RtError):
        ctx.render(cam, depth, tile=(30, 30))
    ctx.close()


@pytest.mark.parametrize("world,pool_div", [(2, 2), (4, 4), (8, 3)])
def test_dynamic_tile_stealing_is_bit_identical(world, pool_div):
    """Logical ranks rendered one after the other into ONE frame: every pool_div-th tile group is not
    owned by anybody; the rank that gets there first claims its 8x4 blocks from the shared cursor.
    The frame must equal the single-rank frame bit for bit, every pixel traced exactly once."""
    import torch
    scene, cam, depth, _ = build_case("blubmixed_d5")
    cam.width, cam.height = 456, 270          # ragged edge tiles, 8 x 9 tiles
    ctx = make_ctx(scene)
    full = ctx.render(cam, depth)[0]
    for frame_index in (0, 1, 70):            # 70: slot wrap-around of the 64-entry cursor array
        frame = torch.zeros(cam.height * cam.width * 3, dtype=torch.uint8, device="cuda")
        primary, stolen = 0, 0
        order = list(range(world)) if frame_index != 1 else list(reversed(range(world)))
        for r in order:
            st = ctx.render_device(cam, depth, frame.data_ptr(), rank=r, world=world,
                                   steal=(pool_div, frame_index, None))
            primary += st["rays_primary"]
            stolen += st["stolen_blocks"]
        torch.cuda.synchronize()
        out = frame.cpu().numpy().reshape(cam.height, cam.width, 3)
        assert primary == cam.width * cam.height
        assert stolen > 0
        assert np.array_equal(out, full), frame_index
    ctx.close()


def test_peer_handshake_and_tile_push_logical_ranks():
    """bench.py's N > 1 frame loop with logical ranks on one GPU and one stream: each rank renders its tiles
    into a local packed buffer (direct RGB8), pushes them into the shared frame and goes through the
    peer-memory handshake.  Rank 0 opens the frame (phase 0 publishes "earlier frames consumed"), ranks 1..N-1
    are enqueued before rank 0's completion wait, so no kernel ever has to spin."""
    import torch
    scene, cam, depth, _ = build_case("synth_small_d1")      # no bounces: direct RGB8 path
    cam.width, cam.height = 328, 200
    world = 4
    ctx = make_ctx(scene)
    full = ctx.render(cam, depth)[0]
    frame_ptr, _ = ctx.shared_buffer_create(cam.width * cam.height * 3)
    sync_ptr, _ = ctx.shared_buffer_create(1024)
    _, _, tb = api.tile_layout(cam.width, cam.height)
    for k in range(3):
        ctx.peer_sync(sync_ptr, 0, world, k, 0)
        for r in list(range(1, world)) + [0]:
            _, owned, _ = api.tile_layout(cam.width, cam.height, 0, 0, r, world)
            packed = torch.zeros(max(owned * tb, 1), dtype=torch.uint8, device="cuda")
            ctx.render_device(cam, depth, packed.data_ptr(), rank=r, world=world, flags=api.FLAG_PACKED_TILES,
                              want_stats=False)
            ctx.peer_sync(sync_ptr, r, world, k, 0)
            ctx.assemble_tiles(packed.data_ptr(), r, world, cam.width, cam.height, frame_ptr)
            ctx.peer_sync(sync_ptr, r, world, k, 1)
        ctx.synchronize()                                     # raises on a handshake time-out
        out = np.zeros((cam.height, cam.width, 3), np.uint8)
        ctx.download(frame_ptr, out)
        assert np.array_equal(out, full), k
    sync = np.zeros(256, np.uint32)
    ctx.download(sync_ptr, sync)
    ctx.close()
    assert sync[64] == 2                                      # rank 0 opened frame 2: frames 0 and 1 consumed


@pytest.mark.parametrize("name", ["synth_small_d1", "blubmixed_d5", "bobtex_d3"])
def test_work_order_feedback_never_changes_a_pixel(name):
    """Heavy-tiles-first re-sort: from the second frame on the kernel walks the tiles in the order of their
    measured cost.  Twelve frames (re-sorts after frames 0, 1 and 8): every frame equals the first, and every
    pixel is traced exactly once."""
    scene, cam, depth, _ = build_case(name)
    cam.width, cam.height = 648, 366
    ctx = make_ctx(scene)
    first = None
    for k in range(12):
        rgb, _, _, st = ctx.render(cam, depth)
        assert st["rays_primary"] == cam.width * cam.height, k
        if first is None:
            first = rgb.copy()
        else:
            assert np.array_equal(rgb, first), k
    # one third of the frame as a logical rank (packed output): same pixels
    import torch
    _, owned, tb = api.tile_layout(cam.width, cam.height, 32, 16, 1, 3)
    packed = torch.zeros(owned * tb, dtype=torch.uint8, device="cuda")
    outs = []
    for k in range(10):
        st = ctx.render_device(cam, depth, packed.data_ptr(), tile=(32, 16), rank=1, world=3, flags=api.FLAG_PACKED_TILES)
        torch.cuda.synchronize()
        outs.append(packed.cpu().numpy().copy())
        assert np.array_equal(outs[-1], outs[0]), k
    frame = torch.zeros(cam.height * cam.width * 3, dtype=torch.uint8, device="cuda")
    ctx.assemble_tiles(packed.data_ptr(), 1, 3, cam.width, cam.height, frame.data_ptr(), tile=(32, 16))
    torch.cuda.synchronize()
    ctx.close()
    got = frame.cpu().numpy().reshape(cam.height, cam.width, 3)
    tx, ty = (cam.width + 31) // 32, (cam.height + 15) // 16
    for t in range(1, tx * ty, 3):                         # rank 1's tiles
        x0, y0 = (t % tx) * 32, (t // tx) * 16
        assert np.array_equal(got[y0:y0 + 16, x0:x0 + 32], first[y0:y0 + 16, x0:x0 + 32]), t


@pytest.mark.parametrize("name", ["synth_small_d1", "analytic_stock_d1", "one_triangle", "two_triangles",
                                  "degenerate_and_duplicate_triangles", "coincident_centroids"])
@pytest.mark.parametrize("div", [1, 3, 0])
def test_heavy_tiles_on_the_wide_view_never_change_a_pixel(name, div, monkeypatch):
    """RT_WIDE_HEAVY: the batches of the heaviest tiles (the first 1/div of the cost-sorted order) walk the 4-wide view
    of the tree inside the same launch, the others the binary tree; div = 0: batches change to the wide view mid-walk
    (RT_WIDE_AFTER_BURSTS).  Both views index the same nodes and tie-break
    equal distances by object order, so frames, first-hit ids and distances are those of the binary walk — for the
    whole frame (k_frame) and for a logical rank's share pushed into a shared frame (k_frame_push)."""
    import torch
    if name in EDGE_CASES:
        scene, cam, depth = build_edge_case(name)
    else:
        scene, cam, depth, _ = build_case(name)
    if name == "analytic_stock_d1":            # bounce-free variant of the stock scene: the one-launch frame needs kr = 0
        scene.materials["kr"] = 0.0
        scene.materials["kt"] = 0.0
    cam.width, cam.height = 328, 203
    monkeypatch.setenv("RT_WIDE_HEAVY", "0")
    ref_ctx = make_ctx(scene)
    ref = ref_ctx.render(cam, depth, aux=True)
    ref_ctx.close()
    monkeypatch.setenv("RT_WIDE_HEAVY", "2")
    if div:
        monkeypatch.setenv("RT_WIDE_HEAVY_DIV", str(div))
        monkeypatch.setenv("RT_WIDE_AFTER_BURSTS", "0")
    else:
        # no tile starts on the wide view; every batch moves to it after 2 bursts of 4 steps, i.e. nearly every ray
        # changes views in the middle of its walk (stack and current node carry over)
        monkeypatch.setenv("RT_WIDE_HEAVY_DIV", "65536")
        monkeypatch.setenv("RT_WIDE_AFTER_BURSTS", "2")
        monkeypatch.setenv("RT_LOOP_PRIMARY", "4")
    ctx = make_ctx(scene)
    for k in range(4):                         # frame 0 runs in spatial order (binary walk only), then the sorted order
        got = ctx.render(cam, depth, tile=(32, 16), aux=True)
        assert np.array_equal(got[0], ref[0]), k
        assert np.array_equal(got[1], ref[1]), k
        assert np.array_equal(got[2], ref[2]), k
        assert got[3]["rays_primary"] == cam.width * cam.height
    # a logical rank's share through rt_render_push (rank 1 of 2; the frame lives on this GPU)
    world, rank, tile = 2, 1, (32, 16)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream or 1)
    frame = torch.zeros(cam.width * cam.height * 3, dtype=torch.uint8, device="cuda")
    sync_ptr, _ = ctx.shared_buffer_create(1024)
    _, owned, tb = api.tile_layout(cam.width, cam.height, tile[0], tile[1], rank, world)
    packed = torch.zeros(max(owned * tb, 1), dtype=torch.uint8, device="cuda")
    cs = api.camera_struct(cam)
    params = api.Context._params(depth, tile=tile, rank=rank, world=world, flags=api.FLAG_PACKED_TILES)
    tx, ty = (cam.width + tile[0] - 1) // tile[0], (cam.height + tile[1] - 1) // tile[1]
    for k in range(4):
        frame.fill_(0x5a)
        ctx.peer_sync(sync_ptr, 0, world, k, 0)
        ctx.render_push(cs, params, packed.data_ptr(), frame.data_ptr(), sync_ptr, k)
        ctx.synchronize()
        got = frame.cpu().numpy().reshape(cam.height, cam.width, 3)
        for t in range(rank, tx * ty, world):
            x0, y0 = (t % tx) * tile[0], (t // tx) * tile[1]
            assert np.array_equal(got[y0:y0 + tile[1], x0:x0 + tile[0]], ref[0][y0:y0 + tile[1], x0:x0 + tile[0]]), (k, t)
    ctx.close()


@pytest.mark.parametrize("name,tile", [("synth_small_d1", (0, 0)), ("synth_small_d1", (32, 16)), ("synth_small_d1", (8, 4)),
                                       ("blubmixed_d5", (0, 0))])
def test_render_push_one_call_per_rank(name, tile):
    """rt_render_push with logical ranks on one GPU.  Bounce-free scene: ONE kernel traces, shades and copies
    every finished tile into the shared frame (the warp that completes a tile pushes it); scene with bounces:
    render + separate push.  One context per logical rank (all on ONE stream, enqueued in an order in which no
    handshake kernel has to wait), five frames, so every rank's heavy-tiles-first re-sort happens in between;
    8x4 tiles make every batch a tile of its own and leave 24-byte rows to the byte path."""
    import torch
    scene, cam, depth, _ = build_case(name)
    cam.width, cam.height = 328, 203
    world = 3
    ctx = make_ctx(scene)
    full = ctx.render(cam, depth)[0]
    ctxs = [ctx] + [make_ctx(scene) for _ in range(world - 1)]
    for c in ctxs:
        c.set_stream(torch.cuda.current_stream().cuda_stream or 1)
    frame = torch.zeros(cam.width * cam.height * 3, dtype=torch.uint8, device="cuda")
    frame_ptr = frame.data_ptr()
    sync_ptr, _ = ctx.shared_buffer_create(1024)
    cs = api.camera_struct(cam)
    packed = []
    for r in range(world):
        _, owned, tb = api.tile_layout(cam.width, cam.height, tile[0], tile[1], r, world)
        packed.append(torch.zeros(max(owned * tb, 1), dtype=torch.uint8, device="cuda"))
    for k in range(5):
        frame.fill_(0x5a)                                     # a tile that is not pushed shows
        ctx.peer_sync(sync_ptr, 0, world, k, 0)
        for r in list(range(1, world)) + [0]:
            params = api.Context._params(depth, tile=tile, rank=r, world=world, flags=api.FLAG_PACKED_TILES)
            ctxs[r].render_push(cs, params, packed[r].data_ptr(), frame_ptr, sync_ptr, k)
        for c in ctxs:
            c.synchronize()                                   # raises on a handshake time-out / queue overflow
        assert np.array_equal(frame.cpu().numpy().reshape(cam.height, cam.width, 3), full), k
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("name,col_step", [("synth1m", 16), ("blub4k", 16), ("bob1080", 16)])
def test_full_size_frames_against_oracle_column_sample(oracle, name, col_step):
    """BASELINE configs 4, 3 and 2 at their full size (3840x2160, 3840x2160, 1920x1080): the GPU frame against the
    CPU oracle (grid as shipped, all host threads) on every 16th column, plus size-independent properties
    (idempotence, ray accounting)."""
    scene, cam, depth, _ = scenes.workload(name)
    ctx = make_ctx(scene)
    rgb, prim, t, st = ctx.render(cam, depth, aux=True)
    rgb2, _, _, st2 = ctx.render(cam, depth)
    ctx.close()
    assert np.array_equal(rgb, rgb2)
    assert st["rays_primary"] == cam.width * cam.height
    assert st["rays_shadow"] == st2["rays_shadow"] and st["rays_secondary"] == st2["rays_secondary"]
    begin = col_step // 2 + 1
    ref_rgb, ref_prim, ref_t, info = oracle.render(scene, cam, depth, ob.MODE_AS_SHIPPED, col_begin=begin, col_step=col_step,
                                                   nthreads=NTHREADS)
    cols = slice(begin, None, col_step)
    m = parity.compare(rgb[:, cols], prim[:, cols], t[:, cols], ref_rgb[:, cols], ref_prim[:, cols], ref_t[:, cols])
    assert m["hit_pixels"] > 4000, m
    parity.assert_parity(m, name + " column sample vs oracle")
    if name == "synth1m":       # no dielectrics: the GPU traces exactly the oracle's rays on those columns
        hits = int((prim[:, cols] >= 0).sum())
        assert info["rays_total"] == m["pixels"] + hits * len(scene.lights)


def test_reference_default_depth_10_with_dielectrics(oracle):
    """RECURSION_DEPTH 10 (world.h:11) on the dielectric/mirror scene: deep level*2 chains through k_paths."""
    scene, cam, _, _ = build_case("blubmixed_d5")
    ctx = make_ctx(scene)
    rgb, prim, t, st = ctx.render(cam, 10, aux=True)
    ctx.close()
    tr = oracle.render(scene, cam, 10, ob.MODE_TRUE_NEAREST)
    parity.assert_parity(parity.compare(rgb, prim, t, tr[0], tr[1], tr[2]), "blub depth 10")
    assert st["rays_secondary"] > 0


def test_nine_lights_take_the_unfused_path(oracle):
    """More than 8 lights: occlusion no longer fits the per-hit bit mask, so the separate any-hit kernel and
    the wave loop are used — same frame as the oracle."""
    scene, cam, depth, _ = build_case("bobtex_d3")
    rng = np.random.default_rng(5)
    lights = np.concatenate([rng.uniform(-40, 40, (9, 3)) + np.array([0, 45, 0]), rng.uniform(0.05, 0.2, (9, 3))], axis=1)
    scene.lights = lights.astype(np.float32)
    scene.normalise()
    ctx = make_ctx(scene)
    rgb, prim, t, st = ctx.render(cam, depth, aux=True)
    ctx.close()
    tr = oracle.render(scene, cam, depth, ob.MODE_TRUE_NEAREST)
    parity.assert_parity(parity.compare(rgb, prim, t, tr[0], tr[1], tr[2]), "nine lights")
    hits = int((prim >= 0).sum())
    assert st["rays_shadow"] >= 9 * hits


# ---- BASELINE config 1 at its stated size -------------------------------------------------------------------------
@pytest.mark.parametrize("camera", ["stock", "close"])
@pytest.mark.parametrize("depth", [1, 5])
def test_config1_full_frame_640x480(oracle, camera, depth):
    """configs[0]: the Serial renderer's built-in scene — spheres, quad, cylinder (the literal of lumina.cpp:312-356)
    + tetrahedron.obj — at 640x480 with the stock camera of lumina.cpp:302-306 and with the close camera SURVEY 8(d)
    adds for coverage; whole frame against the oracle in both modes (depth 1 is the config's, depth 5 exercises the
    mirror sphere, the floor and the dielectric cylinder)."""
    scene = scenes.analytic_scene()
    cam = scenes.stock_camera(640, 480) if camera == "stock" else scenes.close_camera(640, 480)
    ctx = make_ctx(scene)
    rgb, prim, t, st = ctx.render(cam, depth, aux=True)
    ctx.close()
    assert st["rays_primary"] == 640 * 480
    tr = oracle.render(scene, cam, depth, ob.MODE_TRUE_NEAREST, nthreads=NTHREADS)
    m = parity.compare(rgb, prim, t, tr[0], tr[1], tr[2])
    parity.assert_parity(m, f"config 1, {camera} camera, depth {depth}, vs true-nearest oracle")
    sh = oracle.render(scene, cam, depth, ob.MODE_AS_SHIPPED, nthreads=NTHREADS)
    m2 = parity.compare(rgb, prim, t, sh[0], sh[1], sh[2])
    if camera == "close" and depth > 1:
        # camera inside the triangle grid: the shipped walk stops at the first voxel with a hit (Q13) and returns
        # non-nearest hits for a patch of reflections; first-hit ids still agree everywhere
        assert m2["id_match"] == 1.0 and m2["colour_within_1"] >= 0.998, m2
    else:
        parity.assert_parity(m2, f"config 1, {camera} camera, depth {depth}, vs the reference as shipped")
    if camera == "close":
        assert m["hit_pixels"] > 100000, m          # the objects fill the frame


# ---- BASELINE config 5: orbit frames with a refit before each ----------------------------------------------------
def test_config5_orbit_frames_with_refit(oracle):
    """configs[4]: the 120-frame orbit of Parellel/interactive_camera.cu:64-81 over textured bob at 1920x1080,
    rt_scene_commit(REFIT) before every frame.  Frames k = 0, 30, 60, 90 against the oracle on every 8th column
    (BBox initialiser fixed: 26x faster than the shipped grid and the same hits as the linear loop on this mesh,
    SURVEY 8d), and the refit must leave the BVH bit-identical (static geometry)."""
    scene, cam0, depth, _ = scenes.workload("orbit")
    ctx = make_ctx(scene)
    nodes0, order0, _ = ctx.bvh_download()
    for k in (0, 30, 60, 90):
        cam = scenes.orbit_camera(k, width=1920, height=1080)
        ctx.commit(api.COMMIT_REFIT, want_stats=False)
        rgb, prim, t, st = ctx.render(cam, depth, aux=True)
        assert st["rays_primary"] == 1920 * 1080
        step, begin = 8, 3
        ref = oracle.render(scene, cam, depth, ob.MODE_BBOX_FIXED, col_begin=begin, col_step=step, nthreads=NTHREADS)
        cols = slice(begin, None, step)
        m = parity.compare(rgb[:, cols], prim[:, cols], t[:, cols], ref[0][:, cols], ref[1][:, cols], ref[2][:, cols])
        assert m["hit_pixels"] > 2000, (k, m)
        parity.assert_parity(m, f"orbit frame {k}")
    nodes1, order1, _ = ctx.bvh_download()
    ctx.close()
    assert np.array_equal(nodes0.view(np.uint32), nodes1.view(np.uint32)) and np.array_equal(order0, order1)


# ---- advisor findings of round 1 -----------------------------------------------------------------------------------
def test_depth_zero_dielectric_traces_the_refracted_child(oracle):
    """max_depth 0 with a dielectric: the refracted child has level 0 * 2 = 0 (world.cpp:98) and is alive; the
    frame must carry its contribution (and no spurious queue overflow)."""
    from cases import glass_ball_scene
    scene = glass_ball_scene()
    cam = scenes.close_camera(320, 240)
    ctx = make_ctx(scene)
    rgb, prim, t, st = ctx.render(cam, 0, aux=True)
    rgb_nostats = np.zeros_like(rgb)
    ctx._check(ctx.lib.rt_render(ctx.h, api.camera_struct(cam), api.Context._params(0), rgb_nostats.ctypes.data, None, None))
    ctx.close()
    assert st["rays_secondary"] > 0
    assert np.array_equal(rgb, rgb_nostats)
    tr = oracle.render(scene, cam, 0, ob.MODE_TRUE_NEAREST, nthreads=NTHREADS)
    parity.assert_parity(parity.compare(rgb, prim, t, tr[0], tr[1], tr[2]), "depth 0 with a dielectric sphere")
    g = np.load(os.path.join(GOLDEN, "glass_ball_d0.npz"))
    parity.assert_parity(parity.compare(rgb, prim, t, g["rgb"], g["prim_id"], g["t"]), "depth-0 golden (true nearest)")


def test_refit_after_the_mesh_moved_far_keeps_its_boxes_tight(oracle):
    """The absolute pad of the node boxes follows the CURRENT extent: a refit of a mesh that moved far beyond its
    original extent (x 8) must still be watertight — first hits against the oracle on the moved scene (colours are
    not compared at the full bar: that far from the origin FP32 hit points carry 6 x the rounding error)."""
    scene, cam, depth, _ = build_case("bob2000_d10")
    ctx = make_ctx(scene)
    moved = (np.asarray(scene.tri_v, np.float32).reshape(-1, 9) + np.asarray([300, 0, 200] * 3, np.float32)).copy()
    ctx.update_vertices(moved)
    ctx.commit(api.COMMIT_REFIT)
    from realtrace_b200.scene import Camera
    cam2 = Camera(pos=(340.0, 40.0, 200.0), target=(300.0, 0.0, 200.0), up=(0.0, 1.0, 0.0), fovy=45.0, width=200, height=150)
    r = ctx.render(cam2, 3, aux=True)
    ctx.close()
    scene2, _, _, _ = build_case("bob2000_d10")
    scene2.tri_v = moved
    tr = oracle.render(scene2, cam2, 3, ob.MODE_TRUE_NEAREST)
    m = parity.compare(r[0], r[1], r[2], tr[0], tr[1], tr[2])
    assert m["hit_pixels"] > 500, m
    assert m["id_mismatches_on_hits"] <= parity.hit_budget(m, 0.002) and m["t_max_rel"] <= parity.T_REL_TOL, m
    assert m["colour_within_1"] >= 0.99, m


def test_device_vertices_survive_a_scene_call_before_the_rebuild():
    """rt_scene_update_vertices_device, then another rt_scene_set_* call (clears `committed`), then BUILD: the
    rebuild must start from the device-side vertices, not silently revert to the old host copy."""
    import torch
    scene, cam, depth, _ = build_case("bobtex_d3")
    ctx = make_ctx(scene)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream or 1)
    v = torch.from_numpy(np.ascontiguousarray(scene.tri_v, np.float32)).cuda().reshape(-1, 3, 3)
    d = v.clone()
    d[:, :, 1] += 0.8 * torch.sin(0.35 * v[:, :, 0])
    ctx.update_vertices_device(d.data_ptr(), d.shape[0])
    lights = np.ascontiguousarray(scene.lights, np.float32)
    ctx._check(ctx.lib.rt_scene_set_lights(ctx.h, lights.ctypes.data, len(lights)))      # clears `committed`
    ctx.commit(api.COMMIT_BUILD)
    got = ctx.render(cam, depth, aux=True)
    ctx.close()
    scene2, _, _, _ = build_case("bobtex_d3")
    scene2.tri_v = d.cpu().numpy().reshape(-1, 9)
    want_ctx = make_ctx(scene2)
    want = want_ctx.render(cam, depth, aux=True)
    want_ctx.close()
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])


def build_overflow_case():
    """A valid scene that trips a device-side error: one dielectric sphere seen from close by at depth 120.  Inside
    the sphere every internal reflection (level + 1) parks a refracted sibling (level * 2, alive up to level 60), so a
    lane of k_paths needs more than its RT_PATH_STACK = 24 parked rays."""
    from realtrace_b200.scene import Scene, make_materials
    mats = make_materials([dict(color=(1.0, 1.0, 1.0), ka=0.4, kd=0.9, ks=0.4, kr=0.1, kt=0.8, eta=1.5)])
    s = Scene(sph=[(0.0, 0.0, 0.0, 6.0)], sph_material=[0], sph_object_id=[0], materials=mats,
              lights=np.asarray([scenes.STOCK_LIGHT], np.float32), ambient=scenes.STOCK_AMBIENT,
              background=scenes.STOCK_BACKGROUND, name="dielectric_ball").normalise()
    return s, scenes.close_camera(96, 64), 120


def test_render_without_stats_reports_kernel_errors(monkeypatch):
    """rt_render(stats = NULL) must look at the device error word too: a dielectric hall of mirrors at depth 120
    overflows the per-lane stack of parked rays in k_paths.  (With RT_PATH_SHARE=1 the idle lanes of the warp take
    the oldest parked rays and this scene no longer overflows, so the hand-over is switched off here.)"""
    monkeypatch.setenv("RT_PATH_SHARE", "0")
    scene, cam, depth = build_overflow_case()
    ctx = make_ctx(scene)
    rgb = np.zeros((cam.height, cam.width, 3), np.uint8)
    with pytest.raises(api.RtError) as e1:
        ctx.render(cam, depth)
    rc = ctx.lib.rt_render(ctx.h, api.camera_struct(cam), api.Context._params(depth), rgb.ctypes.data, None, None)
    ctx.close()
    assert e1.value.code == -5 and rc == -5
