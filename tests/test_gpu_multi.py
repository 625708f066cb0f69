"""One context over several GPUs (rt_create_multi) and the pipelined host-frame path (rt_render_enqueue /
rt_render_wait): what a program written against the reference's classes gets from RenderEngine::render().

On a box with one GPU the multi-device tests run with a one-device context (same entry points, same code path up to
the tile split); with two or more GPUs they compare the frame the GPUs assembled with the frame one GPU renders —
every pixel is computed by exactly one device with the same code and data, so the two must be byte-identical."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import parity
from cases import build_case
from oracle import binding as ob
from realtrace_b200 import api, scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSETS = os.path.join(ROOT, "assets")


def n_gpus():
    import torch
    return torch.cuda.device_count()


def one_gpu_frame(scene, cam, depth):
    ctx = api.Context(0)
    ctx.set_scene(scene)
    ctx.commit()
    out = ctx.render(cam, depth, aux=True)
    ctx.close()
    return out


@pytest.mark.parametrize("name", ["synth_small_d1", "bobtex_d3", "blubmixed_d5", "analytic_close_d5"])
def test_multi_device_context_renders_the_one_gpu_frame(name):
    scene, cam, depth, _ = build_case(name)
    cam.width, cam.height = 488, 274                      # ragged: not a multiple of any tile size
    want = one_gpu_frame(scene, cam, depth)
    ctx = api.Context(devices="all")
    assert ctx.device_count() == n_gpus()
    ctx.set_scene(scene)
    ctx.commit()
    for k in range(3):                                    # the heavy-tiles-first re-sort happens in between
        got = ctx.render(cam, depth, aux=True)
        assert np.array_equal(got[0], want[0]), (name, k)
        assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
        for key in ("rays_primary", "rays_shadow", "rays_secondary"):
            assert got[3][key] == want[3][key], (key, got[3], want[3])
    ctx.close()


@pytest.mark.parametrize("devices", [[0], "all"])
def test_pipelined_frames_equal_synchronous_frames(devices):
    """rt_render_enqueue / rt_render_wait with two frames in flight: frame k+1 is enqueued before frame k is waited
    for; every frame must equal the one rt_render delivers for the same camera."""
    scene, _, depth, _ = build_case("bobtex_d3")
    cams = [scenes.orbit_camera(k, width=400, height=300) for k in (0, 20, 40, 60, 80)]
    ctx = api.Context(devices=devices) if devices == "all" else api.Context(0)
    ctx.set_scene(scene)
    ctx.commit()
    want = [ctx.render(c, depth)[0].copy() for c in cams]
    bufs = [api.host_alloc(400 * 300 * 3), api.host_alloc(400 * 300 * 3)]
    try:
        ctx.render_enqueue(cams[0], depth, bufs[0], 0)
        for k in range(1, len(cams) + 1):
            if k < len(cams):
                ctx.render_enqueue(cams[k], depth, bufs[k % 2], k % 2)
            ctx.render_wait((k - 1) % 2)
            assert np.array_equal(bufs[(k - 1) % 2].reshape(300, 400, 3), want[k - 1]), k - 1
        # a pageable destination is page-locked on first use and gives the same frame
        plain = np.zeros(400 * 300 * 3, np.uint8)
        ctx.render_enqueue(cams[2], depth, plain, 0)
        ctx.render_wait(0)
        assert np.array_equal(plain.reshape(300, 400, 3), want[2])
    finally:
        ctx.close()
        for b in bufs:
            api.host_free(b)


def test_multi_device_refit_and_device_side_vertices():
    """REFIT commits and rt_scene_update_vertices[_device] reach every device of the context."""
    import torch
    scene, cam, depth, _ = build_case("bobtex_d3")
    moved = np.asarray(scene.tri_v, np.float32).copy()
    moved[:, 1::3] += 1.5
    scene2, _, _, _ = build_case("bobtex_d3")
    scene2.tri_v = moved
    want = one_gpu_frame(scene2, cam, depth)
    ctx = api.Context(devices="all")
    ctx.set_scene(scene)
    ctx.commit()
    ctx.update_vertices(moved)
    ctx.commit(api.COMMIT_REFIT, want_stats=False)
    got = ctx.render(cam, depth)
    assert np.array_equal(got[0], want[0])
    d = torch.from_numpy(np.ascontiguousarray(scene.tri_v, np.float32)).cuda(0)
    torch.cuda.synchronize()
    ctx.update_vertices_device(d.data_ptr(), len(scene.tri_v))          # back to the original mesh, from device memory
    ctx.commit(api.COMMIT_REFIT, want_stats=False)
    got = ctx.render(cam, depth)
    ctx.close()
    assert np.array_equal(got[0], one_gpu_frame(scene, cam, depth)[0])


def test_headless_lumina_uses_every_gpu_and_matches_one_gpu(tmp_path):
    """The lumina-compatible driver goes through World / RenderEngine::render(): with RT_DEVICES=1 and with all
    GPUs it must write the same image."""
    exe = os.path.join(ROOT, "realtrace_b200", "lumina_headless")
    outs = []
    for env_devices in ("1", "0"):                       # 0 = every visible device
        out = tmp_path / f"frame_{env_devices}.ppm"
        env = dict(os.environ, RT_DEVICES=env_devices)
        r = subprocess.run([exe, "640", "480", os.path.join(ASSETS, "bob_tri.obj"), str(out)], capture_output=True, text=True,
                           timeout=180, env=env)
        assert r.returncode == 0, r.stderr
        outs.append(out.read_bytes())
    assert outs[0] == outs[1]


def test_microbench_reports_plausible_peaks():
    ctx = api.Context(0)
    m = ctx.microbench()
    ctx.close()
    assert 2000 < m["hbm_read_gbs"] < 9000, m
    assert m["l2_read_gbs"] > m["hbm_read_gbs"], m
    assert 1000 < m["implied_sm_mhz"] < 2200, m
    assert m["l2_dependent_fetch_ns"] > 50, m
