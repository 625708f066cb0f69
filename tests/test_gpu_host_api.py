"""The C++ mirror of the reference's classes (realtrace_b200/host) on the GPU: scenes built through
World/Triangle/Sphere/.../RenderEngine exactly like Serial/lumina.cpp:302-370 must give the same
frames as the flat-array path, and therefore the same parity against the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import parity
from oracle import binding as ob
from realtrace_b200 import api, scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSETS = os.path.join(ROOT, "assets")


@pytest.fixture(scope="module")
def host():
    import __graft_entry__ as entry
    entry.build()
    lib = C.CDLL(os.path.join(ROOT, "realtrace_b200", "librealtrace_host.so"))
    lib.rt_host_demo.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_char_p, C.c_int]
    lib.rt_host_demo.restype = C.c_int
    return lib


def demo(lib, which, w, h, depth):
    rgb = np.zeros((h, w, 3), np.uint8)
    rays = np.zeros(3, np.uint64)
    ms = C.c_float()
    err = C.create_string_buffer(512)
    rc = lib.rt_host_demo(which.encode(), ASSETS.encode(), w, h, depth, rgb.ctypes.data, rays.ctypes.data,
                          C.byref(ms), err, 512)
    assert rc == 0, (rc, err.value.decode())
    return rgb, rays, ms.value


def flat(scene, cam, depth):
    ctx = api.Context(0)
    ctx.set_scene(scene)
    ctx.commit()
    out = ctx.render(cam, depth)
    ctx.close()
    return out


@pytest.mark.parametrize("which,builder,cam,depth", [
    ("lumina_default", lambda: scenes.obj_scene("bob_tri.obj", 2000), lambda: scenes.stock_camera(320, 240), 10),
    ("analytic_close", lambda: scenes.analytic_scene(), lambda: scenes.close_camera(320, 240), 5),
    ("analytic", lambda: scenes.analytic_scene(), lambda: scenes.stock_camera(320, 240), 1),
    ("bob_textured", lambda: scenes.bob_textured(), lambda: scenes.stock_camera(320, 240), 3),
])
def test_class_api_frames_equal_flat_api_frames(host, which, builder, cam, depth):
    rgb, rays, ms = demo(host, which, 320, 240, depth)
    ref_rgb, _, _, st = flat(builder(), cam(), depth)
    assert np.array_equal(rgb, ref_rgb)
    assert (int(rays[0]), int(rays[1]), int(rays[2])) == (st["rays_primary"], st["rays_shadow"], st["rays_secondary"])
    assert ms > 0


def test_class_api_matches_the_oracle(host):
    import __graft_entry__ as entry
    entry.build_oracle()
    rgb, _, _ = demo(host, "lumina_default", 320, 240, 10)
    tr = ob.load_port().render(scenes.obj_scene("bob_tri.obj", 2000), scenes.stock_camera(320, 240), 10,
                               ob.MODE_AS_SHIPPED, aux=False)
    diff = np.abs(rgb.astype(int) - tr[0].astype(int)).max(axis=-1)
    assert (diff <= 1).mean() >= parity.COLOUR_MATCH_MIN


def test_headless_lumina_binary(host, tmp_path):
    out = tmp_path / "frame.ppm"
    exe = os.path.join(ROOT, "realtrace_b200", "lumina_headless")
    r = subprocess.run([exe, "161", "121", os.path.join(ASSETS, "bob_tri.obj"), str(out)], capture_output=True, text=True,
                       timeout=120)
    assert r.returncode == 0, r.stderr
    data = out.read_bytes()
    header = b"P6\n160 120\n255\n"            # 161x121 rounded down to even like lumina.cpp:484-485
    assert data.startswith(header)
    img = np.frombuffer(data[len(header):], np.uint8).reshape(120, 160, 3)[::-1]
    ref = flat(scenes.obj_scene("bob_tri.obj", 2000), scenes.stock_camera(160, 120), 10)[0]     # the :266 cap is the default
    assert np.array_equal(img, ref)
