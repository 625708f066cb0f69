"""CPU-side checks of the C++ host layer (realtrace_b200/host): what does not need a GPU — the OBJ/PNG
loader against the Python harness loader (the flat arrays both feed to the GPU must be identical), the PNG
writer, the orbit camera against Parellel/interactive_camera.cu's formula, camera rays against the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSETS = os.path.join(ROOT, "assets")


@pytest.fixture(scope="module")
def host():
    import __graft_entry__ as entry
    entry.build()
    lib = C.CDLL(os.path.join(ROOT, "realtrace_b200", "librealtrace_host.so"))
    lib.rt_host_load_obj.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.c_int]
    lib.rt_host_save_png.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
    lib.rt_host_orbit_eye.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p]
    lib.rt_host_camera_ray.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    return lib


def _load(host, obj, tex, max_faces, cap):
    v = np.zeros((cap, 9), np.float32)
    rgb = np.zeros((cap, 9), np.float32)
    err = C.create_string_buffer(256)
    n = host.rt_host_load_obj(os.path.join(ASSETS, obj).encode(), os.path.join(ASSETS, tex).encode() if tex else None,
                              max_faces, v.ctypes.data, rgb.ctypes.data, cap, err, 256)
    assert n >= 0, err.value.decode()
    return n, v[:n], rgb[:n]


@pytest.mark.parametrize("obj,cap", [("tetrahedron.obj", -1), ("bob_tri.obj", -1), ("bob_tri.obj", 2000),
                                     ("blub_triangulated.obj", -1)])
def test_cpp_obj_loader_equals_python_loader(host, obj, cap):
    from realtrace_b200 import objio
    tri, _, _, _ = objio.triangles_from_obj(os.path.join(ASSETS, obj), None if cap < 0 else cap)
    n, v, _ = _load(host, obj, None, cap, 20000)
    assert n == len(tri)
    assert np.array_equal(v, tri)


def test_cpp_png_decoder_and_texel_fetch_equal_python(host):
    from realtrace_b200 import scenes
    s = scenes.bob_textured()
    n, v, rgb = _load(host, "bob_tri.obj", "bob_diffuse.png", -1, 20000)
    assert n == len(s.tri_v)
    assert np.array_equal(v, s.tri_v)
    assert np.array_equal(rgb, s.tri_rgb)


def test_png_writer_roundtrip(host, tmp_path):
    from PIL import Image
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)      # bottom-up bitmap like Camera::getBitmap()
    path = str(tmp_path / "frame.png")
    assert host.rt_host_save_png(path.encode(), img.ctypes.data, 53, 37) == 0
    back = np.asarray(Image.open(path).convert("RGB"))
    assert np.array_equal(back, img[::-1])                       # row 0 of the bitmap is the bottom row


def test_orbit_camera_matches_the_reference_formula(host):
    from realtrace_b200.scene import orbit_camera
    out = np.zeros(3, np.float32)
    for k in (0, 1, 17, 60, 119):
        yaw = np.float32(2.0 * np.pi * k / 120)
        host.rt_host_orbit_eye(float(yaw), 0.3, 84.85281374238570, out.ctypes.data)
        cam = orbit_camera(k)
        assert np.allclose(out, np.asarray(cam.pos, np.float32), rtol=2e-6, atol=2e-5), k


def test_camera_rays_equal_the_oracle_camera(host, port_oracle):
    """Camera::get_ray_direction of the mirror class vs the oracle's (bit-identical FP64)."""
    from realtrace_b200 import scenes
    from kat import kat_scene
    pos, tgt, up = (np.array(x, np.float64) for x in ((60, 60, 0), (0, 0, 0), (0, 1, 0)))
    out = np.zeros(3, np.float64)
    # the oracle exposes primary rays only through rendering; compare against the Python restatement used
    # for the GPU camera (scene.Camera.basis) evaluated like camera.cpp:33-44
    cam = scenes.stock_camera(640, 480)
    u, v, w, focal, aspect = cam.basis()
    for (i, j) in ((0, 0), (319, 240), (639, 479), (123, 45)):
        host.rt_host_camera_ray(pos.ctypes.data, tgt.ctypes.data, up.ctypes.data, 45.0, 640, 480, i, j, out.ctypes.data)
        xw = np.float32(float(aspect) * (i - 640 / 2.0 + 0.5) / 640)
        yw = np.float32((j - 480 / 2.0 + 0.5) / 480)
        d = -w * float(focal) + u * float(xw) + v * float(yw)
        d = d / np.sqrt((d * d).sum())
        assert np.array_equal(out, d), (i, j)
