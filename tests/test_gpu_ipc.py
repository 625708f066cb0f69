"""Frame assembly through peer memory: a second PROCESS maps rank 0's frame (CUDA IPC — across GPUs
this is the NVLink path bench.py uses at N > 1) and its resolve kernel writes its tiles into it."""
import os
import subprocess
import sys

import numpy as np
import pytest

from cases import build_case
from realtrace_b200 import api

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_processes_assemble_one_frame_through_ipc():
    scene, cam, depth, _ = build_case("blubmixed_d5")
    cam.width, cam.height = 200, 150
    ctx = api.Context(0)
    ctx.set_scene(scene)
    ctx.commit()
    full = ctx.render(cam, depth)[0]
    ptr, handle = ctx.shared_buffer_create(cam.width * cam.height * 3)
    st = ctx.render_device(cam, depth, ptr, rank=0, world=2)       # this process: the even tiles
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ipc_child.py"), handle.hex(), "1", "2",
                        str(cam.width), str(cam.height)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "IPC_CHILD_OK" in r.stdout, r.stdout + r.stderr
    out = np.zeros((cam.height, cam.width, 3), np.uint8)
    ctx.download(ptr, out)
    ctx.close()
    assert st["rays_primary"] < cam.width * cam.height
    assert np.array_equal(out, full)


def test_two_processes_steal_from_one_cursor():
    """The stealing cursor is mapped by a second process too (system-scope atomics on IPC memory)."""
    scene, cam, depth, _ = build_case("blubmixed_d5")
    cam.width, cam.height = 456, 270
    ctx = api.Context(0)
    ctx.set_scene(scene)
    ctx.commit()
    full = ctx.render(cam, depth)[0]
    ptr, handle = ctx.shared_buffer_create(cam.width * cam.height * 3)
    cursor, chandle = ctx.shared_buffer_create(256)
    # the child (rank 1) goes first and takes the whole pool; the parent then only finds its own tiles
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ipc_child.py"), handle.hex(), "1", "2",
                        str(cam.width), str(cam.height), chandle.hex(), "2", "5"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "IPC_CHILD_OK" in r.stdout, r.stdout + r.stderr
    st = ctx.render_device(cam, depth, ptr, rank=0, world=2, steal=(2, 5, cursor))
    out = np.zeros((cam.height, cam.width, 3), np.uint8)
    ctx.download(ptr, out)
    ctx.close()
    assert st["stolen_blocks"] == 0 and st["rays_primary"] < cam.width * cam.height // 2
    assert np.array_equal(out, full)


def test_async_frames_then_synchronize():
    scene, cam, depth, _ = build_case("bobtex_d3")
    ctx = api.Context(0)
    ctx.set_scene(scene)
    ctx.commit()
    ref = ctx.render(cam, depth)[0]
    ptr, _ = ctx.shared_buffer_create(cam.width * cam.height * 3)
    for _ in range(5):
        ctx.render_device(cam, depth, ptr, want_stats=False)
    ctx.synchronize()
    out = np.zeros_like(ref)
    ctx.download(ptr, out)
    ctx.close()
    assert np.array_equal(out, ref)
