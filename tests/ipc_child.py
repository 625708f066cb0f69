"""Child process of tests/test_gpu_ipc.py: maps the parent's shared frame (CUDA IPC) and renders the
tiles of one rank straight into it."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import build_case  # noqa: E402
from realtrace_b200 import api  # noqa: E402


def main():
    handle = bytes.fromhex(sys.argv[1])
    rank, world, w, h = (int(x) for x in sys.argv[2:6])
    scene, cam, depth, _ = build_case("blubmixed_d5")
    cam.width, cam.height = w, h
    ctx = api.Context(0)
    ctx.set_scene(scene)
    ctx.commit()
    ptr = ctx.shared_buffer_open(handle)
    steal = None
    if len(sys.argv) > 6:
        cursor = ctx.shared_buffer_open(bytes.fromhex(sys.argv[6]))
        steal = (int(sys.argv[7]), int(sys.argv[8]), cursor)
    ctx.render_device(cam, depth, ptr, rank=rank, world=world, want_stats=False, steal=steal)   # asynchronous frame
    ctx.synchronize()
    ctx.close()
    print("IPC_CHILD_OK", flush=True)


if __name__ == "__main__":
    main()
