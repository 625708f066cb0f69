"""One frame of every shipped kernel inside a cudaProfilerStart/Stop bracket, for the ncu captures under profiles/
(ncu --profile-from-start off).  usage: gpu_profile_r2.py <synth1m|blub4k|bob1080|build>
  synth1m  k_frame (one launch per frame) and k_frame_push (a rank's share of an 8-GPU frame, frame on this GPU)
  blub4k   k_traverse<primary, fused>, k_shade, k_paths<wide>, k_resolve, k_assemble16 (packed tiles + scatter)
  bob1080  the same kernels on the textured mirror scene
  build    LBVH build + refit + 4-wide collapse of the 1 M-triangle scene"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from realtrace_b200 import api, scenes

what = sys.argv[1] if len(sys.argv) > 1 else "synth1m"
name = "synth1m" if what == "build" else what
scene, cam, depth, _ = scenes.workload(name)
if what == "build":
    os.environ["RT_WIDE_BVH"] = "2"          # so that k_collapse4 runs on the bounce-free scene too
ctx = api.Context(0)
ctx.set_scene(scene)
ctx.commit()
W, H = cam.width, cam.height
frame = torch.zeros(W * H * 3, dtype=torch.uint8, device="cuda")
prof = torch.cuda.profiler
if what == "build":
    ctx.commit()
    torch.cuda.synchronize()
    prof.start()
    print("build", ctx.commit())
    print("refit", ctx.commit(api.COMMIT_REFIT))
    torch.cuda.synchronize()
    prof.stop()
else:
    for _ in range(4):
        st = ctx.render_device(cam, depth, frame.data_ptr())
    torch.cuda.synchronize()
    prof.start()
    st = ctx.render_device(cam, depth, frame.data_ptr())
    torch.cuda.synchronize()
    prof.stop()
    print(what, "frame", st["ms_device"], "rays", st["rays_primary"] + st["rays_shadow"] + st["rays_secondary"])
    # a rank's share of an 8-GPU frame through rt_render_push (rank 1; rank 0's "frame open" is published first)
    world, rank, tile = 8, 1, (32, 16)
    _, owned, tb = api.tile_layout(W, H, tile[0], tile[1], rank, world)
    packed = torch.zeros(max(owned * tb, 1), dtype=torch.uint8, device="cuda")
    sync_ptr, _ = ctx.shared_buffer_create(1024)
    cs = api.camera_struct(cam)
    params = api.Context._params(depth, tile=tile, rank=rank, world=world, flags=api.FLAG_PACKED_TILES)
    for k in range(4):
        ctx.peer_sync(sync_ptr, 0, world, k, 0)
        ctx.render_push(cs, params, packed.data_ptr(), frame.data_ptr(), sync_ptr, k)
    ctx.synchronize()
    ctx.peer_sync(sync_ptr, 0, world, 4, 0)
    ctx.synchronize()
    prof.start()
    ctx.render_push(cs, params, packed.data_ptr(), frame.data_ptr(), sync_ptr, 4)
    ctx.synchronize()
    prof.stop()
ctx.close()
