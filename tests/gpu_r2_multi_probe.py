"""Development probe (round 2): where does a multi-device frame's time go?  One context over all GPUs."""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(name):
    import numpy as np
    from realtrace_b200 import api, scenes
    scene, cam, depth, _ = scenes.workload(name)
    W, H = cam.width, cam.height
    ctx = api.Context(devices="all")
    ctx.set_scene(scene)
    ctx.commit()
    n = ctx.device_count()
    bufs = [api.host_alloc(W * H * 3), api.host_alloc(W * H * 3)]
    cs = api.camera_struct(cam)
    out = {"workload": name, "devices": n, "env": {k: v for k, v in os.environ.items() if k.startswith("RT_")}}
    for _ in range(4):
        st = ctx.render(cs, depth, out=bufs[0].reshape(H, W, 3))[3]
    out["ms_device_stats"] = round(st["ms_device"], 4)
    # A: frames into a device frame on rank 0, asynchronous
    import torch
    dev_frame = torch.zeros(W * H * 3, dtype=torch.uint8, device="cuda:0")
    for _ in range(3):
        ctx.render_device(cs, depth, dev_frame.data_ptr(), want_stats=False)
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        ctx.render_device(cs, depth, dev_frame.data_ptr(), want_stats=False)
    ctx.synchronize()
    out["A_device_frame_async_ms"] = round((time.perf_counter() - t0) / 20 * 1e3, 4)
    # B: synchronous host frames
    lat = []
    for _ in range(8):
        t0 = time.perf_counter()
        ctx.render(cs, depth, out=bufs[0].reshape(H, W, 3))
        lat.append((time.perf_counter() - t0) * 1e3)
    out["B_sync_render_ms"] = round(float(np.median(lat)), 4)
    # C: pipelined
    ctx.render_enqueue(cs, depth, bufs[0], 0)
    ctx.render_wait(0)
    t0 = time.perf_counter()
    ctx.render_enqueue(cs, depth, bufs[0], 0)
    steps = 20
    t_enq, t_wait = 0.0, 0.0
    enq_list, wait_list = [], []
    for k in range(1, steps + 1):
        a = time.perf_counter()
        if k < steps:
            ctx.render_enqueue(cs, depth, bufs[k % 2], k % 2)
        b = time.perf_counter()
        ctx.render_wait((k - 1) % 2)
        c = time.perf_counter()
        t_enq += b - a
        t_wait += c - b
        enq_list.append(round((b - a) * 1e3, 3))
        wait_list.append(round((c - b) * 1e3, 3))
    out["C_pipelined_ms"] = round((time.perf_counter() - t0) / steps * 1e3, 4)
    out["C_host_enqueue_ms"] = round(t_enq / steps * 1e3, 4)
    out["C_host_wait_ms"] = round(t_wait / steps * 1e3, 4)
    out["C_enq_each"] = enq_list
    out["C_wait_each"] = wait_list
    print(json.dumps(out), flush=True)
    ctx.close()


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2])
    else:
        names = os.environ.get("PROBE_WORKLOADS", "synth1m,blub4k").split(",")
        variants = ({}, {"RT_NO_VMM": "1"}, {"RT_FRAME_KERNEL": "0"}, {"RT_FRAME_KERNEL": "0", "RT_NO_VMM": "1"})
        if os.environ.get("PROBE_ENVS"):
            variants = [json.loads(a) for a in os.environ["PROBE_ENVS"].split(";")]
        for name in names:
            for v in variants:
                env = dict(os.environ)
                env.update(v)
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", name], env=env, capture_output=True,
                                   text=True, timeout=300)
                print(r.stdout.strip() or f"FAILED {name} {v}: {r.stderr[-1500:]}", flush=True)
