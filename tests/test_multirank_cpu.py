"""N > 1 host path on CPU: world_size-2 and -3 gloo runs of the tile partition + gather + assembly,
and the reference arm of bench.py under torchrun (rank 0 works, the others exit 0)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(nproc, script_args, port, timeout=300):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port)] + script_args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


@pytest.mark.parametrize("world,size", [(2, (200, 150)), (3, (640, 480))])
def test_tile_gather_assembly_gloo(world, size):
    r = _torchrun(world, [os.path.join(ROOT, "tests", "multirank_worker.py"), str(size[0]), str(size[1])], 29611 + world)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "MULTIRANK_OK" in r.stdout


def test_host_pack_assemble_roundtrip_all_world_sizes():
    from realtrace_b200 import multigpu
    rng = np.random.default_rng(1)
    frame = rng.integers(0, 256, (135, 250, 3), dtype=np.uint8)
    for world in (1, 2, 4, 8, 5):
        parts = [multigpu.pack_tiles_host(frame, r, world) for r in range(world)]
        assert sum(len(multigpu.owned_tiles(250, 135, r, world)) for r in range(world)) == 4 * 5
        assert np.array_equal(multigpu.assemble_host(parts, 250, 135), frame)


def test_reference_arm_under_torchrun_prints_one_line():
    r = _torchrun(2, [os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                      "--workload", "analytic"], 29631)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mrays/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["e2e"]["h2d_bytes_per_step"] == 0
