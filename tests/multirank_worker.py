"""Worker for tests/test_multirank_cpu.py: world_size ranks on CPU (gloo) run the N > 1 host path —
interleaved tile ownership, packing, the gather collective, assembly — on a synthetic frame."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from realtrace_b200 import api, multigpu  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    W, H = int(sys.argv[1]), int(sys.argv[2])
    yy, xx = np.mgrid[0:H, 0:W]
    frame = np.stack([(xx * 7 + yy * 3) % 256, (xx ^ yy) % 256, (xx * yy) % 251], axis=-1).astype(np.uint8)
    mine = multigpu.pack_tiles_host(frame, rank, world)
    # the C ABI's layout function must agree with the Python mirror
    total, owned, tile_bytes = api.tile_layout(W, H, 0, 0, rank, world)
    assert owned == len(multigpu.owned_tiles(W, H, rank, world)) and owned * tile_bytes == mine.size
    padded = torch.zeros(multigpu.max_owned(W, H, world) * tile_bytes, dtype=torch.uint8)
    padded[:mine.size] = torch.from_numpy(mine)
    got = multigpu.gather_packed(padded, rank, world)
    ok = 1
    if rank == 0:
        out = multigpu.assemble_host([g.numpy() for g in got], W, H)
        ok = int(np.array_equal(out, frame))
    flag = torch.tensor([ok])
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTIRANK_OK" if ok else "MULTIRANK_MISMATCH", flush=True)
    sys.exit(0 if int(flag[0]) else 1)


if __name__ == "__main__":
    main()
