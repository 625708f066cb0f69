"""Tile sharding across ranks (one process per GPU) — the host-side logic of SURVEY §8(e).

The scene is replicated; rank r renders the tiles with tile_id % world == r (interleaved), writes
them back to back into a packed RGB8 buffer (tile-local row-major), and one NCCL gather brings the
packed buffers to rank 0, which scatters them into the frame (rt_assemble_tiles on the GPU;
assemble_host below is the same mapping in numpy, used by the CPU/gloo tests and as a cross-check).
Every pixel is produced by exactly one rank with identical code and data, so the assembled frame is
bit-identical for any world size.
"""
from __future__ import annotations

import numpy as np

TILE_W, TILE_H = 64, 32


def tile_grid(width, height, tile_w=TILE_W, tile_h=TILE_H):
    return (width + tile_w - 1) // tile_w, (height + tile_h - 1) // tile_h


def owned_tiles(width, height, rank, world, tile_w=TILE_W, tile_h=TILE_H):
    tx, ty = tile_grid(width, height, tile_w, tile_h)
    return list(range(rank, tx * ty, world))


def max_owned(width, height, world, tile_w=TILE_W, tile_h=TILE_H):
    tx, ty = tile_grid(width, height, tile_w, tile_h)
    return (tx * ty + world - 1) // world


def pack_tiles_host(frame, rank, world, tile_w=TILE_W, tile_h=TILE_H):
    """frame (H, W, 3) uint8 -> this rank's packed tiles, flat uint8 (pixels outside the frame are 0)."""
    H, W, _ = frame.shape
    tx, _ = tile_grid(W, H, tile_w, tile_h)
    ids = owned_tiles(W, H, rank, world, tile_w, tile_h)
    out = np.zeros((len(ids), tile_h, tile_w, 3), np.uint8)
    for k, t in enumerate(ids):
        x0, y0 = (t % tx) * tile_w, (t // tx) * tile_h
        h, w = min(tile_h, H - y0), min(tile_w, W - x0)
        out[k, :h, :w] = frame[y0:y0 + h, x0:x0 + w]
    return out.reshape(-1)


def assemble_host(packed_per_rank, width, height, tile_w=TILE_W, tile_h=TILE_H):
    """list of packed buffers (index = rank) -> frame (H, W, 3)."""
    world = len(packed_per_rank)
    tx, _ = tile_grid(width, height, tile_w, tile_h)
    frame = np.zeros((height, width, 3), np.uint8)
    for rank, packed in enumerate(packed_per_rank):
        ids = owned_tiles(width, height, rank, world, tile_w, tile_h)
        tiles = np.asarray(packed, np.uint8)[:len(ids) * tile_h * tile_w * 3].reshape(len(ids), tile_h, tile_w, 3)
        for k, t in enumerate(ids):
            x0, y0 = (t % tx) * tile_w, (t // tx) * tile_h
            h, w = min(tile_h, height - y0), min(tile_w, width - x0)
            frame[y0:y0 + h, x0:x0 + w] = tiles[k, :h, :w]
    return frame


def gather_packed(packed, rank, world, gather_list=None):
    """The frame-assembly collective: gather equally sized packed buffers on rank 0.

    packed: a torch uint8 tensor (CUDA with the nccl backend, CPU with gloo) padded to
    max_owned * tile_bytes.  Returns the list of per-rank buffers on rank 0, None elsewhere."""
    import torch.distributed as dist
    if rank == 0 and gather_list is None:
        import torch
        gather_list = [torch.zeros_like(packed) for _ in range(world)]
    dist.gather(packed, gather_list if rank == 0 else None, dst=0)
    return gather_list if rank == 0 else None
