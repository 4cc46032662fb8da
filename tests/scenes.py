"""Seeded synthetic scenes shared by the tests, smoke() and bench.py."""
import numpy as np


def lattice(nx, ny, nz, origin=(1.5, 1.5, 1.5), spacing=1.0, jitter=0.05, seed=12345):
    """nx*ny*nz block, x-major like init_chunk_from_grid (reference src/Lustrine.cpp:694-719):
    cell centre = index*spacing + origin, plus uniform jitter so no pair sits on a lattice distance."""
    gx = np.arange(nx, dtype=np.float32) * np.float32(spacing) + np.float32(origin[0])
    gy = np.arange(ny, dtype=np.float32) * np.float32(spacing) + np.float32(origin[1])
    gz = np.arange(nz, dtype=np.float32) * np.float32(spacing) + np.float32(origin[2])
    pos = np.stack(np.meshgrid(gx, gy, gz, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    if jitter > 0:
        rng = np.random.default_rng(seed)
        pos += rng.uniform(-jitter, jitter, pos.shape).astype(np.float32)
    return np.ascontiguousarray(pos)


def dam_break(n_side, jitter=0.05, seed=12345):
    """N^3 block at (1,1,1)+0.5 in a (3N, 2N, 2N) domain."""
    return (3 * n_side, 2 * n_side, 2 * n_side), lattice(n_side, n_side, n_side, jitter=jitter, seed=seed)


def floor_plate(nx, nz, y=0.5, origin=(0.5, 0.5)):
    gx = np.arange(nx, dtype=np.float32) + np.float32(origin[0])
    gz = np.arange(nz, dtype=np.float32) + np.float32(origin[1])
    return np.ascontiguousarray(np.stack(np.meshgrid(gx, np.array([y], np.float32), gz, indexing="ij"), -1).reshape(-1, 3).astype(np.float32))


def sand_pile(n_side, drop=3.0, jitter=0.05, seed=777):
    """N^3 sand block dropped onto a solid floor plate with two obstacle boxes (scaled)."""
    D = (3 * n_side, 2 * n_side + 8, 3 * n_side)
    sand = lattice(n_side, n_side, n_side, origin=(n_side + 0.5, drop + 0.5, n_side + 0.5), jitter=jitter, seed=seed)
    floor = floor_plate(3 * n_side, 3 * n_side, y=0.5)
    box = lattice(max(n_side // 3, 1), 2, max(n_side // 3, 1), origin=(n_side + 1.5, 1.5, n_side + 1.5), jitter=0.0)
    return D, sand, np.concatenate([floor, box]).astype(np.float32)
