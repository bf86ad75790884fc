"""Height-field terrain for the Elevation task.

The reference loads ``Terrains/huge_compact.usd`` (a 726 900-triangle mesh: a slab with top at z = 0.2 over
+-20.5 m carrying ramps/boxes up to z = 2.0, slopes mostly 15-25 deg) and collides /
ray-casts against the triangles.  The B200 path works on a regular raster of the TOP surface (what a vertical
ray from z+20 hits): ``procedural_heightfield`` synthesises a terrain with the same envelope (BASELINE config 3
says "procedural terrain"); a raster of the real mesh can be passed instead (same array contract).

Array contract: float32 [ny, pitch] with pitch = nx rounded up to a multiple of 4 (TMA needs 16-byte row strides);
sample (ix, iy) is the height at (x0 + ix*cell, y0 + iy*cell).
"""
from __future__ import annotations

import math

import numpy as np


def pad_pitch(h: np.ndarray) -> np.ndarray:
    ny, nx = h.shape
    pitch = (nx + 3) & ~3
    out = np.zeros((ny, pitch), dtype=np.float32)
    out[:, :nx] = h
    if pitch > nx:
        out[:, nx:] = h[:, -1:]
    return out


def procedural_heightfield(seed: int = 0, half_extent: float = 20.5, cell: float = 0.1, base: float = 0.2, z_max: float = 2.0,
                           n_ramps: int = 28, n_boxes: int = 24):
    """Returns (heights [ny, nx] float32, x0, y0, cell)."""
    n = int(round(2 * half_extent / cell)) + 1
    xs = -half_extent + cell * np.arange(n, dtype=np.float64)
    X, Y = np.meshgrid(xs, xs, indexing="xy")            # X[iy, ix]
    h = np.full((n, n), base, dtype=np.float64)
    rng = np.random.default_rng(seed + 0x7E44A1)
    for _ in range(n_ramps):                              # wedge: rises along `dir` with a 15-25 deg slope, flat top, falls back
        cx, cy = rng.uniform(-17, 17, 2)
        ang = rng.uniform(0, 2 * math.pi)
        slope = math.tan(math.radians(rng.uniform(15, 25)))
        top = rng.uniform(0.3, z_max - base)
        width = rng.uniform(1.0, 3.0)
        plateau = rng.uniform(0.5, 2.5)
        u = (X - cx) * math.cos(ang) + (Y - cy) * math.sin(ang)          # along
        v = -(X - cx) * math.sin(ang) + (Y - cy) * math.cos(ang)         # across
        run = top / slope
        prof = np.clip(np.minimum(u + run + plateau / 2, -(u - run - plateau / 2)) * slope, 0.0, top)
        side = np.clip((width / 2 - np.abs(v)) * 2.0, 0.0, 1.0)          # 0.5 m soft shoulders
        h = np.maximum(h, base + prof * side)
    for _ in range(n_boxes):                              # low platforms with 45 deg chamfers
        cx, cy = rng.uniform(-18, 18, 2)
        sx, sy = rng.uniform(0.6, 2.5, 2)
        top = rng.uniform(0.05, 0.4)
        d = np.minimum(sx / 2 - np.abs(X - cx), sy / 2 - np.abs(Y - cy))
        h = np.maximum(h, base + np.clip(d + top, 0.0, top))
    h = np.clip(h, 0.0, z_max)
    return h.astype(np.float32), -half_extent, -half_extent, cell
