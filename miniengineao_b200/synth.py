"""Deterministic synthetic depth buffers.  Shared by the parity tests (fed to both the
oracle and the CUDA path) and by bench.py; pure numpy, no files.

Common camera: near 0.3, far 100, fovY 60 deg, reversed-Z raw depth
    d = (1/lin01 - 1) / (far/near - 1)      (inverse of Downsample1.compute:40 with AO.cs:565).
"""
from __future__ import annotations

import numpy as np

NEAR, FAR, FOV_Y_DEG = 0.3, 100.0, 60.0


def tan_half_fov(width: int, height: int, fov_y_deg: float = FOV_Y_DEG) -> tuple[float, float]:
    tv = np.tan(np.radians(fov_y_deg) / 2.0)
    return float(width / height * tv), float(tv)


def lin01_to_raw(lin01: np.ndarray, near: float = NEAR, far: float = FAR, reversed_z: bool = True) -> np.ndarray:
    """Raw camera depth whose Linearize() is lin01 (up to rounding)."""
    lin01 = lin01.astype(np.float64)
    fpn = far / near
    if reversed_z:
        d = (1.0 / lin01 - 1.0) / (fpn - 1.0)
    else:
        d = (1.0 / lin01 - fpn) / (1.0 - fpn)
    return d.astype(np.float32)


def _pcg_hash(idx: np.ndarray, seed: int) -> np.ndarray:
    """pcg32-style output hash of (index, seed) -> uint32."""
    with np.errstate(over="ignore"):
        state = (idx.astype(np.uint64) + np.uint64(seed)) * np.uint64(6364136223846793005) + np.uint64(1442695040888963407)
        word = ((state >> ((state >> np.uint64(59)) + np.uint64(5))) ^ state) * np.uint64(12605985483714917081)
        out = (word >> np.uint64(43)) ^ word
    return (out & np.uint64(0xFFFFFFFF)).astype(np.uint32)


def _pixel_rays(width: int, height: int, row0: int = 0, row1: int | None = None):
    row1 = height if row1 is None else row1
    th, tv = tan_half_fov(width, height)
    xs = ((np.arange(width, dtype=np.float64) + 0.5) / width * 2.0 - 1.0) * th
    ys = (1.0 - (np.arange(row0, row1, dtype=np.float64) + 0.5) / height * 2.0) * tv
    return np.meshgrid(xs, ys)      # dx, dy per unit z


def flat_sphere(width: int = 256, height: int = 256, *, plane_z: float = 10.0, sphere=(0.3, 0.0, 8.0, 2.0),
                far: float = FAR) -> np.ndarray:
    """lin01 of a camera-facing plane at z=plane_z with a sphere in front (modelled on Spheres.unity)."""
    dx, dy = _pixel_rays(width, height)
    cx, cy, cz, r = sphere
    # ray p = t * (dx, dy, 1); |p - c|^2 = r^2
    a = dx * dx + dy * dy + 1.0
    b = -2.0 * (dx * cx + dy * cy + cz)
    c = cx * cx + cy * cy + cz * cz - r * r
    disc = b * b - 4 * a * c
    t = np.where(disc > 0, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), np.inf)
    z = np.minimum(np.where(t > 0, t, np.inf), plane_z)
    return (z / far).astype(np.float32)


_corridor_cache: dict = {}


def corridor(width: int, height: int, *, frame: int = 0, seed: int = 0xA0, noise: float = 1e-4,
             row0: int = 0, row1: int | None = None, far: float = FAR) -> np.ndarray:
    """lin01 of the "Sponza-like" corridor: floor y=-2, ceiling y=+6, walls x=+-5, end wall z=60, two rows
    of vertical cylinders r=0.5 every 4 m at x=+-3.5; camera z-offset 0.25*frame; multiplicative hash
    noise so that no two neighbouring depths tie exactly.  Closed box: no sky pixels.
    Rows [row0,row1) only (bands of a large frame can be generated independently)."""
    row1 = height if row1 is None else row1
    zoff = 0.25 * frame
    key = (width, height, row0, row1)
    if _corridor_cache.get("key") != key:       # floor / ceiling / side walls do not depend on the frame index: kept for frame streams
        dx, dy = _pixel_rays(width, height, row0, row1)
        with np.errstate(divide="ignore", invalid="ignore"):
            planes = np.where(dy < 0, -2.0 / dy, np.inf)
            planes = np.minimum(planes, np.where(dy > 0, 6.0 / dy, np.inf))
            planes = np.minimum(planes, np.where(dx != 0, 5.0 / np.abs(dx), np.inf))
        ys, xs = np.meshgrid(np.arange(row0, row1, dtype=np.uint64), np.arange(width, dtype=np.uint64), indexing="ij")
        _corridor_cache.clear()
        _corridor_cache.update(key=key, dx1=dx[0].copy(), planes=planes, idx=xs + np.uint64(width) * ys)
    dx1, planes, idx = _corridor_cache["dx1"], _corridor_cache["planes"], _corridor_cache["idx"]
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.minimum(planes, 60.0 - zoff)     # end wall
        # the cylinders are vertical, so their ray parameter depends on the COLUMN only: evaluate on one row and broadcast
        # (the same float64 operations per element as the full-frame form, hence the same bits)
        a = dx1 * dx1 + 1.0
        tcyl = np.full(dx1.shape, np.inf)
        for cxs in (-3.5, 3.5):
            for cz in np.arange(4.0, 60.0, 4.0) - zoff:
                if cz <= 0.6:
                    continue
                b = -2.0 * (dx1 * cxs + cz)
                c = cxs * cxs + cz * cz - 0.25
                disc = b * b - 4 * a * c
                tc = np.where(disc > 0, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), np.inf)
                tcyl = np.minimum(tcyl, np.where(tc > 0, tc, np.inf))
        t = np.minimum(t, tcyl[None, :])
    h = _pcg_hash(idx, seed + frame * 7919)
    z = t * (1.0 + noise * (h.astype(np.float64) / 4294967296.0 - 0.5))
    return (z / far).astype(np.float32)


def random_depth(width: int, height: int, seed: int = 0, lo: float = 0.02, hi: float = 0.9) -> np.ndarray:
    """Smooth-ish random lin01 field with discontinuities (stress input for parity tests)."""
    rng = np.random.default_rng(seed)
    base = rng.uniform(lo, hi, size=((height + 15) // 16 + 2, (width + 15) // 16 + 2))
    yy = (np.arange(height) / 16.0)[:, None]
    xx = (np.arange(width) / 16.0)[None, :]
    y0, x0 = np.floor(yy).astype(int), np.floor(xx).astype(int)
    fy, fx = yy - y0, xx - x0
    v = (base[y0, x0] * (1 - fy) * (1 - fx) + base[y0 + 1, x0] * fy * (1 - fx)
         + base[y0, x0 + 1] * (1 - fy) * fx + base[y0 + 1, x0 + 1] * fy * fx)
    blocks = rng.uniform(lo, hi, size=(height // 37 + 1, width // 53 + 1))
    mask = rng.uniform(size=blocks.shape) < 0.3
    by, bx = np.arange(height)[:, None] // 37, np.arange(width)[None, :] // 53
    v = np.where(mask[by, bx], blocks[by, bx], v)
    v = v * (1.0 + 1e-3 * (rng.uniform(size=v.shape) - 0.5))
    return np.clip(v, lo / 2, 0.999).astype(np.float32)
