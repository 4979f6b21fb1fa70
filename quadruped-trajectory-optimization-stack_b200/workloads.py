"""Synthetic workloads named in BASELINE.json / SURVEY 8(d): batched multi-start windows on rough
heightfields (config 4) and the terrain-variant replan sweep (config 5).  Host numpy only."""
import numpy as np

from . import PROBLEM_DTYPE
from .heightfield import get_height, rough_terrain

FEET_XY = ((0.21, 0.19), (0.21, -0.19), (-0.21, 0.19), (-0.21, -0.19))   # ref: QTOS/combiner.py:146-152


def multistart_problems(n, grid, res, seed=1234, hf_id=0, group_size=1):
    """n start/goal pairs: start x,y ~ U(0, 2.5), z = h + 0.24, yaw 0, feet at nominal stance on the
    terrain; goal = start + (U(0.2, 0.6), U(-0.1, 0.1)).  `group_size` consecutive problems share a
    group id (candidates that compete in best-plan selection)."""
    rng = np.random.default_rng(seed)
    p = np.zeros(n, dtype=PROBLEM_DTYPE)
    sx, sy = rng.uniform(0, 2.5, n), rng.uniform(0, 2.5, n)
    gx, gy = sx + rng.uniform(0.2, 0.6, n), sy + rng.uniform(-0.1, 0.1, n)
    p["start_pos"][:, 0], p["start_pos"][:, 1] = sx, sy
    p["start_pos"][:, 2] = get_height(grid, res, sx, sy) + 0.24
    p["goal"][:, 0], p["goal"][:, 1], p["goal"][:, 2] = gx, gy, 0.24
    for e, (a, b) in enumerate(FEET_XY):
        p["ee"][:, e, 0], p["ee"][:, e, 1] = sx + a, sy + b
        p["ee"][:, e, 2] = get_height(grid, res, sx + a, sy + b)
    p["hf_id"] = hf_id
    p["group"] = np.arange(n) // max(1, group_size)
    return p


def terrain_variants(n_variants=8):
    """seeds 0..n-1 of the config-4 generator."""
    return [rough_terrain(seed=s) for s in range(n_variants)]
