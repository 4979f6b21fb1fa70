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


def replan_from_rows(prev_problems, row, step=(0.4, 0.0)):
    """vectorised replan_problems: `row` [n, 37] is every plan's CSV row at the hand-over time."""
    p = prev_problems.copy()
    row = np.asarray(row, dtype=np.float64)
    p["start_pos"] = row[:, 1:4]
    p["start_ang"] = row[:, 4:7]
    p["start_vel"] = 0.0
    p["start_ang_vel"] = 0.0
    p["ee"] = row[:, 7:19].reshape(-1, 4, 3)
    p["t_start"] = row[:, 0]
    p["goal"][:, 0] = prev_problems["goal"][:, 0] + step[0]
    p["goal"][:, 1] = prev_problems["goal"][:, 1] + step[1]
    return p


def replan_sweep_problems(n_total, variants, hf_ids, seed=1234, group_size=8):
    """BASELINE config 5, generation 0: n_total windows spread evenly over the terrain variants (multi-start pairs on each
    variant's own grid, groups of `group_size` candidates); the timed generation is made of their successors
    (replan_from_rows on each plan's final, all-stance row)."""
    nv = len(variants)
    per = n_total // nv
    out = []
    for v, (grid, res) in enumerate(variants):
        q = multistart_problems(per, grid, res, seed=seed + v, hf_id=hf_ids[v], group_size=group_size)
        q["group"] += v * ((per + group_size - 1) // group_size)
        out.append(q)
    return np.concatenate(out)


def replan_problems(prev_problems, rows, lookahead_row, step=(0.4, 0.0)):
    """Receding-horizon successors (config 5): the next window of every plan starts from the plan's own state at
    `lookahead_row` of its 1 kHz CSV rows -- what Combiner._state reads back from towr.csv (ref: QTOS/combiner.py:245-296,
    scripts/main.py:177 lookahead) and scripts/main.py turns into -s / -s_ang / -e1..-e4 / -t; the goal moves on by `step`.
    The reference never passes `-n`, so ./main zeroes the start velocities (main.cpp:237-242); mirrored here."""
    from .handoff import state_of_row, EE_NAMES
    n = len(prev_problems)
    p = prev_problems.copy()
    for i in range(n):
        row = rows[i][lookahead_row]
        st = state_of_row(row)
        p["start_pos"][i] = st["CoM"]
        p["start_ang"][i] = st["orientation"]
        p["start_vel"][i] = 0.0
        p["start_ang_vel"][i] = 0.0
        for e, name in enumerate(EE_NAMES):
            p["ee"][i, e] = st[name]
        p["t_start"][i] = row[0]
        p["goal"][i, 0] = prev_problems["goal"][i, 0] + step[0]
        p["goal"][i, 1] = prev_problems["goal"][i, 1] + step[1]
    return p
