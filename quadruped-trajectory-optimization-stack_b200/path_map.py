"""In-process mirror of the reference's batched feasibility probing (`PATH_MAP`).

The reference walks the world height map, queues one probe (start cell -> cell two columns on) for
every pair with a raised neighbour, fans the queue out over 32 OS processes that each run
`docker exec <id> ./main <flags> -r 5.0` and tests `returncode == 0`, and marks a diamond of cells
around the start and goal of every failed probe (ref: QTOS/generateHeightField.py:172-404).  Here the
whole probe set is ONE `Solver.solve` batch on the GPU; the host logic either side of it is restated
so the resulting `bool_map` is the one the reference builds:

  probe_set            ref: generateHeightField.py:303-342 (probe_map), stated as two 1-D coordinate tables (the
                            reference's add-then-round(., 2) walk along each axis, so coordinates are bit-identical)
                            and a mask over the (row, column pair) grid instead of the nested cell loop
  danger_mask          ref: generateHeightField.py:282-301 (neighbors_danger) for all cells at once
  hull_offsets         ref: generateHeightField.py:226-263 (find_convex_hull; scipy ConvexHull, all scan lines at once)
  probe_problems       ref: generateHeightField.py:365-373 (state_config) + QTOS/utils.py:644-670 (cmd_args)
                            + solver/towr/src/main.cpp:163-306 (the flags ./main would parse)
  mark                 ref: generateHeightField.py:386-404 (success clears start/mid/goal, failure sets the
                            start and goal diamonds; the `neighbors_mid` loop of the reference has no
                            assignment and is a no-op, kept as such)
  PathMap              ref: generateHeightField.py:195-224 (constructor flow, flat ground short cut)

Ordering: the reference's 32 workers race on the shared array, so its result depends on scheduling
when a cleared cell of one probe lies in the diamond of another.  This mirror applies the marks in
queue (FIFO) order, i.e. what the reference produces with one worker; tests pin it against the
reference's own `worker_f` run that way.
"""
import numpy as np

from . import make_problems

ORIGIN_SHIFT = 1.0          # ref: generateHeightField.py:205-206
PROBE_RUNTIME = 5.0         # worker_f's `-r 5.0` (ref: generateHeightField.py:373): Ipopt's max_cpu_time, an iteration budget here


_SCAN = ((1, 0), (-1, 0), (0, 1), (0, -1), (1, 1), (1, -1), (-1, -1), (-1, 1))     # ref :282-301: the order decides


def danger_mask(m, sz=1):
    """danger_mask(m)[ix, iy] == the reference's neighbors_danger(ix, iy) for every cell at once: the 8 neighbours are
    looked at in the reference's order and the FIRST decisive one settles the cell -- a neighbour outside the map says
    False (the reference returns from inside its loop), a raised one says True."""
    m = np.asarray(m)
    nx, ny = m.shape
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    verdict = np.zeros(m.shape, dtype=bool)
    open_ = np.ones(m.shape, dtype=bool)
    for dx, dy in _SCAN:
        jx, jy = ix + sz * dx, iy + sz * dy
        inside = (jx >= 0) & (jx < nx) & (jy >= 0) & (jy < ny)
        raised = np.zeros(m.shape, dtype=bool)
        raised[inside] = m[jx[inside], jy[inside]] > 0
        verdict |= open_ & raised
        open_ &= inside & ~raised
    return verdict


def neighbors_danger(m, ix, iy, sz=1):
    """one cell of danger_mask (kept for callers that ask about single cells)."""
    return bool(danger_mask(m, sz)[ix, iy])


def _rounded_walk(first, step, n):
    """the reference's coordinate bookkeeping along one axis: add `step`, round to 2 decimals, n times (plain Python
    floats -- the rounding after every step is what makes the probe coordinates bit-identical to the reference's)."""
    out, v = [], first
    for _ in range(n):
        v = round(v + step, 2)
        out.append(v)
    return out


def probe_set(m, multi_map_shift=1, mesh_resolution=0.1):
    """-> dict of arrays in queue order: start[n,3], goal[n,3] (world x, y, map height), idx_start[n,2], idx_goal[n,2].

    The reference walks rows (outer) and column PAIRS (inner): probe (ix, k) runs from cell (ix, 2k) to cell (ix, 2k + 2)
    and is queued when either end has a raised neighbour.  Its coordinates do not depend on the map: y advances by one
    resolution step per row, x_goal by two per column pair, and every probe starts where the previous one of its row
    ended (the first starts one step in), each value rounded to 2 decimals as it is formed.  So the coordinates are two
    1-D tables and the queue is a mask over the (row, pair) grid, read in row-major order."""
    m = np.asarray(m)
    step = mesh_resolution
    nrow, npair = m.shape[0], max(m.shape[1] // 2 - 1, 0)
    shift = (multi_map_shift - 1) * ORIGIN_SHIFT
    half = mesh_resolution * (m.shape[1] / 2)            # the reference uses shape[1] for both axes
    lo, hi = -half - mesh_resolution / 2 + shift, -half + mesh_resolution / 2 + shift
    ys = _rounded_walk(lo, step, nrow)                   # y of row ix (start and goal alike)
    xg = _rounded_walk(hi, 2 * step, npair)              # x_goal of pair k
    xs = ([round(lo + step, 2)] + xg[:-1]) if npair else []
    risky = danger_mask(m)
    cols = 2 * np.arange(npair)
    queued = risky[:, cols] | risky[:, cols + 2] if npair else np.zeros((nrow, 0), dtype=bool)
    ix, k = np.nonzero(queued)                           # row-major = the reference's queue order
    iy, iy_off = 2 * k, 2 * k + 2
    ys, xs, xg = np.array(ys, dtype=np.float64), np.array(xs, dtype=np.float64), np.array(xg, dtype=np.float64)
    start = np.stack([xs[k], ys[ix], m[ix, iy].astype(np.float64)], axis=1) if len(ix) else np.zeros((0, 3))
    goal = np.stack([xg[k], ys[ix], m[ix, iy_off].astype(np.float64)], axis=1) if len(ix) else np.zeros((0, 3))
    return {"start": start.reshape(-1, 3), "goal": goal.reshape(-1, 3),
            "idx_start": np.stack([ix, iy], axis=1).astype(np.int64).reshape(-1, 2),
            "idx_goal": np.stack([ix, iy_off], axis=1).astype(np.int64).reshape(-1, 2)}


def hull_offsets(points):
    """cells of the filled convex hull of `points`, relative to the centre of its bounding box (ref :226-263): every scan
    line y is filled between the smallest and the largest of its cuts with the hull's non-horizontal edges, a cut being
    the edge's x at y truncated towards zero; lines with fewer than two cuts stay empty.  All lines and edges at once."""
    from scipy.spatial import ConvexHull
    pts = np.asarray(points)
    hv = pts[ConvexHull(pts).vertices]
    a, b = hv, np.roll(hv, -1, axis=0)                   # edges a -> b
    x0, y0 = int(hv[:, 0].min()), int(hv[:, 1].min())
    w, h = int(hv[:, 0].max()) - x0 + 1, int(hv[:, 1].max()) - y0 + 1
    y = (y0 + np.arange(h))[:, None]                     # [line, edge]
    ya, yb, xa, xb = a[None, :, 1], b[None, :, 1], a[None, :, 0], b[None, :, 0]
    slanted = ya != yb
    hit = slanted & (((ya <= y) & (y <= yb)) | ((yb <= y) & (y <= ya)))
    with np.errstate(divide="ignore", invalid="ignore"):
        cut = np.trunc(xa + (xb - xa) * (y - ya) / np.where(slanted, yb - ya, 1)).astype(np.int64)
    big = np.iinfo(np.int64).max
    left = np.where(hit, cut, big).min(axis=1)
    right = np.where(hit, cut, -big).max(axis=1)
    filled = hit.sum(axis=1) >= 2
    xcol = x0 + np.arange(w)[None, :]
    grid = filled[:, None] & (xcol >= left[:, None]) & (xcol <= right[:, None])
    return np.argwhere(grid) - np.array([h // 2, w // 2])


def diamond(scale=1):
    """the start/goal neighbourhood of the reference (generateHeightField.py:215,217,219)."""
    k = scale * 3
    return hull_offsets(((-k, 0), (k, 0), (0, -k), (0, k)))


def probe_problems(probes, hf_id=0):
    """one qtos_problem per probe, with the values ./main would parse from worker_f's command line:
    -s start + (0, 0, 0.24); -e1..-e4 nominal stance + start (x, y, map height); -s_ang 0 0 0;
    -g goal + (0, 0, 0.24).  `str(float)` -> std::stod round-trips doubles exactly, so no text detour."""
    n = len(probes["start"])
    p = make_problems(n)
    if n == 0:
        return p
    s, g = probes["start"], probes["goal"]
    p["start_pos"][:, 0], p["start_pos"][:, 1], p["start_pos"][:, 2] = s[:, 0], s[:, 1], s[:, 2] + 0.24
    p["goal"][:, 0], p["goal"][:, 1], p["goal"][:, 2] = g[:, 0], g[:, 1], g[:, 2] + 0.24
    for e, (a, b) in enumerate(((0.21, 0.19), (0.21, -0.19), (-0.21, 0.19), (-0.21, -0.19))):
        p["ee"][:, e, 0] = a + s[:, 0]
        p["ee"][:, e, 1] = b + s[:, 1]
        p["ee"][:, e, 2] = 0.0 + s[:, 2]
    p["hf_id"] = hf_id
    p["group"] = np.arange(n)
    return p


def _stamp(out, cx, cy, offsets):
    q = offsets + np.array([cx, cy])
    ok = (q[:, 0] >= 0) & (q[:, 0] < out.shape[0]) & (q[:, 1] >= 0) & (q[:, 1] < out.shape[1])
    out[q[ok, 0], q[ok, 1]] = 1


def mark(shape, probes, feasible, offsets_start, offsets_end=None):
    """bool_map after the probes have been applied in queue order (a later probe overwrites an earlier one's cells)."""
    offsets_end = offsets_start if offsets_end is None else offsets_end
    offsets_start, offsets_end = np.asarray(offsets_start).reshape(-1, 2), np.asarray(offsets_end).reshape(-1, 2)
    out = np.zeros(shape, dtype=np.float32)        # the reference's shared int array is used as float32
    for (sx, sy), (gx, gy), ok in zip(probes["idx_start"].tolist(), probes["idx_goal"].tolist(), feasible):
        if ok:
            out[sx, sy] = out[sx, sy + 1] = out[gx, gy] = 0
        else:
            _stamp(out, sx, sy, offsets_start)
            _stamp(out, gx, gy, offsets_end)
    return out.astype("int")


class PathMap:
    """PATH_MAP with the 32-process `docker exec ./main` fan-out replaced by one batched solve.

    map            world height map (Height_Map_Generator.map: row = y, col = x)
    solver         qtos_b200.Solver of the production shape (Custom gait, 5 s), max_batch >= 1
    towr_grid, resolution   the grid ./main would read (towr_heightfield.txt) and the `-resolution` it gets;
                   worker_f passes no -resolution, so main.cpp's default 0.1 applies (main.cpp:293-297)
    """

    def __init__(self, map, solver, towr_grid, multi_map_shift=1, scale=1, resolution=0.1, options=None):
        self.map = np.asarray(map)
        self.mesh_resolution = 0.1 * (1 / scale)
        self.probes = probe_set(self.map, multi_map_shift, self.mesh_resolution)
        self.neighbors_start = diamond(scale)
        self.neighbors_end = diamond(scale)
        self.bool_map = np.zeros(self.map.shape, dtype=int)
        self.results = None
        if np.all(self.map == 0):                   # check_flat_ground: nothing to probe
            return
        n = len(self.probes["start"])
        if n == 0:
            return
        hid = solver.upload_heightfield(np.asarray(towr_grid, dtype=np.float64), resolution)
        if options is None:                          # what `./main ... -r 5.0` would run with
            from . import default_options
            options = default_options(max_cpu_time=PROBE_RUNTIME)
        res, _x, _ = solver.solve(probe_problems(self.probes, hid), options)
        self.results = res
        self.feasible = res["status"] == 0           # `p_status.returncode == 0`
        self.bool_map = mark(self.map.shape, self.probes, self.feasible, self.neighbors_start, self.neighbors_end)
