"""In-process mirror of the reference's batched feasibility probing (`PATH_MAP`).

The reference walks the world height map, queues one probe (start cell -> cell two columns on) for
every pair with a raised neighbour, fans the queue out over 32 OS processes that each run
`docker exec <id> ./main <flags> -r 5.0` and tests `returncode == 0`, and marks a diamond of cells
around the start and goal of every failed probe (ref: QTOS/generateHeightField.py:172-404).  Here the
whole probe set is ONE `Solver.solve` batch on the GPU; the host logic either side of it is restated
so the resulting `bool_map` is the one the reference builds:

  probe_set            ref: generateHeightField.py:303-342 (probe_map; same float accumulation and
                            round(., 2) calls, so coordinates are bit-identical)
  neighbors_danger     ref: generateHeightField.py:282-301
  hull_offsets         ref: generateHeightField.py:226-263 (find_convex_hull; scipy ConvexHull + scan lines)
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
PROBE_RUNTIME = 5.0         # `-r 5.0`, accepted and ignored by the GPU solver


def neighbors_danger(m, ix, iy, sz=1):
    """True when one of the 8 neighbours is raised; a neighbour outside the map ends the scan with
    False (the reference returns from inside the loop, order of the offsets matters)."""
    for dx, dy in ((sz, 0), (-sz, 0), (0, sz), (0, -sz), (sz, sz), (sz, -sz), (-sz, -sz), (-sz, sz)):
        if dx + ix >= m.shape[0] or dx + ix < 0:
            return False
        elif dy + iy >= m.shape[1] or dy + iy < 0:
            return False
        elif m[dx + ix][dy + iy] > 0:
            return True
    return False


def probe_set(m, multi_map_shift=1, mesh_resolution=0.1):
    """-> dict of arrays in queue order: start[n,3], goal[n,3] (world x, y, map height), idx_start[n,2],
    idx_goal[n,2].  Plain Python floats on purpose: the reference accumulates and rounds step by step."""
    m = np.asarray(m)
    step = mesh_resolution
    x_start = -mesh_resolution * (m.shape[1] / 2) - mesh_resolution / 2 + ((multi_map_shift - 1) * ORIGIN_SHIFT)
    y_start = -mesh_resolution * (m.shape[1] / 2) - mesh_resolution / 2 + ((multi_map_shift - 1) * ORIGIN_SHIFT)
    x_goal = -mesh_resolution * (m.shape[1] / 2) + mesh_resolution / 2 + ((multi_map_shift - 1) * ORIGIN_SHIFT)
    y_goal = -mesh_resolution * (m.shape[1] / 2) - mesh_resolution / 2 + ((multi_map_shift - 1) * ORIGIN_SHIFT)
    _x_start, _y_start, _x_goal, _y_goal = x_start, y_start, x_goal, y_goal
    ix, iy, iy_off = 0, 0, 2
    S, G, IS, IG = [], [], [], []
    for _ in range(m.shape[0]):
        _y_start += step
        _y_goal += step
        _x_start, _x_goal = x_start, x_goal
        for y in range(m.shape[1] // 2 - 1):
            if y == 0:
                _x_start += step
                iy, iy_off = 0, 2
            else:
                _x_start = _x_goal
            _x_goal += 2 * step
            _x_start, _y_start = round(_x_start, 2), round(_y_start, 2)
            _x_goal, _y_goal = round(_x_goal, 2), round(_y_goal, 2)
            if neighbors_danger(m, ix, iy) or neighbors_danger(m, ix, iy_off):
                S.append((_x_start, _y_start, float(m[ix][iy])))
                G.append((_x_goal, _y_goal, float(m[ix][iy_off])))
                IS.append((ix, iy))
                IG.append((ix, iy_off))
            iy += 2
            iy_off += 2
        ix += 1
    return {"start": np.array(S, dtype=np.float64).reshape(-1, 3), "goal": np.array(G, dtype=np.float64).reshape(-1, 3),
            "idx_start": np.array(IS, dtype=np.int64).reshape(-1, 2), "idx_goal": np.array(IG, dtype=np.int64).reshape(-1, 2)}


def hull_offsets(points):
    """cells of the filled convex hull of `points`, relative to the centre of its bounding box."""
    from scipy.spatial import ConvexHull
    points = np.array(points)
    hv = points[ConvexHull(points).vertices]
    min_x, max_x = np.min(hv[:, 0]), np.max(hv[:, 0])
    min_y, max_y = np.min(hv[:, 1]), np.max(hv[:, 1])
    grid = np.zeros((max_y - min_y + 1, max_x - min_x + 1), dtype=int)
    for y in range(min_y, max_y + 1):
        cuts = []
        for i in range(len(hv)):
            x1, y1 = hv[i]
            x2, y2 = hv[(i + 1) % len(hv)]
            if y1 == y2:
                continue
            if y1 <= y <= y2 or y2 <= y <= y1:
                cuts.append(int(x1 + (x2 - x1) * (y - y1) / (y2 - y1)))
        cuts.sort()
        if len(cuts) >= 2:
            grid[y - min_y, cuts[0] - min_x:cuts[-1] - min_x + 1] = 1
    return np.argwhere(grid == 1) - np.array([grid.shape[0] // 2, grid.shape[1] // 2])


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


def mark(shape, probes, feasible, offsets_start, offsets_end=None):
    """bool_map after the probes have been applied in queue order."""
    offsets_end = offsets_start if offsets_end is None else offsets_end
    out = np.zeros(shape, dtype=np.float32)        # the reference's shared int array is used as float32
    for k in range(len(probes["idx_start"])):
        sx, sy = (int(v) for v in probes["idx_start"][k])
        gx, gy = (int(v) for v in probes["idx_goal"][k])
        if feasible[k]:
            out[sx, sy] = 0
            out[sx, sy + 1] = 0
            out[gx, gy] = 0
        else:
            for dx, dy in offsets_start:
                if 0 <= sx + dx < shape[0] and 0 <= sy + dy < shape[1]:
                    out[sx + dx, sy + dy] = 1
            for dx, dy in offsets_end:
                if 0 <= gx + dx < shape[0] and 0 <= gy + dy < shape[1]:
                    out[gx + dx, gy + dy] = 1
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
        res, _x, _ = solver.solve(probe_problems(self.probes, hid), options)
        self.results = res
        self.feasible = res["status"] == 0           # `p_status.returncode == 0`
        self.bool_map = mark(self.map.shape, self.probes, self.feasible, self.neighbors_start, self.neighbors_end)
