"""Heightfield producer/consumer side of the local-planner boundary (host, numpy).

Mirrors what QTOS/generateHeightField.py does around the TOWR call so that the grid handed to
the solver is indexed bit-exactly like the reference's:
  tile reader            ref: QTOS/generateHeightField.py:100-118  (file rows -> transpose)
  scale_map              ref: QTOS/generateHeightField.py:39-56    (element/row replication)
  tile concatenation     ref: QTOS/generateHeightField.py:479-491  (side by side along columns)
  towr grid              ref: QTOS/generateHeightField.py:568,607-632 (transpose, rows shifted down by one)
  file format            ref: QTOS/generateHeightField.py:590-605  ("v, v, ...," per row, no final newline)
  resolution             ref: QTOS/generateHeightField.py:574       (1 / (rows / 2))
  text parser            ref: solver/towr/src/custom_terrain.cpp:22-49
  GetHeight              ref: solver/towr/src/custom_terrain.cpp:51-94 (numpy restatement, used by the
                              host only to place synthetic start states on the terrain)
"""
import numpy as np


def read_tile(path, delimiter=","):
    rows = []
    with open(path) as f:
        for line in f.readlines():
            vals = []
            for tok in line.strip().split(delimiter):
                try:
                    vals.append(float(tok))
                except ValueError:
                    pass
            rows.append(vals)
    return np.transpose(np.array(rows))


def scale_map(m, scale_factor=1):
    m = np.asarray(m)
    return np.repeat(np.repeat(m, scale_factor, axis=1), scale_factor, axis=0)


def combine_tiles(tiles):
    return np.concatenate([np.asarray(t) for t in tiles], axis=1)


def towr_grid(world_map):
    """heightfield.txt map[row=y][col=x] -> towr_heightfield grid hf[ix][iy] (rows shifted down by one)."""
    t = np.transpose(np.asarray(world_map, dtype=np.float64))
    out = np.zeros_like(t)
    out[1:] = t[:-1]
    return out


def resolution(world_map):
    return 1.0 / (np.asarray(world_map).shape[0] / 2)


def write_heightfield(path, data):
    data = np.asarray(data)
    with open(path, "w") as f:
        for i, line in enumerate(data):
            f.write(", ".join(str(v) for v in line) + ",")
            if i < len(data) - 1:
                f.write("\n")


def read_towr_heightfield(path):
    """Parse like CustomTerrain::ReadHeightField: one grid row (world x) per text line."""
    rows = []
    with open(path) as f:
        for line in f.read().split("\n"):
            vals = [float(t) for t in line.replace(",", " ").split()]
            rows.append(vals)
    while rows and not rows[-1]:
        rows.pop()
    if not rows:
        raise ValueError("empty heightfield file: %s" % path)
    n = len(rows[0])
    if any(len(r) != n for r in rows):
        raise ValueError("ragged heightfield file: %s" % path)
    return np.array(rows, dtype=np.float64)


def cell_indices(grid_shape, res, x, y):
    """(ix0, iy0, ix1, iy1) exactly as CustomTerrain::GetHeight picks them (negative/NaN -> last cell)."""
    nx, ny = grid_shape
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    xf = np.floor((x - (-1.0)) / res)
    yf = np.floor((y - (-1.0)) / res)

    def clamp(fl, size):
        bad = ~(fl >= 0.0) | (fl >= size - 1)
        return np.where(bad, size - 1, np.where(bad, 0, fl)).astype(np.int64)

    ix0, iy0 = clamp(xf, nx), clamp(yf, ny)
    return ix0, iy0, np.minimum(ix0 + 1, nx - 1), np.minimum(iy0 + 1, ny - 1)


def get_height(grid, res, x, y):
    grid = np.asarray(grid, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    ix0, iy0, ix1, iy1 = cell_indices(grid.shape, res, x, y)
    x0, x1 = ix0 * res + (-1.0), ix1 * res + (-1.0)
    y0, y1 = iy0 * res + (-1.0), iy1 * res + (-1.0)
    z00, z01, z10, z11 = grid[ix0, iy0], grid[ix0, iy1], grid[ix1, iy0], grid[ix1, iy1]
    s = 1 / (res * res)
    u0, u1 = s * (x1 - x), s * (x - x0)
    w0 = u0 * z00 + u1 * z10
    w1 = u0 * z01 + u1 * z11
    return w0 * (y1 - y) + w1 * (y - y0)


def rough_terrain(seed=1234, n_plate=32, up=8, hmax=0.075):
    """Synthetic 256x256 rough grid of SURVEY 8(d) config 4: 32x32 plateaus U(0, 0.075) m rounded to
    1e-4 (range of data/heightfields/random_terrain.txt), nearest-upsampled x8 like scale_map; res 0.02."""
    rng = np.random.default_rng(seed)
    plate = np.round(rng.uniform(0, hmax, (n_plate, n_plate)), 4)
    return scale_map(plate, up).astype(np.float64), 0.02
