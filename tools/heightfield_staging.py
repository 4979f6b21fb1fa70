"""Measures heightfield tile staging through the TMA unit against direct global reads: the terrain queries of
4096 constraint evaluations of the bench workload (52 per evaluation for shape S5: 4 feet x 13 nodes; positions = each window's
initial guess), one CTA per evaluation.  Run under gpurun."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qtos_b200 as Q
from qtos_b200 import heightfield as HF, workloads

grid, res = HF.rough_terrain(1234)
S = Q.Solver(Q.default_shape("Custom", 5.0), max_batch=64)
hid = S.upload_heightfield(grid, res)
n = 4096
p = workloads.multistart_problems(n, grid, res, hf_id=hid)
rng = np.random.default_rng(0)
# 52 foot positions per window: on the segment start -> goal around each nominal foothold, like the node positions of the initial guess
t = np.linspace(0, 1, 13)
xy = np.zeros((n, 4, 13, 2))
for e in range(4):
    s0 = p["ee"][:, e, :2]
    g0 = p["goal"][:, :2] + (p["ee"][:, e, :2] - p["start_pos"][:, :2])
    xy[:, e] = s0[:, None, :] + t[None, :, None] * (g0 - s0)[:, None, :] + 0.01 * rng.standard_normal((n, 13, 2))
xy = xy.reshape(n * 52, 2)
r = S.measure_heightfield_staging(hid, xy, 52)
print("4096 evaluations x 52 queries on the 256x256 grid (res 0.02):", r)
print("direct %.1f us, staged %.1f us per launch; per evaluation %.2f / %.2f ns; identical answers: %s" % (
    1e3 * r["ms_direct"], 1e3 * r["ms_staged"], 1e6 * r["ms_direct"] / n, 1e6 * r["ms_staged"] / n, r["max_diff"] == 0.0))
g5 = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "heightfields.npz"))
grid5, res5 = g5["exp_5_towr"], float(g5["exp_5_res"])
hid5 = S.upload_heightfield(grid5, res5)
xy5 = xy.copy(); xy5[:, 0] = np.clip(xy5[:, 0] * 0.8 - 0.5, -0.9, 2.9); xy5[:, 1] = np.clip(xy5[:, 1] * 0.3, -0.9, 0.9)
r5 = S.measure_heightfield_staging(hid5, xy5, 52)
print("exp_5 stairs grid %s (res 1/110):" % (grid5.shape,), r5)
S.close()
