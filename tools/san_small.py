import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qtos_b200 as Q
from qtos_b200 import heightfield as HF, workloads
grid, res = HF.rough_terrain(1234)
S = Q.Solver(Q.default_shape("C1", 2.0), max_batch=32)
p = workloads.multistart_problems(24, grid, res, hf_id=S.upload_heightfield(grid, res))
r, x, rows = S.solve(p, csv=True)
print(r["status"], r["iters"])
