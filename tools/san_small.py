import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qtos_b200 as Q
from qtos_b200 import heightfield as HF, workloads
grid, res = HF.rough_terrain(1234)
shape = ("Custom", 5.0) if os.environ.get("QTOS_SHAPE") == "S5" else ("C1", 2.0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
S = Q.Solver(Q.default_shape(*shape), max_batch=32)
p = workloads.multistart_problems(n, grid, res, hf_id=S.upload_heightfield(grid, res))
r, x, rows = S.solve(p, csv=True)
print(r["status"], r["iters"])
