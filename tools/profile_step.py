"""One bench step (4096 windows, shape S2) without the CPU arm -- the command profiled under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qtos_b200 as Q
from bench import build_workload, COMBO, DURATION

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
grid, res, p = build_workload(n)
if os.environ.get("QTOS_SHAPE") == "S5":        # production shape (Custom gait, 5 s) on the same workload
    COMBO, DURATION = "Custom", 5.0
S = Q.Solver(Q.default_shape(COMBO, DURATION), max_batch=n)
p["hf_id"] = S.upload_heightfield(grid, res)
S.set_profiling(True)
for _ in range(steps):
    r, x, _ = S.solve(p)
    st = S.last_stats()
    print("converged", int((r["status"] == 0).sum()), "of", n, "iters mean %.2f" % r["iters"].mean(), st)
