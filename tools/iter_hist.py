"""Development: iteration-count histogram and final statuses of the bench workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qtos_b200 as Q
from bench import build_workload, COMBO, DURATION
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
grid, res, p = build_workload(n)
S = Q.Solver(Q.default_shape(COMBO, DURATION), max_batch=n)
p["hf_id"] = S.upload_heightfield(grid, res)
r, x, _ = S.solve(p)
it = r["iters"]
print("iters histogram:", {int(k): int(v) for k, v in zip(*np.unique(it, return_counts=True))})
for st in np.unique(r["status"]):
    m = r["status"] == st
    print("status", int(st), "count", int(m.sum()), "iters", sorted(it[m].tolist())[-12:], "viol max %.2e" % np.nanmax(r["constr_viol"][m]))
