"""Phase times for small batches (latency view) and for the production shape S5."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qtos_b200 as Q
from qtos_b200 import heightfield as HF, workloads
grid, res = HF.rough_terrain(1234)
for combo, T, batches in [("C1", 2.0, [1, 8, 148, 1024]), ("Custom", 5.0, [1, 148, 2048])]:
    for n in batches:
        S = Q.Solver(Q.default_shape(combo, T), max_batch=n)
        p = workloads.multistart_problems(n, grid, res, hf_id=S.upload_heightfield(grid, res))
        S.solve(p)
        S.set_profiling(True)
        t = time.perf_counter(); r, x, _ = S.solve(p); dt = time.perf_counter() - t
        st = S.last_stats()
        S.set_profiling(False)
        t = time.perf_counter(); r, x, _ = S.solve(p); dt2 = time.perf_counter() - t
        print(combo, "batch", n, "wall %.2f ms (profiled %.2f)" % (1e3 * dt2, 1e3 * dt), "conv", int((r["status"] == 0).sum()), "iters mean %.1f max %d" % (r["iters"].mean(), r["iters"].max()),
              {k: round(v, 2) for k, v in st["ms"].items()}, "solves/s %.0f" % (n / dt2))
        S.close()
