"""Development: iterations / convergence of the bench workload under different interior-point options."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qtos_b200 as Q
from bench import build_workload, COMBO, DURATION
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
grid, res, p = build_workload(n)
S = Q.Solver(Q.default_shape(COMBO, DURATION), max_batch=n)
p["hf_id"] = S.upload_heightfield(grid, res)
for mu, sw in itertools.product([0.1, 0.03, 0.01, 1e-3, 1e-4], [0.1, 0.03, 0.01, 0.3]):
    o = Q.default_options(mu_init=mu, sigma_w=sw)
    r, x, _ = S.solve(p, o)
    it = r["iters"]
    print("mu_init %-7g sigma_w %-5g converged %5d/%d iters mean %.2f p50 %d p90 %d max %d viol max %.1e" % (
        mu, sw, int((r["status"] == 0).sum()), n, it.mean(), np.median(it), np.percentile(it, 90), it.max(), r["constr_viol"][r["status"] == 0].max()), flush=True)
