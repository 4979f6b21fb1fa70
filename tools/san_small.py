"""Small solves for compute-sanitizer (memcheck / racecheck / synccheck): batch call (IPOPT with a forced second attempt, FAST),
the streaming session with pool refill, the selection kernels and the heightfield staging kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qtos_b200 as Q
from qtos_b200 import heightfield as HF, workloads, parallel
grid, res = HF.rough_terrain(1234)
shape = ("Custom", 5.0) if os.environ.get("QTOS_SHAPE") == "S5" else ("C1", 2.0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
S = Q.Solver(Q.default_shape(*shape), max_batch=8)
hid = S.upload_heightfield(grid, res)
p = workloads.multistart_problems(n, grid, res, hf_id=hid)
r, x, rows = S.solve(p, Q.default_options(max_iter=12), csv=True)              # chunks of 8
print("ipopt", r["status"], r["iters"])
r, x, _ = S.solve(p[:4], Q.default_options(max_iter=3, retry_failed=1, tol=1e-9))  # nobody converges; exercises the exit paths
r, x, _ = S.solve(p[:4], Q.default_options(algorithm=Q.ALG_FAST, max_iter=8))
print("fast", r["status"], r["iters"])
S.stream_begin(Q.default_options(max_iter=10))
t = [S.stream_submit(p[:8]), S.stream_submit(p[4:n])]
for k in t:
    rr, xx = S.stream_wait(k)
print("stream", rr["status"], S.stream_info())
S.stream_end()
d_res = torch.from_numpy(r.view(np.uint8).reshape(len(r), -1)).cuda()
d_group = torch.zeros(len(r), dtype=torch.int32, device="cuda")
print("winners", parallel.select_best_device(S, d_res, d_group, 0, 1, 2).cpu())
xy = np.random.default_rng(0).uniform(0.2, 1.0, (4 * 52, 2))
print(S.measure_heightfield_staging(hid, xy, 52))
S.close()
