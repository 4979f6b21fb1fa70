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
# this round's additions: trajectory rows delivered by the streaming session, the gradient query, a shape with cost terms and
# terrain gradients (objective branches of kip_prepare / kip_step / k_init, k_jac_tg, force tasks of eval_g_block), base_rom
rows_out = np.zeros((6, S.csv_rows, Q.CSV_COLS))
S.stream_begin(Q.default_options(max_iter=10))
S.stream_wait(S.stream_submit(p[:6], csv_out=rows_out))
S.stream_end()
print("stream csv finite", bool(np.isfinite(rows_out).all()))
print("gradients", [float(np.abs(v).max()) for v in S.heightfield_gradients(hid, np.random.default_rng(1).uniform(-1.2, 4.5, (300, 2)))])
sh2 = Q.default_shape(*shape); sh2.cost_force_z = 1.0; sh2.cost_ee_vel_xy = 0.1; sh2.terrain_gradients = 1; sh2.base_rom = 1
S2 = Q.Solver(sh2, max_batch=4)
hid2 = S2.upload_heightfield(grid, res)
p2 = workloads.multistart_problems(6, grid, res, hf_id=hid2)
r2, x2, _ = S2.solve(p2, Q.default_options(max_iter=10))
g2, J2 = S2.eval(p2[:2], x2[:2])
print("options", r2["status"], r2["iters"], bool(np.isfinite(g2).all()))
S2.close()
xy = np.random.default_rng(0).uniform(0.2, 1.0, (4 * 52, 2))
print(S.measure_heightfield_staging(hid, xy, 52))
S.close()
