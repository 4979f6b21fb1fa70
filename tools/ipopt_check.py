"""Development GPU check of QTOS_ALG_IPOPT (run under gpurun): the reference's logged solves (iteration tables of
logs/towr_log.out, plans of data/traj/towr.csv) and rough-terrain windows against oracle/towr_ipopt.c."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import qtos_b200 as Q
import oracle as O
from qtos_b200 import heightfield as HF, workloads

G = os.path.join(ROOT, "tests", "golden")
log = json.load(open(os.path.join(G, "towr_log.json")))
csv = np.load(os.path.join(G, "gait_csv.npz"))
nwin = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nbig = int(sys.argv[2]) if len(sys.argv) > 2 else 0


def oracle_problem(so, pr, grid, res):
    inst = O.make_instance(start_pos=pr["start_pos"], start_ang=pr["start_ang"], goal=pr["goal"], ee=pr["ee"], t_start=float(pr["t_start"]))
    return O.Problem(so, inst, O.Terrain(grid, res))


# ---- the logged solves: shape S5 of the golden build (m = 3.0, max_dev_x = 0.08)
sh = Q.default_shape("Custom", 5.0, mass=3.0); sh.max_dev[0] = 0.08
so = O.default_shape("Custom", 5.0, mass=3.0); so.max_dev[0] = 0.08
S = Q.Solver(sh, max_batch=4)
flat = np.zeros((600, 200))
hid = S.upload_heightfield(flat, 0.01)
p = Q.make_problems(3)
for k, inp in enumerate(log["inputs"]):
    p[k]["start_pos"] = inp["start_pos"]; p[k]["start_ang"] = inp["start_ang"]; p[k]["goal"] = inp["goal"]
    p[k]["ee"] = inp["ee"]; p[k]["t_start"] = inp["t_start"]; p[k]["hf_id"] = hid
t = time.time(); r, x, rows = S.solve(p, csv=True); dt = time.time() - t
tr = S.trace(3)
print("logged solves: status", r["status"], "iters", r["iters"], "(log:", log["iters"], ") viol", r["constr_viol"], "%.3fs" % dt)
for k in range(3):
    tab = log["iteration_tables"][k]
    for g in tab:
        i = g["iter"]
        print("  s%d it%d  gpu inf_pr %.2e inf_du %.2e lg(mu) %5.1f |d| %.2e a_du %.2e a_pr %.2e%s ls %d | log %s %s %s %s %s %s%s %d" % (
            k, i, tr[k, i, 0], tr[k, i, 1], np.log10(max(tr[k, i, 2], 1e-300)), tr[k, i, 3], tr[k, i, 4], tr[k, i, 5], chr(int(tr[k, i, 7])) if tr[k, i, 7] else " ", tr[k, i, 6],
            g["inf_pr"], g["inf_du"], g["lg_mu"], g["dnorm"], g["alpha_du"], g["alpha_pr"], g["tag"], g["ls"]))
for k, (Gk, row0) in enumerate(((csv["towr_g4"], 2502), (csv["towr_g2"], 0))):
    rk = rows[k][row0::10][:len(Gk)]
    print("  plan %d vs TOWR+Ipopt golden csv: CoM %.2e  angles %.2e  feet %.2e  forces %.2e" % (
        k, np.abs(rk[:, 1:4] - Gk[:, 1:4]).max(), np.abs(rk[:, 4:7] - Gk[:, 4:7]).max(), np.abs(rk[:, 7:19] - Gk[:, 7:19]).max(), np.abs(rk[1:, 25:37] - Gk[1:, 25:37]).max()))
    po = oracle_problem(so, p[k], flat, 0.01)
    xo, ro = po.solve_ipopt()
    print("  oracle C port: status %d iters %d; gpu-vs-oracle plan diff %.2e" % (ro.status, ro.iters, np.abs(rows[k][:, 1:19] - po.csv(xo)[:, 1:19]).max()))
S.close()

# ---- rough terrain, batched shape S2
grid, res = HF.rough_terrain(1234)
S = Q.Solver(Q.default_shape("C1", 2.0), max_batch=max(nwin, nbig, 1))
hid = S.upload_heightfield(grid, res)
pr = workloads.multistart_problems(max(nwin, nbig), grid, res, hf_id=hid)
r, x, _ = S.solve(pr[:nwin])
so2 = O.default_shape("C1", 2.0)
d = []; same = 0; st_same = 0
for i in range(nwin):
    po = oracle_problem(so2, pr[i], grid, res)
    xo, ro = po.solve_ipopt()
    d.append(np.abs(po.csv(x[i])[::10, 1:19] - po.csv(xo)[::10, 1:19]).max())
    same += int(ro.iters == r["iters"][i]); st_same += int(ro.status == r["status"][i])
    if ro.status != r["status"][i] or ro.iters != r["iters"][i]:
        print("   window %d: gpu (%d, %d it) oracle (%d, %d it) diff %.2e" % (i, r["status"][i], r["iters"][i], ro.status, ro.iters, d[-1]))
d = np.array(d)
print("rough terrain S2, %d windows: gpu converged %d, same status %d, same iters %d, plan diff p50 %.2e p90 %.2e max %.2e, viol max (converged) %.2e" % (
    nwin, (r["status"] == 0).sum(), st_same, same, np.percentile(d, 50), np.percentile(d, 90), d.max(), r["constr_viol"][r["status"] == 0].max()))
if nbig:
    S.set_profiling(True)
    for rep in range(2):
        t = time.time(); rb, xb, _ = S.solve(pr[:nbig]); dt = time.time() - t
    st = S.last_stats()
    print("batch %d: %.3fs  %.0f solves/s  converged %d  status counts %s  iters p50 %d max %d" % (
        nbig, dt, nbig / dt, (rb["status"] == 0).sum(), dict(zip(*np.unique(rb["status"], return_counts=True))), np.median(rb["iters"]), rb["iters"].max()))
    print("   ", st)
    of = Q.default_options(algorithm=Q.ALG_FAST)
    for rep in range(2):
        t = time.time(); rb, xb, _ = S.solve(pr[:nbig], options=of); dt = time.time() - t
    print("FAST batch %d: %.3fs  %.0f solves/s  converged %d" % (nbig, dt, nbig / dt, (rb["status"] == 0).sum()), S.last_stats())
S.close()
