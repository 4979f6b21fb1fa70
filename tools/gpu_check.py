"""Development GPU check: eval parity vs the oracle and a few solves (run under gpurun)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qtos_b200 as Q
import oracle as O
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from proto_ipm import rough_terrain

def oracle_problem(shape_o, pr, grid, res):
    ter = O.Terrain(grid, res)
    inst = O.make_instance(start_pos=pr["start_pos"], start_ang=pr["start_ang"], goal=pr["goal"], ee=pr["ee"], t_start=pr["t_start"])
    return O.Problem(shape_o, inst, ter)

def gen(n, grid, res, seed=1234):
    ter = O.Terrain(grid, res)
    p = Q.make_problems(n)
    for i in range(n):
        rng = np.random.default_rng(seed + 1 + i)
        sx, sy = rng.uniform(0, 2.5, 2)
        gx, gy = sx + rng.uniform(0.2, 0.6), sy + rng.uniform(-0.1, 0.1)
        p[i]["start_pos"] = (sx, sy, ter.height(sx, sy) + 0.24)
        p[i]["goal"] = (gx, gy, 0.24)
        p[i]["ee"] = [(sx + a, sy + b, ter.height(sx + a, sy + b)) for a, b in [(0.21, 0.19), (0.21, -0.19), (-0.21, 0.19), (-0.21, -0.19)]]
    return p

for combo, T in [("C1", 2.0), ("Custom", 5.0)]:
    grid, res = rough_terrain()
    S = Q.Solver(Q.default_shape(combo, T), max_batch=64)
    print(combo, "dims", S.n_vars, S.n_cons, S.dims.n_free, "blocks", S.dims.kkt_blocks, "ws/problem", S.dims.workspace_bytes_per_problem)
    hid = S.upload_heightfield(grid, res)
    n = 4
    p = gen(n, grid, res); p["hf_id"] = hid
    x0, xl, xu, gl, gu = S.initial(p)
    so = O.default_shape(combo, T)
    for i in range(n):
        po = oracle_problem(so, p[i], grid, res)
        ox0 = po.x0(); oxl, oxu, ogl, ogu = po.bounds()
        ox0[oxl == oxu] = oxl[oxl == oxu]
        print(" x0 diff", np.abs(x0[i] - ox0).max(), "bounds", np.abs(np.clip(xl[i],-1e20,1e20) - oxl).max(), np.abs(gl[i]-ogl).max(), np.abs(gu[i]-ogu).max())
        rng = np.random.default_rng(i)
        xr = ox0 + 0.05 * rng.standard_normal(len(ox0)); xr[oxl == oxu] = oxl[oxl == oxu]
        g, J = S.eval(p[i:i+1], xr[None])
        og = po.g(xr); oJ = po.jac(xr)
        oJ[:, oxl == oxu] = 0
        print(" g diff", np.abs(g[0] - og).max(), "J diff", np.abs(J[0] - oJ).max(), "J scale", np.abs(oJ).max())
    t = time.time(); res_, x, _ = S.solve(p); dt = time.time() - t
    print(" solve", res_["status"], res_["iters"], res_["constr_viol"], "%.3fs" % dt)
    for i in range(n):
        po = oracle_problem(so, p[i], grid, res)
        xo, ro = po.solve()
        print("  oracle", ro.status, ro.iters, ro.constr_viol, "x diff", np.abs(x[i] - xo).max(), "csv diff", np.abs(po.csv(x[i]) - po.csv(xo))[:, 1:19].max())
    rows = S.sample_csv(p, x)
    po = oracle_problem(so, p[0], grid, res)
    print(" csv sampler diff", np.abs(rows[0] - po.csv(x[0])).max())
    S.set_profiling(True)
    big = gen(64, grid, res); big["hf_id"] = hid
    t = time.time(); r2, x2, _ = S.solve(big); dt = time.time() - t
    print(" batch64 %.3fs" % dt, "conv", (r2["status"] == 0).sum(), "iters", r2["iters"].min(), r2["iters"].max(), S.last_stats())
    S.set_profiling(False)
