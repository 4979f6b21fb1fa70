"""GPU tier (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on identical inputs.

Tolerances (north_star): constraint violation <= 1e-4; CoM / foot trajectories within 1 mm of the reference's
TOWR + Ipopt plans (golden CSVs) and of the oracle running the same algorithm; heightfield heights and cell
indices bit-exact; g and J to round-off (1e-10 absolute on values of order 1e0..1e3).

Both algorithms of the library are covered: QTOS_ALG_IPOPT (default; oracle/towr_ipopt.c) and QTOS_ALG_FAST
(oracle/towr_ipm.c)."""
import os

import numpy as np
import pytest

import qtos_b200 as Q
from qtos_b200 import heightfield as HF
from qtos_b200 import towr_cli, workloads
from conftest import oracle_problem, FEET_19

pytestmark = pytest.mark.gpu

G_TOL = 1e-10
J_TOL = 1e-10
TRAJ_TOL_M = 1e-3


@pytest.fixture(scope="module")
def solvers():
    S = {"S2": Q.Solver(Q.default_shape("C1", 2.0), max_batch=256), "S5": Q.Solver(Q.default_shape("Custom", 5.0), max_batch=64)}
    yield S
    for s in S.values():
        s.close()


SHAPES = {"S2": ("C1", 2.0), "S5": ("Custom", 5.0)}
ALGS = {"ipopt": Q.ALG_IPOPT, "fast": Q.ALG_FAST}


def _opts(alg):
    return Q.default_options(algorithm=ALGS[alg])


def _oracle_solve(po, alg):
    return po.solve_ipopt() if alg == "ipopt" else po.solve()


def _rough(S, n, seed=1234):
    grid, res = HF.rough_terrain(seed)
    hid = S.upload_heightfield(grid, res)
    return workloads.multistart_problems(n, grid, res, seed=seed, hf_id=hid), grid, res


@pytest.mark.parametrize("shape", ["S2", "S5"])
def test_structure_and_initial_point(solvers, oracle, shape):
    S = solvers[shape]
    p, grid, res = _rough(S, 3)
    p["start_ang"][1] = (0.03, -0.02, 0.2)
    x0, xl, xu, gl, gu = S.initial(p)
    so = oracle.default_shape(*SHAPES[shape])
    for i in range(3):
        po = oracle_problem(oracle, so, p[i], grid, res)
        assert (S.n_vars, S.n_cons) == (po.n, po.m)
        oxl, oxu, ogl, ogu = po.bounds()
        ox0 = po.x0(); ox0[oxl == oxu] = oxl[oxl == oxu]
        assert np.array_equal(xl[i], oxl) and np.array_equal(xu[i], oxu)
        assert np.array_equal(gl[i], ogl) and np.array_equal(gu[i], ogu)
        assert np.abs(x0[i] - ox0).max() < 1e-14
    if shape == "S5":
        assert (S.n_vars, S.n_cons, S.dims.n_free, S.dims.n_eq, S.dims.n_ineq) == (1040, 1730, 1005, 706, 1024)


@pytest.mark.parametrize("shape", ["S2", "S5"])
def test_constraints_and_jacobian_match_oracle(solvers, oracle, shape):
    S = solvers[shape]
    p, grid, res = _rough(S, 4, seed=7)
    p["start_ang"][:] = [(0.0, 0.0, 0.0), (0.05, -0.03, 0.3), (-0.1, 0.08, -0.4), (0.02, 0.02, 1.0)]
    so = oracle.default_shape(*SHAPES[shape])
    rng = np.random.default_rng(5)
    x0 = S.initial(p)[0]
    X = x0 + 0.05 * rng.standard_normal(x0.shape)
    g, J = S.eval(p, X)
    for i in range(4):
        po = oracle_problem(oracle, so, p[i], grid, res)
        oxl, oxu, _, _ = po.bounds()
        fixed = oxl == oxu
        x = X[i].copy(); x[fixed] = oxl[fixed]
        og, oJ = po.g(x), po.jac(x)
        oJ[:, fixed] = 0.0
        assert np.abs(g[i] - og).max() < G_TOL
        assert np.abs(J[i] - oJ).max() < J_TOL
        assert np.all(J[i][:, fixed] == 0.0)


def test_heightfield_queries_bit_exact(solvers, oracle, golden_hf):
    S = solvers["S2"]
    rng = np.random.default_rng(11)
    for name in ("exp_1", "exp_3", "exp_5"):
        grid, res = golden_hf[name + "_towr"], float(golden_hf[name + "_res"])
        hid = S.upload_heightfield(grid, res)
        ter = oracle.Terrain(grid, res)
        pts = np.concatenate([rng.uniform(-1.4, 5.0, (2000, 2)),
                              [[-1.0, -1.0], [0.0, 0.0], [-1.5, 0.2], [0.2, -1.5], [9.0, 9.0], [1e6, -1e6], [0.5, -1.0 + 7 * res], [-1.0 + 13 * res, 0.3]]])
        h = S.height(hid, pts)
        cells = S.height_cells(hid, pts)
        want = np.array([ter.height(x, y) for x, y in pts])
        wantc = np.array([ter.cell(x, y) for x, y in pts])[:, [0, 1, 2, 3]]
        assert np.array_equal(h, want)                      # bit-exact, no tolerance
        assert np.array_equal(cells, wantc)
        assert np.array_equal(h, HF.get_height(grid, res, pts[:, 0], pts[:, 1]))
        # the gradient query (the derivative the reference carries commented out, custom_terrain.cpp:96-156): bit-exact against the
        # oracle's restatement, and the slope of the height along each axis inside a cell
        hx, hy = S.heightfield_gradients(hid, pts)
        wg = np.array([ter.height_deriv(x, y) for x, y in pts])
        assert np.array_equal(hx, wg[:, 0]) and np.array_equal(hy, wg[:, 1])
        inner = pts[:2000]
        d = 1e-7
        same = np.all(S.height_cells(hid, inner) == S.height_cells(hid, inner + d), axis=1)
        fdx = (S.height(hid, inner + [d, 0]) - S.height(hid, inner)) / d
        fdy = (S.height(hid, inner + [0, d]) - S.height(hid, inner)) / d
        assert np.abs(fdx - hx[:2000])[same].max() < 1e-5 and np.abs(fdy - hy[:2000])[same].max() < 1e-5
    assert len(S.height(hid, np.zeros((0, 2)))) == 0        # empty query
    assert len(S.heightfield_gradients(hid, np.zeros((0, 2)))[0]) == 0


@pytest.mark.parametrize("alg", ["ipopt", "fast"])
@pytest.mark.parametrize("shape,n", [("S2", 24), ("S5", 8)])
def test_solve_matches_oracle(solvers, oracle, shape, n, alg):
    """Same algorithm on CPU and GPU: same status, same iteration count, node values and the 1 kHz trajectories
    within 1 mm.  On rough terrain the Ipopt path is sensitive to round-off in its last iterations (sigma_w -> 1e-8
    with zero terrain gradients in J; the oracle itself moves a window by 1e-3 m at the 90th percentile when its
    penalty parameter changes from 1e-5 to 1e-6), and the GPU factors with FP64 tensor-core
    block products where the oracle runs a scalar skyline Cholesky.  For that algorithm the bar is therefore: every
    window ends with the oracle's status, at least nine windows in ten take the oracle's iteration count AND stay
    within 1 mm of its plan, and the median deviation is below 1e-6 m (measured 2e-8 m)."""
    S = solvers[shape]
    p, grid, res = _rough(S, n)
    r, x, rows = S.solve(p, options=_opts(alg), csv=True)
    so = oracle.default_shape(*SHAPES[shape])
    devs, close = [], 0
    for i in range(n):
        po = oracle_problem(oracle, so, p[i], grid, res)
        xo, ro = _oracle_solve(po, alg)
        assert r["status"][i] == ro.status
        dev = np.abs(rows[i][:, 1:19] - po.csv(xo)[:, 1:19]).max()
        devs.append(dev)
        if alg == "fast":
            assert abs(int(r["iters"][i]) - ro.iters) <= 1
            assert dev < TRAJ_TOL_M, (i, dev)
        close += int(int(r["iters"][i]) == ro.iters and dev < TRAJ_TOL_M) if alg == "ipopt" else 1
        assert np.abs(rows[i] - po.csv(x[i])).max() < 1e-10          # sampler kernel vs oracle sampler
        if ro.status == 0:
            assert r["constr_viol"][i] <= 1e-4
            g = po.g(x[i]); _, _, gl, gu = po.bounds()
            assert np.maximum(gl - g, g - gu).max() <= 1e-4 + 1e-9   # re-checked by the oracle's own g(x)
    print("trajectory deviation vs oracle [m]: median %.2e p90 %.2e max %.2e, same iterations and within 1 mm: %d/%d" % (
        np.median(devs), np.percentile(devs, 90), max(devs), close, n))
    assert close >= 0.9 * n and np.median(devs) < 1e-6


def test_ipopt_reproduces_reference_log_and_plans(oracle, towr_log, golden_csv):
    """The reference's own Ipopt evidence: the three iteration tables of logs/towr_log.out:55-62,192-199,329-337 and the
    plans solves #1 and #2 wrote to data/traj/towr.csv.  Inputs = the logged command lines on the flat 600x200 grid at
    0.01 m; shape = the build that produced them (m = 3.0 kg, max_dev_x = 0.08; tests/test_ipopt_emulation.py).
    Iteration counts 7, 7, 8; iterations 0-4 print identically (inf_pr, inf_du, lg(mu), ||d||, alpha_du, alpha_pr,
    step tag, ls); the plans agree with TOWR + Ipopt to the CSV's 6 significant digits -- north_star asks 1 mm,
    measured 2.6e-6 m (CoM) and 5.6e-6 m (feet), asserted at 2e-5 m."""
    sh = Q.default_shape("Custom", 5.0, mass=3.0); sh.max_dev[0] = 0.08
    so = oracle.default_shape("Custom", 5.0, mass=3.0); so.max_dev[0] = 0.08
    S = Q.Solver(sh, max_batch=3)
    flat = np.zeros((600, 200))
    hid = S.upload_heightfield(flat, 0.01)
    p = Q.make_problems(3)
    for k, inp in enumerate(towr_log["inputs"]):
        for key in ("start_pos", "start_ang", "goal", "ee", "t_start"):
            p[k][key] = inp[key]
    p["hf_id"] = hid
    r, x, rows = S.solve(p, csv=True)
    tr = S.trace(3)
    assert list(r["status"]) == [0, 0, 0] and list(r["iters"]) == towr_log["iters"]
    for k, table in enumerate(towr_log["iteration_tables"]):
        assert np.all(tr[k, len(table):] == 0)
        for g in table:
            i = g["iter"]
            assert int(tr[k, i, 6]) == g["ls"] and (i == 0 or chr(int(tr[k, i, 7])) == g["tag"])
            assert abs(np.log10(tr[k, i, 2]) - float(g["lg_mu"])) <= 0.051 + (0.1 if i > 4 else 0.0)
            if i <= 4:
                for col, key in enumerate(("inf_pr", "inf_du", None, "dnorm", "alpha_du", "alpha_pr")):
                    if key:
                        assert "%.2e" % tr[k, i, col] == g[key] or abs(tr[k, i, col] - float(g[key])) <= 6e-3 * abs(float(g[key])), (k, i, key)
    for k, (G, row0) in enumerate(((golden_csv["towr_g4"], 2502), (golden_csv["towr_g2"], 0))):
        rk = rows[k][row0::10][:len(G)]
        assert np.allclose(rk[:, 0], G[:, 0], atol=1e-9)
        assert np.abs(rk[:, 1:4] - G[:, 1:4]).max() < 2e-5            # CoM
        assert np.abs(rk[:, 7:19] - G[:, 7:19]).max() < 2e-5          # feet
        assert np.abs(rk[:, 4:7] - G[:, 4:7]).max() < 1e-4            # base Euler angles
        assert np.abs(rk[1:, 25:37] - G[1:, 25:37]).max() < 5e-3      # forces [N]
        po = oracle_problem(oracle, so, p[k], flat, 0.01)
        xo, ro = po.solve_ipopt()
        assert ro.status == 0 and ro.iters == r["iters"][k]
        assert np.abs(rows[k][:, 1:19] - po.csv(xo)[:, 1:19]).max() < 1e-6
    S.close()


@pytest.mark.parametrize("alg", ["ipopt", "fast"])
def test_reference_experiment_windows(solvers, oracle, golden_hf, golden_csv, alg):
    """exp_1 (flat), exp_3 (0.5 m blocks), exp_5 (stairs): the single-window configs of BASELINE.json."""
    S = solvers["S5"]
    so = oracle.default_shape("Custom", 5.0)
    # exp_3: a 0.5 m block covers x in [0.7, 1.1), y in [-0.3, 0.1).  "exp_3" passes beside it; "exp_3_blocked" asks the
    # right front foot to end ON the block (infeasible range of motion) -- the case PATH_MAP probes for
    # (ref: QTOS/generateHeightField.py:387-404): both implementations must report a non-zero status.
    cases = {"exp_1": ((0.0, 0.0), (0.5, 0.0)), "exp_3": ((0.0, 0.35), (0.5, 0.35)), "exp_3_blocked": ((0.0, 0.0), (0.5, 0.0)),
             "exp_5": ((0.0, 0.1), (0.45, 0.1))}
    for name, (s, g) in cases.items():
        grid, res = golden_hf[name[:5] + "_towr"], float(golden_hf[name[:5] + "_res"])
        hid = S.upload_heightfield(grid, res)
        p = Q.make_problems(1)
        h0 = HF.get_height(grid, res, s[0], s[1])
        p["start_pos"][0] = (s[0], s[1], h0 + 0.24); p["goal"][0] = (g[0], g[1], 0.24); p["hf_id"] = hid
        p["ee"][0] = [(s[0] + a, s[1] + b, HF.get_height(grid, res, s[0] + a, s[1] + b)) for a, b, _ in FEET_19]
        r, x, rows = S.solve(p, options=_opts(alg), csv=True)
        po = oracle_problem(oracle, so, p[0], grid, res)
        xo, ro = _oracle_solve(po, alg)
        assert r["status"][0] == ro.status, name
        if r["status"][0] == 0:
            assert np.abs(rows[0][:, 1:19] - po.csv(xo)[:, 1:19]).max() < TRAJ_TOL_M, name
        if name in ("exp_1", "exp_3", "exp_5"):
            assert r["status"][0] == 0 and r["constr_viol"][0] <= 1e-4, name
        if name == "exp_3_blocked":
            assert r["status"][0] != 0 and towr_cli.exit_code(r["status"][0]) != 0
    # G3 inputs with the constants that produced the golden CSV (m = 3.0): feasible, and the measured gap
    # to Ipopt's own plan is bounded (a feasibility problem has no unique solution; DESIGN.md parity tiers)
    S3 = Q.Solver(Q.default_shape("Custom", 5.0, mass=3.0), max_batch=1)
    hid = S3.upload_heightfield(np.zeros((600, 200)), 0.01)
    p = Q.make_problems(1); p["goal"][0] = (0.502222, 0.0, 0.24); p["ee"][0] = FEET_19; p["hf_id"] = hid
    r, x, rows = S3.solve(p, options=_opts(alg), csv=True)
    assert r["status"][0] == 0 and r["constr_viol"][0] <= 1e-4 and rows.shape == (1, 5001, 37)
    gap = np.abs(rows[0][::10, 1:4] - golden_csv["gait"][:, 1:4]).max()
    print("CoM gap to Ipopt golden gait.csv [m]:", gap)
    assert gap < 0.10
    S3.close()


def test_batch_properties_at_bench_size(solvers):
    """Size-independent properties on a 256-problem batch of the bench workload."""
    S = solvers["S2"]
    p, grid, res = _rough(S, 256)
    r, x, _ = S.solve(p)
    ok = r["status"] == 0
    assert ok.mean() >= 0.97
    assert np.all(r["constr_viol"][ok] <= 1e-4)
    assert set(np.unique(r["status"])) <= {0, -1, -2}       # every problem reports a status, none dropped
    x0, xl, xu, gl, gu = S.initial(p)
    fixed = xl == xu
    assert np.array_equal(x[fixed], xl[fixed])              # start / goal bounds are held exactly
    g = S.eval(p, x, jac=False)
    viol = np.maximum(np.maximum(gl - g, g - gu), 0).max(axis=1)
    assert np.all(viol[ok] <= 1e-4 + 1e-9)                  # violation recomputed from scratch by the eval entry
    # solving the same batch twice is deterministic (owner-computes assembly, no atomics in the data path)
    r2, x2, _ = S.solve(p)
    assert np.array_equal(x, x2) and np.array_equal(r["iters"], r2["iters"])
    # order independence: a permuted batch gives the permuted answer
    perm = np.random.default_rng(0).permutation(256)
    r3, x3, _ = S.solve(p[perm])
    assert np.array_equal(x3, x[perm])
    # stance feet sit on the terrain: CSV foot z equals the heightfield under the foot at contact samples
    rows = S.sample_csv(p[:4], x[:4])
    assert rows.shape == (4, 2001, 37) and np.allclose(rows[:, 0, 1:4], p["start_pos"][:4])
    last = rows[:, -1]
    for i in range(4):
        if not ok[i]:
            continue
        for e in range(4):
            fx, fy, fz = last[i, 7 + 3 * e:10 + 3 * e]
            assert abs(fz - HF.get_height(grid, res, fx, fy)) < 2e-4


def test_errors_are_loud(solvers):
    S = solvers["S2"]
    p = Q.make_problems(1); p["hf_id"] = 12345
    with pytest.raises(Q.QtosError, match="heightfield"):
        S.solve(p)
    with pytest.raises(Q.QtosError):
        Q.Solver(Q.default_shape("C1", -1.0))
    bad = Q.default_shape("C1", 2.0); bad.combo = 17
    with pytest.raises(Q.QtosError, match="gait"):
        Q.Solver(bad)


def test_cli_clone_writes_reference_csv(tmp_path, golden_hf):
    """`./main <flags>` drop-in: reads ../data/heightfields/from_pybullet/towr_heightfield.txt, writes traj.csv."""
    build = tmp_path / "towr" / "build"; build.mkdir(parents=True)
    hfdir = tmp_path / "towr" / "data" / "heightfields" / "from_pybullet"; hfdir.mkdir(parents=True)
    HF.write_heightfield(str(hfdir / "towr_heightfield.txt"), golden_hf["exp_1_towr"])
    args = {"-s": [0, 0, 0.24], "-g": [0.5, 0, 0.24], "-e1": [0.21, 0.19, 0.0], "-e2": [0.21, -0.19, 0.0],
            "-e3": [-0.21, 0.19, 0.0], "-e4": [-0.21, -0.19, 0.0], "-s_ang": [0, 0, 0], "-t": 1.25, "-r": 15.0, "-resolution": 0.1}
    rc = towr_cli.towr_main(towr_cli.cmd_args(args).split(), cwd=str(build), quiet=True)
    assert rc == 0
    text = open(build / "traj.csv").read().strip().split("\n")
    assert len(text) == 5001 and all(len(l.split(",")) == 37 for l in text[:50])
    first = [float(v) for v in text[0].split(",")]
    assert first[0] == 1.25 and first[1:4] == [0.0, 0.0, 0.24] and first[7:10] == [0.21, 0.19, 0.0]
    assert float(text[-1].split(",")[0]) == 6.25
    assert towr_cli.towr_main(["-g", "0.5", "0", "0.24"], cwd=str(tmp_path), quiet=True) == 2      # missing heightfield


def test_replan_sweep_over_terrain_variants(solvers, oracle):
    """BASELINE config 5 in small: terrain variants as separate device-resident grids inside ONE batch, then the
    receding-horizon successors started from each plan's own state (contact-phase row), best plan per group of candidates."""
    from qtos_b200 import parallel
    S = solvers["S2"]
    variants = workloads.terrain_variants(3)
    probs = []
    for v, (grid, res) in enumerate(variants):
        hid = S.upload_heightfield(grid, res)
        probs.append(workloads.multistart_problems(8, grid, res, seed=100 + v, hf_id=hid, group_size=4))
        probs[-1]["group"] += 2 * v
    p = np.concatenate(probs)
    r, x, rows = S.solve(p, csv=True)
    assert (r["status"] == 0).mean() >= 0.9
    so = oracle.default_shape(*SHAPES["S2"])
    for i in (0, 9, 17, 23):                                  # one window per variant (+1) against the oracle on ITS grid
        grid, res = variants[i // 8]
        po = oracle_problem(oracle, so, p[i], grid, res)
        xo, ro = po.solve_ipopt()
        assert r["status"][i] == ro.status
        if ro.iters == r["iters"][i]:
            assert np.abs(rows[i][:, 1:19] - po.csv(xo)[:, 1:19]).max() < TRAJ_TOL_M
    # a variant's plan must not depend on which other grids share the batch
    r1, x1, _ = S.solve(p[8:16])
    assert np.array_equal(x1, x[8:16])
    # successors: all four feet are in stance at the last row of a trot window (ref: combiner.check_legs_contact)
    nxt = workloads.replan_problems(p, rows, rows.shape[1] - 1)
    assert np.allclose(nxt["t_start"], 2.0) and np.all(nxt["start_vel"] == 0)
    r2, x2, rows2 = S.solve(nxt, csv=True)
    ok = (r["status"] == 0) & (r2["status"] == 0)
    assert ok.mean() >= 0.75
    assert np.allclose(rows2[ok][:, 0, 1:7], rows[ok][:, -1, 1:7], atol=1e-12)       # CoM pose is continuous across the hand-over
    assert np.allclose(rows2[ok][:, 0, 7:19], rows[ok][:, -1, 7:19], atol=1e-12)     # and so are the footholds
    assert np.allclose(rows2[:, 0, 0], 2.0) and np.allclose(rows2[:, -1, 0], 4.0)
    i = int(np.flatnonzero(ok)[0])
    grid, res = variants[i // 8]
    xo, ro = oracle_problem(oracle, so, nxt[i], grid, res).solve_ipopt()
    assert ro.status == r2["status"][i]
    # best plan per group of 4 candidates: deterministic key, one winner per group, winners are converged when any is
    winners, _ = parallel.select_best(parallel.make_records(r2, np.arange(len(p)), nxt["group"]))
    assert sorted(winners) == list(range(6))
    for g, w in winners.items():
        members = np.flatnonzero(nxt["group"] == g)
        assert w in members
        if (r2["status"][members] == 0).any():
            assert r2["status"][w] == 0 and r2["cost"][w] == r2["cost"][members][r2["status"][members] == 0].min()


def test_chunking_and_argument_errors(solvers):
    """n > max_batch is solved in chunks with identical results; bad calls fail loudly."""
    S = solvers["S2"]
    p, grid, res = _rough(S, 20)
    r, x, _ = S.solve(p)
    S8 = Q.Solver(Q.default_shape(*SHAPES["S2"]), max_batch=8)
    p8 = p.copy(); p8["hf_id"] = S8.upload_heightfield(grid, res)
    r8, x8, rows8 = S8.solve(p8, csv=True)
    assert np.array_equal(x8, x) and np.array_equal(r8["iters"], r["iters"]) and rows8.shape == (20, 2001, 37)
    with pytest.raises(Q.QtosError):
        S8.solve(p8[:0])
    S8.close()


def test_bad_window_does_not_poison_the_batch(solvers):
    """independent windows stay independent: a NaN start and an absurd goal end with a non-zero status inside the
    iteration cap while every other window of the batch returns exactly what it returns alone."""
    S = solvers["S2"]
    p, grid, res = _rough(S, 16)
    r0, x0, _ = S.solve(p)
    q = p.copy()
    q["start_pos"][3, 0] = np.nan
    q["goal"][7] = (1e6, -1e6, 0.24)
    r, x, _ = S.solve(q)
    assert r["status"][3] != 0 and r["status"][7] != 0 and r["iters"].max() <= 200
    keep = np.ones(16, bool); keep[[3, 7]] = False
    assert np.array_equal(x[keep], x0[keep]) and np.array_equal(r["status"][keep], r0["status"][keep])


def test_contexts_of_different_shapes_coexist():
    """the dynamic shared-memory cap is per kernel function, not per context: a narrow shape created after a wide one
    (and the other way round) must still launch."""
    for order in (("S5", "S2"), ("S2", "S5")):
        ctxs = [Q.Solver(Q.default_shape(*SHAPES[k]), max_batch=2) for k in order]
        for S in ctxs:
            p, _, _ = _rough(S, 2)
            r, _, _ = S.solve(p)
            assert set(r["status"]) <= {0, -1, -2}
        for S in ctxs:
            S.close()


def test_caller_supplied_output_buffers(solvers):
    """solve(out=...) writes into the caller's (e.g. page-locked) buffers and returns them; wrong shapes are refused."""
    import torch
    S = solvers["S2"]
    p, _, _ = _rough(S, 8)
    r0, x0, _ = S.solve(p)
    h_res = torch.empty(8 * Q.RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    h_x = torch.empty((8, S.n_vars), dtype=torch.float64).pin_memory()
    res, x = h_res.numpy().view(Q.RESULT_DTYPE).reshape(8), h_x.numpy()
    r1, x1, _ = S.solve(p, out=(res, x))
    assert r1 is res and x1 is x and np.array_equal(x1, x0) and np.array_equal(r1["iters"], r0["iters"])
    with pytest.raises(ValueError):
        S.solve(p, out=(res, x[:4]))


def test_c0_duration_mode_matches_oracle(oracle):
    """`-duration 2.0` selects gait combo C0 (ref: main.cpp:299-306,425-427; the `-t` mode of scripts/main.py:119-121):
    the shape is compiled, solved and compared with the oracle like the production shapes (parsed only)."""
    a = towr_cli.parse_main_argv(["-g", "0.4", "0.05", "0.24", "-s", "0", "0", "0.24", "-duration", "2.0"])
    assert a["combo"] == "C0" and a["duration"] == 2.0
    S = Q.Solver(Q.default_shape("C0", 2.0), max_batch=8)
    grid, res = HF.rough_terrain(5)
    hid = S.upload_heightfield(grid, res)
    p = workloads.multistart_problems(8, grid, res, seed=5, hf_id=hid)
    r, x, rows = S.solve(p, csv=True)
    so = oracle.default_shape("C0", 2.0)
    x0, xl, xu, gl, gu = S.initial(p[:1])
    po = oracle_problem(oracle, so, p[0], grid, res)
    oxl, oxu, ogl, ogu = po.bounds()
    assert np.array_equal(gl[0], ogl) and np.array_equal(gu[0], ogu) and len(ogl) == S.n_cons and po.n == S.n_vars
    close = 0
    for i in range(8):
        po = oracle_problem(oracle, so, p[i], grid, res)
        xo, ro = po.solve_ipopt()
        assert r["status"][i] == ro.status
        close += int(r["iters"][i] == ro.iters and np.abs(rows[i][:, 1:19] - po.csv(xo)[:, 1:19]).max() < TRAJ_TOL_M)
    assert (r["status"] == 0).sum() >= 7 and close >= 7
    S.close()


def test_heightfield_release_and_device_side_validation(solvers):
    """qtos_free_heightfield releases a grid and its id is reused; a device-resident problem that names a released or
    unknown id ends with Invalid_Number_Detected (-13) instead of being solved on grid 0 (VERDICT r1 robustness)."""
    import torch
    S = Q.Solver(Q.default_shape(*SHAPES["S2"]), max_batch=8)
    grid, res = HF.rough_terrain(1234)
    a = S.upload_heightfield(grid, res)
    b = S.upload_heightfield(np.zeros((64, 64)), 0.1)
    assert (a, b) == (0, 1)
    p = workloads.multistart_problems(4, grid, res, hf_id=a)
    r0, x0, _ = S.solve(p)
    S.free_heightfield(b)
    with pytest.raises(Q.QtosError):
        S.free_heightfield(b)
    q = p.copy(); q["hf_id"][1] = b
    with pytest.raises(Q.QtosError, match="heightfield"):
        S.solve(q)                                            # host path: refused before any launch
    q["hf_id"][2] = 77
    d_p = torch.from_numpy(q.view(np.uint8).reshape(4, -1)).cuda()
    d_res = torch.zeros((4, Q.RESULT_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
    d_x = torch.zeros((4, S.n_vars), dtype=torch.float64, device="cuda")
    S.solve_device(d_p.data_ptr(), 4, None, d_res.data_ptr(), d_x.data_ptr())
    r = d_res.cpu().numpy().view(Q.RESULT_DTYPE).reshape(4)
    assert list(r["status"]) == [r0["status"][0], -13, -13, r0["status"][3]] and np.isnan(r["constr_viol"][1])
    assert np.array_equal(d_x.cpu().numpy()[[0, 3]], x0[[0, 3]])     # the good windows are untouched by their neighbours
    c = S.upload_heightfield(grid, res)
    assert c == b                                              # the released id is handed out again
    S.close()


def test_runtime_budget_and_async_calls(solvers):
    """-r / max_cpu_time is an iteration budget (0.1 s of the reference per iteration) and reports Ipopt's
    Maximum_CpuTime_Exceeded (-4); the asynchronous entry points return the synchronous results."""
    S = solvers["S2"]
    p, grid, res = _rough(S, 16)
    r0, x0, _ = S.solve(p)
    r, x, _ = S.solve(p, Q.default_options(max_cpu_time=0.45))          # 4 iterations
    assert set(r["status"]) == {-4} and r["iters"].max() == 4 and towr_cli.exit_code(-4) == 252
    r, x, _ = S.solve(p, Q.default_options(max_cpu_time=15.0))          # QTOS's default budget: 150 iterations
    assert np.array_equal(r["status"], r0["status"]) and np.array_equal(x, x0)
    res_buf, x_buf = np.zeros(16, dtype=Q.RESULT_DTYPE), np.zeros((16, S.n_vars))
    S.solve_async(p, (res_buf, x_buf))
    with pytest.raises(Q.QtosError, match="in flight"):
        S.solve_async(p, (res_buf, x_buf))
    ra, xa = S.wait()
    assert ra is res_buf and np.array_equal(xa, x0) and np.array_equal(ra["iters"], r0["iters"])
    S.wait()                                                            # nothing in flight: returns at once
    rows = S.sample_rows(p[:3], x0[:3], -1)
    full = S.sample_csv(p[:3], x0[:3])
    assert rows.shape == (3, 1, 37) and np.array_equal(rows[:, 0], full[:, -1])
    assert np.array_equal(S.sample_rows(p[:3], x0[:3], 750, 2), full[:, 750:752])


def test_continuous_batching_returns_the_batch_call_results():
    """Streaming session: jobs flow through a pool of 64 slots that is smaller than the work queued (48 + 64 + 20 + 64
    windows, all submitted at once); every window's plan, status and iteration count is BIT-IDENTICAL to what
    qtos_solve_batch returns for it, whatever company it kept in the pool; the pool stays full while work is waiting."""
    import torch
    S = Q.Solver(Q.default_shape(*SHAPES["S2"]), max_batch=64)
    grid, res = HF.rough_terrain(1234)
    hid = S.upload_heightfield(grid, res)
    p = workloads.multistart_problems(196, grid, res, hf_id=hid)
    r0, x0, _ = S.solve(p)
    S.stream_begin()
    with pytest.raises(Q.QtosError, match="stream"):
        S.solve(p[:4])                                              # the context is busy streaming
    cuts = [(0, 48), (48, 112), (112, 132)]
    tickets = [S.stream_submit(p[a:b]) for a, b in cuts]
    d_p = torch.from_numpy(p[132:196].view(np.uint8).reshape(64, -1)).cuda()
    d_res = torch.zeros((64, Q.RESULT_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
    d_x = torch.zeros((64, S.n_vars), dtype=torch.float64, device="cuda")
    td = S.stream_submit_device(d_p.data_ptr(), 64, d_res.data_ptr(), d_x.data_ptr())
    for t, (a, b) in reversed(list(zip(tickets, cuts))):           # waiting out of order is fine
        r, x = S.stream_wait(t)
        assert np.array_equal(x, x0[a:b]) and np.array_equal(r["status"], r0["status"][a:b]) and np.array_equal(r["iters"], r0["iters"][a:b])
        assert np.array_equal(r["cost"], r0["cost"][a:b]) and np.array_equal(r["constr_viol"], r0["constr_viol"][a:b])
    S.stream_wait(td)
    rd = d_res.cpu().numpy().view(Q.RESULT_DTYPE).reshape(64)
    assert np.array_equal(d_x.cpu().numpy(), x0[132:]) and np.array_equal(rd["iters"], r0["iters"][132:])
    info = S.stream_info()
    assert info["windows_done"] == 196 and info["slot_iterations"] >= int(r0["iters"].sum())
    # the tail of one job rides with the next: fewer batch iterations than the four jobs need one after the other
    serial = sum(int(r0["iters"][a:b].max()) + 1 for a, b in cuts + [(132, 196)])
    assert info["iterations"] < serial, (info, serial)
    t2 = S.stream_submit(p[:8])                                     # the session outlives an idle period
    r, x = S.stream_wait(t2)
    assert np.array_equal(x, x0[:8])
    # a job can also deliver the reference's actual output, the 1 kHz rows of every plan (qtos_stream_submit_csv): identical to the
    # rows the sampler entry point returns for the same plans; pinned or pageable destination
    rows_ref = S.sample_csv(p[:40], x0[:40])
    h_rows = torch.empty((40, S.csv_rows, Q.CSV_COLS), dtype=torch.float64).pin_memory()
    t3 = S.stream_submit(p[:40], csv_out=h_rows.numpy())
    pageable = np.zeros((12, S.csv_rows, Q.CSV_COLS))
    t4 = S.stream_submit(p[40:52], csv_out=pageable)
    r, x = S.stream_wait(t3); S.stream_wait(t4)
    assert np.array_equal(x, x0[:40])
    S.stream_end()
    _r, _x, rows_sync = S.solve(p[:52], csv=True)
    assert np.array_equal(h_rows.numpy(), rows_sync[:40]) and np.array_equal(pageable, rows_sync[40:52])
    assert np.array_equal(rows_ref, rows_sync[:40])
    with pytest.raises(ValueError):
        S.stream_begin(); S.stream_submit(p[:4], csv_out=np.zeros((4, 7, Q.CSV_COLS)))
    S.stream_end()
    r1, x1, _ = S.solve(p[:16])                                     # and the context is a batch solver again
    assert np.array_equal(x1, x0[:16])
    S.close()
    # a job larger than one copy chunk (256 windows): the rows of the windows either side of the chunk boundary
    Sb = Q.Solver(Q.default_shape(*SHAPES["S2"]), max_batch=300)
    hid = Sb.upload_heightfield(grid, res)
    pb = workloads.multistart_problems(300, grid, res, hf_id=hid)
    big = np.zeros((300, Sb.csv_rows, Q.CSV_COLS))
    Sb.stream_begin()
    rb, xb = Sb.stream_wait(Sb.stream_submit(pb, csv_out=big))
    Sb.stream_end()
    pick = [0, 255, 256, 299]
    assert np.array_equal(big[pick], Sb.sample_csv(pb[pick], xb[pick])) and np.isfinite(big).all()
    Sb.close()


def test_second_attempt_rescues_failed_line_searches(solvers, oracle):
    """qtos_options.retry_failed: a window whose filter line search fails (Ipopt would enter its restoration phase) is restarted
    once with limited_memory_init_val_min = 1e-2.  The first 1024 windows of the bench workload hold a few such windows
    (90 and 139 in the oracle; which ones fail is itself sensitive to round-off); with the second attempt they converge, report
    the iterations of both attempts, and every other window of the batch is untouched bit for bit.  The oracle makes the
    same second attempt and converges on them too."""
    S = solvers["S2"]
    p, grid, res = _rough(S, 1024)                            # four chunks of the fixture's 256 slots
    r0, x0, _ = S.solve(p, Q.default_options(retry_failed=0))
    r1, x1, _ = S.solve(p)
    failed = np.flatnonzero(r0["status"] == -2)
    print("windows that needed the second attempt:", failed)
    assert 1 <= len(failed) <= 16
    assert np.all(r1["status"][failed] == 0) and np.all(r1["constr_viol"][failed] <= 1e-4)
    assert np.all(r1["iters"][failed] > r0["iters"][failed])
    keep = np.ones(1024, bool); keep[failed] = False
    assert np.array_equal(x1[keep], x0[keep]) and np.array_equal(r1["iters"][keep], r0["iters"][keep])
    so = oracle.default_shape(*SHAPES["S2"])
    for i in failed:
        po = oracle_problem(oracle, so, p[i], grid, res)
        xo, ro = po.solve_ipopt()
        assert ro.status == 0


def test_device_side_selection_equals_the_sorted_reference():
    """k_records + k_select (the hand-written selection kernels) against parallel.argmin_per_group_sorted on records with ties
    on every key, non-converged groups, NaN metrics (status -13) and an empty group."""
    import torch
    from qtos_b200 import parallel
    S = Q.Solver(Q.default_shape(*SHAPES["S2"]), max_batch=1)
    rng = np.random.default_rng(3)
    n, n_groups = 5000, 700
    r = np.zeros(n, dtype=Q.RESULT_DTYPE)
    r["status"] = rng.choice([0, 0, 0, -1, -2, -13], n)
    r["cost"] = np.round(rng.uniform(0, 3, n), 1)                       # many exact ties
    r["constr_viol"] = np.round(rng.uniform(0, 1e-4, n), 5)
    bad = r["status"] == -13
    r["cost"][bad] = np.nan; r["constr_viol"][bad] = np.nan
    group = rng.integers(0, n_groups - 1, n).astype(np.int32)           # group n_groups - 1 stays empty
    group[:40] = 5; r["status"][:40] = -1                                # a group where nobody converged
    d_res = torch.from_numpy(r.view(np.uint8).reshape(n, -1)).cuda()
    d_group = torch.from_numpy(group).cuda()
    win = parallel.select_best_device(S, d_res, d_group, rank=3, world=8, n_groups=n_groups).cpu().numpy()
    want = parallel.argmin_per_group_sorted(parallel.make_records(r, 3 + 8 * np.arange(n), group))
    assert win[n_groups - 1] == -1 and len(want) == n_groups - 1
    for g, w in want.items():
        assert win[g] == w, (g, win[g], w)
    S.close()


def test_base_motion_constraint_option(oracle):
    """qtos_shape.base_rom: TOWR's optional BaseMotionConstraint (Parameters::BaseRom, base_motion_constraint.cc:38-93; SURVEY 8f rank 3,
    the part that does not need duration variables).  82 samples x 6 rows are appended after the swing sets; values, bounds and
    Jacobian equal the oracle's (the z row is stated relative to the fixed start height, so value and bounds shift together),
    and the solves agree with the oracle's Ipopt port."""
    sh = Q.default_shape("C1", 2.0); sh.base_rom = 1
    so = oracle.default_shape("C1", 2.0); so.base_rom = 1
    S = Q.Solver(sh, max_batch=8)
    assert S.n_cons == 892 + 82 * 6
    grid, res = HF.rough_terrain(7)
    hid = S.upload_heightfield(grid, res)
    p = workloads.multistart_problems(8, grid, res, seed=7, hf_id=hid)
    x0, xl, xu, gl, gu = S.initial(p)
    rng = np.random.default_rng(0)
    for i in range(2):
        po = oracle_problem(oracle, so, p[i], grid, res)
        oxl, oxu, ogl, ogu = po.bounds()
        x = x0[i] + 0.02 * rng.standard_normal(S.n_vars); x[oxl == oxu] = oxl[oxl == oxu]
        g, J = S.eval(p[i:i + 1], x[None])
        og, oJ = po.g(x), po.jac(x); oJ[:, oxl == oxu] = 0
        shift = np.zeros(S.n_cons); shift[892 + 5::6] = p["start_pos"][i, 2]          # the LZ rows
        assert np.abs(g[0] + shift - og).max() < G_TOL and np.abs(J[0] - oJ).max() < J_TOL
        assert np.allclose(np.clip(gl[i], -1e20, 1e20) + shift, ogl, atol=1e-15) and np.allclose(np.clip(gu[i], -1e20, 1e20) + shift, ogu, atol=1e-15)
    r, x, rows = S.solve(p, csv=True)
    close = 0
    for i in range(8):
        po = oracle_problem(oracle, so, p[i], grid, res)
        xo, ro = po.solve_ipopt()
        assert r["status"][i] == ro.status
        if ro.status == 0:
            assert r["constr_viol"][i] <= 1e-4
            roll_pitch = np.abs(rows[i][::25, 4:6]).max(); dz = rows[i][::25, 3] - p["start_pos"][i, 2]
            assert roll_pitch <= 0.01 + 1e-4 and dz.min() >= -0.02 - 1e-4 and dz.max() <= 0.1 + 1e-4      # the constraint holds at its samples
        close += int(r["iters"][i] == ro.iters and np.abs(rows[i][:, 1:19] - po.csv(xo)[:, 1:19]).max() < TRAJ_TOL_M)
    assert (r["status"] == 0).sum() >= 6 and close >= 6
    S.close()


def test_more_windows_than_slots_flow_through_the_pool():
    """qtos_solve_batch with n > max_batch runs the windows through the slots as a pool (not chunk after chunk); per-window
    results equal the one-batch solve bit for bit, CSV sampling included."""
    grid, res = HF.rough_terrain(1234)
    big = Q.Solver(Q.default_shape(*SHAPES["S2"]), max_batch=96)
    small = Q.Solver(Q.default_shape(*SHAPES["S2"]), max_batch=20)
    p = workloads.multistart_problems(96, grid, res, hf_id=big.upload_heightfield(grid, res))
    assert small.upload_heightfield(grid, res) == p["hf_id"][0]
    r0, x0, _ = big.solve(p)
    r1, x1, rows = small.solve(p, csv=True)                       # 96 windows through 20 slots
    assert np.array_equal(x1, x0) and np.array_equal(r1["iters"], r0["iters"]) and np.array_equal(r1["status"], r0["status"])
    assert np.array_equal(r1["cost"], r0["cost"]) and rows.shape == (96, 2001, 37)
    assert np.array_equal(rows[5], big.sample_csv(p[5:6], x0[5:6])[0])
    r2, x2, _ = small.solve(p, Q.default_options(algorithm=Q.ALG_FAST))      # the FAST algorithm still goes chunk by chunk
    assert set(r2["status"]) <= {0, -1, -2}
    big.close(); small.close()


def test_cost_terms_option(oracle):
    """qtos_shape.cost_force_z / cost_ee_vel_xy: TOWR's optional NodeCost terms (Parameters::costs_, nlp_formulation.cc:343-376,
    node_cost.cc:53-83) as the objective of QTOS_ALG_IPOPT.  The reference logs no run with an objective, so the
    chain is emulator (pinned to the logged f == 0 tables) -> C oracle (tests/test_cost_terms.py) -> kernels (here): the GPU
    prints the C oracle's iteration table for the first iterations, every converged plan is feasible, and where both converge the
    objectives agree to within 1e-3 of the initial objective (north_star: objective within 1e-3 relative).  With an objective the
    algorithm needs ~100 iterations and its limited-memory steps reach ||d|| ~ 1e3, so late iterations amplify round-off: in
    the oracle alone 1-3 of 8 such windows end at the iteration limit or in a failed line search, and WHICH ones is not stable
    against last-bit changes -- the status comparison is therefore a count, not window by window."""
    sh = Q.default_shape("C1", 2.0); sh.cost_force_z = 1.0
    so = oracle.default_shape("C1", 2.0); so.cost_force_z = 1.0
    S = Q.Solver(sh, max_batch=8)
    grid, res = HF.rough_terrain(11)
    flat = np.zeros((40, 20))
    hf_rough, hf_flat = S.upload_heightfield(grid, res), S.upload_heightfield(flat, 0.1)
    p = workloads.multistart_problems(8, grid, res, seed=11, hf_id=hf_rough)
    p["start_pos"][0] = (0, 0, 0.24); p["goal"][0] = (0.5, 0, 0.24); p["ee"][0] = [(0.21, 0.18, 0), (0.21, -0.18, 0), (-0.21, 0.18, 0), (-0.21, -0.18, 0)]
    p["hf_id"][0] = hf_flat
    r, x, _ = S.solve(p)
    tr = S.trace(8)
    fmt = lambda a: "%.2e %.2e %5.1f %.2e %.2e %.2e%s %d" % (a[0], a[1], np.log10(a[2]), a[3], a[4], a[5], chr(int(a[7])) if a[7] else " ", int(a[6]))
    same_status = 0
    for i in range(8):
        po = oracle_problem(oracle, so, p[i], flat if i == 0 else grid, 0.1 if i == 0 else res)
        f0 = po.cost(po.x0())
        xo, ro = po.solve_ipopt()
        for k in range(min(6, ro.n_trace)):
            o_row = (ro.tr_inf_pr[k], ro.tr_inf_du[k], ro.tr_mu[k], ro.tr_dnorm[k], ro.tr_alpha_du[k], ro.tr_alpha_pr[k], ro.tr_ls[k], ord(ro.tr_tag[k:k + 1].decode()) if ro.tr_tag[k:k + 1] != b" " else 0)
            assert fmt(tr[i, k]) == fmt(o_row), (i, k, fmt(tr[i, k]), fmt(o_row))
        assert abs(tr[i, 0, 1] - 2 * 1.5 * 9.80665 / 4) < 1e-12             # iteration 0: inf_du = ||grad f(x0)||_inf
        same_status += int(r["status"][i] == ro.status)
        if r["status"][i] == 0:
            assert r["constr_viol"][i] <= 1e-4
            if ro.status == 0:
                assert abs(po.cost(x[i]) - ro.objective) <= 1e-3 * f0, (i, po.cost(x[i]), ro.objective, f0)
    print("cost terms: statuses", r["status"], "iters", r["iters"], "same status as the oracle:", same_status, "of 8")
    assert same_status >= 4 and (r["status"] == 0).sum() >= 4
    with pytest.raises(Exception):
        S.solve(p, Q.default_options(algorithm=Q.ALG_FAST))      # the FAST algorithm assumes f == 0
    S.close()


def test_terrain_gradients_option(oracle):
    """qtos_shape.terrain_gradients: the terrain's first derivatives as the bilinear surface gives them -- the code
    the reference carries commented out (custom_terrain.cpp:96-156) -- in the terrain rows' Jacobian (terrain_constraint.cc:90-108) and
    in the contact basis of the force rows (height_map.cc:95-141, force_constraint.cc:67-135).  Values and Jacobian equal the oracle's
    on the bench's plateau terrain; on a smooth hill the solves follow the oracle's Ipopt port like the default path does; on the
    plateaus the option is the regression the oracle measured (the derivative is zero almost everywhere and huge on the 2 cm ramps), and
    the GPU reports it the same way."""
    sh = Q.default_shape("C1", 2.0); sh.terrain_gradients = 1
    so = oracle.default_shape("C1", 2.0); so.terrain_gradients = 1
    S = Q.Solver(sh, max_batch=16)
    assert S.n_vars == 640 and S.n_cons == 892
    rough, res = HF.rough_terrain(5)
    gx, gy = np.meshgrid(np.arange(256) * 0.02 - 1.0, np.arange(256) * 0.02 - 1.0, indexing="ij")
    hill = 0.04 + 0.03 * np.sin(1.7 * gx) * np.cos(1.3 * gy)
    h_rough, h_hill = S.upload_heightfield(rough, res), S.upload_heightfield(hill, 0.02)
    rng = np.random.default_rng(2)
    for grid, hid in ((rough, h_rough), (hill, h_hill)):
        p = workloads.multistart_problems(4, grid, 0.02, seed=5, hf_id=hid)
        x0, xl, xu, gl, gu = S.initial(p)
        for i in range(2):
            po = oracle_problem(oracle, so, p[i], grid, 0.02)
            oxl, oxu, ogl, ogu = po.bounds()
            assert np.array_equal(np.clip(gl[i], -1e20, 1e20), ogl) and np.array_equal(np.clip(gu[i], -1e20, 1e20), ogu)
            x = x0[i] + 0.03 * rng.standard_normal(S.n_vars); x[oxl == oxu] = oxl[oxl == oxu]
            g, J = S.eval(p[i:i + 1], x[None])
            og, oJ = po.g(x), po.jac(x); oJ[:, oxl == oxu] = 0
            assert np.abs(g[0] - og).max() < G_TOL and np.abs(J[0] - oJ).max() < 1e-9, (np.abs(g[0] - og).max(), np.abs(J[0] - oJ).max())
            _, ro_ = po.layout()
            assert np.abs(oJ[ro_[0]:ro_[4]]).sum(axis=1).max() > 1.0 or grid is rough      # the hill's terrain rows do see dh/dx, dh/dy
    # solves on the smooth hill: status, iteration count and plan follow the oracle
    p = workloads.multistart_problems(12, hill, 0.02, seed=9, hf_id=h_hill)
    r, x, rows = S.solve(p, csv=True)
    close = 0
    for i in range(12):
        po = oracle_problem(oracle, so, p[i], hill, 0.02)
        xo, ro = po.solve_ipopt()
        assert r["status"][i] == ro.status, (i, r["status"][i], ro.status)
        if ro.status == 0:
            assert r["constr_viol"][i] <= 1e-4
            close += int(r["iters"][i] == ro.iters and np.abs(rows[i][:, 1:19] - po.csv(xo)[:, 1:19]).max() < TRAJ_TOL_M)
    print("terrain gradients, smooth hill: statuses", r["status"], "iters", r["iters"], "same iterations and within 1 mm:", close, "of 12")
    assert (r["status"] == 0).sum() >= 10 and close >= 9
    # plateaus: the regression, reported like the oracle reports it
    p = workloads.multistart_problems(12, rough, 0.02, seed=5, hf_id=h_rough)
    r, x, _ = S.solve(p)
    same = sum(int(r["status"][i] == oracle_problem(oracle, so, p[i], rough, 0.02).solve_ipopt()[1].status) for i in range(12))
    print("terrain gradients, plateaus: statuses", r["status"], "same status as the oracle:", same, "of 12")
    assert same >= 8
    S.close()
