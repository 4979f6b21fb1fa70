"""SURVEY 8(f) rank 4, cost terms: TOWR's optional NodeCost terms (Parameters::costs_, empty on the reference's path;
ref: solver/towr/src/nlp_formulation.cc:343-376, node_cost.cc:53-83) as an objective of the Ipopt algorithm -- oracle level.

The reference holds no golden data with cost terms (its logged runs have f == 0), so the chain of evidence is: the dense
emulator (oracle/ipopt_emul.py) is pinned to the reference's logged iteration tables with f == 0; the objective enters it
exactly where Ipopt's published algorithm has f (gradient of the Lagrangian, limited-memory pairs, barrier function and its
directional derivative in the filter line search, the (f, theta) filter of the adaptive barrier update, objective scaling);
the C restatement (oracle/towr_ipopt.c) must then print the emulator's iteration table; the GPU must print the C oracle's
(tests/test_gpu_parity.py::test_cost_terms_option)."""
import numpy as np
import pytest

from oracle import ipopt_emul as E


def _problem(O, wf, wv, combo="C1", T=2.0):
    sh = O.default_shape(combo, T)
    sh.cost_force_z, sh.cost_ee_vel_xy = wf, wv
    return O.Problem(sh, O.make_instance(), O.Terrain(np.zeros((40, 20)), 0.1))


def test_cost_is_the_sum_over_nodes_and_its_gradient_is_exact(oracle):
    P = _problem(oracle, 0.7, 0.3)
    vo, _ = P.layout()
    rng = np.random.default_rng(3)
    x = P.x0() + 0.1 * rng.standard_normal(P.n)
    f, g = P.cost(x, with_grad=True)
    # independent statement: force sets hold (pos, vel) x (x, y, z) per optimised node -> f_z is every 6th entry from offset 4;
    # foot-motion swing nodes hold x, vx, y, vy, z (5 entries), stance nodes x, y, z (3 entries): count by the set layout
    fz = np.concatenate([x[vo[6 + e]:vo[7 + e]][4::6] for e in range(4)])
    expect_force = 0.7 * float(fz @ fz)
    P0 = _problem(oracle, 0.7, 0.0)
    assert abs(P0.cost(x) - expect_force) < 1e-12 * max(1.0, expect_force)
    assert f > expect_force                                      # the velocity term adds to it
    fd = np.array([(P.cost(x + 1e-6 * e) - P.cost(x - 1e-6 * e)) / 2e-6 for e in np.eye(P.n)])
    assert np.abs(fd - g).max() < 1e-6
    assert np.count_nonzero(g) == np.count_nonzero(fd > 1e-9) + np.count_nonzero(fd < -1e-9)
    # no weights: no objective
    Pn = _problem(oracle, 0.0, 0.0)
    assert Pn.cost(x) == 0.0 and not Pn.cost(x, with_grad=True)[1].any()


def test_c_oracle_prints_the_dense_emulators_table_with_an_objective(oracle):
    P = _problem(oracle, 1.0, 0.0)
    xc, rc = P.solve_ipopt(max_iter=9)
    o = E.Options(); o.max_iter = 9
    re = E.IpoptEmulator(P, o).solve()
    fmt = lambda *a: "%.2e %.2e %5.1f %.2e %.2e %.2e%s %d" % a
    for k in range(9):
        t = re.trace[k]
        assert fmt(rc.tr_inf_pr[k], rc.tr_inf_du[k], np.log10(rc.tr_mu[k]), rc.tr_dnorm[k], rc.tr_alpha_du[k], rc.tr_alpha_pr[k], rc.tr_tag[k:k + 1].decode(), rc.tr_ls[k]) == \
               fmt(t["inf_pr"], t["inf_du"], np.log10(t["mu"]), t["dnorm"], t["alpha_du"], t["alpha_pr"], t["tag"], t["ls"]), k
    # iteration 0 shows the objective: inf_du = ||grad f(x0)||_inf = 2 w m g / 4
    assert abs(rc.tr_inf_du[0] - 2 * 1.0 * 1.5 * 9.80665 / 4) < 1e-12


@pytest.mark.parametrize("wf,wv", [(1.0, 0.0), (0.05, 0.0)])
def test_objective_is_minimised_and_the_plan_stays_feasible(oracle, wf, wv):
    P = _problem(oracle, wf, wv)
    f0 = P.cost(P.x0())
    x, r = P.solve_ipopt()
    assert r.status == 0 and r.constr_viol <= 1e-4
    assert abs(r.objective - P.cost(x)) < 1e-12
    assert r.objective < 1e-3 * f0          # node values can all vanish: the force between two nodes is carried by their derivatives
    xl, xu, gl, gu = P.bounds()
    g = P.g(x)
    assert np.all(g >= gl - 1e-4) and np.all(g <= gu + 1e-4)
