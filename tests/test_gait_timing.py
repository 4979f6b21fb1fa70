"""SURVEY 8(f) rank 3, gait-timing optimisation at the oracle level: TOWR's optional PhaseDurations variables, the foot splines that
follow them, the duration columns of the dynamic and range-of-motion rows and the TotalDuration rows
(ref: solver/towr/src/phase_durations.cc:44-154, phase_spline.cc:55-93, polynomial.cc:236-257, total_duration_constraint.cc:48-70,
dynamic_constraint.cc:108-114, range_of_motion_constraint.cc:107-109, nlp_formulation.cc:80-83,186-198, parameters.cc:52,77-80).
Off on the reference's path (main.cpp never calls OptimizePhaseDurations) and NOT in the CUDA kernels: see DESIGN.md section 7.

The reference holds no golden data for this option; the Jacobian is pinned by finite differences (what Ipopt's derivative_test
does, main.cpp:454) and the solve by the dense Ipopt emulator, extended with variable bounds for the durations."""
import numpy as np
import pytest

from oracle import ipopt_emul as E


def _pair(O, combo, T, dmin=0.2, goal=(0.5, 0, 0.24)):
    sh, sh0 = O.default_shape(combo, T), O.default_shape(combo, T)
    sh.optimize_timings = 1; sh.phase_dur_min = dmin
    ter = O.Terrain(np.zeros((40, 20)), 0.1)
    inst = O.make_instance(goal=goal)
    return O.Problem(sh, inst, ter), O.Problem(sh0, inst, ter)


@pytest.mark.parametrize("combo,T", [("C1", 2.0), ("Custom", 5.0)])
def test_layout_bounds_and_initial_point(oracle, combo, T):
    P, P0 = _pair(oracle, combo, T)
    so, rt = P.schedule_layout()
    nph = [len(P.phase_durations(e)) for e in range(4)]
    assert P.n == P0.n + sum(nph) - 4 and P.m == P0.m + 4                    # all phases but the last; one TotalDuration row per foot
    assert so[0] == P0.n and rt == [P0.m + e for e in range(4)]              # schedule sets and TotalTime rows come last
    assert P0.schedule_layout() == ([-1] * 4, [-1] * 4)
    x0 = P.x0()
    xl, xu, gl, gu = P.bounds()
    for e in range(4):
        d = P0.phase_durations(e)
        sl = slice(so[e], so[e] + nph[e] - 1)
        assert np.array_equal(x0[sl], d[:-1]) and np.all(xl[sl] == 0.2) and np.all(xu[sl] == 1.0)
        assert gl[rt[e]] == 0.1 and gu[rt[e]] == T - 0.2                      # total_duration_constraint.cc:56-62
        assert abs(P.g(x0)[rt[e]] - d[:-1].sum()) < 1e-15
    # with the gait's own durations the rest of the problem is the fixed-timing one
    assert np.array_equal(x0[:P0.n], P0.x0())
    assert np.abs(P.g(x0)[:P0.m] - P0.g(P0.x0())).max() < 1e-12
    J = P.jac(x0)
    assert np.abs(J[:P0.m, :P0.n] - P0.jac(P0.x0())).max() < 1e-12
    assert np.array_equal(J[P0.m:, so[0]:].sum(axis=1), np.array(nph, dtype=float) - 1) and not J[P0.m:, :so[0]].any()
    # only the dynamic and range-of-motion rows see the durations
    _, ro = P0.layout()
    rows = np.nonzero(np.abs(J[:P0.m, so[0]:]).sum(axis=1))[0]
    assert rows.min() >= ro[4] and np.all((rows < ro[5]) | ((rows >= ro[7]) & (rows < ro[11])))


@pytest.mark.parametrize("combo,T", [("C1", 2.0), ("Custom", 5.0)])
def test_duration_jacobian_against_finite_differences(oracle, combo, T):
    P, P0 = _pair(oracle, combo, T)
    so, _ = P.schedule_layout()
    rng = np.random.default_rng(1)
    x = P.x0() + 0.02 * rng.standard_normal(P.n)
    x[so[0]:] = P.x0()[so[0]:] * (1 + 0.05 * rng.standard_normal(P.n - so[0]))      # durations moved, still positive
    J = P.jac(x)
    cols = list(range(so[0], P.n)) + list(range(0, so[0], 23))
    err = 0.0
    for c in cols:
        e = np.zeros(P.n); e[c] = 1e-7
        err = max(err, np.abs((P.g(x + e) - P.g(x - e)) / 2e-7 - J[:, c]).max())
    assert err < 2e-6, err                                                        # entries up to ~1e2, central differences
    assert np.count_nonzero(J[:, so[0]:]) > 1000


def test_emulated_ipopt_solves_with_free_timings(oracle):
    """C1 has feet with 9 phases in 2 s, so TOWR's default duration bounds (0.2, 1.0) leave 0.2 s of slack in total; with a lower
    bound of 0.1 the emulated Ipopt converges in 9 iterations, the durations move and every bound holds."""
    P, P0 = _pair(oracle, "C1", 2.0, dmin=0.1)
    so, rt = P.schedule_layout()
    res = E.IpoptEmulator(P).solve()
    assert res.status == 0 and res.constr_viol <= 1e-4 and res.iters <= 15
    xl, xu, gl, gu = P.bounds()
    g = P.g(res.x)
    assert np.all(g >= gl - 1e-4) and np.all(g <= gu + 1e-4)
    d = res.x[so[0]:]
    assert np.all(d >= 0.1 - 1e-8) and np.all(d <= 1.0 + 1e-8) and np.abs(d - P.x0()[so[0]:]).max() > 0.02
    for e in range(4):
        n = len(P.phase_durations(e)) - 1
        assert 2.0 - res.x[so[e]:so[e] + n].sum() >= 0.2 - 1e-4              # the last phase keeps its 0.2 s
    # the emulator with variable bounds still solves the fixed-timing problem as before (no bounded variables there)
    r0 = E.IpoptEmulator(P0).solve()
    assert r0.status == 0 and r0.iters == 7
