"""the in-process PATH_MAP mirror against vectors produced by RUNNING the reference's PATH_MAP
(tests/golden/make_golden_path_map.py: fake ./main exit codes, the reference's own probe_map / worker_f / hull)."""
import os
import shlex

import numpy as np
import pytest

import qtos_b200 as Q
from qtos_b200 import heightfield as HF
from qtos_b200 import path_map as PM
from qtos_b200 import towr_cli
from conftest import GOLDEN

GOLD = np.load(os.path.join(GOLDEN, "path_map.npz"))
CASES = ["exp_3", "scale2"]


@pytest.mark.parametrize("name", CASES)
def test_probe_set_and_flags_are_the_references(name):
    m, shift, scale = GOLD[name + "_map"], int(GOLD[name + "_shift"]), int(GOLD[name + "_scale"])
    probes = PM.probe_set(m, shift, 0.1 * (1 / scale))
    cmds = [str(c) for c in GOLD[name + "_cmds"]]
    assert len(probes["start"]) == len(cmds) > 0
    p = PM.probe_problems(probes, hf_id=3)
    for k, c in enumerate(cmds):                      # what ./main parses from the reference's command line
        a = towr_cli.parse_main_argv(shlex.split(c))
        assert a["start"] == list(p["start_pos"][k]) and a["goal"] == list(p["goal"][k])
        assert a["ee"] == [list(e) for e in p["ee"][k]]
        assert a["start_ang"] == [0.0, 0.0, 0.0] and a["runtime"] == PM.PROBE_RUNTIME and a["combo"] == "Custom"
        assert a["resolution"] == 0.1
    assert np.all(p["hf_id"] == 3) and np.all(p["start_vel"] == 0)


@pytest.mark.parametrize("name", CASES)
def test_marking_is_bit_exact(name):
    m, shift, scale = GOLD[name + "_map"], int(GOLD[name + "_shift"]), int(GOLD[name + "_scale"])
    probes = PM.probe_set(m, shift, 0.1 * (1 / scale))
    hull = PM.diamond(scale)
    assert np.array_equal(hull, GOLD[name + "_hull"])
    got = PM.mark(m.shape, probes, GOLD[name + "_status"] == 0, hull)
    assert got.dtype == GOLD[name + "_bool_map"].dtype and np.array_equal(got, GOLD[name + "_bool_map"])
    # all feasible -> empty map; all infeasible -> superset of the golden marks
    assert PM.mark(m.shape, probes, np.ones(len(probes["start"]), bool), hull).sum() == 0
    allbad = PM.mark(m.shape, probes, np.zeros(len(probes["start"]), bool), hull)
    assert np.all(allbad >= got)


def test_edge_cases():
    assert not PM.neighbors_danger(np.zeros((4, 4)), 0, 0)             # border cell: the scan leaves the map first
    m = np.zeros((5, 6)); m[2, 3] = 0.5
    assert PM.neighbors_danger(m, 2, 2) and PM.neighbors_danger(m, 1, 2) and not PM.neighbors_danger(m, 2, 3)
    empty = PM.probe_set(np.zeros((6, 8)))
    assert empty["start"].shape == (0, 3) and empty["idx_goal"].shape == (0, 2)
    assert len(PM.probe_problems(empty)) == 0
    assert PM.mark((6, 8), empty, np.zeros(0, bool), PM.diamond()).sum() == 0

    class NoSolver:                                                    # flat ground never reaches the solver
        def upload_heightfield(self, *a):
            raise AssertionError("solver touched on flat ground")
    pm = PM.PathMap(np.zeros((20, 40)), NoSolver(), np.zeros((40, 20)))
    assert pm.bool_map.shape == (20, 40) and pm.bool_map.sum() == 0 and pm.results is None


@pytest.mark.gpu
def test_path_map_batch_matches_one_by_one_cli_semantics(oracle):
    """exp_3 world map: the whole probe queue as ONE batch.  Each probe's status equals what the oracle (Ipopt algorithm, oracle/towr_ipopt.c) returns
    for the same flags (the `returncode == 0` test of worker_f), and bool_map is the FIFO marking of those."""
    from conftest import oracle_problem
    m, shift = GOLD["exp_3_map"], int(GOLD["exp_3_shift"])
    grid = HF.towr_grid(m)
    S = Q.Solver(Q.default_shape("Custom", 5.0), max_batch=64)
    before = S.launch_count()
    pm = PM.PathMap(m, S, grid, multi_map_shift=shift, scale=1)
    assert S.launch_count() > before and len(pm.feasible) == len(GOLD["exp_3_cmds"])
    so = oracle.default_shape("Custom", 5.0)
    p = PM.probe_problems(pm.probes, 0)
    for k in range(0, len(p), 6):
        xo, ro = oracle_problem(oracle, so, p[k], grid, 0.1).solve_ipopt()
        assert (ro.status == 0) == bool(pm.feasible[k]), k
    assert np.array_equal(pm.bool_map, PM.mark(m.shape, pm.probes, pm.feasible, PM.diamond(1)))
    print("exp_3 probes:", len(pm.feasible), "feasible:", int(pm.feasible.sum()), "marked cells:", int(pm.bool_map.sum()))
    # a single 3 cm bump on otherwise flat ground: every probe of the queue checked against the oracle
    m2 = np.zeros((20, 40)); m2[10, 21] = 0.03; m2[4, 8] = 0.03
    grid2 = HF.towr_grid(m2)
    pm2 = PM.PathMap(m2, S, grid2, multi_map_shift=2, scale=1)
    p2 = PM.probe_problems(pm2.probes, 0)
    assert len(p2) > 0
    for k in range(len(p2)):
        xo, ro = oracle_problem(oracle, so, p2[k], grid2, 0.1).solve_ipopt()
        assert (ro.status == 0) == bool(pm2.feasible[k]), k
    assert np.array_equal(pm2.bool_map, PM.mark(m2.shape, pm2.probes, pm2.feasible, PM.diamond(1)))
    print("bump map probes:", len(p2), "feasible:", int(pm2.feasible.sum()), "marked cells:", int(pm2.bool_map.sum()))
    S.close()
