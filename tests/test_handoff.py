"""CPU tier: the array forms of the reference's plan hand-off (qtos_b200.handoff) against outputs of the reference's OWN
Combiner._state / _truncate_csv / combine, produced by tests/golden/make_golden_handoff.py on two consecutive plans."""
import os
import types

import numpy as np

from conftest import GOLDEN
from qtos_b200 import handoff as H

G = np.load(os.path.join(GOLDEN, "handoff.npz"))
KEYS = ("CoM", "orientation", "FL_FOOT", "FR_FOOT", "HL_FOOT", "HR_FOOT", "CoM_vel", "CoM_vel_ang")


def test_state_and_combine_equal_the_reference_methods():
    plans = H.ArrayPlans(G["rows_a"])
    plans.new_plan(G["rows_b"])
    c = types.SimpleNamespace(height_set={0.0})
    plans.patch(c)
    for rnd in range(2):
        last_t, look, cutoff = G["r%d_in" % rnd]
        c.last_timestep, c.lookahead_original, c.cutoff_idx, c.next_traj_step = float(last_t), int(look), int(cutoff), 0
        st = c._state()
        assert np.array_equal(np.array([st[k] for k in KEYS]), G["r%d_state" % rnd])
        assert c.next_traj_step == int(G["r%d_next_traj_step" % rnd]) and c.lookahead == int(G["r%d_lookahead" % rnd])
        c.combine()
        # all 37 columns; pandas' default float parser is not round-trip exact (-3.46945e-18 reads back one ulp off), the
        # array path holds the correctly rounded value: equal to 1 ulp, and exactly equal on everything a consumer reads
        want = G["r%d_traj_plan" % rnd]
        assert c.traj_plan.shape == want.shape and np.allclose(c.traj_plan, want, rtol=4e-16, atol=0.0)
        big = np.abs(want) > 1e-12
        assert np.array_equal(c.traj_plan[big], want[big])
        assert np.allclose(plans.new, G["r%d_file_after" % rnd], rtol=4e-16, atol=0.0)   # what the reference wrote back to towr.csv
        plans.promote()
        if rnd == 0:
            plans.new_plan(G["rows_b"])


def test_look_ahead_edges():
    v = H.as_csv_values(G["rows_a"])
    assert H.look_ahead_rows(v, 0.0, 1) == (1, 1)                    # the matching row is consumed: the NEXT row is handed out
    assert H.look_ahead_rows(v, 0.0105, 10) == (21, 12)              # 0.0105 <= round(t, 3) first at t = 0.011
    try:
        H.look_ahead_rows(v, 99.0, 10)
        assert False
    except StopIteration:
        pass
    # plan shorter than the look-ahead: the reference starts over and takes the look-ahead row without the contact test
    short = v[:500]
    try:
        H.combiner_state(short, 0.0, 600, {0.0})
        assert False
    except (StopIteration, IndexError):
        pass
