"""Regenerates tests/golden/handoff.npz by RUNNING the reference's Combiner methods (build container only: /root/reference
does not exist on the GPU box).

QTOS.combiner.Combiner reads the planner's output as CSV text: `_state()` (combiner.py:245-296) walks the current plan file
with utils.look_ahead to the first all-feet-in-contact row after the look-ahead, `combine()` (combiner.py:125-135) glues the
truncated old plan to the new one through pandas (dropping the first row of each file as a header) and writes the result
back.  Here a Combiner is created WITHOUT its constructor (which needs the simulator), pointed at two plan files written by
the CPU oracle in the reference's "%g" format, and its own `_state`, `_truncate_csv` and `combine` are run for two replanning
rounds.  Stored: the full-precision rows of both plans, every attribute that went in, every value that came out.
"""
import os, shutil, sys, tempfile
from unittest import mock
import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for name in ["pybullet", "pybullet_data", "matplotlib", "matplotlib.pyplot", "pinocchio", "yaml", "scipy", "scipy.spatial", "scipy.spatial.transform",
             "scipy.interpolate", "scipy.signal"]:
    if name not in sys.modules:
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = mock.MagicMock()
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
import oracle as O                       # noqa: E402
from QTOS.combiner import Combiner       # noqa: E402

tmp = tempfile.mkdtemp()
so = O.default_shape("C1", 2.0)
ter = O.Terrain(np.zeros((300, 300)), 0.02)
FEET = [(0.21, 0.19, 0.0), (0.21, -0.19, 0.0), (-0.21, 0.19, 0.0), (-0.21, -0.19, 0.0)]
pa = O.Problem(so, O.make_instance(start_pos=(0, 0, 0.24), goal=(0.4, 0.0, 0.24), ee=FEET), ter)
xa, ra = pa.solve_ipopt()
rows_a = pa.csv(xa)
last = rows_a[-1]
pb = O.Problem(so, O.make_instance(start_pos=last[1:4], start_ang=last[4:7], goal=(0.8, 0.0, 0.24), ee=last[7:19].reshape(4, 3), t_start=float(last[0])), ter)
xb, rb = pb.solve_ipopt()
rows_b = pb.csv(xb)
assert ra.status == 0 and rb.status == 0
cur, new = os.path.join(tmp, "traj.csv"), os.path.join(tmp, "towr.csv")
pa.write_csv(xa, cur); pb.write_csv(xb, new)

c = object.__new__(Combiner)
c.current_traj, c.new_traj = cur, new
c.decimal_precision = 3
c.height_set = {0.0}
out = {"rows_a": rows_a, "rows_b": rows_b}
for rnd, (last_t, look, cutoff) in enumerate(((0.0, 1500, 0), (1.2, 600, 1100))):
    c.last_timestep, c.lookahead, c.lookahead_original, c.cutoff_idx, c.next_traj_step = last_t, look, look, cutoff, 0
    st = c._state()
    out["r%d_in" % rnd] = np.array([last_t, look, cutoff], dtype=np.float64)
    out["r%d_state" % rnd] = np.array([st[k] for k in ("CoM", "orientation", "FL_FOOT", "FR_FOOT", "HL_FOOT", "HR_FOOT", "CoM_vel", "CoM_vel_ang")], dtype=np.float64)
    out["r%d_next_traj_step" % rnd] = np.array(c.next_traj_step)
    out["r%d_lookahead" % rnd] = np.array(c.lookahead)
    c.combine()
    out["r%d_traj_plan" % rnd] = np.array(c.traj_plan, dtype=np.float64)
    out["r%d_file_after" % rnd] = np.loadtxt(new, delimiter=",")
    print("round", rnd, "state CoM", st["CoM"], "next_traj_step", c.next_traj_step, "combined", c.traj_plan.shape)
    shutil.copyfile(new, cur)            # scripts/main.py:57 `docker cp` + the plan becomes the current one
    if rnd == 0:
        pb.write_csv(xb, new)            # the next window's plan arrives (same file name, like the reference)
np.savez_compressed(os.path.join(HERE, "handoff.npz"), **out)
