"""Replan hand-off latency on the host: the reference's file detour (write traj.csv with "%g", pandas parse of the current and the
new plan, csv.reader walk to the look-ahead row -- QTOS/combiner.py:125-135,245-296) against the array forms in
qtos_b200.handoff, for one production-shape plan (5001 x 37).  CPU only; no solver involved."""
import csv, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pandas as pd
from qtos_b200 import handoff as H

rng = np.random.default_rng(0)
rows = np.cumsum(rng.standard_normal((5001, 37)) * 1e-3, axis=0); rows[:, 0] = np.arange(5001) * 1e-3
rows[:, [9, 12, 15, 18]] = 0.0                                 # feet on flat ground
tmp = tempfile.mkdtemp(); cur, new = os.path.join(tmp, "traj.csv"), os.path.join(tmp, "towr.csv")


def write(path, r):
    with open(path, "w") as f:
        for row in r:
            f.write(",".join("%g" % v for v in row) + "\n")


def file_path():
    write_new()                                                # stands for ./main's ofstream (22 ms measured) + docker cp
    with open(cur, newline="") as f:                           # Combiner._state
        rd = csv.reader(f); k = 0
        for r in rd:
            k += 1
            if 0.0 <= round(float(r[0]), 3):
                break
        for _ in range(3750 - 1):
            next(rd)
        st = [float(v) for v in next(rd)[1:]]
    old = pd.read_csv(cur).iloc[0:3750].to_numpy()             # Combiner.combine
    nw = pd.read_csv(new).to_numpy()
    comb = np.concatenate((old, nw), axis=0)
    pd.DataFrame(comb).to_csv(new, index=False, header=None)
    return comb


def array_path(plans):
    plans.new_plan(rows)
    st, nts, look = H.combiner_state(plans.current, 0.0, 3750, {0.0})
    return H.combine_rows(plans.current, plans.new, 0, nts)


write(cur, rows)
write(os.path.join(tmp, "fresh.csv"), rows)


def write_new():
    import shutil
    shutil.copyfile(os.path.join(tmp, "fresh.csv"), new)


plans, plans_full = H.ArrayPlans(rows), H.ArrayPlans(rows, rounded=False)
for name, fn in (("file detour (parse + walk + combine + write back)", file_path), ("arrays, %g-rounded values", lambda: array_path(plans)),
                 ("arrays, full precision", lambda: array_path(plans_full))):
    fn()
    t = time.perf_counter()
    for _ in range(5):
        out = fn()
    print("%-52s %.1f ms per replan (combined plan %s)" % (name, 1e3 * (time.perf_counter() - t) / 5, out.shape))
