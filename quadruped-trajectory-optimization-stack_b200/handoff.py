"""CSV-free handoff of a plan to the reference's consumers.

The reference writes traj.csv (37 columns, default ostream precision, solver/towr/src/main.cpp:83-131), copies it
out of the container and re-parses ~1.6 MB of text at every replan: `pd.read_csv(...).to_numpy()` in
Combiner.combine (QTOS/combiner.py:125-135, the FIRST ROW IS TAKEN AS THE HEADER and dropped), csv.reader +
float() in Combiner._state (combiner.py:245-296), np.loadtxt-style reads in scripts/run.py:184-188.  Here the
sampler kernel's rows go to those consumers as arrays; the functions below reproduce what the text detour would
have delivered, value for value, so a consumer cannot tell the difference.
"""
import numpy as np

EE_NAMES = ('FL_FOOT', 'FR_FOOT', 'HL_FOOT', 'HR_FOOT')          # ref: QTOS/utils.py:12


def as_csv_values(rows):
    """rows [n, 37] (full precision) -> the doubles a reader of traj.csv gets: every value rounded through
    "%g" (6 significant digits), exactly like ofstream << double followed by strtod."""
    rows = np.asarray(rows, dtype=np.float64)
    return np.char.mod("%g", rows).astype(np.float64)


def read_csv_frame(rows, rounded=True):
    """what `pd.read_csv(traj.csv).to_numpy()` returns: row 0 became the header, so the frame starts at row 1
    (ref: combiner.py:130-131,305)."""
    v = as_csv_values(rows) if rounded else np.asarray(rows, dtype=np.float64)
    return v[1:]


def state_of_row(row):
    """Combiner._state's dictionary for one CSV row (time in column 0 is skipped; ref: combiner.py:266-274)."""
    r = [float(v) for v in row[1:]]
    st = {"CoM": r[0:3], "orientation": r[3:6]}
    for i, name in enumerate(EE_NAMES):
        st[name] = r[6 + 3 * i:9 + 3 * i]
    st["CoM_vel"] = r[18:21]
    st["CoM_vel_ang"] = r[21:24]
    return st
