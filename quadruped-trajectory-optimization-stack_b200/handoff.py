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


# ---------------------------------------------------------------------------------------------------------------------
# Combiner on arrays.  The four functions below are the array forms of utils.look_ahead,
# Combiner._state, Combiner._truncate_csv and Combiner.combine (ref: QTOS/utils.py:495-521, QTOS/combiner.py:78-92,125-135,
# 245-312); `values` is always what a reader of the CSV file would hold (as_csv_values of the sampler rows, or the result
# of an earlier combine).  tests/test_handoff.py checks them against outputs of the reference's own methods
# (tests/golden/make_golden_handoff.py).

def look_ahead_rows(values, start_time=0.0, timesteps=6000, ndigits=3):
    """-> (index of the row the reference's reader would hand out next, stop_idx).  Raises StopIteration like the
    reference's csv reader when the plan ends first."""
    t = np.round(np.asarray(values)[:, 0], ndigits)
    hit = np.flatnonzero(start_time <= t)
    if len(hit) == 0:
        raise StopIteration
    stop_idx = int(hit[0]) + 1
    nxt = stop_idx + timesteps - 1
    if nxt - 1 > len(t):                      # the skip loop itself ran off the end of the file
        raise StopIteration
    return nxt, stop_idx


def feet_in_contact(row, height_set, tol=6):
    """Combiner.check_legs_contact on one row: every foot z, rounded to `tol` decimals, is a height of the map."""
    return all(round(float(row[9 + 3 * e]), tol) in height_set for e in range(4))


def combiner_state(values, last_timestep, lookahead, height_set, ndigits=3):
    """Combiner._state: the first row at or after the look-ahead with all four feet on a terrain height (or, when the plan
    ends first, the look-ahead row itself).  -> (state dict, next_traj_step, lookahead actually used)."""
    values = np.asarray(values, dtype=np.float64)
    look = lookahead
    idx, step = look_ahead_rows(values, last_timestep, look, ndigits)
    while True:
        if idx >= len(values):                # StopIteration branch of the reference: start over, take the row as it is
            look = lookahead
            idx, step = look_ahead_rows(values, last_timestep, look, ndigits)
            row = values[idx]
            break
        row = values[idx]
        if feet_in_contact(row, height_set):
            break
        look += 1
        idx += 1
    st = state_of_row(row)
    st = {k: [0 if abs(x) < 1e-4 else x for x in v] for k, v in st.items()}          # utils.zero_filter
    return st, step + look - 1, look


def truncate_rows(values, cutoff_idx, next_traj_step):
    """Combiner._truncate_csv: pandas took row 0 as the header; iloc[start:end] of what is left."""
    start = 0 if cutoff_idx <= 0 else cutoff_idx - 1
    return np.asarray(values)[1:][start:next_traj_step]


def combine_rows(current_values, new_values, cutoff_idx, next_traj_step):
    """Combiner.combine: the kept part of the current plan followed by the new plan (whose first row went as a header).
    The result is both `traj_plan` and the content of the file the reference writes back."""
    return np.concatenate((truncate_rows(current_values, cutoff_idx, next_traj_step), np.asarray(new_values)[1:]), axis=0)


class ArrayPlans:
    """Array-backed stand-in for Combiner's two plan files.  `patch(combiner)` replaces the file-reading methods of a live
    reference Combiner (its attributes last_timestep, lookahead_original, cutoff_idx, next_traj_step, height_set are used
    and updated exactly as the originals do); `new_plan(rows)` is what `docker cp ... towr.csv` was."""

    def __init__(self, current_rows=None, rounded=True):
        """rounded=True: every plan goes through the "%g" rounding of traj.csv, so consumers hold exactly the values the file
        path gives them (190 ms per 5001-row plan in numpy); rounded=False keeps the sampler's full precision (1 ms)."""
        self.rounded = rounded
        self.current = None if current_rows is None else self._values(current_rows)
        self.new = None

    def _values(self, rows):
        return as_csv_values(rows) if self.rounded else np.array(rows, dtype=np.float64)

    def new_plan(self, rows):
        self.new = self._values(rows)

    def promote(self):
        """scripts/main.py:54-57: the combined plan becomes the current one"""
        self.current = self.new

    def patch(self, combiner):
        plans = self

        def _state():
            st, nts, look = combiner_state(plans.current, combiner.last_timestep, combiner.lookahead_original, combiner.height_set)
            combiner.lookahead = look
            combiner.next_traj_step = nts
            return st

        def combine():
            if combiner.cutoff_idx <= 0:
                combiner.cutoff_idx = 0
            out = combine_rows(plans.current, plans.new, combiner.cutoff_idx, combiner.next_traj_step)
            combiner.traj_plan = out
            plans.new = out

        combiner._state, combiner.combine = _state, combine
        return combiner
