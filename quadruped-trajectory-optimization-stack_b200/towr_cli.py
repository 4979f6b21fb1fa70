"""Host-side mirror of the reference's local-planner boundary: the `./main` command line.

QTOS reaches TOWR as `docker exec <id> ./main ` + cmd_args(args) (ref: QTOS/utils.py:15-26,644-670;
scripts/main.py:48-50,90-92,125-127; scripts/run.py:294-295; QTOS/generateHeightField.py:385-386).
This module keeps that surface:

  _flags, cmd_args(args)   same whitelist / string building as QTOS/utils.py:26,644-670
  parse_main_argv(argv)    the flag semantics of solver/towr/src/main.cpp:133-346
  towr_main(argv, cwd)     one solve: reads ../data/heightfields/from_pybullet/towr_heightfield.txt
                           relative to cwd (main.cpp:364), writes traj.csv into cwd (main.cpp:15,470),
                           returns the process exit code (= solver status, main.cpp:463,471)
  `python -m qtos_b200.towr_cli <flags>` is the drop-in for `./main <flags>`.
"""
import os
import sys

import numpy as np

from . import Solver, default_options, default_shape, make_problems, write_csv, NOMINAL_FEET
from .heightfield import read_towr_heightfield

# ref: QTOS/utils.py:26
_flags = ['-g', '-s', '-s_ang', '-s_vel', '-e1', '-e2', '-e3', '-e4', '-t', '-r', '-resolution', 's_vel',
          's_ang_vel', '-duration']

HEIGHTFIELD_REL = os.path.join("..", "data", "heightfields", "from_pybullet", "towr_heightfield.txt")
TRAJ_FILE = "traj.csv"


def _plain(v):
    """numpy >= 2 prints np.float64(1.0) inside str(list); the reference pins numpy 1.24 where it
    prints 1.0.  Cast so the command line stays what main.cpp can parse."""
    if isinstance(v, np.ndarray):
        return [_plain(x) for x in v.tolist()]
    if isinstance(v, (list, tuple)):
        return [_plain(x) for x in v]
    if isinstance(v, np.generic):
        return v.item()
    return v


def cmd_args(args):
    """dict -> argv string, same rules as QTOS.utils.cmd_args: only whitelisted keys, falsy values
    dropped, brackets and commas stripped, trailing space kept."""
    cmd = ""
    for key, value in args.items():
        value = _plain(value) if not isinstance(value, (dict, str)) else value
        if key in _flags and value:
            cmd += key + " " + str(value).replace(",", "").replace("[", "").replace("]", "") + " "
    return cmd


def _stod(tok):
    """std::stod: longest numeric prefix; raises ValueError (std::invalid_argument) if none."""
    tok = tok.strip()
    for end in range(len(tok), 0, -1):
        try:
            return float(tok[:end])
        except ValueError:
            continue
    raise ValueError("stod: no conversion: %r" % tok)


def _grab(argv, opt):
    """getcmdParser: the (up to) three tokens after the first occurrence of opt, or None when the
    option is absent or is the last token (ref: main.cpp:61-81)."""
    if opt in argv:
        i = argv.index(opt)
        if i + 1 < len(argv):
            return argv[i + 1:i + 4]
    return None


def parse_main_argv(argv):
    """argv (without argv[0]) -> dict(start, start_ang, start_vel, start_ang_vel, goal, ee, t_start,
    runtime, resolution, duration, combo).  Unknown tokens are ignored like main.cpp does."""
    argv = list(argv)
    out = dict(goal=[0.5, 0.0, 0.24], runtime=15.0, start=[0.0, 0.0, 0.24], start_ang=[0.0, 0.0, 0.0],
               start_ang_vel=[0.0, 0.0, 0.0], start_vel=[0.0, 0.0, 0.0], normalize=False,
               ee=[[f[0], f[1], 0.0] for f in NOMINAL_FEET], t_start=0.0, resolution=0.1, duration=5.0,
               combo="Custom")
    if not argv:
        return out

    def vec3(opt, key):
        t = _grab(argv, opt)
        if t is not None:
            out[key] = [_stod(t[0]), _stod(t[1]), _stod(t[2])]

    def scalar(opt, key):
        t = _grab(argv, opt)
        if t is not None:
            out[key] = _stod(t[0])

    try:
        vec3("-g", "goal"); scalar("-r", "runtime"); vec3("-s", "start"); vec3("-s_ang", "start_ang")
        vec3("-s_ang_vel", "start_ang_vel"); vec3("-s_vel", "start_vel")
        t = _grab(argv, "-n")
        if t is not None:
            out["normalize"] = " ".join(t) == "t"       # the reference compares the JOINED tokens (main.cpp:66,228-232): only a trailing `-n t` normalises
        else:
            out["start_vel"] = [0.0, 0.0, 0.0]          # main.cpp:237-242: no -n => start velocity zeroed
        for i in range(4):
            t = _grab(argv, "-e%d" % (i + 1))
            if t is not None:
                out["ee"][i] = [_stod(t[0]), _stod(t[1]), _stod(t[2])]
        scalar("-t", "t_start"); scalar("-resolution", "resolution")
        t = _grab(argv, "-duration")
        if t is not None:
            out["duration"] = _stod(t[0]); out["combo"] = "C0"   # main.cpp:299-306,425-427
    except (ValueError, IndexError) as e:
        # main.cpp:308-311 prints and carries on with whatever was parsed so far
        sys.stderr.write("Argument input error\nError: %s\n" % e)
    if out["normalize"]:
        out["goal"][0] -= out["start"][0]; out["goal"][1] -= out["start"][1]
        out["start"][0] = 0.0; out["start"][1] = 0.0
    return out


def problem_from_args(a, hf_id=0):
    p = make_problems(1)
    p["start_pos"][0] = a["start"]; p["start_ang"][0] = a["start_ang"]; p["start_vel"][0] = a["start_vel"]
    p["start_ang_vel"][0] = a["start_ang_vel"]; p["goal"][0] = a["goal"]; p["ee"][0] = a["ee"]
    p["t_start"][0] = a["t_start"]; p["hf_id"][0] = hf_id
    return p


_SOLVERS = {}
_HEIGHTFIELDS = {}                 # insertion-ordered: (solver, path, mtime, size, resolution) -> device heightfield id
HEIGHTFIELD_CACHE = 8              # grids kept on the device per solver of a long-lived process


SOLVER_BATCH = 64                  # windows one cached solver takes per launch (the daemon batches queued requests)


def _solver(combo, duration, device=0):
    key = (combo, float(duration), device)
    if key not in _SOLVERS:
        _SOLVERS[key] = Solver(default_shape(combo, duration), device=device, max_batch=SOLVER_BATCH)
    return _SOLVERS[key]


def exit_code(status):
    """what the calling shell sees for `return status` (ref: main.cpp:471)."""
    return int(status) & 0xFF


def towr_main_many(requests, device=0):
    """`requests` = [(argv, cwd), ...] -> [(exit code, stdout text, stderr text), ...].  Every request is one `./main`
    call; requests that share a shape and a runtime budget are solved as ONE batch (what the 32 concurrent PATH_MAP
    workers of the reference amount to, ref: QTOS/generateHeightField.py:344-404).  traj.csv is written per request,
    in request order, exactly where the one-by-one calls would have written it."""
    out = [None] * len(requests)
    groups = {}
    for k, (argv, cwd) in enumerate(requests):
        a = parse_main_argv(argv)
        hf_path = os.path.join(cwd, HEIGHTFIELD_REL)
        if not os.path.exists(hf_path):
            # the reference prints and then reads an empty grid (UB); the replacement fails loudly
            out[k] = (2, "", "Could not open file %s\n" % hf_path)
            continue
        try:
            S = _solver(a["combo"], a["duration"], device)
            st = os.stat(hf_path)
            key = (id(S), os.path.abspath(hf_path), st.st_mtime_ns, st.st_size, float(a["resolution"]))
            hid = _HEIGHTFIELDS.get(key)
            if hid is None:         # a long-lived process (serve.py) re-reads and re-uploads only a CHANGED terrain file
                mine = [q for q in _HEIGHTFIELDS if q[0] == id(S)]
                busy = {g[1] for gk, gl in groups.items() if gk[0] == id(S) for g in gl}
                if len(mine) >= HEIGHTFIELD_CACHE and _HEIGHTFIELDS[mine[0]] not in busy:
                    S.free_heightfield(_HEIGHTFIELDS.pop(mine[0]))     # oldest grid of this solver leaves the device, its id is reused
                grid = read_towr_heightfield(hf_path)
                hid = _HEIGHTFIELDS[key] = S.upload_heightfield(grid, a["resolution"])
        except Exception as e:
            out[k] = (3, "", "qtos: %s\n" % e)
            continue
        # -r (max_cpu_time, ref: main.cpp:180,459-460) becomes an iteration budget, see qtos_options.max_cpu_time
        groups.setdefault((id(S), float(a["runtime"])), []).append((k, hid, a, S, cwd))
    for (_, runtime), members in groups.items():
        S = members[0][3]
        p = np.concatenate([problem_from_args(a, hid) for _, hid, a, _, _ in members])
        try:
            res, x, rows = S.solve(p, default_options(max_cpu_time=runtime), csv=True)
        except Exception as e:
            for k, *_ in members:
                out[k] = (3, "", "qtos: %s\n" % e)
            continue
        for j, (k, _, _, _, cwd) in enumerate(members):
            try:
                write_csv(rows[j], os.path.join(cwd, TRAJ_FILE))
            except Exception as e:
                out[k] = (3, "", "qtos: %s\n" % e)
                continue
            out[k] = (exit_code(res["status"][j]),
                      "status -> %d  iterations %d  constraint violation %.3e\n" % (res["status"][j], res["iters"][j], res["constr_viol"][j]), "")
    return out


def towr_main(argv, cwd=".", device=0, quiet=False):
    rc, text, err = towr_main_many([(argv, cwd)], device)[0]
    if err:
        sys.stderr.write(err)
    if text and not quiet:
        sys.stdout.write(text)
    return rc


if __name__ == "__main__":
    sys.exit(towr_main(sys.argv[1:]))
