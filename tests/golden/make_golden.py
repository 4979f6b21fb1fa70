"""Regenerates tests/golden/*.npz / *.json from the READ-ONLY reference checkout (run in the build
container only: /root/reference does not exist on the GPU box).

  heightfields.npz   towr_heightfield grids for exp_1 / exp_3 / exp_5 produced by IMPORTING the
                     reference's QTOS/generateHeightField.py (pybullet / matplotlib stubbed: they are
                     not installed here and are not touched by the code that runs) in a scratch cwd
  gait_csv.npz       every 10th row of test/data/traj/gait.csv (G3) and of rows 1254..6254 of
                     data/traj/towr.csv (G2)  -- the reference's only golden plans
  towr_log.json      known answers read off logs/towr_log.out (G1): sizes, nnz, iteration-0 inf_pr
  cmd_args.json      QTOS.utils.cmd_args outputs for sample argument dicts
"""
import json, os, random, shutil, sys, tempfile, types
from unittest import mock
import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

for name in ["pybullet", "pybullet_data", "matplotlib", "matplotlib.pyplot", "pinocchio", "yaml", "pandas"]:
    if name not in sys.modules:
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = mock.MagicMock()
sys.path.insert(0, REF)

tmp = tempfile.mkdtemp()
shutil.copytree(os.path.join(REF, "data", "heightfields"), os.path.join(tmp, "data", "heightfields"))
os.makedirs(os.path.join(tmp, "data", "heightfields", "from_pybullet"), exist_ok=True)
os.chdir(tmp)
import QTOS.generateHeightField as G   # noqa: E402
import QTOS.utils as U                 # noqa: E402


def parse(path):
    rows = []
    for line in open(path).read().split("\n"):
        vals = [float(t) for t in line.replace(",", " ").split()]
        if vals:
            rows.append(vals)
    return np.array(rows)


out = {}
cfgs = {"exp_1": (["plane", "plane"], 1, False), "exp_3": (["feasibility", "feasibility_1", "plane"], 1, True),
        "exp_5": (["climb_2", "climb_1"], 11, False)}
for name, (maps, scale, rnd) in cfgs.items():
    random.seed(0)
    g = G.Height_Map_Generator(maps=maps, scale_factor=scale, randomize_env=rnd, bool_map_search=False)
    out[name + "_towr"] = parse(G.TOWR_HEIGHTFIELD_OUT)
    out[name + "_world"] = parse(G.HEIGHT_FIELD_OUT)
    out[name + "_res"] = np.array(g.resolution)
    print(name, out[name + "_towr"].shape, float(g.resolution), out[name + "_towr"].max())
# tiles used by the host-side heightfield mirror test
for t in ["plane", "feasibility_test", "climb_1", "random_terrain"]:
    out["tile_" + t] = G.heighmap_2_np_reader(os.path.join("data", "heightfields", t + ".txt"))
out["scale_map_3"] = G.scale_map(out["tile_climb_1"], 3)
np.savez_compressed(os.path.join(HERE, "heightfields.npz"), **out)

gait = np.loadtxt(os.path.join(REF, "test/data/traj/gait.csv"), delimiter=",")
towr = np.loadtxt(os.path.join(REF, "data/traj/towr.csv"), delimiter=",")
# towr_g4 = rows t = 2.502 .. 3.755 of the plan the FIRST logged solve wrote (the head of towr.csv before the
# splice at t = 3.756): with towr_g2 these are the input -> output pairs of the logged solves #1 and #2
np.savez_compressed(os.path.join(HERE, "gait_csv.npz"), gait=gait[::10], towr_g2=towr[1254:6255][::10],
                    towr_g4=towr[:1254][::10],
                    gait_rows=np.array(gait.shape[0]), towr_rows=np.array(towr.shape[0]))

import re  # noqa: E402


def iteration_tables(path):
    """the three Ipopt iteration tables (logs/towr_log.out:54-62,191-199,328-337) as printed strings"""
    tabs, cur = [], None
    for line in open(path).read().split("\n"):
        if line.startswith("iter    objective"):
            cur = []
            continue
        if cur is None:
            continue
        m = re.match(r"\s*(\d+)\s+(\S+)\s+(\S+)\s+(\S+)\s+(\S+)\s+(\S+)\s+(\S+)\s+(\S+)\s+(\S+?)([a-zA-Z]?)\s+(\d+)\s*$", line)
        if m:
            cur.append({"iter": int(m.group(1)), "inf_pr": m.group(3), "inf_du": m.group(4), "lg_mu": m.group(5),
                        "dnorm": m.group(6), "lg_rg": m.group(7), "alpha_du": m.group(8), "alpha_pr": m.group(9),
                        "tag": m.group(10), "ls": int(m.group(11))})
        elif not line.strip() and cur:
            tabs.append(cur)
            cur = None
    return tabs


def logged_inputs(path):
    """the argument echo in front of each solve (logs/towr_log.out:8-29,140-166,278-304): goal, start, start_ang,
    four feet; then (solves 2, 3) t_start ... resolution"""
    lines = open(path).read().split("\n")
    out = []
    for i, line in enumerate(lines):
        if line.startswith("This is Ipopt version"):
            j = i
            while not lines[j].startswith("****") or "TOWR" not in lines[j - 3]:
                j -= 1
            nums, replan = [], False
            for t in lines[j + 1:i]:
                try:
                    nums.append(float(t))
                except ValueError:
                    if nums and len(nums) >= 21:
                        replan = t.strip() == "s_vel"   # "<t_start> s_vel ..." follows the feet only on replans
                        break
            out.append({"goal": nums[0:3], "start_pos": nums[3:6], "start_ang": nums[6:9],
                        "ee": [nums[9:12], nums[12:15], nums[15:18], nums[18:21]],
                        "t_start": nums[21] if replan else 0.0})
    return out


log = {"n_vars_free": 1005, "n_vars_total": 1040, "n_fixed": 35, "n_eq": 706, "n_ineq": 1024,
       "nnz_eq": 11557, "nnz_ineq": 20605, "ineq_lower_only": 112, "ineq_both": 816, "ineq_upper_only": 96,
       "inf_pr_iter0": 19.4, "iters": [7, 7, 8],
       "iteration_tables": iteration_tables(os.path.join(REF, "logs/towr_log.out")),
       "inputs": logged_inputs(os.path.join(REF, "logs/towr_log.out")),
       "source": "logs/towr_log.out:8-29,40-64,98-129,140-166,191-201,278-304,328-339"}
json.dump(log, open(os.path.join(HERE, "towr_log.json"), "w"), indent=1)

samples = [
    {"-s": [0, 0, 0.24], "-g": [0.5, 0, 0.24], "-e1": [0.21, 0.19, 0.0], "-e2": [0.21, -0.19, 0.0],
     "-e3": [-0.21, 0.19, 0.0], "-e4": [-0.21, -0.19, 0.0], "-s_ang": [0, 0, 0], "-t": 0.0, "-r": 15.0,
     "-resolution": 0.1, "s_vel": [0, 0, 0]},
    {"-s": [0.335266, -0.0123145, 0.221551], "-g": [0.9100042764, 0.0, 0.24], "-t": 3.756, "-resolution": 0.01,
     "-s_ang": [-0.0422299, -0.0416417, 0.00880732], "scripts": {"run": "x"}, "step_size": 0.5},
]
cases = []
for a in samples:
    cases.append({"args": a, "out": U.cmd_args(dict(a))})
    print(cases[-1]["out"])
json.dump({"flags": list(U._flags), "cases": cases}, open(os.path.join(HERE, "cmd_args.json"), "w"), indent=1)
