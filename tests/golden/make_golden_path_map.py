"""Regenerates tests/golden/path_map.npz by RUNNING the reference's PATH_MAP (build container only:
/root/reference does not exist on the GPU box).

The reference needs a TOWR docker container; here `DockerInfo` is stubbed, `subprocess.run` is replaced by a
fake `./main` whose exit code is a fixed function of the `-s` flag it receives (so the marking logic is exercised
on both branches), and `run()` drives the reference's own `worker_f` with ONE worker in this process, which makes
the order of the marks the queue order.  Stored: the world map, every command line the reference built, the fake
exit codes, the probe indices and the resulting bool_map, for exp_3 (scale 1, three tiles) and a scale-2 map.
"""
import os, random, shlex, shutil, sys, tempfile, time, types
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
os.makedirs(os.path.join(tmp, "data", "plots"), exist_ok=True)
os.chdir(tmp)
import QTOS.generateHeightField as G   # noqa: E402

# the reference pins numpy 1.24 (requirements.txt), where str([np.float64(0.5)]) prints 0.5; numpy 2 would leak
# "np.float64(0.5)" into the command line -- reproduce the pinned behaviour
np.set_printoptions(legacy="1.25")

LOG = []


def fake_status(cmd):
    """exit code of the fake ./main: 0 unless hash(start x, y) % 3 == 0"""
    i = cmd.index("-s")
    x, y = float(cmd[i + 1]), float(cmd[i + 2])
    return 0 if (int(round(x * 100)) * 7 + int(round(y * 100)) * 13) % 3 else 1


def fake_run(cmd, *a, **k):
    if "./main" in cmd:
        LOG.append(" ".join(cmd[cmd.index("./main") + 1:]))
        return types.SimpleNamespace(returncode=fake_status(cmd))
    return types.SimpleNamespace(returncode=0)


def one_worker_run(self):
    time.sleep(0.5)                      # let the queue's feeder thread flush before worker_f polls empty()
    self.worker_f(self.shared_arr, self.data_queue, self.num_cols)
    self.bool_map = np.frombuffer(self.shared_arr.get_obj(), dtype=np.float32).reshape(self.map.shape[0], self.map.shape[1]).astype('int')


out = {}
with mock.patch.object(G, "DockerInfo", lambda: "0123abcd"), mock.patch.object(G.subprocess, "run", fake_run), \
        mock.patch.object(G.PATH_MAP, "run", one_worker_run):
    for name, maps, scale in (("exp_3", ["feasibility", "feasibility_1", "plane"], 1), ("scale2", ["feasibility", "step_1"], 2)):
        random.seed(0)
        g = G.Height_Map_Generator(maps=maps, scale_factor=scale, randomize_env=(name == "exp_3"), bool_map_search=False)
        LOG.clear()
        pm = G.PATH_MAP(g.map, multi_map_shift=g.multi_map_shift, scale=scale)
        out[name + "_map"] = np.array(g.map, dtype=np.float64)
        out[name + "_shift"] = np.array(g.multi_map_shift)
        out[name + "_scale"] = np.array(scale)
        out[name + "_cmds"] = np.array(list(LOG))
        out[name + "_status"] = np.array([fake_status(shlex.split("x " + c)) for c in LOG])
        out[name + "_bool_map"] = np.array(pm.bool_map)
        out[name + "_hull"] = np.array(pm.neighbors_start)
        print(name, g.map.shape, "probes", len(LOG), "failed", int(out[name + "_status"].sum()), "marked", int(pm.bool_map.sum()))
np.savez_compressed(os.path.join(HERE, "path_map.npz"), **out)
