"""SURVEY 8(b) zero-Python-edit option: the `docker` stand-in speaks the four verbs QTOS uses, driven here through
the reference's own command templates (QTOS/utils.py:14-24) and its DockerInfo parsing rule (utils.py:686-692)."""
import os
import shlex
import subprocess

import numpy as np
import pytest

from qtos_b200 import heightfield as HF
from qtos_b200 import towr_cli
from conftest import ROOT

SHIM_DIR = os.path.join(ROOT, "quadruped-trajectory-optimization-stack_b200", "shim")
# verbatim command templates of the reference (QTOS/utils.py:14-24)
SCRIPTS = {
    'copy': 'docker cp <id>:root/catkin_ws/src/towr_solo12/towr/build/traj.csv ./data/traj/towr.csv',
    'run': 'docker exec <id> ./main',
    'info': 'docker ps -f ancestor=towr',
    'data': 'docker cp <id>:root/catkin_ws/src/towr_solo12/towr/build/traj.csv /tmp/towr_shim_test.csv',
    'heightfield_rm': 'docker exec -t <id> rm /root/catkin_ws/src/towr_solo12/towr/data/heightfields/from_pybullet/towr_heightfield.txt',
    'heightfield_copy': 'docker cp ./data/heightfields/from_pybullet/towr_heightfield.txt <id>:root/catkin_ws/src/towr_solo12/towr/data/heightfields/from_pybullet/towr_heightfield.txt',
}


def _env(tmp_path):
    env = dict(os.environ)
    env["PATH"] = SHIM_DIR + os.pathsep + env["PATH"]
    env["QTOS_SHIM_ROOT"] = str(tmp_path / "container")
    return env


def _docker_info(env):
    """DockerInfo(): shell out, flatten, take the token before 'towr'."""
    p = subprocess.run([SCRIPTS['info']], shell=True, capture_output=True, text=True, env=env)
    out = p.stdout.replace('\n', ' ').split()
    return out[out.index('towr') - 1]


def _host_tree(tmp_path, grid):
    host = tmp_path / "host"
    (host / "data" / "heightfields" / "from_pybullet").mkdir(parents=True)
    (host / "data" / "traj").mkdir(parents=True)
    HF.write_heightfield(str(host / "data" / "heightfields" / "from_pybullet" / "towr_heightfield.txt"), grid)
    return host


def test_ps_cp_rm_verbs(tmp_path, golden_hf):
    env = _env(tmp_path)
    cid = _docker_info(env)
    s = {k: v.replace("<id>", cid) for k, v in SCRIPTS.items()}
    host = _host_tree(tmp_path, golden_hf["exp_3_towr"])
    inside = tmp_path / "container" / "data" / "heightfields" / "from_pybullet" / "towr_heightfield.txt"
    # Height_Map_Generator / PATH_MAP.setup order: rm (file may be absent) then copy
    assert subprocess.run(shlex.split(s['heightfield_rm']), env=env, cwd=host, capture_output=True).returncode == 1
    assert subprocess.run(shlex.split(s['heightfield_copy']), env=env, cwd=host).returncode == 0
    assert np.array_equal(HF.read_towr_heightfield(str(inside)), golden_hf["exp_3_towr"])
    assert subprocess.run(shlex.split(s['heightfield_rm']), env=env, cwd=host).returncode == 0 and not inside.exists()
    # copying a plan that does not exist yet fails like docker cp does
    assert subprocess.run(shlex.split(s['copy']), env=env, cwd=host, capture_output=True).returncode == 1
    assert subprocess.run(["docker", "exec", "nope", "./main"], env=env, capture_output=True).returncode == 1


@pytest.mark.gpu
def test_exec_main_through_the_shim(tmp_path, golden_hf):
    pkg = os.path.join(ROOT, "quadruped-trajectory-optimization-stack_b200")
    if not os.path.exists(os.path.join(pkg, "main")):          # normally built by __graft_entry__.build(); host C++ only
        subprocess.check_call(["make", "-s", "-C", os.path.join(pkg, "csrc"), "qtos_main"])
    env = _env(tmp_path)
    cid = _docker_info(env)
    s = {k: v.replace("<id>", cid) for k, v in SCRIPTS.items()}
    host = _host_tree(tmp_path, golden_hf["exp_1_towr"])
    subprocess.run(shlex.split(s['heightfield_copy']), env=env, cwd=host, check=True)
    args = {"-s": [0, 0, 0.24], "-g": [0.5, 0, 0.24], "-e1": [0.21, 0.19, 0.0], "-e2": [0.21, -0.19, 0.0],
            "-e3": [-0.21, 0.19, 0.0], "-e4": [-0.21, -0.19, 0.0], "-s_ang": [0, 0, 0], "-t": 2.5, "-r": 15.0, "-resolution": 0.1}
    p = subprocess.run(shlex.split(s['run'] + " " + towr_cli.cmd_args(args)), env=env, cwd=host, capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert subprocess.run(shlex.split(s['copy']), env=env, cwd=host).returncode == 0
    plan = np.loadtxt(host / "data" / "traj" / "towr.csv", delimiter=",")
    assert plan.shape == (5001, 37) and plan[0, 0] == 2.5 and abs(plan[-1, 1] - 0.5) < 1e-6
    # a blocked probe (right front foot ends on the 0.5 m block of exp_3) reports a non-zero exit code
    HF.write_heightfield(str(host / "data" / "heightfields" / "from_pybullet" / "towr_heightfield.txt"), golden_hf["exp_3_towr"])
    subprocess.run(shlex.split(s['heightfield_copy']), env=env, cwd=host, check=True)
    args["-t"] = 0.0
    p = subprocess.run(shlex.split(s['run'] + " " + towr_cli.cmd_args(args)), env=env, cwd=host, capture_output=True, text=True)
    assert p.returncode != 0


@pytest.mark.gpu
def test_daemon_behind_the_same_command_line(tmp_path, golden_hf):
    """`python -m qtos_b200.serve` keeps the CUDA context, the compiled shape and the uploaded terrain; the docker
    stand-in asks it when its socket answers.  Same traj.csv as the native ./main, a fraction of its latency."""
    import sys
    import time
    from qtos_b200 import serve
    pkg = os.path.join(ROOT, "quadruped-trajectory-optimization-stack_b200")
    if not os.path.exists(os.path.join(pkg, "main")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(pkg, "csrc"), "qtos_main"])
    env = _env(tmp_path)
    cid = _docker_info(env)
    s = {k: v.replace("<id>", cid) for k, v in SCRIPTS.items()}
    host = _host_tree(tmp_path, golden_hf["exp_1_towr"])
    subprocess.run(shlex.split(s['heightfield_copy']), env=env, cwd=host, check=True)
    args = {"-s": [0, 0, 0.24], "-g": [0.5, 0, 0.24], "-e1": [0.21, 0.19, 0.0], "-e2": [0.21, -0.19, 0.0],
            "-e3": [-0.21, 0.19, 0.0], "-e4": [-0.21, -0.19, 0.0], "-s_ang": [0, 0, 0], "-t": 1.0, "-r": 15.0, "-resolution": 0.1}
    cmd = shlex.split(s['run'] + " " + towr_cli.cmd_args(args))
    # native process first (no daemon yet): the reference csv
    t0 = time.perf_counter(); p = subprocess.run(cmd, env=env, cwd=host, capture_output=True, text=True); t_native = time.perf_counter() - t0
    assert p.returncode == 0, p.stdout + p.stderr
    build = tmp_path / "container" / "build"
    native_csv = open(build / "traj.csv").read()
    os.remove(build / "traj.csv")
    sock = str(tmp_path / "container" / "qtos.sock")
    env_d = dict(env, PYTHONPATH=ROOT + os.pathsep + env.get("PYTHONPATH", ""))
    d = subprocess.Popen([sys.executable, "-m", "qtos_b200.serve", "--socket", sock], env=env_d, cwd=ROOT, stdout=subprocess.PIPE, text=True)
    try:
        assert "ready" in d.stdout.readline()
        times = []
        for _ in range(3):
            t0 = time.perf_counter(); p = subprocess.run(cmd, env=env, cwd=host, capture_output=True, text=True); times.append(time.perf_counter() - t0)
            assert p.returncode == 0 and "status -> 0" in p.stdout, p.stdout + p.stderr
        assert open(build / "traj.csv").read() == native_csv            # byte-identical plan file
        # a blocked probe still reports a non-zero exit code through the daemon
        HF.write_heightfield(str(host / "data" / "heightfields" / "from_pybullet" / "towr_heightfield.txt"), golden_hf["exp_3_towr"])
        subprocess.run(shlex.split(s['heightfield_copy']), env=env, cwd=host, check=True)
        args["-t"] = 0.0
        p = subprocess.run(shlex.split(s['run'] + " " + towr_cli.cmd_args(args)), env=env, cwd=host, capture_output=True, text=True)
        assert p.returncode != 0
        print("native ./main %.3f s; through the daemon %s s" % (t_native, ["%.3f" % t for t in times]))
        assert min(times[1:]) < 0.5 * t_native
    finally:
        serve.request(sock, {"cmd": "shutdown"}, timeout=10.0)
        d.wait(timeout=30)


def test_daemon_protocol_without_a_device(tmp_path):
    """host side of the resident solver (no GPU needed): the stand-in routes `exec <id> ./main` to the socket when it
    answers, exit code and stderr come back through it, a bad request does not end the daemon, shutdown removes the socket."""
    import sys
    from qtos_b200 import serve
    env = _env(tmp_path)
    cid = _docker_info(env)
    (tmp_path / "container" / "build").mkdir(parents=True)
    sock = str(tmp_path / "container" / "qtos.sock")
    assert serve.request(sock, {"cmd": "ping"}) is None              # nobody there: the client says so, no exception
    env_d = dict(env, PYTHONPATH=ROOT + os.pathsep + env.get("PYTHONPATH", ""))
    d = subprocess.Popen([sys.executable, "-m", "qtos_b200.serve", "--socket", sock], env=env_d, cwd=ROOT, stdout=subprocess.PIPE, text=True)
    try:
        assert "ready" in d.stdout.readline()
        assert serve.request(sock, {"cmd": "ping"}) == {"rc": 0, "out": "pong"}
        with pytest.raises(RuntimeError):                            # a second daemon refuses the socket of a live one
            serve.serve(sock)
        # no heightfield in the container tree: the clone's loud exit code 2 (main.cpp:364 reads the file; missing -> UB there)
        p = subprocess.run(shlex.split(SCRIPTS['run'].replace("<id>", cid) + " -t 0.0"), env=env, capture_output=True, text=True)
        assert p.returncode == 2 and "Could not open file" in p.stderr
        import socket as _s
        c = _s.socket(_s.AF_UNIX, _s.SOCK_STREAM); c.connect(sock); c.sendall(b"not json\n")
        assert b"bad request" in c.recv(4096); c.close()
        assert serve.request(sock, {"cmd": "ping"})["out"] == "pong"
    finally:
        serve.request(sock, {"cmd": "shutdown"}, timeout=10.0)
        d.wait(timeout=30)
    assert not os.path.exists(sock)
