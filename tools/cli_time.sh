set -e
R=$(mktemp -d); mkdir -p $R/towr/build $R/towr/data/heightfields/from_pybullet
python - <<PY
import sys; sys.path.insert(0,'.')
import numpy as np
from qtos_b200 import heightfield as HF
g=np.load('tests/golden/heightfields.npz')
HF.write_heightfield('$R/towr/data/heightfields/from_pybullet/towr_heightfield.txt', g['exp_1_towr'])
PY
cd $R/towr/build
M=$GRAFT_REPO_ROOT/quadruped-trajectory-optimization-stack_b200/main
python - "$M" <<'PY'
import subprocess, sys, time
args = "-s 0 0 0.24 -g 0.5 0 0.24 -e1 0.21 0.19 0 -e2 0.21 -0.19 0 -e3 -0.21 0.19 0 -e4 -0.21 -0.19 0 -s_ang 0 0 0 -t 0 -r 15 -resolution 0.1".split()
for i in range(4):
    t = time.perf_counter(); p = subprocess.run([sys.argv[1]] + args, capture_output=True, text=True, env=dict(__import__("os").environ, QTOS_TIMING="1")); dt = time.perf_counter() - t
    print("run %d: exit %d, wall %.3f s, %s" % (i, p.returncode, dt, p.stdout.strip().splitlines()[-1]))
    if i == 3: print(p.stderr)
PY
wc -l traj.csv
