import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, qtos_b200 as Q
from qtos_b200 import heightfield as HF, workloads
grid, res = HF.rough_terrain(1234)
S = Q.Solver(Q.default_shape("C1", 2.0), max_batch=256)
hid = S.upload_heightfield(grid, res)
p = workloads.multistart_problems(256, grid, res, hf_id=hid)
r, x, _ = S.solve(p)
rows = S.sample_csv(p, x)
print(rows.shape)
