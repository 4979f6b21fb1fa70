"""Development: where does an N-GPU bench step lose time?  Solve every rank's shard of the N x 4096-window workload of
bench.py on ONE GPU, one after the other, and print the device time and the iteration count of each: a step of the
N-GPU job ends at the all-gather, i.e. it costs the slowest shard.  usage: python tools/shard_tail.py [world=8]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import qtos_b200 as Q
from qtos_b200 import parallel
from bench import build_workload, COMBO, DURATION, PER_GPU

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
grid, res, p_all = build_workload(PER_GPU * world)
S = Q.Solver(Q.default_shape(COMBO, DURATION), max_batch=PER_GPU)
hid = S.upload_heightfield(grid, res)
S.set_profiling(True)
stream = torch.cuda.ExternalStream(S.stream)
rows = []
for rank in range(world):
    p = np.ascontiguousarray(p_all[parallel.shard_indices(PER_GPU * world, rank, world)])
    p["hf_id"] = hid
    ms = []
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        r, x, _ = S.solve(p)
        e1.record(stream)
        torch.cuda.synchronize()
        if rep:
            ms.append(e0.elapsed_time(e1))
    st = S.last_stats()
    rows.append((rank, float(np.median(ms)), st["iterations"], int(r["iters"].max()), float(r["iters"].mean()), int((r["status"] == 0).sum())))
    print("rank %d: %.2f ms  batch iterations %d  iters max %d mean %.2f  converged %d" % rows[-1], flush=True)
t = [r[1] for r in rows]
print("world %d: fastest shard %.2f ms, slowest %.2f ms, mean %.2f ms" % (world, min(t), max(t), sum(t) / len(t)))
