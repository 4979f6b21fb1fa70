"""Experiment: one streaming pool of 4096 slots against NP pools of 4096/NP slots on their own contexts / streams (phase-shifted
kernel mixes on one GPU).  usage: two_pools.py [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qtos_b200 as Q
from bench import build_workload, COMBO, DURATION

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
N = 4096
grid, res, p = build_workload(N)
dev = torch.device("cuda", 0)
for npools in (1, 2, 4):
    per = N // npools
    ctxs = [Q.Solver(Q.default_shape(COMBO, DURATION), max_batch=per) for _ in range(npools)]
    for c in ctxs:
        p["hf_id"] = c.upload_heightfield(grid, res)
    d_p = [torch.from_numpy(np.ascontiguousarray(p[k * per:(k + 1) * per]).view(np.uint8).reshape(per, -1)).to(dev) for k in range(npools)]
    depth = 8
    d_res = [[torch.zeros((per, Q.RESULT_DTYPE.itemsize), dtype=torch.uint8, device=dev) for _ in range(depth)] for _ in range(npools)]
    d_x = [[torch.zeros((per, ctxs[0].n_vars), dtype=torch.float64, device=dev) for _ in range(depth)] for _ in range(npools)]
    for c in ctxs:
        c.stream_begin()
    def run(steps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); conv = 0
        inflight = []
        for s in range(steps):
            if len(inflight) == depth:
                tk, slot = inflight.pop(0)
                for k, c in enumerate(ctxs):
                    c.stream_wait(tk[k]); conv += int((d_res[k][slot].cpu().numpy().view(Q.RESULT_DTYPE)["status"] == 0).sum())
            slot = s % depth
            inflight.append(([c.stream_submit_device(d_p[k].data_ptr(), per, d_res[k][slot].data_ptr(), d_x[k][slot].data_ptr()) for k, c in enumerate(ctxs)], slot))
        for tk, slot in inflight:
            for k, c in enumerate(ctxs):
                c.stream_wait(tk[k]); conv += int((d_res[k][slot].cpu().numpy().view(Q.RESULT_DTYPE)["status"] == 0).sum())
        torch.cuda.synchronize()
        return conv, time.perf_counter() - t0
    run(depth + 1)
    conv, dt = run(steps)
    print("pools %d x %d slots: %d steps %.3f s  %.0f solves/s  %.1f ms/step converged %d/%d" % (npools, per, steps, dt, conv / dt, 1e3 * dt / steps, conv, steps * N))
    for c in ctxs:
        c.stream_end(); c.close()
