"""Experiment: the streaming session's throughput against the size of its pool of workspace slots (jobs stay 4096 windows;
more slots = more windows per launch = the partially filled last wave of every kernel weighs less).
usage: pool_probe.py [steps] [pool sizes ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qtos_b200 as Q
from bench import build_workload, COMBO, DURATION

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
pools = [int(a) for a in sys.argv[2:]] or [4096, 6144, 8192]
N = 4096
grid, res, p = build_workload(N)
dev = torch.device("cuda", 0)
for pool in pools:
    S = Q.Solver(Q.default_shape(COMBO, DURATION), max_batch=pool)
    p["hf_id"] = S.upload_heightfield(grid, res)
    d_p = torch.from_numpy(np.ascontiguousarray(p).view(np.uint8).reshape(N, -1)).to(dev)
    depth = max(8, 2 * pool // N + 4)
    d_res = [torch.zeros((N, Q.RESULT_DTYPE.itemsize), dtype=torch.uint8, device=dev) for _ in range(depth)]
    d_x = [torch.zeros((N, S.n_vars), dtype=torch.float64, device=dev) for _ in range(depth)]
    S.stream_begin()
    def run(steps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); conv = 0
        inflight = []
        for s in range(steps):
            if len(inflight) == depth:
                tk, slot = inflight.pop(0)
                S.stream_wait(tk); conv += int((d_res[slot].cpu().numpy().view(Q.RESULT_DTYPE)["status"] == 0).sum())
            slot = s % depth
            inflight.append((S.stream_submit_device(d_p.data_ptr(), N, d_res[slot].data_ptr(), d_x[slot].data_ptr()), slot))
        for tk, slot in inflight:
            S.stream_wait(tk); conv += int((d_res[slot].cpu().numpy().view(Q.RESULT_DTYPE)["status"] == 0).sum())
        torch.cuda.synchronize()
        return conv, time.perf_counter() - t0
    run(depth + 1)
    conv, dt = run(steps)
    print("pool of %d slots, %d jobs queued: %d steps %.3f s  %.0f solves/s  %.1f ms/step converged %d/%d" % (pool, depth, steps, dt, conv / dt, 1e3 * dt / steps, conv, steps * N), flush=True)
    S.stream_end(); S.close()
