"""Experiment: throughput of the streaming session when every plan's 1 kHz rows are delivered to page-locked host memory
(qtos_stream_submit_csv), against the number of half-step jobs queued.  usage: rows_probe.py [steps] [depths ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qtos_b200 as Q
from bench import build_workload, COMBO, DURATION

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
depths = [int(a) for a in sys.argv[2:]] or [16]
N = 4096; n_sub = N // 2
grid, res, p = build_workload(N)
S = Q.Solver(Q.default_shape(COMBO, DURATION), max_batch=N)
p["hf_id"] = S.upload_heightfield(grid, res)
h_rows = [torch.empty((n_sub, S.csv_rows, Q.CSV_COLS), dtype=torch.float64).pin_memory() for _ in range(max(depths))]
h_res = [torch.empty(n_sub * Q.RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory() for _ in range(max(depths))]
h_x = [torch.empty((n_sub, S.n_vars), dtype=torch.float64).pin_memory() for _ in range(max(depths))]
h_p = torch.from_numpy(p.view(np.uint8).reshape(N, -1)).pin_memory()
pp = h_p.numpy().view(Q.PROBLEM_DTYPE).reshape(N)
S.stream_begin()
for depth in depths:
    def run(nsteps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); conv = 0; inflight = []
        for s in range(nsteps):
            if len(inflight) == depth:
                r, x = S.stream_wait(inflight.pop(0)); conv += int((r["status"] == 0).sum())
            k = s % depth
            inflight.append(S.stream_submit(pp[(s % 2) * n_sub:(s % 2 + 1) * n_sub], (h_res[k].numpy().view(Q.RESULT_DTYPE).reshape(n_sub), h_x[k].numpy()), csv_out=h_rows[k].numpy()))
        for t in inflight:
            r, x = S.stream_wait(t); conv += int((r["status"] == 0).sum())
        torch.cuda.synchronize()
        return conv, time.perf_counter() - t0
    run(depth)
    conv, dt = run(2 * steps)
    print("rows delivered, %d half-step jobs queued: %d steps %.3f s  %.0f solves/s  %.1f ms/step  finite %s" % (depth, steps, dt, conv / dt, 1e3 * dt / steps, bool(np.isfinite(h_rows[0].numpy()[::64, ::50]).all())), flush=True)
S.stream_end(); S.close()
