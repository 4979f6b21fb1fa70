#!/usr/bin/env python
"""bench.py -- converged gait-plan NLP solves/s on B200 (BASELINE.json metric), one JSON line.

  python bench.py [--gpus N --steps K --warmup W]          this repo's CUDA path
  python bench.py --impl reference [...]                   CPU arm: the oracle port of the reference's
                                                           algorithm (Ipopt 3.11.9 as TOWR runs it,
                                                           oracle/towr_ipopt.c) on the box's host cores --
                                                           the reference binary itself needs
                                                           Eigen3/ifopt/Ipopt/MUMPS: unbuildable here

A step = one pass of the hot path over one batch: 4096 independent local-plan windows per GPU, solved with
the reference's algorithm (QTOS_ALG_IPOPT) to the reference's convergence criteria (tol 1e-3, constr_viol
1e-4, compl 1e-4, max_iter 200).
  N = 1: BASELINE.json configs[3] -- batched multi-start, 4096 start/goal pairs on ONE 256x256 rough
         heightfield (seed 1234), trot gait C1, T = 2 s, 640 variables / 892 constraints.
  N > 1: BASELINE.json configs[4] -- replan sweep, 4096 x N receding-horizon windows over EIGHT terrain
         variants (seeds 0..7), each window started from the final state of a previously solved plan on its
         own grid; one process per GPU (torchrun), windows sharded statically (weak scaling), one all-gather
         of per-candidate records for best-plan selection inside the timed region.
The steps run through the library's continuous-batching session (qtos_stream_*): the 4096 workspace slots of the
context are a pool, a step is one job of 4096 windows, eight jobs are queued at a time (the slowest windows of a job take ~70 iterations, a job's bulk ~10), and the windows of the next
job enter the pool as the windows of the current one finish -- so every launch works on a full pool and the last slow
windows of a step do not run as a chain of near-empty launches.  `serial` in the JSON line is the same work through
the synchronous batch call (one step at a time, stragglers exposed); the per-phase times come from that pass, the
roofline figures from CUDA events the streaming session records around its k_factor and kip_solve launches.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PER_GPU = 4096
GROUP = 8
IN_FLIGHT = int(os.environ.get("QTOS_IN_FLIGHT", "8"))       # jobs queued in the streaming session
CSV_DEPTH = int(os.environ.get("QTOS_CSV_DEPTH", "16"))      # half-step jobs queued when every plan's 1 kHz rows are delivered (at most the library's 16 staging areas)
N_VARIANTS = 8
COMBO, DURATION = "C1", 2.0
CPU_SAMPLE = 192
WORKLOAD_1 = ("batched multi-start (BASELINE configs[3]): %d start/goal pairs per GPU on a 256x256 rough heightfield "
              "(seed 1234), trot C1, T=2s, 640 vars / 892 cons; groups of 8 candidates, best-plan selection" % PER_GPU)
WORKLOAD_N = ("replan sweep (BASELINE configs[4]): %d receding-horizon windows per GPU over 8 terrain variants (256x256, seeds "
              "0..7), each started from the final state of a solved plan on its grid, trot C1, T=2s; groups of 8 candidates, "
              "best-plan all-gather" % PER_GPU)
# algorithmic work of ONE factorization of the primal normal matrix of shape S2
# (n_free = 605, RCM envelope 50 728 entries): sum_i w_i^2 = 6.3e6 FP64 flop; the IPOPT path adds the forward
# substitution of its 16 right-hand sides (2 x envelope x 16 = 1.62e6) and their Gram matrix (2 x 605 x 16 x 16 = 0.31e6)
ALG_FLOP_FACTOR = 6.3e6
ALG_FLOP_RHS = 2.0 * 50728 * 16 + 2.0 * 605 * 16 * 16
TRAFFIC_JSON = os.path.join(ROOT, "profiles", "r02_traffic.json")     # ncu dram bytes per problem, written by tools/ncu_traffic.py


SWEEP_AT_N1 = os.environ.get("QTOS_BENCH_SWEEP") == "1"     # development: the N > 1 workload on one GPU (weak-scaling reference point)


def is_sweep(world):
    return world > 1 or SWEEP_AT_N1


def workload_name(world):
    return WORKLOAD_N if is_sweep(world) else WORKLOAD_1


def build_workload(n_total, seed=1234):
    """config 4: one terrain"""
    from qtos_b200 import heightfield as HF, workloads
    grid, res = HF.rough_terrain(seed)
    p = workloads.multistart_problems(n_total, grid, res, seed=seed, group_size=GROUP)
    return grid, res, p


# ------------------------------------------------------------------ CPU arm (oracle port of the reference algorithm)

def _cpu_solve(args):
    import oracle as O
    rec, grid, res, want_row = args
    so = O.default_shape(COMBO, DURATION)
    inst = O.make_instance(start_pos=rec["start_pos"], start_ang=rec["start_ang"], goal=rec["goal"], ee=rec["ee"],
                           t_start=float(rec["t_start"]))
    po = O.Problem(so, inst, O.Terrain(grid, res))
    t = time.perf_counter()
    x, r = po.solve_ipopt()
    dt = time.perf_counter() - t
    return r.status, dt, (po.csv(x)[-1] if want_row else None)


def cpu_sample_jobs(world, n_sample):
    """the first n_sample windows of rank 0's shard of the GPU arm's workload (for N > 1 the successors are produced by
    solving generation 0 with the oracle first, untimed)."""
    import multiprocessing as mp
    from qtos_b200 import parallel, workloads
    if not is_sweep(world):
        grid, res, p = build_workload(PER_GPU)
        return [(p[i], grid, res, False) for i in range(n_sample)]
    variants = workloads.terrain_variants(N_VARIANTS)
    p0 = workloads.replan_sweep_problems(PER_GPU * world, variants, list(range(N_VARIANTS)), group_size=GROUP)
    # spread the sample over the variants like the shard itself is
    idx = parallel.shard_indices(PER_GPU * world, 0, world)[:: max(1, PER_GPU // n_sample)][:n_sample]
    gen0 = [(p0[i], variants[p0[i]["hf_id"]][0], variants[p0[i]["hf_id"]][1], True) for i in idx]
    with mp.get_context("fork").Pool(os.cpu_count()) as pool:
        out = pool.map(_cpu_solve, gen0, chunksize=1)
    rows = np.array([o[2] for o in out])
    nxt = workloads.replan_from_rows(p0[idx], rows)
    return [(nxt[k], gen0[k][1], gen0[k][2], False) for k in range(len(idx))]


def cpu_arm(jobs, cores=None):
    """oracle Ipopt port on the host cores, one process per core; returns (converged solves/s, cores, p50 latency s, converged)."""
    import multiprocessing as mp
    cores = cores or os.cpu_count()
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_cpu_solve, jobs[:cores])                     # warm the workers
        t = time.perf_counter()
        out = pool.map(_cpu_solve, jobs, chunksize=1)
        dt = time.perf_counter() - t
    conv = sum(1 for s, _, _ in out if s == 0)
    lat = sorted(t_ for _, t_, _ in out)
    return conv / dt, cores, lat[len(lat) // 2], conv


CPU_KIND_NOTE = ("oracle/towr_ipopt.c: the reference's algorithm (Ipopt 3.11.9 as TOWR configures it) restated in C and pinned to "
                 "the reference's logged iteration tables and plans; TOWR+Ipopt itself cannot be built here")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as O
    O.build()
    world = args.gpus
    jobs = cpu_sample_jobs(world, CPU_SAMPLE)
    for _ in range(min(args.warmup, 1)):
        cpu_arm(jobs[:16])
    vals = []
    for _ in range(args.steps):
        v, cores, p50, conv = cpu_arm(jobs)
        vals.append(v)
    value = float(np.mean(vals))
    sample = "%d windows of rank 0's shard per step (of %d per GPU; the first ones at N = 1, every %d-th at N > 1), one process per host core" % (len(jobs), PER_GPU, max(1, PER_GPU // CPU_SAMPLE))
    line = {"impl": "reference", "metric": "converged gait-plan NLP solves/sec", "value": value, "unit": "solves/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * len(jobs) / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(world), "problems_per_gpu": PER_GPU, "parallelism": "host cores", "algorithm": "ipopt",
                       "sample": sample},
            "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample + "; " + CPU_KIND_NOTE,
                             "p50_latency_ms": 1e3 * p50, "converged": conv},
            "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


# ------------------------------------------------------------------ GPU arm

def clocks_sampler(path, device):
    q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    try:
        return subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"],
                                stdout=open(path, "w"), stderr=subprocess.DEVNULL)
    except Exception:
        return None


def parse_clocks(path):
    sm, mx, reasons = [], [], set()
    try:
        for line in open(path):
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9:
                continue
            sm.append(float(f[1])); mx.append(float(f[2]))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
    except Exception:
        pass
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
    return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons)}


def golden_parity(Q, device):
    """the reference's own logged solves (logs/towr_log.out, data/traj/towr.csv) on this device: distance of the GPU plans
    from the TOWR + Ipopt plans.  Fixtures travel in tests/golden/ (generated by tests/golden/make_golden.py)."""
    try:
        G = os.path.join(ROOT, "tests", "golden")
        log = json.load(open(os.path.join(G, "towr_log.json")))
        csv = np.load(os.path.join(G, "gait_csv.npz"))
        sh = Q.default_shape("Custom", 5.0, mass=3.0); sh.max_dev[0] = 0.08        # the build that produced the log
        S = Q.Solver(sh, device=device, max_batch=3)
        hid = S.upload_heightfield(np.zeros((600, 200)), 0.01)
        p = Q.make_problems(3)
        for k, inp in enumerate(log["inputs"]):
            for key in ("start_pos", "start_ang", "goal", "ee", "t_start"):
                p[k][key] = inp[key]
        p["hf_id"] = hid
        r, x, rows = S.solve(p, csv=True)
        com = feet = 0.0
        for k, (Gk, row0) in enumerate(((csv["towr_g4"], 2502), (csv["towr_g2"], 0))):
            rk = rows[k][row0::10][:len(Gk)]
            com = max(com, float(np.abs(rk[:, 1:4] - Gk[:, 1:4]).max())); feet = max(feet, float(np.abs(rk[:, 7:19] - Gk[:, 7:19]).max()))
        S.close()
        return {"vs": "TOWR+Ipopt plans of data/traj/towr.csv (solves #1, #2 of logs/towr_log.out), 6-digit CSV", "iters": [int(v) for v in r["iters"]],
                "iters_logged": log["iters"], "com_gap_m": com, "feet_gap_m": feet, "tolerance_m": 1e-3}
    except Exception as e:                                  # fixtures missing: say so, never guess
        return {"error": repr(e)}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import qtos_b200 as Q
    from qtos_b200 import heightfield as HF, parallel, workloads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the solver has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        host_group = dist.new_group(backend="gloo")          # host-side line-up before every NCCL all-gather (parallel.select_best_device)
    dev = torch.device("cuda", local)
    n_total = PER_GPU * world
    idx = parallel.shard_indices(n_total, rank, world)
    n = len(idx)
    opts = Q.default_options()                               # QTOS_ALG_IPOPT

    ctxs = [Q.Solver(Q.default_shape(COMBO, DURATION), device=local, max_batch=PER_GPU)]
    S = ctxs[0]
    if not is_sweep(world):
        grid, res, p_all = build_workload(n_total)
        for c in ctxs:
            hid = c.upload_heightfield(grid, res)
        p = np.ascontiguousarray(p_all[idx]); p["hf_id"] = hid
    else:
        variants = workloads.terrain_variants(N_VARIANTS)
        for c in ctxs:
            hids = [c.upload_heightfield(g, r_) for g, r_ in variants]
        p0 = workloads.replan_sweep_problems(n_total, variants, hids, group_size=GROUP)[idx]
        r0, x0, _ = S.solve(p0, opts)                        # generation 0 (untimed): the plans the timed windows start from
        p = np.ascontiguousarray(workloads.replan_from_rows(p0, S.sample_rows(p0, x0, -1)[:, 0]))
    streams = [torch.cuda.ExternalStream(S.stream, device=dev)] * max(IN_FLIGHT, CSV_DEPTH)
    p_all_groups = n_total // GROUP                          # group ids are dense over the whole job: 8 candidates per group

    # device-resident inputs/outputs for `value`; pinned host buffers for `e2e`
    d_p = torch.from_numpy(p.view(np.uint8).reshape(n, -1)).to(dev)
    d_res = [torch.zeros((n, Q.RESULT_DTYPE.itemsize), dtype=torch.uint8, device=dev) for _ in range(IN_FLIGHT)]
    d_x = [torch.zeros((n, S.n_vars), dtype=torch.float64, device=dev) for _ in range(IN_FLIGHT)]
    h_p = torch.from_numpy(p.view(np.uint8).reshape(n, -1)).pin_memory()
    h_res = [torch.empty(n * Q.RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory() for _ in range(IN_FLIGHT)]
    h_x = [torch.empty((n, S.n_vars), dtype=torch.float64).pin_memory() for _ in range(IN_FLIGHT)]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def finish(r):
        """the step's result: per-candidate records -> best plan per group (all-gather when N > 1)"""
        rec = parallel.make_records(r, idx, p["group"])
        winners, _ = parallel.select_best(rec, device=dev, host_group=host_group)
        return winners

    tickets = {}

    def submit_device(k):                                    # streaming, inputs and outputs resident in HBM
        tickets[k] = S.stream_submit_device(d_p.data_ptr(), n, d_res[k].data_ptr(), d_x[k].data_ptr())

    d_group = torch.from_numpy(np.ascontiguousarray(p["group"], dtype=np.int32)).to(dev)
    n_groups = int(p_all_groups)

    def collect_device(k):
        S.stream_wait(tickets.pop(k))
        # the step's result: best plan per group, selected on the device (k_records -> all-gather -> k_select), 32 KB of winners
        win = parallel.select_best_device(S, d_res[k], d_group, rank, world, n_groups, host_group).cpu()
        assert int((win < 0).sum()) == 0
        return d_res[k].cpu().numpy().view(Q.RESULT_DTYPE).reshape(n)     # statuses / iteration counts for the line's statistics

    def submit_host(k):                                      # streaming, pinned host buffers through the public API
        pp = h_p.numpy().view(Q.PROBLEM_DTYPE).reshape(n)
        tickets[k] = S.stream_submit(pp, (h_res[k].numpy().view(Q.RESULT_DTYPE).reshape(n), h_x[k].numpy()))

    def collect_host(k):
        r, x = S.stream_wait(tickets.pop(k))
        finish(r)
        return r

    def submit_serial(k):
        S.solve_device(d_p.data_ptr(), n, opts, d_res[k].data_ptr(), d_x[k].data_ptr())

    def collect_serial(k):
        r = d_res[k].cpu().numpy().view(Q.RESULT_DTYPE).reshape(n)
        finish(r)
        return r

    def timed(steps, submit, collect, depth, on_result=None):
        """`steps` steps with `depth` batches in flight; returns (ms per step = max of device time and wall time, converged)"""
        barrier()
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(depth)]
        t0 = time.perf_counter()
        ev0.record(streams[0])
        conv = 0; inflight = []
        for s in range(steps):
            k = s % depth
            if len(inflight) == depth:
                r = collect(inflight.pop(0)); conv += int((r["status"] == 0).sum())
                if on_result:
                    on_result(r, k)
            submit(k); inflight.append(k)
        while inflight:
            k = inflight.pop(0)
            r = collect(k); conv += int((r["status"] == 0).sum())
            if on_result:
                on_result(r, k)
        for k in range(depth):
            ev1[k].record(streams[k])
        barrier()
        wall = time.perf_counter() - t0
        dev_ms = max(ev0.elapsed_time(e) for e in ev1)
        return max(dev_ms, 1e3 * wall) / steps, conv

    def reduce_max_sum(ms, conv):
        t = torch.tensor([ms, float(conv)], dtype=torch.float64, device=dev)
        if world > 1:
            a = t.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
            b = t.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
            return float(a[0]), float(b[1])
        return ms, float(conv)

    for _ in range(max(args.warmup, 3)):
        submit_serial(0); collect_serial(0)

    # ---- serial pass with per-kernel events (synchronous batch call, one step at a time): phase times
    prof_steps = min(args.steps, 3)
    stats = {"fact": 0, "launches": 0, "phase": {}, "iters": []}

    def on_prof(r, k):
        st = S.last_stats()
        stats["fact"] += st["factorizations"]; stats["launches"] += st["factor_launches"]
        for name, v in st["ms"].items():
            stats["phase"][name] = stats["phase"].get(name, 0.0) + v
        stats["iters"].append(r["iters"].copy())

    S.set_profiling(True)
    serial_ms, serial_conv = timed(prof_steps, submit_serial, collect_serial, 1, on_prof)
    S.set_profiling(False)
    serial_ms, serial_conv = reduce_max_sum(serial_ms, serial_conv)

    # ---- the measurement: K steps through the streaming session, IN_FLIGHT jobs queued, inputs resident in HBM
    S.stream_begin(opts)
    timed(IN_FLIGHT + 1, submit_device, collect_device, IN_FLIGHT)      # the pool is warm and full
    clk_path = os.path.join(tempfile.gettempdir(), "qtos_clocks_%d.csv" % rank)
    sampler = clocks_sampler(clk_path, local) if rank == 0 else None
    info0 = S.stream_info()
    step_ms, conv = timed(args.steps, submit_device, collect_device, IN_FLIGHT)
    info1 = S.stream_info()
    launches = info1["launches"] - info0["launches"]
    if sampler is not None:
        sampler.terminate()
    step_ms, conv_all = reduce_max_sum(step_ms, conv)
    value = conv_all / args.steps / (step_ms * 1e-3)

    # ---- e2e: host buffers through the public API (H2D problems from pinned memory, D2H results + node values)
    timed(IN_FLIGHT, submit_host, collect_host, IN_FLIGHT)
    e_ms, conv_e = timed(args.steps, submit_host, collect_host, IN_FLIGHT)
    e_ms, conv_e_all = reduce_max_sum(e_ms, conv_e)
    e2e_value = conv_e_all / args.steps / (e_ms * 1e-3)

    # ---- e2e with the reference's actual output: the 1 kHz rows of every plan (what ./main writes to traj.csv) delivered to
    #      page-locked host memory by the same session (qtos_stream_submit_csv), 0.59 MB per window over PCIe
    e2e_csv_stream = None
    if world == 1:
        # a job ends with its slowest window (~70 batch iterations after admission), so throughput needs several jobs in flight:
        # half-steps (2048 windows, 1.2 GB of rows each) are queued, as many windows as the headline measurement keeps in flight
        # (a job's last window ends ~0.8 s after its admission; at 40k solves/s that is ~16 jobs of 2048 windows -- 1.2 GB of page-locked rows each)
        n_sub = n // 2
        h_rows = []
        for _ in range(CSV_DEPTH):
            try:
                h_rows.append(torch.empty((n_sub, S.csv_rows, Q.CSV_COLS), dtype=torch.float64).pin_memory())
            except RuntimeError:                                # the host does not lock that much memory: fewer jobs in flight
                break
        depth_c = len(h_rows)
        h_res_c = [torch.empty(n_sub * Q.RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory() for _ in range(depth_c)]
        h_x_c = [torch.empty((n_sub, S.n_vars), dtype=torch.float64).pin_memory() for _ in range(depth_c)]
        pp_all = h_p.numpy().view(Q.PROBLEM_DTYPE).reshape(n)

        def submit_csv(k):
            half = slice((k % 2) * n_sub, (k % 2 + 1) * n_sub)
            tickets[k] = S.stream_submit(pp_all[half], (h_res_c[k].numpy().view(Q.RESULT_DTYPE).reshape(n_sub), h_x_c[k].numpy()),
                                         csv_out=h_rows[k].numpy())

        def collect_csv(k):
            r, x = S.stream_wait(tickets.pop(k))
            return r

        timed(depth_c, submit_csv, collect_csv, depth_c)
        c_steps = max(4, args.steps)
        c_ms, conv_c = timed(2 * c_steps, submit_csv, collect_csv, depth_c)
        c_ms *= 2.0                                             # two half-steps per 4096-window step
        rows_ok = bool(np.isfinite(h_rows[0].numpy()[::64, ::50]).all()) and float(h_rows[0][0, -1, 0]) > 0.0
        e2e_csv_stream = {"value": conv_c / c_steps / (c_ms * 1e-3), "ms_per_step": c_ms, "steps": c_steps, "jobs_queued": depth_c,
                          "windows_per_job": n_sub, "rows_finite": rows_ok}
        del h_rows
    S.stream_end()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rank 0 extras: CSV-returning e2e, single-window latency, FAST algorithm, shape S5, golden parity, CPU arm
    rows_bytes = n * S.csv_rows * Q.CSV_COLS * 8
    t0 = time.perf_counter()
    r_c, x_c, rows_c = S.solve(p[:512], opts, csv=True)       # the reference's actual output: 1 kHz rows, 0.59 MB per window
    e2e_csv = int((r_c["status"] == 0).sum()) / (time.perf_counter() - t0)
    del rows_c
    S1 = Q.Solver(Q.default_shape(COMBO, DURATION), device=local, max_batch=1)
    if not is_sweep(world):
        S1.upload_heightfield(grid, res)
    else:
        for g, r_ in variants:
            S1.upload_heightfield(g, r_)
    lat = []
    for i in range(36):
        t0 = time.perf_counter(); S1.solve(p[i:i + 1], opts); lat.append(time.perf_counter() - t0)
    lat = sorted(lat[4:])
    S1.close()
    fast = None; s5 = None
    if not is_sweep(world):
        of = Q.default_options(algorithm=Q.ALG_FAST)
        S.solve(p, of)
        t0 = time.perf_counter(); rf, _, _ = S.solve(p, of); dtf = time.perf_counter() - t0
        fast = {"value": float((rf["status"] == 0).sum() / dtf), "unit": "solves/s", "converged_fraction": float((rf["status"] == 0).mean()),
                "note": "QTOS_ALG_FAST (sigma I Hessian, monotone mu, l1 merit): feasible plans, not Ipopt's plans; one batch in flight, host buffers"}
    fp64_peak = S.fp64_peak_tflops()
    if not is_sweep(world):
        S5 = Q.Solver(Q.default_shape("Custom", 5.0), device=local, max_batch=PER_GPU)
        p5 = p.copy(); p5["hf_id"] = S5.upload_heightfield(grid, res)
        S5.solve(p5[:256], opts)
        S5.set_profiling(True)
        t0 = time.perf_counter(); r5, _, _ = S5.solve(p5, opts); dt5 = time.perf_counter() - t0
        st5 = S5.last_stats()
        s5 = {"workload": "the same 4096 start/goal pairs with the production shape S5 (Custom gait, T=5s, 1040 vars / 1730 cons)",
              "value": float((r5["status"] == 0).sum() / dt5), "unit": "solves/s", "ms_per_step": 1e3 * dt5,
              "converged_fraction": float((r5["status"] == 0).mean()), "iters_mean": float(r5["iters"].mean()),
              "phase_ms_per_step": st5["ms"], "factorizations": int(st5["factorizations"]),
              "factor_tflops": st5["factorizations"] * S5.dims.flops_factor / (st5["ms"]["factor"] * 1e-3) / 1e12 if st5["ms"]["factor"] > 0 else None,
              "note": "synchronous batch call, host buffers, one step"}
        # the same shape through the streaming session: six jobs of 4096 windows queued at once
        S5.set_profiling(False)
        bufs = [(np.zeros(len(p5), dtype=Q.RESULT_DTYPE), np.zeros((len(p5), S5.n_vars))) for _ in range(6)]
        S5.stream_begin(opts)
        t0 = time.perf_counter()
        tk = [S5.stream_submit(p5, b) for b in bufs]
        conv5 = sum(int((S5.stream_wait(t)[0]["status"] == 0).sum()) for t in tk)
        dts = time.perf_counter() - t0
        i5 = S5.stream_info()
        S5.stream_end()
        s5["streaming"] = {"value": conv5 / dts, "unit": "solves/s", "ms_per_step": 1e3 * dts / len(bufs), "converged_fraction": conv5 / (len(bufs) * len(p5)),
                           "factor_tflops": i5["timed_slot_iterations"] * S5.dims.flops_factor / (i5["factor_ms"] * 1e-3) / 1e12 if i5["factor_ms"] > 0 else None,
                           "note": "six 4096-window jobs queued at once (pageable host buffers), pool of 4096 slots; flops = sum of w_i^2 over the scalar envelope of the compiled ordering (dims.flops_factor), factorization only"}
        S5.close()
        S51 = Q.Solver(Q.default_shape("Custom", 5.0), device=local, max_batch=1)
        q5 = p[:36].copy(); q5["hf_id"] = S51.upload_heightfield(grid, res)
        lat5 = []
        for i in range(36):
            t0 = time.perf_counter(); S51.solve(q5[i:i + 1], opts); lat5.append(time.perf_counter() - t0)
        lat5 = sorted(lat5[4:])
        s5["p50_latency_ms"] = 1e3 * lat5[len(lat5) // 2]; s5["p99_latency_ms"] = 1e3 * lat5[-1]
        S51.close()
    parity = golden_parity(Q, local)
    import oracle as O
    O.build()
    cpu_v, cpu_cores, cpu_p50, cpu_conv = cpu_arm(cpu_sample_jobs(world, CPU_SAMPLE))

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    traffic = {}
    try:
        traffic = json.load(open(TRAFFIC_JSON))
    except Exception:
        pass
    dims = S.dims
    ph = {k: v / prof_steps for k, v in stats["phase"].items()}
    # roofline: the streaming session's own CUDA events around k_factor / kip_solve, over the timed steps
    fact = info1["timed_slot_iterations"] - info0["timed_slot_iterations"]
    fl = max(1, info1["timed_iterations"] - info0["timed_iterations"])
    f_ms, s_ms = info1["factor_ms"] - info0["factor_ms"], info1["solve_ms"] - info0["solve_ms"]
    pool_occupancy = (info1["slot_iterations"] - info0["slot_iterations"]) / max(1, info1["iterations"] - info0["iterations"]) / PER_GPU
    flop_fact = ALG_FLOP_FACTOR + ALG_FLOP_RHS
    achieved = fact * flop_fact / (f_ms * 1e-3) / 1e12 if f_ms > 0 else None
    # kip_solve: algorithmic bytes per problem = the factor read once per sweep (2 n_refine + 1 sweeps: stored blocks + inverses of
    # the diagonal blocks) + the Jacobian values read once per J product (n_refine + 1 row products, n_refine gathers, 1 expansion)
    # + Q read once per Woodbury correction (n_refine + 1) -- DESIGN.md section 5
    nref = opts.n_refine
    l_bytes = (dims.kkt_blocks + dims.kkt_order // dims.kkt_block) * dims.kkt_block * dims.kkt_block * 8
    j_bytes = dims.nnz_jac * 8
    q_bytes = 12 * dims.kkt_order * 8
    solve_bytes = (2 * nref + 1) * l_bytes + (2 * nref + 2) * j_bytes + (nref + 1) * q_bytes
    hbm = peaks.get("hbm_gbs")
    solve_gbs = fact * solve_bytes / (s_ms * 1e-3) / 1e9 if s_ms > 0 else None
    iters = np.concatenate(stats["iters"])
    tr_f = traffic.get("k_factor_ipopt", {}).get("dram_bytes_per_problem")
    tr_s = traffic.get("kip_solve", {}).get("dram_bytes_per_problem")
    line = {
        "metric": "converged gait-plan NLP solves/sec", "value": value, "unit": "solves/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(world), "problems_per_gpu": PER_GPU, "parallelism": "shard%d" % world, "algorithm": "ipopt",
                   "mode": "continuous batching (qtos_stream_*): pool of %d slots, %d jobs queued" % (PER_GPU, IN_FLIGHT), "pool_occupancy": pool_occupancy,
                   "pool_occupancy_note": "occupied slots averaged over the batch ITERATIONS of the timed region, the ~70 near-empty iterations that drain the last job included; by time the pool is full except for that drain",
                   "l2": "per-step working set %.1f GB per GPU > 126 MB L2" % (PER_GPU * dims.workspace_bytes_per_problem / 1e9)},
        "converged_fraction": conv_all / (args.steps * n_total), "iters_mean": float(iters.mean()), "iters_max": int(iters.max()),
        "p50_latency_ms": 1e3 * lat[len(lat) // 2], "p99_latency_ms": 1e3 * lat[-1],
        "serial": {"value": serial_conv / prof_steps / (serial_ms * 1e-3), "ms_per_step": serial_ms,
                   "note": "the same steps through the synchronous batch call, one at a time (straggler iterations exposed)"},
        "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": int(n * Q.PROBLEM_DTYPE.itemsize),
                "d2h_bytes_per_step": int(n * (Q.RESULT_DTYPE.itemsize + 8 * S.n_vars)), "ms_per_step": e_ms,
                "returns": "per-window status records + spline node values (the plan); the reference's 1 kHz CSV rows are sampled on demand",
                "e2e_csv": {"value": e2e_csv_stream["value"] if e2e_csv_stream else e2e_csv, "unit": "solves/s",
                            "d2h_bytes_per_window": int(S.csv_rows * Q.CSV_COLS * 8 + Q.RESULT_DTYPE.itemsize + 8 * S.n_vars),
                            "d2h_bytes_per_step": int(rows_bytes + n * (Q.RESULT_DTYPE.itemsize + 8 * S.n_vars)),
                            "streaming": e2e_csv_stream,
                            "synchronous_call": {"value": e2e_csv, "note": "512 windows through Solver.solve(csv=True), rows into a fresh pageable array"},
                            "note": ("the reference's actual output: every window's 1 kHz rows (%.1f GB per 4096-window step) delivered to page-locked host memory "
                                     "by the streaming session (qtos_stream_submit_csv), the copies overlapping the next job's iterations" % (rows_bytes / 1e9))
                                    if e2e_csv_stream else "512 windows with all 1 kHz rows copied to pageable host memory"}},
        "gpu_launches": int(launches),
        "phase_ms_per_step": ph,
        "roofline": {"kernel": "k_factor<.,1> (block-skyline Cholesky of the condensed KKT matrix + forward substitution of 16 right-hand sides)",
                     "bound": "tensor",
                     "bound_detail": "FP64: 16x16 block updates on the FP64 tensor-core path (DMMA m8n8k4); SURVEY 8(d) names the "
                                     "FP64 FMA rate as the bounding roofline, and B200's FP64 tensor and FMA peaks coincide",
                     "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": (achieved / fp64_peak) if achieved else None,
                     "peak_source": "FP64 FMA loop measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
                     "alg_flop_per_factorization": flop_fact, "factorizations_per_step": fact / args.steps,
                     "avg_launch_ms": f_ms / fl,
                     "traffic": (tr_f * fact / fl) if tr_f else None,
                     "traffic_source": "profiles/r02_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one 4096-problem launch) x problems per average launch",
                     "algorithmic_bytes": (2.0 * dims.kkt_blocks * 256 * 8 + 2.0 * 16 * dims.kkt_order * 8) * fact / fl,
                     "hbm_peak_gbs": hbm},
        "roofline_solve": {"kernel": "kip_solve (triangular sweeps over L, Woodbury term, multiplier passes)", "bound": "hbm",
                           "achieved": solve_gbs, "peak": hbm, "unit": "GB/s", "frac": (solve_gbs / hbm) if (solve_gbs and hbm) else None,
                           "algorithmic_bytes_per_problem": solve_bytes, "avg_launch_ms": s_ms / fl,
                           "traffic": (tr_s * fact / fl) if tr_s else None},
        "parity": parity,
        "cpu_baseline": {"value": cpu_v, "unit": "solves/s", "cores": cpu_cores, "kind": "port",
                         "sample": "%d windows of rank 0's shard (the same sample as --impl reference), one process per host core; %s" % (CPU_SAMPLE, CPU_KIND_NOTE),
                         "p50_latency_ms": 1e3 * cpu_p50, "converged": cpu_conv},
        "clocks": parse_clocks(clk_path),
    }
    if fast:
        line["fast_algorithm"] = fast
    if s5:
        line["shapes"] = {"S5": s5}
    _emit(line)
    S.close()
    if world > 1:
        dist.destroy_process_group()


def _emit(line):
    """the one JSON line goes to the process's real stdout; everything else that lands on fd 1 (NCCL's version banner,
    library chatter) was redirected to stderr by main()"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="qtos_b200", choices=["qtos_b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
