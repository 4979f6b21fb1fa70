#!/usr/bin/env python
"""bench.py -- converged gait-plan NLP solves/s on B200 (BASELINE.json metric), one JSON line.

  python bench.py [--gpus N --steps K --warmup W]          this repo's CUDA path
  python bench.py --impl reference [...]                   CPU arm: the oracle port of the reference
                                                           algorithm on the box's host cores (the reference
                                                           binary needs Eigen3/ifopt/Ipopt/MUMPS: unbuildable)

A step = one pass of the hot path over one batch: 4096 independent local-plan windows per GPU
(BASELINE.json configs[3]: trot gait C1, T = 2 s, 640 variables / 892 constraints, 256x256 rough
heightfield, seed 1234), solved to the reference's convergence criteria.  N > 1: one process per GPU
(torchrun), candidates sharded statically (weak scaling), one all-gather of per-candidate records for
best-plan selection inside the timed region.
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
COMBO, DURATION = "C1", 2.0
WORKLOAD = ("batched multi-start: %d start/goal pairs per GPU on a 256x256 rough heightfield (seed 1234), trot C1, "
            "T=2s, 640 vars / 892 cons; groups of 8 candidates, best-plan all-gather" % PER_GPU)
# algorithmic work of ONE factorization of the primal normal matrix of shape S2
# (n_free = 605, RCM envelope 50 728 entries): sum_i w_i^2 = 6.3e6 FP64 flop
ALG_FLOP_PER_FACTORIZATION = 6.3e6
# DRAM traffic of k_factor per factorization, from the ncu --set full capture of one launch with all 4096 problems
# active (profiles/r01k_summary.md: dram__bytes_read.sum + dram__bytes_write.sum over 4096 factorizations); the
# algorithmic bytes are the assembled matrix read once and the factor written once: 2 x 202 blocks x 256 doubles
# = 0.827 MB (taken from the compiled shape at run time)
NCU_DRAM_BYTES_PER_FACTORIZATION = 1.1702e6


def build_workload(n_total, seed=1234):
    from qtos_b200 import heightfield as HF, workloads
    grid, res = HF.rough_terrain(seed)
    p = workloads.multistart_problems(n_total, grid, res, seed=seed, group_size=GROUP)
    return grid, res, p


# ------------------------------------------------------------------ CPU arm (oracle port)

def _cpu_solve(args):
    import oracle as O
    rec, grid, res = args
    so = O.default_shape(COMBO, DURATION)
    inst = O.make_instance(start_pos=rec["start_pos"], start_ang=rec["start_ang"], goal=rec["goal"], ee=rec["ee"])
    po = O.Problem(so, inst, O.Terrain(grid, res))
    t = time.perf_counter()
    _, r = po.solve()
    return r.status, time.perf_counter() - t


def cpu_arm(n_sample, cores=None):
    """oracle IPM on the host cores, one process per core; returns (solves/s, cores, p50 latency s, converged)."""
    import multiprocessing as mp
    import oracle as O
    O.build()
    cores = cores or os.cpu_count()
    grid, res, p = build_workload(n_sample)
    jobs = [(p[i], grid, res) for i in range(n_sample)]
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_cpu_solve, jobs[:cores])                     # warm the workers
        t = time.perf_counter()
        out = pool.map(_cpu_solve, jobs, chunksize=1)
        dt = time.perf_counter() - t
    conv = sum(1 for s, _ in out if s == 0)
    lat = sorted(t_ for _, t_ in out)
    return conv / dt, cores, lat[len(lat) // 2], conv


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = 192
    for _ in range(min(args.warmup, 1)):
        cpu_arm(16)
    vals = []
    for _ in range(args.steps):
        v, cores, p50, conv = cpu_arm(n_sample)
        vals.append(v)
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": "converged gait-plan NLP solves/sec", "value": value, "unit": "solves/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n_sample / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "problems_per_gpu": PER_GPU, "parallelism": "host cores",
                       "sample": "%d of the %d windows per step" % (n_sample, PER_GPU)},
            "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": "port",
                             "sample": "%d windows of the bench workload per step, one process per core; oracle/towr_ipm.c "
                                       "(same IPM as the GPU path; TOWR+Ipopt itself cannot be built here)" % n_sample,
                             "p50_latency_ms": 1e3 * p50},
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


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import qtos_b200 as Q
    from qtos_b200 import parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the solver has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    n_total = PER_GPU * world
    grid, res, p_all = build_workload(n_total)
    idx = parallel.shard_indices(n_total, rank, world)
    p = np.ascontiguousarray(p_all[idx])
    S = Q.Solver(Q.default_shape(COMBO, DURATION), device=local, max_batch=PER_GPU)
    hid = S.upload_heightfield(grid, res)
    p["hf_id"] = hid
    opts = Q.default_options()
    n = len(p)
    stream = torch.cuda.ExternalStream(S.stream, device=dev)

    # device-resident inputs/outputs for `value`
    d_p = torch.from_numpy(p.view(np.uint8).reshape(n, -1)).to(dev)
    d_res = torch.zeros((n, Q.RESULT_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    d_x = torch.zeros((n, S.n_vars), dtype=torch.float64, device=dev)
    # pinned host buffers for `e2e`
    h_p = torch.from_numpy(p.view(np.uint8).reshape(n, -1)).pin_memory()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        S.solve_device(d_p.data_ptr(), n, opts, d_res.data_ptr(), d_x.data_ptr())
        r = d_res.cpu().numpy().view(Q.RESULT_DTYPE).reshape(n)      # 229 KB of records: the step's result
        rec = parallel.make_records(r, idx, p["group"])
        winners, _ = parallel.select_best(rec, device=dev)
        return r, winners

    h_res = torch.empty(n * Q.RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    h_x = torch.empty((n, S.n_vars), dtype=torch.float64).pin_memory()

    def step_e2e():
        pp = h_p.numpy().view(Q.PROBLEM_DTYPE).reshape(n)
        # host buffers through the public API: H2D problems from pinned memory, D2H results + node values into pinned memory
        r, x, _ = S.solve(pp, opts, out=(h_res.numpy().view(Q.RESULT_DTYPE).reshape(n), h_x.numpy()))
        rec = parallel.make_records(r, idx, p["group"])
        winners, _ = parallel.select_best(rec, device=dev)
        return r, x

    S.set_profiling(True)
    for _ in range(args.warmup):
        step_device()
    clk_path = os.path.join(tempfile.gettempdir(), "qtos_clocks_%d.csv" % rank)
    sampler = clocks_sampler(clk_path, local) if rank == 0 else None
    barrier()
    launches0 = S.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    conv_total = 0; fact = 0; fact_ms = 0.0; fact_launches = 0; phase = {}
    iters_hist = []
    t_wall = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        r, winners = step_device()
        conv_total += int((r["status"] == 0).sum())
        st = S.last_stats()
        fact += st["factorizations"]; fact_ms += st["ms"]["factor"]; fact_launches += st["factor_launches"]
        for k, v in st["ms"].items():
            phase[k] = phase.get(k, 0.0) + v
        iters_hist.append(r["iters"].copy())
    ev1.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = S.launch_count() - launches0
    if sampler is not None:
        sampler.terminate()
    dev_ms = ev0.elapsed_time(ev1)
    # device time covers the solver kernels; the host part of a step (record D2H + selection) is inside the
    # wall clock between the same barriers -- report the slower of the two so nothing is hidden
    step_ms = max(dev_ms, 1e3 * t_wall) / args.steps
    tm = torch.tensor([step_ms, float(conv_total)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = tm.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = tm.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        step_ms = float(tmax[0]); conv_all = float(tsum[1])
    else:
        conv_all = float(conv_total)
    value = conv_all / args.steps / (step_ms * 1e-3)

    # e2e: host buffers through the public API
    S.set_profiling(False)
    step_e2e()
    barrier()
    t0 = time.perf_counter(); conv_e = 0
    for _ in range(args.steps):
        r_e, x_e = step_e2e()
        conv_e += int((r_e["status"] == 0).sum())
    barrier()
    e_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    te = torch.tensor([e_ms, float(conv_e)], dtype=torch.float64, device=dev)
    if world > 1:
        a = te.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
        b = te.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
        e_ms = float(a[0]); conv_e_all = float(b[1])
    else:
        conv_e_all = float(conv_e)
    e2e_value = conv_e_all / args.steps / (e_ms * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # single-window latency (batch of 1, host call), p50 over 32 windows
    S1 = Q.Solver(Q.default_shape(COMBO, DURATION), device=local, max_batch=1)
    hid1 = S1.upload_heightfield(grid, res)
    lat = []
    for i in range(36):
        q = p[i:i + 1].copy(); q["hf_id"] = hid1
        t0 = time.perf_counter(); S1.solve(q, opts); lat.append(time.perf_counter() - t0)
    lat = sorted(lat[4:])
    fp64_peak = S.fp64_peak_tflops()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    achieved = fact * ALG_FLOP_PER_FACTORIZATION / (fact_ms * 1e-3) / 1e12 if fact_ms > 0 else None
    iters = np.concatenate(iters_hist)
    cpu_v, cpu_cores, cpu_p50, _ = cpu_arm(128)
    dims = S.dims
    line = {
        "metric": "converged gait-plan NLP solves/sec", "value": value, "unit": "solves/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "problems_per_gpu": PER_GPU, "parallelism": "shard%d" % world,
                   "l2": "per-step working set %.1f GB per GPU > 126 MB L2" % (PER_GPU * dims.workspace_bytes_per_problem / 1e9)},
        "converged_fraction": conv_all / (args.steps * n_total), "iters_mean": float(iters.mean()), "iters_max": int(iters.max()),
        "p50_latency_ms": 1e3 * lat[len(lat) // 2], "p99_latency_ms": 1e3 * lat[-1],
        "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": int(n * Q.PROBLEM_DTYPE.itemsize),
                "d2h_bytes_per_step": int(n * (Q.RESULT_DTYPE.itemsize + 8 * S.n_vars)), "ms_per_step": e_ms},
        "gpu_launches": int(launches),
        "phase_ms_per_step": {k: v / args.steps for k, v in phase.items()},
        "roofline": {"kernel": "k_factor (block-skyline Cholesky of the condensed KKT matrix)", "bound": "tensor",
                     "bound_detail": "FP64: 16x16 block updates on the FP64 tensor-core path (DMMA m8n8k4); SURVEY 8(d) names the "
                                     "FP64 FMA rate as the bounding roofline, and B200's FP64 tensor and FMA peaks coincide",
                     "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": (achieved / fp64_peak) if achieved else None,
                     "peak_source": "FP64 FMA loop measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
                     "alg_flop_per_factorization": ALG_FLOP_PER_FACTORIZATION, "factorizations_per_step": fact / args.steps,
                     "avg_launch_ms": fact_ms / max(1, fact_launches),
                     "traffic": NCU_DRAM_BYTES_PER_FACTORIZATION * fact / max(1, fact_launches),
                     "traffic_unit": "bytes per average launch (ncu dram bytes per factorization x factorizations per launch)",
                     "algorithmic_bytes": 2.0 * dims.kkt_blocks * dims.kkt_block * dims.kkt_block * 8 * fact / max(1, fact_launches),
                     "hbm_peak_gbs": peaks.get("hbm_gbs")},
        "cpu_baseline": {"value": cpu_v, "unit": "solves/s", "cores": cpu_cores, "kind": "port",
                         "sample": "128 windows of the same workload, one process per core, oracle/towr_ipm.c",
                         "p50_latency_ms": 1e3 * cpu_p50},
        "clocks": parse_clocks(clk_path),
    }
    _emit(line)
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
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="qtos_b200", choices=["qtos_b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
