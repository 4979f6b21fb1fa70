"""qtos_b200 -- B200-native batched gait-planning NLP solver (drop-in for the QTOS local planner).

Host side of the C ABI in include/qtos_b200.h (ctypes; no torch types cross the boundary).
The reference reaches this path as `docker exec <id> ./main <flags>` built by
QTOS/utils.py:15-26,644-670 and called from scripts/main.py:48-50,90-92 and
QTOS/generateHeightField.py:385-386; `towr_main` mirrors that command line, `Solver` is the
batched in-process form of it.

There is NO CPU fallback: importing works anywhere (so the CPU test tier can check symbols and
host logic), but every compute call needs the CUDA library and a GPU and raises otherwise.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_SO = os.path.join(_HERE, "libqtos_b200.so")
_ROOT = os.path.dirname(_HERE)
NEE = 4
CSV_COLS = 37
COMBOS = {"C0": 0, "C1": 1, "C2": 2, "C3": 3, "C4": 4, "Custom": 5}
STATUS_RUNNING = 99

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def _sources():
    return [os.path.join(_CSRC, f) for f in ("qtos_kernels.cu", "qtos_compile.cpp")]


def _deps():
    d = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cu", ".cuh", ".cpp", ".h", ".inc"))]
    d.append(os.path.join(_ROOT, "include", "qtos_b200.h"))
    return d


def build(force=False, verbose=False):
    """Compile csrc/ for sm_100a with nvcc into libqtos_b200.so (in-tree)."""
    if os.environ.get("QTOS_LIB"):          # development: load an experimental build of the same sources
        return os.environ["QTOS_LIB"]
    import fcntl
    with open(os.path.join(_HERE, ".build.lock"), "w") as lock:      # one process of a torchrun group compiles, the others wait
        fcntl.flock(lock, fcntl.LOCK_EX)
        stale = force or not os.path.exists(_SO) or any(
            os.path.getmtime(s) > os.path.getmtime(_SO) for s in _deps() if os.path.exists(s))
        if stale:
            nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
            if not os.path.exists(nvcc):
                if os.path.exists(_SO):
                    return _SO          # GPU box without sources newer than the shipped library
                raise RuntimeError("nvcc not found and libqtos_b200.so is missing")
            tmp = _SO + ".tmp%d" % os.getpid()
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + _sources() + ["-o", tmp]
            subprocess.check_call(cmd, cwd=_CSRC)
            os.replace(tmp, _SO)        # readers never see a half-written library
    return _SO


class Shape(C.Structure):
    _fields_ = [("mass", C.c_double), ("I_b", C.c_double * 9),
                ("nominal", (C.c_double * 3) * NEE), ("max_dev", C.c_double * 3),
                ("mu", C.c_double), ("force_limit", C.c_double), ("t_swing_avg", C.c_double),
                ("dt_base_poly", C.c_double), ("force_polys_per_stance", C.c_int),
                ("ee_polys_per_swing", C.c_int), ("dt_dynamic", C.c_double),
                ("dt_rom", C.c_double), ("combo", C.c_int), ("duration", C.c_double),
                ("base_rom", C.c_int), ("dt_base_rom", C.c_double),
                ("cost_force_z", C.c_double), ("cost_ee_vel_xy", C.c_double), ("terrain_gradients", C.c_int)]


class Options(C.Structure):
    _fields_ = [("tol", C.c_double), ("constr_viol_tol", C.c_double), ("compl_inf_tol", C.c_double),
                ("dual_inf_tol", C.c_double), ("max_iter", C.c_int), ("mu_init", C.c_double),
                ("sigma_w", C.c_double), ("delta_c", C.c_double), ("feas_exit", C.c_int),
                ("algorithm", C.c_int), ("n_refine", C.c_int), ("lm_history", C.c_int), ("max_cpu_time", C.c_double),
                ("lm_init_val_min", C.c_double), ("retry_failed", C.c_int), ("retry_lm_init_val_min", C.c_double)]


ALG_IPOPT, ALG_FAST = 0, 1          # qtos_options.algorithm
TRACE_ITERS, TRACE_COLS = 48, 8


class Dims(C.Structure):
    _fields_ = [("n_vars", C.c_int), ("n_cons", C.c_int), ("n_free", C.c_int), ("n_eq", C.c_int),
                ("n_ineq", C.c_int), ("nnz_jac", C.c_int), ("csv_rows", C.c_int),
                ("kkt_order", C.c_int), ("kkt_block", C.c_int), ("kkt_blocks", C.c_int),
                ("flops_factor", C.c_double), ("workspace_bytes_per_problem", C.c_longlong)]


class StreamInfo(C.Structure):
    _fields_ = [("iterations", C.c_longlong), ("slot_iterations", C.c_longlong), ("windows_done", C.c_longlong),
                ("launches", C.c_longlong), ("factor_ms", C.c_double), ("solve_ms", C.c_double),
                ("timed_iterations", C.c_longlong), ("timed_slot_iterations", C.c_longlong), ("retried", C.c_longlong)]


class Stats(C.Structure):
    _fields_ = [("ms", C.c_float * 8), ("factorizations", C.c_longlong), ("factor_launches", C.c_longlong),
                ("iterations", C.c_int), ("retried", C.c_int)]


# numpy mirrors of qtos_problem / qtos_result (C layout, checked against ctypes sizes below)
PROBLEM_DTYPE = np.dtype([("start_pos", "f8", 3), ("start_ang", "f8", 3), ("start_vel", "f8", 3),
                          ("start_ang_vel", "f8", 3), ("goal", "f8", 3), ("ee", "f8", (NEE, 3)),
                          ("t_start", "f8"), ("hf_id", "i4"), ("group", "i4")], align=True)
RESULT_DTYPE = np.dtype([("status", "i4"), ("iters", "i4"), ("constr_viol", "f8"), ("dual_inf", "f8"),
                         ("compl_inf", "f8"), ("nlp_error", "f8"), ("mu", "f8"), ("cost", "f8")], align=True)
assert PROBLEM_DTYPE.itemsize == 8 * 28 + 8 and RESULT_DTYPE.itemsize == 56

EXPORTS = ["qtos_default_shape", "qtos_default_options", "qtos_create", "qtos_destroy", "qtos_last_error",
           "qtos_get_dims", "qtos_upload_heightfield", "qtos_free_heightfield", "qtos_heightfield_query", "qtos_heightfield_cells", "qtos_heightfield_gradients",
           "qtos_get_initial", "qtos_eval", "qtos_solve_batch", "qtos_solve_batch_device", "qtos_solve_batch_async",
           "qtos_solve_batch_device_async", "qtos_wait", "qtos_stream_begin", "qtos_stream_submit", "qtos_stream_submit_csv", "qtos_stream_submit_device",
           "qtos_stream_wait", "qtos_stream_stats", "qtos_stream_end", "qtos_make_records", "qtos_select_best", "qtos_get_trace", "qtos_sample_csv", "qtos_sample_csv_rows",
           "qtos_write_csv", "qtos_launch_count", "qtos_set_profiling", "qtos_last_stats", "qtos_stream",
           "qtos_measure_fp64_peak", "qtos_measure_heightfield_staging", "qtos_assembly_table_stats"]

_LIB = None


def lib():
    """Load libqtos_b200.so (building it when nvcc and newer sources are present)."""
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        vp = C.c_void_p
        L.qtos_default_shape.argtypes = [C.POINTER(Shape)]
        L.qtos_default_options.argtypes = [C.POINTER(Options)]
        L.qtos_create.argtypes = [C.c_int, C.POINTER(Shape), C.c_int, C.POINTER(vp)]
        L.qtos_destroy.argtypes = [vp]
        L.qtos_last_error.argtypes = [vp]
        L.qtos_last_error.restype = C.c_char_p
        L.qtos_get_dims.argtypes = [vp, C.POINTER(Dims)]
        L.qtos_upload_heightfield.argtypes = [vp, dp, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_int)]
        L.qtos_free_heightfield.argtypes = [vp, C.c_int]
        L.qtos_heightfield_query.argtypes = [vp, C.c_int, dp, C.c_int, dp]
        L.qtos_heightfield_cells.argtypes = [vp, C.c_int, dp, C.c_int, C.POINTER(C.c_longlong)]
        L.qtos_heightfield_gradients.argtypes = [vp, C.c_int, dp, C.c_int, dp, dp]
        L.qtos_get_initial.argtypes = [vp, vp, C.c_int, dp, dp, dp, dp, dp]
        L.qtos_eval.argtypes = [vp, vp, C.c_int, dp, dp, dp]
        L.qtos_solve_batch.argtypes = [vp, vp, C.c_int, C.POINTER(Options), vp, dp, dp]
        L.qtos_solve_batch_device.argtypes = [vp, vp, C.c_int, C.POINTER(Options), vp, vp]
        L.qtos_solve_batch_async.argtypes = [vp, vp, C.c_int, C.POINTER(Options), vp, dp, dp]
        L.qtos_solve_batch_device_async.argtypes = [vp, vp, C.c_int, C.POINTER(Options), vp, vp]
        L.qtos_wait.argtypes = [vp]
        L.qtos_stream_begin.argtypes = [vp, C.POINTER(Options)]
        L.qtos_stream_submit.argtypes = [vp, vp, C.c_int, vp, dp, C.POINTER(C.c_int)]
        L.qtos_stream_submit_csv.argtypes = [vp, vp, C.c_int, vp, dp, dp, C.POINTER(C.c_int)]
        L.qtos_stream_submit_device.argtypes = [vp, vp, C.c_int, vp, vp, C.POINTER(C.c_int)]
        L.qtos_stream_wait.argtypes = [vp, C.c_int]
        L.qtos_stream_stats.argtypes = [vp, C.POINTER(StreamInfo)]
        L.qtos_stream_end.argtypes = [vp]
        L.qtos_make_records.argtypes = [vp, vp, vp, C.c_longlong, C.c_longlong, C.c_int, vp]
        L.qtos_select_best.argtypes = [vp, vp, C.c_int, C.c_int, vp]
        L.qtos_get_trace.argtypes = [vp, C.c_int, dp]
        L.qtos_sample_csv.argtypes = [vp, vp, C.c_int, dp, dp]
        L.qtos_sample_csv_rows.argtypes = [vp, vp, C.c_int, dp, C.c_int, C.c_int, dp]
        L.qtos_write_csv.argtypes = [dp, C.c_int, C.c_char_p]
        L.qtos_launch_count.argtypes = [vp]
        L.qtos_launch_count.restype = C.c_longlong
        L.qtos_set_profiling.argtypes = [vp, C.c_int]
        L.qtos_last_stats.argtypes = [vp, C.POINTER(Stats)]
        L.qtos_stream.argtypes = [vp]
        L.qtos_stream.restype = vp
        L.qtos_measure_fp64_peak.argtypes = [vp, dp]
        L.qtos_measure_heightfield_staging.argtypes = [vp, C.c_int, dp, C.c_int, C.c_int, dp, dp, dp, C.POINTER(C.c_int)]
        L.qtos_assembly_table_stats.argtypes = [C.POINTER(Shape), C.c_int, C.POINTER(C.c_ulonglong)]
        _LIB = L
    return _LIB


def default_shape(combo="Custom", duration=5.0, mass=None):
    """Solo12 constants as vendored in the reference (m = 1.5, F5 quirk inertia tensor)."""
    s = Shape()
    lib().qtos_default_shape(C.byref(s))
    s.combo = COMBOS[combo] if isinstance(combo, str) else int(combo)
    s.duration = float(duration)
    if mass is not None:
        s.mass = float(mass)
    return s


def assembly_table_stats(shape, rows_dealt=-1):
    """The assembly term streams the shape compiler builds for k_asm, inspected on the host (no device needed): terms, slots,
    slots of the slowest warp summed over the block rows, steps holding a target twice, an order-sensitive hash of every target's
    terms, whether the panel rows were dealt by term count (rows_dealt: 0 / 1 force a deal, -1 = what Solver would use)."""
    out = (C.c_ulonglong * 6)()
    rc = lib().qtos_assembly_table_stats(C.byref(shape), int(rows_dealt), out)
    if rc:
        raise RuntimeError("qtos_assembly_table_stats failed with %d" % rc)
    return dict(zip(("terms", "slots", "slowest_warp_slots", "duplicate_targets", "order_hash", "rows_dealt"), (int(v) for v in out)))


def default_options(**kw):
    o = Options()
    lib().qtos_default_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


NOMINAL_FEET = ((0.21, 0.18, 0.0), (0.21, -0.18, 0.0), (-0.21, 0.18, 0.0), (-0.21, -0.18, 0.0))


def make_problems(n):
    """n default windows: the no-flag defaults of ./main (ref: solver/towr/src/main.cpp:306-346)."""
    p = np.zeros(n, dtype=PROBLEM_DTYPE)
    p["start_pos"] = (0.0, 0.0, 0.24)
    p["goal"] = (0.5, 0.0, 0.24)
    p["ee"] = np.array(NOMINAL_FEET)
    return p


class QtosError(RuntimeError):
    pass


class Solver:
    """One device, one compiled shape, workspace for `max_batch` concurrent problems."""

    def __init__(self, shape=None, device=0, max_batch=1):
        self._L = lib()
        self.shape = shape if shape is not None else default_shape()
        self._h = C.c_void_p()
        rc = self._L.qtos_create(int(device), C.byref(self.shape), int(max_batch), C.byref(self._h))
        if rc != 0:
            raise QtosError("qtos_create failed (%d): %s" % (rc, self._L.qtos_last_error(None).decode()))
        self.device, self.max_batch = device, max_batch
        d = Dims()
        self._L.qtos_get_dims(self._h, C.byref(d))
        self.dims = d
        self.n_vars, self.n_cons, self.csv_rows = d.n_vars, d.n_cons, d.csv_rows

    def close(self):
        if getattr(self, "_h", None):
            self._L.qtos_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise QtosError("qtos error %d: %s" % (rc, self._L.qtos_last_error(self._h).decode()))

    @staticmethod
    def _probs(p):
        p = np.ascontiguousarray(p, dtype=PROBLEM_DTYPE)
        return p, p.ctypes.data_as(C.c_void_p)

    def upload_heightfield(self, grid, res):
        """grid[ix, iy] as parsed from towr_heightfield.txt (row = world x)."""
        g = np.ascontiguousarray(grid, dtype=np.float64)
        if g.ndim != 2:
            raise ValueError("heightfield must be 2-D")
        hid = C.c_int(-1)
        self._ck(self._L.qtos_upload_heightfield(self._h, _dp(g), g.shape[0], g.shape[1], float(res), C.byref(hid)))
        return hid.value

    def free_heightfield(self, hf_id):
        """release a grid on the device; the id is handed out again by a later upload"""
        self._ck(self._L.qtos_free_heightfield(self._h, int(hf_id)))

    def height(self, hf_id, xy):
        xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        out = np.zeros(len(xy))
        self._ck(self._L.qtos_heightfield_query(self._h, int(hf_id), _dp(xy), len(xy), _dp(out)))
        return out

    def height_cells(self, hf_id, xy):
        xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        out = np.zeros((len(xy), 4), dtype=np.int64)
        self._ck(self._L.qtos_heightfield_cells(self._h, int(hf_id), _dp(xy), len(xy), out.ctypes.data_as(C.POINTER(C.c_longlong))))
        return out

    def heightfield_gradients(self, hf_id, xy):
        """(dh/dx, dh/dy) of the bilinear surface at the points xy[n, 2] (the derivative the reference carries commented out)"""
        xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        hx, hy = np.zeros(len(xy)), np.zeros(len(xy))
        self._ck(self._L.qtos_heightfield_gradients(self._h, int(hf_id), _dp(xy), len(xy), _dp(hx), _dp(hy)))
        return hx, hy

    def initial(self, problems):
        p, pp = self._probs(problems)
        n = len(p)
        x0, xl, xu = (np.zeros((n, self.n_vars)) for _ in range(3))
        gl, gu = (np.zeros((n, self.n_cons)) for _ in range(2))
        self._ck(self._L.qtos_get_initial(self._h, pp, n, _dp(x0), _dp(xl), _dp(xu), _dp(gl), _dp(gu)))
        return x0, xl, xu, gl, gu

    def eval(self, problems, x=None, jac=True):
        p, pp = self._probs(problems)
        n = len(p)
        xx = None if x is None else np.ascontiguousarray(x, dtype=np.float64).reshape(n, self.n_vars)
        g = np.zeros((n, self.n_cons))
        J = np.zeros((n, self.n_cons, self.n_vars)) if jac else None
        self._ck(self._L.qtos_eval(self._h, pp, n, _dp(xx), _dp(g), _dp(J)))
        return (g, J) if jac else g

    def solve(self, problems, options=None, csv=False, out=None):
        """Solve a batch of windows.  Returns (results[RESULT_DTYPE], x[n, n_vars], csv or None).
        `out` = (results, x) lets the caller supply the host buffers the C ABI writes into -- e.g. page-locked
        memory, which the device-to-host copies then reach at full PCIe speed (a fresh pageable array is faulted in
        page by page under the copy)."""
        p, pp = self._probs(problems)
        n = len(p)
        o = options if options is not None else default_options()
        if out is not None:
            res, x = out
            if res.dtype != RESULT_DTYPE or res.shape != (n,) or not res.flags.c_contiguous:
                raise ValueError("out[0] must be a C-contiguous RESULT_DTYPE array of length %d" % n)
            if x.dtype != np.float64 or x.shape != (n, self.n_vars) or not x.flags.c_contiguous:
                raise ValueError("out[1] must be a C-contiguous float64 array of shape (%d, %d)" % (n, self.n_vars))
        else:
            res = np.zeros(n, dtype=RESULT_DTYPE)
            x = np.zeros((n, self.n_vars))
        rows = np.zeros((n, self.csv_rows, CSV_COLS)) if csv else None
        self._ck(self._L.qtos_solve_batch(self._h, pp, n, C.byref(o), res.ctypes.data_as(C.c_void_p), _dp(x), _dp(rows)))
        return res, x, rows

    def solve_async(self, problems, out, options=None):
        """Asynchronous solve into caller-supplied host buffers out = (results, x) (see solve); returns at once.
        `wait()` returns the same (results, x).  One call in flight per Solver."""
        p, pp = self._probs(problems)
        n = len(p)
        res, x = out
        if res.dtype != RESULT_DTYPE or res.shape != (n,) or x.dtype != np.float64 or x.shape != (n, self.n_vars) \
                or not res.flags.c_contiguous or not x.flags.c_contiguous:
            raise ValueError("out must be (RESULT_DTYPE[n], float64[n, n_vars]), C-contiguous")
        o = options if options is not None else default_options()
        self._inflight = (p, res, x, o)            # keep the buffers alive until wait()
        self._ck(self._L.qtos_solve_batch_async(self._h, pp, n, C.byref(o), res.ctypes.data_as(C.c_void_p), _dp(x), None))

    def solve_device_async(self, d_problems_ptr, n, options=None, d_results_ptr=None, d_x_ptr=None):
        o = options if options is not None else default_options()
        self._inflight = (None, None, None, o)
        self._ck(self._L.qtos_solve_batch_device_async(self._h, C.c_void_p(d_problems_ptr), int(n), C.byref(o),
                                                       C.c_void_p(d_results_ptr) if d_results_ptr else None,
                                                       C.c_void_p(d_x_ptr) if d_x_ptr else None))

    def wait(self):
        """Join the asynchronous call of this Solver; raises if the solve failed."""
        self._ck(self._L.qtos_wait(self._h))
        _, res, x, _ = getattr(self, "_inflight", None) or (None, None, None, None)
        self._inflight = None
        return res, x

    # ---- continuous batching: the max_batch workspace slots as a pool that jobs flow through
    def stream_begin(self, options=None):
        o = options if options is not None else default_options()
        self._stream_keep = {}
        self._ck(self._L.qtos_stream_begin(self._h, C.byref(o)))

    def stream_submit(self, problems, out=None, csv_out=None):
        """queue a job of <= max_batch windows (host buffers); returns a ticket for stream_wait.  `out` = (results, x) as
        in solve(); fresh arrays otherwise.  `csv_out` = float64[n, csv_rows, 37] (C-contiguous, ideally page-locked): the job also
        delivers the 1 kHz rows of every plan -- what the reference writes to traj.csv -- copied while the pool keeps iterating."""
        p, pp = self._probs(problems)
        n = len(p)
        res, x = out if out is not None else (np.zeros(n, dtype=RESULT_DTYPE), np.zeros((n, self.n_vars)))
        if res.dtype != RESULT_DTYPE or res.shape != (n,) or x.dtype != np.float64 or x.shape != (n, self.n_vars) \
                or not res.flags.c_contiguous or not x.flags.c_contiguous:
            raise ValueError("out must be (RESULT_DTYPE[n], float64[n, n_vars]), C-contiguous")
        t = C.c_int(-1)
        if csv_out is not None:
            if csv_out.dtype != np.float64 or csv_out.shape != (n, self.csv_rows, CSV_COLS) or not csv_out.flags.c_contiguous:
                raise ValueError("csv_out must be float64[n, csv_rows, %d], C-contiguous" % CSV_COLS)
            self._ck(self._L.qtos_stream_submit_csv(self._h, pp, n, res.ctypes.data_as(C.c_void_p), _dp(x), _dp(csv_out), C.byref(t)))
        else:
            self._ck(self._L.qtos_stream_submit(self._h, pp, n, res.ctypes.data_as(C.c_void_p), _dp(x), C.byref(t)))
        self._stream_keep[t.value] = (p, res, x, csv_out)
        return t.value

    def stream_submit_device(self, d_problems_ptr, n, d_results_ptr, d_x_ptr):
        t = C.c_int(-1)
        self._ck(self._L.qtos_stream_submit_device(self._h, C.c_void_p(d_problems_ptr), int(n), C.c_void_p(d_results_ptr),
                                                   C.c_void_p(d_x_ptr), C.byref(t)))
        self._stream_keep[t.value] = (None, None, None, None)
        return t.value

    def stream_wait(self, ticket):
        """block until every window of the job has ended; returns its (results, x) (None, None for device jobs)"""
        self._ck(self._L.qtos_stream_wait(self._h, int(ticket)))
        _, res, x, _rows = self._stream_keep.pop(ticket, (None, None, None, None))
        return res, x

    def stream_info(self):
        s = StreamInfo()
        self._ck(self._L.qtos_stream_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in StreamInfo._fields_}

    def stream_end(self):
        self._ck(self._L.qtos_stream_end(self._h))
        self._stream_keep = {}

    def make_records(self, d_results_ptr, d_group_ptr, id0, id_stride, n, d_rec_ptr):
        """selection records of n device-resident results (see qtos_make_records); d_group: int32, d_rec: float64 [n, 5]"""
        self._ck(self._L.qtos_make_records(self._h, C.c_void_p(d_results_ptr), C.c_void_p(d_group_ptr), int(id0), int(id_stride), int(n), C.c_void_p(d_rec_ptr)))

    def select_best(self, d_rec_ptr, n_rec, n_groups, d_winner_ptr):
        """winning global id per group (int64 [n_groups] on the device) of n_rec gathered records"""
        self._ck(self._L.qtos_select_best(self._h, C.c_void_p(d_rec_ptr), int(n_rec), int(n_groups), C.c_void_p(d_winner_ptr)))

    def solve_device(self, d_problems_ptr, n, options=None, d_results_ptr=None, d_x_ptr=None):
        """Device-resident variant: raw device pointers (e.g. torch tensor .data_ptr())."""
        o = options if options is not None else default_options()
        self._ck(self._L.qtos_solve_batch_device(self._h, C.c_void_p(d_problems_ptr), int(n), C.byref(o),
                                                 C.c_void_p(d_results_ptr) if d_results_ptr else None,
                                                 C.c_void_p(d_x_ptr) if d_x_ptr else None))

    def trace(self, n):
        """Ipopt-style iteration table of the first n problems of the last solve (algorithm IPOPT):
        [n, TRACE_ITERS, (inf_pr, inf_du, mu, ||d||, alpha_du, alpha_pr, ls, tag)]."""
        t = np.zeros((n, TRACE_ITERS, TRACE_COLS))
        self._ck(self._L.qtos_get_trace(self._h, int(n), _dp(t)))
        return t

    def sample_csv(self, problems, x):
        p, pp = self._probs(problems)
        n = len(p)
        xx = np.ascontiguousarray(x, dtype=np.float64).reshape(n, self.n_vars)
        rows = np.zeros((n, self.csv_rows, CSV_COLS))
        self._ck(self._L.qtos_sample_csv(self._h, pp, n, _dp(xx), _dp(rows)))
        return rows

    def sample_rows(self, problems, x, row0, n_rows=1):
        """rows [row0, row0 + n_rows) of the 1 kHz trajectories: [n, n_rows, 37] (negative row0 counts from the end)."""
        p, pp = self._probs(problems)
        n = len(p)
        if row0 < 0:
            row0 += self.csv_rows
        xx = np.ascontiguousarray(x, dtype=np.float64).reshape(n, self.n_vars)
        rows = np.zeros((n, n_rows, CSV_COLS))
        self._ck(self._L.qtos_sample_csv_rows(self._h, pp, n, _dp(xx), int(row0), int(n_rows), _dp(rows)))
        return rows

    def launch_count(self):
        return int(self._L.qtos_launch_count(self._h))

    def set_profiling(self, on=True):
        self._ck(self._L.qtos_set_profiling(self._h, int(bool(on))))

    def last_stats(self):
        """device ms per phase of the last solve (profiling on), problems factored, iterations."""
        st = Stats()
        self._L.qtos_last_stats(self._h, C.byref(st))
        names = ["init", "jac", "prepare", "assemble", "factor", "step", "solve"]
        return {"ms": dict(zip(names, list(st.ms)[:7])), "factorizations": st.factorizations,
                "factor_launches": st.factor_launches, "iterations": st.iterations, "retried": st.retried}

    def measure_heightfield_staging(self, hf_id, xy, group_size):
        """xy [n_groups * group_size, 2]: ms per launch answering from global memory / from a bulk-copied shared-memory tile,
        the largest difference of the answers, groups that did not fit the tile"""
        xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        a, b, d, f = C.c_double(), C.c_double(), C.c_double(), C.c_int()
        self._ck(self._L.qtos_measure_heightfield_staging(self._h, int(hf_id), _dp(xy), len(xy) // group_size, int(group_size),
                                                          C.byref(a), C.byref(b), C.byref(d), C.byref(f)))
        return {"ms_direct": a.value, "ms_staged": b.value, "max_diff": d.value, "fallbacks": f.value}

    def fp64_peak_tflops(self):
        v = C.c_double()
        self._ck(self._L.qtos_measure_fp64_peak(self._h, C.byref(v)))
        return v.value

    @property
    def stream(self):
        return self._L.qtos_stream(self._h)


def write_csv(rows, path):
    rows = np.ascontiguousarray(rows, dtype=np.float64).reshape(-1, CSV_COLS)
    rc = lib().qtos_write_csv(_dp(rows), len(rows), os.fsencode(path))
    if rc != 0:
        raise QtosError("cannot write %s" % path)
