"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(qtos_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

NEE = 4
MAX_PHASES = 64
COMBOS = {"C0": 0, "C1": 1, "C2": 2, "C3": 3, "C4": 4, "Custom": 5}


class Heightfield(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("res", C.c_double),
                ("h", C.POINTER(C.c_double))]


class Shape(C.Structure):
    _fields_ = [("mass", C.c_double), ("I_b", C.c_double * 9),
                ("nominal", (C.c_double * 3) * NEE), ("max_dev", C.c_double * 3),
                ("mu", C.c_double), ("force_limit", C.c_double), ("t_swing_avg", C.c_double),
                ("dt_base_poly", C.c_double), ("force_polys_per_stance", C.c_int),
                ("ee_polys_per_swing", C.c_int), ("dt_dynamic", C.c_double),
                ("dt_rom", C.c_double), ("combo", C.c_int), ("duration", C.c_double),
                ("base_rom", C.c_int), ("dt_base_rom", C.c_double), ("terrain_gradients", C.c_int),
                ("cost_force_z", C.c_double), ("cost_ee_vel_xy", C.c_double),
                ("optimize_timings", C.c_int), ("phase_dur_min", C.c_double), ("phase_dur_max", C.c_double)]


class Instance(C.Structure):
    _fields_ = [("start_pos", C.c_double * 3), ("start_ang", C.c_double * 3),
                ("start_vel", C.c_double * 3), ("start_ang_vel", C.c_double * 3),
                ("goal", C.c_double * 3), ("ee", (C.c_double * 3) * NEE),
                ("t_start", C.c_double)]


class IpmOptions(C.Structure):
    _fields_ = [("tol", C.c_double), ("constr_viol_tol", C.c_double),
                ("compl_inf_tol", C.c_double), ("dual_inf_tol", C.c_double),
                ("max_iter", C.c_int), ("mu_init", C.c_double), ("mu_strategy", C.c_int),
                ("sigma_w", C.c_double), ("verbose", C.c_int), ("delta_c", C.c_double), ("feas_exit", C.c_int)]


class IpmResult(C.Structure):
    _fields_ = [("status", C.c_int), ("iters", C.c_int),
                ("constr_viol", C.c_double), ("dual_inf", C.c_double),
                ("compl_inf", C.c_double), ("nlp_error", C.c_double), ("mu", C.c_double),
                ("n_trace", C.c_int),
                ("tr_inf_pr", C.c_double * 256), ("tr_inf_du", C.c_double * 256),
                ("tr_mu", C.c_double * 256), ("tr_dnorm", C.c_double * 256),
                ("tr_alpha_pr", C.c_double * 256), ("tr_alpha_du", C.c_double * 256),
                ("tr_ls", C.c_int * 256), ("chol_fix", C.c_int)]


class IpoptOptions(C.Structure):
    _fields_ = [("tol", C.c_double), ("constr_viol_tol", C.c_double), ("compl_inf_tol", C.c_double),
                ("dual_inf_tol", C.c_double), ("max_iter", C.c_int), ("delta_c", C.c_double),
                ("n_refine", C.c_int), ("lm_history", C.c_int), ("verbose", C.c_int), ("sigma_floor", C.c_double),
                ("retry_failed", C.c_int), ("retry_sigma_floor", C.c_double)]


class IpoptResult(C.Structure):
    _fields_ = [("status", C.c_int), ("iters", C.c_int),
                ("constr_viol", C.c_double), ("dual_inf", C.c_double), ("compl_inf", C.c_double),
                ("nlp_error", C.c_double), ("mu", C.c_double), ("n_trace", C.c_int),
                ("tr_inf_pr", C.c_double * 256), ("tr_inf_du", C.c_double * 256), ("tr_mu", C.c_double * 256),
                ("tr_dnorm", C.c_double * 256), ("tr_alpha_pr", C.c_double * 256), ("tr_alpha_du", C.c_double * 256),
                ("tr_ls", C.c_int * 256), ("tr_pairs", C.c_int * 256), ("tr_free", C.c_int * 256),
                ("tr_tag", C.c_char * 256), ("chol_fix", C.c_int), ("n_regularized", C.c_int), ("retried", C.c_int),
                ("objective", C.c_double)]


def build(force=False):
    """Compile oracle/liboracle.so with the committed Makefile (gcc)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        L.orc_problem_create.restype = C.c_void_p
        L.orc_problem_create.argtypes = [C.POINTER(Shape), C.POINTER(Instance), C.POINTER(Heightfield)]
        L.orc_problem_free.argtypes = [C.c_void_p]
        L.orc_n.argtypes = [C.c_void_p]
        L.orc_m.argtypes = [C.c_void_p]
        L.orc_get_x0.argtypes = [C.c_void_p, dp]
        L.orc_get_bounds.argtypes = [C.c_void_p, dp, dp, dp, dp]
        L.orc_get_phase_durations.argtypes = [C.c_void_p, C.c_int, dp]
        L.orc_get_layout.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_get_schedule_layout.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_height.restype = C.c_double
        L.orc_height.argtypes = [C.POINTER(Heightfield), C.c_double, C.c_double]
        L.orc_height_cell.argtypes = [C.POINTER(Heightfield), C.c_double, C.c_double, C.POINTER(C.c_longlong)]
        L.orc_height_deriv.argtypes = [C.POINTER(Heightfield), C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_eval_g.argtypes = [C.c_void_p, dp, dp]
        L.orc_eval_jac.argtypes = [C.c_void_p, dp, dp, C.POINTER(C.c_ubyte)]
        L.orc_eval_cost.restype = C.c_double
        L.orc_eval_cost.argtypes = [C.c_void_p, dp, dp]
        L.orc_csv_rows.argtypes = [C.c_void_p, C.c_double]
        L.orc_sample_csv.argtypes = [C.c_void_p, dp, C.c_double, dp]
        L.orc_write_csv.argtypes = [C.c_void_p, dp, C.c_double, C.c_char_p]
        L.orc_default_shape.argtypes = [C.POINTER(Shape)]
        L.orc_ipm_default_options.argtypes = [C.POINTER(IpmOptions)]
        L.orc_ipm_solve.argtypes = [C.c_void_p, C.POINTER(IpmOptions), dp, C.POINTER(IpmResult)]
        L.orc_ipopt_default_options.argtypes = [C.POINTER(IpoptOptions)]
        L.orc_ipopt_solve.argtypes = [C.c_void_p, C.POINTER(IpoptOptions), dp, C.POINTER(IpoptResult)]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def default_shape(combo="Custom", duration=5.0, mass=None):
    s = Shape()
    lib().orc_default_shape(C.byref(s))
    s.combo = COMBOS[combo] if isinstance(combo, str) else int(combo)
    s.duration = float(duration)
    if mass is not None:
        s.mass = float(mass)
    return s


class Terrain:
    """hf[ix][iy] grid with world offset (-1,-1) (ref: custom_terrain.hpp:32-35)."""

    def __init__(self, grid, res):
        self.grid = np.ascontiguousarray(grid, dtype=np.float64)
        assert self.grid.ndim == 2
        self.res = float(res)
        self.c = Heightfield(self.grid.shape[0], self.grid.shape[1], self.res, _dp(self.grid))

    def height(self, x, y):
        return lib().orc_height(C.byref(self.c), float(x), float(y))

    def height_deriv(self, x, y):
        hx, hy = C.c_double(), C.c_double()
        lib().orc_height_deriv(C.byref(self.c), float(x), float(y), C.byref(hx), C.byref(hy))
        return hx.value, hy.value

    def cell(self, x, y):
        idx = (C.c_longlong * 4)()
        lib().orc_height_cell(C.byref(self.c), float(x), float(y), idx)
        return tuple(idx)


def make_instance(start_pos=(0, 0, 0.24), start_ang=(0, 0, 0), goal=(0.5, 0, 0.24),
                  ee=None, t_start=0.0, start_vel=(0, 0, 0), start_ang_vel=(0, 0, 0)):
    if ee is None:
        ee = [(0.21, 0.18, 0.0), (0.21, -0.18, 0.0), (-0.21, 0.18, 0.0), (-0.21, -0.18, 0.0)]
    inst = Instance()
    for i in range(3):
        inst.start_pos[i] = start_pos[i]
        inst.start_ang[i] = start_ang[i]
        inst.start_vel[i] = start_vel[i]
        inst.start_ang_vel[i] = start_ang_vel[i]
        inst.goal[i] = goal[i]
        for e in range(NEE):
            inst.ee[e][i] = ee[e][i]
    inst.t_start = t_start
    return inst


class Problem:
    def __init__(self, shape, inst, terrain):
        self.shape, self.inst, self.terrain = shape, inst, terrain
        self.h = lib().orc_problem_create(C.byref(shape), C.byref(inst), C.byref(terrain.c))
        if not self.h:
            raise ValueError("orc_problem_create failed")
        self.n = lib().orc_n(self.h)
        self.m = lib().orc_m(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_problem_free(self.h)
            self.h = None

    def x0(self):
        x = np.zeros(self.n)
        lib().orc_get_x0(self.h, _dp(x))
        return x

    def bounds(self):
        xl, xu = np.zeros(self.n), np.zeros(self.n)
        gl, gu = np.zeros(self.m), np.zeros(self.m)
        lib().orc_get_bounds(self.h, _dp(xl), _dp(xu), _dp(gl), _dp(gu))
        return xl, xu, gl, gu

    def phase_durations(self, ee):
        out = np.zeros(MAX_PHASES)
        n = lib().orc_get_phase_durations(self.h, ee, _dp(out))
        return out[:n].copy()

    def layout(self):
        vo = (C.c_int * 11)()
        ro = (C.c_int * 20)()
        lib().orc_get_layout(self.h, vo, ro)
        return list(vo), list(ro)

    def schedule_layout(self):
        """(variable offset of every foot's phase durations, TotalDuration row of every foot); -1 when timings are fixed"""
        so, rt = (C.c_int * 4)(), (C.c_int * 4)()
        lib().orc_get_schedule_layout(self.h, so, rt)
        return list(so), list(rt)

    def g(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        g = np.zeros(self.m)
        lib().orc_eval_g(self.h, _dp(x), _dp(g))
        return g

    def jac(self, x, with_mask=False):
        x = np.ascontiguousarray(x, dtype=np.float64)
        J = np.zeros((self.m, self.n))
        if with_mask:
            mask = np.zeros((self.m, self.n), dtype=np.uint8)
            lib().orc_eval_jac(self.h, _dp(x), _dp(J), mask.ctypes.data_as(C.POINTER(C.c_ubyte)))
            return J, mask
        lib().orc_eval_jac(self.h, _dp(x), _dp(J), None)
        return J

    def cost(self, x, with_grad=False):
        """objective (the NodeCost terms of the shape; 0 when none) and optionally its gradient"""
        x = np.ascontiguousarray(x, dtype=np.float64)
        if with_grad:
            g = np.zeros(self.n)
            return lib().orc_eval_cost(self.h, _dp(x), _dp(g)), g
        return lib().orc_eval_cost(self.h, _dp(x), None)

    def csv(self, x, dt=0.001):
        x = np.ascontiguousarray(x, dtype=np.float64)
        n = lib().orc_csv_rows(self.h, dt)
        rows = np.zeros((n, 37))
        lib().orc_sample_csv(self.h, _dp(x), dt, _dp(rows))
        return rows

    def write_csv(self, x, path, dt=0.001):
        x = np.ascontiguousarray(x, dtype=np.float64)
        return lib().orc_write_csv(self.h, _dp(x), dt, path.encode())

    def solve_ipopt(self, x0=None, **opts):
        """the reference's Ipopt 3.11.9 algorithm in the form the CUDA kernels implement (towr_ipopt.c)"""
        o = IpoptOptions()
        lib().orc_ipopt_default_options(C.byref(o))
        for k, v in opts.items():
            setattr(o, k, v)
        x = self.x0() if x0 is None else np.array(x0, dtype=np.float64)
        res = IpoptResult()
        lib().orc_ipopt_solve(self.h, C.byref(o), _dp(x), C.byref(res))
        return x, res

    def solve(self, x0=None, **opts):
        o = IpmOptions()
        lib().orc_ipm_default_options(C.byref(o))
        for k, v in opts.items():
            setattr(o, k, v)
        x = self.x0() if x0 is None else np.array(x0, dtype=np.float64)
        res = IpmResult()
        lib().orc_ipm_solve(self.h, C.byref(o), _dp(x), C.byref(res))
        return x, res
