"""Algorithm prototype (numpy, dense) for the batched IPM -- development tool only.

Uses the CPU oracle's g/J to settle the interior-point design that oracle/towr_ipm.c and
the CUDA kernels then both implement.  Not part of the product path.
"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O

INF = 1e19


def solve(p, opts=None, verbose=True):
    o = dict(tol=1e-3, constr_viol_tol=1e-4, compl_inf_tol=1e-4, dual_inf_tol=1.0, max_iter=200,
             mu_init=0.1, sigma=1.0, delta_c=1e-8, kappa_eps=10.0, kappa_mu=0.2, theta_mu=1.5,
             tau_min=0.99, bound_push=0.01, bound_frac=0.01, relax=1e-8, max_ls=12, eta=1e-4)
    if opts:
        o.update(opts)
    x = p.x0()
    xl, xu, gl, gu = p.bounds()
    free = np.where(xl != xu)[0]
    x[xl == xu] = xl[xl == xu]
    eq = np.where(gl == gu)[0]
    iq = np.where(gl != gu)[0]
    J0 = p.jac(x)[:, free]
    sc = np.minimum(1.0, 100.0 / np.maximum(np.abs(J0).max(axis=1), 1e-300))
    sc = np.maximum(sc, 1e-8)
    dL = gl[iq] * sc[iq]
    dU = gu[iq] * sc[iq]
    hasL = gl[iq] > -INF
    hasU = gu[iq] < INF
    dL = np.where(hasL, dL - o['relax'] * np.maximum(1, np.abs(dL)), -np.inf)
    dU = np.where(hasU, dU + o['relax'] * np.maximum(1, np.abs(dU)), np.inf)
    ceq = gl[eq] * sc[eq]

    def evalg(x):
        g = p.g(x) * sc
        return g[eq] - ceq, g[iq]

    def evalJ(x):
        J = p.jac(x)[:, free] * sc[:, None]
        return J[eq], J[iq]

    c, d = evalg(x)
    # slack init
    s = d.copy()
    both = hasL & hasU
    pL = np.where(hasL, np.minimum(o['bound_push'] * np.maximum(1, np.abs(np.where(hasL, dL, 0))),
                                   np.where(both, o['bound_frac'] * (dU - dL), np.inf)), 0)
    pU = np.where(hasU, np.minimum(o['bound_push'] * np.maximum(1, np.abs(np.where(hasU, dU, 0))),
                                   np.where(both, o['bound_frac'] * (dU - dL), np.inf)), 0)
    s = np.where(hasL, np.maximum(s, dL + pL), s)
    s = np.where(hasU, np.minimum(s, dU - pU), s)
    zL = np.where(hasL, 1.0, 0.0)
    zU = np.where(hasU, 1.0, 0.0)
    yc = np.zeros(len(eq))
    yd = np.zeros(len(iq))
    mu = o['mu_init']
    nu = 1.0
    n = len(free)
    trace = []
    status = -1

    def barrier(s, mu):
        return -mu * (np.log(np.where(hasL, s - dL, 1.0)).sum() + np.log(np.where(hasU, dU - s, 1.0)).sum())

    for it in range(o['max_iter'] + 1):
        Jc, Jd = evalJ(x)
        sl = np.where(hasL, s - dL, 1.0)
        su = np.where(hasU, dU - s, 1.0)
        rx = Jc.T @ yc + Jd.T @ yd
        rs = -yd - zL + zU
        rd = d - s
        theta_inf = max(np.abs(c).max(), np.abs(rd).max())
        dual_inf = max(np.abs(rx).max(), np.abs(rs).max())
        compl0 = max((zL * sl * hasL).max(), (zU * su * hasU).max())

        def err(mu_):
            cm = max(np.abs((zL * sl - mu_) * hasL).max(), np.abs((zU * su - mu_) * hasU).max())
            sd = max(100.0, (np.abs(yc).sum() + np.abs(yd).sum() + zL.sum() + zU.sum()) / (len(yc) + len(yd) + hasL.sum() + hasU.sum())) / 100.0
            sc_ = max(100.0, (zL.sum() + zU.sum()) / (hasL.sum() + hasU.sum())) / 100.0
            return max(dual_inf / sd, theta_inf, cm / sc_)
        E0 = err(0.0)
        # unscaled violation
        g_un = p.g(x)
        viol = max(np.maximum(gl - g_un, 0).max(), np.maximum(g_un - gu, 0).max())
        trace.append((it, theta_inf, dual_inf, mu, viol, E0))
        if verbose:
            print(f"{it:3d} inf_pr={theta_inf:9.3e} inf_du={dual_inf:9.3e} lg(mu)={np.log10(mu):5.1f} viol={viol:9.3e} E0={E0:9.3e}", end='')
        if E0 <= o['tol'] and viol <= o['constr_viol_tol'] and compl0 <= o['compl_inf_tol'] and dual_inf <= o['dual_inf_tol']:
            status = 0
            if verbose: print()
            break
        if it == o['max_iter']:
            if verbose: print()
            break
        # barrier update (monotone)
        mu_min = min(o['tol'], o['compl_inf_tol']) / (o['kappa_eps'] + 1)
        while err(mu) <= o['kappa_eps'] * mu and mu > mu_min:
            mu = max(mu_min, min(o['kappa_mu'] * mu, mu ** o['theta_mu']))
        tau = max(o['tau_min'], 1 - mu)
        Sig = zL / sl * hasL + zU / su * hasU
        rsm = -yd - mu / sl * hasL + mu / su * hasU
        rho = 1.0 / o['delta_c']
        M = o['sigma'] * np.eye(n) + Jd.T @ (Sig[:, None] * Jd) + rho * (Jc.T @ Jc)
        rhs = -rx - Jd.T @ (Sig * rd + rsm) - rho * (Jc.T @ c)
        L = np.linalg.cholesky(M)
        dx = np.linalg.solve(L.T, np.linalg.solve(L, rhs))
        ds = Jd @ dx + rd
        dyd = Sig * ds + rsm
        dyc = rho * (Jc @ dx + c)
        dzL = (mu / sl - zL - zL / sl * ds) * hasL
        dzU = (mu / su - zU + zU / su * ds) * hasU
        # fraction to boundary
        def ftb(v, dv, mask):
            neg = mask & (dv < 0)
            return min(1.0, (-tau * v[neg] / dv[neg]).min()) if neg.any() else 1.0
        a_pr = min(ftb(sl, ds, hasL), ftb(su, -ds, hasU))
        a_du = min(ftb(zL, dzL, hasL), ftb(zU, dzU, hasU))
        # merit line search
        theta1 = np.abs(c).sum() + np.abs(rd).sum()
        gphi_d = (-mu / sl * hasL + mu / su * hasU) @ ds
        quad = o['sigma'] * dx @ dx + ds @ (Sig * ds)
        if theta1 > 1e-14:
            nu_trial = (gphi_d + 0.5 * quad) / (0.7 * theta1)
            if nu < nu_trial:
                nu = nu_trial + 1.0
        phi0 = barrier(s, mu) + nu * theta1
        Dphi = gphi_d - nu * theta1
        alpha = a_pr
        ls = 0
        ok = False
        while ls < o['max_ls']:
            xt = x.copy(); xt[free] += alpha * dx
            st = s + alpha * ds
            ct, dt = evalg(xt)
            th = np.abs(ct).sum() + np.abs(dt - st).sum()
            phit = barrier(st, mu) + nu * th
            ls += 1
            if phit <= phi0 + o['eta'] * alpha * Dphi or phit <= phi0 - 1e-13*abs(phi0) and False:
                ok = True
                break
            alpha *= 0.5
        if verbose:
            print(f" |dx|={np.abs(dx).max():8.2e} a_pr={alpha:8.2e} a_du={a_du:8.2e} ls={ls} nu={nu:8.2e} {'ok' if ok else 'FAIL'}")
        x = xt; s = st; c = ct; d = dt
        yc = yc + alpha * dyc
        yd = yd + alpha * dyd
        zL = zL + a_du * dzL
        zU = zU + a_du * dzU
        # multiplier safeguard
        sl = np.where(hasL, s - dL, 1.0); su = np.where(hasU, dU - s, 1.0)
        ks = 1e10
        zL = np.where(hasL, np.clip(zL, mu / (ks * sl), ks * mu / sl), 0)
        zU = np.where(hasU, np.clip(zU, mu / (ks * su), ks * mu / su), 0)
    return x, status, it, trace


def flat_problem(combo="Custom", T=5.0, mass=3.0, goal=(0.5, 0, 0.24)):
    sh = O.default_shape(combo, T, mass=mass)
    ter = O.Terrain(np.zeros((600, 200)), 0.01)
    inst = O.make_instance(goal=goal, ee=[(0.21, 0.19, 0), (0.21, -0.19, 0), (-0.21, 0.19, 0), (-0.21, -0.19, 0)])
    return O.Problem(sh, inst, ter)


if __name__ == "__main__":
    combo = sys.argv[1] if len(sys.argv) > 1 else "Custom"
    T = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0
    p = flat_problem(combo, T)
    x, st, it, tr = solve(p)
    print("status", st, "iters", it)


def rough_terrain(seed=1234):
    rng = np.random.default_rng(seed)
    plate = np.round(rng.uniform(0, 0.075, (32, 32)), 4)
    return np.kron(plate, np.ones((8, 8))), 0.02


def rough_problem(idx, combo="C1", T=2.0, seed=1234, mass=1.5):
    grid, res = rough_terrain(seed)
    ter = O.Terrain(grid, res)
    rng = np.random.default_rng(seed + 1 + idx)
    sx, sy = rng.uniform(0, 2.5, 2)
    gx, gy = sx + rng.uniform(0.2, 0.6), sy + rng.uniform(-0.1, 0.1)
    ee = [(sx + a, sy + b, ter.height(sx + a, sy + b)) for a, b in [(0.21, 0.19), (0.21, -0.19), (-0.21, 0.19), (-0.21, -0.19)]]
    inst = O.make_instance(start_pos=(sx, sy, ter.height(sx, sy) + 0.24), goal=(gx, gy, 0.24), ee=ee)
    return O.Problem(O.default_shape(combo, T, mass=mass), inst, ter)
