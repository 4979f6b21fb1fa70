"""ipopt_emul.py -- oracle: restatement of the Ipopt 3.11.9 algorithm the reference runs (numpy, dense).

TEST INFRASTRUCTURE ONLY (see towr_oracle.h).  The reference solves its NLP with ifopt -> Ipopt 3.11.9 +
MUMPS (ref: solver/towr/src/main.cpp:444-463; logs/towr_log.out:37,88).  Ipopt's source is NOT vendored
under /root/reference, so this file restates the published algorithm (Waechter & Biegler, Math. Prog. 106,
2006; Nocedal, Waechter & Waltz, "Adaptive barrier update strategies", SIAM J. Optim. 19, 2009) with the
option values in force on the reference path:

    ifopt sets      tol 1e-3, hessian_approximation limited-memory, linear_solver mumps
    main.cpp sets   max_iter 200, max_cpu_time -r, jacobian_approximation exact
    the logs show   mu_strategy adaptive (iteration-0 lg(mu) = 0.0, non-monotone mu; towr_log.out:55-62)
    all others      Ipopt 3.11 defaults (quality-function mu oracle, obj-constr-filter globalization,
                    L-BFGS history 6 / scalar1, filter line search, bound_push/frac 0.01,
                    gradient-based scaling with max gradient 100, bound_relax_factor 1e-8)

The anchor is the reference's own golden data: the three iteration tables of logs/towr_log.out (columns
inf_pr, inf_du, lg(mu), ||d||, alpha_du, alpha_pr, step type) and the plans data/traj/towr.csv those
solves wrote.  tests/test_ipopt_emulation.py holds the comparison.

g(x) and J(x) come from the C oracle (towr_eval.c); everything here is dense linear algebra on the
1005 + 706 (+1024 condensed) system, so one S5 solve takes a few seconds.
"""
import numpy as np

INF = 1e19


class Options:
    tol = 1e-3
    constr_viol_tol = 1e-4
    compl_inf_tol = 1e-4
    dual_inf_tol = 1.0
    max_iter = 200
    s_max = 100.0
    bound_push = 0.01
    bound_frac = 0.01
    bound_relax_factor = 1e-8
    nlp_scaling_max_gradient = 100.0
    nlp_scaling_min_value = 1e-8
    constr_mult_init_max = 1000.0
    lsq_init = False            # least-squares y0: singular here (duplicated final sample) -> Ipopt falls back to y = 0
    tau_min = 0.99
    kappa_sigma = 1e10
    # adaptive mu
    mu_max_fact = 1e3
    mu_min = 1e-11
    sigma_max = 100.0
    sigma_min = 1e-6
    qf_section_sigma_tol = 1e-2
    qf_section_qf_tol = 0.0
    qf_max_section_steps = 8
    filter_margin_fact = 1e-5
    filter_max_margin = 1.0
    adaptive_mu_monotone_init_factor = 0.8
    barrier_tol_factor = 10.0
    mu_linear_decrease_factor = 0.2
    mu_superlinear_decrease_power = 1.5
    # L-BFGS
    lm_history = 6
    lm_init_val = 1.0
    lm_init_val_max = 1e8
    lm_init_val_min = 1e-8
    lm_max_skipping = 2
    # filter line search
    theta_max_fact = 1e4
    theta_min_fact = 1e-4
    eta_phi = 1e-8
    delta = 1.0
    s_phi = 2.3
    s_theta = 1.1
    gamma_phi = 1e-8
    gamma_theta = 1e-5
    alpha_min_frac = 0.05
    max_soc = 4
    kappa_soc = 0.99
    alpha_red_factor = 0.5
    obj_max_inc = 5.0
    # perturbation
    delta_cd_val = 1e-8
    delta_cd_exp = 0.25
    first_hessian_perturbation = 1e-4
    verbose = False


class Result:
    pass


def _frac_to_bound(slack, dslack, tau):
    """largest alpha in (0,1] with slack + alpha*dslack >= (1-tau)*slack  (DenseVector::FracToBound)."""
    neg = dslack < 0
    if not neg.any():
        return 1.0
    return float(min(1.0, (-tau * slack[neg] / dslack[neg]).min()))


class IpoptEmulator:
    def __init__(self, problem, options=None):
        self.p = problem
        self.o = options or Options()
        o = self.o
        xl, xu, gl, gu = problem.bounds()
        self.x_full = problem.x0()
        fixed = xl == xu
        self.x_full[fixed] = xl[fixed]                    # fixed_variable_treatment make_parameter
        self.free = np.where(~fixed)[0]
        self.eq = np.where(gl == gu)[0]
        self.iq = np.where(gl != gu)[0]
        self.n, self.mc, self.md = len(self.free), len(self.eq), len(self.iq)
        # gradient-based scaling at the starting point
        J0 = problem.jac(self.x_full)[:, self.free]
        amax = np.abs(J0).max(axis=1)
        sc = np.ones(problem.m)
        big = amax > o.nlp_scaling_max_gradient
        sc[big] = np.maximum(o.nlp_scaling_max_gradient / amax[big], o.nlp_scaling_min_value)
        self.sc = sc
        self.gl_eq = gl[self.eq]
        dl, du = gl[self.iq].copy(), gu[self.iq].copy()
        self.hasL, self.hasU = dl > -INF, du < INF
        dl = np.where(self.hasL, dl - o.bound_relax_factor * np.maximum(1.0, np.abs(dl)), -np.inf)
        du = np.where(self.hasU, du + o.bound_relax_factor * np.maximum(1.0, np.abs(du)), np.inf)
        self.dL = np.where(self.hasL, dl * sc[self.iq], -np.inf)
        self.dU = np.where(self.hasU, du * sc[self.iq], np.inf)
        self.gl_raw, self.gu_raw = gl, gu
        # bounds on the variables themselves (none on the reference's path, where every bounded variable is fixed; the gait-timing
        # option bounds the phase durations): multipliers z_L, z_U, relaxed like the slack bounds
        fl, fu = xl[self.free], xu[self.free]
        self.xhasL, self.xhasU = fl > -INF, fu < INF
        self.xL = np.where(self.xhasL, fl - o.bound_relax_factor * np.maximum(1.0, np.abs(fl)), -np.inf)
        self.xU = np.where(self.xhasU, fu + o.bound_relax_factor * np.maximum(1.0, np.abs(fu)), np.inf)
        self.n_bounds = int(self.hasL.sum() + self.hasU.sum() + self.xhasL.sum() + self.xhasU.sum())
        # optional cost terms (f == 0 on the reference's path): gradient-based objective scaling at the starting point
        self.has_cost = hasattr(problem, "cost") and (problem.shape.cost_force_z != 0.0 or problem.shape.cost_ee_vel_xy != 0.0)
        self.df = 1.0
        if self.has_cost:
            g0 = np.abs(problem.cost(self.x_full, with_grad=True)[1][self.free]).max()
            if g0 > o.nlp_scaling_max_gradient:
                self.df = max(o.nlp_scaling_max_gradient / g0, o.nlp_scaling_min_value)

    # ---- problem functions in Ipopt's scaled space
    def full(self, x):
        xf = self.x_full.copy()
        xf[self.free] = x
        return xf

    def cd(self, x):
        g = self.p.g(self.full(x))
        return self.sc[self.eq] * (g[self.eq] - self.gl_eq), self.sc[self.iq] * g[self.iq], g

    def f(self, x, with_grad=False):
        if not self.has_cost:
            return (0.0, np.zeros(self.n)) if with_grad else 0.0
        if with_grad:
            v, g = self.p.cost(self.full(x), with_grad=True)
            return self.df * v, self.df * g[self.free]
        return self.df * self.p.cost(self.full(x))

    def jac(self, x):
        J = self.p.jac(self.full(x))[:, self.free] * self.sc[:, None]
        return J[self.eq], J[self.iq]

    def unscaled_violation(self, g):
        return float(max(0.0, np.maximum(self.gl_raw - g, g - self.gu_raw).max()))

    # ---- linear algebra: condensed augmented system, dense
    def factor(self, W, Sig, Jc, Jd, delta_c):
        n, mc = self.n, self.mc
        K = np.zeros((n + mc, n + mc))
        K[:n, :n] = W + Jd.T @ (Sig[:, None] * Jd)
        K[:n, n:] = Jc.T
        K[n:, :n] = Jc
        K[n:, n:] = -delta_c * np.eye(mc)
        import scipy.linalg as sla
        lu = sla.lu_factor(K)
        return (lu, K)

    def kkt_solve(self, fac, Sig, Jd, rx, rs, rc, rd, rvL, rvU, sL, sU, vL, vU, xb=None):
        """solves the 8-block primal-dual system  K * sol = rhs  (PDFullSpaceSolver::SolveOnce).
        xb = (rzL, rzU, xsL, xsU, zL, zU): the blocks of the variable bounds (their Sigma is part of the factored W)."""
        import scipy.linalg as sla
        lu, K = fac
        n = self.n
        aug_s = rs.copy()
        aug_s[self.hasL] += rvL / sL
        aug_s[self.hasU] -= rvU / sU
        rx = rx.copy()
        if xb is not None:
            rzL, rzU, xsL, xsU, zL, zU = xb
            rx[self.xhasL] += rzL / xsL
            rx[self.xhasU] -= rzU / xsU
        b = np.concatenate([rx + Jd.T @ (Sig * rd + aug_s), rc])
        sol = sla.lu_solve(lu, b)
        r = b - K @ sol                                   # one refinement step on the condensed system
        sol += sla.lu_solve(lu, r)
        dx, dyc = sol[:n], sol[n:]
        ds = Jd @ dx - rd
        dyd = Sig * ds - aug_s
        dvL = (rvL - vL * ds[self.hasL]) / sL
        dvU = (rvU + vU * ds[self.hasU]) / sU
        if xb is not None:
            return dx, ds, dyc, dyd, dvL, dvU, (rzL - zL * dx[self.xhasL]) / xsL, (rzU + zU * dx[self.xhasU]) / xsU
        return dx, ds, dyc, dyd, dvL, dvU, np.zeros(0), np.zeros(0)

    # ---- the algorithm
    def solve(self):
        o = self.o
        n, mc, md = self.n, self.mc, self.md
        hasL, hasU, dL, dU = self.hasL, self.hasU, self.dL, self.dU
        x = self.x_full[self.free].copy()
        xhasL, xhasU, xL, xU = self.xhasL, self.xhasU, self.xL, self.xU
        if xhasL.any() or xhasU.any():                     # DefaultIterateInitializer::push_variables on x
            xw = np.where(xhasL & xhasU, xU - xL, np.inf)
            qL = np.minimum(o.bound_push * np.maximum(1.0, np.abs(np.where(xhasL, xL, 0.0))), o.bound_frac * xw)
            qU = np.minimum(o.bound_push * np.maximum(1.0, np.abs(np.where(xhasU, xU, 0.0))), o.bound_frac * xw)
            x = np.where(xhasL, np.maximum(x, xL + qL), x)
            x = np.where(xhasU, np.minimum(x, xU - qU), x)
        zL, zU = np.ones(int(xhasL.sum())), np.ones(int(xhasU.sum()))
        c, d, graw = self.cd(x)
        Jc, Jd = self.jac(x)
        fval, gf = self.f(x, with_grad=True)
        # slack initialisation (DefaultIterateInitializer::push_variables)
        both = hasL & hasU
        width = np.where(both, dU - dL, np.inf)
        pL = np.minimum(o.bound_push * np.maximum(1.0, np.abs(np.where(hasL, dL, 0.0))), o.bound_frac * width)
        pU = np.minimum(o.bound_push * np.maximum(1.0, np.abs(np.where(hasU, dU, 0.0))), o.bound_frac * width)
        s = d.copy()
        s = np.where(hasL, np.maximum(s, dL + pL), s)
        s = np.where(hasU, np.minimum(s, dU - pU), s)
        vL = np.ones(int(hasL.sum()))
        vU = np.ones(int(hasU.sum()))
        yc, yd = np.zeros(mc), np.zeros(md)

        res = Result()
        res.trace = []
        # L-BFGS memory
        S, Y = [], []
        sigma_w = o.lm_init_val
        lm_skipped = 0
        last = None
        # adaptive mu state
        mu, tau = 1.0, 0.0
        free_mode = True
        mu_max = -1.0
        mu_min = min(o.mu_min, 0.5 * min(o.tol, o.compl_inf_tol))
        amu_filter = []                                    # (f, theta) entries of AdaptiveMuUpdate's own filter
        ls_filter = []
        theta_max = theta_min = -1.0
        status, it = -1, 0
        alpha_pr = alpha_du = 0.0
        dnorm = 0.0
        step_tag = " "
        ls_count = 0

        def slacks(s):
            return s[hasL] - dL[hasL], dU[hasU] - s[hasU]

        def xslacks(x_):
            return x_[xhasL] - xL[xhasL], xU[xhasU] - x_[xhasU]

        while True:
            sL, sU = slacks(s)
            xsL, xsU = xslacks(x)
            # ---- L-BFGS update (LimMemQuasiNewtonUpdater::UpdateHessian)
            if last is not None:
                lx, lJc, lJd, lgf = last
                s_new = x - lx
                y_new = (gf + Jc.T @ yc + Jd.T @ yd) - (lgf + lJc.T @ yc + lJd.T @ yd)
                sTy = float(s_new @ y_new)
                snrm, ynrm = np.linalg.norm(s_new), np.linalg.norm(y_new)
                skipping = sTy <= np.sqrt(np.finfo(float).eps) * snrm * ynrm
                if skipping:
                    lm_skipped += 1
                    if lm_skipped >= o.lm_max_skipping:
                        S, Y, sigma_w, lm_skipped = [], [], o.lm_init_val, 0
                else:
                    lm_skipped = 0
                    S.append(s_new); Y.append(y_new)
                    if len(S) > o.lm_history:
                        S.pop(0); Y.pop(0)
                    sigma_w = min(max(sTy / float(s_new @ s_new), o.lm_init_val_min), o.lm_init_val_max)
            last = (x.copy(), Jc, Jd, gf)
            W = sigma_w * np.eye(n)
            self._sigma_w, self._Bl, self._Mid = sigma_w, None, None
            if S:
                Sm, Ym = np.array(S).T, np.array(Y).T
                StY = Sm.T @ Ym
                Lm = np.tril(StY, -1)
                Dm = np.diag(np.diag(StY))
                Mid = np.block([[sigma_w * Sm.T @ Sm, Lm], [Lm.T, -Dm]])
                Bl = np.hstack([sigma_w * Sm, Ym])
                W = W - Bl @ np.linalg.solve(Mid, Bl.T)
                self._Bl, self._Mid = Bl, Mid

            # ---- error measures (IpoptCalculatedQuantities)
            glx = gf + Jc.T @ yc + Jd.T @ yd
            glx[xhasL] -= zL
            glx[xhasU] += zU
            gls = -yd.copy()
            gls[hasL] -= vL
            gls[hasU] += vU
            dms = d - s
            dual_inf = max(np.abs(glx).max(), np.abs(gls).max())
            primal_inf = max(np.abs(c).max() if mc else 0.0, np.abs(dms).max())
            compl = max((sL * vL).max() if len(sL) else 0.0, (sU * vU).max() if len(sU) else 0.0,
                        (xsL * zL).max() if len(xsL) else 0.0, (xsU * zU).max() if len(xsU) else 0.0)
            sum_y = np.abs(yc).sum() + np.abs(yd).sum()
            sum_z = vL.sum() + vU.sum() + zL.sum() + zU.sum()
            s_d = max(o.s_max, (sum_y + sum_z) / (mc + md + self.n_bounds)) / o.s_max
            s_c = max(o.s_max, sum_z / max(self.n_bounds, 1)) / o.s_max
            nlp_error = max(dual_inf / s_d, primal_inf, compl / s_c)
            viol = self.unscaled_violation(graw)
            theta = np.abs(c).sum() + np.abs(dms).sum()    # constraint_violation_norm_type 1-norm
            res.trace.append(dict(iter=it, inf_pr=viol, inf_du=dual_inf, mu=mu, dnorm=dnorm, alpha_du=alpha_du,
                                  alpha_pr=alpha_pr, tag=step_tag, ls=ls_count, nlp_error=nlp_error, theta=theta,
                                  sigma_w=sigma_w, n_pairs=len(S), free_mode=free_mode))
            if o.verbose:
                print("%4d %.2e %.2e %5.1f %.2e %.2e %.2e%s %2d  E=%.2e sw=%.3g np=%d %s" % (
                    it, viol, dual_inf, np.log10(mu), dnorm, alpha_du, alpha_pr, step_tag, ls_count, nlp_error,
                    sigma_w, len(S), "" if free_mode else "F"))
            # the absolute tolerances apply to the unscaled problem: dual infeasibility and complementarity carry the objective's scale
            if nlp_error <= o.tol and dual_inf / self.df <= o.dual_inf_tol and viol <= o.constr_viol_tol and compl / self.df <= o.compl_inf_tol:
                status = 0
                break
            if it >= o.max_iter:
                status = -1
                break

            avrg_compl = (float(sL @ vL) + float(sU @ vU) + float(xsL @ zL) + float(xsU @ zU)) / self.n_bounds
            Sig = np.zeros(md)
            Sig[hasL] += vL / sL
            Sig[hasU] += vU / sU
            Sig_x = np.zeros(n)
            Sig_x[xhasL] += zL / xsL
            Sig_x[xhasU] += zU / xsU
            delta_c = o.delta_cd_val * mu ** o.delta_cd_exp
            fac = self.factor(W + np.diag(Sig_x), Sig, Jc, Jd, delta_c)

            # ---- barrier parameter (AdaptiveMuUpdate::UpdateBarrierParameter)
            if mu_max < 0:
                mu_max = o.mu_max_fact * avrg_compl

            def amu_acceptable(th):
                # with f == 0 an entry (-margin, theta_k - margin) is passed only by a smaller violation
                for (ef, eth) in amu_filter:
                    if not (fval <= ef or th <= eth):
                        return False
                return True

            def remember():
                m_ = o.filter_margin_fact * min(o.filter_max_margin, theta)
                amu_filter.append((fval - m_, theta - m_))

            def barrier_error():
                cm = max(np.abs(sL * vL - mu).max() if len(sL) else 0.0, np.abs(sU * vU - mu).max() if len(sU) else 0.0,
                         np.abs(xsL * zL - mu).max() if len(xsL) else 0.0, np.abs(xsU * zU - mu).max() if len(xsU) else 0.0)
                return max(dual_inf / s_d, primal_inf, cm / s_c)

            if not free_mode:
                if amu_acceptable(theta):
                    free_mode = True
                    remember()
                else:
                    if barrier_error() <= o.barrier_tol_factor * mu:
                        new_mu = min(o.mu_linear_decrease_factor * mu, mu ** o.mu_superlinear_decrease_power)
                        new_mu = max(new_mu, min(o.compl_inf_tol, o.tol) / (o.barrier_tol_factor + 1.0))
                        mu, tau = new_mu, max(o.tau_min, 1.0 - new_mu)
                        ls_filter = []
            else:
                if amu_acceptable(theta):
                    remember()
                else:
                    free_mode = False
                    mu = o.adaptive_mu_monotone_init_factor * avrg_compl
                    mu = min(max(mu, mu_min), mu_max)
                    tau = max(o.tau_min, 1.0 - mu)
                    ls_filter = []

            rvL0, rvU0 = sL * vL, sU * vU                  # curr_compl_s_L/U
            rzL0, rzU0 = xsL * zL, xsU * zU                # curr_compl_x_L/U
            zero_n, zero_d = np.zeros(n), np.zeros(md)
            if free_mode:
                tau = max(o.tau_min, 1.0 - nlp_error)
                # QualityFunctionMuOracle::CalculateMu
                aff = self.kkt_solve(fac, Sig, Jd, glx, gls, c, dms, rvL0, rvU0, sL, sU, vL, vU, (rzL0, rzU0, xsL, xsU, zL, zU))
                aff = [-a for a in aff]
                cen = self.kkt_solve(fac, Sig, Jd, zero_n, zero_d, np.zeros(mc), zero_d,
                                     np.full(len(sL), avrg_compl), np.full(len(sU), avrg_compl), sL, sU, vL, vU,
                                     (np.full(len(xsL), avrg_compl), np.full(len(xsU), avrg_compl), xsL, xsU, zL, zU))
                n_dual, n_pri, n_comp = n + md, mc + md, self.n_bounds
                gl2 = float(glx @ glx + gls @ gls)
                pr2 = float(c @ c + dms @ dms)

                def qf(sig):
                    ds_ = aff[1] + sig * cen[1]
                    dsl, dsu = ds_[hasL], -ds_[hasU]
                    dvl, dvu = aff[4] + sig * cen[4], aff[5] + sig * cen[5]
                    dx_ = aff[0] + sig * cen[0]
                    dxl, dxu = dx_[xhasL], -dx_[xhasU]
                    dzl, dzu = aff[6] + sig * cen[6], aff[7] + sig * cen[7]
                    a_p = min(_frac_to_bound(sL, dsl, tau), _frac_to_bound(sU, dsu, tau),
                              _frac_to_bound(xsL, dxl, tau), _frac_to_bound(xsU, dxu, tau))
                    a_d = min(_frac_to_bound(vL, dvl, tau), _frac_to_bound(vU, dvu, tau),
                              _frac_to_bound(zL, dzl, tau), _frac_to_bound(zU, dzu, tau))
                    cl = (sL + a_p * dsl) * (vL + a_d * dvl)
                    cu = (sU + a_p * dsu) * (vU + a_d * dvu)
                    xl_ = (xsL + a_p * dxl) * (zL + a_d * dzl)
                    xu_ = (xsU + a_p * dxu) * (zU + a_d * dzu)
                    return ((1 - a_d) ** 2 * gl2 / n_dual + (1 - a_p) ** 2 * pr2 / n_pri
                            + (float(cl @ cl) + float(cu @ cu) + float(xl_ @ xl_) + float(xu_ @ xu_)) / n_comp)

                def golden(s_up_in, q_up, s_lo_in, q_lo):
                    s_up, s_lo = s_up_in, s_lo_in            # ScaleSigma is the identity (linear search)
                    s_up0, s_lo0 = s_up, s_lo
                    gfac = (3.0 - np.sqrt(5.0)) / 2.0
                    m1 = s_lo + gfac * (s_up - s_lo)
                    m2 = s_lo + (1 - gfac) * (s_up - s_lo)
                    q1, q2 = qf(m1), qf(m2)
                    k = 0
                    while ((s_up - s_lo) >= o.qf_section_sigma_tol * s_up
                           and (1 - min(q_lo, q_up, q1, q2) / max(q_lo, q_up, q1, q2)) >= o.qf_section_qf_tol
                           and k < o.qf_max_section_steps):
                        k += 1
                        if q1 > q2:
                            s_lo, q_lo = m1, q1
                            m1, q1 = m2, q2
                            m2 = s_lo + (1 - gfac) * (s_up - s_lo)
                            q2 = qf(m2)
                        else:
                            s_up, q_up = m2, q2
                            m2, q2 = m1, q1
                            m1 = s_lo + gfac * (s_up - s_lo)
                            q1 = qf(m1)
                    if ((s_up - s_lo) >= o.qf_section_sigma_tol * s_up
                            and (1 - min(q_lo, q_up, q1, q2) / max(q_lo, q_up, q1, q2)) < o.qf_section_qf_tol):
                        qm = min(q_lo, q_up, q1, q2)
                        sg = s_lo if qm == q_lo else m1 if qm == q1 else m2 if qm == q2 else s_up
                    else:
                        if q1 < q2:
                            sg, q = m1, q1
                        else:
                            sg, q = m2, q2
                        if s_up == s_up0:
                            qt = qf(s_up) if q_up < 0 else q_up
                            if qt < q:
                                sg, q = s_up, qt
                        elif s_lo == s_lo0:
                            qt = qf(s_lo) if q_lo < 0 else q_lo
                            if qt < q:
                                sg, q = s_lo, qt
                    return float(sg)

                qf_1 = qf(1.0)
                s_1m = 1.0 - max(1e-4, o.qf_section_sigma_tol)
                qf_1m = qf(s_1m)
                mu_lo = max(mu_min, 0.0)
                if qf_1m > qf_1:
                    s_up = min(o.sigma_max, mu_max / avrg_compl)
                    s_lo = 1.0
                    sig = s_up if s_lo >= s_up else golden(s_up, -100.0, s_lo, qf_1)
                else:
                    s_lo = max(o.sigma_min, mu_lo / avrg_compl)
                    s_up = min(max(s_lo, s_1m), mu_max / avrg_compl)
                    sig = s_lo if s_lo >= s_up else golden(s_up, qf_1m, s_lo, -100.0)
                mu = min(max(sig * avrg_compl, mu_lo), mu_max)
                mu = max(mu, mu_min)
                ls_filter = []
                res.trace[-1]["sigma"] = sig
                res.trace[-1]["avrg_compl"] = avrg_compl

            # ---- search direction (PDSearchDirCalc)
            step = self.kkt_solve(fac, Sig, Jd, glx, gls, c, dms, rvL0 - mu, rvU0 - mu, sL, sU, vL, vU,
                                  (rzL0 - mu, rzU0 - mu, xsL, xsU, zL, zU))
            dx, ds, dyc, dyd, dvL, dvU, dzL, dzU = [-a for a in step]
            dnorm = max(np.abs(dx).max(), np.abs(ds).max())

            # ---- filter line search (BacktrackingLineSearch + FilterLSAcceptor)
            alpha_max = min(_frac_to_bound(sL, ds[hasL], tau), _frac_to_bound(sU, -ds[hasU], tau),
                            _frac_to_bound(xsL, dx[xhasL], tau), _frac_to_bound(xsU, -dx[xhasU], tau))
            alpha_du = min(_frac_to_bound(vL, dvL, tau), _frac_to_bound(vU, dvU, tau),
                           _frac_to_bound(zL, dzL, tau), _frac_to_bound(zU, dzU, tau))

            def barrier(s_, x_):
                a, b = slacks(s_)
                xa, xb_ = xslacks(x_)
                return -mu * (np.log(a).sum() + np.log(b).sum() + np.log(xa).sum() + np.log(xb_).sum())

            phi0 = barrier(s, x) + fval
            gBD = (-mu * (float((ds[hasL] / sL).sum()) - float((ds[hasU] / sU).sum()))
                   - mu * (float((dx[xhasL] / xsL).sum()) - float((dx[xhasU] / xsU).sum())) + float(gf @ dx))
            if theta_max < 0:
                theta_max = o.theta_max_fact * max(1.0, theta)
                theta_min = o.theta_min_fact * max(1.0, theta)
            # minimal step size (FilterLSAcceptor::CalculateAlphaMin)
            gamma_theta, gamma_phi = o.gamma_theta, o.gamma_phi
            alpha_min = gamma_theta
            if gBD < 0:
                alpha_min = min(gamma_theta, gamma_phi * theta / (-gBD))
                if theta <= theta_min:
                    alpha_min = min(alpha_min, o.delta * theta ** o.s_theta / (-gBD) ** o.s_phi)
            alpha_min *= o.alpha_min_frac

            def acceptable_to_filter(th, ph):
                for (eph, eth) in ls_filter:
                    if not (ph <= eph or th <= eth):
                        return False
                return True

            alpha = alpha_max
            ls_count = 0
            accepted = False
            while alpha > alpha_min or ls_count == 0:
                ls_count += 1
                xt, st = x + alpha * dx, s + alpha * ds
                ct, dt_, gt = self.cd(xt)
                th_t = np.abs(ct).sum() + np.abs(dt_ - st).sum()
                ph_t = barrier(st, xt) + self.f(xt)
                ok = False
                if th_t <= theta_max and np.isfinite(ph_t):
                    switching = gBD < 0 and alpha * (-gBD) ** o.s_phi > o.delta * theta ** o.s_theta
                    if theta <= theta_min and switching:
                        ok = ph_t - phi0 <= o.eta_phi * alpha * gBD + 10 * np.finfo(float).eps * abs(phi0)
                        tag = "f"
                    else:
                        eps10 = 10 * np.finfo(float).eps
                        ok = (th_t - (1 - gamma_theta) * theta <= eps10 * abs(theta) or
                              ph_t - phi0 + gamma_phi * theta <= eps10 * abs(phi0))
                        if ok and ph_t > phi0:             # obj_max_inc
                            bas = np.log10(abs(phi0)) if abs(phi0) > 10 else 1.0
                            ok = np.log10(ph_t - phi0) <= o.obj_max_inc + bas
                        tag = "h"
                    if ok:
                        ok = acceptable_to_filter(th_t, ph_t)
                if ok:
                    accepted = True
                    break
                alpha *= o.alpha_red_factor
            if not accepted:
                status = -2                                # restoration phase not restated
                res.trace[-1]["fail"] = "line search"
                break
            # step-type tag as printed: f if switching condition and Armijo hold at the accepted alpha
            switching = gBD < 0 and alpha * (-gBD) ** o.s_phi > o.delta * theta ** o.s_theta
            armijo = ph_t - phi0 <= o.eta_phi * alpha * gBD + 10 * np.finfo(float).eps * abs(phi0)
            step_tag = "f" if (switching and armijo) else "h"
            if not (switching and armijo):
                ls_filter.append((phi0 - gamma_phi * theta, (1 - gamma_theta) * theta))
            alpha_pr = alpha
            # ---- accept
            x, s = xt, st
            c, d, graw = ct, dt_, gt
            yc = yc + alpha * dyc
            yd = yd + alpha * dyd
            vL = vL + alpha_du * dvL
            vU = vU + alpha_du * dvU
            zL = zL + alpha_du * dzL
            zU = zU + alpha_du * dzU
            sL, sU = slacks(s)
            xsL, xsU = xslacks(x)
            # IpoptAlgorithm::correct_bound_multiplier: free mode uses the trial average complementarity
            mu_c = min((float(sL @ vL) + float(sU @ vU) + float(xsL @ zL) + float(xsU @ zU)) / self.n_bounds, 1e3) if free_mode else mu
            vL = np.minimum(np.maximum(vL, mu_c / (o.kappa_sigma * sL)), o.kappa_sigma * mu_c / sL)
            vU = np.minimum(np.maximum(vU, mu_c / (o.kappa_sigma * sU)), o.kappa_sigma * mu_c / sU)
            zL = np.minimum(np.maximum(zL, mu_c / (o.kappa_sigma * xsL)), o.kappa_sigma * mu_c / xsL)
            zU = np.minimum(np.maximum(zU, mu_c / (o.kappa_sigma * xsU)), o.kappa_sigma * mu_c / xsU)
            Jc, Jd = self.jac(x)
            fval, gf = self.f(x, with_grad=True)
            it += 1

        res.status, res.iters = status, it
        res.x = self.full(x)
        res.constr_viol = viol
        res.nlp_error = nlp_error
        res.objective = fval / self.df
        return res
