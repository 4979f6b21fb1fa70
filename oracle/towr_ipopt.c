/*
 * towr_ipopt.c -- oracle: the Ipopt 3.11.9 algorithm the reference runs (ref: src/main.cpp:444-463;
 * /root/reference/logs/towr_log.out:37), restated in plain C in the form the CUDA product implements.
 * TEST INFRASTRUCTURE (see towr_oracle.h).
 *
 * Ipopt's source is not vendored under /root/reference.  The algorithm below is the published one
 * (Waechter & Biegler, Math. Prog. 106, 2006; Nocedal, Waechter & Waltz, SIAM J. Optim. 19, 2009) with the
 * options in force on the reference path (ifopt: tol 1e-3, limited-memory Hessian, mumps; the logs show
 * mu_strategy adaptive).  oracle/ipopt_emul.py states the same algorithm with a dense, exact KKT solve and is
 * PINNED to the reference's golden data: it reproduces the three iteration tables of logs/towr_log.out to the
 * printed digits and the plans of data/traj/towr.csv to 2e-6 m (tests/test_ipopt_emulation.py).  This file is
 * pinned to that emulator and to the same golden data (tests/test_oracle_ipopt_c.py).
 *
 * Differences from the emulator are confined to the linear algebra, chosen to match the CUDA kernels:
 *   - inequality slacks and their multipliers are condensed; the equality block is treated by a penalty
 *     rho = 1/delta_c plus `n_refine` multiplier-method passes (each one more solve with the same factor):
 *         M = sigma_w I + Jd' Sigma Jd + rho Jc' Jc          (SPD, skyline Cholesky in RCM order)
 *         dx_{k+1} = Mf^-1 (b1 + Jc' (rho b2 - dy_k)),  dy_{k+1} = dy_k + rho (Jc dx_{k+1} - b2)
 *   - the limited-memory term  W = sigma_w I - Bl Mid^-1 Bl'  (compact BFGS, Bl = [sigma_w S, Y]) enters by the
 *     Woodbury identity, written between the two triangular solves of M = L L' (the kernels get Q = L^-1 Bl as
 *     extra right-hand sides of the factorization's forward substitution, and never need M^-1 Bl):
 *         Mf^-1 v = L'^-1 (p + Q (Mid - Q'Q)^-1 Q' p),   p = L^-1 v,   Q = L^-1 Bl;
 *   - the affine-scaling and centering directions of the quality-function mu oracle are two right-hand sides
 *     of the same factorization; the search direction is aff + (mu / avg_compl) cen (exact: the KKT
 *     right-hand side is affine in mu).
 * Not restated: second-order correction, watchdog, restoration phase (never entered on the reference's
 * logged runs: every logged iteration has ls = 1); a line search that fails ends the solve with status -2.
 */
#include "towr_sparse.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define IPM_INF 1e19
#define LM_MAX 6

void orc_ipopt_default_options(orc_ipopt_options *o)
{
	o->tol = 1e-3; o->constr_viol_tol = 1e-4; o->compl_inf_tol = 1e-4; o->dual_inf_tol = 1.0;
	o->max_iter = 200;
	o->delta_c = 1e-6; o->n_refine = 1;
	o->lm_history = 6;
	o->sigma_floor = 0.0;
	o->verbose = 0;
	o->retry_failed = 1; o->retry_sigma_floor = 1e-2;
}

static double dmax(double a, double b) { return a > b ? a : b; }
static double dmin(double a, double b) { return a < b ? a : b; }

/* dense LU with partial pivoting, n <= 12; A (row-major, n*n) is overwritten; returns 0 if singular */
static int lu_factor(int n, double *A, int *piv)
{
	for (int k = 0; k < n; ++k) {
		int p = k;
		for (int i = k + 1; i < n; ++i) if (fabs(A[i * n + k]) > fabs(A[p * n + k])) p = i;
		piv[k] = p;
		if (A[p * n + k] == 0.0) return 0;
		if (p != k) for (int j = 0; j < n; ++j) { double t = A[k * n + j]; A[k * n + j] = A[p * n + j]; A[p * n + j] = t; }
		for (int i = k + 1; i < n; ++i) {
			A[i * n + k] /= A[k * n + k];
			for (int j = k + 1; j < n; ++j) A[i * n + j] -= A[i * n + k] * A[k * n + j];
		}
	}
	return 1;
}

static void lu_solve(int n, const double *A, const int *piv, double *b)
{
	for (int k = 0; k < n; ++k) { double t = b[k]; b[k] = b[piv[k]]; b[piv[k]] = t; }     /* P b (full-row swaps were applied to L too) */
	for (int k = 0; k < n; ++k) for (int i = k + 1; i < n; ++i) b[i] -= A[i * n + k] * b[k];
	for (int k = n - 1; k >= 0; --k) { for (int j = k + 1; j < n; ++j) b[k] -= A[k * n + j] * b[j]; b[k] /= A[k * n + k]; }
}

typedef struct {
	const ipm_struct *S;
	int n, m;
	const double *jv;                 /* scaled Jacobian values (CSR) */
	const unsigned char *iseq, *hasL, *hasU;
	double *Lsky;                     /* factor of M */
	int nlr;                          /* columns of Bl (2 x stored pairs) */
	double *Bl, *Z;                   /* [nlr][n]; Z = Q = L^-1 Bl in permuted order */
	double Clu[4 * LM_MAX * LM_MAX]; int Cpiv[2 * LM_MAX];
	double rho;
	int n_refine;
	double *tmp;                      /* [n] */
} kkt_t;

/* u = Mf^-1 v with the limited-memory term (Woodbury between the triangular solves); free order in and out */
static void mf_solve(const kkt_t *K, const double *v, double *u)
{
	const ipm_struct *S = K->S;
	const int n = K->n;
	for (int i = 0; i < n; ++i) K->tmp[S->iperm[i]] = v[i];
	orc_sky_fwd(S, K->Lsky, K->tmp);                              /* p = L^-1 v (permuted order, like Q) */
	if (K->nlr) {
		double t[2 * LM_MAX];
		for (int a = 0; a < K->nlr; ++a) { double s = 0; for (int i = 0; i < n; ++i) s += K->Z[(size_t)a * n + i] * K->tmp[i]; t[a] = s; }
		lu_solve(K->nlr, K->Clu, K->Cpiv, t);
		for (int a = 0; a < K->nlr; ++a) for (int i = 0; i < n; ++i) K->tmp[i] += K->Z[(size_t)a * n + i] * t[a];
	}
	orc_sky_bwd(S, K->Lsky, K->tmp);
	for (int i = 0; i < n; ++i) u[i] = K->tmp[S->iperm[i]];
}

/* One primal-dual solve (PDFullSpaceSolver::SolveOnce in condensed form).  Row-indexed inputs: rc_d[i] = rhs of the
 * c row (eq) or d row (ineq), rs / rvL / rvU on inequality rows.  Outputs dx[n], and per row ds, dy, dvL, dvU. */
static void kkt_solve(const kkt_t *K, const double *rx, const double *rs, const double *rcd, const double *rvL, const double *rvU,
                      const double *Sig, const double *sL, const double *sU, const double *vL, const double *vU,
                      double *dx, double *ds, double *dy, double *dvL, double *dvU, double *w /* [m] scratch */, double *v /* [n] scratch */)
{
	const ipm_struct *S = K->S;
	const int n = K->n, m = K->m;
	for (int i = 0; i < m; ++i) dy[i] = 0.0;
	for (int pass = 0; pass <= K->n_refine; ++pass) {
		for (int i = 0; i < m; ++i) {
			if (K->iseq[i]) { w[i] = K->rho * rcd[i] - dy[i]; continue; }
			double aug = rs[i];
			if (K->hasL[i]) aug += rvL[i] / sL[i];
			if (K->hasU[i]) aug -= rvU[i] / sU[i];
			w[i] = Sig[i] * rcd[i] + aug;
		}
		for (int j = 0; j < n; ++j) v[j] = rx ? rx[j] : 0.0;
		for (int i = 0; i < m; ++i) for (int a = S->rowptr[i]; a < S->rowptr[i + 1]; ++a) v[S->col[a]] += K->jv[a] * w[i];
		mf_solve(K, v, dx);
		double rmax = 0, bmax = 0; int imax = -1;
		for (int i = 0; i < m; ++i) {
			if (!K->iseq[i]) continue;
			double jdx = 0; for (int a = S->rowptr[i]; a < S->rowptr[i + 1]; ++a) jdx += K->jv[a] * dx[S->col[a]];
			dy[i] += K->rho * (jdx - rcd[i]);
			if (fabs(jdx - rcd[i]) > rmax) { rmax = fabs(jdx - rcd[i]); imax = i; }
			bmax = dmax(bmax, fabs(rcd[i]));
		}
		if (getenv("ORC_DEBUG_KKT")) printf("      pass %d: max |Jc dx - b2| = %.3e (row %d) max|b2| = %.3e\n", pass, rmax, imax, bmax);
	}
	for (int i = 0; i < m; ++i) {
		if (K->iseq[i]) { ds[i] = dvL[i] = dvU[i] = 0.0; continue; }
		double jdx = 0; for (int a = S->rowptr[i]; a < S->rowptr[i + 1]; ++a) jdx += K->jv[a] * dx[S->col[a]];
		double aug = rs[i];
		if (K->hasL[i]) aug += rvL[i] / sL[i];
		if (K->hasU[i]) aug -= rvU[i] / sU[i];
		ds[i] = jdx - rcd[i];
		dy[i] = Sig[i] * ds[i] - aug;
		dvL[i] = K->hasL[i] ? (rvL[i] - vL[i] * ds[i]) / sL[i] : 0.0;
		dvU[i] = K->hasU[i] ? (rvU[i] + vU[i] * ds[i]) / sU[i] : 0.0;
	}
}

/* DenseVector::FracToBound over both bound sides: slack + alpha * dslack >= (1 - tau) * slack */
static double frac_to_bound(int m, const unsigned char *hasL, const unsigned char *hasU, const double *aL, const double *dL_,
                            const double *aU, const double *dU_, double tau)
{
	double al = 1.0;
	for (int i = 0; i < m; ++i) {
		if (hasL[i] && dL_[i] < 0) al = dmin(al, -tau * aL[i] / dL_[i]);
		if (hasU[i] && dU_[i] < 0) al = dmin(al, -tau * aU[i] / dU_[i]);
	}
	return al;
}

static int ipopt_attempt(orc_problem *p, const orc_ipopt_options *o, double *x, orc_ipopt_result *res)
{
	const int na = p->n, m = p->m;
	memset(res, 0, sizeof(*res));
	for (int i = 0; i < na; ++i) if (p->xl[i] == p->xu[i]) x[i] = p->xl[i];
	ipm_struct *S = orc_build_struct(p, x);
	const int n = S->n, nnz = S->rowptr[m];
	const double *gl = p->gl, *gu = p->gu;
	const int hist = o->lm_history < LM_MAX ? o->lm_history : LM_MAX;

#define DV(name, len) double *name = (double *)calloc((size_t)(len) > 0 ? (size_t)(len) : 1, sizeof(double))
	DV(Jd, (size_t)m * na); DV(jv, nnz); DV(jv_old, nnz); DV(sc, m); DV(g, m); DV(gt, m);
	DV(r, m); DV(rt, m); DV(s, m); DV(st, m); DV(y, m); DV(vL, m); DV(vU, m); DV(dL, m); DV(dU, m);
	DV(sL, m); DV(sU, m); DV(Sig, m); DV(w, m);
	DV(rs, m); DV(rcd, m); DV(rvL, m); DV(rvU, m);
	DV(a_ds, m); DV(a_dy, m); DV(a_dvL, m); DV(a_dvU, m); DV(c_ds, m); DV(c_dy, m); DV(c_dvL, m); DV(c_dvU, m);
	DV(ds, m); DV(dy, m); DV(dvL, m); DV(dvU, m); DV(tL, m); DV(tU, m); DV(uL, m); DV(uU, m);
	DV(glx, n); DV(a_dx, n); DV(c_dx, n); DV(dx, n); DV(vtmp, n); DV(tmp, n); DV(last_x, n); DV(gJold, n); DV(xf, n);
	DV(M, S->skyptr[n]); DV(xt, na);
	DV(gfull, na); DV(gf, n);
	DV(Sm, (size_t)LM_MAX * n); DV(Ym, (size_t)LM_MAX * n); DV(Bl, (size_t)2 * LM_MAX * n); DV(Z, (size_t)2 * LM_MAX * n);
	unsigned char *iseq = (unsigned char *)calloc(m, 1), *hasL = (unsigned char *)calloc(m, 1), *hasU = (unsigned char *)calloc(m, 1);

#define GATHER_J(dst) do { orc_eval_jac(p, x, Jd, NULL); \
	for (int r_ = 0; r_ < m; ++r_) for (int a_ = S->rowptr[r_]; a_ < S->rowptr[r_ + 1]; ++a_) \
		(dst)[a_] = sc[r_] * Jd[(size_t)r_ * na + S->var_of[S->col[a_]]]; } while (0)
#define JT_TIMES(J_, vec_, out_) do { memset((out_), 0, sizeof(double) * n); \
	for (int r_ = 0; r_ < m; ++r_) for (int a_ = S->rowptr[r_]; a_ < S->rowptr[r_ + 1]; ++a_) (out_)[S->col[a_]] += (J_)[a_] * (vec_)[r_]; } while (0)

	/* gradient-based scaling at x0 (nlp_scaling_max_gradient 100, min value 1e-8) */
	for (int i = 0; i < m; ++i) sc[i] = 1.0;
	GATHER_J(jv);
	for (int i = 0; i < m; ++i) {
		double mx = 0.0;
		for (int a = S->rowptr[i]; a < S->rowptr[i + 1]; ++a) mx = dmax(mx, fabs(jv[a]));
		sc[i] = mx > 100.0 ? dmax(100.0 / mx, 1e-8) : 1.0;
	}
	GATHER_J(jv);
	int n_bounds = 0, n_eq = 0, n_iq = 0;
	for (int i = 0; i < m; ++i) {
		iseq[i] = gl[i] == gu[i];
		if (iseq[i]) { n_eq++; continue; }
		n_iq++;
		hasL[i] = gl[i] > -IPM_INF; hasU[i] = gu[i] < IPM_INF;
		/* bound_relax_factor 1e-8 on the unscaled bound, then scaled */
		if (hasL[i]) { dL[i] = sc[i] * (gl[i] - 1e-8 * dmax(1.0, fabs(gl[i]))); n_bounds++; }
		if (hasU[i]) { dU[i] = sc[i] * (gu[i] + 1e-8 * dmax(1.0, fabs(gu[i]))); n_bounds++; }
	}
	orc_eval_g(p, x, g);
	for (int i = 0; i < m; ++i) {
		if (iseq[i]) { r[i] = sc[i] * (g[i] - gl[i]); continue; }
		r[i] = sc[i] * g[i];
		double v = r[i];
		const double width = hasL[i] && hasU[i] ? dU[i] - dL[i] : 1e300;
		if (hasL[i]) v = dmax(v, dL[i] + dmin(0.01 * dmax(1.0, fabs(dL[i])), 0.01 * width));
		if (hasU[i]) v = dmin(v, dU[i] - dmin(0.01 * dmax(1.0, fabs(dU[i])), 0.01 * width));
		s[i] = v;
		vL[i] = hasL[i] ? 1.0 : 0.0; vU[i] = hasU[i] ? 1.0 : 0.0;
	}
	for (int i = 0; i < n; ++i) xf[i] = x[S->var_of[i]];
	/* optional cost terms (orc_shape.cost_*; f == 0 on the reference's path): objective scaling like the rows',
	 * min(1, 100 / ||grad f(x0)||_inf) over the free variables */
	const int has_cost = p->shape.cost_force_z != 0.0 || p->shape.cost_ee_vel_xy != 0.0;
	double df = 1.0, fval = 0.0;
	if (has_cost) {
		orc_eval_cost(p, x, gfull);
		double mx = 0.0;
		for (int i = 0; i < n; ++i) mx = dmax(mx, fabs(gfull[S->var_of[i]]));
		if (mx > 100.0) df = dmax(100.0 / mx, 1e-8);
	}

	kkt_t K; memset(&K, 0, sizeof(K));
	K.S = S; K.n = n; K.m = m; K.jv = jv; K.iseq = iseq; K.hasL = hasL; K.hasU = hasU; K.Lsky = M; K.Bl = Bl; K.Z = Z;
	K.rho = 1.0 / o->delta_c; K.n_refine = o->n_refine; K.tmp = tmp;

	/* algorithm state */
	int n_pairs = 0, lm_skipped = 0, have_last = 0;
	double sigma_w = 1.0;                                  /* limited_memory_init_val */
	double mu = 1.0, tau = 0.0;                            /* AdaptiveMuUpdate::InitializeImpl */
	int free_mode = 1;
	double mu_max = -1.0;
	const double mu_min = dmin(1e-11, 0.5 * dmin(o->tol, o->compl_inf_tol));
	/* AdaptiveMuUpdate's own filter of (f, theta) pairs: a point passes when, against every entry, it is no larger in at least
	 * one coordinate.  With f == 0 an entry (-margin, theta_k - margin) is passed only by theta <= theta_k - margin. */
	double amu_f[ORC_FILTER_MAX], amu_th[ORC_FILTER_MAX]; int n_amu = 0;
	double fphi[ORC_FILTER_MAX], fth[ORC_FILTER_MAX]; int nfilter = 0;   /* line-search filter */
	double theta_max = -1.0, theta_min = -1.0;
	int status = -1, it = 0, ls_count = 0;
	double delta_w_last = 0.0;
	double alpha_pr = 0, alpha_du = 0, dnorm = 0;
	char tag = ' ';
	const double eps10 = 10.0 * 2.220446049250313e-16;

	for (;;) {
		/* ---- slacks of the bounds */
		for (int i = 0; i < m; ++i) { if (hasL[i]) sL[i] = s[i] - dL[i]; if (hasU[i]) sU[i] = dU[i] - s[i]; }
		/* ---- limited-memory update (LimMemQuasiNewtonUpdater::UpdateHessian): s = x+ - x, y = (J+ - J)' lambda+ */
		JT_TIMES(jv, y, glx);
		if (has_cost) {                                       /* grad_x L = grad f + J' lambda */
			fval = df * orc_eval_cost(p, x, gfull);
			for (int i = 0; i < n; ++i) { gf[i] = df * gfull[S->var_of[i]]; glx[i] += gf[i]; }
		}
		if (have_last) {
			double sTy = 0, sTs = 0, yTy = 0;
			for (int i = 0; i < n; ++i) { const double sn = xf[i] - last_x[i], yn = glx[i] - gJold[i]; vtmp[i] = sn; tmp[i] = yn; sTy += sn * yn; sTs += sn * sn; yTy += yn * yn; }
			const int skipping = sTy <= sqrt(2.220446049250313e-16) * sqrt(sTs) * sqrt(yTy);
			if (skipping) {
				if (++lm_skipped >= 2) { n_pairs = 0; sigma_w = 1.0; lm_skipped = 0; }
			} else {
				lm_skipped = 0;
				if (n_pairs == hist) {
					memmove(Sm, Sm + n, sizeof(double) * (size_t)(hist - 1) * n); memmove(Ym, Ym + n, sizeof(double) * (size_t)(hist - 1) * n);
					n_pairs--;
				}
				memcpy(Sm + (size_t)n_pairs * n, vtmp, sizeof(double) * n); memcpy(Ym + (size_t)n_pairs * n, tmp, sizeof(double) * n);
				n_pairs++;
				sigma_w = dmin(dmax(sTy / sTs, 1e-8), 1e8);
			}
		}
		memcpy(last_x, xf, sizeof(double) * n); have_last = 1;

		/* ---- error measures (IpoptCalculatedQuantities) */
		double dual_inf = 0, primal_inf = 0, compl = 0, sum_y = 0, sum_z = 0, viol = 0, theta = 0, cs = 0;
		for (int i = 0; i < n; ++i) dual_inf = dmax(dual_inf, fabs(glx[i]));
		for (int i = 0; i < m; ++i) {
			sum_y += fabs(y[i]);
			if (iseq[i]) { primal_inf = dmax(primal_inf, fabs(r[i])); theta += fabs(r[i]); viol = dmax(viol, fabs(r[i]) / sc[i]); continue; }
			primal_inf = dmax(primal_inf, fabs(r[i] - s[i])); theta += fabs(r[i] - s[i]);
			dual_inf = dmax(dual_inf, fabs(-y[i] - vL[i] + vU[i]));
			if (hasL[i]) { compl = dmax(compl, sL[i] * vL[i]); cs += sL[i] * vL[i]; sum_z += vL[i]; viol = dmax(viol, gl[i] - r[i] / sc[i]); }
			if (hasU[i]) { compl = dmax(compl, sU[i] * vU[i]); cs += sU[i] * vU[i]; sum_z += vU[i]; viol = dmax(viol, r[i] / sc[i] - gu[i]); }
		}
		const double s_d = dmax(100.0, (sum_y + sum_z) / (double)(m + n_bounds)) / 100.0;
		const double s_c = dmax(100.0, sum_z / (double)(n_bounds > 0 ? n_bounds : 1)) / 100.0;
		const double nlp_error = dmax(dmax(dual_inf / s_d, primal_inf), compl / s_c);
		const double avrg_compl = cs / (double)(n_bounds > 0 ? n_bounds : 1);
		if (it < ORC_TRACE_MAX) {
			res->tr_inf_pr[it] = viol; res->tr_inf_du[it] = dual_inf; res->tr_mu[it] = mu; res->tr_dnorm[it] = dnorm;
			res->tr_alpha_pr[it] = alpha_pr; res->tr_alpha_du[it] = alpha_du; res->tr_ls[it] = ls_count; res->tr_tag[it] = tag;
			res->tr_pairs[it] = n_pairs; res->tr_free[it] = free_mode; res->n_trace = it + 1;
		}
		res->constr_viol = viol; res->dual_inf = dual_inf; res->compl_inf = compl; res->nlp_error = nlp_error; res->mu = mu;
		if (o->verbose) printf("%4d %.2e %.2e %5.1f %.2e %.2e %.2e%c %2d  E=%.2e sw=%.3g np=%d %s\n", it, viol, dual_inf, log10(mu), dnorm,
		                       alpha_du, alpha_pr, tag, ls_count, nlp_error, sigma_w, n_pairs, free_mode ? "" : "F");
		if (!(nlp_error == nlp_error) || !(theta == theta)) { status = -13; break; }
		/* the absolute tolerances apply to the unscaled problem: dual infeasibility and complementarity carry the objective's scale */
		if (nlp_error <= o->tol && dual_inf / df <= o->dual_inf_tol && viol <= o->constr_viol_tol && compl / df <= o->compl_inf_tol) { status = 0; break; }
		if (it >= o->max_iter) { status = -1; break; }

		/* ---- barrier parameter (AdaptiveMuUpdate::UpdateBarrierParameter) */
		if (mu_max < 0) mu_max = 1e3 * avrg_compl;
		int acceptable = 1;
		for (int k = 0; k < n_amu; ++k) if (!(fval <= amu_f[k] || theta <= amu_th[k])) { acceptable = 0; break; }
		if (!free_mode) {
			if (acceptable) { free_mode = 1; }
			else {
				double cm = 0;
				for (int i = 0; i < m; ++i) { if (hasL[i]) cm = dmax(cm, fabs(sL[i] * vL[i] - mu)); if (hasU[i]) cm = dmax(cm, fabs(sU[i] * vU[i] - mu)); }
				const double berr = dmax(dmax(dual_inf / s_d, primal_inf), cm / s_c);
				if (berr <= 10.0 * mu) {
					double nm = dmin(0.2 * mu, pow(mu, 1.5));
					nm = dmax(nm, dmin(o->compl_inf_tol, o->tol) / 11.0);
					mu = nm; tau = dmax(0.99, 1.0 - mu); nfilter = 0;
				}
			}
		} else if (!acceptable) {
			free_mode = 0;
			mu = dmin(dmax(0.8 * avrg_compl, mu_min), mu_max);
			tau = dmax(0.99, 1.0 - mu); nfilter = 0;
		}
		if (free_mode && acceptable) {                        /* RememberCurrentPointAsAccepted */
			const double mg = 1e-5 * dmin(1.0, theta);
			if (mg > 0.0) {
				const double ef = fval - mg, eth = theta - mg;
				int w = 0;                                        /* entries the new one dominates leave (Filter::AddEntry) */
				for (int k = 0; k < n_amu; ++k) if (!(ef <= amu_f[k] && eth <= amu_th[k])) { amu_f[w] = amu_f[k]; amu_th[w] = amu_th[k]; w++; }
				n_amu = w;
				if (n_amu == ORC_FILTER_MAX) { memmove(amu_f, amu_f + 1, sizeof(double) * (ORC_FILTER_MAX - 1)); memmove(amu_th, amu_th + 1, sizeof(double) * (ORC_FILTER_MAX - 1)); n_amu--; }
				amu_f[n_amu] = ef; amu_th[n_amu] = eth; n_amu++;
			}
		}

		/* ---- factorization and the two directions; a non-positive pivot or a non-finite direction repeats them with
		 *      W + delta_w I (PDPerturbationHandler: 1e-4 the first time, a third of the last successful value later, then
		 *      x100 / x8, give up beyond 1e40) */
		const double sigma_f = dmax(sigma_w, o->sigma_floor);
		double delta_w = 0.0;
		int gave_up = 0;
		for (;;) {
			/* ---- factor M = (sigma_w + delta_w) I + Jd' Sigma Jd + rho Jc' Jc, Q = L^-1 Bl, C = Mid - Q'Q */
			for (int i = 0; i < m; ++i) {
				if (iseq[i]) { Sig[i] = K.rho; continue; }
				Sig[i] = (hasL[i] ? vL[i] / sL[i] : 0.0) + (hasU[i] ? vU[i] / sU[i] : 0.0);
			}
			memset(M, 0, sizeof(double) * S->skyptr[n]);
			for (int i = 0; i < n; ++i) M[S->skyptr[i] + i - S->first[i]] = sigma_f + delta_w;
			for (int i = 0; i < m; ++i) {
				const double D = Sig[i];
				for (int a = S->rowptr[i]; a < S->rowptr[i + 1]; ++a) {
					const int pa = S->iperm[S->col[a]]; const double va = jv[a];
					if (va == 0.0) continue;
					for (int b = S->rowptr[i]; b < S->rowptr[i + 1]; ++b) {
						const int pb = S->iperm[S->col[b]];
						if (pb <= pa) M[S->skyptr[pa] + pb - S->first[pa]] += D * va * jv[b];
					}
				}
			}
			const int bad_pivot = orc_sky_chol(S, M);
			if (bad_pivot) res->chol_fix++;
			K.nlr = 2 * n_pairs;
			if (n_pairs) {
				double Mid[4 * LM_MAX * LM_MAX];
				const int q = 2 * n_pairs;
				for (int a = 0; a < n_pairs; ++a) for (int i = 0; i < n; ++i) { Bl[(size_t)a * n + i] = sigma_f * Sm[(size_t)a * n + i]; Bl[(size_t)(n_pairs + a) * n + i] = Ym[(size_t)a * n + i]; }
				for (int a = 0; a < n_pairs; ++a) for (int b = 0; b < n_pairs; ++b) {
					double ss = 0, sy = 0;
					for (int i = 0; i < n; ++i) { ss += Sm[(size_t)a * n + i] * Sm[(size_t)b * n + i]; sy += Sm[(size_t)a * n + i] * Ym[(size_t)b * n + i]; }
					Mid[a * q + b] = sigma_f * ss;
					Mid[a * q + n_pairs + b] = a > b ? sy : 0.0;             /* L: strictly lower part of S'Y */
					Mid[(n_pairs + b) * q + a] = a > b ? sy : 0.0;           /* L' */
					Mid[(n_pairs + a) * q + n_pairs + b] = a == b ? -sy : 0.0;   /* -D */
				}
				for (int a = 0; a < q; ++a) {
					for (int i = 0; i < n; ++i) Z[(size_t)a * n + S->iperm[i]] = Bl[(size_t)a * n + i];
					orc_sky_fwd(S, M, Z + (size_t)a * n);                 /* Q = L^-1 Bl */
				}
				for (int a = 0; a < q; ++a) for (int b = 0; b < q; ++b) {
					double t = 0; for (int i = 0; i < n; ++i) t += Z[(size_t)a * n + i] * Z[(size_t)b * n + i];
					K.Clu[a * q + b] = Mid[a * q + b] - t;
				}
				if (getenv("ORC_DEBUG_LM")) { printf("  Mid:"); for (int a = 0; a < q * q; ++a) printf(" %.6e", Mid[a]); printf("\n  C:"); for (int a = 0; a < q * q; ++a) printf(" %.6e", K.Clu[a]); printf("\n"); }
				if (!lu_factor(q, K.Clu, K.Cpiv)) { K.nlr = 0; n_pairs = 0; sigma_w = 1.0; }
			}

			/* ---- affine-scaling and centering directions (two right-hand sides of one factorization) */
			for (int i = 0; i < n; ++i) vtmp[i] = -glx[i];
			for (int i = 0; i < m; ++i) {
				if (iseq[i]) { rcd[i] = -r[i]; rs[i] = rvL[i] = rvU[i] = 0.0; continue; }
				rcd[i] = -(r[i] - s[i]);
				rs[i] = -(-y[i] - vL[i] + vU[i]);
				rvL[i] = hasL[i] ? -sL[i] * vL[i] : 0.0; rvU[i] = hasU[i] ? -sU[i] * vU[i] : 0.0;
			}
			kkt_solve(&K, vtmp, rs, rcd, rvL, rvU, Sig, sL, sU, vL, vU, a_dx, a_ds, a_dy, a_dvL, a_dvU, w, xt /* scratch >= n */);
			for (int i = 0; i < m; ++i) { rcd[i] = rs[i] = 0.0; rvL[i] = hasL[i] ? avrg_compl : 0.0; rvU[i] = hasU[i] ? avrg_compl : 0.0; }
			kkt_solve(&K, NULL, rs, rcd, rvL, rvU, Sig, sL, sU, vL, vU, c_dx, c_ds, c_dy, c_dvL, c_dvU, w, xt);

			double nf = 0.0;
			for (int i = 0; i < n; ++i) nf += 0.0 * a_dx[i] + 0.0 * c_dx[i];
			if (nf == 0.0 && !bad_pivot) break;
			delta_w = delta_w == 0.0 ? (delta_w_last == 0.0 ? 1e-4 : dmax(1e-20, delta_w_last / 3.0)) : delta_w * (delta_w_last == 0.0 ? 100.0 : 8.0);
			if (delta_w > 1e40) { gave_up = 1; break; }
			res->n_regularized++;
			if (o->verbose) printf("  regularisation: delta_w = %.3e\n", delta_w);
		}
		if (gave_up) { status = -2; break; }
		if (delta_w > 0.0) delta_w_last = delta_w;

		double sigma = mu / avrg_compl;
		if (free_mode) {
			tau = dmax(0.99, 1.0 - nlp_error);
			/* QualityFunctionMuOracle::CalculateMu: 2-norm-squared quality function, golden section on sigma (linear) */
			double gl2 = 0, pr2 = 0;
			for (int i = 0; i < n; ++i) gl2 += glx[i] * glx[i];
			for (int i = 0; i < m; ++i) {
				if (iseq[i]) { pr2 += r[i] * r[i]; continue; }
				const double gs = -y[i] - vL[i] + vU[i]; gl2 += gs * gs; pr2 += (r[i] - s[i]) * (r[i] - s[i]);
			}
			const double n_dual = n + n_iq, n_pri = m, n_comp = n_bounds;
#define QF(sig_, out_) do { const double sg_ = (sig_); \
			for (int i = 0; i < m; ++i) { if (iseq[i]) continue; const double d_ = a_ds[i] + sg_ * c_ds[i]; tL[i] = d_; tU[i] = -d_; \
				uL[i] = a_dvL[i] + sg_ * c_dvL[i]; uU[i] = a_dvU[i] + sg_ * c_dvU[i]; } \
			const double ap_ = frac_to_bound(m, hasL, hasU, sL, tL, sU, tU, tau), ad_ = frac_to_bound(m, hasL, hasU, vL, uL, vU, uU, tau); \
			double cc_ = 0; for (int i = 0; i < m; ++i) { \
				if (hasL[i]) { const double t_ = (sL[i] + ap_ * tL[i]) * (vL[i] + ad_ * uL[i]); cc_ += t_ * t_; } \
				if (hasU[i]) { const double t_ = (sU[i] + ap_ * tU[i]) * (vU[i] + ad_ * uU[i]); cc_ += t_ * t_; } } \
			(out_) = (1 - ad_) * (1 - ad_) * gl2 / n_dual + (1 - ap_) * (1 - ap_) * pr2 / n_pri + cc_ / n_comp; } while (0)
			double qf_1, qf_1m; const double s_1m = 1.0 - 1e-2;
			QF(1.0, qf_1); QF(s_1m, qf_1m);
			double s_up, s_lo, q_up, q_lo; int search = 1;
			if (qf_1m > qf_1) { s_up = dmin(100.0, mu_max / avrg_compl); s_lo = 1.0; q_up = -100.0; q_lo = qf_1; if (s_lo >= s_up) { sigma = s_up; search = 0; } }
			else { s_lo = dmax(1e-6, mu_min / avrg_compl); s_up = dmin(dmax(s_lo, s_1m), mu_max / avrg_compl); q_up = qf_1m; q_lo = -100.0; if (s_lo >= s_up) { sigma = s_lo; search = 0; } }
			if (search) {
				const double s_up0 = s_up, s_lo0 = s_lo, gfac = (3.0 - sqrt(5.0)) / 2.0;
				double m1 = s_lo + gfac * (s_up - s_lo), m2 = s_lo + (1 - gfac) * (s_up - s_lo), q1, q2;
				QF(m1, q1); QF(m2, q2);
				int k = 0;
				while ((s_up - s_lo) >= 1e-2 * s_up && k < 8) {      /* quality_function_section_qf_tol 0: no early exit */
					k++;
					if (q1 > q2) { s_lo = m1; q_lo = q1; m1 = m2; q1 = q2; m2 = s_lo + (1 - gfac) * (s_up - s_lo); QF(m2, q2); }
					else { s_up = m2; q_up = q2; m2 = m1; q2 = q1; m1 = s_lo + gfac * (s_up - s_lo); QF(m1, q1); }
				}
				double q;
				if (q1 < q2) { sigma = m1; q = q1; } else { sigma = m2; q = q2; }
				if (s_up == s_up0) { double qt = q_up; if (qt < 0) QF(s_up, qt); if (qt < q) { sigma = s_up; q = qt; } }
				else if (s_lo == s_lo0) { double qt = q_lo; if (qt < 0) QF(s_lo, qt); if (qt < q) { sigma = s_lo; q = qt; } }
			}
			mu = dmax(dmin(dmax(sigma * avrg_compl, mu_min), mu_max), mu_min);
			sigma = mu / avrg_compl;
			nfilter = 0;
		}

		/* ---- search direction aff + sigma cen */
		dnorm = 0;
		for (int i = 0; i < n; ++i) { dx[i] = a_dx[i] + sigma * c_dx[i]; dnorm = dmax(dnorm, fabs(dx[i])); }
		for (int i = 0; i < m; ++i) {
			dy[i] = a_dy[i] + sigma * c_dy[i];
			if (iseq[i]) continue;
			ds[i] = a_ds[i] + sigma * c_ds[i]; dvL[i] = a_dvL[i] + sigma * c_dvL[i]; dvU[i] = a_dvU[i] + sigma * c_dvU[i];
			dnorm = dmax(dnorm, fabs(ds[i]));
			tL[i] = ds[i]; tU[i] = -ds[i];
		}
		const double alpha_max = frac_to_bound(m, hasL, hasU, sL, tL, sU, tU, tau);
		alpha_du = frac_to_bound(m, hasL, hasU, vL, dvL, vU, dvU, tau);

		/* ---- filter line search (BacktrackingLineSearch + FilterLSAcceptor) */
		double phi0 = 0, gBD = 0;
		for (int i = 0; i < m; ++i) {
			if (hasL[i]) { phi0 -= mu * log(sL[i]); gBD -= mu * ds[i] / sL[i]; }
			if (hasU[i]) { phi0 -= mu * log(sU[i]); gBD += mu * ds[i] / sU[i]; }
		}
		if (has_cost) { phi0 += fval; for (int i = 0; i < n; ++i) gBD += gf[i] * dx[i]; }
		if (theta_max < 0) { theta_max = 1e4 * dmax(1.0, theta); theta_min = 1e-4 * dmax(1.0, theta); }
		double alpha_min = 1e-5;
		if (gBD < 0) {
			alpha_min = dmin(1e-5, 1e-8 * theta / (-gBD));
			if (theta <= theta_min) alpha_min = dmin(alpha_min, pow(theta, 1.1) / pow(-gBD, 2.3));
		}
		alpha_min *= 0.05;
		double alpha = alpha_max, th_t = 0, ph_t = 0;
		int accepted = 0;
		ls_count = 0;
		while (alpha > alpha_min || ls_count == 0) {
			ls_count++;
			memcpy(xt, x, sizeof(double) * na);
			for (int i = 0; i < n; ++i) xt[S->var_of[i]] += alpha * dx[i];
			orc_eval_g(p, xt, gt);
			th_t = 0; ph_t = 0;
			for (int i = 0; i < m; ++i) {
				if (iseq[i]) { rt[i] = sc[i] * (gt[i] - gl[i]); th_t += fabs(rt[i]); continue; }
				rt[i] = sc[i] * gt[i]; st[i] = s[i] + alpha * ds[i];
				th_t += fabs(rt[i] - st[i]);
				if (hasL[i]) ph_t -= mu * log(st[i] - dL[i]);
				if (hasU[i]) ph_t -= mu * log(dU[i] - st[i]);
			}
			if (has_cost) ph_t += df * orc_eval_cost(p, xt, NULL);
			int ok = 0;
			if (th_t <= theta_max && ph_t == ph_t && fabs(ph_t) < 1e300) {
				const int switching = gBD < 0 && alpha * pow(-gBD, 2.3) > pow(theta, 1.1);
				if (theta <= theta_min && switching) ok = ph_t - phi0 - 1e-8 * alpha * gBD <= eps10 * fabs(phi0);
				else {
					ok = th_t - (1 - 1e-5) * theta <= eps10 * fabs(theta) || ph_t - phi0 + 1e-8 * theta <= eps10 * fabs(phi0);
					if (ok && ph_t > phi0) {                         /* obj_max_inc 5 */
						const double bas = fabs(phi0) > 10.0 ? log10(fabs(phi0)) : 1.0;
						ok = log10(ph_t - phi0) <= 5.0 + bas;
					}
				}
				if (ok) for (int k = 0; k < nfilter; ++k) if (!(ph_t <= fphi[k] || th_t <= fth[k])) { ok = 0; break; }
			}
			if (ok) { accepted = 1; break; }
			if (o->verbose > 1) printf("   trial alpha %.3e th_t %.6e ph_t %.6e (theta %.6e phi0 %.6e)\n", alpha, th_t, ph_t, theta, phi0);
			alpha *= 0.5;
		}
		if (!accepted) { if (o->verbose) printf("  LS failed: alpha_max %.3e alpha_min %.3e theta %.3e th_t %.3e phi0 %.6e ph_t %.6e gBD %.3e theta_min %.3e nfilter %d dnorm %.3e\n", alpha_max, alpha_min, theta, th_t, phi0, ph_t, gBD, theta_min, nfilter, dnorm); status = -2; break; }
		{
			const int switching = gBD < 0 && alpha * pow(-gBD, 2.3) > pow(theta, 1.1);
			const int armijo = ph_t - phi0 - 1e-8 * alpha * gBD <= eps10 * fabs(phi0);
			tag = switching && armijo ? 'f' : 'h';
			if (!(switching && armijo)) {
				if (nfilter == ORC_FILTER_MAX) { memmove(fphi, fphi + 1, sizeof(double) * (ORC_FILTER_MAX - 1)); memmove(fth, fth + 1, sizeof(double) * (ORC_FILTER_MAX - 1)); nfilter--; }
				fphi[nfilter] = phi0 - 1e-8 * theta; fth[nfilter] = (1 - 1e-5) * theta; nfilter++;
			}
		}
		alpha_pr = alpha;
		/* ---- accept the trial point */
		memcpy(x, xt, sizeof(double) * na);
		for (int i = 0; i < n; ++i) xf[i] = x[S->var_of[i]];
		double cs_t = 0;
		for (int i = 0; i < m; ++i) {
			r[i] = rt[i];
			y[i] += alpha * dy[i];
			if (iseq[i]) continue;
			s[i] = st[i];
			if (hasL[i]) { sL[i] = s[i] - dL[i]; vL[i] += alpha_du * dvL[i]; cs_t += sL[i] * vL[i]; }
			if (hasU[i]) { sU[i] = dU[i] - s[i]; vU[i] += alpha_du * dvU[i]; cs_t += sU[i] * vU[i]; }
		}
		/* IpoptAlgorithm::correct_bound_multiplier (kappa_sigma 1e10): free mode uses the trial average complementarity */
		const double mu_c = free_mode ? dmin(cs_t / (double)(n_bounds > 0 ? n_bounds : 1), 1e3) : mu;
		for (int i = 0; i < m; ++i) {
			if (hasL[i]) vL[i] = dmin(dmax(vL[i], mu_c / (1e10 * sL[i])), 1e10 * mu_c / sL[i]);
			if (hasU[i]) vU[i] = dmin(dmax(vU[i], mu_c / (1e10 * sU[i])), 1e10 * mu_c / sU[i]);
		}
		JT_TIMES(jv, y, gJold);                              /* grad_x L(x_k, lambda_{k+1}) for the next limited-memory pair */
		if (has_cost) for (int i = 0; i < n; ++i) gJold[i] += gf[i];
		GATHER_J(jv);
		it++;
	}
	res->status = status; res->iters = it;
	res->objective = has_cost ? orc_eval_cost(p, x, NULL) : 0.0;
	(void)n_eq; (void)jv_old;
	free(Jd); free(jv); free(jv_old); free(sc); free(g); free(gt); free(r); free(rt); free(s); free(st); free(y); free(vL); free(vU);
	free(dL); free(dU); free(sL); free(sU); free(Sig); free(w); free(rs); free(rcd); free(rvL); free(rvU);
	free(a_ds); free(a_dy); free(a_dvL); free(a_dvU); free(c_ds); free(c_dy); free(c_dvL); free(c_dvU);
	free(ds); free(dy); free(dvL); free(dvU); free(tL); free(tU); free(uL); free(uU);
	free(glx); free(a_dx); free(c_dx); free(dx); free(vtmp); free(tmp); free(last_x); free(gJold); free(xf); free(M); free(xt);
	free(gfull); free(gf); free(Sm); free(Ym); free(Bl); free(Z); free(iseq); free(hasL); free(hasU);
	orc_free_struct(S);
	return status;
}

/* One attempt, and -- like the CUDA product (qtos_options.retry_failed) -- a second one from the same x0 with Ipopt's
 * limited_memory_init_val_min raised to retry_sigma_floor when the filter line search of the first fails (status -2, where
 * Ipopt would enter its restoration phase).  The iteration count is the sum of both attempts. */
int orc_ipopt_solve(orc_problem *p, const orc_ipopt_options *o, double *x, orc_ipopt_result *res)
{
	double *x0 = (double *)malloc(sizeof(double) * p->n);
	memcpy(x0, x, sizeof(double) * p->n);
	int st = ipopt_attempt(p, o, x, res);
	if (st == -2 && o->retry_failed && o->sigma_floor < o->retry_sigma_floor) {
		const int first = res->iters;
		orc_ipopt_options o2 = *o;
		o2.sigma_floor = o->retry_sigma_floor;
		memcpy(x, x0, sizeof(double) * p->n);
		st = ipopt_attempt(p, &o2, x, res);
		res->iters += first;
		res->retried = 1;
	}
	free(x0);
	return st;
}
