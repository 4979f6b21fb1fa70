/*
 * towr_ipm.c -- oracle: primal-dual interior-point loop on the restated TOWR NLP.
 * TEST INFRASTRUCTURE (see towr_oracle.h).
 *
 * What the reference runs here is Ipopt 3.11.9 + MUMPS through ifopt
 * (ref: src/main.cpp:444-463; /root/reference/logs/towr_log.out:37,88), whose
 * source is NOT vendored under /root/reference.  This file therefore restates
 * the PUBLISHED algorithm (Waechter & Biegler, Math. Prog. 106, 2006) in the
 * form the CUDA product implements, so that "same inputs -> same plan" can be
 * checked to round-off between CPU and GPU:
 *
 *   - fixed variables (xl == xu) are eliminated (Ipopt make_parameter);
 *   - rows with gl == gu are equalities c(x)=0, the others get slacks
 *     d(x) - s = 0, dL <= s <= dU (|bound| >= 1e19 is infinite);
 *   - gradient-based row scaling min(1, 100/||grad||_inf) at x0;
 *   - bound_relax 1e-8, slack push 0.01/0.01, bound multipliers 1, y = 0;
 *   - Hessian of the Lagrangian replaced by sigma_w * I (the reference runs
 *     L-BFGS with the same identity start; f == 0, ref: src/parameters.cc:62-63);
 *   - equality block regularised by -delta_c I and condensed, so one SPD
 *     system  (sigma I + Jd' Sigma Jd + Jc'Jc/delta_c) dx = rhs  is factored
 *     per iteration with a skyline Cholesky in reverse-Cuthill-McKee order;
 *   - fraction-to-boundary tau = max(0.99, 1-mu); l1-merit backtracking line
 *     search; monotone (Fiacco-McCormick) barrier update; Ipopt's scaled
 *     termination test (tol 1e-3, constr_viol 1e-4, compl 1e-4, dual_inf 1).
 *
 * Iterate-level parity with Ipopt itself is UNPINNED (different Hessian
 * model / barrier schedule; a feasibility problem has no unique answer).
 */
#include "towr_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define IPM_INF 1e19

void orc_ipm_default_options(orc_ipm_options *o)
{
	o->tol = 1e-3;
	o->constr_viol_tol = 1e-4;
	o->compl_inf_tol = 1e-4;
	o->dual_inf_tol = 1.0;
	o->max_iter = 200;
	o->mu_init = 0.1;
	o->mu_strategy = 0;
	o->sigma_w = 0.1;
	o->verbose = 0;
	o->delta_c = 1e-5;
	o->feas_exit = 1;
}

/* ---------------------------------------------------------------- structure */

#include "towr_sparse.h"

static void rcm_order(int n, const unsigned char *adj /* n*n */, int *perm)
{
	int *deg = (int *)calloc(n, sizeof(int));
	int *visited = (int *)calloc(n, sizeof(int));
	int *queue = (int *)malloc(sizeof(int) * n);
	int *nb = (int *)malloc(sizeof(int) * n);
	for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) if (i != j && adj[(size_t)i * n + j]) deg[i]++;
	int cnt = 0;
	while (cnt < n) {
		int start = -1;
		for (int i = 0; i < n; ++i) if (!visited[i] && (start < 0 || deg[i] < deg[start])) start = i;
		/* pseudo-peripheral refinement: repeat BFS from the farthest min-degree node */
		for (int rep = 0; rep < 3; ++rep) {
			int *lvl = (int *)malloc(sizeof(int) * n);
			for (int i = 0; i < n; ++i) lvl[i] = -1;
			int qh = 0, qt = 0; queue[qt++] = start; lvl[start] = 0;
			while (qh < qt) {
				int u = queue[qh++];
				for (int v = 0; v < n; ++v)
					if (adj[(size_t)u * n + v] && lvl[v] < 0 && !visited[v]) { lvl[v] = lvl[u] + 1; queue[qt++] = v; }
			}
			int far = start, maxl = 0;
			for (int k = 0; k < qt; ++k) {
				int u = queue[k];
				if (lvl[u] > maxl || (lvl[u] == maxl && deg[u] < deg[far])) { maxl = lvl[u]; far = u; }
			}
			free(lvl);
			if (far == start) break;
			start = far;
		}
		int qh = cnt, qt = cnt;
		perm[qt++] = start; visited[start] = 1;
		while (qh < qt) {
			int u = perm[qh++], k = 0;
			for (int v = 0; v < n; ++v) if (adj[(size_t)u * n + v] && !visited[v]) { nb[k++] = v; visited[v] = 1; }
			for (int a = 1; a < k; ++a) {       /* insertion sort by degree */
				int v = nb[a], b = a - 1;
				while (b >= 0 && deg[nb[b]] > deg[v]) { nb[b + 1] = nb[b]; b--; }
				nb[b + 1] = v;
			}
			for (int a = 0; a < k; ++a) perm[qt++] = nb[a];
		}
		cnt = qt;
	}
	for (int i = 0; i < n / 2; ++i) { int t = perm[i]; perm[i] = perm[n - 1 - i]; perm[n - 1 - i] = t; }
	free(deg); free(visited); free(queue); free(nb);
}

ipm_struct *orc_build_struct(orc_problem *p, const double *x0)
{
	ipm_struct *S = (ipm_struct *)calloc(1, sizeof(ipm_struct));
	const int na = p->n, m = p->m;
	S->m = m;
	S->free_of = (int *)malloc(sizeof(int) * na);
	S->var_of = (int *)malloc(sizeof(int) * na);
	int n = 0;
	for (int i = 0; i < na; ++i) {
		if (p->xl[i] != p->xu[i]) { S->free_of[i] = n; S->var_of[n] = i; n++; } else S->free_of[i] = -1;
	}
	S->n = n;
	double *J = (double *)malloc(sizeof(double) * (size_t)m * na);
	unsigned char *mask = (unsigned char *)malloc((size_t)m * na);
	orc_eval_jac(p, x0, J, mask);
	S->rowptr = (int *)malloc(sizeof(int) * (m + 1));
	int nnz = 0;
	for (int r = 0; r < m; ++r) for (int c = 0; c < na; ++c) if (mask[(size_t)r * na + c] && S->free_of[c] >= 0) nnz++;
	S->col = (int *)malloc(sizeof(int) * nnz);
	nnz = 0;
	for (int r = 0; r < m; ++r) {
		S->rowptr[r] = nnz;
		for (int c = 0; c < na; ++c) if (mask[(size_t)r * na + c] && S->free_of[c] >= 0) S->col[nnz++] = S->free_of[c];
	}
	S->rowptr[m] = nnz;
	free(J); free(mask);
	/* pattern of J'J */
	unsigned char *adj = (unsigned char *)calloc((size_t)n * n, 1);
	for (int r = 0; r < m; ++r)
		for (int a = S->rowptr[r]; a < S->rowptr[r + 1]; ++a)
			for (int b = S->rowptr[r]; b < S->rowptr[r + 1]; ++b)
				adj[(size_t)S->col[a] * n + S->col[b]] = 1;
	S->perm = (int *)malloc(sizeof(int) * n);
	S->iperm = (int *)malloc(sizeof(int) * n);
	rcm_order(n, adj, S->perm);
	for (int i = 0; i < n; ++i) S->iperm[S->perm[i]] = i;
	S->first = (int *)malloc(sizeof(int) * n);
	S->skyptr = (long *)malloc(sizeof(long) * (n + 1));
	for (int i = 0; i < n; ++i) S->first[i] = i;
	for (int i = 0; i < n; ++i)
		for (int j = 0; j < n; ++j)
			if (adj[(size_t)i * n + j]) {
				int pi = S->iperm[i], pj = S->iperm[j];
				if (pj < S->first[pi]) S->first[pi] = pj;
			}
	S->skyptr[0] = 0;
	for (int i = 0; i < n; ++i) S->skyptr[i + 1] = S->skyptr[i] + (i - S->first[i] + 1);
	free(adj);
	return S;
}

void orc_free_struct(ipm_struct *S)
{
	free(S->free_of); free(S->var_of); free(S->rowptr); free(S->col);
	free(S->perm); free(S->iperm); free(S->first); free(S->skyptr); free(S);
}

/* skyline Cholesky, row oriented: A[i][j], first[i] <= j <= i at sky[skyptr[i] + j - first[i]] */
int orc_sky_chol(const ipm_struct *S, double *A)
{
	const int n = S->n; int bad = 0;
	for (int i = 0; i < n; ++i) {
		double *ri = A + S->skyptr[i]; const int fi = S->first[i];
		for (int j = fi; j <= i; ++j) {
			const double *rj = A + S->skyptr[j]; const int fj = S->first[j];
			int k0 = fi > fj ? fi : fj;
			double s = ri[j - fi];
			for (int k = k0; k < j; ++k) s -= ri[k - fi] * rj[k - fj];
			if (j < i) ri[j - fi] = s / rj[j - fj];
			else {
				if (!(s > 0.0)) { s = 1e-30; bad = 1; }
				ri[j - fi] = sqrt(s);
			}
		}
	}
	return bad;
}

void orc_sky_fwd(const ipm_struct *S, const double *L, double *b)
{
	const int n = S->n;
	for (int i = 0; i < n; ++i) {
		const double *ri = L + S->skyptr[i]; const int fi = S->first[i];
		double s = b[i];
		for (int k = fi; k < i; ++k) s -= ri[k - fi] * b[k];
		b[i] = s / ri[i - fi];
	}
}

void orc_sky_bwd(const ipm_struct *S, const double *L, double *b)
{
	const int n = S->n;
	for (int i = n - 1; i >= 0; --i) {
		const double *ri = L + S->skyptr[i]; const int fi = S->first[i];
		b[i] /= ri[i - fi];
		const double bi = b[i];
		for (int k = fi; k < i; ++k) b[k] -= ri[k - fi] * bi;
	}
}

void orc_sky_solve(const ipm_struct *S, const double *L, double *b)
{
	orc_sky_fwd(S, L, b);
	orc_sky_bwd(S, L, b);
}

/* ---------------------------------------------------------------- the loop */

static double dmax(double a, double b) { return a > b ? a : b; }
static double dmin(double a, double b) { return a < b ? a : b; }

int orc_ipm_solve(orc_problem *p, const orc_ipm_options *o, double *x, orc_ipm_result *res)
{
	const int na = p->n, m = p->m;
	memset(res, 0, sizeof(*res));
	for (int i = 0; i < na; ++i) if (p->xl[i] == p->xu[i]) x[i] = p->xl[i];
	ipm_struct *S = orc_build_struct(p, x);
	const int n = S->n, nnz = S->rowptr[m];
	const double *gl = p->gl, *gu = p->gu;

	double *Jd = (double *)malloc(sizeof(double) * (size_t)m * na);   /* dense scratch */
	double *jv = (double *)malloc(sizeof(double) * nnz);
	double *sc = (double *)malloc(sizeof(double) * m);
	double *g = (double *)malloc(sizeof(double) * m), *gt = (double *)malloc(sizeof(double) * m);
	/* per-row state; equality rows use c=res, y; inequality rows use all */
	double *r = (double *)calloc(m, sizeof(double));     /* c (eq) or d (ineq), scaled */
	double *rt = (double *)calloc(m, sizeof(double));
	double *s = (double *)calloc(m, sizeof(double)), *st = (double *)calloc(m, sizeof(double));
	double *y = (double *)calloc(m, sizeof(double));
	double *zL = (double *)calloc(m, sizeof(double)), *zU = (double *)calloc(m, sizeof(double));
	double *dL = (double *)calloc(m, sizeof(double)), *dU = (double *)calloc(m, sizeof(double));
	double *Sig = (double *)calloc(m, sizeof(double)), *w = (double *)calloc(m, sizeof(double));
	double *ds = (double *)calloc(m, sizeof(double)), *dy = (double *)calloc(m, sizeof(double));
	double *dzL = (double *)calloc(m, sizeof(double)), *dzU = (double *)calloc(m, sizeof(double));
	unsigned char *iseq = (unsigned char *)calloc(m, 1), *hasL = (unsigned char *)calloc(m, 1), *hasU = (unsigned char *)calloc(m, 1);
	double *rx = (double *)calloc(n, sizeof(double)), *dx = (double *)calloc(n, sizeof(double));
	double *xt = (double *)malloc(sizeof(double) * na);
	double *M = (double *)malloc(sizeof(double) * S->skyptr[n]);

#define GATHER_J() do { orc_eval_jac(p, x, Jd, NULL); \
	for (int r_ = 0; r_ < m; ++r_) for (int a_ = S->rowptr[r_]; a_ < S->rowptr[r_ + 1]; ++a_) \
		jv[a_] = sc[r_] * Jd[(size_t)r_ * na + S->var_of[S->col[a_]]]; } while (0)

	/* scaling (gradient based at x0) */
	for (int i = 0; i < m; ++i) sc[i] = 1.0;
	GATHER_J();
	for (int i = 0; i < m; ++i) {
		double mx = 0.0;
		for (int a = S->rowptr[i]; a < S->rowptr[i + 1]; ++a) mx = dmax(mx, fabs(jv[a]));
		sc[i] = mx > 100.0 ? dmax(100.0 / mx, 1e-8) : 1.0;
	}
	int n_bounds = 0, n_eq = 0, n_iq = 0;
	for (int i = 0; i < m; ++i) {
		iseq[i] = gl[i] == gu[i];
		if (iseq[i]) { n_eq++; continue; }
		n_iq++;
		hasL[i] = gl[i] > -IPM_INF; hasU[i] = gu[i] < IPM_INF;
		if (hasL[i]) { double b = sc[i] * gl[i]; dL[i] = b - 1e-8 * dmax(1.0, fabs(b)); n_bounds++; }
		if (hasU[i]) { double b = sc[i] * gu[i]; dU[i] = b + 1e-8 * dmax(1.0, fabs(b)); n_bounds++; }
	}
	orc_eval_g(p, x, g);
	for (int i = 0; i < m; ++i) {
		if (iseq[i]) { r[i] = sc[i] * (g[i] - gl[i]); continue; }
		r[i] = sc[i] * g[i];
		double v = r[i];
		if (hasL[i]) {
			double push = 0.01 * dmax(1.0, fabs(dL[i]));
			if (hasU[i]) push = dmin(push, 0.01 * (dU[i] - dL[i]));
			v = dmax(v, dL[i] + push);
		}
		if (hasU[i]) {
			double push = 0.01 * dmax(1.0, fabs(dU[i]));
			if (hasL[i]) push = dmin(push, 0.01 * (dU[i] - dL[i]));
			v = dmin(v, dU[i] - push);
		}
		s[i] = v;
		zL[i] = hasL[i] ? 1.0 : 0.0; zU[i] = hasU[i] ? 1.0 : 0.0;
	}
	double mu = o->mu_init, nu = 1.0;
	const double rho = 1.0 / o->delta_c, sigma = o->sigma_w;
	const double kappa_eps = 10.0, kappa_mu = 0.2, theta_mu = 1.5;
	const double mu_min = dmin(o->tol, o->compl_inf_tol) / (kappa_eps + 1.0);
	int status = -1, it = 0, nfail = 0;

	for (it = 0; ; ++it) {
		GATHER_J();
		/* residuals and error measures */
		memset(rx, 0, sizeof(double) * n);
		for (int i = 0; i < m; ++i) for (int a = S->rowptr[i]; a < S->rowptr[i + 1]; ++a) rx[S->col[a]] += jv[a] * y[i];
		double theta_inf = 0, dual_inf = 0, compl0 = 0, sum_y = 0, sum_z = 0, viol = 0;
		for (int i = 0; i < n; ++i) dual_inf = dmax(dual_inf, fabs(rx[i]));
		for (int i = 0; i < m; ++i) {
			sum_y += fabs(y[i]);
			if (iseq[i]) { theta_inf = dmax(theta_inf, fabs(r[i])); viol = dmax(viol, fabs(r[i]) / sc[i]); continue; }
			theta_inf = dmax(theta_inf, fabs(r[i] - s[i]));
			dual_inf = dmax(dual_inf, fabs(-y[i] - zL[i] + zU[i]));
			if (hasL[i]) { compl0 = dmax(compl0, zL[i] * (s[i] - dL[i])); sum_z += zL[i]; viol = dmax(viol, gl[i] - r[i] / sc[i]); }
			if (hasU[i]) { compl0 = dmax(compl0, zU[i] * (dU[i] - s[i])); sum_z += zU[i]; viol = dmax(viol, r[i] / sc[i] - gu[i]); }
		}
		const double s_d = dmax(100.0, (sum_y + sum_z) / (double)(m + n_bounds)) / 100.0;
		const double s_c = dmax(100.0, sum_z / (double)(n_bounds > 0 ? n_bounds : 1)) / 100.0;
		const double E0 = dmax(dmax(dual_inf / s_d, theta_inf), compl0 / s_c);
		if (it < 256) {
			res->tr_inf_pr[it] = theta_inf; res->tr_inf_du[it] = dual_inf; res->tr_mu[it] = mu;
			res->n_trace = it + 1;
		}
		res->constr_viol = viol; res->dual_inf = dual_inf; res->compl_inf = compl0; res->nlp_error = E0; res->mu = mu;
		if (o->verbose)
			printf("%3d inf_pr=%9.3e inf_du=%9.3e lg(mu)=%5.1f viol=%9.3e E0=%9.3e", it, theta_inf, dual_inf, log10(mu), viol, E0);
		/* f == 0 on this path (ref: src/parameters.cc:62-63): every feasible point is a KKT point with
		 * zero multipliers, so E0 evaluated at (x, y=0, z=0) is just the primal infeasibility */
		const int feas = o->feas_exit && viol <= o->constr_viol_tol && theta_inf <= o->tol;
		if (feas || (E0 <= o->tol && viol <= o->constr_viol_tol && compl0 <= o->compl_inf_tol && dual_inf <= o->dual_inf_tol)) {
			status = 0; if (o->verbose) printf("\n"); break;
		}
		if (it >= o->max_iter) { if (o->verbose) printf("\n"); break; }
		/* monotone barrier update */
		for (;;) {
			double cm = 0;
			for (int i = 0; i < m; ++i) {
				if (hasL[i]) cm = dmax(cm, fabs(zL[i] * (s[i] - dL[i]) - mu));
				if (hasU[i]) cm = dmax(cm, fabs(zU[i] * (dU[i] - s[i]) - mu));
			}
			double Emu = dmax(dmax(dual_inf / s_d, theta_inf), cm / s_c);
			if (Emu <= kappa_eps * mu && mu > mu_min) mu = dmax(mu_min, dmin(kappa_mu * mu, pow(mu, theta_mu)));
			else break;
		}
		const double tau = dmax(0.99, 1.0 - mu);
		/* Sigma, condensed rhs weights w_i; rhs = -J' w */
		for (int i = 0; i < m; ++i) {
			if (iseq[i]) { Sig[i] = rho; w[i] = y[i] + rho * r[i]; continue; }
			double sg = 0, rsm = -y[i];
			if (hasL[i]) { sg += zL[i] / (s[i] - dL[i]); rsm -= mu / (s[i] - dL[i]); }
			if (hasU[i]) { sg += zU[i] / (dU[i] - s[i]); rsm += mu / (dU[i] - s[i]); }
			Sig[i] = sg;
			dy[i] = rsm;                               /* stash r_s^mu */
			w[i] = y[i] + sg * (r[i] - s[i]) + rsm;
		}
		memset(M, 0, sizeof(double) * S->skyptr[n]);
		for (int i = 0; i < n; ++i) M[S->skyptr[i] + i - S->first[i]] = sigma;
		memset(dx, 0, sizeof(double) * n);
		for (int i = 0; i < m; ++i) {
			const double D = Sig[i];
			for (int a = S->rowptr[i]; a < S->rowptr[i + 1]; ++a) {
				const int pa = S->iperm[S->col[a]]; const double va = jv[a];
				dx[pa] -= va * w[i];
				if (va == 0.0) continue;
				for (int b = S->rowptr[i]; b < S->rowptr[i + 1]; ++b) {
					const int pb = S->iperm[S->col[b]];
					if (pb > pa) continue;
					M[S->skyptr[pa] + pb - S->first[pa]] += D * va * jv[b];
				}
			}
		}
		if (orc_sky_chol(S, M)) res->chol_fix++;
		orc_sky_solve(S, M, dx);                               /* dx in permuted order */
		/* recover ds, dy, dz */
		double a_pr = 1.0, a_du = 1.0, theta1 = 0, gphi_d = 0, quad = 0, dxmax = 0;
		for (int i = 0; i < n; ++i) { quad += sigma * dx[i] * dx[i]; dxmax = dmax(dxmax, fabs(dx[i])); }
		for (int i = 0; i < m; ++i) {
			double jdx = 0;
			for (int a = S->rowptr[i]; a < S->rowptr[i + 1]; ++a) jdx += jv[a] * dx[S->iperm[S->col[a]]];
			if (iseq[i]) { dy[i] = rho * (jdx + r[i]); theta1 += fabs(r[i]); continue; }
			const double rsm = dy[i];
			ds[i] = jdx + (r[i] - s[i]);
			dy[i] = Sig[i] * ds[i] + rsm;
			theta1 += fabs(r[i] - s[i]);
			quad += ds[i] * Sig[i] * ds[i];
			if (hasL[i]) {
				const double sl = s[i] - dL[i];
				dzL[i] = mu / sl - zL[i] - zL[i] / sl * ds[i];
				gphi_d += -mu / sl * ds[i];
				if (ds[i] < 0) a_pr = dmin(a_pr, -tau * sl / ds[i]);
				if (dzL[i] < 0) a_du = dmin(a_du, -tau * zL[i] / dzL[i]);
			}
			if (hasU[i]) {
				const double su = dU[i] - s[i];
				dzU[i] = mu / su - zU[i] + zU[i] / su * ds[i];
				gphi_d += mu / su * ds[i];
				if (ds[i] > 0) a_pr = dmin(a_pr, tau * su / ds[i]);
				if (dzU[i] < 0) a_du = dmin(a_du, -tau * zU[i] / dzU[i]);
			}
		}
		if (theta1 > 1e-14) {
			const double nu_trial = (gphi_d + 0.5 * quad) / (0.7 * theta1);
			if (nu < nu_trial) nu = nu_trial + 1.0;
		}
		double bar0 = 0;
		for (int i = 0; i < m; ++i) {
			if (hasL[i]) bar0 -= mu * log(s[i] - dL[i]);
			if (hasU[i]) bar0 -= mu * log(dU[i] - s[i]);
		}
		const double phi0 = bar0 + nu * theta1, Dphi = gphi_d - nu * theta1;
		double alpha = a_pr; int ls = 0;
		for (;;) {
			memcpy(xt, x, sizeof(double) * na);
			for (int i = 0; i < n; ++i) xt[S->var_of[S->perm[i]]] += alpha * dx[i];
			orc_eval_g(p, xt, gt);
			double th = 0, bar = 0;
			for (int i = 0; i < m; ++i) {
				if (iseq[i]) { rt[i] = sc[i] * (gt[i] - gl[i]); th += fabs(rt[i]); continue; }
				rt[i] = sc[i] * gt[i];
				st[i] = s[i] + alpha * ds[i];
				th += fabs(rt[i] - st[i]);
				if (hasL[i]) bar -= mu * log(st[i] - dL[i]);
				if (hasU[i]) bar -= mu * log(dU[i] - st[i]);
			}
			ls++;
			if (bar + nu * th <= phi0 + 1e-4 * alpha * Dphi || ls >= 12) break;
			alpha *= 0.5;
		}
		if (o->verbose) printf(" |dx|=%8.2e a_pr=%8.2e a_du=%8.2e ls=%d nu=%8.2e\n", dxmax, alpha, a_du, ls, nu);
		nfail = ls >= 12 ? nfail + 1 : 0;
		if (it < 256) { res->tr_dnorm[it] = dxmax; res->tr_alpha_pr[it] = alpha; res->tr_alpha_du[it] = a_du; res->tr_ls[it] = ls; }
		memcpy(x, xt, sizeof(double) * na);
		for (int i = 0; i < m; ++i) {
			r[i] = rt[i];
			y[i] += alpha * dy[i];
			if (iseq[i]) continue;
			s[i] = st[i];
			if (hasL[i]) {
				const double sl = s[i] - dL[i];
				zL[i] = dmin(dmax(zL[i] + a_du * dzL[i], mu / (1e10 * sl)), 1e10 * mu / sl);
			}
			if (hasU[i]) {
				const double su = dU[i] - s[i];
				zU[i] = dmin(dmax(zU[i] + a_du * dzU[i], mu / (1e10 * su)), 1e10 * mu / su);
			}
		}
		if (nfail >= 3) { status = -2; break; }      /* line search stalled three times in a row */
	}
	res->status = status; res->iters = it;
	(void)n_eq; (void)n_iq;
	free(Jd); free(jv); free(sc); free(g); free(gt); free(r); free(rt); free(s); free(st); free(y);
	free(zL); free(zU); free(dL); free(dU); free(Sig); free(w); free(ds); free(dy); free(dzL); free(dzU);
	free(iseq); free(hasL); free(hasU); free(rx); free(dx); free(xt); free(M);
	orc_free_struct(S);
	return status;
}
