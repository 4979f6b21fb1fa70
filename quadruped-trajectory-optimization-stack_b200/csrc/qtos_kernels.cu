/*
 * qtos_kernels.cu -- sm_100a kernels of the batched interior-point gait-plan solver.
 *
 * One thread block per problem in every kernel; a batch iteration is five launches
 * (jac -> prepare -> assemble -> factor -> step), all on the context's stream:
 *   k_init      x0, fixed variables, g(x0), J(x0), row scaling, slack/multiplier init
 *               (ref: nlp_formulation.cc:100-190; Ipopt initialisation)
 *   k_jac       dynamics + range-of-motion Jacobian element blocks at x
 *   k_prepare   J'y, error measures, termination test, barrier update, Sigma, rhs = -J'w
 *   k_assemble  M = sigma I + J' D J into block-skyline storage (owner-computes gather)
 *   k_factor    blocked left-looking Cholesky, 16x16 blocks, panels staged in shared memory
 *   k_step      triangular solves, step recovery, fraction-to-boundary, l1-merit backtracking
 *               line search with in-kernel g(x) evaluations, iterate update
 *   k_csv       1 kHz trajectory sampler (ref: main.cpp:92-131)
 *   k_height    batched heightfield queries (ref: custom_terrain.cpp:51-94)
 * The interior-point algorithm is the one restated in oracle/towr_ipm.c (test oracle).
 */
#include "qtos_device.cuh"

#define QTOS_THREADS 256
#define NB QTOS_NB

/* ------------------------------------------------------------------ reductions */

template <int N>
__device__ __forceinline__ void block_reduce(double (&v)[N], const int (&op)[N], double *smem /* >= N*32 */)
{
	/* op: 0 sum, 1 max, 2 min; result broadcast to all threads */
	for (int i = 0; i < N; ++i)
		for (int o = 16; o > 0; o >>= 1) {
			double t = __shfl_xor_sync(0xffffffffu, v[i], o);
			v[i] = op[i] == 0 ? v[i] + t : (op[i] == 1 ? fmax(v[i], t) : fmin(v[i], t));
		}
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
	__syncthreads();
	if (lane == 0) for (int i = 0; i < N; ++i) smem[i * 32 + warp] = v[i];
	__syncthreads();
	for (int i = 0; i < N; ++i) {
		double a = smem[i * 32];
		for (int w = 1; w < nw; ++w) { double t = smem[i * 32 + w]; a = op[i] == 0 ? a + t : (op[i] == 1 ? fmax(a, t) : fmin(a, t)); }
		v[i] = a;
	}
	__syncthreads();
}

#define WS(arr, len) (W.arr + (size_t)pid * (len))

__device__ __forceinline__ void scaled_rows(const DevTables &T, const double *sc, double *r /* in: raw g, out: scaled */)
{
	for (int i = threadIdx.x; i < T.m; i += blockDim.x) {
		const double g = r[i];
		r[i] = (T.row_flags[i] & ROW_EQ) ? sc[i] * (g - T.gl[i]) : sc[i] * g;
	}
}

/* ------------------------------------------------------------------ k_init */

__global__ void __launch_bounds__(QTOS_THREADS)
k_init(DevTables T, DevWork W, const qtos_problem *probs, const DevHeightfield *hfs, int n_hf, qtos_options opt, int do_solver_init)
{
	const int pid = blockIdx.x;
	const qtos_problem &pr = probs[pid];
	double *x = WS(x, T.n_all), *P = WS(P, 32), *sc = WS(sc, T.m), *r = WS(r, T.m), *Jv = WS(Jv, T.nJ);
	__shared__ double sfin[6][3];
	const int hid = pr.hf_id >= 0 && pr.hf_id < n_hf ? pr.hf_id : 0;
	const DevHeightfield hf = hfs[hid];
	if (threadIdx.x == 0) {
		for (int d = 0; d < 3; ++d) {
			P[QP_START_POS + d] = pr.start_pos[d]; P[QP_START_ANG + d] = pr.start_ang[d];
			P[QP_START_VEL + d] = pr.start_vel[d]; P[QP_START_ANGVEL + d] = pr.start_ang_vel[d];
			P[QP_GOAL + d] = pr.goal[d];
			for (int e = 0; e < QTOS_NEE; ++e) P[QP_EE + 3 * e + d] = pr.ee[e][d];
		}
		P[QP_ZERO] = 0.0;
		/* end points of the linear initial guess (ref: nlp_formulation.cc:100-190) */
		sfin[0][0] = pr.goal[0]; sfin[0][1] = pr.goal[1];
		sfin[0][2] = qtos_height(hf, pr.goal[0], pr.goal[1]) - T.nominal[0][2];
		sfin[1][0] = sfin[1][1] = sfin[1][2] = 0.0;
		for (int e = 0; e < QTOS_NEE; ++e) {
			const double fx = pr.goal[0] + T.nominal[e][0], fy = pr.goal[1] + T.nominal[e][1];
			sfin[2 + e][0] = fx; sfin[2 + e][1] = fy; sfin[2 + e][2] = qtos_height(hf, fx, fy);
		}
	}
	__syncthreads();
	for (int v = threadIdx.x; v < T.n_all; v += blockDim.x) {
		const int fs = T.fix_src[v];
		if (fs >= 0) { x[v] = P[fs]; continue; }
		const int s = T.x0_spline[v], dim = T.x0_dim[v], node = T.x0_node[v];
		double a, b;
		if (s == 0) { a = pr.start_pos[dim]; b = sfin[0][dim]; }
		else if (s == 1) { a = pr.start_ang[dim]; b = 0.0; }
		else if (s < 6) { a = pr.ee[s - 2][dim]; b = sfin[s][dim]; }
		else { a = b = dim == 2 ? T.mass * T.grav / QTOS_NEE : 0.0; }
		const double dp = b - a;
		x[v] = T.x0_deriv[v] == 0 ? a + node / (double)(T.n_nodes[s] - 1) * dp : dp / T.T;
	}
	if (!do_solver_init) return;
	for (int i = threadIdx.x; i < T.m; i += blockDim.x) sc[i] = 1.0;
	__syncthreads();
	eval_g_block(T, hf, x, r);
	for (int i = threadIdx.x; i < T.nJ; i += blockDim.x) Jv[i] = T.Jconst[i];
	__syncthreads();
	eval_jac_block(T, x, sc, Jv);
	__syncthreads();
	/* gradient-based row scaling min(1, 100/||row||_inf) at x0 */
	for (int i = threadIdx.x; i < T.m; i += blockDim.x) {
		const Element &E = T.elems[T.row_elem[i]];
		const int rr = i - E.row0;
		double mx = 0.0;
		for (int a = 0; a < E.ncols; ++a) mx = fmax(mx, fabs(Jv[E.valoff + a * E.nrows + rr]));
		sc[i] = mx > 100.0 ? fmax(100.0 / mx, 1e-8) : 1.0;
	}
	__syncthreads();
	for (int e = threadIdx.x; e < T.n_elem; e += blockDim.x) {
		const Element &E = T.elems[e];
		for (int a = 0; a < E.ncols; ++a) for (int rr = 0; rr < E.nrows; ++rr) Jv[E.valoff + a * E.nrows + rr] *= sc[E.row0 + rr];
	}
	double *s = WS(s, T.m), *y = WS(y, T.m), *zL = WS(zL, T.m), *zU = WS(zU, T.m), *dL = WS(dL, T.m), *dU = WS(dU, T.m);
	for (int i = threadIdx.x; i < T.m; i += blockDim.x) {
		const int fl = T.row_flags[i];
		const double g = r[i];
		y[i] = 0.0; zL[i] = zU[i] = 0.0; dL[i] = dU[i] = 0.0; s[i] = 0.0;
		if (fl & ROW_EQ) { r[i] = sc[i] * (g - T.gl[i]); continue; }
		r[i] = sc[i] * g;
		double v = r[i], lo = 0, hi = 0;
		if (fl & ROW_HASL) { const double b = sc[i] * T.gl[i]; lo = b - 1e-8 * fmax(1.0, fabs(b)); dL[i] = lo; zL[i] = 1.0; }
		if (fl & ROW_HASU) { const double b = sc[i] * T.gu[i]; hi = b + 1e-8 * fmax(1.0, fabs(b)); dU[i] = hi; zU[i] = 1.0; }
		if (fl & ROW_HASL) {
			double push = 0.01 * fmax(1.0, fabs(lo));
			if (fl & ROW_HASU) push = fmin(push, 0.01 * (hi - lo));
			v = fmax(v, lo + push);
		}
		if (fl & ROW_HASU) {
			double push = 0.01 * fmax(1.0, fabs(hi));
			if (fl & ROW_HASL) push = fmin(push, 0.01 * (hi - lo));
			v = fmin(v, hi - push);
		}
		s[i] = v;
	}
	if (threadIdx.x == 0) {
		double *scal = WS(scal, 16);
		scal[SC_MU] = opt.mu_init; scal[SC_NU] = 1.0; scal[SC_NFAIL] = 0.0;
		W.status[pid] = QTOS_RUNNING; W.iters[pid] = 0; W.flags[pid] = 0;
	}
}

/* ------------------------------------------------------------------ k_jac */

__global__ void __launch_bounds__(QTOS_THREADS)
k_jac(DevTables T, DevWork W)
{
	const int pid = blockIdx.x;
	if (W.status[pid] != QTOS_RUNNING) return;
	eval_jac_block(T, WS(x, T.n_all), WS(sc, T.m), WS(Jv, T.nJ));
}

/* ------------------------------------------------------------------ k_prepare */

__device__ __forceinline__ double jt_gather(const DevTables &T, const double *Jv, const double *wrow, int i)
{
	double acc = 0.0;
	for (int q = T.jt_ptr[i]; q < T.jt_ptr[i + 1]; ++q) {
		const uint32_t t = T.jt_terms[q];
		const Element &E = T.elems[t >> 8];
		const double *col = Jv + E.valoff + (t & 255u) * E.nrows;
		const double *wr = wrow + E.row0;
		for (int rr = 0; rr < E.nrows; ++rr) acc += col[rr] * wr[rr];
	}
	return acc;
}

__global__ void __launch_bounds__(QTOS_THREADS)
k_prepare(DevTables T, DevWork W, qtos_options opt, int it)
{
	const int pid = blockIdx.x;
	if (W.status[pid] != QTOS_RUNNING) return;
	__shared__ double red[8 * 32];
	const double *Jv = WS(Jv, T.nJ), *r = WS(r, T.m), *s = WS(s, T.m), *zL = WS(zL, T.m), *zU = WS(zU, T.m);
	const double *dL = WS(dL, T.m), *dU = WS(dU, T.m), *sc = WS(sc, T.m);
	double *y = WS(y, T.m), *Sig = WS(Sig, T.m), *w = WS(w, T.m), *dy = WS(dy, T.m), *vec = WS(vec, T.npad), *scal = WS(scal, 16);
	/* v: 0 dual_inf(max) 1 theta_inf(max) 2 compl max 3 compl min 4 sum|y| 5 sum z 6 viol(max) */
	double v[7] = {0, 0, 0, 1e300, 0, 0, 0};
	for (int i = threadIdx.x; i < T.n_free; i += blockDim.x) v[0] = fmax(v[0], fabs(jt_gather(T, Jv, y, i)));
	for (int i = threadIdx.x; i < T.m; i += blockDim.x) {
		const int fl = T.row_flags[i];
		v[4] += fabs(y[i]);
		if (fl & ROW_EQ) { v[1] = fmax(v[1], fabs(r[i])); v[6] = fmax(v[6], fabs(r[i]) / sc[i]); continue; }
		v[1] = fmax(v[1], fabs(r[i] - s[i]));
		v[0] = fmax(v[0], fabs(-y[i] - zL[i] + zU[i]));
		if (fl & ROW_HASL) { const double c = zL[i] * (s[i] - dL[i]); v[2] = fmax(v[2], c); v[3] = fmin(v[3], c); v[5] += zL[i]; v[6] = fmax(v[6], T.gl[i] - r[i] / sc[i]); }
		if (fl & ROW_HASU) { const double c = zU[i] * (dU[i] - s[i]); v[2] = fmax(v[2], c); v[3] = fmin(v[3], c); v[5] += zU[i]; v[6] = fmax(v[6], r[i] / sc[i] - T.gu[i]); }
	}
	const int ops[7] = {1, 1, 1, 2, 0, 0, 1};
	block_reduce<7>(v, ops, red);
	const double dual_inf = v[0], theta_inf = v[1], cmax = v[2], cmin = T.n_bounds > 0 ? v[3] : 0.0;
	const double s_d = fmax(100.0, (v[4] + v[5]) / (double)(T.m + T.n_bounds)) / 100.0;
	const double s_c = fmax(100.0, v[5] / (double)(T.n_bounds > 0 ? T.n_bounds : 1)) / 100.0;
	const double E0 = fmax(fmax(dual_inf / s_d, theta_inf), cmax / s_c);
	double mu = scal[SC_MU];
	/* f == 0 on this path (ref: parameters.cc:62-63): a feasible point is a KKT point with zero multipliers */
	const bool feas = opt.feas_exit && v[6] <= opt.constr_viol_tol && theta_inf <= opt.tol;
	const bool conv = feas || (E0 <= opt.tol && v[6] <= opt.constr_viol_tol && cmax <= opt.compl_inf_tol && dual_inf <= opt.dual_inf_tol);
	__syncthreads();
	if (threadIdx.x == 0) {
		scal[SC_DUAL] = dual_inf; scal[SC_THETA] = theta_inf; scal[SC_COMPL] = cmax; scal[SC_VIOL] = v[6]; scal[SC_E0] = E0;
		W.iters[pid] = it;
		if (conv) { W.status[pid] = QTOS_SOLVE_SUCCEEDED; atomicSub(W.n_running, 1); }
		else if (it >= opt.max_iter) { W.status[pid] = QTOS_MAX_ITER; atomicSub(W.n_running, 1); }
	}
	if (conv || it >= opt.max_iter) return;
	/* monotone barrier update: max_i |z_i s_i - mu| = max(cmax - mu, mu - cmin) */
	const double kappa_eps = 10.0, mu_min = fmin(opt.tol, opt.compl_inf_tol) / (kappa_eps + 1.0);
	for (;;) {
		const double cm = fmax(cmax - mu, mu - cmin);
		const double Emu = fmax(fmax(dual_inf / s_d, theta_inf), cm / s_c);
		if (Emu <= kappa_eps * mu && mu > mu_min) mu = fmax(mu_min, fmin(0.2 * mu, mu * sqrt(mu)));
		else break;
	}
	if (threadIdx.x == 0) scal[SC_MU] = mu;
	const double rho = 1.0 / opt.delta_c;
	for (int i = threadIdx.x; i < T.m; i += blockDim.x) {
		const int fl = T.row_flags[i];
		if (fl & ROW_EQ) { Sig[i] = rho; w[i] = y[i] + rho * r[i]; continue; }
		double sg = 0.0, rsm = -y[i];
		if (fl & ROW_HASL) { const double sl = s[i] - dL[i]; sg += zL[i] / sl; rsm -= mu / sl; }
		if (fl & ROW_HASU) { const double su = dU[i] - s[i]; sg += zU[i] / su; rsm += mu / su; }
		Sig[i] = sg; dy[i] = rsm;
		w[i] = y[i] + sg * (r[i] - s[i]) + rsm;
	}
	__syncthreads();
	for (int i = threadIdx.x; i < T.npad; i += blockDim.x) vec[i] = i < T.n_free ? -jt_gather(T, Jv, w, i) : 0.0;
}

/* ------------------------------------------------------------------ k_assemble */

__global__ void __launch_bounds__(QTOS_THREADS)
k_assemble(DevTables T, DevWork W, qtos_options opt)
{
	const int pid = blockIdx.x, chunk = blockIdx.y;
	if (W.status[pid] != QTOS_RUNNING) return;
	const double *Jv = WS(Jv, T.nJ), *Sig = WS(Sig, T.m);
	double *M = WS(M, T.nM);
	const int per = (T.nM / NB / NB + T.n_chunks - 1) / T.n_chunks * NB * NB;
	const int lo = chunk * per, hi = min(T.nM, lo + per);
	for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) M[i] = 0.0;
	__syncthreads();
	for (int t = T.asm_chunk[chunk] + threadIdx.x; t < T.asm_chunk[chunk + 1]; t += blockDim.x) {
		double acc = 0.0;
		for (int q = T.asm_ptr[t]; q < T.asm_ptr[t + 1]; ++q) {
			const uint32_t term = T.asm_terms[q];
			const Element &E = T.elems[term >> 16];
			const double *ca = Jv + E.valoff + ((term >> 8) & 255u) * E.nrows, *cb = Jv + E.valoff + (term & 255u) * E.nrows;
			const double *D = Sig + E.row0;
			for (int rr = 0; rr < E.nrows; ++rr) acc += D[rr] * ca[rr] * cb[rr];
		}
		M[T.asm_off[t]] = acc;
	}
	__syncthreads();
	for (int i = threadIdx.x; i < T.npad; i += blockDim.x) {
		const int off = T.diag_off[i];
		if (off >= lo && off < hi) M[off] = i < T.n_free ? M[off] + opt.sigma_w : 1.0;
	}
}

/* ------------------------------------------------------------------ k_factor */

#define OPLD 17          /* padded row stride of operand blocks in shared memory */

__global__ void __launch_bounds__(QTOS_THREADS)
k_factor(DevTables T, DevWork W, int max_w /* max blocks per block row */)
{
	const int pid = blockIdx.x;
	if (W.status[pid] != QTOS_RUNNING) return;
	extern __shared__ double sm[];
	double *rp = sm;                               /* [max_w][16][16] current block row */
	double *op = rp + (size_t)max_w * 256;         /* [max_w][16][OPLD] operand panel */
	double *tmp = op + (size_t)max_w * 16 * OPLD;  /* [16][OPLD] */
	double *inv = tmp + 16 * OPLD;                 /* [16][OPLD] inverse of a diagonal block */
	double *M = WS(M, T.nM), *Dinv = WS(Dinv, T.nb * 256);
	const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15;
	__shared__ int bad;
	if (tid == 0) bad = 0;
	for (int I = 0; I < T.nb; ++I) {
		const int fI = T.fb[I], wI = I - fI + 1;
		double *rowg = M + (size_t)T.blkptr[I] * 256;
		__syncthreads();
		for (int q = tid; q < wI * 256; q += QTOS_THREADS) rp[q] = rowg[q];
		for (int J = fI; J <= I; ++J) {
			const int K0 = max(fI, T.fb[J]), nK = J - K0;
			__syncthreads();
			if (J < I) {
				const double *src = M + (size_t)(T.blkptr[J] + K0 - T.fb[J]) * 256;
				for (int q = tid; q < nK * 256; q += QTOS_THREADS) op[(q >> 4) * OPLD + (q & 15)] = src[q];
				inv[ti * OPLD + tj] = Dinv[(size_t)J * 256 + tid];
			} else {
				for (int q = tid; q < nK * 256; q += QTOS_THREADS) op[(q >> 4) * OPLD + (q & 15)] = rp[(size_t)(K0 - fI) * 256 + q];
			}
			__syncthreads();
			double acc = rp[(size_t)(J - fI) * 256 + ti * 16 + tj];
			const double *a = rp + (size_t)(K0 - fI) * 256 + ti * 16;
			const double *b = op + tj * OPLD;
			for (int K = 0; K < nK; ++K) {
#pragma unroll
				for (int k = 0; k < 16; ++k) acc -= a[K * 256 + k] * b[K * 16 * OPLD + k];
			}
			tmp[ti * OPLD + tj] = acc;
			__syncthreads();
			if (J < I) {
				/* L[I,J] = S * inv(L[J,J])' */
				double v = 0.0;
				for (int k = 0; k <= tj; ++k) v += tmp[ti * OPLD + k] * inv[tj * OPLD + k];
				rp[(size_t)(J - fI) * 256 + ti * 16 + tj] = v;
			} else {
				/* Cholesky of the 16x16 diagonal block in tmp (lower), then its inverse */
				for (int j = 0; j < 16; ++j) {
					if (tid == 0) {
						double d = tmp[j * OPLD + j];
						if (!(d > 0.0)) { d = 1e-30; bad = 1; }
						tmp[j * OPLD + j] = sqrt(d);
					}
					__syncthreads();
					if (tj == j && ti > j) tmp[ti * OPLD + j] /= tmp[j * OPLD + j];
					__syncthreads();
					if (tj > j && ti >= tj) tmp[ti * OPLD + tj] -= tmp[ti * OPLD + j] * tmp[tj * OPLD + j];
					__syncthreads();
				}
				if (tid < 16) {
					/* column tid of inv(L): forward substitution of e_tid */
					const int c = tid;
					double xcol[16];
					for (int i = 0; i < 16; ++i) {
						double sacc = i == c ? 1.0 : 0.0;
						for (int k = c; k < i; ++k) sacc -= tmp[i * OPLD + k] * xcol[k];
						xcol[i] = i < c ? 0.0 : sacc / tmp[i * OPLD + i];
					}
					for (int i = 0; i < 16; ++i) inv[i * OPLD + c] = xcol[i];
				}
				__syncthreads();
				rp[(size_t)(J - fI) * 256 + ti * 16 + tj] = ti >= tj ? tmp[ti * OPLD + tj] : 0.0;
				Dinv[(size_t)I * 256 + tid] = inv[ti * OPLD + tj];
			}
		}
		__syncthreads();
		for (int q = tid; q < wI * 256; q += QTOS_THREADS) rowg[q] = rp[q];
	}
	__syncthreads();
	if (tid == 0 && bad) W.flags[pid] |= 1;
}

/* ------------------------------------------------------------------ k_step */

__global__ void __launch_bounds__(QTOS_THREADS)
k_step(DevTables T, DevWork W, const qtos_problem *probs, const DevHeightfield *hfs, int n_hf, qtos_options opt)
{
	const int pid = blockIdx.x;
	if (W.status[pid] != QTOS_RUNNING) return;
	extern __shared__ double sm[];
	double *b = sm;                       /* [npad] solution vector (permuted order) */
	double *part = b + T.npad;            /* [256] scratch */
	double *red = part + 256;             /* [8*32] */
	const double *M = WS(M, T.nM), *Dinv = WS(Dinv, T.nb * 256), *Jv = WS(Jv, T.nJ);
	const int tid = threadIdx.x;
	for (int i = tid; i < T.npad; i += blockDim.x) b[i] = W.vec[(size_t)pid * T.npad + i];
	__syncthreads();
	/* forward: L z = b */
	{
		const int k = tid & 15, ti = tid >> 4;
		for (int I = 0; I < T.nb; ++I) {
			const int fI = T.fb[I];
			const double *rowg = M + (size_t)T.blkptr[I] * 256;
			double acc = 0.0;
			for (int J = fI; J < I; ++J) acc += rowg[(size_t)(J - fI) * 256 + ti * 16 + k] * b[J * 16 + k];
			for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
			if (k == 0) part[ti] = b[I * 16 + ti] - acc;
			__syncthreads();
			/* z_I = inv(L_II) * part */
			if (tid < 16) {
				double v = 0.0;
				const double *iv = Dinv + (size_t)I * 256 + tid * 16;
				for (int q = 0; q <= tid; ++q) v += iv[q] * part[q];
				b[I * 16 + tid] = v;
			}
			__syncthreads();
		}
		/* backward: L' x = z */
		for (int I = T.nb - 1; I >= 0; --I) {
			const int fI = T.fb[I];
			const double *rowg = M + (size_t)T.blkptr[I] * 256;
			if (tid < 16) {
				double v = 0.0;
				const double *iv = Dinv + (size_t)I * 256;
				for (int q = tid; q < 16; ++q) v += iv[q * 16 + tid] * b[I * 16 + q];
				part[tid] = v;
			}
			__syncthreads();
			if (tid < 16) b[I * 16 + tid] = part[tid];
			/* b_J -= L[I,J]' x_I : thread per (J,k) */
			for (int c = tid; c < (I - fI) * 16; c += blockDim.x) {
				const int Jr = c >> 4, kk = c & 15;
				const double *blk = rowg + (size_t)Jr * 256 + kk;
				double acc = 0.0;
#pragma unroll
				for (int q = 0; q < 16; ++q) acc += blk[q * 16] * part[q];
				b[(fI + Jr) * 16 + kk] -= acc;
			}
			__syncthreads();
		}
	}
	/* step recovery */
	const double *r = WS(r, T.m), *Sig = WS(Sig, T.m), *dLb = WS(dL, T.m), *dUb = WS(dU, T.m);
	double *s = WS(s, T.m), *y = WS(y, T.m), *zL = WS(zL, T.m), *zU = WS(zU, T.m);
	double *ds = WS(ds, T.m), *dy = WS(dy, T.m), *dzL = WS(dzL, T.m), *dzU = WS(dzU, T.m);
	double *rt = WS(rt, T.m), *st = WS(st, T.m), *x = WS(x, T.n_all), *xt = WS(xt, T.n_all), *scal = WS(scal, 16);
	const double *sc = WS(sc, T.m);
	const double mu = scal[SC_MU], rho = 1.0 / opt.delta_c, tau = fmax(0.99, 1.0 - mu);
	/* v: 0 a_pr(min) 1 a_du(min) 2 theta1(sum) 3 gphi_d(sum) 4 quad(sum) 5 bar0(sum) */
	double v[6] = {1.0, 1.0, 0, 0, 0, 0};
	for (int i = tid; i < T.n_free; i += blockDim.x) v[4] += opt.sigma_w * b[i] * b[i];
	for (int i = tid; i < T.m; i += blockDim.x) {
		const int fl = T.row_flags[i];
		const Element &E = T.elems[T.row_elem[i]];
		const int rr = i - E.row0;
		const int16_t *cols = T.elem_cols + E.coloff;
		double jdx = 0.0;
		for (int a = 0; a < E.ncols; ++a) jdx += Jv[E.valoff + a * E.nrows + rr] * b[cols[a]];
		if (fl & ROW_EQ) { dy[i] = rho * (jdx + r[i]); v[2] += fabs(r[i]); continue; }
		const double rsm = dy[i];
		const double dsi = jdx + (r[i] - s[i]);
		ds[i] = dsi; dy[i] = Sig[i] * dsi + rsm;
		v[2] += fabs(r[i] - s[i]);
		v[4] += dsi * Sig[i] * dsi;
		if (fl & ROW_HASL) {
			const double sl = s[i] - dLb[i];
			const double dz = mu / sl - zL[i] - zL[i] / sl * dsi;
			dzL[i] = dz; v[3] += -mu / sl * dsi; v[5] -= mu * log(sl);
			if (dsi < 0) v[0] = fmin(v[0], -tau * sl / dsi);
			if (dz < 0) v[1] = fmin(v[1], -tau * zL[i] / dz);
		}
		if (fl & ROW_HASU) {
			const double su = dUb[i] - s[i];
			const double dz = mu / su - zU[i] + zU[i] / su * dsi;
			dzU[i] = dz; v[3] += mu / su * dsi; v[5] -= mu * log(su);
			if (dsi > 0) v[0] = fmin(v[0], tau * su / dsi);
			if (dz < 0) v[1] = fmin(v[1], -tau * zU[i] / dz);
		}
	}
	{ const int ops[6] = {2, 2, 0, 0, 0, 0}; block_reduce<6>(v, ops, red); }
	const double a_pr = v[0], a_du = v[1], theta1 = v[2], gphi_d = v[3], quad = v[4], bar0 = v[5];
	double nu = scal[SC_NU];
	if (theta1 > 1e-14) {
		const double nu_trial = (gphi_d + 0.5 * quad) / (0.7 * theta1);
		if (nu < nu_trial) nu = nu_trial + 1.0;
	}
	const double phi0 = bar0 + nu * theta1, Dphi = gphi_d - nu * theta1;
	const int hid = probs[pid].hf_id >= 0 && probs[pid].hf_id < n_hf ? probs[pid].hf_id : 0;
	const DevHeightfield hf = hfs[hid];
	double alpha = a_pr;
	int ls = 0;
	for (;;) {
		for (int i = tid; i < T.n_all; i += blockDim.x) {
			const int p = T.perm_of_var[i];
			xt[i] = p >= 0 ? x[i] + alpha * b[p] : x[i];
		}
		__syncthreads();
		eval_g_block(T, hf, xt, rt);
		__syncthreads();
		double u[2] = {0, 0};     /* theta, barrier */
		for (int i = tid; i < T.m; i += blockDim.x) {
			const int fl = T.row_flags[i];
			const double g = rt[i];
			if (fl & ROW_EQ) { const double c = sc[i] * (g - T.gl[i]); rt[i] = c; u[0] += fabs(c); continue; }
			const double d = sc[i] * g, sn = s[i] + alpha * ds[i];
			rt[i] = d; st[i] = sn;
			u[0] += fabs(d - sn);
			if (fl & ROW_HASL) u[1] -= mu * log(sn - dLb[i]);
			if (fl & ROW_HASU) u[1] -= mu * log(dUb[i] - sn);
		}
		{ const int ops[2] = {0, 0}; block_reduce<2>(u, ops, red); }
		ls++;
		if (u[1] + nu * u[0] <= phi0 + 1e-4 * alpha * Dphi || ls >= 12) break;
		alpha *= 0.5;
	}
	/* accept */
	for (int i = tid; i < T.n_all; i += blockDim.x) x[i] = xt[i];
	for (int i = tid; i < T.m; i += blockDim.x) {
		const int fl = T.row_flags[i];
		WS(r, T.m)[i] = rt[i];
		y[i] += alpha * dy[i];
		if (fl & ROW_EQ) continue;
		s[i] = st[i];
		if (fl & ROW_HASL) { const double sl = st[i] - dLb[i]; zL[i] = fmin(fmax(zL[i] + a_du * dzL[i], mu / (1e10 * sl)), 1e10 * mu / sl); }
		if (fl & ROW_HASU) { const double su = dUb[i] - st[i]; zU[i] = fmin(fmax(zU[i] + a_du * dzU[i], mu / (1e10 * su)), 1e10 * mu / su); }
	}
	if (tid == 0) {
		scal[SC_NU] = nu;
		const double nfail = ls >= 12 ? scal[SC_NFAIL] + 1.0 : 0.0;
		scal[SC_NFAIL] = nfail;
		if (nfail >= 3.0) { W.status[pid] = QTOS_STEP_FAILED; atomicSub(W.n_running, 1); }   /* line search stalled */
	}
}

/* ------------------------------------------------------------------ results, sampler, queries */

__global__ void k_results(DevTables T, DevWork W, qtos_result *res, double *x_out, int n)
{
	const int pid = blockIdx.x;
	if (pid >= n) return;
	if (threadIdx.x == 0 && res) {
		const double *scal = WS(scal, 16);
		qtos_result R;
		R.status = W.status[pid]; R.iters = W.iters[pid];
		R.constr_viol = scal[SC_VIOL]; R.dual_inf = scal[SC_DUAL]; R.compl_inf = scal[SC_COMPL]; R.nlp_error = scal[SC_E0]; R.mu = scal[SC_MU];
		R.cost = 0.0;
		res[pid] = R;
	}
	if (x_out) for (int i = threadIdx.x; i < T.n_all; i += blockDim.x) x_out[(size_t)pid * T.n_all + i] = WS(x, T.n_all)[i];
}

/* post-hoc plan cost for best-plan selection: sum over optimised nodes of f_z^2 (force) and
 * v_x^2 + v_y^2 (foot motion) -- the reference's optional NodeCost terms
 * (ref: nlp_formulation.cc:354-376, node_cost.cc:53-63), evaluated, not optimised */
__global__ void k_cost(DevTables T, const double *x_all, qtos_result *res, int n)
{
	const int pid = blockIdx.x;
	if (pid >= n) return;
	__shared__ double red[32];
	const double *x = x_all + (size_t)pid * T.n_all;
	double acc = 0.0;
	for (int v = T.var_off[2] + threadIdx.x; v < T.n_all; v += blockDim.x) {
		const int s = T.x0_spline[v], deriv = T.x0_deriv[v], dim = T.x0_dim[v];
		if (s >= 6 && deriv == 0 && dim == 2) acc += x[v] * x[v];
		if (s >= 2 && s < 6 && deriv == 1 && dim < 2) acc += x[v] * x[v];
	}
	double u[1] = {acc}; const int ops[1] = {0};
	block_reduce<1>(u, ops, red);
	if (threadIdx.x == 0) res[pid].cost = u[0];
}

__global__ void k_csv(DevTables T, const double *x_all, const qtos_problem *probs, double *rows, int n)
{
	const int pid = blockIdx.y;
	const int row = blockIdx.x * blockDim.x + threadIdx.x;
	if (row >= T.csv_rows || pid >= n) return;
	const double *x = x_all + (size_t)pid * T.n_all;
	double *c = rows + ((size_t)pid * T.csv_rows + row) * QTOS_CSV_COLS;
	c[0] = T.csv_t[row] + probs[pid].t_start;
	for (int s = 0; s < 10; ++s) {
		const int id = T.csv_id[(size_t)row * 10 + s];
		const double t = T.csv_tl[(size_t)row * 10 + s], D = T.dur[s * T.dur_ld + id];
		for (int k = 0; k < 3; ++k) {
			double p0, v0, p1, v1;
			if (s < 2) { const double *xs = x + T.var_off[s] + id * 6; p0 = xs[k]; v0 = xs[3 + k]; p1 = xs[6 + k]; v1 = xs[9 + k]; }
			else {
				const int16_t *nv = T.node_var + ((size_t)s * T.max_nodes + id) * 6;
				p0 = node_val(x, nv[k]); v0 = node_val(x, nv[3 + k]); p1 = node_val(x, nv[6 + k]); v1 = node_val(x, nv[9 + k]);
			}
			/* cubic Hermite coefficients (ref: polynomial.cc:89-104) */
			const double C = -(3 * (p0 - p1) + D * (2 * v0 + v1)) / (D * D);
			const double Dd = (2 * (p0 - p1) + D * (v0 + v1)) / (D * D * D);
			const double pos = p0 + t * v0 + t * t * C + t * t * t * Dd;
			if (s == 0) { c[1 + k] = pos; c[19 + k] = v0 + 2 * t * C + 3 * t * t * Dd; }
			else if (s == 1) { c[4 + k] = pos; c[22 + k] = v0 + 2 * t * C + 3 * t * t * Dd; }
			else if (s < 6) c[7 + (s - 2) * 3 + k] = pos;
			else c[25 + (s - 6) * 3 + k] = pos;
		}
	}
}

__global__ void k_height(DevHeightfield hf, const double *xy, int n, double *h_out, long long *idx_out)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	if (h_out) h_out[i] = qtos_height(hf, xy[2 * i], xy[2 * i + 1]);
	if (idx_out) { long long c[4]; qtos_height_cell(hf, xy[2 * i], xy[2 * i + 1], c); for (int q = 0; q < 4; ++q) idx_out[4 * i + q] = c[q]; }
}

/* test/parity entry: g(x) and dense Jacobian at caller-provided x (fixed entries of x are overwritten) */
__global__ void __launch_bounds__(QTOS_THREADS)
k_eval_dense(DevTables T, DevWork W, const qtos_problem *probs, const DevHeightfield *hfs, int n_hf,
             const double *x_in, double *g_out, double *jac_out)
{
	const int pid = blockIdx.x;
	double *x = WS(x, T.n_all), *sc = WS(sc, T.m), *Jv = WS(Jv, T.nJ), *P = WS(P, 32);
	const int hid = probs[pid].hf_id >= 0 && probs[pid].hf_id < n_hf ? probs[pid].hf_id : 0;
	const DevHeightfield hf = hfs[hid];
	if (x_in) for (int v = threadIdx.x; v < T.n_all; v += blockDim.x) x[v] = T.fix_src[v] >= 0 ? P[T.fix_src[v]] : x_in[(size_t)pid * T.n_all + v];
	for (int i = threadIdx.x; i < T.m; i += blockDim.x) sc[i] = 1.0;
	__syncthreads();
	if (g_out) eval_g_block(T, hf, x, g_out + (size_t)pid * T.m);
	if (!jac_out) return;
	for (int i = threadIdx.x; i < T.nJ; i += blockDim.x) Jv[i] = T.Jconst[i];
	__syncthreads();
	eval_jac_block(T, x, sc, Jv);
	__syncthreads();
	double *J = jac_out + (size_t)pid * T.m * T.n_all;
	for (int i = threadIdx.x; i < T.m * T.n_all; i += blockDim.x) J[i] = 0.0;
	__syncthreads();
	for (int e = threadIdx.x; e < T.n_elem; e += blockDim.x) {
		const Element &E = T.elems[e];
		for (int a = 0; a < E.ncols; ++a) {
			const int var = T.var_of_perm[T.elem_cols[E.coloff + a]];
			for (int rr = 0; rr < E.nrows; ++rr) J[(size_t)(E.row0 + rr) * T.n_all + var] = Jv[E.valoff + a * E.nrows + rr];
		}
	}
}

/* FP64 FMA throughput probe (roofline denominator measured on the box; MEASURED_PEAKS.json has no FP64 figure) */
__global__ void k_fp64_peak(double *out, int iters)
{
	double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
	const double b = 1.0000001, c = 1e-9;
	for (int i = 0; i < iters; ++i) {
		a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
		a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

#include "qtos_capi.inc"
