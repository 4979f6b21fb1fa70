/*
 * qtos_ipopt.cuh -- kernels of QTOS_ALG_IPOPT: the interior-point algorithm the reference runs (Ipopt 3.11.9 as configured
 * by ifopt: limited-memory BFGS Hessian with history 6, adaptive quality-function barrier update, filter line search;
 * ref: solver/towr/src/main.cpp:444-463, logs/towr_log.out:37-64).  Ipopt's source is not under /root/reference; the
 * algorithm is the published one (Waechter & Biegler, Math. Prog. 106, 2006; Nocedal, Waechter & Waltz, SIAM J. Optim. 19,
 * 2009), restated for one thread block per problem.  The test oracle of exactly this form is oracle/towr_ipopt.c, which is
 * pinned to the reference's logged iteration tables and plans.
 *
 * One batch iteration = k_jac_dyn | k_jac_rom -> kip_prepare -> k_asm -> k_factor<.,1> -> kip_solve -> kip_step:
 *   kip_prepare  J'y, limited-memory update (pairs s = x+ - x, y = (J+ - J)'lambda+), error measures and termination,
 *                barrier-update bookkeeping (free / fixed mode), Sigma, the right-hand sides of the affine-scaling and
 *                centering directions and the limited-memory columns Bl = [sigma S, Y] as 16 right-hand sides (W.RB)
 *   k_asm        M = sigma I + Jd' Sigma Jd + rho Jc' Jc
 *   k_factor     M = L L' and P = L^-1 RB (one more block row of the factorization)
 *   kip_solve    Woodbury term between the triangular solves: Mf^-1 v = L'^-1 (p + Q (Mid - Q'Q)^-1 Q'p), Q = L^-1 Bl;
 *                n_refine multiplier-method passes on the equality block; both directions at once (two right-hand sides
 *                through every sweep over L, the rows of L streamed by bulk asynchronous copies)
 *   kip_step     quality-function barrier oracle (golden section over sigma), search direction aff + sigma cen,
 *                fraction to the boundary, filter line search with in-kernel g(x), iterate and multiplier update
 */
#ifndef QTOS_IPOPT_CUH_
#define QTOS_IPOPT_CUH_

#define IP_EPS 2.220446049250313e-16

/* ------------------------------------------------------------------ kip_prepare */

/* (J' wa)_i and (J' wb)_i together for permuted variable i over a term table (see jt_gather): the column is read once */
__device__ __forceinline__ void jt_gather2(const int *ptr, const uint2_t *tab, const double *Jv, const double *wa, const double *wb, int i, double &ga, double &gb)
{
	const int g = i >> 5, base = ptr[g], ns = (ptr[g + 1] - base) >> 5;
	const uint2 *tk = reinterpret_cast<const uint2 *>(tab) + base + (i & 31);
	double a0 = 0.0, a1 = 0.0;
	for (int s = 0; s < ns; s += 2) {
		uint2 d[2];
		double c[2][6], u[2][6], w[2][6];
#pragma unroll
		for (int q = 0; q < 2; ++q) d[q] = s + q < ns ? __ldg(tk + 32 * (s + q)) : make_uint2(0u, 0u);
#pragma unroll
		for (int q = 0; q < 2; ++q) {
			const int nr = d[q].x >> 20;
			const double *col = Jv + (d[q].x & 0xfffffu), *pa = wa + d[q].y, *pb = wb + d[q].y;
#pragma unroll
			for (int rr = 0; rr < 6; ++rr) { c[q][rr] = rr < nr ? col[rr] : 0.0; u[q][rr] = rr < nr ? pa[rr] : 0.0; w[q][rr] = rr < nr ? pb[rr] : 0.0; }
		}
#pragma unroll
		for (int q = 0; q < 2; ++q)
#pragma unroll
			for (int rr = 0; rr < 6; ++rr) { a0 += c[q][rr] * u[q][rr]; a1 += c[q][rr] * w[q][rr]; }
	}
	ga = a0; gb = a1;
}


__global__ void __launch_bounds__(QTOS_THREADS, PREP_MINB)
kip_prepare(DevTables T, DevWork W, qtos_options opt)
{
	const int pid = blockIdx.x;
	if (W.status[pid] != QTOS_RUNNING) return;
	__shared__ double red[15 * 32];
	__shared__ double sdots[2 * IP_LM * IP_LM];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const double *Jv = WS(Jv, T.nJ), *r = WS(r, T.m), *s = WS(s, T.m), *vL = WS(zL, T.m), *vU = WS(zU, T.m);
	const double *dL = WS(dL, T.m), *dU = WS(dU, T.m), *sc = WS(sc, T.m), *y = WS(y, T.m), *x = WS(x, T.n_all);
	double *Sig = WS(Sig, T.m), *wA = WS(wA, T.m), *wC = WS(wC, T.m), *scal = WS(scal, 16), *ip = WS(ipst, IP_N);
	double *glx = WS(glx, T.npad), *lastx = WS(lastx, T.npad), *gJold = WS(gJold, T.npad);
	double *tS = WS(adx, T.npad), *tY = WS(cdx, T.npad);          /* candidate pair (the direction buffers are free here) */
	double *lmS = WS(lmS, IP_LM * T.npad), *lmY = WS(lmY, IP_LM * T.npad), *RB = WS(RB, T.npad * IP_NRHS);
	const int hist = opt.lm_history < 1 ? 1 : (opt.lm_history > IP_LM ? IP_LM : opt.lm_history);
	/* state of the previous iteration (every thread reads it before thread 0 rewrites it at the end) */
	double mu = ip[IP_MU], tau = ip[IP_TAU], sigma_w = ip[IP_SIGMA_W], mu_max = ip[IP_MU_MAX], amu_thmin = ip[IP_AMU_THMIN];
	int free_mode = (int)ip[IP_FREE], n_pairs = (int)ip[IP_NPAIRS], skipped = (int)ip[IP_SKIPPED], head = (int)ip[IP_HEAD];
	int nfilter = (int)ip[IP_NFILTER];
	/* a retry repeats the iteration at the same point with more regularisation on the diagonal (see kip_step): the
	 * limited-memory update, the termination test and the barrier bookkeeping of this iteration are already done */
	const int retry = (int)ip[IP_RETRY], it = (int)ip[IP_ITER];
	const int have_last = retry ? 0 : (int)ip[IP_HAVE_LAST];
	const double delta_w = retry ? ip[IP_DELTA_W] : 0.0;
	/* v: 0 dual_inf (max) 1 primal_inf (max) 2 compl (max) 3 sum|y| 4 sum z 5 viol (max) 6 theta (sum) 7 sum s z 8 sum 0*r
	 *    9 |grad L|^2 10 |c|^2 11 max |s z - mu| 12 s'y 13 s's 14 y'y */
	double v[15] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
	/* optional cost terms (qtos_shape.cost_*): f = df sum c_v x_v^2, grad_x L = grad f + J' lambda; f == 0 otherwise */
	const double df = T.cost_c ? ip[IP_OBJ_SCALE] : 1.0;
	double fval = 0.0;
	if (T.cost_c) {
		double u[1] = {0.0};
		for (int i = tid; i < T.n_all; i += blockDim.x) u[0] += T.cost_c[i] * x[i] * x[i];
		const int ops[1] = {0};
		block_reduce<1>(u, ops, red);
		fval = df * u[0];
	}
	for (int i = tid; i < T.npad; i += blockDim.x) {
		double g = jt_gather(T, Jv, y, i);
		const int var = T.var_of_perm[i];
		const double xi = var >= 0 ? x[var] : 0.0;
		if (T.cost_c && var >= 0) g += 2.0 * df * T.cost_c[var] * xi;
		glx[i] = g;
		v[0] = fmax(v[0], fabs(g)); v[9] += g * g;
		if (have_last) {
			const double sn = xi - lastx[i], yn = g - gJold[i];
			tS[i] = sn; tY[i] = yn;
			v[12] += sn * yn; v[13] += sn * sn; v[14] += yn * yn;
		}
		lastx[i] = xi;
	}
	for (int i = tid; i < T.m; i += blockDim.x) {
		const int fl = T.row_flags[i];
		v[3] += fabs(y[i]);
		v[8] += 0.0 * r[i];
		if (fl & ROW_EQ) { const double a = fabs(r[i]); v[1] = fmax(v[1], a); v[6] += a; v[5] = fmax(v[5], a / sc[i]); v[10] += r[i] * r[i]; continue; }
		const double c = r[i] - s[i], gs = -y[i] - vL[i] + vU[i];
		v[1] = fmax(v[1], fabs(c)); v[6] += fabs(c); v[10] += c * c;
		v[0] = fmax(v[0], fabs(gs)); v[9] += gs * gs;
		if (fl & ROW_HASL) { const double p = (s[i] - dL[i]) * vL[i]; v[2] = fmax(v[2], p); v[7] += p; v[4] += vL[i]; v[11] = fmax(v[11], fabs(p - mu)); v[5] = fmax(v[5], T.gl[i] - r[i] / sc[i]); }
		if (fl & ROW_HASU) { const double p = (dU[i] - s[i]) * vU[i]; v[2] = fmax(v[2], p); v[7] += p; v[4] += vU[i]; v[11] = fmax(v[11], fabs(p - mu)); v[5] = fmax(v[5], r[i] / sc[i] - T.gu[i]); }
	}
	{ const int ops[15] = {1, 1, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 0}; block_reduce<15>(v, ops, red); }
	const double dual_inf = v[0], primal_inf = v[1], compl_ = v[2], viol = v[5], theta = v[6];
	const int nbnd = T.n_bounds > 0 ? T.n_bounds : 1;
	const double s_d = fmax(100.0, (v[3] + v[4]) / (double)(T.m + T.n_bounds)) / 100.0;
	const double s_c = fmax(100.0, v[4] / (double)nbnd) / 100.0;
	const double nlp_error = fmax(fmax(dual_inf / s_d, primal_inf), compl_ / s_c);
	const double avrg_compl = v[7] / (double)nbnd;
	const bool invalid = !(v[8] == 0.0) || !(nlp_error == nlp_error) || !(theta == theta);
	/* the absolute tolerances apply to the unscaled problem: dual infeasibility and complementarity carry the objective's scale */
	const bool conv = !invalid && nlp_error <= opt.tol && dual_inf / df <= opt.dual_inf_tol && viol <= opt.constr_viol_tol && compl_ / df <= opt.compl_inf_tol;
	const bool out_of_time = it >= cpu_budget(opt);
	const bool stop = !retry && (invalid || conv || it >= opt.max_iter || out_of_time);
	if (tid == 0 && !retry) {
		scal[SC_DUAL] = dual_inf; scal[SC_THETA] = primal_inf; scal[SC_COMPL] = compl_; scal[SC_VIOL] = viol; scal[SC_E0] = nlp_error; scal[SC_MU] = mu;
		W.iters[pid] = (int)ip[IP_ITER_BASE] + it;
		if (it < QTOS_TRACE_ITERS) {
			double *tr = WS(trace, QTOS_TRACE_ITERS * QTOS_TRACE_COLS) + it * QTOS_TRACE_COLS;
			tr[0] = viol; tr[1] = dual_inf; tr[2] = mu; tr[3] = ip[IP_DNORM]; tr[4] = ip[IP_ALPHA_DU]; tr[5] = ip[IP_ALPHA_PR]; tr[6] = ip[IP_LS]; tr[7] = ip[IP_TAG];
		}
		if (stop) {
			if (invalid) { const double qnan = v[8] - v[8] + (nlp_error - nlp_error); scal[SC_VIOL] = scal[SC_E0] = qnan; }
			W.status[pid] = invalid ? QTOS_INVALID_NUMBER : (conv ? QTOS_SOLVE_SUCCEEDED : (out_of_time ? QTOS_MAX_CPUTIME : QTOS_MAX_ITER));
			atomicSub(W.n_running, 1);
		}
	}
	if (stop) return;

	/* ---- limited-memory update (LimMemQuasiNewtonUpdater::UpdateHessian) */
	int new_slot = -1;
	if (have_last) {
		const double sTy = v[12], sTs = v[13], yTy = v[14];
		const bool skipping = sTy <= sqrt(IP_EPS) * sqrt(sTs) * sqrt(yTy);
		if (skipping) { if (++skipped >= 2) { n_pairs = 0; head = 0; sigma_w = 1.0; skipped = 0; } }
		else {
			skipped = 0;
			if (n_pairs == hist) { head = (head + 1) % hist; n_pairs--; }
			new_slot = (head + n_pairs) % hist;
			n_pairs++;
			sigma_w = fmin(fmax(sTy / sTs, 1e-8), 1e8);
		}
	}
	if (new_slot >= 0) {
		for (int i = tid; i < T.npad; i += blockDim.x) { lmS[(size_t)new_slot * T.npad + i] = tS[i]; lmY[(size_t)new_slot * T.npad + i] = tY[i]; }
	}
	const double sigma_b = fmax(sigma_w, ip[IP_SIGMA_MIN]);   /* limited_memory_init_val_min above Ipopt's 1e-8 (second attempts) */
	const double sigma_f = sigma_b + delta_w;         /* diagonal of M; the limited-memory columns keep sigma_b (W + delta_w I) */

	/* ---- barrier parameter, part one (AdaptiveMuUpdate::UpdateBarrierParameter): mode switches and the fixed-mode update;
	 *      the free-mode oracle needs the two directions and runs in kip_step */
	const double mu_min = fmin(1e-11, 0.5 * fmin(opt.tol, opt.compl_inf_tol));
	if (mu_max < 0.0) mu_max = 1e3 * avrg_compl;
	/* AdaptiveMuUpdate's own filter of (f, theta) pairs: a point passes when, against every entry, it is no larger in at least one
	 * coordinate.  With f == 0 the entries (-margin, theta_k - margin) collapse to their smallest theta (amu_thmin). */
	bool acceptable = theta <= amu_thmin;
	const int n_amu = T.cost_c ? (int)ip[IP_NAMU] : 0;
	if (T.cost_c) {
		acceptable = true;
		for (int k = 0; k < n_amu; ++k) if (!(fval <= ip[IP_AMUF + k] || theta <= ip[IP_AMUT + k])) { acceptable = false; break; }
	}
	if (retry) { /* done in the first attempt */ }
	else if (!free_mode) {
		if (acceptable) free_mode = 1;
		else {
			const double berr = fmax(fmax(dual_inf / s_d, primal_inf), v[11] / s_c);
			if (berr <= 10.0 * mu) {
				double nm = fmin(0.2 * mu, pow(mu, 1.5));
				nm = fmax(nm, fmin(opt.compl_inf_tol, opt.tol) / 11.0);
				mu = nm; tau = fmax(0.99, 1.0 - mu); nfilter = 0;
			}
		}
	} else if (!acceptable) {
		free_mode = 0;
		mu = fmin(fmax(0.8 * avrg_compl, mu_min), mu_max);
		tau = fmax(0.99, 1.0 - mu); nfilter = 0;
	}
	const double amu_mg = 1e-5 * fmin(1.0, theta);
	const bool remember = !retry && free_mode && acceptable && amu_mg > 0.0;     /* RememberCurrentPointAsAccepted */
	if (remember && theta - amu_mg < amu_thmin) amu_thmin = theta - amu_mg;

	/* ---- Sigma and the first-pass row weights of the two right-hand sides (PDFullSpaceSolver in condensed form) */
	const double rho = 1.0 / opt.delta_c;
	for (int i = tid; i < T.m; i += blockDim.x) {
		const int fl = T.row_flags[i];
		if (fl & ROW_EQ) { Sig[i] = rho; wA[i] = rho * -r[i]; wC[i] = 0.0; continue; }
		double sg = 0.0, augA = -(-y[i] - vL[i] + vU[i]), augC = 0.0;
		if (fl & ROW_HASL) { const double sl = s[i] - dL[i]; sg += vL[i] / sl; augA += (-sl * vL[i]) / sl; augC += avrg_compl / sl; }
		if (fl & ROW_HASU) { const double su = dU[i] - s[i]; sg += vU[i] / su; augA -= (-su * vU[i]) / su; augC -= avrg_compl / su; }
		Sig[i] = sg;
		wA[i] = sg * -(r[i] - s[i]) + augA;
		wC[i] = augC;
	}
	__syncthreads();
	/* ---- right-hand-side block: columns 0..5 sigma S, 6..11 Y (chronological, unused ones zero), 12 affine, 13 centering */
	for (int i = tid; i < T.npad; i += blockDim.x) {
		double ga, vc;
		jt_gather2(T.jg_ptr, T.jg, Jv, wA, wC, i, ga, vc);          /* both right-hand sides from one pass over the columns */
		const double va = -glx[i] + ga;
		double *o = RB + (size_t)(i >> 4) * 256 + (i & 15);
#pragma unroll
		for (int a = 0; a < IP_LM; ++a) {
			const int sl = (head + a) % hist;
			o[a * 16] = a < n_pairs ? sigma_b * lmS[(size_t)sl * T.npad + i] : 0.0;
			o[(IP_LM + a) * 16] = a < n_pairs ? lmY[(size_t)sl * T.npad + i] : 0.0;
		}
		o[12 * 16] = i < T.n_free ? va : 0.0; o[13 * 16] = i < T.n_free ? vc : 0.0; o[14 * 16] = 0.0; o[15 * 16] = 0.0;
	}
	/* ---- middle matrix of the compact representation: [[sigma S'S, L], [L', -D]], L = strictly lower part of S'Y */
	for (int q = warp; q < n_pairs * n_pairs; q += blockDim.x >> 5) {
		const int a = q / n_pairs, b = q - a * n_pairs;
		const double *Sa = lmS + (size_t)((head + a) % hist) * T.npad, *Sb = lmS + (size_t)((head + b) % hist) * T.npad, *Yb = lmY + (size_t)((head + b) % hist) * T.npad;
		double ss = 0.0, sy = 0.0;
		for (int i = lane; i < T.npad; i += 32) { const double sa = Sa[i]; ss += sa * Sb[i]; sy += sa * Yb[i]; }
		for (int o = 16; o > 0; o >>= 1) { ss += __shfl_xor_sync(0xffffffffu, ss, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); }
		if (lane == 0) { sdots[a * IP_LM + b] = ss; sdots[IP_LM * IP_LM + a * IP_LM + b] = sy; }
	}
	__syncthreads();
	if (tid < 4 * IP_LM * IP_LM) {
		const int q2 = 2 * IP_LM, ra = tid / q2, cb = tid - ra * q2;
		const int a = ra % IP_LM, b = cb % IP_LM;
		double val = 0.0;
		if (a < n_pairs && b < n_pairs) {
			if (ra < IP_LM && cb < IP_LM) val = sigma_b * sdots[a * IP_LM + b];
			else if (ra < IP_LM) val = a > b ? sdots[IP_LM * IP_LM + a * IP_LM + b] : 0.0;              /* L[a][b] */
			else if (cb < IP_LM) val = b > a ? sdots[IP_LM * IP_LM + b * IP_LM + a] : 0.0;              /* L'[a][b] = L[b][a] */
			else val = a == b ? -sdots[IP_LM * IP_LM + a * IP_LM + a] : 0.0;
		}
		ip[IP_MID + tid] = val;
	}
	if (tid == 0) {
		ip[IP_MU] = mu; ip[IP_TAU] = tau; ip[IP_FREE] = free_mode; ip[IP_MU_MAX] = mu_max; ip[IP_AMU_THMIN] = amu_thmin;
		ip[IP_SIGMA_W] = sigma_w; ip[IP_SIGMA_F] = sigma_f; ip[IP_NPAIRS] = n_pairs; ip[IP_SKIPPED] = skipped; ip[IP_HEAD] = head;
		ip[IP_HAVE_LAST] = 1.0; ip[IP_NFILTER] = nfilter; ip[IP_AVRG] = avrg_compl; ip[IP_ERR] = nlp_error; ip[IP_THETA] = theta;
		ip[IP_GL2] = v[9]; ip[IP_PR2] = v[10];
		if (T.cost_c) {
			ip[IP_FVAL] = fval;
			if (remember) {
				const double ef = fval - amu_mg, eth = theta - amu_mg;
				int w = 0;                                 /* entries the new one dominates leave (Filter::AddEntry) */
				for (int k = 0; k < n_amu; ++k) if (!(ef <= ip[IP_AMUF + k] && eth <= ip[IP_AMUT + k])) { ip[IP_AMUF + w] = ip[IP_AMUF + k]; ip[IP_AMUT + w] = ip[IP_AMUT + k]; w++; }
				if (w == IP_FILTER_MAX) { for (int k = 1; k < w; ++k) { ip[IP_AMUF + k - 1] = ip[IP_AMUF + k]; ip[IP_AMUT + k - 1] = ip[IP_AMUT + k]; } w--; }
				ip[IP_AMUF + w] = ef; ip[IP_AMUT + w] = eth; ip[IP_NAMU] = w + 1;
			}
		}
		W.flags[pid] = 0;                              /* k_factor reports a non-positive pivot here */
		W.active[atomicAdd(W.n_active, 1)] = pid;
	}
}

/* ------------------------------------------------------------------ kip_solve */

#define KS_T 128
/* resident CTAs per SM the kernel is compiled for: the most of 6 / 5 / 4 for which the ring of L blocks the CTAs leave room for still holds
 * a block row + 1.  S2: kip_solve<6> (80 registers, no spills, 11-block ring) -- 42.0 ms per bench step with four CTAs and a 20-block ring,
 * 39.5 with five and 15 blocks, 38.0 with six; wide shapes (S5: a ring of at least 17 blocks) get four CTAs whatever the bound and keep kip_solve<4>'s 121 registers
 * (S5: 62.6 ms, 63.5 under the five-CTA register budget) */
#define KS_MINB_MAX 6

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
	asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra WAIT_%=;\n}\n"
	             :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
/* 1-D bulk asynchronous copy global -> shared (TMA unit, no tensor map), completion counted on the mbarrier */
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

/* The rows of L stream through a shared-memory RING of 2 KB blocks, several rows ahead of the row being solved: the
 * five (2 n_refine + 1) sweeps of one kernel form ONE fetch sequence (backward, forward, backward, ...), thread 0 keeps
 * issuing bulk copies while ring space and one of KS_NBAR mbarriers are free, so the prefetch also runs across the
 * J-products between two sweeps.  A row = inv(L_II) (from Dinv) + its off-diagonal blocks fI..I-1 (contiguous in M),
 * all fragment-major (see frag_off); a row may wrap around the end of the ring. */
#define KS_NBAR 8
struct SweepCtx {
	const int *fb, *blkptr;
	int nb, npad;
	const double *M, *Dinv;
	double *z;              /* [2][npad] */
	double *ring;           /* [rc][256] */
	double *part;           /* [4][16][2] */
	double *xi;             /* [2][16] */
	uint64_t *mbar;         /* [KS_NBAR] */
	int rc, total;          /* ring capacity in blocks; rows of the whole fetch sequence */
	int p_seq, p_pos, p_free;   /* producer (thread 0) */
	int c_seq, c_pos;           /* consumer (all threads alike) */
};

__device__ __forceinline__ int sweep_row_of(const SweepCtx &C, int seq)
{
	const int sw = seq / C.nb, k = seq - sw * C.nb;
	return (sw & 1) ? k : C.nb - 1 - k;               /* even sweeps run backward */
}

/* thread 0: issue the fetches of as many rows as fit */
__device__ __forceinline__ void sweep_pump(SweepCtx &C)
{
	while (C.p_seq < C.total && C.p_seq - C.c_seq < KS_NBAR) {
		const int row = sweep_row_of(C, C.p_seq), nblk = row - C.fb[row] + 1;
		if (nblk > C.p_free) break;
		uint64_t *bar = &C.mbar[C.p_seq % KS_NBAR];
		mbar_expect_tx(bar, (unsigned)(nblk * 2048));
		bulk_g2s(C.ring + (size_t)C.p_pos * 256, C.Dinv + (size_t)row * 256, 2048u, bar);
		const int pos = C.p_pos + 1 == C.rc ? 0 : C.p_pos + 1, n = nblk - 1;
		const double *src = C.M + (size_t)C.blkptr[row] * 256;
		const int first = n < C.rc - pos ? n : C.rc - pos;
		if (first) bulk_g2s(C.ring + (size_t)pos * 256, src, (unsigned)(first * 2048), bar);
		if (n - first) bulk_g2s(C.ring, src + (size_t)first * 256, (unsigned)((n - first) * 2048), bar);
		C.p_pos = (C.p_pos + nblk) % C.rc; C.p_free -= nblk; C.p_seq++;
	}
}

/* all threads, after the barrier that ends a row: its blocks are free again */
__device__ __forceinline__ void sweep_release(SweepCtx &C, int nblk)
{
	C.c_pos = (C.c_pos + nblk) % C.rc; C.c_seq++;
	if (threadIdx.x == 0) { C.p_free += nblk; sweep_pump(C); }
}

__device__ __forceinline__ const double *sweep_blk(const SweepCtx &C, int j)      /* block j of the current row (0 = inverse) */
{
	int p = C.c_pos + j; if (p >= C.rc) p -= C.rc;
	return C.ring + (size_t)p * 256;
}

/* L' x = z for the two right-hand sides in C.z, in place; all KS_T threads */
__device__ __forceinline__ void sweep_backward(SweepCtx &C)
{
	const int tid = threadIdx.x, npad = C.npad;
	for (int I = C.nb - 1; I >= 0; --I) {
		mbar_wait(&C.mbar[C.c_seq % KS_NBAR], (unsigned)((C.c_seq / KS_NBAR) & 1));
		const int fI = C.fb[I];
		{
			/* x_I = inv(L_II)' z_I for both right-hand sides: four lanes per entry */
			const double *inv = sweep_blk(C, 0);
			const int k = tid >> 6, c = (tid >> 2) & 15, pt = tid & 3;
			const double *zI = C.z + k * npad + I * 16;
			double acc = 0.0;
			for (int q = c + pt; q < 16; q += 4) acc += inv[frag_off(q, c)] * zI[q];
			acc += __shfl_xor_sync(0xffffffffu, acc, 1); acc += __shfl_xor_sync(0xffffffffu, acc, 2);
			if (pt == 0) C.xi[k * 16 + c] = acc;
		}
		__syncthreads();
		if (tid < 32) C.z[(tid >> 4) * npad + I * 16 + (tid & 15)] = C.xi[tid];
		for (int c = tid; c < (I - fI) * 16; c += KS_T) {
			const double *blk = sweep_blk(C, 1 + (c >> 4)) + frag_off(0, c & 15);
			double a0 = 0.0, a1 = 0.0;
#pragma unroll
			for (int q = 0; q < 16; ++q) { const double l = blk[((q >> 3) << 7) + ((q & 7) << 3)]; a0 += l * C.xi[q]; a1 += l * C.xi[16 + q]; }
			C.z[fI * 16 + c] -= a0; C.z[npad + fI * 16 + c] -= a1;
		}
		__syncthreads();
		sweep_release(C, I - fI + 1);
	}
}

/* L z = v for the two right-hand sides in C.z, in place */
__device__ __forceinline__ void sweep_forward(SweepCtx &C)
{
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, npad = C.npad;
	const int fr = lane >> 2, fc = lane & 3;
	for (int I = 0; I < C.nb; ++I) {
		mbar_wait(&C.mbar[C.c_seq % KS_NBAR], (unsigned)((C.c_seq / KS_NBAR) & 1));
		const int fI = C.fb[I], nbk = I - fI;
		/* every warp takes every fourth block of the row; a lane reads its operand-fragment chunks (conflict-free 128-bit
		 * loads): rows fr and 8 + fr, columns 4 fc + 2 h + {0, 1} */
		double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};      /* [row half][right-hand side] */
		for (int jb = warp; jb < nbk; jb += KS_T / 32) {
			const double2 *blk = reinterpret_cast<const double2 *>(sweep_blk(C, 1 + jb)) + lane;
			const double *z0 = C.z + (fI + jb) * 16 + 4 * fc, *z1 = z0 + npad;
#pragma unroll
			for (int h = 0; h < 2; ++h) {
				const double2 za = *reinterpret_cast<const double2 *>(z0 + 2 * h), zb = *reinterpret_cast<const double2 *>(z1 + 2 * h);
#pragma unroll
				for (int tn = 0; tn < 2; ++tn) {
					const double2 l = blk[tn * 64 + h * 32];
					acc[tn][0] += l.x * za.x + l.y * za.y;
					acc[tn][1] += l.x * zb.x + l.y * zb.y;
				}
			}
		}
#pragma unroll
		for (int tn = 0; tn < 2; ++tn)
#pragma unroll
			for (int k = 0; k < 2; ++k) {
				double a = acc[tn][k];
				a += __shfl_xor_sync(0xffffffffu, a, 1); a += __shfl_xor_sync(0xffffffffu, a, 2);
				if (fc == 0) C.part[(warp * 16 + tn * 8 + fr) * 2 + k] = a;
			}
		__syncthreads();
		{
			/* z_I = inv(L_II) (v_I - L[I,<I] z): four lanes per (right-hand side, row), each rebuilds the entries of the
			 * reduced right-hand side it needs from the warps' partial sums */
			const double *inv = sweep_blk(C, 0);
			const int k = tid >> 6, q = (tid >> 2) & 15, pt = tid & 3;
			double a = 0.0;
			for (int j = pt; j <= q; j += 4) {
				double sv = C.z[k * npad + I * 16 + j];
#pragma unroll
				for (int w = 0; w < KS_T / 32; ++w) sv -= C.part[(w * 16 + j) * 2 + k];
				a += inv[frag_off(q, j)] * sv;
			}
			a += __shfl_xor_sync(0xffffffffu, a, 1); a += __shfl_xor_sync(0xffffffffu, a, 2);
			C.xi[tid >> 2] = a;                           /* entry (k, q) = index k * 16 + q; the four lanes write the same value */
		}
		__syncthreads();
		if (tid < 32) C.z[(tid >> 4) * npad + I * 16 + (tid & 15)] = C.xi[tid];
		__syncthreads();
		sweep_release(C, nbk + 1);
	}
}

/* dense LU with partial pivoting of the 12 x 12 Woodbury matrix (one thread; row-major, in place) */
__device__ inline int lu12_factor(double *A, int *piv)
{
	const int n = 2 * IP_LM;
	for (int k = 0; k < n; ++k) {
		int p = k;
		for (int i = k + 1; i < n; ++i) if (fabs(A[i * n + k]) > fabs(A[p * n + k])) p = i;
		piv[k] = p;
		if (A[p * n + k] == 0.0) return 0;
		if (p != k) for (int j = 0; j < n; ++j) { const double t = A[k * n + j]; A[k * n + j] = A[p * n + j]; A[p * n + j] = t; }
		for (int i = k + 1; i < n; ++i) {
			A[i * n + k] /= A[k * n + k];
			for (int j = k + 1; j < n; ++j) A[i * n + j] -= A[i * n + k] * A[k * n + j];
		}
	}
	return 1;
}
__device__ inline void lu12_solve(const double *A, const int *piv, double *b)
{
	const int n = 2 * IP_LM;
	for (int k = 0; k < n; ++k) { const double t = b[k]; b[k] = b[piv[k]]; b[piv[k]] = t; }
	for (int k = 0; k < n; ++k) for (int i = k + 1; i < n; ++i) b[i] -= A[i * n + k] * b[k];
	for (int k = n - 1; k >= 0; --k) { for (int j = k + 1; j < n; ++j) b[k] -= A[k * n + j] * b[j]; b[k] /= A[k * n + k]; }
}

/* (J z0)_row and (J z1)_row: eight lanes share a row (columns strided), all 32 lanes of the warp call it together */
__device__ __forceinline__ void row_dot2(const DevTables &T, const double *Jv, const double *z0, const double *z1, int row, bool valid, double &j0, double &j1)
{
	double a0 = 0.0, a1 = 0.0;
	if (valid) {
		const Element &E = T.elems[T.row_elem[row]];
		const double *jr = Jv + E.valoff + (row - E.row0);
		const int16_t *cols = T.elem_cols + E.coloff;
		/* two columns per lane in flight (the loop is bound by its dependent loads: value, column index, then z); the sums keep
		 * their order */
		const int ncols = E.ncols, ld = E.ld;
		int a = threadIdx.x & 7;
		for (; a + 8 < ncols; a += 16) {
			const double jv = jr[a * ld], jw = jr[(a + 8) * ld];
			const int c = cols[a], d = cols[a + 8];
			a0 += jv * z0[c]; a1 += jv * z1[c];
			a0 += jw * z0[d]; a1 += jw * z1[d];
		}
		if (a < ncols) { const double jv = jr[a * ld]; const int c = cols[a]; a0 += jv * z0[c]; a1 += jv * z1[c]; }
	}
#pragma unroll
	for (int o = 1; o < 8; o <<= 1) { a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); }
	j0 = a0; j1 = a1;
}

template <int MINB>
__global__ void __launch_bounds__(KS_T, MINB)
kip_solve(DevTables T, DevWork W, qtos_options opt, int rc)
{
	const int pid = blockIdx.x;
	if (W.status[pid] != QTOS_RUNNING) return;
	extern __shared__ __align__(16) double sm[];
	const int npad = T.npad, q12 = 2 * IP_LM;
	double *z = sm;                                /* [2][npad] */
	double *ring = z + 2 * npad;                   /* [rc][256] */
	double *G = ring + (size_t)rc * 256;           /* [12][14]  Q'[Q, p_aff, p_cen] */
	double *Clu = G + q12 * 14;                    /* [12][12] */
	double *tt = Clu + q12 * q12;                  /* [2][12] */
	double *part = tt + 2 * q12;                   /* [4][16][2] */
	double *xi = part + 128;                       /* [2][16] */
	uint64_t *mbar = reinterpret_cast<uint64_t *>(xi + 32);   /* [KS_NBAR] */
	int *piv = reinterpret_cast<int *>(mbar + KS_NBAR);  /* [12] + flag */
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const double *Jv = WS(Jv, T.nJ), *r = WS(r, T.m), *s = WS(s, T.m), *vL = WS(zL, T.m), *vU = WS(zU, T.m);
	const double *dL = WS(dL, T.m), *dU = WS(dU, T.m), *y = WS(y, T.m), *Sig = WS(Sig, T.m);
	const double *RB = WS(RB, npad * IP_NRHS), *PB = WS(PB, npad * IP_NRHS);
	double *ip = WS(ipst, IP_N);
	double *ady = WS(ady, T.m), *cdy = WS(cdy, T.m), *eA = WS(rt, T.m), *eC = WS(st, T.m);
	const int n_pairs = (int)ip[IP_NPAIRS];
	const double avrg_compl = ip[IP_AVRG], rho = 1.0 / opt.delta_c;
	SweepCtx C; C.fb = T.fb; C.blkptr = T.blkptr; C.nb = T.nb; C.npad = npad; C.M = WS(M, T.nM); C.Dinv = WS(Dinv, T.nb * 256);
	C.z = z; C.ring = ring; C.part = part; C.xi = xi; C.mbar = mbar; C.rc = rc; C.total = (2 * opt.n_refine + 1) * T.nb;
	C.p_seq = C.p_pos = C.c_seq = C.c_pos = 0; C.p_free = rc;
	if (tid == 0) {
		for (int b = 0; b < KS_NBAR; ++b) mbar_init(&mbar[b], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		sweep_pump(C);                                 /* the first rows of L are on their way while the Woodbury term is set up */
	}
	for (int i = tid; i < npad; i += KS_T) {
		const double *o = PB + (size_t)(i >> 4) * 256 + (i & 15);
		z[i] = o[12 * 16]; z[npad + i] = o[13 * 16];
	}
	for (int i = tid; i < T.m; i += KS_T) { eA[i] = 0.0; eC[i] = 0.0; }          /* inequality rows of the Jc' dy operand stay zero */
	__syncthreads();
	int nlr = 2 * n_pairs;
	for (int pass = 0; pass <= opt.n_refine; ++pass) {
		if (nlr) {
			if (pass == 0) {
				/* G = Q'[Q | p_aff | p_cen] came with the factorization (k_factor's Gram matrix of the right-hand-side row) */
				const double *Gg = WS(G, 256);
				for (int q = tid; q < q12 * 14; q += KS_T) G[q] = Gg[(q / 14) * 16 + (q % 14)];
			} else {
				/* later passes: only Q'p of the two new right-hand sides; a warp takes every fourth column of Q */
				for (int a = warp; a < q12; a += KS_T / 32) {
					if ((a % IP_LM) >= n_pairs) continue;
					double a0 = 0.0, a1 = 0.0;
					for (int i = lane; i < npad; i += 32) {
						const double qa = PB[(size_t)(i >> 4) * 256 + a * 16 + (i & 15)];
						a0 += qa * z[i]; a1 += qa * z[npad + i];
					}
					for (int o = 16; o > 0; o >>= 1) { a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); }
					if (lane == 0) { G[a * 14 + 12] = a0; G[a * 14 + 13] = a1; }
				}
			}
			__syncthreads();
			if (pass == 0) {
				/* C = Mid - Q'Q on the slots in use, identity on the others */
				for (int q = tid; q < q12 * q12; q += KS_T) {
					const int a = q / q12, b = q - a * q12;
					const bool used = (a % IP_LM) < n_pairs && (b % IP_LM) < n_pairs;
					Clu[q] = used ? ip[IP_MID + q] - G[a * 14 + b] : (a == b ? 1.0 : 0.0);
				}
				__syncthreads();
				if (tid == 0) {
					piv[q12] = lu12_factor(Clu, piv);
					if (!piv[q12]) { ip[IP_NPAIRS] = 0.0; ip[IP_HEAD] = 0.0; ip[IP_SIGMA_W] = 1.0; }     /* singular: drop the pairs */
				}
				__syncthreads();
				if (!piv[q12]) nlr = 0;
			}
			if (nlr) {
				if (tid < 2) {
					double *t = tt + tid * q12;
					for (int a = 0; a < q12; ++a) t[a] = (a % IP_LM) < n_pairs ? G[a * 14 + 12 + tid] : 0.0;
					lu12_solve(Clu, piv, t);
				}
				__syncthreads();
				for (int i = tid; i < npad; i += KS_T) {
					const double *o = PB + (size_t)(i >> 4) * 256 + (i & 15);
					double a0 = 0.0, a1 = 0.0;
					for (int a = 0; a < q12; ++a) {
						if ((a % IP_LM) >= n_pairs) continue;
						const double qa = o[a * 16];
						a0 += qa * tt[a]; a1 += qa * tt[q12 + a];
					}
					z[i] += a0; z[npad + i] += a1;
				}
			}
			__syncthreads();
		}
		sweep_backward(C);                         /* z = dx of both directions, permuted order */
		/* multiplier-method update on the equality rows: dy += rho (Jc dx - b2) */
		for (int base = 0; base < T.n_eq; base += KS_T / 8) {
			const int idx = base + (tid >> 3);
			const bool valid = idx < T.n_eq;
			const int i = valid ? T.eq_rows[idx] : 0;
			double j0, j1;
			row_dot2(T, Jv, z, z + npad, i, valid, j0, j1);
			if (valid && (tid & 7) == 0) {
				const double da = (pass ? ady[i] : 0.0) + rho * (j0 + r[i]), dc = (pass ? cdy[i] : 0.0) + rho * j1;      /* b2 = -r (affine), 0 (centering) */
				ady[i] = da; cdy[i] = dc; eA[i] = da; eC[i] = dc;
			}
		}
		if (pass == opt.n_refine) break;
		__syncthreads();
		/* next right-hand side: v = v0 - Jc' dy, then p = L^-1 v */
		for (int i = tid; i < npad; i += KS_T) {
			const double *o = RB + (size_t)(i >> 4) * 256 + (i & 15);
			double ga, gc;
			jt_gather2(T.jgc_ptr, T.jgc, Jv, eA, eC, i, ga, gc);
			z[i] = i < T.n_free ? o[12 * 16] - ga : 0.0; z[npad + i] = i < T.n_free ? o[13 * 16] - gc : 0.0;
		}
		__syncthreads();
		sweep_forward(C);
	}
	__syncthreads();
	/* ---- expansion of the condensed rows: ds, dy, dvL, dvU of both directions; dx out */
	double *adx = WS(adx, npad), *cdx = WS(cdx, npad);
	for (int i = tid; i < npad; i += KS_T) { adx[i] = z[i]; cdx[i] = z[npad + i]; }
	double *ads = WS(ads, T.m), *advL = WS(advL, T.m), *advU = WS(advU, T.m), *cds = WS(cds, T.m), *cdvL = WS(cdvL, T.m), *cdvU = WS(cdvU, T.m);
	/* Jd dx of both directions first (eight lanes per row), parked in the ring's shared memory (the sweeps are over), then
	 * the row-wise expansion with one thread per row: all lanes busy, neighbouring rows read neighbouring addresses */
	double *jd = ring;                             /* [2][n_ineq] */
	for (int base = 0; base < T.n_ineq; base += KS_T / 8) {
		const int idx = base + (tid >> 3);
		const bool valid = idx < T.n_ineq;
		double j0, j1;
		row_dot2(T, Jv, z, z + npad, valid ? T.iq_rows[idx] : 0, valid, j0, j1);
		if (valid && (tid & 7) == 0) { jd[idx] = j0; jd[T.n_ineq + idx] = j1; }
	}
	__syncthreads();
	for (int idx = tid; idx < T.n_ineq; idx += KS_T) {
		const int i = T.iq_rows[idx];
		const double j0 = jd[idx], j1 = jd[T.n_ineq + idx];
		const int fl = T.row_flags[i];
		double augA = -(-y[i] - vL[i] + vU[i]), augC = 0.0, sl = 1.0, su = 1.0;
		if (fl & ROW_HASL) { sl = s[i] - dL[i]; augA += (-sl * vL[i]) / sl; augC += avrg_compl / sl; }
		if (fl & ROW_HASU) { su = dU[i] - s[i]; augA -= (-su * vU[i]) / su; augC -= avrg_compl / su; }
		const double dsa = j0 + (r[i] - s[i]), dsc = j1;                     /* ds = Jd dx - b2, b2 = -(d - s) / 0 */
		ads[i] = dsa; cds[i] = dsc;
		ady[i] = Sig[i] * dsa - augA; cdy[i] = Sig[i] * dsc - augC;
		advL[i] = (fl & ROW_HASL) ? (-sl * vL[i] - vL[i] * dsa) / sl : 0.0;
		cdvL[i] = (fl & ROW_HASL) ? (avrg_compl - vL[i] * dsc) / sl : 0.0;
		advU[i] = (fl & ROW_HASU) ? (-su * vU[i] + vU[i] * dsa) / su : 0.0;
		cdvU[i] = (fl & ROW_HASU) ? (avrg_compl + vU[i] * dsc) / su : 0.0;
	}
}

/* ------------------------------------------------------------------ kip_step */

/* QualityFunctionMuOracle::CalculateQualityFunction (2-norm squared, no centrality term) at sigma */
__device__ __forceinline__ double ip_quality(const DevTables &T, const DevWork &W, int pid, double sg, double tau, double gl2, double pr2, double *red)
{
	const double *s = WS(s, T.m), *vL = WS(zL, T.m), *vU = WS(zU, T.m), *dL = WS(dL, T.m), *dU = WS(dU, T.m);
	const double *ads = WS(ads, T.m), *advL = WS(advL, T.m), *advU = WS(advU, T.m), *cds = WS(cds, T.m), *cdvL = WS(cdvL, T.m), *cdvU = WS(cdvU, T.m);
	double a[2] = {1.0, 1.0};
	for (int i = threadIdx.x; i < T.m; i += blockDim.x) {
		const int fl = T.row_flags[i];
		if (fl & ROW_EQ) continue;
		const double d = ads[i] + sg * cds[i];
		if (fl & ROW_HASL) { const double uL = advL[i] + sg * cdvL[i]; if (d < 0) a[0] = fmin(a[0], -tau * (s[i] - dL[i]) / d); if (uL < 0) a[1] = fmin(a[1], -tau * vL[i] / uL); }
		if (fl & ROW_HASU) { const double uU = advU[i] + sg * cdvU[i]; if (-d < 0) a[0] = fmin(a[0], -tau * (dU[i] - s[i]) / -d); if (uU < 0) a[1] = fmin(a[1], -tau * vU[i] / uU); }
	}
	{ const int ops[2] = {2, 2}; block_reduce<2>(a, ops, red); }
	const double ap = a[0], ad = a[1];
	double cc[1] = {0.0};
	for (int i = threadIdx.x; i < T.m; i += blockDim.x) {
		const int fl = T.row_flags[i];
		if (fl & ROW_EQ) continue;
		const double d = ads[i] + sg * cds[i];
		if (fl & ROW_HASL) { const double t = ((s[i] - dL[i]) + ap * d) * (vL[i] + ad * (advL[i] + sg * cdvL[i])); cc[0] += t * t; }
		if (fl & ROW_HASU) { const double t = ((dU[i] - s[i]) + ap * -d) * (vU[i] + ad * (advU[i] + sg * cdvU[i])); cc[0] += t * t; }
	}
	{ const int ops[1] = {0}; block_reduce<1>(cc, ops, red); }
	const double n_dual = T.n_free + T.n_ineq, n_pri = T.m, n_comp = T.n_bounds;
	return (1 - ad) * (1 - ad) * gl2 / n_dual + (1 - ap) * (1 - ap) * pr2 / n_pri + cc[0] / n_comp;
}

template <int MINB>
__global__ void __launch_bounds__(QTOS_THREADS, MINB)
kip_step(DevTables T, DevWork W, const qtos_problem *probs, const DevHeightfield *hfs, int n_hf, qtos_options opt)
{
	const int pid = blockIdx.x;
	if (W.status[pid] != QTOS_RUNNING) return;
	extern __shared__ __align__(16) double sm[];
	double *b = sm;                       /* [npad] dx (permuted order) */
	double *red = b + T.npad;             /* [8*32] */
	__shared__ double sfphi[IP_FILTER_MAX], sfth[IP_FILTER_MAX];
	const int tid = threadIdx.x;
	const double *Jv = WS(Jv, T.nJ), *sc = WS(sc, T.m), *dLb = WS(dL, T.m), *dUb = WS(dU, T.m);
	const double *adx = WS(adx, T.npad), *cdx = WS(cdx, T.npad);
	const double *ads = WS(ads, T.m), *ady = WS(ady, T.m), *advL = WS(advL, T.m), *advU = WS(advU, T.m);
	const double *cds = WS(cds, T.m), *cdy = WS(cdy, T.m), *cdvL = WS(cdvL, T.m), *cdvU = WS(cdvU, T.m);
	double *s = WS(s, T.m), *y = WS(y, T.m), *vL = WS(zL, T.m), *vU = WS(zU, T.m), *r = WS(r, T.m);
	double *rt = WS(rt, T.m), *st = WS(st, T.m), *x = WS(x, T.n_all), *xt = WS(xt, T.n_all), *ip = WS(ipst, IP_N);
	double mu = ip[IP_MU], tau = ip[IP_TAU];
	const int free_mode = (int)ip[IP_FREE];
	int nfilter = (int)ip[IP_NFILTER];
	const double avrg_compl = ip[IP_AVRG], nlp_error = ip[IP_ERR], theta = ip[IP_THETA], gl2 = ip[IP_GL2], pr2 = ip[IP_PR2], mu_max = ip[IP_MU_MAX];
	double theta_max = ip[IP_TH_MAX], theta_min = ip[IP_TH_MIN];
	if (tid < IP_FILTER_MAX) { sfphi[tid] = ip[IP_FPHI + tid]; sfth[tid] = ip[IP_FTH + tid]; }
	{
		/* A factorization that met a non-positive pivot or produced a non-finite direction is repeated with W + delta_w I
		 * (Ipopt's PDPerturbationHandler: 1e-4 the first time, a third of the last successful value later, then x100 / x8):
		 * the problem sits this step out and kip_prepare re-enters the same iteration with the larger diagonal. */
		double nf[1] = {0.0};
		for (int i = tid; i < T.npad; i += blockDim.x) nf[0] += 0.0 * adx[i] + 0.0 * cdx[i];
		{ const int ops[1] = {0}; block_reduce<1>(nf, ops, red); }
		if (!(nf[0] == 0.0) || (W.flags[pid] & 1)) {
			if (tid == 0) {
				const double last = ip[IP_DELTA_LAST];
				double dw = ip[IP_RETRY] != 0.0 ? ip[IP_DELTA_W] : 0.0;
				dw = dw == 0.0 ? (last == 0.0 ? 1e-4 : fmax(1e-20, last / 3.0)) : dw * (last == 0.0 ? 100.0 : 8.0);
				if (dw > 1e40) { W.status[pid] = QTOS_STEP_FAILED; atomicSub(W.n_running, 1); }
				ip[IP_DELTA_W] = dw; ip[IP_RETRY] = 1.0;
			}
			return;
		}
	}
	const double mu_min = fmin(1e-11, 0.5 * fmin(opt.tol, opt.compl_inf_tol));
	double sigma = mu / avrg_compl;
	if (free_mode) {
		tau = fmax(0.99, 1.0 - nlp_error);
		/* QualityFunctionMuOracle::CalculateMu: golden section on sigma (linear scale), at most 8 steps */
#define QF(sig_) ip_quality(T, W, pid, (sig_), tau, gl2, pr2, red)
		const double s_1m = 1.0 - 1e-2;
		const double qf_1 = QF(1.0), qf_1m = QF(s_1m);
		double s_up, s_lo, q_up, q_lo; bool search = true;
		if (qf_1m > qf_1) { s_up = fmin(100.0, mu_max / avrg_compl); s_lo = 1.0; q_up = -100.0; q_lo = qf_1; if (s_lo >= s_up) { sigma = s_up; search = false; } }
		else { s_lo = fmax(1e-6, mu_min / avrg_compl); s_up = fmin(fmax(s_lo, s_1m), mu_max / avrg_compl); q_up = qf_1m; q_lo = -100.0; if (s_lo >= s_up) { sigma = s_lo; search = false; } }
		if (search) {
			const double s_up0 = s_up, s_lo0 = s_lo, gfac = (3.0 - sqrt(5.0)) / 2.0;
			double m1 = s_lo + gfac * (s_up - s_lo), m2 = s_lo + (1 - gfac) * (s_up - s_lo);
			double q1 = QF(m1), q2 = QF(m2);
			int k = 0;
			while ((s_up - s_lo) >= 1e-2 * s_up && k < 8) {
				k++;
				if (q1 > q2) { s_lo = m1; q_lo = q1; m1 = m2; q1 = q2; m2 = s_lo + (1 - gfac) * (s_up - s_lo); q2 = QF(m2); }
				else { s_up = m2; q_up = q2; m2 = m1; q2 = q1; m1 = s_lo + gfac * (s_up - s_lo); q1 = QF(m1); }
			}
			double q;
			if (q1 < q2) { sigma = m1; q = q1; } else { sigma = m2; q = q2; }
			if (s_up == s_up0) { double qt = q_up; if (qt < 0) qt = QF(s_up); if (qt < q) { sigma = s_up; q = qt; } }
			else if (s_lo == s_lo0) { double qt = q_lo; if (qt < 0) qt = QF(s_lo); if (qt < q) { sigma = s_lo; q = qt; } }
		}
#undef QF
		mu = fmax(fmin(fmax(sigma * avrg_compl, mu_min), mu_max), mu_min);
		sigma = mu / avrg_compl;
		nfilter = 0;
	}
	/* ---- search direction aff + sigma cen; fraction to the boundary; barrier objective and its directional derivative */
	/* v: 0 dnorm (max) 1 alpha_max (min) 2 alpha_du (min) 3 phi0 (sum) 4 gBD (sum) */
	double v[5] = {0.0, 1.0, 1.0, 0.0, 0.0};
	const double df = T.cost_c ? ip[IP_OBJ_SCALE] : 1.0;
	for (int i = tid; i < T.npad; i += blockDim.x) {
		const double d = adx[i] + sigma * cdx[i]; b[i] = d; v[0] = fmax(v[0], fabs(d));
		if (T.cost_c) { const int var = T.var_of_perm[i]; if (var >= 0) v[4] += 2.0 * df * T.cost_c[var] * x[var] * d; }   /* grad f . dx */
	}
	for (int i = tid; i < T.m; i += blockDim.x) {
		const int fl = T.row_flags[i];
		if (fl & ROW_EQ) continue;
		const double d = ads[i] + sigma * cds[i];
		v[0] = fmax(v[0], fabs(d));
		if (fl & ROW_HASL) {
			const double sl = s[i] - dLb[i], u = advL[i] + sigma * cdvL[i];
			if (d < 0) v[1] = fmin(v[1], -tau * sl / d);
			if (u < 0) v[2] = fmin(v[2], -tau * vL[i] / u);
			v[3] -= mu * log(sl); v[4] -= mu * d / sl;
		}
		if (fl & ROW_HASU) {
			const double su = dUb[i] - s[i], u = advU[i] + sigma * cdvU[i];
			if (-d < 0) v[1] = fmin(v[1], -tau * su / -d);
			if (u < 0) v[2] = fmin(v[2], -tau * vU[i] / u);
			v[3] -= mu * log(su); v[4] += mu * d / su;
		}
	}
	{ const int ops[5] = {1, 2, 2, 0, 0}; block_reduce<5>(v, ops, red); }
	const double dnorm = v[0], alpha_max = v[1], alpha_du = v[2], phi0 = v[3] + (T.cost_c ? ip[IP_FVAL] : 0.0), gBD = v[4];
	/* ---- filter line search (BacktrackingLineSearch + FilterLSAcceptor) */
	if (theta_max < 0.0) { theta_max = 1e4 * fmax(1.0, theta); theta_min = 1e-4 * fmax(1.0, theta); }
	double alpha_min = 1e-5;
	if (gBD < 0) {
		alpha_min = fmin(1e-5, 1e-8 * theta / (-gBD));
		if (theta <= theta_min) alpha_min = fmin(alpha_min, pow(theta, 1.1) / pow(-gBD, 2.3));
	}
	alpha_min *= 0.05;
	const double eps10 = 10.0 * IP_EPS;
	const int hid = probs[pid].hf_id >= 0 && probs[pid].hf_id < n_hf ? probs[pid].hf_id : 0;
	const DevHeightfield hf = hfs[hid];
	double alpha = alpha_max, th_t = 0.0, ph_t = 0.0;
	bool accepted = false;
	int ls = 0;
	while (alpha > alpha_min || ls == 0) {
		ls++;
		for (int i = tid; i < T.n_all; i += blockDim.x) {
			const int p = T.perm_of_var[i];
			xt[i] = p >= 0 ? x[i] + alpha * b[p] : x[i];
		}
		__syncthreads();
		eval_g_block(T, hf, xt, rt);
		__syncthreads();
		double u[2] = {0.0, 0.0};
		for (int i = tid; i < T.m; i += blockDim.x) {
			const int fl = T.row_flags[i];
			const double g = rt[i];
			if (fl & ROW_EQ) { const double c = sc[i] * (g - T.gl[i]); rt[i] = c; u[0] += fabs(c); continue; }
			const double d = sc[i] * g, sn = s[i] + alpha * (ads[i] + sigma * cds[i]);
			rt[i] = d; st[i] = sn;
			u[0] += fabs(d - sn);
			if (fl & ROW_HASL) u[1] -= mu * log(sn - dLb[i]);
			if (fl & ROW_HASU) u[1] -= mu * log(dUb[i] - sn);
		}
		if (T.cost_c) for (int i = tid; i < T.n_all; i += blockDim.x) u[1] += df * T.cost_c[i] * xt[i] * xt[i];
		{ const int ops[2] = {0, 0}; block_reduce<2>(u, ops, red); }
		th_t = u[0]; ph_t = u[1];
		bool ok = false;
		if (th_t <= theta_max && ph_t == ph_t && fabs(ph_t) < 1e300) {
			const bool switching = gBD < 0 && alpha * pow(-gBD, 2.3) > pow(theta, 1.1);
			if (theta <= theta_min && switching) ok = ph_t - phi0 - 1e-8 * alpha * gBD <= eps10 * fabs(phi0);
			else {
				ok = th_t - (1 - 1e-5) * theta <= eps10 * fabs(theta) || ph_t - phi0 + 1e-8 * theta <= eps10 * fabs(phi0);
				if (ok && ph_t > phi0) {                         /* obj_max_inc 5 */
					const double bas = fabs(phi0) > 10.0 ? log10(fabs(phi0)) : 1.0;
					ok = log10(ph_t - phi0) <= 5.0 + bas;
				}
			}
			if (ok) for (int k = 0; k < nfilter; ++k) if (!(ph_t <= sfphi[k] || th_t <= sfth[k])) { ok = false; break; }
		}
		if (ok) { accepted = true; break; }
		alpha *= 0.5;
	}
	if (!accepted) {                                              /* Ipopt would enter the restoration phase here */
		if (tid == 0) { W.status[pid] = QTOS_STEP_FAILED; atomicSub(W.n_running, 1); ip[IP_LS] = ls; }
		return;
	}
	const bool switching = gBD < 0 && alpha * pow(-gBD, 2.3) > pow(theta, 1.1);
	const bool armijo = ph_t - phi0 - 1e-8 * alpha * gBD <= eps10 * fabs(phi0);
	const bool ftype = switching && armijo;
	/* ---- accept the trial point */
	for (int i = tid; i < T.n_all; i += blockDim.x) x[i] = xt[i];
	double cs[1] = {0.0};
	for (int i = tid; i < T.m; i += blockDim.x) {
		const int fl = T.row_flags[i];
		r[i] = rt[i];
		y[i] += alpha * (ady[i] + sigma * cdy[i]);
		if (fl & ROW_EQ) continue;
		s[i] = st[i];
		if (fl & ROW_HASL) { const double nv = vL[i] + alpha_du * (advL[i] + sigma * cdvL[i]); vL[i] = nv; cs[0] += (st[i] - dLb[i]) * nv; }
		if (fl & ROW_HASU) { const double nv = vU[i] + alpha_du * (advU[i] + sigma * cdvU[i]); vU[i] = nv; cs[0] += (dUb[i] - st[i]) * nv; }
	}
	{ const int ops[1] = {0}; block_reduce<1>(cs, ops, red); }
	/* IpoptAlgorithm::correct_bound_multiplier (kappa_sigma 1e10): free mode uses the trial average complementarity */
	const double mu_c = free_mode ? fmin(cs[0] / (double)(T.n_bounds > 0 ? T.n_bounds : 1), 1e3) : mu;
	for (int i = tid; i < T.m; i += blockDim.x) {
		const int fl = T.row_flags[i];
		if (fl & ROW_HASL) { const double sl = s[i] - dLb[i]; vL[i] = fmin(fmax(vL[i], mu_c / (1e10 * sl)), 1e10 * mu_c / sl); }
		if (fl & ROW_HASU) { const double su = dUb[i] - s[i]; vU[i] = fmin(fmax(vU[i], mu_c / (1e10 * su)), 1e10 * mu_c / su); }
	}
	__syncthreads();
	/* grad_x L(x_k, lambda_{k+1}) = grad f(x_k) + J(x_k)' lambda_{k+1} for the next limited-memory pair (the Jacobian values are
	 * still those of x_k, and kip_prepare left x_k in W.lastx) */
	double *gJold = WS(gJold, T.npad);
	const double *lastx = WS(lastx, T.npad);
	for (int i = tid; i < T.npad; i += blockDim.x) {
		double g = jt_gather(T, Jv, y, i);
		if (T.cost_c) { const int var = T.var_of_perm[i]; if (var >= 0) g += 2.0 * df * T.cost_c[var] * lastx[i]; }
		gJold[i] = g;
	}
	if (tid == 0) {
		if (!ftype) {
			if (nfilter == IP_FILTER_MAX) { for (int k = 1; k < IP_FILTER_MAX; ++k) { ip[IP_FPHI + k - 1] = sfphi[k]; ip[IP_FTH + k - 1] = sfth[k]; } nfilter--; }
			ip[IP_FPHI + nfilter] = phi0 - 1e-8 * theta; ip[IP_FTH + nfilter] = (1 - 1e-5) * theta; nfilter++;
		}
		ip[IP_NFILTER] = nfilter; ip[IP_MU] = mu; ip[IP_TAU] = tau; ip[IP_TH_MAX] = theta_max; ip[IP_TH_MIN] = theta_min;
		ip[IP_ALPHA_PR] = alpha; ip[IP_ALPHA_DU] = alpha_du; ip[IP_DNORM] = dnorm; ip[IP_LS] = ls; ip[IP_TAG] = ftype ? 'f' : 'h';
		ip[IP_ITER] += 1.0;
		if (ip[IP_RETRY] != 0.0) { ip[IP_DELTA_LAST] = ip[IP_DELTA_W]; ip[IP_RETRY] = 0.0; ip[IP_DELTA_W] = 0.0; }
	}
}

/* ------------------------------------------------------------------ heightfield tile staging (measurement only)
 *
 * north_star asks for heightfield tiles "staged through TMA/shared memory".  The product reads the four corner heights of a
 * query straight from global memory (read-only path): an evaluation makes ~50-100 queries scattered over the <= 1 m^2 a window
 * covers.  These two kernels measure the alternative on exactly that access pattern -- one CTA per GROUP of queries (the
 * queries of one evaluation): k_height_group reads global memory, k_height_staged first brings the bounding box of the group's
 * cells into shared memory with one 1-D bulk asynchronous copy per grid row (TMA unit, mbarrier completion) and then answers
 * from there with the same bit-exact arithmetic.  qtos_measure_heightfield_staging times both. */
__global__ void k_height_group(DevHeightfield hf, const double *xy, int group, double *h_out)
{
	const int q = blockIdx.x * group + threadIdx.x;
	if ((int)threadIdx.x < group) h_out[q] = qtos_height(hf, xy[2 * q], xy[2 * q + 1]);
}

#define HS_MAX_ROWS 72
#define HS_MAX_COLS 80
__global__ void k_height_staged(DevHeightfield hf, const double *xy, int group, double *h_out, int *fallbacks)
{
	__shared__ __align__(16) double tile[HS_MAX_ROWS * HS_MAX_COLS];
	__shared__ long long box[4];
	__shared__ uint64_t bar;
	const int tid = threadIdx.x, q = blockIdx.x * group + tid;
	long long c[4] = {0, 0, 0, 0};
	double x = 0.0, y = 0.0;
	if (tid < group) { x = xy[2 * q]; y = xy[2 * q + 1]; qtos_height_cell(hf, x, y, c); }
	/* bounding box of the group's cells (warp shuffles + one shared-memory round for two warps) */
	long long lo0 = tid < group ? c[0] : (1LL << 40), hi0 = tid < group ? c[2] : -1, lo1 = tid < group ? c[1] : (1LL << 40), hi1 = tid < group ? c[3] : -1;
	for (int o = 16; o > 0; o >>= 1) {
		lo0 = min(lo0, __shfl_xor_sync(0xffffffffu, lo0, o)); hi0 = max(hi0, __shfl_xor_sync(0xffffffffu, hi0, o));
		lo1 = min(lo1, __shfl_xor_sync(0xffffffffu, lo1, o)); hi1 = max(hi1, __shfl_xor_sync(0xffffffffu, hi1, o));
	}
	__shared__ long long part[2][4];
	if ((tid & 31) == 0) { part[tid >> 5][0] = lo0; part[tid >> 5][1] = hi0; part[tid >> 5][2] = lo1; part[tid >> 5][3] = hi1; }
	if (tid == 0) mbar_init(&bar, 1);
	__syncthreads();
	if (tid == 0) {
		const int nw = (blockDim.x + 31) >> 5;
		long long a = part[0][0], b = part[0][1], cc = part[0][2], d = part[0][3];
		for (int w = 1; w < nw; ++w) { a = min(a, part[w][0]); b = max(b, part[w][1]); cc = min(cc, part[w][2]); d = max(d, part[w][3]); }
		cc &= ~1LL;                                        /* 16-byte aligned row starts */
		long long cols = ((d - cc + 1) + 1) & ~1LL;         /* and sizes */
		if (cc + cols > hf.ny) cols = (hf.ny - cc) & ~1LL;
		box[0] = a; box[1] = b - a + 1; box[2] = cc; box[3] = cols;
		const bool fits = box[1] <= HS_MAX_ROWS && cols <= HS_MAX_COLS && cols > 0 && d < cc + cols && (hf.ny & 1) == 0;
		if (!fits) { box[3] = 0; atomicAdd(fallbacks, 1); }
		else {
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
			mbar_expect_tx(&bar, (unsigned)(box[1] * cols * 8));
			for (long long r = 0; r < box[1]; ++r) bulk_g2s(tile + r * cols, hf.h + (a + r) * hf.ny + cc, (unsigned)(cols * 8), &bar);
		}
	}
	__syncthreads();
	if (box[3] == 0) { if (tid < group) h_out[q] = qtos_height(hf, x, y); return; }
	mbar_wait(&bar, 0u);
	if (tid >= group) return;
	const long long cols = box[3];
	const double res = hf.res;
	const double x0 = __dadd_rn(__dmul_rn((double)c[0], res), -1.0), x1 = __dadd_rn(__dmul_rn((double)c[2], res), -1.0);
	const double y0 = __dadd_rn(__dmul_rn((double)c[1], res), -1.0), y1 = __dadd_rn(__dmul_rn((double)c[3], res), -1.0);
	const double z00 = tile[(c[0] - box[0]) * cols + (c[1] - box[2])], z01 = tile[(c[0] - box[0]) * cols + (c[3] - box[2])];
	const double z10 = tile[(c[2] - box[0]) * cols + (c[1] - box[2])], z11 = tile[(c[2] - box[0]) * cols + (c[3] - box[2])];
	const double sxy = __ddiv_rn(1.0, __dmul_rn(res, res));
	const double u0 = __dmul_rn(sxy, __dsub_rn(x1, x)), u1 = __dmul_rn(sxy, __dsub_rn(x, x0));
	const double w0 = __dadd_rn(__dmul_rn(u0, z00), __dmul_rn(u1, z10));
	const double w1 = __dadd_rn(__dmul_rn(u0, z01), __dmul_rn(u1, z11));
	h_out[q] = __dadd_rn(__dmul_rn(w0, __dsub_rn(y1, y)), __dmul_rn(w1, __dsub_rn(y, y0)));
}

#endif
