/*
 * qtos_kernels.cu -- sm_100a kernels of the batched interior-point gait-plan solver.
 *
 * One thread block per problem (k_jac: four adjacent lanes per sample; k_asm: one block per problem and block row).  All launches of a
 * solve go to the context's stream (k_jac_rom beside k_jac_dyn on a second one).  Kernels shared by both algorithms:
 *   k_init      x0, fixed variables, g(x0), constant J elements; after the first k_jac: row scaling, slack / multiplier
 *               and algorithm-state initialisation (ref: nlp_formulation.cc:100-190; Ipopt initialisation)
 *   k_jac_*     dynamics + range-of-motion Jacobian element blocks at x
 *   k_asm       M = sigma I + J' D J, one block row per CTA, owner-computes gather into shared memory
 *   k_factor    left-looking block-skyline Cholesky on 16x16 blocks with FP64 tensor-core MMAs (DMMA m8n8k4);
 *               <., 0>: forward substitution of one right-hand side fused, then the backward substitution -> dx
 *               <., 1>: the 16 right-hand sides of the IPOPT path forward-substituted as one more block row, plus their Gram matrix
 *               <., ., 1>: the same code with an L1-rich carve-out, for launches with at most one active problem per SM
 *   k_csv       1 kHz trajectory sampler, any row range (ref: main.cpp:92-131)
 *   k_height    batched heightfield queries (ref: custom_terrain.cpp:51-94)
 *   k_results, k_cost, k_admit, k_records, k_select   results, post-hoc plan cost, pool admission, best-plan selection
 * QTOS_ALG_FAST (oracle/towr_ipm.c):  k_prepare (J'y, error measures, termination, monotone barrier update, Sigma, rhs = -J'w)
 *   and k_step (step recovery, fraction to the boundary, l1-merit backtracking with in-kernel g(x), iterate update).
 * QTOS_ALG_IPOPT (oracle/towr_ipopt.c), the default: kip_prepare, kip_solve, kip_step in qtos_ipopt.cuh.
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

__device__ __forceinline__ double qnan_fill() { return nan(""); }

__global__ void __launch_bounds__(QTOS_THREADS)
k_init(DevTables T, DevWork W, const qtos_problem *probs, const DevHeightfield *hfs, int n_hf, qtos_options opt, int do_solver_init, const int *slots)
{
	const int pid = slots ? slots[blockIdx.x] : blockIdx.x;      /* streaming: the workspace slots that were just refilled */
	const qtos_problem &pr = probs[pid];
	double *x = WS(x, T.n_all), *P = WS(P, 32), *sc = WS(sc, T.m), *r = WS(r, T.m), *Jv = WS(Jv, T.nJ);
	__shared__ double sfin[6][3];
	const bool hf_ok = pr.hf_id >= 0 && pr.hf_id < n_hf && hfs[pr.hf_id].h != nullptr;
	const DevHeightfield hf = hfs[hf_ok ? pr.hf_id : 0];
	if (!hf_ok && do_solver_init) {
		/* unknown or released heightfield: the window ends here with Invalid_Number_Detected instead of being solved on some
		 * other grid (host-resident problems are refused before the launch; device-resident ones can only be caught here) */
		if (do_solver_init == 1 && threadIdx.x == 0) {
			double *scal = WS(scal, 16);
			const double qnan = nan("");
			scal[SC_DUAL] = scal[SC_THETA] = scal[SC_COMPL] = scal[SC_VIOL] = scal[SC_E0] = qnan; scal[SC_MU] = 0.0;
			W.status[pid] = QTOS_INVALID_NUMBER; W.iters[pid] = 0; W.flags[pid] = 0; atomicSub(W.n_running, 1);
		}
		if (do_solver_init == 1) for (int v = threadIdx.x; v < T.n_all; v += blockDim.x) x[v] = qnan_fill();
		return;
	}
	/* do_solver_init: 0 = x0 only (qtos_get_initial / qtos_eval); 1 = x0, g(x0), constant Jacobian elements, problem
	 * marked running -- the x-dependent elements then come from k_jac_dyn / k_jac_rom (one thread per sample instead
	 * of one block per problem); 2 = row scaling from J(x0), scaled J, slacks and multipliers */
	/* 3 = 1 for the second attempt of a window (qtos_options.retry_failed): the iterations spent so far are kept */
	const bool restart = do_solver_init == 3;
	if (restart) do_solver_init = 1;
	if (do_solver_init == 2) goto scaling;
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
	if (threadIdx.x == 0) { W.status[pid] = QTOS_RUNNING; W.iters[pid] = restart ? W.iters[pid] : 0; W.flags[pid] = 0; }
	return;
scaling:
	/* gradient-based row scaling min(1, 100/||row||_inf) at x0 */
	for (int i = threadIdx.x; i < T.m; i += blockDim.x) {
		const Element &E = T.elems[T.row_elem[i]];
		const int rr = i - E.row0;
		double mx = 0.0;
		for (int a = 0; a < E.ncols; ++a) mx = fmax(mx, fabs(Jv[E.valoff + a * E.ld + rr]));
		sc[i] = mx > 100.0 ? fmax(100.0 / mx, 1e-8) : 1.0;
	}
	__syncthreads();
	for (int e = threadIdx.x; e < T.n_elem; e += blockDim.x) {
		const Element &E = T.elems[e];
		for (int a = 0; a < E.ncols; ++a) for (int rr = 0; rr < E.nrows; ++rr) Jv[E.valoff + a * E.ld + rr] *= sc[E.row0 + rr];
	}
	double *s = WS(s, T.m), *y = WS(y, T.m), *zL = WS(zL, T.m), *zU = WS(zU, T.m), *dL = WS(dL, T.m), *dU = WS(dU, T.m);
	for (int i = threadIdx.x; i < T.m; i += blockDim.x) {
		const int fl = T.row_flags[i];
		const double g = r[i];
		y[i] = 0.0; zL[i] = zU[i] = 0.0; dL[i] = dU[i] = 0.0; s[i] = 0.0;
		if (fl & ROW_EQ) { r[i] = sc[i] * (g - T.gl[i]); continue; }
		r[i] = sc[i] * g;
		double v = r[i], lo = 0, hi = 0;
		if (opt.algorithm == QTOS_ALG_IPOPT) {            /* bound_relax_factor acts on the unscaled bound */
			if (fl & ROW_HASL) { lo = sc[i] * (T.gl[i] - 1e-8 * fmax(1.0, fabs(T.gl[i]))); dL[i] = lo; zL[i] = 1.0; }
			if (fl & ROW_HASU) { hi = sc[i] * (T.gu[i] + 1e-8 * fmax(1.0, fabs(T.gu[i]))); dU[i] = hi; zU[i] = 1.0; }
		} else {
			if (fl & ROW_HASL) { const double b = sc[i] * T.gl[i]; lo = b - 1e-8 * fmax(1.0, fabs(b)); dL[i] = lo; zL[i] = 1.0; }
			if (fl & ROW_HASU) { const double b = sc[i] * T.gu[i]; hi = b + 1e-8 * fmax(1.0, fabs(b)); dU[i] = hi; zU[i] = 1.0; }
		}
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
	}
	if (opt.algorithm == QTOS_ALG_IPOPT) {
		/* AdaptiveMuUpdate::InitializeImpl, LimMemQuasiNewtonUpdater (limited_memory_init_val 1), empty filters */
		double *ip = WS(ipst, IP_N), *tr = WS(trace, QTOS_TRACE_ITERS * QTOS_TRACE_COLS);
		const double base = (double)W.iters[pid];          /* 0, or the iterations of the first attempt */
		for (int i = threadIdx.x; i < IP_N; i += blockDim.x)
			ip[i] = i == IP_MU || i == IP_FREE || i == IP_SIGMA_W || i == IP_OBJ_SCALE ? 1.0 : (i == IP_MU_MAX || i == IP_TH_MAX || i == IP_TH_MIN ? -1.0 : (i == IP_AMU_THMIN ? 1e300 :
			        (i == IP_SIGMA_MIN ? opt.lm_init_val_min : (i == IP_ITER_BASE ? base : 0.0))));
		for (int i = threadIdx.x; i < QTOS_TRACE_ITERS * QTOS_TRACE_COLS; i += blockDim.x) tr[i] = 0.0;
		if (T.cost_c) {
			/* objective scaling like the rows': min(1, 100 / ||grad f(x0)||_inf) over the free variables (f = sum c_v x_v^2) */
			__shared__ double redc[32];
			double u[1] = {0.0};
			for (int v = threadIdx.x; v < T.n_all; v += blockDim.x) if (T.fix_src[v] < 0) u[0] = fmax(u[0], fabs(2.0 * T.cost_c[v] * x[v]));
			const int ops[1] = {1};
			block_reduce<1>(u, ops, redc);
			if (threadIdx.x == 0 && u[0] > 100.0) ip[IP_OBJ_SCALE] = fmax(100.0 / u[0], 1e-8);
		}
	}
}

/* ------------------------------------------------------------------ k_jac */

/* every lane of a warp evaluates the same kind of sample.  A dynamics sample is shared by JAC_DYN_SPLIT adjacent lanes: each computes the
 * sample's state (a few hundred flops, the same in every lane) and writes every JAC_DYN_SPLIT-th of its up to 112 columns -- four times the
 * threads for a kernel that is bound by the latency of 22 serial samples per problem, and adjacent lanes store adjacent columns */
#ifndef JAC_DYN_SPLIT
#define JAC_DYN_SPLIT 4
#endif
__global__ void __launch_bounds__(64)
k_jac_dyn(DevTables T, DevWork W, int n)
{
	const int tt = blockIdx.x * blockDim.x + threadIdx.x;
	const int t = tt / JAC_DYN_SPLIT, q = tt - t * JAC_DYN_SPLIT;
	if (t >= n * T.n_dyn) return;
	const int pid = t / T.n_dyn, k = t - pid * T.n_dyn;
	if (W.status[pid] != QTOS_RUNNING) return;
	const DynSample &D = T.dyn[k];
	const Element &E = T.elems[D.elem];
	DynState S; dyn_state(T, D, WS(x, T.n_all), S);
	dyn_jac(T, D, S, WS(sc, T.m) + E.row0, WS(Jv, T.nJ) + E.valoff, E.ncols, q, JAC_DYN_SPLIT);
}

#ifndef JAC_ROM_SPLIT
#define JAC_ROM_SPLIT 4      /* adjacent lanes per range-of-motion sample, like JAC_DYN_SPLIT (a sample has ~36 columns): jac 6.10 (1) / 5.29 (2) / 5.00 (4) ms per step */
#endif
__global__ void __launch_bounds__(128)
k_jac_rom(DevTables T, DevWork W, int n)
{
	const int tt = blockIdx.x * blockDim.x + threadIdx.x;
	const int t = tt / JAC_ROM_SPLIT, q = tt - t * JAC_ROM_SPLIT;
	if (t >= n * T.n_rom4) return;
	const int pid = t / T.n_rom4, k = t - pid * T.n_rom4;
	if (W.status[pid] != QTOS_RUNNING) return;
	const RomSample &R = T.rom[k];
	const Element &E = T.elems[R.elem];
	RomState S; rom_state(T, R, WS(x, T.n_all), S);
	rom_jac(T, R, S, WS(sc, T.m) + E.row0, WS(Jv, T.nJ) + E.valoff, E.ncols, q, JAC_ROM_SPLIT);
}

/* qtos_shape.terrain_gradients: the x-dependent Jacobian values of the terrain rows and the force nodes, one thread per task */
__global__ void __launch_bounds__(128)
k_jac_tg(DevTables T, DevWork W, const qtos_problem *probs, const DevHeightfield *hfs, int n_hf, int n)
{
	const int nt = T.n_ter + T.n_frc;
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n * nt) return;
	const int pid = t / nt, k = t - pid * nt;
	if (W.status[pid] != QTOS_RUNNING) return;
	const int hid = probs[pid].hf_id >= 0 && probs[pid].hf_id < n_hf ? probs[pid].hf_id : 0;
	tg_jac_task(T, hfs[hid], WS(x, T.n_all), WS(sc, T.m), WS(Jv, T.nJ), k);
}

/* ------------------------------------------------------------------ k_prepare */

/* iterations the reference completes within max_cpu_time (`-r`), see qtos_options */
__device__ __forceinline__ int cpu_budget(const qtos_options &opt)
{
	return opt.max_cpu_time > 0.0 ? (int)fmin(1e9, floor(opt.max_cpu_time / QTOS_REF_SECONDS_PER_ITERATION)) : 0x7fffffff;
}

#ifndef JG_U
#define JG_U 2          /* terms in flight per lane of the J' v gather: more costs the fourth resident CTA and loses */
#endif
#ifndef PREP_MINB
#define PREP_MINB 4
#endif
#ifndef STEP_MINB
#define STEP_MINB 4      /* four CTAs per SM: the line search is latency-bound, the fourth CTA pays for ~1.7 KB of spills */
#endif
/* (J' v)_i for variable i (all 32 lanes of a warp call it together, i = 32 g + lane): self-contained term descriptors,
 * read coalesced, four terms in flight; the sum runs in term order, rows ascending */
__device__ __forceinline__ double jt_gather(const DevTables &T, const double *Jv, const double *wrow, int i)
{
	const int g = i >> 5, base = T.jg_ptr[g], ns = (T.jg_ptr[g + 1] - base) >> 5;
	const uint2 *tk = reinterpret_cast<const uint2 *>(T.jg) + base + (i & 31);
	double acc = 0.0;
	for (int s = 0; s < ns; s += JG_U) {
		uint2 d[JG_U];
		double c[JG_U][6], w[JG_U][6];
#pragma unroll
		for (int u = 0; u < JG_U; ++u) d[u] = s + u < ns ? __ldg(tk + 32 * (s + u)) : make_uint2(0u, 0u);
#pragma unroll
		for (int u = 0; u < JG_U; ++u) {
			const int nr = d[u].x >> 20;
			const double *col = Jv + (d[u].x & 0xfffffu), *wr = wrow + d[u].y;
#pragma unroll
			for (int rr = 0; rr < 6; ++rr) { c[u][rr] = rr < nr ? col[rr] : 0.0; w[u][rr] = rr < nr ? wr[rr] : 0.0; }
		}
#pragma unroll
		for (int u = 0; u < JG_U; ++u) {
			const int nr = d[u].x >> 20;
#pragma unroll
			for (int rr = 0; rr < 6; ++rr) if (rr < nr) acc += c[u][rr] * w[u][rr];
		}
	}
	return acc;
}

__global__ void __launch_bounds__(QTOS_THREADS, PREP_MINB)
k_prepare(DevTables T, DevWork W, qtos_options opt, int it)
{
	const int pid = blockIdx.x;
	if (W.status[pid] != QTOS_RUNNING) return;
	__shared__ double red[8 * 32];
	const double *Jv = WS(Jv, T.nJ), *r = WS(r, T.m), *s = WS(s, T.m), *zL = WS(zL, T.m), *zU = WS(zU, T.m);
	const double *dL = WS(dL, T.m), *dU = WS(dU, T.m), *sc = WS(sc, T.m);
	double *y = WS(y, T.m), *Sig = WS(Sig, T.m), *w = WS(w, T.m), *dy = WS(dy, T.m), *vec = WS(vec, T.npad), *scal = WS(scal, 16);
	/* v: 0 dual_inf(max) 1 theta_inf(max) 2 compl max 3 compl min 4 sum|y| 5 sum z 6 viol(max) 7 sum 0*r (NaN/Inf probe:
	 * fmax/fmin drop NaN operands, so a non-finite constraint value would otherwise pass for feasible) */
	double v[8] = {0, 0, 0, 1e300, 0, 0, 0, 0};
	for (int i = threadIdx.x; i < T.npad; i += blockDim.x) v[0] = fmax(v[0], fabs(jt_gather(T, Jv, y, i)));   /* padding variables have no terms */
	for (int i = threadIdx.x; i < T.m; i += blockDim.x) {
		const int fl = T.row_flags[i];
		v[4] += fabs(y[i]);
		v[7] += 0.0 * r[i];
		if (fl & ROW_EQ) { v[1] = fmax(v[1], fabs(r[i])); v[6] = fmax(v[6], fabs(r[i]) / sc[i]); continue; }
		v[1] = fmax(v[1], fabs(r[i] - s[i]));
		v[0] = fmax(v[0], fabs(-y[i] - zL[i] + zU[i]));
		if (fl & ROW_HASL) { const double c = zL[i] * (s[i] - dL[i]); v[2] = fmax(v[2], c); v[3] = fmin(v[3], c); v[5] += zL[i]; v[6] = fmax(v[6], T.gl[i] - r[i] / sc[i]); }
		if (fl & ROW_HASU) { const double c = zU[i] * (dU[i] - s[i]); v[2] = fmax(v[2], c); v[3] = fmin(v[3], c); v[5] += zU[i]; v[6] = fmax(v[6], r[i] / sc[i] - T.gu[i]); }
	}
	const int ops[8] = {1, 1, 1, 2, 0, 0, 1, 0};
	block_reduce<8>(v, ops, red);
	if (!(v[7] == 0.0) || !(v[4] == v[4])) {            /* non-finite constraints or multipliers: stop this window */
		if (threadIdx.x == 0) {
			const double qnan = v[7] - v[7] + (v[4] - v[4]);       /* NaN without a literal */
			scal[SC_DUAL] = scal[SC_THETA] = scal[SC_COMPL] = scal[SC_VIOL] = scal[SC_E0] = qnan;
			W.iters[pid] = it; W.status[pid] = QTOS_INVALID_NUMBER; atomicSub(W.n_running, 1);
		}
		return;
	}
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
		else if (it >= cpu_budget(opt)) { W.status[pid] = QTOS_MAX_CPUTIME; atomicSub(W.n_running, 1); }
		else if (it >= opt.max_iter) { W.status[pid] = QTOS_MAX_ITER; atomicSub(W.n_running, 1); }
	}
	if (conv || it >= opt.max_iter || it >= cpu_budget(opt)) return;
	/* monotone barrier update: max_i |z_i s_i - mu| = max(cmax - mu, mu - cmin) */
	const double kappa_eps = 10.0, mu_min = fmin(opt.tol, opt.compl_inf_tol) / (kappa_eps + 1.0);
	for (;;) {
		const double cm = fmax(cmax - mu, mu - cmin);
		const double Emu = fmax(fmax(dual_inf / s_d, theta_inf), cm / s_c);
		if (Emu <= kappa_eps * mu && mu > mu_min) mu = fmax(mu_min, fmin(0.2 * mu, mu * sqrt(mu)));
		else break;
	}
	if (threadIdx.x == 0) { scal[SC_MU] = mu; W.active[atomicAdd(W.n_active, 1)] = pid; }   /* order is irrelevant: the CTAs of k_asm are independent */
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
	for (int i = threadIdx.x; i < ((T.npad + 31) & ~31); i += blockDim.x) {
		const double gi = jt_gather(T, Jv, w, i);
		if (i < T.npad) vec[i] = i < T.n_free ? -gi : 0.0;
	}
}

/* ------------------------------------------------------------------ k_factor */

#define FT 128           /* threads of the factor kernel: 4 warps, one 8x8 tile of a 16x16 block each */
#ifndef FACTOR_MINB
#define FACTOR_MINB 4    /* resident CTAs per SM (two panel buffers: 54 KB of shared memory per CTA for shape S2) */
#endif
#ifndef ASM_MINB
#define ASM_MINB 6
#endif
#define TLD 18           /* leading dimension of the diagonal block's hand-off tile (rows 4 banks apart for the row-per-lane reads) */
#define TLT 24           /* leading dimension of the off-diagonal S tile: rows 16 banks apart, like the panel (see SWZ) */

/* Blocks of L and the inverses of its diagonal blocks live in global memory in FRAGMENT-MAJOR order: the four
 * contraction values lane (fr, fc) of warp tile tn feeds to its four DMMA steps (k = 4 fc + kk) are stored as two
 * 16-byte chunks, chunk h of all 32 lanes contiguous, so a B operand is two fully coalesced 128-bit loads:
 *   offset(r, c) = (r >> 3) * 128 + ((c >> 1) & 1) * 64 + ((r & 7) * 4 + (c >> 2)) * 2 + (c & 1) */
__device__ __forceinline__ int frag_off(int r, int c)
{
	return ((r >> 3) << 7) + (((c >> 1) & 1) << 6) + ((((r & 7) << 2) + (c >> 2)) << 1) + (c & 1);
}

/* Shared-memory tiles of k_factor (the panel and the off-diagonal S tile) keep rows 16 banks (64 bytes mod 128) apart
 * and store ODD rows with the two 16-byte halves of every 32-byte group swapped (column pair p lives at p ^ 1).
 * Accumulator fragments (one pair per lane, four lanes per row) and operand fragments (two pairs per lane, four lanes
 * per row, 32 bytes apart) then both cover all 32 banks with every quarter-warp: no bank conflicts on either access,
 * and the swap costs nothing at run time (a loop-invariant pointer offset). */
#define SWZ(pair, row) ((pair) ^ ((row) & 1))

/* D = A(8x4) * B(4x8) + C, FP64 tensor-core MMA; fragment layout (PTX ISA, m8n8k4):
 * a = A[lane>>2][lane&3], b = B[lane&3][lane>>2], c/d = C[lane>>2][2*(lane&3) + {0,1}] */
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

/* Assembly of one block row of the condensed KKT matrix A = sigma I + J' D J, one CTA per (problem, block row):
 * stage A_a = D J[:, a] for every (element, column a) of the block row in shared memory, then every warp walks its
 * flat term stream, one term per lane: panel[i(a)][perm(b)] += A_a . J[:, b].  Every panel row belongs to one warp
 * (dealt by term count at shape-compile time, qtos_compile.cpp) and the targets inside one step are distinct, so the sums are
 * race-free and their order is fixed (elements ascending per target).  Terms of an element run in (b, a) order: a step of 32 lanes touches ~8 J columns
 * and the warp's <= 4 staged columns, a third of the L1 wavefronts of the (a, b) order.
 * Block rows are independent, so the grid is problems x block rows and no CTA carries a serial chain; the finished
 * panel (16 x 16 w_I, row-major) goes to the block row's slot of M, where k_factor picks it up. */
__global__ void __launch_bounds__(FT, ASM_MINB)
k_asm(DevTables T, DevWork W, qtos_options opt, int rp_ld)
{
	const int slot = blockIdx.x / T.nb, I = blockIdx.x - slot * T.nb;
	if (slot >= *W.n_active) return;               /* the grid is an upper bound of the active count */
	const int pid = W.active[slot];
	extern __shared__ __align__(16) double sm[];
	double *rp = sm;                               /* [16][rp_ld]  the block row, row-major over the whole panel */
	double *As = rp + 16 * rp_ld;                  /* [as_max][6] staged D J columns of the block row */
	int *av = reinterpret_cast<int *>(As + 6 * T.as_max);         /* [as_max] value offsets of the staged columns */
	const double *Jv = WS(Jv, T.nJ), *Sig = WS(Sig, T.m);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int fI = T.fb[I], wI = I - fI + 1, wcols = wI * 16;
	{
		const int s0 = T.as_ptr[I], nst = T.as_ptr[I + 1] - s0;
		/* one thread per staged column: its (up to) six values are three 128-bit loads, all in flight together */
		for (int k = tid; k < nst; k += FT) {
			const AsmCol C = T.as_col[s0 + k];
			const double2 *src = reinterpret_cast<const double2 *>(Jv + C.voff);
			const int n2 = (C.nrows + 1) >> 1;
			double2 v0 = src[0], v1 = make_double2(0.0, 0.0), v2 = v1;
			if (n2 > 1) v1 = src[1];
			if (n2 > 2) v2 = src[2];
			const double *sg = Sig + C.row0;
			double s6[6];
#pragma unroll
			for (int r = 0; r < 6; ++r) s6[r] = r < C.nrows ? sg[r] : 0.0;
			double2 *dst = reinterpret_cast<double2 *>(As + 6 * k);
			dst[0] = make_double2(s6[0] * v0.x, s6[1] * v0.y);
			dst[1] = make_double2(s6[2] * v1.x, s6[3] * v1.y);
			dst[2] = make_double2(s6[4] * v2.x, s6[5] * v2.y);
			av[k] = C.voff;
		}
		for (int r = warp; r < 16; r += FT / 32)
			for (int c = lane; c < wcols; c += 32) rp[r * rp_ld + c] = 0.0;
	}
	__syncthreads();
	{
		const int t0 = T.at_ptr[I * 4 + warp], t1 = T.at_ptr[I * 4 + warp + 1];
		/* four steps in flight: the J columns of all four, then the sums in step order; the descriptors of the NEXT four
		 * steps are fetched one round trip ahead (the loop is bound by its dependent round trips descriptor -> J column,
		 * not by the L1 data pipe: deeper register pipelines cost occupancy and lose) */
		uint32_t dn[4];
#pragma unroll
		for (int u = 0; u < 4; ++u) dn[u] = t0 + lane + 32 * u < t1 ? __ldg(T.at + t0 + lane + 32 * u) : 0u;
		for (int t = t0 + lane; t < t1; t += 128) {
			uint32_t d[4];
			double2 b[4][3];
#pragma unroll
			for (int u = 0; u < 4; ++u) { d[u] = dn[u]; dn[u] = t + 128 + 32 * u < t1 ? __ldg(T.at + t + 128 + 32 * u) : 0u; }
#pragma unroll
			for (int u = 0; u < 4; ++u) {
				const int n2 = d[u] & 3;
				b[u][0] = b[u][1] = b[u][2] = make_double2(0.0, 0.0);
				if (n2) {
					const double2 *B = reinterpret_cast<const double2 *>(Jv + av[(d[u] >> 2) & 511]) - ((d[u] >> 11) & 127) * n2;
					b[u][0] = B[0];
					if (n2 > 1) b[u][1] = B[1];
					if (n2 > 2) b[u][2] = B[2];
				}
			}
#pragma unroll
			for (int u = 0; u < 4; ++u) {
				if (d[u] & 3) {
					const double2 *A = reinterpret_cast<const double2 *>(As + 6 * ((d[u] >> 2) & 511));
					const double2 a0 = A[0], a1 = A[1], a2 = A[2];      /* staged columns are zero padded to 6 rows */
					const double acc = a0.x * b[u][0].x + a1.x * b[u][1].x + a2.x * b[u][2].x;
					const double acc2 = a0.y * b[u][0].y + a1.y * b[u][1].y + a2.y * b[u][2].y;
					rp[d[u] >> 18] += acc + acc2;
				}
				__syncwarp();
			}
		}
	}
	__syncthreads();
	if (tid < 16) {
		const int i = I * 16 + tid;
		double *d = rp + tid * rp_ld + (wI - 1) * 16 + tid;
		*d = i < T.n_free ? *d + (opt.algorithm == QTOS_ALG_IPOPT ? W.ipst[(size_t)pid * IP_N + IP_SIGMA_F] : opt.sigma_w) : 1.0;
	}
	__syncthreads();
	double2 *out = reinterpret_cast<double2 *>(WS(M, T.nM) + (size_t)T.blkptr[I] * 256);
	const int w2 = wcols >> 1;
	for (int q = tid; q < 16 * w2; q += FT) {
		const int r = q / w2, c2 = q - r * w2;
		out[q] = *reinterpret_cast<const double2 *>(rp + r * rp_ld + 2 * c2);
	}
}

/* one assembled block row (16 x 2 w2 doubles, row-major in global memory) into a shared-memory panel, swizzled */
__device__ __forceinline__ void panel_fetch(double *panel, int rp_ld, const double *src, int w2, int warp, int lane)
{
	const double2 *in = reinterpret_cast<const double2 *>(src);
	for (int r = warp; r < 16; r += 4)
		for (int c2 = lane; c2 < w2; c2 += 32) {
			const unsigned dst = (unsigned)__cvta_generic_to_shared(reinterpret_cast<double2 *>(panel + r * rp_ld) + SWZ(c2, r));
			asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(in + r * w2 + c2) : "memory");
		}
}

/* named barriers of k_factor: the four tile warps among themselves, and the two hand-offs with the diagonal warp */
#define FTT 160          /* threads of the factor kernel: four tile warps + one diagonal warp */
__device__ __forceinline__ void tile_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void diag_ready_arrive() { __threadfence_block(); asm volatile("bar.arrive 2, 160;" ::: "memory"); }
__device__ __forceinline__ void diag_ready_wait() { asm volatile("bar.sync 2, 160;" ::: "memory"); }
__device__ __forceinline__ void diag_done_arrive() { __threadfence_block(); asm volatile("bar.arrive 3, 160;" ::: "memory"); }
__device__ __forceinline__ void diag_done_wait() { asm volatile("bar.sync 3, 160;" ::: "memory"); }

/* Factorization and solve of the condensed KKT system of one problem in one CTA (left-looking over block rows),
 * software-pipelined between four TILE warps (one 8x8 tile of a 16x16 block each) and one DIAGONAL warp:
 *   tile warps, block row I:   load the assembled row A[I, fI..I] (k_asm) into shared memory
 *     for J < I   L[I,J] = (A[I,J] - sum_K L[I,K] L[J,K]') inv(L[J,J])'   two DMMA products per 16x16 block
 *     J = I       S = A[I,I] - sum_K L[I,K] L[I,K]' and p = b_I - L[I,<I] z  handed to the diagonal warp,
 *                 then the row of L is stored (fragment-major) and the next row starts at once
 *   diagonal warp, block row I: Cholesky of S in registers + inv(L[I,I]) (by shuffles), inv -> global (the only form
 *                 of the diagonal block anybody needs: later rows and both substitutions multiply by it),
 *                 z_I = inv(L[I,I]) p
 * Row I+1 needs inv(L[I,I]) and z_I only for its LAST off-diagonal block, so the serial 16x16 factorization -- a third
 * of the kernel's critical path when it sat between barriers of all warps -- runs under the sweep of the next row.
 * Then the backward substitution over the finished factor -> dx (tile warps).
 * IPM = 1 (QTOS_ALG_IPOPT): instead of one right-hand side carried through the sweep and the backward substitution, the
 * IP_NRHS right-hand sides of W.RB (limited-memory columns, affine and centering directions) are forward-substituted as ONE
 * MORE BLOCK ROW of the factorization: with RB' (16 x n) appended below A, row nb of L is (L^-1 RB)' -- the same two DMMA
 * products per block as every other row, A operand from a shared-memory ring of the row's last max_w blocks. */
/* LAT = 1 is the same code under a second symbol: launches with at most one active problem per SM use it, and its preferred
 * shared-memory carve-out leaves the rest of the SM's memory to L1, where the rows of L that come back as B operands then hit
 * (one window: 1.86 -> 1.66 ms per solve for S2, 4.61 -> 3.99 for S5); a kernel attribute belongs to the function, so a separate
 * instantiation keeps the full-size launches of other contexts at the maximum carve-out. */
template <int nbuf, int IPM, int LAT = 0>
__global__ void __launch_bounds__(FTT, FACTOR_MINB)
k_factor(DevTables T, DevWork W, qtos_options opt, int rp_ld, int max_w)
{
	const int pid = blockIdx.x;
	if (W.status[pid] != QTOS_RUNNING) return;
	extern __shared__ __align__(16) double sm[];
	double *rp0 = sm;                              /* [nbuf][16][rp_ld]  block rows I (being swept) and I+1 (arriving), row-major panels;
	                                                  nbuf = 1 where two buffers would cost resident CTAs (wide shapes): no prefetch then */
	double *zs = rp0 + nbuf * 16 * rp_ld;          /* [npad] rhs -> z -> dx */
	double *tmp = zs + T.npad;                     /* [16][TLT] S of an off-diagonal block (tile warps) */
	double *dS = tmp + 16 * TLT;                   /* [16][TLD] S of the diagonal block -> inv(L_II) (hand-off buffer) */
	double *part = dS + 16 * TLD;                  /* [16] rhs of the diagonal solves */
	double *pp = part + 16;                        /* [2][16] L[I,<I] z, one half of the columns per tile column */
	double *M = WS(M, T.nM), *Dinv = WS(Dinv, T.nb * 256);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	__shared__ int bad;
	if (tid == 0) bad = 0;
	if (!IPM) for (int i = tid; i < T.npad; i += FTT) zs[i] = W.vec[(size_t)pid * T.npad + i];
	else for (int i = tid; i < T.npad; i += FTT) zs[i] = 0.0;
	if (IPM && tid < 32) pp[tid] = 0.0;
	__syncthreads();
	if (warp == 4) {
		/* ---------------- diagonal warp ---------------- */
		for (int I = 0; I < T.nb; ++I) {
			diag_ready_wait();
			/* Cholesky of the 16x16 diagonal block and its inverse, in registers: lane r (and r + 16)
			 * holds row r; pivots, columns and the rows of L travel by shuffle */
			const int r = lane & 15;
			double a[16], x[16];
#pragma unroll
			for (int c = 0; c < 16; ++c) a[c] = dS[r * TLD + c];
			/* step j finishes column j of L, then entry j of this lane's column of inv(L) by forward substitution of e_r:
			 * x_j = (delta_jr - sum_{k<j} L[j][k] x_k) / L[j][j]; the two are independent instruction streams, so the
			 * substitution's FMA chain (split in two) runs under the latency of the next pivot's rsqrt */
#pragma unroll
			for (int j = 0; j < 16; ++j) {
				double d = __shfl_sync(0xffffffffu, a[j], j);
				if (!(d > 0.0)) { d = 1e-30; if (lane == 0) bad = 1; }
				const double rl = rsqrt(d);
				const double lij = r == j ? d * rl : a[j] * rl;      /* L[r][j] for r >= j */
				a[j] = lij;
#pragma unroll
				for (int k = j + 1; k < 16; ++k) {
					const double lkj = __shfl_sync(0xffffffffu, lij, k);
					a[k] -= lij * lkj;                                  /* only r >= k is read later */
				}
				double s0 = j == r ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
				for (int k = 0; k < j; ++k) {
					const double ljk = __shfl_sync(0xffffffffu, a[k], j);
					if (k & 1) s1 -= ljk * x[k]; else s0 -= ljk * x[k];
				}
				x[j] = j < r ? 0.0 : (s0 + s1) * rl;
			}
			__syncwarp();
			if (lane < 16) {
#pragma unroll
				for (int c = 0; c < 16; ++c) dS[c * TLD + r] = x[c];
			}
			__syncwarp();
			/* inv(L[I,I]) to global (fragment-major) for later rows and the backward solve; z_I */
			for (int q = lane; q < 256; q += 32) Dinv[(size_t)I * 256 + frag_off(q >> 4, q & 15)] = dS[(q >> 4) * TLD + (q & 15)];
			if (lane < 16) part[lane] = zs[I * 16 + lane] - (pp[lane] + pp[16 + lane]);   /* p = b_I - L[I,<I] z */
			__syncwarp();
			if (lane < 16) {
				double v = 0.0;
				for (int q = 0; q <= lane; ++q) v += dS[lane * TLD + q] * part[q];
				zs[I * 16 + lane] = v;
			}
			diag_done_arrive();
		}
	} else {
		/* ---------------- tile warps ---------------- */
		const int tm = warp >> 1, tn = warp & 1;       /* 8x8 tile of the 16x16 block owned by this warp */
		const int fr = lane >> 2, fc = lane & 3;       /* fragment row / column */
		const int odd = fr & 1;
		if (nbuf == 2) {
			panel_fetch(rp0, rp_ld, M, (T.blkptr[1] - T.blkptr[0]) * 8, warp, lane);
			asm volatile("cp.async.commit_group;" ::: "memory");
		}
		for (int I = 0; I < T.nb; ++I) {
			const int fI = T.fb[I], wI = I - fI + 1;
			const int rowbase = T.blkptr[I] * 256;
			double pacc = 0.0;                         /* this lane's share of L[I,<I] z (forward substitution) */
			double *rp = rp0 + (nbuf == 2 ? (I & 1) * 16 * rp_ld : 0);
			tile_sync();                               /* row I-1 has left the other panel buffer */
			if (nbuf == 2) {
				/* next row's assembled panel (k_asm) starts its way from HBM now and lands under this row's sweep:
				 * asynchronous 16-byte copies, no register staging; this row's copies were issued one row ago */
				if (I + 1 < T.nb) panel_fetch(rp0 + ((I + 1) & 1) * 16 * rp_ld, rp_ld, M + (size_t)T.blkptr[I + 1] * 256, (T.blkptr[I + 2] - T.blkptr[I + 1]) * 8, warp, lane);
				asm volatile("cp.async.commit_group;" ::: "memory");
				asm volatile("cp.async.wait_group 1;" ::: "memory");
			} else {
				panel_fetch(rp0, rp_ld, M + (size_t)T.blkptr[I] * 256, wI * 8, warp, lane);
				asm volatile("cp.async.commit_group;" ::: "memory");
				asm volatile("cp.async.wait_group 0;" ::: "memory");
			}
			if (I > 0 && fI == I) diag_done_wait();    /* no off-diagonal block: keep the hand-off in step */
			for (int J = fI; J <= I; ++J) {
				const int K0 = max(fI, T.fb[J]), nK = J - K0;
				/* inv(L[J,J]) for the second product, requested before the sweep of the block so its latency hides under
				 * it (J = I-1 has to wait for the diagonal warp first) */
				double2 i01 = make_double2(0.0, 0.0), i23 = i01;
				if (J < I - 1) {
					const double2 *bi = reinterpret_cast<const double2 *>(Dinv + (size_t)J * 256 + tn * 128) + lane;
					i01 = bi[0]; i23 = bi[32];
				}
				tile_sync();
				/* four independent accumulator chains (one per k-step of a block) instead of one chain of 4 nK MMAs;
				 * DMMA step kk of lane (fr, fc) contracts k = 4 fc + kk, so both operands are 128-bit loads */
				double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0, c4 = 0.0, c5 = 0.0, c6 = 0.0, c7 = 0.0;
				const double2 *a = reinterpret_cast<const double2 *>(rp + (tm * 8 + fr) * rp_ld + (K0 - fI) * 16 + 4 * fc);
				const double2 *a_lo = a + odd, *a_hi = a + (odd ^ 1);       /* columns 4 fc, 4 fc + 1 / 4 fc + 2, 4 fc + 3 */
				if (J < I) {
					const double2 *b = reinterpret_cast<const double2 *>(M + (size_t)(T.blkptr[J] + K0 - T.fb[J]) * 256 + tn * 128) + lane;
					for (int K = 0; K < nK; ++K) {
						const double2 b01 = b[K * 128], b23 = b[K * 128 + 32];   /* plain loads: written earlier in this kernel */
						const double2 a01 = a_lo[K * 8], a23 = a_hi[K * 8];
						dmma(c0, c1, a01.x, b01.x); dmma(c2, c3, a01.y, b01.y);
						dmma(c4, c5, a23.x, b23.x); dmma(c6, c7, a23.y, b23.y);
					}
				} else {
					const double2 *b = reinterpret_cast<const double2 *>(rp + (tn * 8 + fr) * rp_ld + (K0 - fI) * 16 + 4 * fc);
					const double2 *b_lo = b + odd, *b_hi = b + (odd ^ 1);
					for (int K = 0; K < nK; ++K) {
						const double2 b01 = b_lo[K * 8], b23 = b_hi[K * 8];
						const double2 a01 = a_lo[K * 8], a23 = a_hi[K * 8];
						dmma(c0, c1, a01.x, b01.x); dmma(c2, c3, a01.y, b01.y);
						dmma(c4, c5, a23.x, b23.x); dmma(c6, c7, a23.y, b23.y);
					}
				}
				c0 = (c0 + c2) + (c4 + c6); c1 = (c1 + c3) + (c5 + c7);
				/* S = A[I,J] - sum */
				const int cp = SWZ(tn * 4 + fc, fr);       /* this lane's column pair of the block, as stored */
				const double2 cA = reinterpret_cast<const double2 *>(rp + (tm * 8 + fr) * rp_ld + (J - fI) * 16)[cp];
				const double2 sij = make_double2(cA.x - c0, cA.y - c1);
				if (J < I) reinterpret_cast<double2 *>(tmp + (tm * 8 + fr) * TLT)[cp] = sij;
				else *reinterpret_cast<double2 *>(dS + (tm * 8 + fr) * TLD + tn * 8 + 2 * fc) = sij;   /* plain layout for the diagonal warp */
				if (J < I) {
					/* L[I,J] = S * inv(L[J,J])' : second DMMA product, B = inv(L[J,J]) in fragment order from global */
					if (J == I - 1) {                  /* inv(L[I-1,I-1]) and z_{I-1} are needed from here on, not earlier */
						diag_done_wait();
#ifdef FACTOR_INV_FROM_GLOBAL
						const double2 *bi = reinterpret_cast<const double2 *>(Dinv + (size_t)J * 256 + tn * 128) + lane;
						i01 = bi[0]; i23 = bi[32];
#else
						/* the diagonal warp left inv(L[I-1,I-1]) in the hand-off tile as well: this lane's operand fragment (row
						 * tn * 8 + fr, columns 4 fc .. 4 fc + 3) comes from shared memory instead of through L2 -- the load sits on
						 * the critical path of every block row (the tile is rewritten only after the next tile_sync) */
						const double2 *bi = reinterpret_cast<const double2 *>(dS + (tn * 8 + fr) * TLD + 4 * fc);
						i01 = bi[0]; i23 = bi[1];
#endif
					}
					tile_sync();
					const double2 *ta = reinterpret_cast<const double2 *>(tmp + (tm * 8 + fr) * TLT + 4 * fc);
					const double2 t01 = ta[odd], t23 = ta[odd ^ 1];
					double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
					dmma(x0, x1, t01.x, i01.x); dmma(x2, x3, t01.y, i01.y); dmma(x0, x1, t23.x, i23.x); dmma(x2, x3, t23.y, i23.y);
					const double2 lij = make_double2(x0 + x2, x1 + x3);
					reinterpret_cast<double2 *>(rp + (tm * 8 + fr) * rp_ld + (J - fI) * 16)[cp] = lij;
					/* forward substitution fused into the sweep: z_J is final (J <= I-1, hand-off passed above) */
					const double2 zj = *reinterpret_cast<const double2 *>(zs + J * 16 + tn * 8 + 2 * fc);
					pacc += lij.x * zj.x + lij.y * zj.y;
				}
			}
			/* ---- L[I,<I] z of this warp's tile column: four lanes per row ---- */
			pacc += __shfl_xor_sync(0xffffffffu, pacc, 1); pacc += __shfl_xor_sync(0xffffffffu, pacc, 2);
			if (fc == 0) pp[tn * 16 + tm * 8 + fr] = pacc;
			diag_ready_arrive();                       /* S and L[I,<I] z are with the diagonal warp now */
			/* ---- store the off-diagonal blocks of the row of L (fragment-major in global memory) ---- */
			for (int q = tid; q < (wI - 1) * 128; q += 128) {   /* chunk ((block * 2 + tn) * 2 + h) * 32 + lane, 16 bytes each */
				const int r = ((q >> 6) & 1) * 8 + ((q & 31) >> 2);
				reinterpret_cast<double2 *>(M + rowbase)[q] = reinterpret_cast<const double2 *>(rp + r * rp_ld + (q >> 7) * 16)[SWZ(2 * (q & 3) + ((q >> 5) & 1), r)];
			}
		}
		diag_done_wait();                              /* last diagonal block */
		if (IPM) {
			/* ---- block row nb: Pt[J] = (RB[J]' - sum_K Pt[K] L[J,K]') inv(L[J,J])', Pt[J] = 16 right-hand sides x 16 variables.
			 *      Block K of the row lives in ring slot K % max_w of panel buffer 0 (step J reads K in [fb[J], J), fewer than
			 *      max_w blocks, and writes J last); the finished block also goes to W.PB, plain [J][rhs][variable] ---- */
			const double *RB = WS(RB, T.npad * IP_NRHS);
			double *PB = WS(PB, T.npad * IP_NRHS);
			double *rp = rp0;
			const int row = tm * 8 + fr, col = tn * 8 + 2 * fc;
			const int cp = SWZ(tn * 4 + fc, fr);
			const double2 *a = reinterpret_cast<const double2 *>(rp + row * rp_ld + 4 * fc);
			const double2 *a_lo = a + odd, *a_hi = a + (odd ^ 1);
			/* Gram matrix of the row, G = P'P (16 x 16: Q'Q, Q'p and p'p of kip_solve's Woodbury term): one more product per
			 * block while it sits in the ring, B operand = the same block read by rows */
			const double2 *gb = reinterpret_cast<const double2 *>(rp + (tn * 8 + fr) * rp_ld + 4 * fc);
			const double2 *gb_lo = gb + odd, *gb_hi = gb + (odd ^ 1);
			double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0, g4 = 0.0, g5 = 0.0, g6 = 0.0, g7 = 0.0;
			auto gram = [&](int slot) {
				const double2 a01 = a_lo[slot * 8], a23 = a_hi[slot * 8], b01 = gb_lo[slot * 8], b23 = gb_hi[slot * 8];
				dmma(g0, g1, a01.x, b01.x); dmma(g2, g3, a01.y, b01.y);
				dmma(g4, g5, a23.x, b23.x); dmma(g6, g7, a23.y, b23.y);
			};
			for (int J = 0; J < T.nb; ++J) {
				const int K0 = T.fb[J], nK = J - K0;
				const double2 *bi = reinterpret_cast<const double2 *>(Dinv + (size_t)J * 256 + tn * 128) + lane;
				const double2 i01 = bi[0], i23 = bi[32];
				const double2 cA = *reinterpret_cast<const double2 *>(RB + (size_t)(J * 16 + row) * 16 + col);
				tile_sync();                           /* block J-1 of the row is in the ring */
				if (J > 0) gram((J - 1) % max_w);
				double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0, c4 = 0.0, c5 = 0.0, c6 = 0.0, c7 = 0.0;
				const double2 *b = reinterpret_cast<const double2 *>(M + (size_t)T.blkptr[J] * 256 + tn * 128) + lane;
				int slot = K0 % max_w;
				for (int K = 0; K < nK; ++K) {
					const double2 b01 = b[K * 128], b23 = b[K * 128 + 32];
					const double2 a01 = a_lo[slot * 8], a23 = a_hi[slot * 8];
					dmma(c0, c1, a01.x, b01.x); dmma(c2, c3, a01.y, b01.y);
					dmma(c4, c5, a23.x, b23.x); dmma(c6, c7, a23.y, b23.y);
					if (++slot == max_w) slot = 0;
				}
				c0 = (c0 + c2) + (c4 + c6); c1 = (c1 + c3) + (c5 + c7);
				reinterpret_cast<double2 *>(tmp + row * TLT)[cp] = make_double2(cA.x - c0, cA.y - c1);
				tile_sync();
				const double2 *ta = reinterpret_cast<const double2 *>(tmp + row * TLT + 4 * fc);
				const double2 t01 = ta[odd], t23 = ta[odd ^ 1];
				double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
				dmma(x0, x1, t01.x, i01.x); dmma(x2, x3, t01.y, i01.y); dmma(x0, x1, t23.x, i23.x); dmma(x2, x3, t23.y, i23.y);
				const double2 pj = make_double2(x0 + x2, x1 + x3);
				reinterpret_cast<double2 *>(rp + row * rp_ld + (J % max_w) * 16)[cp] = pj;
				*reinterpret_cast<double2 *>(PB + (size_t)(J * 16 + row) * 16 + col) = pj;
			}
			tile_sync();
			gram((T.nb - 1) % max_w);
			*reinterpret_cast<double2 *>(WS(G, 256) + row * 16 + col) = make_double2((g0 + g2) + (g4 + g6), (g1 + g3) + (g5 + g7));
			if (tid == 0 && bad) W.flags[pid] |= 1;
			return;
		}
		/* ---- backward substitution: L' x = z.  The factor left L2 long ago (888 resident problems x 0.5 MB), so row I-1
		 *      (inv(L_II) and its off-diagonal blocks, fragment-major, <= max_w blocks = one panel buffer) is copied
		 *      asynchronously into the idle panel buffers while row I is solved ---- */
		auto row_fetch = [&](int I) {
			double *dst = rp0 + (nbuf == 2 ? (I & 1) * 16 * rp_ld : 0);
			const int nch = (I - T.fb[I] + 1) * 128;                 /* 16-byte chunks: block 0 = inv(L_II), then the row */
			const double2 *srow = reinterpret_cast<const double2 *>(M + (size_t)T.blkptr[I] * 256) - 128;
			const double2 *sinv = reinterpret_cast<const double2 *>(Dinv + (size_t)I * 256);
			for (int q = tid; q < nch; q += 128) {
				const unsigned d = (unsigned)__cvta_generic_to_shared(reinterpret_cast<double2 *>(dst) + q);
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(q < 128 ? sinv + q : srow + q) : "memory");
			}
		};
		tile_sync();                                   /* every warp is done with both panel buffers */
		row_fetch(T.nb - 1);
		asm volatile("cp.async.commit_group;" ::: "memory");
		asm volatile("cp.async.wait_group 0;" ::: "memory");
		tile_sync();
		for (int I = T.nb - 1; I >= 0; --I) {
			const int fI = T.fb[I];
			const double *buf = rp0 + (nbuf == 2 ? (I & 1) * 16 * rp_ld : 0);
			if (nbuf == 2) {
				if (I > 0) row_fetch(I - 1);           /* its buffer was last read two rows ago */
				asm volatile("cp.async.commit_group;" ::: "memory");
			}
			if (tid < 16) {
				double v = 0.0;
				for (int q = tid; q < 16; ++q) v += buf[frag_off(q, tid)] * zs[I * 16 + q];
				part[tid] = v;
			}
			tile_sync();
			if (tid < 16) zs[I * 16 + tid] = part[tid];
			for (int c = tid; c < (I - fI) * 16; c += 128) {
				const double *blk = buf + 256 + (c >> 4) * 256 + frag_off(0, c & 15);
				double acc = 0.0;
#pragma unroll
				for (int q = 0; q < 16; ++q) acc += blk[((q >> 3) << 7) + ((q & 7) << 3)] * part[q];
				zs[fI * 16 + c] -= acc;
			}
			if (nbuf == 1 && I > 0) {                  /* one buffer: the next row can only be fetched once this one is done */
				tile_sync();
				row_fetch(I - 1);
				asm volatile("cp.async.commit_group;" ::: "memory");
			}
			asm volatile("cp.async.wait_group 0;" ::: "memory");
			tile_sync();
		}
		for (int i = tid; i < T.npad; i += 128) W.vec[(size_t)pid * T.npad + i] = zs[i];
		if (tid == 0 && bad) W.flags[pid] |= 1;
	}
}

/* ------------------------------------------------------------------ k_step */

/* MINB resident CTAs per SM (four for narrow shapes, three for wide ones; see the launch site): the evaluation inside the line
 * search is latency-bound, occupancy pays more than the spills cost (17.8 -> 9.1 ms per bench step at three, 9.0 at four) */
template <int MINB>
__global__ void __launch_bounds__(QTOS_THREADS, MINB)
k_step(DevTables T, DevWork W, const qtos_problem *probs, const DevHeightfield *hfs, int n_hf, qtos_options opt)
{
	const int pid = blockIdx.x;
	if (W.status[pid] != QTOS_RUNNING) return;
	extern __shared__ __align__(16) double sm[];
	double *b = sm;                       /* [npad] dx from k_factor (permuted order) */
	double *part = b + T.npad;            /* [256] scratch */
	double *red = part + 256;             /* [8*32] */
	const double *Jv = WS(Jv, T.nJ);
	const int tid = threadIdx.x;
	for (int i = tid; i < T.npad; i += blockDim.x) b[i] = W.vec[(size_t)pid * T.npad + i];
	__syncthreads();
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
		for (int a = 0; a < E.ncols; ++a) jdx += Jv[E.valoff + a * E.ld + rr] * b[cols[a]];
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

/* streaming: slot k of the list receives problem src[src_idx[k]] */
__global__ void k_admit(const qtos_problem *src, const int *src_idx, const int *slots, qtos_problem *dst, int n)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k < n) dst[slots[k]] = src[src_idx[k]];
}

/* results of workspace slot `pid` to position `out` of the caller's arrays (batch call: both are the block index) */
__global__ void k_results(DevTables T, DevWork W, qtos_result *res, double *x_out, int n, const int *slots, const int *dst)
{
	if ((int)blockIdx.x >= n) return;
	const int pid = slots ? slots[blockIdx.x] : blockIdx.x, out = dst ? dst[blockIdx.x] : blockIdx.x;
	if (threadIdx.x == 0 && res) {
		const double *scal = WS(scal, 16);
		qtos_result R;
		R.status = W.status[pid] == QTOS_RUNNING ? QTOS_MAX_ITER : W.status[pid]; R.iters = W.iters[pid];
		R.constr_viol = scal[SC_VIOL]; R.dual_inf = scal[SC_DUAL]; R.compl_inf = scal[SC_COMPL]; R.nlp_error = scal[SC_E0]; R.mu = scal[SC_MU];
		R.cost = 0.0;
		res[out] = R;
	}
	if (x_out) for (int i = threadIdx.x; i < T.n_all; i += blockDim.x) x_out[(size_t)out * T.n_all + i] = WS(x, T.n_all)[i];
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

/* rows [row0, row0 + n_rows) of every plan's 1 kHz trajectory */
__global__ void k_csv(DevTables T, const double *x_all, const qtos_problem *probs, double *rows, int n, int row0, int n_rows)
{
	const int pid = blockIdx.y;
	const int lr = blockIdx.x * blockDim.x + threadIdx.x, row = row0 + lr;
	if (lr >= n_rows || pid >= n) return;
	const double *x = x_all + (size_t)pid * T.n_all;
	double *c = rows + ((size_t)pid * n_rows + lr) * QTOS_CSV_COLS;
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

__global__ void k_height_grad(DevHeightfield hf, const double *xy, int n, double *hx_out, double *hy_out)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	double hx, hy;
	qtos_height_grad(hf, xy[2 * i], xy[2 * i + 1], hx, hy);
	hx_out[i] = hx; hy_out[i] = hy;
}

/* ------------------------------------------------------------------ best-plan selection (the one exchange step of the path)
 *
 * A record = 5 doubles (group, not converged, cost, violation, global id).  The winner of a group is the lexicographic minimum of
 * (not converged, cost, violation, id) -- identical on every rank and for every GPU count.  k_records builds the records of this
 * rank's results on the device (non-finite metrics of a status -13 window become +inf); after the all-gather k_select finds the
 * winners: one CTA, one key after the other over the still-tied candidates, per-group minima by 64-bit atomicMin on the
 * order-preserving integer image of the doubles (deterministic: a minimum does not depend on the order of the atomics). */
__device__ __forceinline__ unsigned long long ord_of(double x)
{
	const unsigned long long b = (unsigned long long)__double_as_longlong(x);
	return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord_inv(unsigned long long o)
{
	return __longlong_as_double((long long)((o >> 63) ? (o & 0x7fffffffffffffffull) : ~o));
}

__global__ void k_records(const qtos_result *res, const int *group, long long id0, long long id_stride, int n, double *rec)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const qtos_result R = res[i];
	const double inf = __longlong_as_double(0x7ff0000000000000ll);
	double *r = rec + 5 * (size_t)i;
	r[0] = (double)group[i];
	r[1] = R.status != 0 ? 1.0 : 0.0;
	r[2] = R.cost == R.cost ? R.cost : inf;
	r[3] = R.constr_viol == R.constr_viol ? R.constr_viol : inf;
	r[4] = (double)(id0 + id_stride * i);
}

__global__ void __launch_bounds__(1024)
k_select(const double *rec, int n_rec, int n_groups, unsigned long long *best, unsigned char *tied, long long *winner)
{
	const int tid = threadIdx.x;
	for (int i = tid; i < n_rec; i += blockDim.x) { const int g = (int)rec[5 * (size_t)i]; tied[i] = g >= 0 && g < n_groups; }
	for (int col = 1; col <= 4; ++col) {
		for (int g = tid; g < n_groups; g += blockDim.x) best[g] = ~0ull;
		__threadfence_block(); __syncthreads();
		for (int i = tid; i < n_rec; i += blockDim.x)
			if (tied[i]) atomicMin(&best[(int)rec[5 * (size_t)i]], ord_of(rec[5 * (size_t)i + col]));
		__threadfence_block(); __syncthreads();
		for (int i = tid; i < n_rec; i += blockDim.x)
			if (tied[i] && ord_of(rec[5 * (size_t)i + col]) != best[(int)rec[5 * (size_t)i]]) tied[i] = 0;
		__syncthreads();
	}
	for (int g = tid; g < n_groups; g += blockDim.x) winner[g] = best[g] == ~0ull ? -1 : (long long)ord_inv(best[g]);
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
	eval_jac_block(T, hf, x, sc, Jv);
	__syncthreads();
	double *J = jac_out + (size_t)pid * T.m * T.n_all;
	for (int i = threadIdx.x; i < T.m * T.n_all; i += blockDim.x) J[i] = 0.0;
	__syncthreads();
	for (int e = threadIdx.x; e < T.n_elem; e += blockDim.x) {
		const Element &E = T.elems[e];
		for (int a = 0; a < E.ncols; ++a) {
			const int var = T.var_of_perm[T.elem_cols[E.coloff + a]];
			for (int rr = 0; rr < E.nrows; ++rr) J[(size_t)(E.row0 + rr) * T.n_all + var] = Jv[E.valoff + a * E.ld + rr];
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

#include "qtos_ipopt.cuh"
#include "qtos_capi.inc"
