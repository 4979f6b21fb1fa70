/*
 * qtos_device.cuh -- device-side view of the compiled shape tables and the per-problem
 * workspace, plus the evaluation device functions (spline sampling, SRBD dynamics,
 * range of motion, terrain, linear rows) shared by all kernels.
 *
 * Arithmetic restated for the GPU (not a translation): because phase durations are fixed,
 * every spline sample is a constant linear map of node values (Hermite weights compiled on
 * the host), and every Jacobian block is (d row / d sampled quantity) x (Hermite weight),
 * formed directly in the element's dense column-major block.
 *   spline sampling   ref: solver/towr/src/polynomial.cc:47-257, spline.cc:48-123, node_spline.cc:45-112
 *   Euler ZYX         ref: solver/towr/src/euler_converter.cc:58-310
 *   SRBD dynamics     ref: solver/towr/src/single_rigid_body_dynamics.cc:76-192, dynamic_constraint.cc:73-137
 *   range of motion   ref: solver/towr/src/range_of_motion_constraint.cc:59-109
 *   terrain           ref: solver/towr/src/terrain_constraint.cc:59-108, custom_terrain.cpp:51-159
 */
#ifndef QTOS_DEVICE_CUH_
#define QTOS_DEVICE_CUH_

#include <cstdint>
#include <cuda_runtime.h>
#include "qtos_tables.h"

struct DevHeightfield { const double *h; int nx, ny; double res; };

struct DevTables {
	/* dims */
	int n_all, n_free, npad, m, n_eq, n_ineq, n_bounds, n_dyn, n_rom4, n_lin, n_ter, n_elem;
	int nJ, nb, nM, as_max, csv_rows, max_nodes;
	int n_nodes[10], var_off[11];
	double T, mass, Ib[9], grav;
	/* tables */
	const int16_t *node_var;
	const uint8_t *x0_spline, *x0_deriv, *x0_dim; const int16_t *x0_node; const int8_t *fix_src;
	const int16_t *perm_of_var, *var_of_perm;
	const uint8_t *row_flags; const double *gl, *gu; const int *row_elem;
	const Element *elems; const int16_t *elem_cols; const double *Jconst; const int16_t *jrow;
	const DynSample *dyn; const RomSample *rom; const JCol *jcols;
	const int *lin_row, *lin_ptr; const int16_t *lin_col; const double *lin_val;
	const int *ter_row; const int16_t *ter_var;
	const int *fb, *blkptr, *diag_off;
	const int *as_ptr, *at_ptr; const AsmCol *as_col; const uint32_t *at;
	const int *jg_ptr; const uint2_t *jg;
	const int *jgc_ptr; const uint2_t *jgc; const int16_t *eq_rows, *iq_rows;
	const double *csv_t, *csv_tl; const uint8_t *csv_id;
	const double *dur; int dur_ld;          /* [10][dur_ld] */
	const double *cost_c;                   /* [n_all] objective coefficients (f = sum cost_c[v] x_v^2), nullptr without cost terms */
	const int *tg_ter, *tg_frc;             /* qtos_shape.terrain_gradients: Jacobian tasks of terrain rows [n_ter][8] and force nodes [n_frc][10] */
	int n_frc, mu_pad_;                     /* n_frc = 0 without the option */
	double mu;                              /* friction coefficient (force rows in the contact basis) */
	double nominal[QTOS_NEE][3];
};

/* per-problem workspace: arrays of [n_problems][len] */
struct DevWork {
	double *x, *xt;                         /* [n_all] */
	double *r, *rt, *s, *st, *y, *zL, *zU, *dL, *dU, *Sig, *w, *ds, *dy, *dzL, *dzU, *sc;   /* [m] */
	double *vec, *rx;                       /* [npad] */
	double *P;                              /* [32] */
	double *scal;                           /* [16]: 0 mu, 1 nu, 2 sd, 3 sc_, 4 dual_inf, 5 theta_inf, ... */
	double *Jv;                             /* [nJ] */
	double *M;                              /* [nM] */
	double *Dinv;                           /* [nb*256] inverses of the diagonal blocks of L */
	int *status, *iters, *flags;            /* [1] each */
	int *n_running;                         /* single counter */
	int *active, *n_active;                 /* [n] problems that go on to assemble / factor / step in this iteration, and their count */
	/* QTOS_ALG_IPOPT (qtos_ipopt.cuh) */
	double *ipst;                           /* [IP_N] algorithm state */
	double *glx, *lastx, *gJold, *adx, *cdx;/* [npad] J'y, previous x, J(x_k)'y_{k+1}, affine / centering dx (permuted order) */
	double *lmS, *lmY;                      /* [6][npad] limited-memory pairs (ring) */
	double *RB, *PB;                        /* [nb][16][16] right-hand sides of the factorization's forward substitution, and L^-1 of them */
	double *G;                              /* [16][16] Gram matrix PB' PB */
	double *wA, *wC;                        /* [m] first-pass row weights of the affine / centering right-hand side */
	double *ads, *ady, *advL, *advU, *cds, *cdy, *cdvL, *cdvU;   /* [m] the two directions, row part */
	double *trace;                          /* [QTOS_TRACE_ITERS][QTOS_TRACE_COLS] */
};

/* IPOPT algorithm state per problem (doubles) */
enum { IP_MU = 0, IP_TAU, IP_FREE, IP_MU_MAX, IP_AMU_THMIN, IP_TH_MAX, IP_TH_MIN, IP_SIGMA_W, IP_NPAIRS, IP_SKIPPED, IP_HAVE_LAST,
       IP_NFILTER, IP_SIGMA_F, IP_AVRG, IP_ERR, IP_THETA, IP_GL2, IP_PR2, IP_ALPHA_PR, IP_ALPHA_DU, IP_DNORM, IP_LS, IP_TAG,
       IP_HEAD, IP_ITER, IP_RETRY, IP_DELTA_W, IP_DELTA_LAST, IP_SIGMA_MIN, IP_ITER_BASE, IP_OBJ_SCALE, IP_FVAL, IP_FPHI = 32, IP_FTH = 64, IP_MID = 96,
       IP_NAMU = 240, IP_AMUF = 256, IP_AMUT = 288, IP_N = 320 };   /* IP_AMUF / IP_AMUT: the barrier update's (f, theta) filter, used with cost terms */
#define IP_FILTER_MAX 32
#define IP_LM 6                             /* limited-memory history capacity */
#define IP_NRHS 16                          /* right-hand sides of the factorization: 6 S + 6 Y columns, affine, centering, 2 spare */

enum { SC_MU = 0, SC_NU, SC_SD, SC_SC, SC_DUAL, SC_THETA, SC_COMPL, SC_VIOL, SC_E0, SC_NFAIL, SC_N };

/* ------------------------------------------------------------------ heightfield */

/* CustomTerrain::GetHeight, bit-exact: same operation order as the reference, no FMA contraction.
 * ref: solver/towr/src/custom_terrain.cpp:51-94; custom_terrain.hpp:32-35 (offsets -1,-1,0; z scale 1) */
__device__ __forceinline__ long long qtos_clamp_index(double fl, long long size)
{
	/* static_cast<size_t>(negative) wraps, std::min(.., size-1) then clamps to the LAST cell */
	if (!(fl >= 0.0)) return size - 1;
	if (fl >= (double)(size - 1)) return size - 1;
	return (long long)fl;
}

__device__ __forceinline__ void qtos_height_cell(const DevHeightfield &hf, double x, double y, long long c[4])
{
	const double xf = floor(__ddiv_rn(__dadd_rn(x, 1.0), hf.res));
	const double yf = floor(__ddiv_rn(__dadd_rn(y, 1.0), hf.res));
	c[0] = qtos_clamp_index(xf, hf.nx);
	c[1] = qtos_clamp_index(yf, hf.ny);
	c[2] = c[0] + 1 < hf.nx - 1 ? c[0] + 1 : hf.nx - 1;
	c[3] = c[1] + 1 < hf.ny - 1 ? c[1] + 1 : hf.ny - 1;
}

/* First derivatives of the bilinear surface: the code CustomTerrain::GetHeightDerivWrtX / WrtY carry commented out (they return 0
 * on the reference's path), same cell selection as the height and the reference's operation order, no FMA contraction.
 * ref: solver/towr/src/custom_terrain.cpp:96-156.  Served by qtos_heightfield_gradients; the solver's terrain rows keep the
 * reference's zero derivatives. */
__device__ __forceinline__ void qtos_height_grad(const DevHeightfield &hf, double x, double y, double &hx, double &hy)
{
	long long c[4];
	qtos_height_cell(hf, x, y, c);
	const double res = hf.res;
	const double x0 = __dadd_rn(__dmul_rn((double)c[0], res), -1.0), x1 = __dadd_rn(__dmul_rn((double)c[2], res), -1.0);
	const double y0 = __dadd_rn(__dmul_rn((double)c[1], res), -1.0), y1 = __dadd_rn(__dmul_rn((double)c[3], res), -1.0);
	const double z00 = __ldg(hf.h + c[0] * hf.ny + c[1]), z01 = __ldg(hf.h + c[0] * hf.ny + c[3]);
	const double z10 = __ldg(hf.h + c[2] * hf.ny + c[1]), z11 = __ldg(hf.h + c[2] * hf.ny + c[3]);
	const double s = __ddiv_rn(1.0, __dmul_rn(res, res));
	hx = __dmul_rn(s, __dadd_rn(__dmul_rn(__dadd_rn(-z00, z10), __dsub_rn(y1, y)), __dmul_rn(__dadd_rn(-z01, z11), __dsub_rn(y, y0))));
	hy = __dmul_rn(s, __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(z00, __dsub_rn(x, x1)), __dmul_rn(z10, __dsub_rn(x0, x))),
	                                      __dmul_rn(z01, __dsub_rn(x1, x))), __dmul_rn(z11, __dsub_rn(x, x0))));
}

__device__ __forceinline__ double qtos_height(const DevHeightfield &hf, double x, double y)
{
	long long c[4];
	qtos_height_cell(hf, x, y, c);
	const double res = hf.res;
	const double x0 = __dadd_rn(__dmul_rn((double)c[0], res), -1.0), x1 = __dadd_rn(__dmul_rn((double)c[2], res), -1.0);
	const double y0 = __dadd_rn(__dmul_rn((double)c[1], res), -1.0), y1 = __dadd_rn(__dmul_rn((double)c[3], res), -1.0);
	const double z00 = __ldg(hf.h + c[0] * hf.ny + c[1]), z01 = __ldg(hf.h + c[0] * hf.ny + c[3]);
	const double z10 = __ldg(hf.h + c[2] * hf.ny + c[1]), z11 = __ldg(hf.h + c[2] * hf.ny + c[3]);
	const double s = __ddiv_rn(1.0, __dmul_rn(res, res));
	const double u0 = __dmul_rn(s, __dsub_rn(x1, x)), u1 = __dmul_rn(s, __dsub_rn(x, x0));
	const double w0 = __dadd_rn(__dmul_rn(u0, z00), __dmul_rn(u1, z10));
	const double w1 = __dadd_rn(__dmul_rn(u0, z01), __dmul_rn(u1, z11));
	return __dadd_rn(__dmul_rn(w0, __dsub_rn(y1, y)), __dmul_rn(w1, __dsub_rn(y, y0)));
}

/* ------------------------------------------------------------------ small algebra */

__device__ __forceinline__ void mat3_vec(const double *A, const double *v, double *o)
{
	for (int i = 0; i < 3; ++i) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}
__device__ __forceinline__ void mat3T_vec(const double *A, const double *v, double *o)
{
	for (int i = 0; i < 3; ++i) o[i] = A[i] * v[0] + A[3 + i] * v[1] + A[6 + i] * v[2];
}
__device__ __forceinline__ void cross3(const double *a, const double *b, double *o)
{
	o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}

struct EulerState {
	double sx, cx, sy, cy, sz, cz;
	double R[9], M[9], Md[9];
};

__device__ __forceinline__ void euler_R(EulerState &E, const double *e)
{
	sincos(e[0], &E.sx, &E.cx); sincos(e[1], &E.sy, &E.cy); sincos(e[2], &E.sz, &E.cz);
	const double sx = E.sx, cx = E.cx, sy = E.sy, cy = E.cy, sz = E.sz, cz = E.cz;
	E.R[0] = cy * cz; E.R[1] = cz * sx * sy - cx * sz; E.R[2] = sx * sz + cx * cz * sy;
	E.R[3] = cy * sz; E.R[4] = cx * cz + sx * sy * sz; E.R[5] = cx * sy * sz - cz * sx;
	E.R[6] = -sy;     E.R[7] = cy * sx;                E.R[8] = cx * cy;
}

/* dR/d(roll|pitch|yaw) */
__device__ __forceinline__ void euler_dR(const EulerState &E, int k, double *D)
{
	const double sx = E.sx, cx = E.cx, sy = E.sy, cy = E.cy, sz = E.sz, cz = E.cz;
	if (k == 0) {
		D[0] = 0; D[1] = sx * sz + cx * cz * sy; D[2] = cx * sz - cz * sx * sy;
		D[3] = 0; D[4] = cx * sy * sz - cz * sx; D[5] = -cx * cz - sx * sy * sz;
		D[6] = 0; D[7] = cx * cy;                D[8] = -cy * sx;
	} else if (k == 1) {
		D[0] = -cz * sy; D[1] = cy * cz * sx; D[2] = cx * cy * cz;
		D[3] = -sy * sz; D[4] = cy * sx * sz; D[5] = cx * cy * sz;
		D[6] = -cy;      D[7] = -sx * sy;     D[8] = -cx * sy;
	} else {
		D[0] = -cy * sz; D[1] = -cx * cz - sx * sy * sz; D[2] = cz * sx - cx * sy * sz;
		D[3] = cy * cz;  D[4] = cz * sx * sy - cx * sz;  D[5] = sx * sz + cx * cz * sy;
		D[6] = 0;        D[7] = 0;                       D[8] = 0;
	}
}

__device__ __forceinline__ void euler_M(EulerState &E, const double *ed)
{
	const double sy = E.sy, cy = E.cy, sz = E.sz, cz = E.cz, yd = ed[1], zd = ed[2];
	E.M[0] = cy * cz; E.M[1] = -sz; E.M[2] = 0;
	E.M[3] = cy * sz; E.M[4] = cz;  E.M[5] = 0;
	E.M[6] = -sy;     E.M[7] = 0;   E.M[8] = 1;
	E.Md[0] = -cz * sy * yd - cy * sz * zd; E.Md[1] = -cz * zd; E.Md[2] = 0;
	E.Md[3] = cy * cz * zd - sy * sz * yd;  E.Md[4] = -sz * zd; E.Md[5] = 0;
	E.Md[6] = -cy * yd;                     E.Md[7] = 0;        E.Md[8] = 0;
}

/* dM/d(e_k) (k = 1 pitch, 2 yaw; roll gives 0) -- also equals dMd/d(ed_k) */
__device__ __forceinline__ void euler_dM(const EulerState &E, int k, double *D)
{
	const double sy = E.sy, cy = E.cy, sz = E.sz, cz = E.cz;
	for (int i = 0; i < 9; ++i) D[i] = 0;
	if (k == 1) { D[0] = -sy * cz; D[3] = -sy * sz; D[6] = -cy; }
	else if (k == 2) { D[0] = -cy * sz; D[1] = -cz; D[3] = cy * cz; D[4] = -sz; }
}

/* dMd/d(e_k) */
__device__ __forceinline__ void euler_dMd(const EulerState &E, const double *ed, int k, double *D)
{
	const double sy = E.sy, cy = E.cy, sz = E.sz, cz = E.cz, yd = ed[1], zd = ed[2];
	for (int i = 0; i < 9; ++i) D[i] = 0;
	if (k == 1) { D[0] = -cz * cy * yd + sy * sz * zd; D[3] = -sy * cz * zd - cy * sz * yd; D[6] = sy * yd; }
	else if (k == 2) { D[0] = sz * sy * yd - cy * cz * zd; D[1] = sz * zd; D[3] = -cy * sz * zd - sy * cz * yd; D[4] = -cz * zd; }
}

/* ------------------------------------------------------------------ spline sampling */

__device__ __forceinline__ double node_val(const double *x, int v) { return v >= 0 ? x[v] : 0.0; }

/* value of one dimension of a phase-based spline: sum_q w[q] * node(id + q/2, deriv q%2, dim) */
__device__ __forceinline__ double phase_sample(const DevTables &T, const double *x, int spline, int id, const double *w, int dim)
{
	const int16_t *nv = T.node_var + ((size_t)spline * T.max_nodes + id) * 6;
	return w[0] * node_val(x, nv[dim]) + w[1] * node_val(x, nv[3 + dim]) + w[2] * node_val(x, nv[6 + dim]) + w[3] * node_val(x, nv[9 + dim]);
}

__device__ __forceinline__ double base_sample(const double *xs /* x + var_off + id*6 */, const double *w, int dim)
{
	return w[0] * xs[dim] + w[1] * xs[3 + dim] + w[2] * xs[6 + dim] + w[3] * xs[9 + dim];
}

struct DynState {
	double c[3], cdd[3], e[3], ed[3], edd[3];
	double p[QTOS_NEE][3], f[QTOS_NEE][3];
	EulerState E;
	double w[3], wd[3], Iw[9];
};

__device__ __forceinline__ void dyn_state(const DevTables &T, const DynSample &D, const double *x, DynState &S)
{
	const double *bl = x + T.var_off[0] + D.base_id * 6, *ba = x + T.var_off[1] + D.base_id * 6;
	for (int d = 0; d < 3; ++d) {
		S.c[d] = base_sample(bl, D.W[0], d); S.cdd[d] = base_sample(bl, D.W[2], d);
		S.e[d] = base_sample(ba, D.W[0], d); S.ed[d] = base_sample(ba, D.W[1], d); S.edd[d] = base_sample(ba, D.W[2], d);
	}
	for (int i = 0; i < QTOS_NEE; ++i)
		for (int d = 0; d < 3; ++d) {
			S.p[i][d] = phase_sample(T, x, 2 + i, D.mo_id[i], D.mo_w[i], d);
			S.f[i][d] = phase_sample(T, x, 6 + i, D.fo_id[i], D.fo_w[i], d);
		}
	euler_R(S.E, S.e); euler_M(S.E, S.ed);
	mat3_vec(S.E.M, S.ed, S.w);
	double a[3], b[3];
	mat3_vec(S.E.Md, S.ed, a); mat3_vec(S.E.M, S.edd, b);
	for (int d = 0; d < 3; ++d) S.wd[d] = a[d] + b[d];
	/* I_w = R I_b R' */
	double RI[9];
	for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
		RI[3 * i + j] = S.E.R[3 * i] * T.Ib[j] + S.E.R[3 * i + 1] * T.Ib[3 + j] + S.E.R[3 * i + 2] * T.Ib[6 + j];
	for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
		S.Iw[3 * i + j] = RI[3 * i] * S.E.R[3 * j] + RI[3 * i + 1] * S.E.R[3 * j + 1] + RI[3 * i + 2] * S.E.R[3 * j + 2];
}

/* column descriptor through the read-only path: the loop that consumes it also stores Jacobian values, and a plain
 * load could alias those stores as far as the compiler knows -- every iteration would wait for its own descriptor */
__device__ __forceinline__ JCol ldg_jcol(const JCol *p)
{
	static_assert(sizeof(JCol) == 32, "JCol is two 16-byte words");
	union { int4 q[2]; JCol c; } u;
	u.q[0] = __ldg(reinterpret_cast<const int4 *>(p)); u.q[1] = __ldg(reinterpret_cast<const int4 *>(p) + 1);
	return u.c;
}

/* 6 rows [AX AY AZ LX LY LZ] (ref: single_rigid_body_dynamics.cc:76-103) */
__device__ __forceinline__ void dyn_rows(const DevTables &T, const DynState &S, double *g6)
{
	double Iwd[3], Iww[3], gyro[3], tau[3] = {0, 0, 0}, fs[3] = {0, 0, 0};
	mat3_vec(S.Iw, S.wd, Iwd); mat3_vec(S.Iw, S.w, Iww); cross3(S.w, Iww, gyro);
	for (int i = 0; i < QTOS_NEE; ++i) {
		double r[3] = {S.c[0] - S.p[i][0], S.c[1] - S.p[i][1], S.c[2] - S.p[i][2]}, t[3];
		cross3(S.f[i], r, t);
		for (int d = 0; d < 3; ++d) { tau[d] += t[d]; fs[d] += S.f[i][d]; }
	}
	for (int d = 0; d < 3; ++d) g6[d] = Iwd[d] + gyro[d] - tau[d];
	g6[3] = T.mass * S.cdd[0] - fs[0];
	g6[4] = T.mass * S.cdd[1] - fs[1];
	g6[5] = T.mass * S.cdd[2] - fs[2] - (-T.mass * T.grav);
}

/* one column of d(angular rows)/d(.) given the variation of R, omega and omega_dot */
__device__ __forceinline__ void ang_column(const DevTables &T, const DynState &S, const double *dR /* nullable */,
                                           const double *dw, const double *dwd, const double *Ibu, const double *Ibh,
                                           const double *Iww, double *col)
{
	double a[3], b[3] = {0, 0, 0}, t[3], q[3];
	mat3_vec(S.Iw, dwd, a);                         /* I_w d(wd) */
	mat3_vec(S.Iw, dw, t);                          /* I_w d(w) */
	for (int d = 0; d < 3; ++d) b[d] = t[d];
	if (dR) {
		mat3_vec(dR, Ibu, t); for (int d = 0; d < 3; ++d) a[d] += t[d];        /* dR I_b R' wd */
		mat3T_vec(dR, S.wd, t); mat3_vec(T.Ib, t, q);                          /* R I_b dR' wd */
		mat3_vec(S.E.R, q, t); for (int d = 0; d < 3; ++d) a[d] += t[d];
		mat3_vec(dR, Ibh, t); for (int d = 0; d < 3; ++d) b[d] += t[d];        /* dR I_b R' w */
		mat3T_vec(dR, S.w, t); mat3_vec(T.Ib, t, q);
		mat3_vec(S.E.R, q, t); for (int d = 0; d < 3; ++d) b[d] += t[d];
	}
	cross3(S.w, b, t);
	cross3(dw, Iww, q);
	for (int d = 0; d < 3; ++d) col[d] = a[d] + t[d] + q[d];
}

/* dense column-major element block of one dynamics sample; sc6 = row scales.  Every column is
 * written exactly once (6 contiguous doubles), driven by the host-compiled column descriptors. */
__device__ __forceinline__ void dyn_jac(const DevTables &T, const DynSample &D, const DynState &S, const double *sc6, double *blk, int ncols,
                                        int first = 0, int stride = 1 /* this thread's share of the columns: first, first + stride, ... */)
{
	/* angular rows wrt Euler angle / rate / acceleration */
	double Gp[9], Gv[9], Ga[9];    /* [r*3+k] */
	{
		double u[3], h[3], Ibu[3], Ibh[3], Iww[3], zero[3] = {0, 0, 0};
		mat3T_vec(S.E.R, S.wd, u); mat3T_vec(S.E.R, S.w, h);
		mat3_vec(T.Ib, u, Ibu); mat3_vec(T.Ib, h, Ibh); mat3_vec(S.Iw, S.w, Iww);
		for (int k = 0; k < 3; ++k) {
			double dR[9], dM[9], dMd[9], dw[3], dwd[3], t1[3], t2[3], col[3];
			euler_dR(S.E, k, dR); euler_dM(S.E, k, dM); euler_dMd(S.E, S.ed, k, dMd);
			mat3_vec(dM, S.ed, dw);
			mat3_vec(dMd, S.ed, t1); mat3_vec(dM, S.edd, t2);
			for (int d = 0; d < 3; ++d) dwd[d] = t1[d] + t2[d];
			ang_column(T, S, dR, dw, dwd, Ibu, Ibh, Iww, col);
			for (int r = 0; r < 3; ++r) Gp[3 * r + k] = col[r];
			/* rate: dw = M[:,k], dwd = Md[:,k] + dM_k ed */
			for (int d = 0; d < 3; ++d) { dw[d] = S.E.M[3 * d + k]; dwd[d] = S.E.Md[3 * d + k] + (dM[3 * d] * S.ed[0] + dM[3 * d + 1] * S.ed[1] + dM[3 * d + 2] * S.ed[2]); }
			ang_column(T, S, nullptr, dw, dwd, Ibu, Ibh, Iww, col);
			for (int r = 0; r < 3; ++r) Gv[3 * r + k] = col[r];
			for (int d = 0; d < 3; ++d) dwd[d] = S.E.M[3 * d + k];
			ang_column(T, S, nullptr, zero, dwd, Ibu, Ibh, Iww, col);
			for (int r = 0; r < 3; ++r) Ga[3 * r + k] = col[r];
		}
	}
	/* vectors whose cross-product matrix fills the angular rows: -F (base-lin), f_i (foot), c - p_i (force) */
	double V[9][3];
	for (int d = 0; d < 3; ++d) {
		double F = 0.0;
		for (int i = 0; i < QTOS_NEE; ++i) { F += S.f[i][d]; V[1 + i][d] = S.f[i][d]; V[5 + i][d] = S.c[d] - S.p[i][d]; }
		V[0][d] = -F;
	}
	const double s0 = sc6[0], s1 = sc6[1], s2 = sc6[2], s3 = sc6[3], s4 = sc6[4], s5 = sc6[5];
	const JCol *cols = T.jcols + D.col0;
	for (int sl = first; sl < ncols; sl += stride) {
		const JCol C = ldg_jcol(cols + sl);
		const int k = C.dim;
		double a0, a1, a2, l0 = 0.0, l1 = 0.0, l2 = 0.0;
		if (C.kind == 1) {
			a0 = Gp[k] * C.w[0] + Gv[k] * C.w[1] + Ga[k] * C.w[2];
			a1 = Gp[3 + k] * C.w[0] + Gv[3 + k] * C.w[1] + Ga[3 + k] * C.w[2];
			a2 = Gp[6 + k] * C.w[0] + Gv[6 + k] * C.w[1] + Ga[6 + k] * C.w[2];
		} else {
			/* column k of Cross(v) = (e_k x v) * (-1) = v x e_k */
			const double *v = V[C.kind == 0 ? 0 : (C.kind == 2 ? 1 + C.foot : 5 + C.foot)];
			const double w = C.w[0];
			a0 = (k == 1 ? -v[2] : (k == 2 ? v[1] : 0.0)) * w;
			a1 = (k == 0 ? v[2] : (k == 2 ? -v[0] : 0.0)) * w;
			a2 = (k == 0 ? -v[1] : (k == 1 ? v[0] : 0.0)) * w;
			const double lin = C.kind == 0 ? T.mass * C.w[2] : (C.kind == 3 ? -w : 0.0);
			l0 = k == 0 ? lin : 0.0; l1 = k == 1 ? lin : 0.0; l2 = k == 2 ? lin : 0.0;
		}
		double2 *o = reinterpret_cast<double2 *>(blk + 6 * sl);
		o[0] = make_double2(s0 * a0, s1 * a1); o[1] = make_double2(s2 * a2, s3 * l0); o[2] = make_double2(s4 * l1, s5 * l2);
	}
}

struct RomState { double c[3], e[3], p[3]; EulerState E; };

__device__ __forceinline__ void rom_state(const DevTables &T, const RomSample &R, const double *x, RomState &S)
{
	const double *bl = x + T.var_off[0] + R.base_id * 6, *ba = x + T.var_off[1] + R.base_id * 6;
	for (int d = 0; d < 3; ++d) {
		S.c[d] = base_sample(bl, R.wp, d); S.e[d] = base_sample(ba, R.wp, d);
		S.p[d] = phase_sample(T, x, 2 + R.ee, R.mo_id, R.mo_w, d);
	}
	euler_R(S.E, S.e);
}

__device__ __forceinline__ void rom_rows(const RomState &S, double *g3)
{
	const double r[3] = {S.p[0] - S.c[0], S.p[1] - S.c[1], S.p[2] - S.c[2]};
	mat3T_vec(S.E.R, r, g3);
}

__device__ __forceinline__ void rom_jac(const DevTables &T, const RomSample &R, const RomState &S, const double *sc3, double *blk, int ncols,
                                        int first = 0, int stride = 1 /* this thread's share of the columns */)
{
	const double rW[3] = {S.p[0] - S.c[0], S.p[1] - S.c[1], S.p[2] - S.c[2]};
	double Gp[9];
	for (int k = 0; k < 3; ++k) {
		double dR[9], col[3];
		euler_dR(S.E, k, dR); mat3T_vec(dR, rW, col);
		for (int r = 0; r < 3; ++r) Gp[3 * r + k] = col[r];
	}
	const JCol *cols = T.jcols + R.col0;
	const double s0 = sc3[0], s1 = sc3[1], s2 = sc3[2];     /* before the stores below: no reload per column */
	for (int sl = first; sl < ncols; sl += stride) {          /* column stride 4: 3 rows + zero pad */
		const JCol C = ldg_jcol(cols + sl);
		const int k = C.dim;
		const double w = C.kind == 0 ? -C.w[0] : C.w[0];
		double v0, v1, v2;
		if (C.kind == 1) { v0 = Gp[k]; v1 = Gp[3 + k]; v2 = Gp[6 + k]; }
		else { v0 = S.E.R[3 * k]; v1 = S.E.R[3 * k + 1]; v2 = S.E.R[3 * k + 2]; }
		double2 *o = reinterpret_cast<double2 *>(blk + 4 * sl);
		o[0] = make_double2(s0 * v0 * w, s1 * v1 * w); o[1] = make_double2(s2 * v2 * w, 0.0);
	}
}

/* contact basis at a foothold from the terrain's first derivatives (ref: height_map.cc:95-141, GetNormalizedBasis of Normal,
 * Tangent1, Tangent2); with zero derivatives it is (ez, ex, ey) */
__device__ __forceinline__ void contact_basis(const DevHeightfield &hf, double x, double y, double n[3], double t1[3], double t2[3])
{
	double hx, hy;
	qtos_height_grad(hf, x, y, hx, hy);
	const double nn = sqrt(hx * hx + hy * hy + 1.0), n1 = sqrt(1.0 + hx * hx), n2 = sqrt(1.0 + hy * hy);
	n[0] = -hx / nn; n[1] = -hy / nn; n[2] = 1.0 / nn;
	t1[0] = 1.0 / n1; t1[1] = 0.0; t1[2] = hx / n1;
	t2[0] = 0.0; t2[1] = 1.0 / n2; t2[2] = hy / n2;
}

/* Jacobian values of one terrain row (k < n_ter) or one force node (k - n_ter) with terrain gradients, scaled by sc
 * (ref: terrain_constraint.cc:90-108, force_constraint.cc:110-171) */
__device__ __forceinline__ void tg_jac_task(const DevTables &T, const DevHeightfield &hf, const double *x, const double *sc, double *Jv, int k)
{
	if (k < T.n_ter) {
		const int *t = T.tg_ter + 8 * k;
		const Element &E = T.elems[t[7]];
		double hx, hy;
		qtos_height_grad(hf, x[t[1]], x[t[2]], hx, hy);
		const double s = sc[t[0]];
		double *v = Jv + E.valoff;
		if (t[4] >= 0) v[t[4] * E.ld] = s * -hx;
		if (t[5] >= 0) v[t[5] * E.ld] = s * -hy;
		if (t[6] >= 0) v[t[6] * E.ld] = s;
		return;
	}
	const int *t = T.tg_frc + 10 * (k - T.n_ter);
	const Element &E = T.elems[t[9]];
	double n[3], t1[3], t2[3];
	contact_basis(hf, x[t[4]], x[t[5]], n, t1, t2);
	const double *s = sc + t[0];
	for (int d = 0; d < 3; ++d) {
		if (t[6 + d] < 0) continue;
		double *v = Jv + E.valoff + t[6 + d] * E.ld;
		v[0] = s[0] * n[d];
		v[1] = s[1] * (t1[d] - T.mu * n[d]); v[2] = s[2] * (t1[d] + T.mu * n[d]);
		v[3] = s[3] * (t2[d] - T.mu * n[d]); v[4] = s[4] * (t2[d] + T.mu * n[d]);
	}
}

/* all constraint values g(x) (unscaled), block-cooperative; no sync inside */
__device__ __forceinline__ void eval_g_block(const DevTables &T, const DevHeightfield &hf, const double *x, double *g)
{
	const int n_tasks = T.n_dyn + T.n_rom4 + T.n_lin + T.n_ter + T.n_frc;
	for (int t = threadIdx.x; t < n_tasks; t += blockDim.x) {
		int k = t;
		if (k < T.n_dyn) {
			const DynSample &D = T.dyn[k];
			DynState S; dyn_state(T, D, x, S);
			double g6[6]; dyn_rows(T, S, g6);
			const int r0 = T.elems[D.elem].row0;
			for (int r = 0; r < 6; ++r) g[r0 + r] = g6[r];
			continue;
		}
		k -= T.n_dyn;
		if (k < T.n_rom4) {
			const RomSample &R = T.rom[k];
			RomState S; rom_state(T, R, x, S);
			double g3[3]; rom_rows(S, g3);
			const int r0 = T.elems[R.elem].row0;
			for (int r = 0; r < 3; ++r) g[r0 + r] = g3[r];
			continue;
		}
		k -= T.n_rom4;
		if (k < T.n_lin) {
			double s = 0.0;
			for (int a = T.lin_ptr[k]; a < T.lin_ptr[k + 1]; ++a) s += T.lin_val[a] * x[T.lin_col[a]];
			g[T.lin_row[k]] = s;
			continue;
		}
		k -= T.n_lin;
		if (k < T.n_ter) {
			const int16_t *v = T.ter_var + 3 * k;
			g[T.ter_row[k]] = x[v[2]] - qtos_height(hf, x[v[0]], x[v[1]]);
			continue;
		}
		k -= T.n_ter;
		{
			/* force rows in the contact basis of the foothold (qtos_shape.terrain_gradients; ref: force_constraint.cc:67-93) */
			const int *q = T.tg_frc + 10 * k;
			double n[3], t1[3], t2[3];
			contact_basis(hf, x[q[4]], x[q[5]], n, t1, t2);
			double fn = 0, a1 = 0, a2 = 0, b1 = 0, b2 = 0;
			for (int i = 0; i < 3; ++i) {
				const double f = x[q[1 + i]];
				fn += f * n[i];
				a1 += f * (t1[i] - T.mu * n[i]); a2 += f * (t1[i] + T.mu * n[i]);
				b1 += f * (t2[i] - T.mu * n[i]); b2 += f * (t2[i] + T.mu * n[i]);
			}
			double *o = g + q[0];
			o[0] = fn; o[1] = a1; o[2] = a2; o[3] = b1; o[4] = b2;
		}
	}
}

/* dynamics + range-of-motion element blocks at x, scaled by sc; block-cooperative */
__device__ __forceinline__ void eval_jac_block(const DevTables &T, const DevHeightfield &hf, const double *x, const double *sc, double *Jv)
{
	const int n_tg = T.tg_ter ? T.n_ter + T.n_frc : 0;
	const int n_tasks = T.n_dyn + T.n_rom4 + n_tg;
	for (int t = threadIdx.x; t < n_tasks; t += blockDim.x) {
		if (t >= T.n_dyn + T.n_rom4) { tg_jac_task(T, hf, x, sc, Jv, t - T.n_dyn - T.n_rom4); continue; }
		if (t < T.n_dyn) {
			const DynSample &D = T.dyn[t];
			const Element &E = T.elems[D.elem];
			DynState S; dyn_state(T, D, x, S);
			dyn_jac(T, D, S, sc + E.row0, Jv + E.valoff, E.ncols);
		} else {
			const RomSample &R = T.rom[t - T.n_dyn];
			const Element &E = T.elems[R.elem];
			RomState S; rom_state(T, R, x, S);
			rom_jac(T, R, S, sc + E.row0, Jv + E.valoff, E.ncols);
		}
	}
}

#endif
