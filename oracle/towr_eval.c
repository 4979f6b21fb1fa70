/*
 * towr_eval.c -- oracle: constraint values g(x) and dense Jacobian J(x).
 * TEST INFRASTRUCTURE (see towr_oracle.h).  Restates, in plain C:
 *   ref: src/polynomial.cc:47-257, src/spline.cc:48-123, src/node_spline.cc:45-112  (A8)
 *   ref: src/euler_converter.cc:58-310                                               (A9)
 *   ref: src/dynamic_constraint.cc:37-137, src/single_rigid_body_dynamics.cc:76-192  (A10)
 *   ref: src/range_of_motion_constraint.cc:35-109                                    (A11)
 *   ref: src/terrain_constraint.cc:44-108                                            (A12)
 *   ref: src/force_constraint.cc:37-171                                              (A13)
 *   ref: src/swing_constraint.cc:41-108                                              (A14)
 *   ref: src/spline_acc_constraint.cc:34-86                                          (A15)
 * The Jacobian is returned dense (m x n, row-major) plus an optional
 * structural mask reproducing which entries the reference's
 * FillJacobianBlock touches (known answer: 11557 + 20605 non-zeros after the
 * 35 fixed columns are dropped, ref: /root/reference/logs/towr_log.out:40-41).
 */
#include "towr_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define OPT(s, node, d, k) ((s)->opt[(node) * 6 + (d) * 3 + (k)])
#define VAL(s, node, d, k) ((s)->val[(node) * 6 + (d) * 3 + (k)])
enum { kPos = 0, kVel = 1, kAcc = 2 };
enum { X = 0, Y = 1, Z = 2 };

/* ------------------------------------------------------------ set variables */

/* ref: src/nodes_variables.cc:67-75 (SetVariables): every node value mapped
 * to an optimisation index receives x(idx); the rest keep their constants */
static void spline_set(orc_spline *s, const double *x)
{
	for (int i = 0; i < s->n_nodes * 6; ++i)
		if (s->opt[i] >= 0) s->val[i] = x[s->offset + s->opt[i]];
}

/* polynomials of a phase share its duration equally, ref: src/nodes_variables_phase_based.cc:74-84 */
static int polys_in_phase(const orc_spline *s, int phase)
{
	int n = 0;
	for (int i = 0; i < s->n_polys; ++i) n += s->poly_phase[i] == phase;
	return n;
}

/* ref: src/phase_durations.cc:57-66 (SetVariables: the last phase fills up the total time) and
 * src/phase_spline.cc:55-65 (UpdatePolynomialDurations of every observing spline) */
static void schedule_set(orc_problem *p, const double *x)
{
	for (int ee = 0; ee < ORC_NEE; ++ee) {
		if (p->sched_off[ee] < 0) continue;
		const int np = p->n_phases[ee];
		double sum = 0.0;
		for (int k = 0; k < np - 1; ++k) { p->phase_dur[ee][k] = x[p->sched_off[ee] + k]; sum += x[p->sched_off[ee] + k]; }
		p->phase_dur[ee][np - 1] = p->T - sum;
		for (int w = 0; w < 2; ++w) {
			orc_spline *s = w ? &p->ee_force[ee] : &p->ee_motion[ee];
			for (int i = 0; i < s->n_polys; ++i) s->dur[i] = p->phase_dur[ee][s->poly_phase[i]] / polys_in_phase(s, s->poly_phase[i]);
		}
	}
}

void orc_set_x(orc_problem *p, const double *x)
{
	spline_set(&p->base_lin, x); spline_set(&p->base_ang, x);
	for (int ee = 0; ee < ORC_NEE; ++ee) { spline_set(&p->ee_motion[ee], x); spline_set(&p->ee_force[ee], x); }
	schedule_set(p, x);
}

/* ------------------------------------------------------------ spline eval */

/* ref: src/spline.cc:48-79 (GetSegmentID, GetLocalTime) */
static void locate(const orc_spline *s, double t, int *id, double *tl)
{
	const double eps = 1e-10;
	double acc = 0.0; int found = s->n_polys - 1;
	for (int i = 0; i < s->n_polys; ++i) {
		acc += s->dur[i];
		if (acc >= t - eps) { found = i; break; }
	}
	double loc = t;
	for (int i = 0; i < found; ++i) loc -= s->dur[i];
	*id = found; *tl = loc;
}

/* ref: src/polynomial.cc:89-104 (UpdateCoeff) + :47-72 (GetPoint) */
static void poly_point(const orc_spline *s, int id, double t, double p[3], double v[3], double a[3])
{
	const double T = s->dur[id];
	for (int k = 0; k < 3; ++k) {
		const double p0 = VAL(s, id, 0, k), v0 = VAL(s, id, 1, k);
		const double p1 = VAL(s, id + 1, 0, k), v1 = VAL(s, id + 1, 1, k);
		const double A = p0, B = v0;
		const double C = -(3 * (p0 - p1) + T * (2 * v0 + v1)) / pow(T, 2);
		const double D = (2 * (p0 - p1) + T * (v0 + v1)) / pow(T, 3);
		if (p) p[k] = 0.0 + 1.0 * A + t * B + pow(t, 2) * C + pow(t, 3) * D;
		if (v) v[k] = 0.0 + 1 * 1.0 * B + 2 * t * C + 3 * pow(t, 2) * D;
		if (a) a[k] = 0.0 + 2 * 1 * 1.0 * C + 3 * 2 * t * D;
	}
}

static void spline_point(const orc_spline *s, double t, double p[3], double v[3], double a[3])
{
	int id; double tl;
	locate(s, t, &id, &tl);
	poly_point(s, id, tl, p, v, a);
}

/* ref: src/polynomial.cc:140-234; side 0 = start node, 1 = end node */
static double dnode(double T, double t, int dx, int side, int nd)
{
	const double t2 = pow(t, 2), t3 = pow(t, 3), T2 = pow(T, 2), T3 = pow(T, 3);
	if (side == 0) {
		if (dx == kPos) return nd == kPos ? (2 * t3) / T3 - (3 * t2) / T2 + 1 : t - (2 * t2) / T + t3 / T2;
		if (dx == kVel) return nd == kPos ? (6 * t2) / T3 - (6 * t) / T2 : (3 * t2) / T2 - (4 * t) / T + 1;
		return nd == kPos ? (12 * t) / T3 - 6 / T2 : (6 * t) / T2 - 4 / T;
	}
	if (dx == kPos) return nd == kPos ? (3 * t2) / T2 - (2 * t3) / T3 : t3 / T2 - t2 / T;
	if (dx == kVel) return nd == kPos ? (6 * t) / T2 - (6 * t2) / T3 : (3 * t2) / T2 - (2 * t) / T;
	return nd == kPos ? 6 / T2 - (12 * t) / T3 : (6 * t) / T2 - 2 / T;
}

/* one row (dimension) of NodeSpline::GetJacobianWrtNodes: <=4 entries,
 * duplicates (shared stance variable) merged.  ref: src/node_spline.cc:85-111 */
typedef struct { int n; int col[4]; double v[4]; } jrow;

static void spline_jac_row(const orc_spline *s, int id, double tl, int dx, int dim, jrow *r)
{
	r->n = 0;
	for (int side = 0; side < 2; ++side)
		for (int nd = 0; nd < 2; ++nd) {
			int idx = OPT(s, id + side, nd, dim);
			if (idx < 0) continue;
			double val = dnode(s->dur[id], tl, dx, side, nd);
			int k;
			for (k = 0; k < r->n; ++k) if (r->col[k] == s->offset + idx) break;
			if (k == r->n) { r->col[k] = s->offset + idx; r->v[k] = 0.0; r->n++; }
			r->v[k] += val;
		}
}

/* base (all-optimised) spline: dense local 12-vector, col = side*6 + nd*3 + dim */
static void base_jac_local(const orc_spline *s, int id, double tl, int dx, int dim, double out[12])
{
	memset(out, 0, sizeof(double) * 12);
	for (int side = 0; side < 2; ++side)
		for (int nd = 0; nd < 2; ++nd)
			out[side * 6 + nd * 3 + dim] = dnode(s->dur[id], tl, dx, side, nd);
}

/* ref: src/polynomial.cc:236-257 (CubicHermitePolynomial::GetDerivativeOfPosWrtDuration) */
static double dpos_dT(const orc_spline *s, int id, double t, int k)
{
	const double x0 = VAL(s, id, 0, k), v0 = VAL(s, id, 1, k), x1 = VAL(s, id + 1, 0, k), v1 = VAL(s, id + 1, 1, k);
	const double t2 = pow(t, 2), t3 = pow(t, 3), T = s->dur[id], T2 = pow(T, 2), T3 = pow(T, 3), T4 = pow(T, 4);
	return (t3 * (v0 + v1)) / T3 - (t2 * (2 * v0 + v1)) / T2 - (3 * t3 * (2 * x0 - 2 * x1 + T * v0 + T * v1)) / T4
	       + (2 * t2 * (3 * x0 - 3 * x1 + 2 * T * v0 + T * v1)) / T3;
}

/* Jacobian of a foot spline's position at global time t with respect to the foot's phase durations: out[dim][col],
 * col < n_phases - 1.  ref: src/phase_spline.cc:67-93 (GetJacobianOfPosWrtDurations, GetDerivativeOfPosWrtPhaseDuration)
 * and src/phase_durations.cc:126-154 (GetJacobianOfPos): the current phase's duration stretches its polynomials, every
 * earlier duration shifts the spline along the time axis, and in the LAST phase (whose duration is what the others leave)
 * the earlier ones also compress it. */
static void dur_jac(const orc_problem *p, int ee, const orc_spline *s, double t, double out[3][ORC_MAX_PHASES])
{
	const int np = p->n_phases[ee];
	int id; double tl;
	locate(s, t, &id, &tl);
	double vel[3];
	poly_point(s, id, tl, NULL, vel, NULL);
	/* current phase by the PHASE durations (Spline::GetSegmentID on them, ref: src/spline.cc:48-63) */
	int phase = np - 1; { double acc = 0.0; for (int i = 0; i < np; ++i) { acc += p->phase_dur[ee][i]; if (acc >= t - 1e-10) { phase = i; break; } } }
	const double inner = 1.0 / polys_in_phase(s, s->poly_phase[id]);
	int prev = 0; for (int i = 0; i < id; ++i) prev += s->poly_phase[i] == s->poly_phase[id];
	const int in_last = phase == np - 1;
	for (int k = 0; k < 3; ++k) {
		const double dx_dT = inner * (dpos_dT(s, id, tl, k) - prev * vel[k]);
		for (int c = 0; c < np - 1; ++c) {
			double j = 0.0;
			if (!in_last && c == phase) j = dx_dT;
			if (c < phase) { j = -1 * vel[k]; if (in_last) j -= dx_dT; }
			out[k][c] = j;
		}
	}
}

/* ------------------------------------------------------------ Euler ZYX */

/* ref: src/euler_converter.cc:133-148 */
static void euler_M(const double e[3], double M[3][3])
{
	const double y = e[Y], z = e[Z];
	memset(M, 0, sizeof(double) * 9);
	M[0][Y] = -sin(z); M[0][X] = cos(y) * cos(z);
	M[1][Y] = cos(z);  M[1][X] = cos(y) * sin(z);
	M[2][Z] = 1.0;     M[2][X] = -sin(y);
}

/* ref: src/euler_converter.cc:150-166 */
static void euler_Mdot(const double e[3], const double ed[3], double Md[3][3])
{
	const double z = e[Z], zd = ed[Z], y = e[Y], yd = ed[Y];
	memset(Md, 0, sizeof(double) * 9);
	Md[0][Y] = -cos(z) * zd; Md[0][X] = -cos(z) * sin(y) * yd - cos(y) * sin(z) * zd;
	Md[1][Y] = -sin(z) * zd; Md[1][X] = cos(y) * cos(z) * zd - sin(y) * sin(z) * yd;
	Md[2][X] = -cos(y) * yd;
}

/* ref: src/euler_converter.cc:207-221 */
static void euler_R(const double e[3], double R[3][3])
{
	const double x = e[X], y = e[Y], z = e[Z];
	R[0][0] = cos(y) * cos(z); R[0][1] = cos(z) * sin(x) * sin(y) - cos(x) * sin(z); R[0][2] = sin(x) * sin(z) + cos(x) * cos(z) * sin(y);
	R[1][0] = cos(y) * sin(z); R[1][1] = cos(x) * cos(z) + sin(x) * sin(y) * sin(z); R[1][2] = cos(x) * sin(y) * sin(z) - cos(z) * sin(x);
	R[2][0] = -sin(y);         R[2][1] = cos(y) * sin(x);                            R[2][2] = cos(x) * cos(y);
}

typedef struct {
	int id; double tl;
	double e[3], ed[3], edd[3];
	double jp[3][12], jv[3][12], ja[3][12];   /* rows of d(pos|vel|acc)/d(local nodes) */
	double R[3][3], M[3][3], Md[3][3], w[3], wd[3];
} ang_state;

static void ang_eval(const orc_spline *s, double t, ang_state *a)
{
	locate(s, t, &a->id, &a->tl);
	poly_point(s, a->id, a->tl, a->e, a->ed, a->edd);
	for (int d = 0; d < 3; ++d) {
		base_jac_local(s, a->id, a->tl, kPos, d, a->jp[d]);
		base_jac_local(s, a->id, a->tl, kVel, d, a->jv[d]);
		base_jac_local(s, a->id, a->tl, kAcc, d, a->ja[d]);
	}
	euler_R(a->e, a->R); euler_M(a->e, a->M); euler_Mdot(a->e, a->ed, a->Md);
	for (int r = 0; r < 3; ++r) {
		a->w[r] = 0; a->wd[r] = 0;
		for (int c = 0; c < 3; ++c) {
			a->w[r] += a->M[r][c] * a->ed[c];                                   /* :65-69  */
			a->wd[r] += a->Md[r][c] * a->ed[c] + a->M[r][c] * a->edd[c];         /* :78-83  */
		}
	}
}

/* d(row `dim` of M)/d(nodes): out[c] = d M[dim][c]/du.  ref: src/euler_converter.cc:168-199 */
static void dM_du(const ang_state *a, int dim, double out[3][12])
{
	const double y = a->e[Y], z = a->e[Z];
	memset(out, 0, sizeof(double) * 36);
	for (int k = 0; k < 12; ++k) {
		const double jy = a->jp[Y][k], jz = a->jp[Z][k];
		if (dim == X) { out[Y][k] = -cos(z) * jz; out[X][k] = -cos(z) * sin(y) * jy - cos(y) * sin(z) * jz; }
		if (dim == Y) { out[Y][k] = -sin(z) * jz; out[X][k] = cos(y) * cos(z) * jz - sin(y) * sin(z) * jy; }
		if (dim == Z) { out[X][k] = -cos(y) * jy; }
	}
}

/* ref: src/euler_converter.cc:270-304 */
static void dMdot_du(const ang_state *a, int dim, double out[3][12])
{
	const double z = a->e[Z], zd = a->ed[Z], y = a->e[Y], yd = a->ed[Y];
	memset(out, 0, sizeof(double) * 36);
	for (int k = 0; k < 12; ++k) {
		const double jz = a->jp[Z][k], jy = a->jp[Y][k], jzd = a->jv[Z][k], jyd = a->jv[Y][k];
		if (dim == X) {
			out[Y][k] = sin(z) * zd * jz - cos(z) * jzd;
			out[X][k] = sin(y) * sin(z) * yd * jz - cos(y) * sin(z) * jzd - cos(y) * cos(z) * yd * jy
			          - cos(y) * cos(z) * zd * jz - cos(z) * sin(y) * jyd + sin(y) * sin(z) * jy * zd;
		}
		if (dim == Y) {
			out[Y][k] = -sin(z) * jzd - cos(z) * zd * jz;
			out[X][k] = cos(y) * cos(z) * jzd - sin(y) * sin(z) * jyd - cos(y) * sin(z) * yd * jy
			          - cos(z) * sin(y) * yd * jz - cos(z) * sin(y) * jy * zd - cos(y) * sin(z) * zd * jz;
		}
		if (dim == Z) out[X][k] = sin(y) * yd * jy - cos(y) * jyd;
	}
}

/* ref: src/euler_converter.cc:85-108 */
static void dangvel_du(const ang_state *a, double out[3][12])
{
	double dM[3][12];
	for (int dim = 0; dim < 3; ++dim) {
		dM_du(a, dim, dM);
		for (int k = 0; k < 12; ++k) {
			double s = 0.0;
			for (int c = 0; c < 3; ++c) s += a->ed[c] * dM[c][k];
			for (int c = 0; c < 3; ++c) s += a->M[dim][c] * a->jv[c][k];
			out[dim][k] = s;
		}
	}
}

/* ref: src/euler_converter.cc:110-131 */
static void dangacc_du(const ang_state *a, double out[3][12])
{
	double dM[3][12], dMd[3][12];
	for (int dim = 0; dim < 3; ++dim) {
		dMdot_du(a, dim, dMd); dM_du(a, dim, dM);
		for (int k = 0; k < 12; ++k) {
			double s = 0.0;
			for (int c = 0; c < 3; ++c) s += a->ed[c] * dMd[c][k];
			for (int c = 0; c < 3; ++c) s += a->Md[dim][c] * a->jv[c][k];
			for (int c = 0; c < 3; ++c) s += a->edd[c] * dM[c][k];
			for (int c = 0; c < 3; ++c) s += a->M[dim][c] * a->ja[c][k];
			out[dim][k] = s;
		}
	}
}

/* ref: src/euler_converter.cc:241-268 */
static void dR_du(const ang_state *a, double Rd[3][3][12])
{
	const double x = a->e[X], y = a->e[Y], z = a->e[Z];
	for (int k = 0; k < 12; ++k) {
		const double jx = a->jp[X][k], jy = a->jp[Y][k], jz = a->jp[Z][k];
		Rd[X][X][k] = -cos(z) * sin(y) * jy - cos(y) * sin(z) * jz;
		Rd[X][Y][k] = sin(x) * sin(z) * jx - cos(x) * cos(z) * jz - sin(x) * sin(y) * sin(z) * jz + cos(x) * cos(z) * sin(y) * jx + cos(y) * cos(z) * sin(x) * jy;
		Rd[X][Z][k] = cos(x) * sin(z) * jx + cos(z) * sin(x) * jz - cos(z) * sin(x) * sin(y) * jx - cos(x) * sin(y) * sin(z) * jz + cos(x) * cos(y) * cos(z) * jy;
		Rd[Y][X][k] = cos(y) * cos(z) * jz - sin(y) * sin(z) * jy;
		Rd[Y][Y][k] = cos(x) * sin(y) * sin(z) * jx - cos(x) * sin(z) * jz - cos(z) * sin(x) * jx + cos(y) * sin(x) * sin(z) * jy + cos(z) * sin(x) * sin(y) * jz;
		Rd[Y][Z][k] = sin(x) * sin(z) * jz - cos(x) * cos(z) * jx - sin(x) * sin(y) * sin(z) * jx + cos(x) * cos(y) * sin(z) * jy + cos(x) * cos(z) * sin(y) * jz;
		Rd[Z][X][k] = -cos(y) * jy;
		Rd[Z][Y][k] = cos(x) * cos(y) * jx - sin(x) * sin(y) * jy;
		Rd[Z][Z][k] = -cos(y) * sin(x) * jx - cos(x) * sin(y) * jy;
	}
}

/* ref: src/euler_converter.cc:223-239 (DerivOfRotVecMult) */
static void drotvec_du(const ang_state *a, const double v[3], int inverse, double out[3][12])
{
	double Rd[3][3][12];
	dR_du(a, Rd);
	memset(out, 0, sizeof(double) * 36);
	for (int row = 0; row < 3; ++row)
		for (int col = 0; col < 3; ++col)
			for (int k = 0; k < 12; ++k)
				out[row][k] += v[col] * (inverse ? Rd[col][row][k] : Rd[row][col][k]);
}

/* ------------------------------------------------------------ small algebra */
static void mat3_mul(const double A[3][3], const double B[3][3], double C[3][3])
{
	for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
		double s = 0; for (int k = 0; k < 3; ++k) s += A[i][k] * B[k][j]; C[i][j] = s; }
}
static void mat3_T(const double A[3][3], double B[3][3])
{ for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) B[i][j] = A[j][i]; }
static void mat3_vec(const double A[3][3], const double v[3], double o[3])
{ for (int i = 0; i < 3; ++i) o[i] = A[i][0] * v[0] + A[i][1] * v[1] + A[i][2] * v[2]; }
static void cross3(const double a[3], const double b[3], double o[3])
{ o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; }
/* ref: src/single_rigid_body_dynamics.cc:47-55 (Cross) */
static void crossmat(const double v[3], double C[3][3])
{
	memset(C, 0, sizeof(double) * 9);
	C[0][1] = -v[2]; C[0][2] = v[1];
	C[1][0] = v[2];  C[1][2] = -v[0];
	C[2][0] = -v[1]; C[2][1] = v[0];
}
static void mat3_mul12(const double A[3][3], const double B[3][12], double C[3][12])
{
	for (int i = 0; i < 3; ++i) for (int k = 0; k < 12; ++k)
		C[i][k] = A[i][0] * B[0][k] + A[i][1] * B[1][k] + A[i][2] * B[2][k];
}

/* ------------------------------------------------------------ phase helpers */
static int is_const_node(const orc_spline *s, int node)
{
	if (node == 0) return s->poly_const[0];
	if (node == s->n_nodes - 1) return s->poly_const[s->n_polys - 1];
	return s->poly_const[node - 1] || s->poly_const[node];
}
/* ref: src/nodes_variables_phase_based.cc:121-141 */
static int node_at_start_of_phase(const orc_spline *s, int phase)
{
	for (int i = 0; i < s->n_polys; ++i) if (s->poly_phase[i] == phase) return i;
	return 0;
}
/* ref: src/nodes_variables_phase_based.cc:113-119 (GetPhase of a non-constant node) */
static int phase_of_node(const orc_spline *s, int node)
{
	int poly = node == 0 ? 0 : node - 1;
	return s->poly_phase[poly];
}

/* ------------------------------------------------------------ g(x) */

typedef struct {
	double com[3], com_acc[3];
	ang_state ang;
	double f[ORC_NEE][3], pee[ORC_NEE][3];
	double Iw[3][3];
} dyn_state;

/* ref: src/dynamic_constraint.cc:118-137 (UpdateModel) */
static void dyn_update(orc_problem *p, double t, dyn_state *d)
{
	spline_point(&p->base_lin, t, d->com, NULL, d->com_acc);
	ang_eval(&p->base_ang, t, &d->ang);
	for (int ee = 0; ee < ORC_NEE; ++ee) {
		spline_point(&p->ee_force[ee], t, d->f[ee], NULL, NULL);
		spline_point(&p->ee_motion[ee], t, d->pee[ee], NULL, NULL);
	}
	double Ib[3][3], RI[3][3], Rt[3][3];
	memcpy(Ib, p->shape.I_b, sizeof(Ib));
	mat3_mul(d->ang.R, Ib, RI); mat3_T(d->ang.R, Rt); mat3_mul(RI, Rt, d->Iw);
}


/* contact basis at a foothold, ref: src/height_map.cc:95-141 (GetNormalizedBasis of Normal, Tangent1, Tangent2); with the
 * reference's zero height derivatives this is (ez, ex, ey) */
static void contact_basis(const orc_problem *p, double x, double y, double n[3], double t1[3], double t2[3])
{
	double hx = 0.0, hy = 0.0;
	if (p->shape.terrain_gradients) orc_height_deriv(&p->hf, x, y, &hx, &hy);
	const double nn = sqrt(hx * hx + hy * hy + 1.0), n1 = sqrt(1.0 + hx * hx), n2 = sqrt(1.0 + hy * hy);
	n[0] = -hx / nn; n[1] = -hy / nn; n[2] = 1.0 / nn;
	t1[0] = 1.0 / n1; t1[1] = 0.0; t1[2] = hx / n1;
	t2[0] = 0.0; t2[1] = 1.0 / n2; t2[2] = hy / n2;
}

/* One NodeCost term: weight * value^2 summed over the NODES of a spline (ref: src/node_cost.cc:53-63), gradient
 * 2 * weight * value added to the node's variable once per node that maps to it (ref: src/node_cost.cc:66-83). */
static double node_cost(const orc_spline *s, int deriv, int dim, double weight, double *grad)
{
	double c = 0.0;
	for (int nd = 0; nd < s->n_nodes; ++nd) {
		const double v = VAL(s, nd, deriv, dim);
		c += weight * pow(v, 2);
		if (grad && OPT(s, nd, deriv, dim) >= 0) grad[s->offset + OPT(s, nd, deriv, dim)] += weight * 2.0 * v;
	}
	return c;
}

/* ref: src/nlp_formulation.cc:343-376 (GetCost: ForcesCostID -> force z, EEMotionCostID -> foot velocity x and y) */
double orc_eval_cost(orc_problem *p, const double *x, double *grad)
{
	orc_set_x(p, x);
	if (grad) memset(grad, 0, sizeof(double) * p->n);
	const orc_shape *sh = &p->shape;
	double f = 0.0;
	for (int ee = 0; ee < ORC_NEE; ++ee) {
		if (sh->cost_force_z != 0.0) f += node_cost(&p->ee_force[ee], 0, Z, sh->cost_force_z, grad);
		if (sh->cost_ee_vel_xy != 0.0) {
			f += node_cost(&p->ee_motion[ee], 1, X, sh->cost_ee_vel_xy, grad);
			f += node_cost(&p->ee_motion[ee], 1, Y, sh->cost_ee_vel_xy, grad);
		}
	}
	return f;
}

void orc_eval_g(orc_problem *p, const double *x, double *g)
{
	orc_set_x(p, x);
	const orc_shape *sh = &p->shape;
	const double grav = 9.80665;   /* ref: src/dynamic_model.cc:38 */
	/* terrain, ref: src/terrain_constraint.cc:59-71 */
	for (int ee = 0; ee < ORC_NEE; ++ee) {
		const orc_spline *s = &p->ee_motion[ee];
		for (int nd = 1; nd < s->n_nodes; ++nd)
			g[p->row_terrain[ee] + nd - 1] = VAL(s, nd, 0, Z) - orc_height(&p->hf, VAL(s, nd, 0, X), VAL(s, nd, 0, Y));
	}
	/* dynamic, ref: src/single_rigid_body_dynamics.cc:76-103 */
	for (int k = 0; k < p->n_dyn; ++k) {
		dyn_state d; dyn_update(p, p->t_dyn[k], &d);
		double fsum[3] = {0, 0, 0}, tau[3] = {0, 0, 0};
		for (int ee = 0; ee < ORC_NEE; ++ee) {
			double r[3] = {d.com[0] - d.pee[ee][0], d.com[1] - d.pee[ee][1], d.com[2] - d.pee[ee][2]}, c[3];
			cross3(d.f[ee], r, c);
			for (int i = 0; i < 3; ++i) { tau[i] += c[i]; fsum[i] += d.f[ee][i]; }
		}
		double Iwd[3], Iww[3], wxIw[3];
		mat3_vec(d.Iw, d.ang.wd, Iwd); mat3_vec(d.Iw, d.ang.w, Iww); cross3(d.ang.w, Iww, wxIw);
		double *gk = g + p->row_dynamic + 6 * k;
		for (int i = 0; i < 3; ++i) gk[i] = Iwd[i] + wxIw[i] - tau[i];
		const double gvec[3] = {0.0, 0.0, -sh->mass * grav};
		for (int i = 0; i < 3; ++i) gk[3 + i] = sh->mass * d.com_acc[i] - fsum[i] - gvec[i];
	}
	/* spline acc, ref: src/spline_acc_constraint.cc:48-63 */
	for (int w = 0; w < 2; ++w) {
		const orc_spline *s = w ? &p->base_ang : &p->base_lin;
		int row0 = w ? p->row_acc_ang : p->row_acc_lin;
		for (int j = 0; j < s->n_polys - 1; ++j) {
			double a0[3], a1[3];
			poly_point(s, j, s->dur[j], NULL, NULL, a0);
			poly_point(s, j + 1, 0.0, NULL, NULL, a1);
			for (int d = 0; d < 3; ++d) g[row0 + 3 * j + d] = a0[d] - a1[d];
		}
	}
	/* total duration (optional), ref: src/total_duration_constraint.cc:48-54: the sum EXCLUDES the last phase */
	for (int ee = 0; ee < ORC_NEE; ++ee) {
		if (p->row_total[ee] < 0) continue;
		double sum = 0.0;
		for (int k = 0; k < p->n_phases[ee] - 1; ++k) sum += x[p->sched_off[ee] + k];
		g[p->row_total[ee]] = sum;
	}
	/* base motion (optional), ref: src/base_motion_constraint.cc:60-66: rows AX AY AZ = Euler angles, LX LY LZ = position */
	for (int k = 0; k < p->n_brom; ++k) {
		double b[3], e[3];
		spline_point(&p->base_lin, p->t_brom[k], b, NULL, NULL);
		spline_point(&p->base_ang, p->t_brom[k], e, NULL, NULL);
		for (int d = 0; d < 3; ++d) { g[p->row_base_rom + 6 * k + d] = e[d]; g[p->row_base_rom + 6 * k + 3 + d] = b[d]; }
	}
	/* range of motion, ref: src/range_of_motion_constraint.cc:59-69 */
	for (int ee = 0; ee < ORC_NEE; ++ee)
		for (int k = 0; k < p->n_rom; ++k) {
			double t = p->t_rom[k], b[3], pe[3], e[3], R[3][3];
			spline_point(&p->base_lin, t, b, NULL, NULL);
			spline_point(&p->ee_motion[ee], t, pe, NULL, NULL);
			spline_point(&p->base_ang, t, e, NULL, NULL);
			euler_R(e, R);
			double r[3] = {pe[0] - b[0], pe[1] - b[1], pe[2] - b[2]};
			for (int i = 0; i < 3; ++i)
				g[p->row_rom[ee] + 3 * k + i] = R[0][i] * r[0] + R[1][i] * r[1] + R[2][i] * r[2];
		}
	/* force, ref: src/force_constraint.cc:67-93; basis is (n,t1,t2) = (ez,ex,ey) */
	for (int ee = 0; ee < ORC_NEE; ++ee) {
		const orc_spline *s = &p->ee_force[ee];
		int row = p->row_force[ee];
		for (int nd = 0; nd < s->n_nodes; ++nd) {
			if (is_const_node(s, nd)) continue;
			double n[3] = {-0.0, -0.0, 1.0}, t1[3] = {1.0, 0.0, 0.0}, t2[3] = {0.0, 1.0, 0.0};
			if (sh->terrain_gradients) {
				const orc_spline *mo = &p->ee_motion[ee];
				const int en = node_at_start_of_phase(mo, phase_of_node(s, nd));
				contact_basis(p, VAL(mo, en, 0, X), VAL(mo, en, 0, Y), n, t1, t2);
			}
			const double *f = &VAL(s, nd, 0, 0);
			double fn = 0, a1 = 0, a2 = 0, b1 = 0, b2 = 0;
			for (int i = 0; i < 3; ++i) {
				fn += f[i] * n[i];
				a1 += f[i] * (t1[i] - sh->mu * n[i]); a2 += f[i] * (t1[i] + sh->mu * n[i]);
				b1 += f[i] * (t2[i] - sh->mu * n[i]); b2 += f[i] * (t2[i] + sh->mu * n[i]);
			}
			g[row++] = fn; g[row++] = a1; g[row++] = a2; g[row++] = b1; g[row++] = b2;
		}
	}
	/* swing, ref: src/swing_constraint.cc:59-82 */
	for (int ee = 0; ee < ORC_NEE; ++ee) {
		const orc_spline *s = &p->ee_motion[ee];
		int row = p->row_swing[ee];
		for (int nd = 0; nd < s->n_nodes; ++nd) {
			if (is_const_node(s, nd)) continue;
			for (int d = 0; d < 2; ++d) {
				double prev = VAL(s, nd - 1, 0, d), next = VAL(s, nd + 1, 0, d);
				double dist = next - prev;
				g[row++] = VAL(s, nd, 0, d) - (prev + 0.5 * dist);
				g[row++] = VAL(s, nd, 1, d) - dist / sh->t_swing_avg;
			}
		}
	}
}

/* ------------------------------------------------------------ J(x) */

typedef struct { double *J; unsigned char *mask; int n; } jac_out;

static void jadd(jac_out *o, int row, int col, double v)
{
	o->J[(size_t)row * o->n + col] += v;
	if (o->mask) o->mask[(size_t)row * o->n + col] = 1;
}
static void jset(jac_out *o, int row, int col, double v)
{
	o->J[(size_t)row * o->n + col] = v;
	if (o->mask) o->mask[(size_t)row * o->n + col] = 1;
}

/* rows A (3) = C(3x3 cross-type, structural zeros on the diagonal) * jac(3 x n) */
static void add_cross_times_jac(jac_out *o, int row0, const double C[3][3], const jrow jr[3], double sign)
{
	for (int r = 0; r < 3; ++r)
		for (int c = 0; c < 3; ++c) {
			if (c == r) continue;                 /* Cross() never creates the diagonal */
			for (int k = 0; k < jr[c].n; ++k) jadd(o, row0 + r, jr[c].col[k], sign * C[r][c] * jr[c].v[k]);
		}
}

void orc_eval_jac(orc_problem *p, const double *x, double *J, unsigned char *mask)
{
	orc_set_x(p, x);
	const orc_shape *sh = &p->shape;
	jac_out o = {J, mask, p->n};
	memset(J, 0, sizeof(double) * (size_t)p->m * p->n);
	if (mask) memset(mask, 0, (size_t)p->m * p->n);

	/* terrain, ref: src/terrain_constraint.cc:90-108; height derivatives are 0 (F4) */
	for (int ee = 0; ee < ORC_NEE; ++ee) {
		const orc_spline *s = &p->ee_motion[ee];
		for (int nd = 1; nd < s->n_nodes; ++nd) {
			int row = p->row_terrain[ee] + nd - 1;
			double hx = 0.0, hy = 0.0;
			if (p->shape.terrain_gradients) orc_height_deriv(&p->hf, VAL(s, nd, 0, X), VAL(s, nd, 0, Y), &hx, &hy);
			jset(&o, row, s->offset + OPT(s, nd, 0, Z), 1.0);
			jset(&o, row, s->offset + OPT(s, nd, 0, X), -hx);
			jset(&o, row, s->offset + OPT(s, nd, 0, Y), -hy);
		}
	}
	/* dynamic, ref: src/dynamic_constraint.cc:73-116 */
	for (int k = 0; k < p->n_dyn; ++k) {
		const double t = p->t_dyn[k];
		const int rowA = p->row_dynamic + 6 * k, rowL = rowA + 3;
		dyn_state d; dyn_update(p, t, &d);
		int id; double tl;
		/* base-lin, ref: src/single_rigid_body_dynamics.cc:105-121 */
		locate(&p->base_lin, t, &id, &tl);
		jrow jp[3], ja[3];
		for (int dim = 0; dim < 3; ++dim) {
			spline_jac_row(&p->base_lin, id, tl, kPos, dim, &jp[dim]);
			spline_jac_row(&p->base_lin, id, tl, kAcc, dim, &ja[dim]);
		}
		for (int ee = 0; ee < ORC_NEE; ++ee) {
			double C[3][3]; crossmat(d.f[ee], C);
			add_cross_times_jac(&o, rowA, C, jp, -1.0);
		}
		for (int dim = 0; dim < 3; ++dim)
			for (int q = 0; q < ja[dim].n; ++q) jadd(&o, rowL + dim, ja[dim].col[q], sh->mass * ja[dim].v[q]);
		/* base-ang, ref: src/single_rigid_body_dynamics.cc:123-165 */
		{
			const ang_state *a = &d.ang;
			double Ib[3][3], Rt[3][3], RIb[3][3], tmp[3], v11[3], v21[3];
			memcpy(Ib, sh->I_b, sizeof(Ib)); mat3_T(a->R, Rt); mat3_mul(a->R, Ib, RIb);
			double j11[3][12], j12[3][12], j13[3][12], j21[3][12], j22[3][12], j23[3][12], t12[3][12];
			double jacc[3][12], jvel[3][12];
			mat3_vec(Rt, a->wd, tmp); mat3_vec(Ib, tmp, v11);
			drotvec_du(a, v11, 0, j11);
			drotvec_du(a, a->wd, 1, t12); mat3_mul12(RIb, t12, j12);
			dangacc_du(a, jacc); mat3_mul12(d.Iw, jacc, j13);
			mat3_vec(Rt, a->w, tmp); mat3_vec(Ib, tmp, v21);
			drotvec_du(a, v21, 0, j21);
			drotvec_du(a, a->w, 1, t12); mat3_mul12(RIb, t12, j22);
			dangvel_du(a, jvel); mat3_mul12(d.Iw, jvel, j23);
			double Cw[3][3], CIw[3][3], Iww[3], s2[3][12], c1[3][12], c2[3][12];
			crossmat(a->w, Cw); mat3_vec(d.Iw, a->w, Iww); crossmat(Iww, CIw);
			for (int r = 0; r < 3; ++r) for (int q = 0; q < 12; ++q) s2[r][q] = j21[r][q] + j22[r][q] + j23[r][q];
			mat3_mul12(Cw, s2, c1); mat3_mul12(CIw, jvel, c2);
			for (int r = 0; r < 3; ++r)
				for (int q = 0; q < 12; ++q) {
					int col = p->base_ang.offset + (a->id + q / 6) * 6 + q % 6;
					jadd(&o, rowA + r, col, (j11[r][q] + j12[r][q] + j13[r][q]) + (c1[r][q] - c2[r][q]));
				}
		}
		for (int ee = 0; ee < ORC_NEE; ++ee) {
			/* force, ref: src/single_rigid_body_dynamics.cc:167-179 */
			jrow jf[3], je[3];
			locate(&p->ee_force[ee], t, &id, &tl);
			for (int dim = 0; dim < 3; ++dim) spline_jac_row(&p->ee_force[ee], id, tl, kPos, dim, &jf[dim]);
			double r[3] = {d.com[0] - d.pee[ee][0], d.com[1] - d.pee[ee][1], d.com[2] - d.pee[ee][2]}, C[3][3];
			crossmat(r, C);
			add_cross_times_jac(&o, rowA, C, jf, 1.0);      /* -(-Cross(r)*jac_force) */
			for (int dim = 0; dim < 3; ++dim)
				for (int q = 0; q < jf[dim].n; ++q) jadd(&o, rowL + dim, jf[dim].col[q], -jf[dim].v[q]);
			/* ee position, ref: src/single_rigid_body_dynamics.cc:181-192 */
			locate(&p->ee_motion[ee], t, &id, &tl);
			for (int dim = 0; dim < 3; ++dim) spline_jac_row(&p->ee_motion[ee], id, tl, kPos, dim, &je[dim]);
			crossmat(d.f[ee], C);
			add_cross_times_jac(&o, rowA, C, je, 1.0);      /* -(Cross(f)*(-jac_ee_pos)) */
			/* contact schedule, ref: src/dynamic_constraint.cc:108-114: both of the above with the durations' Jacobians */
			if (p->sched_off[ee] >= 0) {
				double jfT[3][ORC_MAX_PHASES], jxT[3][ORC_MAX_PHASES], Cr[3][3];
				dur_jac(p, ee, &p->ee_force[ee], t, jfT); dur_jac(p, ee, &p->ee_motion[ee], t, jxT);
				crossmat(r, Cr);
				for (int c = 0; c < p->n_phases[ee] - 1; ++c) {
					const int col = p->sched_off[ee] + c;
					for (int i = 0; i < 3; ++i) {
						double a = 0.0;
						for (int q = 0; q < 3; ++q) a += Cr[i][q] * jfT[q][c] + C[i][q] * jxT[q][c];
						jadd(&o, rowA + i, col, a);
						jadd(&o, rowL + i, col, -jfT[i][c]);
					}
				}
			}
		}
	}
	/* spline acc, ref: src/spline_acc_constraint.cc:65-80 */
	for (int w = 0; w < 2; ++w) {
		const orc_spline *s = w ? &p->base_ang : &p->base_lin;
		int row0 = w ? p->row_acc_ang : p->row_acc_lin;
		for (int j = 0; j < s->n_polys - 1; ++j)
			for (int dim = 0; dim < 3; ++dim) {
				jrow a0, a1;
				spline_jac_row(s, j, s->dur[j], kAcc, dim, &a0);
				spline_jac_row(s, j + 1, 0.0, kAcc, dim, &a1);
				for (int q = 0; q < a0.n; ++q) jadd(&o, row0 + 3 * j + dim, a0.col[q], a0.v[q]);
				for (int q = 0; q < a1.n; ++q) jadd(&o, row0 + 3 * j + dim, a1.col[q], -a1.v[q]);
			}
	}
	/* base motion (optional), ref: src/base_motion_constraint.cc:76-86: GetJacobianWrtNodes(t, kPos) of both base splines */
	for (int k = 0; k < p->n_brom; ++k)
		for (int w = 0; w < 2; ++w) {
			const orc_spline *s = w ? &p->base_lin : &p->base_ang;
			int id; double tl;
			locate(s, p->t_brom[k], &id, &tl);
			for (int dim = 0; dim < 3; ++dim) {
				jrow jr;
				spline_jac_row(s, id, tl, kPos, dim, &jr);
				for (int q = 0; q < jr.n; ++q) jadd(&o, p->row_base_rom + 6 * k + 3 * w + dim, jr.col[q], jr.v[q]);
			}
		}
	/* range of motion, ref: src/range_of_motion_constraint.cc:85-109 */
	for (int ee = 0; ee < ORC_NEE; ++ee)
		for (int k = 0; k < p->n_rom; ++k) {
			const double t = p->t_rom[k];
			const int row0 = p->row_rom[ee] + 3 * k;
			ang_state a; ang_eval(&p->base_ang, t, &a);
			double b[3], pe[3];
			spline_point(&p->base_lin, t, b, NULL, NULL);
			spline_point(&p->ee_motion[ee], t, pe, NULL, NULL);
			int id; double tl;
			jrow jb[3], je[3];
			locate(&p->base_lin, t, &id, &tl);
			for (int dim = 0; dim < 3; ++dim) spline_jac_row(&p->base_lin, id, tl, kPos, dim, &jb[dim]);
			locate(&p->ee_motion[ee], t, &id, &tl);
			for (int dim = 0; dim < 3; ++dim) spline_jac_row(&p->ee_motion[ee], id, tl, kPos, dim, &je[dim]);
			for (int r = 0; r < 3; ++r)
				for (int c = 0; c < 3; ++c) {            /* b_R_w[r][c] = R[c][r], stored dense */
					for (int q = 0; q < jb[c].n; ++q) jadd(&o, row0 + r, jb[c].col[q], -1 * a.R[c][r] * jb[c].v[q]);
					for (int q = 0; q < je[c].n; ++q) jadd(&o, row0 + r, je[c].col[q], a.R[c][r] * je[c].v[q]);
				}
			if (p->sched_off[ee] >= 0) {                         /* ref: src/range_of_motion_constraint.cc:107-109 */
				double jxT[3][ORC_MAX_PHASES];
				dur_jac(p, ee, &p->ee_motion[ee], t, jxT);
				for (int c = 0; c < p->n_phases[ee] - 1; ++c)
					for (int r = 0; r < 3; ++r)
						jadd(&o, row0 + r, p->sched_off[ee] + c, a.R[0][r] * jxT[0][c] + a.R[1][r] * jxT[1][c] + a.R[2][r] * jxT[2][c]);
			}
			double rW[3] = {pe[0] - b[0], pe[1] - b[1], pe[2] - b[2]}, dr[3][12];
			drotvec_du(&a, rW, 1, dr);
			for (int r = 0; r < 3; ++r)
				for (int q = 0; q < 12; ++q) {
					/* row X of the inverse product only ever sees d/d(pitch), d/d(yaw) */
					if (r == X && q % 3 == X) continue;
					jadd(&o, row0 + r, p->base_ang.offset + (a.id + q / 6) * 6 + q % 6, dr[r][q]);
				}
		}
	/* force, ref: src/force_constraint.cc:110-171 */
	for (int ee = 0; ee < ORC_NEE; ++ee) {
		const orc_spline *s = &p->ee_force[ee], *mo = &p->ee_motion[ee];
		int row = p->row_force[ee];
		for (int nd = 0; nd < s->n_nodes; ++nd) {
			if (is_const_node(s, nd)) continue;
			double n[3] = {-0.0, -0.0, 1.0}, t1[3] = {1.0, 0.0, 0.0}, t2[3] = {0.0, 1.0, 0.0};
			if (sh->terrain_gradients) {
				const int en = node_at_start_of_phase(mo, phase_of_node(s, nd));
				contact_basis(p, VAL(mo, en, 0, X), VAL(mo, en, 0, Y), n, t1, t2);
			}
			for (int dim = 0; dim < 3; ++dim) {
				int col = s->offset + OPT(s, nd, 0, dim);
				jset(&o, row + 0, col, n[dim]);
				jset(&o, row + 1, col, t1[dim] - sh->mu * n[dim]);
				jset(&o, row + 2, col, t1[dim] + sh->mu * n[dim]);
				jset(&o, row + 3, col, t2[dim] - sh->mu * n[dim]);
				jset(&o, row + 4, col, t2[dim] + sh->mu * n[dim]);
			}
			int ee_node = node_at_start_of_phase(mo, phase_of_node(s, nd));
			for (int dim = 0; dim < 2; ++dim) {
				int col = mo->offset + OPT(mo, ee_node, 0, dim);
				for (int q = 0; q < 5; ++q) jset(&o, row + q, col, 0.0);   /* f . d(basis) == 0 */
			}
			row += 5;
		}
	}
	/* swing, ref: src/swing_constraint.cc:84-108 */
	for (int ee = 0; ee < ORC_NEE; ++ee) {
		const orc_spline *s = &p->ee_motion[ee];
		int row = p->row_swing[ee];
		for (int nd = 0; nd < s->n_nodes; ++nd) {
			if (is_const_node(s, nd)) continue;
			for (int d = 0; d < 2; ++d) {
				jset(&o, row, s->offset + OPT(s, nd, 0, d), 1.0);
				jset(&o, row, s->offset + OPT(s, nd + 1, 0, d), -0.5);
				jset(&o, row, s->offset + OPT(s, nd - 1, 0, d), -0.5);
				row++;
				jset(&o, row, s->offset + OPT(s, nd, 1, d), 1.0);
				jset(&o, row, s->offset + OPT(s, nd + 1, 0, d), -1.0 / sh->t_swing_avg);
				jset(&o, row, s->offset + OPT(s, nd - 1, 0, d), +1.0 / sh->t_swing_avg);
				row++;
			}
		}
	}	/* total duration (optional), ref: src/total_duration_constraint.cc:64-70 */
	for (int ee = 0; ee < ORC_NEE; ++ee)
		if (p->row_total[ee] >= 0)
			for (int c = 0; c < p->n_phases[ee] - 1; ++c) jset(&o, p->row_total[ee], p->sched_off[ee] + c, 1.0);
}
