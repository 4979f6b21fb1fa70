/*
 * towr_problem.c -- oracle: gait tables, variable sets, initial guess, bounds.
 * TEST INFRASTRUCTURE (see towr_oracle.h).  Restates, in plain C:
 *   ref: src/quadruped_gait_generator.cc:39-369, src/gait_generator.cc:54-150  (A4)
 *   ref: src/parameters.cc:40-135                                               (A5)
 *   ref: src/nlp_formulation.cc:63-198, src/nodes_variables.cc:126-181          (A6)
 *   ref: src/nodes_variables_phase_based.cc:38-58,197-298                       (A7)
 *   ref: src/time_discretization_constraint.cc:41-49 (sample times)
 */
#include "towr_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ gait */

/* contact code: first char hind legs, second char front legs
 * (I none, P left only, b right only, B both)
 * ref: src/quadruped_gait_generator.cc:48-74 */
static void contact_from_code(const char *c, int out[ORC_NEE])
{
	out[0] = out[1] = out[2] = out[3] = 0;
	if (c[0] == 'P' || c[0] == 'B') out[2] = 1;  /* LH */
	if (c[0] == 'b' || c[0] == 'B') out[3] = 1;  /* RH */
	if (c[1] == 'P' || c[1] == 'B') out[0] = 1;  /* LF */
	if (c[1] == 'b' || c[1] == 'B') out[1] = 1;  /* RF */
}

typedef struct { int n; double t[8]; const char *c[8]; } stride;

enum { G_STAND, G_FLIGHT, G_WALK1, G_WALK2, G_WALK2E, G_RUN1, G_RUN2, G_RUN2E,
       G_RUN3, G_RUN3E, G_HOP1, G_HOP1E, G_HOP2, G_HOP3, G_HOP3E, G_HOP5 };

/* ref: src/gait_generator.cc:132-146 (RemoveTransition) */
static stride remove_transition(stride g)
{
	stride r = g;
	r.n = g.n - 1;
	r.t[r.n - 1] += g.t[g.n - 1];
	return r;
}

/* ref: src/quadruped_gait_generator.cc:90-367 */
static stride get_gait(int gait)
{
	stride walk    = {8, {0.3,0.2,0.3,0.2,0.3,0.2,0.3,0.2}, {"bB","BB","Bb","BB","PB","BB","BP","BB"}};
	stride overlap = {8, {0.25,0.13,0.25,0.13,0.25,0.13,0.25,0.13}, {"bB","bb","Bb","Pb","PB","PP","BP","bP"}};
	stride gallop  = {8, {0.2,0.3,0.2,0.2,0.2,0.3,0.2,0.2}, {"Bb","BI","BP","bP","bB","IB","PB","Pb"}};
	switch (gait) {
	case G_STAND:  { stride s = {1, {0.3}, {"BB"}}; return s; }
	case G_FLIGHT: { stride s = {1, {0.3}, {"Bb"}}; return s; }
	case G_WALK1:  return walk;
	case G_WALK2:  return overlap;
	case G_WALK2E: return remove_transition(overlap);
	case G_RUN1:   { stride s = {4, {0.3,0.2,0.3,0.2}, {"bP","BB","Pb","BB"}}; return s; }
	case G_RUN2:   { stride s = {4, {0.4,0.1,0.4,0.1}, {"bP","II","Pb","II"}}; return s; }
	case G_RUN2E:  { stride s = {1, {0.4}, {"bP"}}; return s; }
	case G_RUN3:   { stride s = {4, {0.3,0.1,0.3,0.1}, {"PP","II","bb","II"}}; return s; }
	case G_RUN3E:  { stride s = {1, {0.3}, {"PP"}}; return s; }
	case G_HOP1:   { stride s = {4, {0.3,0.1,0.3,0.1}, {"BI","II","IB","II"}}; return s; }
	case G_HOP1E:  { stride s = {1, {0.3}, {"BI"}}; return s; }
	case G_HOP2:   { stride s = {3, {0.3,0.4,0.3}, {"BB","II","BB"}}; return s; }
	case G_HOP3:   return gallop;
	case G_HOP3E:  return remove_transition(gallop);
	default:       { stride s = {6, {0.1,0.2,0.1,0.1,0.2,0.1}, {"Bb","BB","IP","Bb","BB","IP"}}; return s; }
	}
}

/* ref: src/quadruped_gait_generator.cc:76-88 (SetCombo), src/gait_generator.cc:65-111 */
static int build_gait(int combo, double T, int n_phases[ORC_NEE],
                      double dur[ORC_NEE][ORC_MAX_PHASES], int at_start[ORC_NEE])
{
	static const int combos[6][6] = {
		{G_STAND, G_WALK2, G_WALK2, G_WALK2, G_WALK2E, G_STAND},
		{G_STAND, G_RUN2,  G_RUN2,  G_RUN2,  G_RUN2E,  G_STAND},
		{G_STAND, G_RUN3,  G_RUN3,  G_RUN3,  G_RUN3E,  G_STAND},
		{G_STAND, G_HOP1,  G_HOP1,  G_HOP1,  G_HOP1E,  G_STAND},
		{G_STAND, G_HOP3,  G_HOP3,  G_HOP3,  G_HOP3E,  G_STAND},
		{G_STAND, G_WALK1, G_WALK1, G_WALK1, G_WALK2E, G_STAND}};
	double times[64]; int contacts[64][ORC_NEE]; int np = 0;
	if (combo < 0 || combo > 5) return -1;
	for (int gi = 0; gi < 6; ++gi) {
		stride s = get_gait(combos[combo][gi]);
		for (int k = 0; k < s.n; ++k) {
			times[np] = s.t[k];
			contact_from_code(s.c[k], contacts[np]);
			np++;
		}
	}
	for (int ee = 0; ee < ORC_NEE; ++ee) {
		double acc = 0.0; int n = 0; double raw[ORC_MAX_PHASES];
		for (int ph = 0; ph < np - 1; ++ph) {
			acc += times[ph];
			if (contacts[ph][ee] != contacts[ph + 1][ee]) { raw[n++] = acc; acc = 0.0; }
		}
		raw[n++] = acc + times[np - 1];
		double total = 0.0;
		for (int k = 0; k < n; ++k) total += raw[k];
		for (int k = 0; k < n; ++k) dur[ee][k] = (raw[k] / total) * T;   /* normalise then scale */
		n_phases[ee] = n;
		at_start[ee] = contacts[0][ee];
	}
	return 0;
}

/* ------------------------------------------------------------------ splines */

static void spline_alloc(orc_spline *s, int n_polys)
{
	s->n_polys = n_polys;
	s->n_nodes = n_polys + 1;
	s->dur = (double *)calloc(n_polys, sizeof(double));
	s->opt = (int *)malloc(sizeof(int) * s->n_nodes * 6);
	s->val = (double *)calloc(s->n_nodes * 6, sizeof(double));
	s->poly_phase = (int *)calloc(n_polys, sizeof(int));
	s->poly_const = (int *)calloc(n_polys, sizeof(int));
	for (int i = 0; i < s->n_nodes * 6; ++i) s->opt[i] = -1;
	s->n_vars = 0;
}

static void spline_free(orc_spline *s)
{
	free(s->dur); free(s->opt); free(s->val); free(s->poly_phase); free(s->poly_const);
}

/* ref: src/nodes_variables_all.cc:45-61: idx = node*6 + deriv*3 + dim */
static void build_base_spline(orc_spline *s, const double *durs, int n_polys)
{
	spline_alloc(s, n_polys);
	memcpy(s->dur, durs, sizeof(double) * n_polys);
	for (int i = 0; i < s->n_nodes * 6; ++i) s->opt[i] = i;
	s->n_vars = s->n_nodes * 6;
}

/* ref: src/nodes_variables_phase_based.cc:38-58 (BuildPolyInfos) and :66-75 */
static void build_phase_polys(orc_spline *s, int n_phases, const double *phase_dur,
                              int first_const, int n_polys_changing)
{
	int n_polys = 0, c = first_const;
	for (int i = 0; i < n_phases; ++i) { n_polys += c ? 1 : n_polys_changing; c = !c; }
	spline_alloc(s, n_polys);
	int k = 0; c = first_const;
	for (int i = 0; i < n_phases; ++i) {
		int np = c ? 1 : n_polys_changing;
		for (int j = 0; j < np; ++j) {
			s->poly_phase[k] = i;
			s->poly_const[k] = c;
			s->dur[k] = phase_dur[i] / np;
			k++;
		}
		c = !c;
	}
}

/* ref: src/nodes_variables_phase_based.cc:99-111 */
static int is_const_node(const orc_spline *s, int node)
{
	if (node == 0) return s->poly_const[0];
	if (node == s->n_nodes - 1) return s->poly_const[s->n_polys - 1];
	return s->poly_const[node - 1] || s->poly_const[node];
}

#define OPT(s, node, d, k) ((s)->opt[(node) * 6 + (d) * 3 + (k)])
#define VAL(s, node, d, k) ((s)->val[(node) * 6 + (d) * 3 + (k)])

/* ref: src/nodes_variables_phase_based.cc:197-246 */
static void param_ee_motion(orc_spline *s)
{
	int idx = 0;
	for (int node = 0; node < s->n_nodes; ++node) {
		if (!is_const_node(s, node)) {
			for (int dim = 0; dim < 3; ++dim) {
				OPT(s, node, 0, dim) = idx++;
				if (dim != 2) OPT(s, node, 1, dim) = idx++;   /* vz fixed to 0 */
			}
		} else {
			for (int dim = 0; dim < 3; ++dim) {
				OPT(s, node, 0, dim) = idx;
				OPT(s, node + 1, 0, dim) = idx;
				idx++;
			}
			node += 1;
		}
	}
	s->n_vars = idx;
}

/* ref: src/nodes_variables_phase_based.cc:264-296 */
static void param_ee_force(orc_spline *s)
{
	int idx = 0;
	for (int node = 0; node < s->n_nodes; ++node) {
		if (!is_const_node(s, node)) {
			for (int dim = 0; dim < 3; ++dim) {
				OPT(s, node, 0, dim) = idx++;
				OPT(s, node, 1, dim) = idx++;
			}
		} else {
			node += 1;   /* both swing nodes stay zero */
		}
	}
	s->n_vars = idx;
}

/* ref: src/nodes_variables.cc:126-148 (SetByLinearInterpolation); iterates
 * opt indices ascending, every node mapped to the index gets ITS OWN
 * interpolated value, GetValues() (:57-65) then reports the last one. */
static void set_linear(orc_spline *s, const double *a, const double *b, double T)
{
	for (int node = 0; node < s->n_nodes; ++node)
		for (int dim = 0; dim < 3; ++dim) {
			double dp = b[dim] - a[dim];
			if (OPT(s, node, 0, dim) >= 0)
				VAL(s, node, 0, dim) = a[dim] + node / (double)(s->n_nodes - 1) * dp;
			if (OPT(s, node, 1, dim) >= 0)
				VAL(s, node, 1, dim) = dp / T;
		}
}

/* x0 of one set = GetValues(): for shared indices the later node wins */
static void spline_get_values(const orc_spline *s, double *x)
{
	for (int node = 0; node < s->n_nodes; ++node)
		for (int d = 0; d < 2; ++d)
			for (int dim = 0; dim < 3; ++dim) {
				int i = OPT(s, node, d, dim);
				if (i >= 0) x[s->offset + i] = VAL(s, node, d, dim);
			}
}

static void add_bound(orc_problem *p, const orc_spline *s, int node, int d, int dim, double v)
{
	int i = OPT(s, node, d, dim);
	if (i >= 0) { p->xl[s->offset + i] = v; p->xu[s->offset + i] = v; }
}

/* ------------------------------------------------------------------ public */

void orc_default_shape(orc_shape *s)
{
	memset(s, 0, sizeof(*s));
	s->mass = 1.5;
	/* ctor is (m, Ixx,Iyy,Izz, Ixy,Ixz,Iyz) but is called with the YAML order
	 * (ixx,ixy,ixz,iyy,iyz,izz) -> F5 quirk tensor.
	 * ref: include/towr/models/examples/solo12_model.h:35-37,
	 *      src/single_rigid_body_dynamics.cc:36-44 */
	const double Ixx = 0.00578574, Iyy = 0.0, Izz = 0.0, Ixy = 0.01938108, Ixz = 0.0, Iyz = 0.02476124;
	double I[9] = { Ixx, -Ixy, -Ixz,  -Ixy, Iyy, -Iyz,  -Ixz, -Iyz, Izz };
	memcpy(s->I_b, I, sizeof(I));
	const double xn = 0.21, yn = 0.18, zn = -0.24;
	double nom[4][3] = {{xn, yn, zn}, {xn, -yn, zn}, {-xn, yn, zn}, {-xn, -yn, zn}};
	memcpy(s->nominal, nom, sizeof(nom));
	s->max_dev[0] = 0.10; s->max_dev[1] = 0.08; s->max_dev[2] = 0.10;
	s->mu = 0.5;
	s->force_limit = 1000.0;
	s->t_swing_avg = 0.3;
	s->dt_base_poly = 0.1;
	s->force_polys_per_stance = 3;
	s->ee_polys_per_swing = 2;
	s->dt_dynamic = 0.1;
	s->dt_rom = 0.08;
	s->base_rom = 0; s->dt_base_rom = 0.1 / 4.0; s->terrain_gradients = 0;
	s->cost_force_z = 0.0; s->cost_ee_vel_xy = 0.0;
	s->optimize_timings = 0; s->phase_dur_min = 0.2; s->phase_dur_max = 1.0;     /* ref: src/parameters.cc:52 */
	s->combo = ORC_CUSTOM;
	s->duration = 5.0;
}

/* ref: src/time_discretization_constraint.cc:41-49 */
static double *make_times(double T, double dt, int *n_out)
{
	int nf = (int)floor(T / dt);
	double *t = (double *)malloc(sizeof(double) * (nf + 2));
	double acc = 0.0; int n = 0;
	t[n++] = acc;
	for (int i = 0; i < nf; ++i) { acc += dt; t[n++] = acc; }
	t[n++] = T;
	*n_out = n;
	return t;
}

orc_problem *orc_problem_create(const orc_shape *shape, const orc_instance *inst,
                                const orc_heightfield *hf)
{
	orc_problem *p = (orc_problem *)calloc(1, sizeof(orc_problem));
	p->shape = *shape; p->inst = *inst; p->hf = *hf;
	if (build_gait(shape->combo, shape->duration, p->n_phases, p->phase_dur, p->contact_at_start)) {
		free(p); return NULL;
	}
	/* ref: src/parameters.cc:115-127 GetTotalTime = sum of foot 0 durations */
	p->T = 0.0;
	for (int k = 0; k < p->n_phases[0]; ++k) p->T += p->phase_dur[0][k];

	/* ref: src/parameters.cc:82-98 GetBasePolyDurations */
	double bd[4096]; int nb = 0;
	{
		double dt = shape->dt_base_poly, t_left = p->T, eps = 1e-10;
		while (t_left > eps) { bd[nb++] = t_left > dt ? dt : t_left; t_left -= dt; }
	}
	build_base_spline(&p->base_lin, bd, nb);
	build_base_spline(&p->base_ang, bd, nb);
	for (int ee = 0; ee < ORC_NEE; ++ee) {
		build_phase_polys(&p->ee_motion[ee], p->n_phases[ee], p->phase_dur[ee],
		                  p->contact_at_start[ee], shape->ee_polys_per_swing);
		param_ee_motion(&p->ee_motion[ee]);
		build_phase_polys(&p->ee_force[ee], p->n_phases[ee], p->phase_dur[ee],
		                  !p->contact_at_start[ee], shape->force_polys_per_stance);
		param_ee_force(&p->ee_force[ee]);
	}
	/* ifopt variable order, ref: src/nlp_formulation.cc:63-98 */
	int off = 0;
	p->base_lin.offset = off; off += p->base_lin.n_vars;
	p->base_ang.offset = off; off += p->base_ang.n_vars;
	for (int ee = 0; ee < ORC_NEE; ++ee) { p->ee_motion[ee].offset = off; off += p->ee_motion[ee].n_vars; }
	for (int ee = 0; ee < ORC_NEE; ++ee) { p->ee_force[ee].offset = off; off += p->ee_force[ee].n_vars; }
	/* contact schedule sets come last, ref: src/nlp_formulation.cc:80-83; all phases but the last, ref: src/phase_durations.cc:44-46 */
	for (int ee = 0; ee < ORC_NEE; ++ee) {
		p->sched_off[ee] = -1; p->row_total[ee] = -1;
		if (shape->optimize_timings) { p->sched_off[ee] = off; off += p->n_phases[ee] - 1; }
	}
	p->n = off;

	p->x0 = (double *)calloc(p->n, sizeof(double));
	p->xl = (double *)malloc(sizeof(double) * p->n);
	p->xu = (double *)malloc(sizeof(double) * p->n);
	for (int i = 0; i < p->n; ++i) { p->xl[i] = -ORC_INF; p->xu[i] = ORC_INF; }

	/* ---- initial guess + variable bounds ---- */
	const double zero3[3] = {0, 0, 0};
	{   /* ref: src/nlp_formulation.cc:100-134 (MakeBaseVariables) */
		double fx = inst->goal[0], fy = inst->goal[1];
		double fz = orc_height(hf, fx, fy) - shape->nominal[0][2];
		double fin[3] = {fx, fy, fz};
		set_linear(&p->base_lin, inst->start_pos, fin, p->T);
		int last = p->base_lin.n_nodes - 1;
		for (int d = 0; d < 3; ++d) {
			add_bound(p, &p->base_lin, 0, 0, d, inst->start_pos[d]);
			add_bound(p, &p->base_lin, 0, 1, d, inst->start_vel[d]);
			add_bound(p, &p->base_lin, last, 1, d, 0.0);          /* final lin vel {X,Y,Z} */
		}
		add_bound(p, &p->base_lin, last, 0, 0, inst->goal[0]);    /* final lin pos {X,Y} only */
		add_bound(p, &p->base_lin, last, 0, 1, inst->goal[1]);
		/* final base angles forced to 0, ref: src/main.cpp:420 */
		set_linear(&p->base_ang, inst->start_ang, zero3, p->T);
		for (int d = 0; d < 3; ++d) {
			add_bound(p, &p->base_ang, 0, 0, d, inst->start_ang[d]);
			add_bound(p, &p->base_ang, 0, 1, d, inst->start_ang_vel[d]);
			add_bound(p, &p->base_ang, last, 0, d, 0.0);
			add_bound(p, &p->base_ang, last, 1, d, 0.0);
		}
	}
	for (int ee = 0; ee < ORC_NEE; ++ee) {
		/* ref: src/nlp_formulation.cc:136-166; yaw = final_base_.ang.z = 0 -> R = I */
		double x = inst->goal[0] + shape->nominal[ee][0];
		double y = inst->goal[1] + shape->nominal[ee][1];
		double fin[3] = {x, y, orc_height(hf, x, y)};
		set_linear(&p->ee_motion[ee], inst->ee[ee], fin, p->T);
		for (int d = 0; d < 3; ++d) add_bound(p, &p->ee_motion[ee], 0, 0, d, inst->ee[ee][d]);
		/* ref: src/nlp_formulation.cc:168-190 */
		double f[3] = {0.0, 0.0, shape->mass * 9.80665 / ORC_NEE};
		set_linear(&p->ee_force[ee], f, f, p->T);
	}
	spline_get_values(&p->base_lin, p->x0);
	spline_get_values(&p->base_ang, p->x0);
	for (int ee = 0; ee < ORC_NEE; ++ee) spline_get_values(&p->ee_motion[ee], p->x0);
	for (int ee = 0; ee < ORC_NEE; ++ee) spline_get_values(&p->ee_force[ee], p->x0);
	if (shape->optimize_timings)                             /* ref: src/phase_durations.cc:68-76,100-109 */
		for (int ee = 0; ee < ORC_NEE; ++ee)
			for (int k = 0; k < p->n_phases[ee] - 1; ++k) {
				p->x0[p->sched_off[ee] + k] = p->phase_dur[ee][k];
				p->xl[p->sched_off[ee] + k] = shape->phase_dur_min; p->xu[p->sched_off[ee] + k] = shape->phase_dur_max;
			}

	/* ---- constraint layout, ref: src/parameters.cc:55-60 order ---- */
	p->t_dyn = make_times(p->T, shape->dt_dynamic, &p->n_dyn);
	p->t_rom = make_times(p->T, shape->dt_rom, &p->n_rom);
	int row = 0;
	for (int ee = 0; ee < ORC_NEE; ++ee) { p->row_terrain[ee] = row; row += p->ee_motion[ee].n_nodes - 1; }
	p->row_dynamic = row; row += 6 * p->n_dyn;
	p->row_acc_lin = row; row += 3 * (p->base_lin.n_polys - 1);
	p->row_acc_ang = row; row += 3 * (p->base_ang.n_polys - 1);
	for (int ee = 0; ee < ORC_NEE; ++ee) { p->row_rom[ee] = row; row += 3 * p->n_rom; }
	for (int ee = 0; ee < ORC_NEE; ++ee) {
		int cnt = 0;
		for (int nd = 0; nd < p->ee_force[ee].n_nodes; ++nd) cnt += !is_const_node(&p->ee_force[ee], nd);
		p->row_force[ee] = row; row += 5 * cnt;
	}
	for (int ee = 0; ee < ORC_NEE; ++ee) {
		int cnt = 0;
		for (int nd = 0; nd < p->ee_motion[ee].n_nodes; ++nd) cnt += !is_const_node(&p->ee_motion[ee], nd);
		p->row_swing[ee] = row; row += 4 * cnt;
	}
	p->row_base_rom = -1; p->n_brom = 0; p->t_brom = NULL;
	if (shape->base_rom) {                                   /* appended like constraints_.push_back(BaseRom) would */
		p->t_brom = make_times(p->T, shape->dt_base_rom, &p->n_brom);
		p->row_base_rom = row; row += 6 * p->n_brom;
	}
	if (shape->optimize_timings)                             /* constraints_.push_back(TotalTime), ref: src/parameters.cc:77-80 */
		for (int ee = 0; ee < ORC_NEE; ++ee) p->row_total[ee] = row++;
	p->m = row;
	p->gl = (double *)calloc(p->m, sizeof(double));
	p->gu = (double *)calloc(p->m, sizeof(double));
	/* bounds: everything is an equality (0,0) except:
	 * terrain swing nodes [0,1e20]  ref: src/terrain_constraint.cc:73-88
	 * RoM box                       ref: src/range_of_motion_constraint.cc:71-83
	 * force                         ref: src/force_constraint.cc:95-108 */
	for (int ee = 0; ee < ORC_NEE; ++ee) {
		const orc_spline *s = &p->ee_motion[ee];
		for (int nd = 1; nd < s->n_nodes; ++nd)
			if (!is_const_node(s, nd)) p->gu[p->row_terrain[ee] + nd - 1] = ORC_INF;
		for (int k = 0; k < p->n_rom; ++k)
			for (int d = 0; d < 3; ++d) {
				int r = p->row_rom[ee] + 3 * k + d;
				p->gl[r] = shape->nominal[ee][d] - shape->max_dev[d];
				p->gu[r] = shape->nominal[ee][d] + shape->max_dev[d];
			}
		int r = p->row_force[ee];
		for (int nd = 0; nd < p->ee_force[ee].n_nodes; ++nd) {
			if (is_const_node(&p->ee_force[ee], nd)) continue;
			p->gl[r] = 0.0;       p->gu[r] = shape->force_limit; r++;
			p->gl[r] = -ORC_INF;  p->gu[r] = 0.0;  r++;
			p->gl[r] = 0.0;       p->gu[r] = ORC_INF; r++;
			p->gl[r] = -ORC_INF;  p->gu[r] = 0.0;  r++;
			p->gl[r] = 0.0;       p->gu[r] = ORC_INF; r++;
		}
	}
	/* ref: src/total_duration_constraint.cc:56-62 (the hard-coded 0.1 and min_duration_last_phase = 0.2) */
	if (shape->optimize_timings)
		for (int ee = 0; ee < ORC_NEE; ++ee) { p->gl[p->row_total[ee]] = 0.1; p->gu[p->row_total[ee]] = p->T - 0.2; }
	/* BaseMotionConstraint bounds, ref: src/base_motion_constraint.cc:47-57: roll, pitch within +-0.01 rad, yaw and x, y free,
	 * z within [z_init - 0.02, z_init + 0.1], z_init = the base spline's initial height */
	for (int k = 0; k < p->n_brom; ++k) {
		double *lo = p->gl + p->row_base_rom + 6 * k, *hi = p->gu + p->row_base_rom + 6 * k;
		lo[0] = lo[1] = -0.01; hi[0] = hi[1] = 0.01;
		lo[2] = lo[3] = lo[4] = -ORC_INF; hi[2] = hi[3] = hi[4] = ORC_INF;
		lo[5] = inst->start_pos[2] - 0.02; hi[5] = inst->start_pos[2] + 0.1;
	}
	return p;
}

void orc_problem_free(orc_problem *p)
{
	if (!p) return;
	spline_free(&p->base_lin); spline_free(&p->base_ang);
	for (int ee = 0; ee < ORC_NEE; ++ee) { spline_free(&p->ee_motion[ee]); spline_free(&p->ee_force[ee]); }
	free(p->x0); free(p->xl); free(p->xu); free(p->gl); free(p->gu); free(p->t_dyn); free(p->t_rom); free(p->t_brom);
	free(p);
}

int orc_n(const orc_problem *p) { return p->n; }
int orc_m(const orc_problem *p) { return p->m; }
void orc_get_x0(const orc_problem *p, double *x0) { memcpy(x0, p->x0, sizeof(double) * p->n); }
void orc_get_bounds(const orc_problem *p, double *xl, double *xu, double *gl, double *gu)
{
	memcpy(xl, p->xl, sizeof(double) * p->n); memcpy(xu, p->xu, sizeof(double) * p->n);
	memcpy(gl, p->gl, sizeof(double) * p->m); memcpy(gu, p->gu, sizeof(double) * p->m);
}
int orc_get_phase_durations(const orc_problem *p, int ee, double *out)
{
	memcpy(out, p->phase_dur[ee], sizeof(double) * p->n_phases[ee]);
	return p->n_phases[ee];
}
void orc_get_schedule_layout(const orc_problem *p, int *so, int *rt)
{
	for (int ee = 0; ee < 4; ++ee) { so[ee] = p->sched_off[ee]; rt[ee] = p->row_total[ee]; }
}
void orc_get_layout(const orc_problem *p, int *vo, int *ro)
{
	vo[0] = p->base_lin.offset; vo[1] = p->base_ang.offset;
	for (int ee = 0; ee < 4; ++ee) { vo[2 + ee] = p->ee_motion[ee].offset; vo[6 + ee] = p->ee_force[ee].offset; }
	vo[10] = p->ee_force[3].offset + p->ee_force[3].n_vars;      /* schedule sets (if any) follow: orc_get_schedule_layout */
	for (int ee = 0; ee < 4; ++ee) ro[ee] = p->row_terrain[ee];
	ro[4] = p->row_dynamic; ro[5] = p->row_acc_lin; ro[6] = p->row_acc_ang;
	for (int ee = 0; ee < 4; ++ee) { ro[7 + ee] = p->row_rom[ee]; ro[11 + ee] = p->row_force[ee]; ro[15 + ee] = p->row_swing[ee]; }
	ro[19] = p->m;
}
