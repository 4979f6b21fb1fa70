/*
 * towr_oracle.h -- CPU oracle for the QTOS local planner hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This directory is a plain-C, double precision
 * restatement of the reference's algorithm (TOWR formulation + an
 * interior-point NLP loop).  It exists to check the CUDA product path and to
 * be timed as the CPU baseline.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (qtos_b200/) never includes, links or calls anything from here.
 *
 * Parity status: the TOWR formulation part (variables, x0, bounds, g(x),
 * J(x), heightfield, CSV sampling) is PINNED against the reference's golden
 * artefacts (logs/towr_log.out structure + iteration-0 inf_pr, the golden
 * CSVs; see tests/test_oracle_golden.py).  The interior-point loop restates
 * Ipopt's published algorithm (Waechter & Biegler 2006; Ipopt 3.11.9 + MUMPS
 * is what the reference ran, source NOT vendored under /root/reference), so
 * iterate-level parity with Ipopt is "unpinned" beyond the logged iteration
 * tables; see DESIGN.md.
 *
 * Citations "ref:" are paths relative to /root/reference/solver/towr/.
 */
#ifndef TOWR_ORACLE_H_
#define TOWR_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NEE 4          /* ref: include/towr/models/endeffector_mappings.h:44  LF,RF,LH,RH */
#define ORC_MAX_PHASES 64
#define ORC_INF 1e20       /* ifopt "inf" */

/* gait combos, ref: src/quadruped_gait_generator.cc:76-88 */
enum { ORC_C0 = 0, ORC_C1, ORC_C2, ORC_C3, ORC_C4, ORC_CUSTOM };

typedef struct {
	int nx, ny;          /* hf[ix][iy], ix = row of the text file = world x */
	double res;          /* x_step_length_ == y_step_length_ (mesh_scale) */
	const double *h;     /* nx*ny, row-major [ix*ny + iy] */
} orc_heightfield;

/* everything main.cpp + Parameters + Solo12 model hard-code, made explicit */
typedef struct {
	/* model, ref: include/towr/models/examples/solo12_model.h:17-37 */
	double mass;
	double I_b[9];               /* row-major body inertia (incl. the F5 quirk ordering) */
	double nominal[ORC_NEE][3];
	double max_dev[3];
	double mu;                   /* ref: include/towr/terrain/height_map.h:137 */
	double force_limit;          /* ref: src/parameters.cc:48 */
	double t_swing_avg;          /* ref: include/towr/constraints/swing_constraint.h:68 */
	/* discretisation, ref: src/parameters.cc:40-73 */
	double dt_base_poly;
	int    force_polys_per_stance;
	int    ee_polys_per_swing;
	double dt_dynamic;
	double dt_rom;
	/* gait + horizon, ref: src/main.cpp:299-306,424-433 */
	int    combo;
	double duration;
	/* optional BaseMotionConstraint (Parameters::BaseRom, off on the reference's path), ref: src/base_motion_constraint.cc:38-93,
	 * src/parameters.cc:51: rows appended after the swing sets, dt = duration_base_polynomial / 4 */
	int    base_rom;
	double dt_base_rom;
	/* optional: the terrain derivatives the reference carries commented out (ref: src/custom_terrain.cpp:101-124,133-156) in the
	 * terrain rows' Jacobian and in the force rows' contact basis (ref: src/height_map.cc:95-141, src/force_constraint.cc:67-135) */
	int    terrain_gradients;
	/* optional cost terms (Parameters::costs_, empty on the reference's path; ref: src/parameters.cc:62-63, src/nlp_formulation.cc:343-376,
	 * src/node_cost.cc:53-83): weight of ForcesCostID (force z of every force node, squared) and of EEMotionCostID (foot velocity
	 * x and y of every motion node, squared); 0 = the term is absent */
	double cost_force_z, cost_ee_vel_xy;
	/* optional gait-timing optimisation (Parameters::OptimizePhaseDurations, off on the reference's path; ref: src/parameters.cc:52,77-80,
	 * src/phase_durations.cc, src/phase_spline.cc:67-93, src/total_duration_constraint.cc): every foot's phase durations but the last
	 * become variables (sets appended after the force sets, bounds [phase_dur_min, phase_dur_max]), the foot splines follow them, the
	 * dynamic and range-of-motion rows gain columns for them and one TotalDuration row per foot is appended.  ORACLE ONLY. */
	int    optimize_timings;
	double phase_dur_min, phase_dur_max;     /* bound_phase_duration_ = (0.2, 1.0) */
} orc_shape;

typedef struct {
	double start_pos[3], start_ang[3], start_vel[3], start_ang_vel[3];
	double goal[3];
	double ee[ORC_NEE][3];
	double t_start;
} orc_instance;

/* node-value spline: nodes hold pos+vel per dim; opt[] maps to the set-local
 * optimisation index or -1 (ref: src/nodes_variables*.cc) */
typedef struct {
	int n_nodes, n_polys, n_vars, offset;
	double *dur;     /* [n_polys] */
	int    *opt;     /* [n_nodes][2][3] */
	double *val;     /* [n_nodes][2][3] */
	int    *poly_phase;   /* [n_polys] phase id (phase-based splines) */
	int    *poly_const;   /* [n_polys] 1 if in a constant phase */
} orc_spline;

typedef struct orc_problem {
	orc_shape shape;
	orc_instance inst;
	orc_heightfield hf;
	int n_phases[ORC_NEE];
	double phase_dur[ORC_NEE][ORC_MAX_PHASES];
	int contact_at_start[ORC_NEE];
	double T;
	orc_spline base_lin, base_ang, ee_motion[ORC_NEE], ee_force[ORC_NEE];
	int n, m;                 /* variables, constraint rows */
	double *x0, *xl, *xu, *gl, *gu;
	/* constraint-set row offsets in ifopt order */
	int row_terrain[ORC_NEE], row_dynamic, row_acc_lin, row_acc_ang,
	    row_rom[ORC_NEE], row_force[ORC_NEE], row_swing[ORC_NEE];
	int n_dyn, n_rom;         /* sample counts */
	double *t_dyn, *t_rom;
	int row_base_rom, n_brom; /* BaseMotionConstraint rows (6 per sample: AX AY AZ LX LY LZ), -1 / 0 when off */
	double *t_brom;
	int sched_off[ORC_NEE];   /* variable offset of foot ee's phase durations (n_phases - 1 of them), -1 when timings are fixed */
	int row_total[ORC_NEE];   /* TotalDuration row of foot ee, -1 when off */
} orc_problem;

/* ---- problem construction (towr_problem.c) ---- */
void orc_default_shape(orc_shape *s);   /* Solo12 constants as vendored (m=1.5) */
orc_problem *orc_problem_create(const orc_shape *s, const orc_instance *inst,
                                const orc_heightfield *hf);
void orc_problem_free(orc_problem *p);
int  orc_n(const orc_problem *p);
int  orc_m(const orc_problem *p);
void orc_get_x0(const orc_problem *p, double *x0);
void orc_get_bounds(const orc_problem *p, double *xl, double *xu, double *gl, double *gu);
int  orc_get_phase_durations(const orc_problem *p, int ee, double *out);
void orc_get_layout(const orc_problem *p, int *var_offsets /*10+1*/, int *row_offsets /*19+1*/);
void orc_get_schedule_layout(const orc_problem *p, int *sched_off /*4*/, int *row_total /*4*/);

/* ---- terrain (towr_terrain.c) ---- */
double orc_height(const orc_heightfield *hf, double x, double y);
void   orc_height_cell(const orc_heightfield *hf, double x, double y, long long idx[4]);
void   orc_height_deriv(const orc_heightfield *hf, double x, double y, double *hx, double *hy);

/* ---- evaluation (towr_eval.c) ---- */
void orc_set_x(orc_problem *p, const double *x);
void orc_eval_g(orc_problem *p, const double *x, double *g);
/* dense row-major m*n Jacobian; mask (nullable, m*n bytes) gets 1 where the
 * reference's FillJacobianBlock creates a structural entry */
void orc_eval_jac(orc_problem *p, const double *x, double *J, unsigned char *mask);
/* objective = sum of the NodeCost terms of orc_shape (0 when none); grad (nullable, n entries) is its gradient */
double orc_eval_cost(orc_problem *p, const double *x, double *grad);

/* ---- 1 kHz sampler (towr_csv.c) ---- */
int  orc_csv_rows(const orc_problem *p, double dt);
void orc_sample_csv(orc_problem *p, const double *x, double dt, double *rows /* n_rows*37 */);
int  orc_write_csv(orc_problem *p, const double *x, double dt, const char *path);

/* ---- interior point (towr_ipm.c) ---- */
typedef struct {
	double tol;               /* 1e-3 (ifopt) */
	double constr_viol_tol;   /* 1e-4 */
	double compl_inf_tol;     /* 1e-4 */
	double dual_inf_tol;      /* 1.0 */
	int    max_iter;          /* 200 (main.cpp:461) */
	double mu_init;           /* 0.1 */
	int    mu_strategy;       /* 0 monotone, 1 adaptive (loqo-type) */
	double sigma_w;           /* W = sigma_w * I */
	int    verbose;
	double delta_c;           /* equality-block regularisation (condensed as Jc'Jc/delta_c) */
	int    feas_exit;         /* f == 0: accept any point with viol <= constr_viol_tol (zero multipliers are optimal) */
} orc_ipm_options;

typedef struct {
	int status;               /* 0 solved, 1 acceptable, -1 max iter, -2 restoration/step failure, 2 infeasible */
	int iters;
	double constr_viol, dual_inf, compl_inf, nlp_error;
	double mu;
	/* per-iteration trace (first 256) */
	int n_trace;
	double tr_inf_pr[256], tr_inf_du[256], tr_mu[256], tr_dnorm[256],
	       tr_alpha_pr[256], tr_alpha_du[256];
	int tr_ls[256];
	int chol_fix;             /* factorizations that hit a non-positive pivot */
} orc_ipm_result;

void orc_ipm_default_options(orc_ipm_options *o);
int  orc_ipm_solve(orc_problem *p, const orc_ipm_options *o, double *x /* n, out */,
                   orc_ipm_result *res);

/* ---- the Ipopt 3.11.9 algorithm of the reference path (towr_ipopt.c; see its header) ---- */
#define ORC_TRACE_MAX 256
#define ORC_FILTER_MAX 32
typedef struct {
	double tol, constr_viol_tol, compl_inf_tol, dual_inf_tol;   /* 1e-3 (ifopt), 1e-4, 1e-4, 1 */
	int    max_iter;          /* 200 (main.cpp:461) */
	double delta_c;           /* penalty 1/delta_c on the equality block of the condensed system */
	int    n_refine;          /* multiplier-method passes that remove the penalty's bias */
	int    lm_history;        /* limited_memory_max_history 6 */
	int    verbose;
	double sigma_floor;       /* lower bound of the limited-memory scalar sigma_w (0: Ipopt's 1e-8) */
	int    retry_failed;      /* 1: a second attempt with sigma_floor = retry_sigma_floor after a failed line search */
	double retry_sigma_floor; /* 1e-2 */
} orc_ipopt_options;

typedef struct {
	int status, iters;
	double constr_viol, dual_inf, compl_inf, nlp_error, mu;
	int n_trace;
	double tr_inf_pr[ORC_TRACE_MAX], tr_inf_du[ORC_TRACE_MAX], tr_mu[ORC_TRACE_MAX], tr_dnorm[ORC_TRACE_MAX],
	       tr_alpha_pr[ORC_TRACE_MAX], tr_alpha_du[ORC_TRACE_MAX];
	int tr_ls[ORC_TRACE_MAX], tr_pairs[ORC_TRACE_MAX], tr_free[ORC_TRACE_MAX];
	char tr_tag[ORC_TRACE_MAX];
	int chol_fix;
	int n_regularized;        /* factorizations repeated with W + delta_w I */
	int retried;              /* the second attempt ran */
	double objective;         /* unscaled f at the returned point (0 without cost terms) */
} orc_ipopt_result;

void orc_ipopt_default_options(orc_ipopt_options *o);
int  orc_ipopt_solve(orc_problem *p, const orc_ipopt_options *o, double *x /* n, in: x0, out */, orc_ipopt_result *res);

#ifdef __cplusplus
}
#endif
#endif
