/*
 * qtos_b200.h -- C ABI of the B200-native batched gait-planning NLP solver.
 *
 * Drop-in boundary for the QTOS local planner.  The reference crosses this boundary as a
 * process call `docker exec <id> ./main <flags>` (ref: QTOS/utils.py:15-26,644-670;
 * scripts/main.py:48-50,90-92; QTOS/generateHeightField.py:385-386) into
 * solver/towr/src/main.cpp:133-471.  The entry points below are what an in-process FFI for
 * that path binds; INTEGRATION.md shows the ctypes stub on the reference side.
 * Plain pointers and sizes only; every array is caller-owned.  All functions return 0 on
 * success or a negative QTOS_E* code; qtos_last_error() gives the message.  There is no CPU
 * fallback: without a CUDA device qtos_create fails with QTOS_ENODEV.
 */
#ifndef QTOS_B200_H_
#define QTOS_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define QTOS_NEE 4                /* LF, RF, LH, RH (ref: towr/models/endeffector_mappings.h:44) */
#define QTOS_CSV_COLS 37          /* ref: solver/towr/src/main.cpp:92-131 */

enum { QTOS_OK = 0, QTOS_EINVAL = -1, QTOS_ENODEV = -2, QTOS_ECUDA = -3, QTOS_ENOMEM = -4, QTOS_ESHAPE = -5 };

/* gait combos, ref: solver/towr/src/quadruped_gait_generator.cc:76-88 */
enum { QTOS_C0 = 0, QTOS_C1, QTOS_C2, QTOS_C3, QTOS_C4, QTOS_CUSTOM };

/* per-problem solver status; exit code of ./main (ref: main.cpp:463,471) is status & 0xff */
enum {
	QTOS_SOLVE_SUCCEEDED = 0,         /* Ipopt Solve_Succeeded */
	QTOS_MAX_ITER = -1,               /* Ipopt Maximum_Iterations_Exceeded */
	QTOS_STEP_FAILED = -2,            /* line search could not make progress (Ipopt: Restoration_Failed) */
	QTOS_MAX_CPUTIME = -4,            /* Ipopt Maximum_CpuTime_Exceeded: the iteration budget derived from max_cpu_time (-r) ran out */
	QTOS_INVALID_NUMBER = -13,        /* Ipopt Invalid_Number_Detected: NaN/Inf in the constraints (bad inputs), or a problem that
	                                     names a heightfield id that does not exist (device-resident problems cannot be checked on the host) */
	QTOS_RUNNING = 99
};

/* Everything main.cpp + Parameters + the Solo12 model hard-code, made explicit
 * (ref: solver/towr/src/parameters.cc:40-73; models/examples/solo12_model.h:17-37;
 *  height_map.h:137; swing_constraint.h:68; main.cpp:299-306,424-433). */
typedef struct {
	double mass;
	double I_b[9];                    /* row-major body inertia */
	double nominal[QTOS_NEE][3];
	double max_dev[3];
	double mu;                        /* friction coefficient */
	double force_limit;
	double t_swing_avg;
	double dt_base_poly;
	int    force_polys_per_stance;
	int    ee_polys_per_swing;
	double dt_dynamic;
	double dt_rom;
	int    combo;                     /* QTOS_C0 .. QTOS_CUSTOM */
	double duration;
	int    base_rom;                  /* 1: add TOWR's optional BaseMotionConstraint (Parameters::BaseRom, off on the reference's path;
	                                     ref: base_motion_constraint.cc:38-93) after the swing sets: roll, pitch within +-0.01 rad and
	                                     z(t) - z(0) within [-0.02, 0.1] at every dt_base_rom; yaw, x, y rows are kept unbounded like the
	                                     reference's.  The z row is stated relative to the start height, so its bounds are the shape's */
	double dt_base_rom;               /* duration_base_polynomial / 4 (ref: parameters.cc:51) */
	double cost_force_z;              /* weight of TOWR's optional ForcesCostID term: sum over the force nodes of weight * f_z^2, and */
	double cost_ee_vel_xy;            /* of EEMotionCostID: sum over the foot-motion nodes of weight * (v_x^2 + v_y^2)
	                                     (Parameters::costs_, empty on the reference's path: parameters.cc:62-63; nlp_formulation.cc:343-376;
	                                     node_cost.cc:53-83).  0 = the term is absent.  With a weight set the objective is minimised by
	                                     QTOS_ALG_IPOPT exactly as Ipopt would (gradient in the dual infeasibility, the limited-memory pairs,
	                                     the barrier function of the filter line search); QTOS_ALG_FAST refuses such a shape */
	int    terrain_gradients;         /* 1: the terrain's first derivatives are what the bilinear surface gives (the code the reference
	                                     carries commented out, custom_terrain.cpp:96-156) instead of the reference's zeros: the terrain
	                                     rows' Jacobian gets -dh/dx, -dh/dy (terrain_constraint.cc:90-108) and the force rows are stated
	                                     in the contact basis of the foothold (height_map.cc:95-141, force_constraint.cc:67-135; second
	                                     derivatives stay zero like the reference's).  0 (default) = the reference's path.  On plateau
	                                     terrain this is a regression */
} qtos_shape;

/* one local-plan window = the flags of ./main (ref: main.cpp:163-306) */
typedef struct {
	double start_pos[3];              /* -s */
	double start_ang[3];              /* -s_ang */
	double start_vel[3];              /* -s_vel (zeroed by main.cpp when -n is absent) */
	double start_ang_vel[3];          /* -s_ang_vel */
	double goal[3];                   /* -g (z unused downstream) */
	double ee[QTOS_NEE][3];           /* -e1..-e4 */
	double t_start;                   /* -t */
	int    hf_id;                     /* heightfield handle from qtos_upload_heightfield */
	int    group;                     /* multi-start group id (best-plan selection), else 0 */
} qtos_problem;

typedef struct {
	double tol;                       /* 1e-3 (ifopt default in force, logs/towr_log.out) */
	double constr_viol_tol;           /* 1e-4 */
	double compl_inf_tol;             /* 1e-4 */
	double dual_inf_tol;              /* 1.0  */
	int    max_iter;                  /* 200 (ref: main.cpp:461) */
	double mu_init;                   /* 0.1 */
	double sigma_w;                   /* Hessian model sigma_w * I */
	double delta_c;                   /* equality-block penalty 1/delta_c of the condensed system; 0 = the algorithm's default
	                                     (IPOPT 1e-6 with n_refine multiplier passes, FAST 1e-5) */
	int    feas_exit;                 /* FAST only. 1: f == 0 on this path, so any point with violation <= constr_viol_tol is
	                                     optimal with zero multipliers and terminates the solve (default) */
	int    algorithm;                 /* QTOS_ALG_IPOPT (default) or QTOS_ALG_FAST */
	int    n_refine;                  /* IPOPT: multiplier-method passes on the equality block per direction (1) */
	int    lm_history;                /* IPOPT: limited_memory_max_history (6, the maximum) */
	double max_cpu_time;              /* the reference's `-r` / Ipopt max_cpu_time in seconds (ref: main.cpp:180,459-460); 0 = none.
	                                     Wall-clock limits are not deterministic on a batched device, so the budget is counted in
	                                     iterations: floor(max_cpu_time / QTOS_REF_SECONDS_PER_ITERATION), i.e. the iterations the
	                                     reference itself completes in that time; reaching it returns QTOS_MAX_CPUTIME */
	double lm_init_val_min;           /* IPOPT: Ipopt's limited_memory_init_val_min, the floor of the L-BFGS scalar sigma_w; 0 = Ipopt's 1e-8 */
	int    retry_failed;              /* IPOPT: 1 (default) = a window whose filter line search fails (where Ipopt would enter its
	                                     restoration phase) is restarted once from x0 with lm_init_val_min = retry_lm_init_val_min;
	                                     its iteration count is the sum of both attempts.  0 = report -2 at once */
	double retry_lm_init_val_min;     /* 1e-2 */
} qtos_options;
#define QTOS_REF_SECONDS_PER_ITERATION 0.1   /* 0.75-0.88 s for 7-8 iterations, ref: logs/towr_log.out:64,81-82 */

/* QTOS_ALG_IPOPT: the algorithm the reference runs (Ipopt 3.11.9 as configured by ifopt, ref: solver/towr/src/main.cpp:444-463,
 *   logs/towr_log.out:37-64): limited-memory BFGS Hessian (history 6), adaptive quality-function barrier update, filter line
 *   search; reproduces the reference's logged iteration tables and plans (tests/test_gpu_parity.py).  mu_init, sigma_w and
 *   feas_exit are ignored.
 * QTOS_ALG_FAST: Hessian model sigma_w I, monotone barrier update, l1-merit line search: a feasible plan, not Ipopt's plan. */
enum { QTOS_ALG_IPOPT = 0, QTOS_ALG_FAST = 1 };
#define QTOS_TRACE_ITERS 48       /* iterations kept by qtos_get_trace */
#define QTOS_TRACE_COLS 8         /* inf_pr, inf_du, mu, ||d||, alpha_du, alpha_pr, line-search trials, step tag ('f' / 'h' as a number) */

typedef struct {
	int    status;
	int    iters;
	double constr_viol;               /* unscaled max violation of g_L <= g(x) <= g_U */
	double dual_inf;
	double compl_inf;
	double nlp_error;                 /* Ipopt's scaled overall error E_0 */
	double mu;
	double cost;                      /* post-hoc plan cost used for best-plan selection */
} qtos_result;

typedef struct {
	int n_vars;                       /* ifopt variable count (1040 for T=5 s Custom) */
	int n_cons;                       /* constraint rows (1730) */
	int n_free;                       /* after fixed-variable elimination (1005) */
	int n_eq, n_ineq;
	int nnz_jac;                      /* structural non-zeros kept on the device */
	int csv_rows;                     /* rows of the 1 kHz trajectory (5001) */
	int kkt_order, kkt_block, kkt_blocks;    /* condensed system: padded order, block size, stored blocks */
	double flops_factor;              /* algorithmic flops of one factorization (sum of w_i^2) */
	long long workspace_bytes_per_problem;
} qtos_dims;

typedef struct qtos_ctx qtos_ctx;

void qtos_default_shape(qtos_shape *s);            /* Solo12 constants as vendored, Custom gait, T = 5 s */
void qtos_default_options(qtos_options *o);

/* One context = one device + one compiled shape + workspace for max_batch concurrent problems. */
int  qtos_create(int device, const qtos_shape *shape, int max_batch, qtos_ctx **out);
void qtos_destroy(qtos_ctx *ctx);
const char *qtos_last_error(const qtos_ctx *ctx);  /* ctx may be NULL: error of the last failed qtos_create */
int  qtos_get_dims(const qtos_ctx *ctx, qtos_dims *d);

/* hf[ix*ny + iy], ix = row of towr_heightfield.txt = world x (ref: custom_terrain.cpp:12-49) */
int  qtos_upload_heightfield(qtos_ctx *ctx, const double *hf, int nx, int ny, double res, int *hf_id);
/* release a grid; its id is reused by a later upload; problems that still name it end with QTOS_INVALID_NUMBER */
int  qtos_free_heightfield(qtos_ctx *ctx, int hf_id);
/* batched CustomTerrain::GetHeight (ref: custom_terrain.cpp:51-94), bit-exact; xy = n pairs */
int  qtos_heightfield_query(qtos_ctx *ctx, int hf_id, const double *xy, int n, double *h_out);
int  qtos_heightfield_cells(qtos_ctx *ctx, int hf_id, const double *xy, int n, long long *idx4_out);
/* dh/dx and dh/dy of the bilinear surface at n points: the derivative CustomTerrain::GetHeightDerivWrtX / WrtY carry commented out
 * (ref: custom_terrain.cpp:96-156), same cells and operation order as the height, bit-exact against the oracle's restatement.
 * The solver's terrain rows keep the reference's zero derivatives; this is the query alone */
int  qtos_heightfield_gradients(qtos_ctx *ctx, int hf_id, const double *xy, int n, double *hx_out, double *hy_out);

/* problem structure for one instance: x0, bounds (host arrays, n_vars / n_cons long, nullable) */
int  qtos_get_initial(qtos_ctx *ctx, const qtos_problem *p, int n, double *x0, double *xl, double *xu,
                      double *gl, double *gu);
/* constraint values and dense row-major Jacobian (n_cons x n_vars, fixed columns zero) at x;
 * x = n * n_vars; g_out, jac_out nullable */
int  qtos_eval(qtos_ctx *ctx, const qtos_problem *p, int n, const double *x, double *g_out, double *jac_out);

/* solve n independent windows (host buffers; H2D/D2H inside).  x_out: n * n_vars node values;
 * csv_out (nullable): n * csv_rows * 37 doubles */
int  qtos_solve_batch(qtos_ctx *ctx, const qtos_problem *p, int n, const qtos_options *o,
                      qtos_result *res, double *x_out, double *csv_out);
/* same with device-resident buffers (problems, results, x) on the context's stream */
int  qtos_solve_batch_device(qtos_ctx *ctx, const qtos_problem *d_p, int n, const qtos_options *o,
                             qtos_result *d_res, double *d_x_out);
/* Asynchronous variants (SURVEY 8b "batch call is synchronous by default with an async variant"): the call returns at once,
 * a worker thread of the context drives the solve on the context's stream; one call in flight per context; every buffer
 * must stay valid until qtos_wait returns (its return value is the solve's).  Two contexts on one device overlap: the
 * straggler iterations of one batch run beside the full launches of the next. */
int  qtos_solve_batch_async(qtos_ctx *ctx, const qtos_problem *p, int n, const qtos_options *o,
                            qtos_result *res, double *x_out, double *csv_out);
int  qtos_solve_batch_device_async(qtos_ctx *ctx, const qtos_problem *d_p, int n, const qtos_options *o,
                                   qtos_result *d_res, double *d_x_out);
int  qtos_wait(qtos_ctx *ctx);
/* Continuous batching (IPOPT algorithm): the context's max_batch workspace slots become a pool; jobs are submitted at any
 * time and their windows enter the pool as slots fall free, so every launch of the iteration kernels works on a full pool
 * and the last slow windows of one job run beside the fresh windows of the next (the reference's counterpart: 32 worker
 * processes polling one queue, ref: QTOS/generateHeightField.py:344-404).  A window's result is bit-identical to the one
 * qtos_solve_batch returns.  Host-buffer jobs hold at most max_batch windows, and 16 of them are staged on the device at a time
 * (further ones wait their turn in the queue); all buffers stay valid until the job's qtos_stream_wait returns.  Heightfields are uploaded before qtos_stream_begin. */
typedef struct {
	long long iterations;             /* batch iterations launched */
	long long slot_iterations;        /* occupied slots summed over them (= problem-iterations) */
	long long windows_done;
	long long launches;               /* kernel launches of the session */
	double factor_ms, solve_ms;       /* device time of k_factor / kip_solve (CUDA events on the context's stream) ... */
	long long timed_iterations, timed_slot_iterations;   /* ... over this many iterations / problem-iterations */
	long long retried;                /* windows that needed the second attempt (qtos_options.retry_failed) */
} qtos_stream_info;
int  qtos_stream_begin(qtos_ctx *ctx, const qtos_options *o);
int  qtos_stream_submit(qtos_ctx *ctx, const qtos_problem *p, int n, qtos_result *res, double *x_out, int *ticket);
/* like qtos_stream_submit, and the job also delivers the reference's actual output -- the 1 kHz rows of every plan, rows_out =
 * [n][csv_rows][37] doubles on the host (ref: main.cpp:92-131, what getTrajectory writes to traj.csv) -- sampled on the device and
 * copied on the session's copy stream while the pool keeps iterating; page-locked rows_out lets the copies run at PCIe speed */
int  qtos_stream_submit_csv(qtos_ctx *ctx, const qtos_problem *p, int n, qtos_result *res, double *x_out, double *rows_out, int *ticket);
int  qtos_stream_submit_device(qtos_ctx *ctx, const qtos_problem *d_p, int n, qtos_result *d_res, double *d_x_out, int *ticket);
int  qtos_stream_wait(qtos_ctx *ctx, int ticket);
int  qtos_stream_stats(const qtos_ctx *ctx, qtos_stream_info *out);
int  qtos_stream_end(qtos_ctx *ctx);              /* drains every submitted job, then closes the session */
/* Best-plan selection, the one exchange step of the path (multi-start candidates of the same window compete).  A record is five
 * doubles (group, not converged, cost, violation, global id); the winner of a group is the lexicographic minimum of the last
 * four.  qtos_make_records builds this rank's records on the device (candidate i gets id0 + id_stride * i), the caller
 * all-gathers them (NCCL), qtos_select_best returns the winning id of every group in [0, n_groups) (-1: no candidate). */
int  qtos_make_records(qtos_ctx *ctx, const qtos_result *d_res, const int *d_group, long long id0, long long id_stride, int n, double *d_rec);
int  qtos_select_best(qtos_ctx *ctx, const double *d_rec, int n_rec, int n_groups, long long *d_winner);
/* IPOPT algorithm: the per-iteration table Ipopt prints (ref: logs/towr_log.out:55-62), for the first n problems of the
 * last solve: trace_out = n * QTOS_TRACE_ITERS * QTOS_TRACE_COLS doubles; rows past a problem's last iteration are zero */
int  qtos_get_trace(qtos_ctx *ctx, int n, double *trace_out);
/* 1 kHz sampler (ref: main.cpp:92-131): rows_out = n * csv_rows * 37 */
int  qtos_sample_csv(qtos_ctx *ctx, const qtos_problem *p, int n, const double *x, double *rows_out);
/* rows [row0, row0 + n_rows) only: rows_out = n * n_rows * 37 (e.g. the one look-ahead row Combiner._state reads,
 * ref: QTOS/combiner.py:245-296) */
int  qtos_sample_csv_rows(qtos_ctx *ctx, const qtos_problem *p, int n, const double *x, int row0, int n_rows, double *rows_out);
/* write one trajectory as the reference's traj.csv text ("%g", comma separated) */
int  qtos_write_csv(const double *rows, int n_rows, const char *path);

/* instrumentation.  With profiling on, CUDA events are recorded between the kernels of every iteration
 * on the context's stream (no extra synchronisation) and resolved when the solve ends. */
typedef struct {
	float ms[8];                      /* device time of the last solve: init, jac, prepare, assemble, factor, step, solve (IPOPT), 0 */
	long long factorizations;         /* problems factored, summed over the iterations of the last solve */
	long long factor_launches;        /* k_factor launches of the last solve */
	int iterations;                   /* batch iterations of the last solve */
	int retried;                      /* windows that needed the second attempt (qtos_options.retry_failed) */
} qtos_stats;
long long qtos_launch_count(const qtos_ctx *ctx);  /* kernel launches issued by this context so far */
int  qtos_set_profiling(qtos_ctx *ctx, int on);
int  qtos_last_stats(const qtos_ctx *ctx, qtos_stats *st);
void *qtos_stream(const qtos_ctx *ctx);            /* cudaStream_t the context launches on */
/* measurement behind DESIGN.md section 4: the heightfield queries of n_groups evaluations (group_size <= 64 queries each)
 * answered from global memory and from a shared-memory tile filled by 1-D bulk asynchronous copies (TMA unit) */
int  qtos_measure_heightfield_staging(qtos_ctx *ctx, int hf_id, const double *xy, int n_groups, int group_size,
                                      double *ms_direct, double *ms_staged, double *max_diff, int *fallbacks);
/* FP64 FMA throughput of the device measured with a register-resident FMA loop, TFLOP/s */
int  qtos_measure_fp64_peak(qtos_ctx *ctx, double *tflops);
/* host only, no device needed: the assembly term streams the shape compiler builds for k_asm.  rows_dealt:
 * 0 = warp w owns panel rows 4 w .. 4 w + 3, 1 = rows dealt to the warps by term count, -1 = what qtos_create would choose.
 * out[0] terms, out[1] slots (32 x steps, padding included), out[2] slots of the slowest warp summed over the block rows,
 * out[3] steps that hold a target twice (must be 0: the sums are race-free), out[4] a hash of every target's terms IN ORDER
 * (equal hashes = the same floating-point sums, whoever owns the row), out[5] 1 if the rows were dealt */
int  qtos_assembly_table_stats(const qtos_shape *shape, int rows_dealt, unsigned long long out[6]);

#ifdef __cplusplus
}
#endif
#endif
