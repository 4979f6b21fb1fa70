/*
 * qtos_tables.h -- static per-shape tables produced by the host problem compiler
 * (qtos_compile.cpp) and consumed by the device kernels (qtos_kernels.cu).
 *
 * A "shape" is (gait combo, horizon, Parameters, model constants).  Because the reference
 * never optimises phase durations on this path (ref: solver/towr/src/main.cpp:441 commented,
 * parameters.cc IsOptimizeTimings() false) every spline sample time maps to a FIXED polynomial
 * and fixed Hermite weights, so the whole index structure of the NLP is compiled once per shape
 * and shared by every problem of a batch; only start/goal/feet/heightfield vary per problem.
 */
#ifndef QTOS_TABLES_H_
#define QTOS_TABLES_H_

#include <cstdint>
#include <vector>
#include "../../include/qtos_b200.h"

#define QTOS_NB 16                 /* block size of the block-skyline KKT storage */
#define QTOS_DYN_CANON 120         /* canonical local columns of a dynamics element */
#define QTOS_ROM_CANON 36          /* canonical local columns of a range-of-motion element */
#define QTOS_NPARAM 28             /* per-problem parameter vector, see QP_* */

/* layout of the per-problem parameter vector P[] used for x0 / fixed variables */
enum { QP_START_POS = 0, QP_START_ANG = 3, QP_START_VEL = 6, QP_START_ANGVEL = 9, QP_GOAL = 12,
       QP_EE = 15, QP_ZERO = 27 };

enum { ROW_EQ = 1, ROW_HASL = 2, ROW_HASU = 4 };
enum { EL_DYN = 0, EL_ROM = 1, EL_CONST = 2, EL_TG = 3 };   /* EL_TG: terrain / force rows with terrain gradients (values depend on x) */

struct DynSample {                 /* one dynamics sample time (6 rows) */
	int    base_id;                /* base polynomial (nodes base_id, base_id+1) */
	double W[3][4];                /* d{p,v,a}/d{p0,v0,p1,v1} of the base polynomial at the sample */
	int    mo_id[QTOS_NEE];        /* foot-motion polynomial */
	double mo_w[QTOS_NEE][4];      /* position weights */
	int    fo_id[QTOS_NEE];        /* force polynomial */
	double fo_w[QTOS_NEE][4];
	int    elem;
	int    col0;                   /* first JCol of this sample */
	int8_t slot[QTOS_DYN_CANON];   /* canonical column -> dense column of the element, -1 = absent */
};

struct RomSample {                 /* one (foot, time) range-of-motion sample (3 rows) */
	int    base_id;
	double wp[4];
	int    ee;
	int    mo_id;
	double mo_w[4];
	int    elem;
	int    col0;
	int8_t slot[QTOS_ROM_CANON];
};

struct JCol {                      /* one dense column of a dynamics / range-of-motion element */
	uint8_t kind, foot, dim, pad_[5];      /* kind: 0 base-lin, 1 base-ang, 2 foot motion, 3 foot force */
	double  w[3];                          /* Hermite weights (pos, vel, acc), merged over nodes sharing the variable */
};

struct Element {                   /* dense Jacobian block: rows [row0,row0+nrows) x cols[coloff..+ncols), column-major,
                                      column stride ld = nrows rounded up to even (zero padded, 16-byte aligned columns) */
	int row0, nrows, ncols, valoff, coloff, type, ld, pad_;
};

struct AsmCol {                    /* one staged column (element e, local column a) of a block row's assembly */
	int      voff;                 /* offset of the column's values in Jv */
	int16_t  row0;                 /* first constraint row of the element */
	uint8_t  nrows, i;             /* rows of the element; panel row (variable perm % 16) the column adds into */
};

struct uint2_t { uint32_t x, y; };   /* 64-bit table entry, read as uint2 on the device */

struct HostTables {
	qtos_shape shape;
	/* dimensions */
	int n_all = 0, n_free = 0, npad = 0, m = 0, n_eq = 0, n_ineq = 0, n_bounds = 0;
	int n_dyn = 0, n_rom = 0;
	int nJ = 0, nb = 0, nM = 0, nnz_jac = 0;
	int csv_rows = 0;
	double T = 0;
	double flops_factor = 0;
	int max_nodes = 0;             /* leading dimension of node_var */
	int n_nodes[10] = {0}, n_polys[10] = {0}, var_off[11] = {0};
	int row_off[24] = {0};
	/* splines: 0 base-lin, 1 base-ang, 2..5 foot motion, 6..9 foot force */
	std::vector<double> dur[10];
	std::vector<int16_t> node_var;          /* [10][max_nodes][6] full variable index or -1 */
	/* variables */
	std::vector<uint8_t> x0_spline, x0_deriv, x0_dim;   /* [n_all] */
	std::vector<int16_t> x0_node;                        /* node whose interpolated value wins */
	std::vector<int8_t>  fix_src;                        /* [n_all] index into P[] or -1 if free */
	std::vector<double>  cost_c;                         /* [n_all] objective f(x) = sum_v cost_c[v] x_v^2 (NodeCost terms); empty = none */
	std::vector<int16_t> perm_of_var;                    /* [n_all] permuted free index or -1 */
	std::vector<int16_t> var_of_perm;                    /* [npad] full index or -1 (padding) */
	/* rows */
	std::vector<uint8_t> row_flags;                      /* [m] */
	std::vector<double>  gl, gu;                         /* [m] */
	std::vector<int>     row_elem;                       /* [m] */
	/* elements */
	std::vector<Element> elems;
	std::vector<int16_t> elem_cols;                      /* permuted free indices */
	std::vector<double>  Jconst;                         /* [nJ] unscaled values of constant elements */
	std::vector<int16_t> jrow;                           /* [nJ] constraint row of every stored value, -1 = padding */
	/* evaluation */
	std::vector<DynSample> dyn;
	std::vector<RomSample> rom;
	std::vector<JCol>    jcols;                          /* column descriptors of dyn / rom elements */
	std::vector<int>     lin_row, lin_ptr;               /* linear rows: g = sum val * x[col] */
	std::vector<int16_t> lin_col;
	std::vector<double>  lin_val;
	std::vector<int>     ter_row;                        /* terrain rows: g = x[vz] - h(x[vx], x[vy]) */
	std::vector<int16_t> ter_var;                        /* [n_ter][3] */
	/* qtos_shape.terrain_gradients: Jacobian tasks of the terrain rows (row, vx, vy, vz, slot x, slot y, slot z, element) and
	 * the force nodes (row0, fx, fy, fz, foothold x, foothold y, slot fx, slot fy, slot fz, element); slot = column of the element
	 * or -1 (fixed variable); empty without the option */
	std::vector<int>     tg_ter, tg_frc;
	/* condensed KKT structure */
	std::vector<int>     fb, blkptr;                     /* [nb], [nb+1] (in blocks) */
	std::vector<int>     diag_off;                       /* [npad] offset of (i,i) in M */
	/* assembly of sigma I + J' D J, one block row at a time (owner-computes: warp w owns panel rows i % 4 == w):
	 * staged columns A = D J[:, a] of every (element, column a in the block row), then per warp a flat stream of
	 * terms, one per lane: n2 (2 bits, 0 = no-op) | staged column k << 2 (9) | (a - b) << 11 (7) | panel offset << 18 */
	std::vector<int>     as_ptr;                         /* [nb+1] first staged column of every block row */
	std::vector<AsmCol>  as_col;
	std::vector<int>     at_ptr;                         /* [nb*4+1] first term of (block row, warp), multiples of 32 */
	std::vector<uint32_t> at;
	int as_max = 0;                                      /* most staged columns of one block row */
	int max_w = 0, rp_ld = 0;                            /* widest block row (blocks); leading dimension of the shared-memory panel */
	long long asm_terms_total = 0;                       /* (a, b) pairs summed = scalar dot products per assembly */
	/* J' v gather (k_prepare), one thread per variable: the (element, column) terms of 32 neighbouring variables are
	 * interleaved (term s of lane l at jg_ptr[g] + 32 s + l, zero = none) so a warp reads its descriptors coalesced and
	 * keeps several columns in flight: lo = value offset of the column | rows << 20, hi = first constraint row */
	std::vector<int>     jg_ptr;                         /* [ceil(npad/32)+1], multiples of 32 */
	std::vector<uint2_t> jg;
	/* the same gather restricted to elements that hold equality rows (Jc' v: the multiplier passes of kip_solve) */
	std::vector<int>     jgc_ptr;
	std::vector<uint2_t> jgc;
	std::vector<int16_t> eq_rows, iq_rows;               /* [n_eq], [n_ineq] row indices */
	/* 1 kHz sampler */
	std::vector<double>  csv_t;                          /* [csv_rows] accumulated sample times */
	std::vector<uint8_t> csv_id;                         /* [csv_rows][10] */
	std::vector<double>  csv_tl;                         /* [csv_rows][10] */
};

/* returns 0 or QTOS_ESHAPE; err receives a message */
int qtos_compile_shape(const qtos_shape *shape, HostTables *out, char *err, int errlen);
extern int qtos_asm_rows_dealt;    /* how k_asm's panel rows go to its warps: -1 = by the shape's panel width (default), 0 = rows 4 w .. 4 w + 3,
                                      1 = dealt by term count (qtos_assembly_table_stats and development only) */

#endif
