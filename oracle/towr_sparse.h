/*
 * towr_sparse.h -- oracle-internal: structure of the condensed KKT matrix (free variables, CSR Jacobian
 * pattern, reverse Cuthill-McKee order, skyline) shared by towr_ipm.c and towr_ipopt.c.
 * TEST INFRASTRUCTURE (see towr_oracle.h).
 */
#ifndef TOWR_SPARSE_H_
#define TOWR_SPARSE_H_
#include "towr_oracle.h"

typedef struct {
	int n, m;              /* free variables, rows */
	int *free_of;          /* [n_all] -> free index or -1 */
	int *var_of;           /* [n] -> full index */
	int *rowptr, *col;     /* CSR over free columns (structure = reference mask) */
	int *perm, *iperm;     /* RCM: perm[new] = old free index */
	int *first;            /* skyline: first column of row i (permuted) */
	long *skyptr;          /* [n+1] */
} ipm_struct;

ipm_struct *orc_build_struct(orc_problem *p, const double *x0);
void orc_free_struct(ipm_struct *S);
/* skyline Cholesky in place, row oriented; returns 1 if a pivot had to be fixed */
int  orc_sky_chol(const ipm_struct *S, double *A);
/* solves L L' x = b in place (b in permuted order) */
void orc_sky_solve(const ipm_struct *S, const double *L, double *b);
/* the two halves: L z = b and L' x = z */
void orc_sky_fwd(const ipm_struct *S, const double *L, double *b);
void orc_sky_bwd(const ipm_struct *S, const double *L, double *b);

#endif
