#include "towr_oracle.h"
void orc_ipm_default_options(orc_ipm_options *o) { (void)o; }
int orc_ipm_solve(orc_problem *p, const orc_ipm_options *o, double *x, orc_ipm_result *res) { (void)p;(void)o;(void)x;(void)res; return -99; }
