/*
 * towr_csv.c -- oracle: 1 kHz trajectory sampler in the 37-column layout.
 * TEST INFRASTRUCTURE.  Restates ref: src/main.cpp:83-131 (entry, getTrajectory):
 * row = [t+t_start, base xyz, base rpy, LF xyz, RF xyz, LH xyz, RH xyz,
 *        base lin vel, base Euler-rate, LF f, RF f, LH f, RH f];
 * t accumulates `t += timestep` while t <= T + 1e-4; text is default
 * ostream formatting (== "%g", 6 significant digits), comma separated.
 */
#include "towr_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#define VAL(s, node, d, k) ((s)->val[(node) * 6 + (d) * 3 + (k)])

static void locate(const orc_spline *s, double t, int *id, double *tl)
{
	const double eps = 1e-10;
	double acc = 0.0; int found = s->n_polys - 1;
	for (int i = 0; i < s->n_polys; ++i) { acc += s->dur[i]; if (acc >= t - eps) { found = i; break; } }
	double loc = t;
	for (int i = 0; i < found; ++i) loc -= s->dur[i];
	*id = found; *tl = loc;
}

static void point(const orc_spline *s, double tg, double p[3], double v[3])
{
	int id; double t;
	locate(s, tg, &id, &t);
	const double T = s->dur[id];
	for (int k = 0; k < 3; ++k) {
		const double p0 = VAL(s, id, 0, k), v0 = VAL(s, id, 1, k), p1 = VAL(s, id + 1, 0, k), v1 = VAL(s, id + 1, 1, k);
		const double C = -(3 * (p0 - p1) + T * (2 * v0 + v1)) / pow(T, 2);
		const double D = (2 * (p0 - p1) + T * (v0 + v1)) / pow(T, 3);
		p[k] = 0.0 + 1.0 * p0 + t * v0 + pow(t, 2) * C + pow(t, 3) * D;
		if (v) v[k] = 0.0 + 1 * 1.0 * v0 + 2 * t * C + 3 * pow(t, 2) * D;
	}
}

static double total_time(const orc_problem *p)
{
	double T = 0.0;
	for (int i = 0; i < p->base_lin.n_polys; ++i) T += p->base_lin.dur[i];
	return T;
}

int orc_csv_rows(const orc_problem *p, double dt)
{
	double T = total_time(p), t = 0.0; int n = 0;
	while (t <= T + 1e-4) { n++; t += dt; }
	return n;
}

void orc_sample_csv(orc_problem *p, const double *x, double dt, double *rows)
{
	orc_set_x(p, x);
	double T = total_time(p), t = 0.0; int n = 0;
	while (t <= T + 1e-4) {
		double *c = rows + (size_t)37 * n;
		double bl[3], blv[3], ba[3], bav[3];
		point(&p->base_lin, t, bl, blv);
		point(&p->base_ang, t, ba, bav);
		c[0] = t + p->inst.t_start;
		for (int i = 0; i < 3; ++i) { c[1 + i] = bl[i]; c[4 + i] = ba[i]; c[19 + i] = blv[i]; c[22 + i] = bav[i]; }
		for (int ee = 0; ee < ORC_NEE; ++ee) {
			double m[3], f[3];
			point(&p->ee_motion[ee], t, m, NULL);
			point(&p->ee_force[ee], t, f, NULL);
			for (int i = 0; i < 3; ++i) { c[ee * 3 + i + 7] = m[i]; c[ee * 3 + i + 25] = f[i]; }
		}
		n++; t += dt;
	}
}

int orc_write_csv(orc_problem *p, const double *x, double dt, const char *path)
{
	int n = orc_csv_rows(p, dt);
	double *rows = (double *)malloc(sizeof(double) * 37 * (size_t)n);
	orc_sample_csv(p, x, dt, rows);
	FILE *f = fopen(path, "w");
	if (!f) { free(rows); return -1; }
	for (int r = 0; r < n; ++r) {
		for (int i = 0; i < 36; ++i) fprintf(f, "%g,", rows[37 * r + i]);
		fprintf(f, "%g\n", rows[37 * r + 36]);
	}
	fclose(f); free(rows);
	return n;
}
