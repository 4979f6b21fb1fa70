/*
 * towr_terrain.c -- oracle: bilinear heightfield query.  TEST INFRASTRUCTURE.
 * Restates ref: src/custom_terrain.cpp:51-94 and include/towr/terrain/custom_terrain.hpp:32-35
 * (offsets -1/-1/0, z scale 1).  GetHeightDerivWrtX/Y are hard zero in the
 * reference (ref: src/custom_terrain.cpp:96-159), so the terrain basis is
 * always n=(0,0,1), t1=(1,0,0), t2=(0,1,0) (ref: src/height_map.cc:95-141).
 *
 * Compile with -ffp-contract=off: the GPU kernel reproduces this operation
 * order with explicit __dmul_rn/__dadd_rn so heights are bit-identical.
 */
#include "towr_oracle.h"
#include <math.h>
#include <stddef.h>

#define MESH_X_OFFSET (-1.0)
#define MESH_Y_OFFSET (-1.0)

/* static_cast<size_t>(negative double) wraps to a huge value on x86-64, so
 * std::min(..., size-1) clamps it to the LAST cell; NaN/huge clamp likewise. */
static long long clamp_index(double fl, long long size)
{
	if (!(fl >= 0.0)) return size - 1;
	if (fl >= (double)(size - 1)) return size - 1;
	return (long long)fl;
}

void orc_height_cell(const orc_heightfield *hf, double x, double y, long long idx[4])
{
	const double xf = floor((x - MESH_X_OFFSET) / hf->res);
	const double yf = floor((y - MESH_Y_OFFSET) / hf->res);
	idx[0] = clamp_index(xf, hf->nx);
	idx[1] = clamp_index(yf, hf->ny);
	idx[2] = idx[0] + 1 < hf->nx - 1 ? idx[0] + 1 : hf->nx - 1;
	idx[3] = idx[1] + 1 < hf->ny - 1 ? idx[1] + 1 : hf->ny - 1;
}

double orc_height(const orc_heightfield *hf, double x, double y)
{
	long long c[4];
	orc_height_cell(hf, x, y, c);
	const double res = hf->res;
	const double x0 = (double)c[0] * res + MESH_X_OFFSET;
	const double x1 = (double)c[2] * res + MESH_X_OFFSET;
	const double y0 = (double)c[1] * res + MESH_Y_OFFSET;
	const double y1 = (double)c[3] * res + MESH_Y_OFFSET;
	const double z00 = hf->h[c[0] * hf->ny + c[1]] * 1.0 + 0.0;
	const double z01 = hf->h[c[0] * hf->ny + c[3]] * 1.0 + 0.0;
	const double z10 = hf->h[c[2] * hf->ny + c[1]] * 1.0 + 0.0;
	const double z11 = hf->h[c[2] * hf->ny + c[3]] * 1.0 + 0.0;
	/* z = 1/(dx*dy) * u^T * A * v, evaluated left to right:
	 * ((s*u)^T A) v   ref: src/custom_terrain.cpp:91 */
	const double s = 1 / (res * res);
	const double u0 = s * (x1 - x), u1 = s * (x - x0);
	const double w0 = u0 * z00 + u1 * z10;
	const double w1 = u0 * z01 + u1 * z11;
	return w0 * (y1 - y) + w1 * (y - y0);
}

/* First derivatives of the bilinear surface: the code the reference carries commented out in GetHeightDerivWrtX / WrtY
 * (ref: src/custom_terrain.cpp:101-124,133-156), same cell selection as the height.  Used only with orc_shape.terrain_gradients
 * (SURVEY 8f rank 4); second derivatives stay zero like the reference's HeightMap base class. */
void orc_height_deriv(const orc_heightfield *hf, double x, double y, double *hx, double *hy)
{
	long long c[4];
	orc_height_cell(hf, x, y, c);
	const double res = hf->res;
	const double x0 = (double)c[0] * res + MESH_X_OFFSET, x1 = (double)c[2] * res + MESH_X_OFFSET;
	const double y0 = (double)c[1] * res + MESH_Y_OFFSET, y1 = (double)c[3] * res + MESH_Y_OFFSET;
	const double z00 = hf->h[c[0] * hf->ny + c[1]], z01 = hf->h[c[0] * hf->ny + c[3]];
	const double z10 = hf->h[c[2] * hf->ny + c[1]], z11 = hf->h[c[2] * hf->ny + c[3]];
	const double s = 1 / (res * res);
	*hx = s * ((-z00 + z10) * (y1 - y) + (-z01 + z11) * (y - y0));
	*hy = s * (z00 * (x - x1) + z10 * (x0 - x) + z01 * (x1 - x) + z11 * (x - x0));
}
