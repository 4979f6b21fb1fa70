/*
 * qtos_compile.cpp -- host "problem compiler": (gait, horizon, Parameters, model) -> static tables.
 *
 * Replaces, for the QTOS path, the structure-building half of the reference:
 *   gait tables            ref: solver/towr/src/quadruped_gait_generator.cc:39-369, gait_generator.cc:54-150
 *   Parameters             ref: solver/towr/src/parameters.cc:40-135
 *   variable sets          ref: solver/towr/src/nlp_formulation.cc:63-198, nodes_variables_all.cc:45-61,
 *                               nodes_variables_phase_based.cc:38-58,197-298
 *   constraint sets/order  ref: solver/towr/src/nlp_formulation.cc:200-331, parameters.cc:55-60
 *   sample times           ref: solver/towr/src/time_discretization_constraint.cc:41-49
 *   spline segment lookup  ref: solver/towr/src/spline.cc:48-79, polynomial.cc:140-234
 * and adds what a GPU solve needs on top: dense Jacobian "elements", a bandwidth-reducing
 * ordering of the condensed KKT matrix, its block-skyline layout, and owner-computes gather
 * lists for J'DJ assembly and J'w products.
 */
#include "qtos_tables.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <queue>
#include <set>

namespace {

/* ------------------------------------------------------------------ gait */

struct Stride { std::vector<double> t; std::vector<const char *> c; };

enum Gait { Stand, Flight, Walk1, Walk2, Walk2E, Run1, Run2, Run2E, Run3, Run3E, Hop1, Hop1E, Hop2, Hop3, Hop3E, Hop5 };

/* two-letter contact code: [hind][front]; I none, P left, b right, B both */
void decode(const char *c, bool on[QTOS_NEE])
{
	on[2] = c[0] == 'P' || c[0] == 'B';
	on[3] = c[0] == 'b' || c[0] == 'B';
	on[0] = c[1] == 'P' || c[1] == 'B';
	on[1] = c[1] == 'b' || c[1] == 'B';
}

Stride drop_transition(Stride s)
{
	double last = s.t.back();
	s.t.pop_back(); s.c.pop_back();
	s.t.back() += last;
	return s;
}

Stride stride_of(Gait g)
{
	switch (g) {
	case Stand:  return {{0.3}, {"BB"}};
	case Flight: return {{0.3}, {"Bb"}};
	case Walk1:  return {{0.3, 0.2, 0.3, 0.2, 0.3, 0.2, 0.3, 0.2}, {"bB", "BB", "Bb", "BB", "PB", "BB", "BP", "BB"}};
	case Walk2:  return {{0.25, 0.13, 0.25, 0.13, 0.25, 0.13, 0.25, 0.13}, {"bB", "bb", "Bb", "Pb", "PB", "PP", "BP", "bP"}};
	case Walk2E: return drop_transition(stride_of(Walk2));
	case Run1:   return {{0.3, 0.2, 0.3, 0.2}, {"bP", "BB", "Pb", "BB"}};
	case Run2:   return {{0.4, 0.1, 0.4, 0.1}, {"bP", "II", "Pb", "II"}};
	case Run2E:  return {{0.4}, {"bP"}};
	case Run3:   return {{0.3, 0.1, 0.3, 0.1}, {"PP", "II", "bb", "II"}};
	case Run3E:  return {{0.3}, {"PP"}};
	case Hop1:   return {{0.3, 0.1, 0.3, 0.1}, {"BI", "II", "IB", "II"}};
	case Hop1E:  return {{0.3}, {"BI"}};
	case Hop2:   return {{0.3, 0.4, 0.3}, {"BB", "II", "BB"}};
	case Hop3:   return {{0.2, 0.3, 0.2, 0.2, 0.2, 0.3, 0.2, 0.2}, {"Bb", "BI", "BP", "bP", "bB", "IB", "PB", "Pb"}};
	case Hop3E:  return drop_transition(stride_of(Hop3));
	default:     return {{0.1, 0.2, 0.1, 0.1, 0.2, 0.1}, {"Bb", "BB", "IP", "Bb", "BB", "IP"}};
	}
}

bool phase_durations(int combo, double T, std::vector<double> dur[QTOS_NEE], bool contact0[QTOS_NEE])
{
	static const Gait table[6][6] = {
		{Stand, Walk2, Walk2, Walk2, Walk2E, Stand}, {Stand, Run2, Run2, Run2, Run2E, Stand},
		{Stand, Run3, Run3, Run3, Run3E, Stand},     {Stand, Hop1, Hop1, Hop1, Hop1E, Stand},
		{Stand, Hop3, Hop3, Hop3, Hop3E, Stand},     {Stand, Walk1, Walk1, Walk1, Walk2E, Stand}};
	if (combo < 0 || combo > 5) return false;
	std::vector<double> times; std::vector<const char *> codes;
	for (Gait g : table[combo]) {
		Stride s = stride_of(g);
		times.insert(times.end(), s.t.begin(), s.t.end());
		codes.insert(codes.end(), s.c.begin(), s.c.end());
	}
	const int np = (int)times.size();
	for (int ee = 0; ee < QTOS_NEE; ++ee) {
		std::vector<double> raw; double acc = 0.0;
		for (int k = 0; k + 1 < np; ++k) {
			bool a[QTOS_NEE], b[QTOS_NEE];
			decode(codes[k], a); decode(codes[k + 1], b);
			acc += times[k];
			if (a[ee] != b[ee]) { raw.push_back(acc); acc = 0.0; }
		}
		raw.push_back(acc + times[np - 1]);
		double total = 0.0;
		for (double v : raw) total += v;
		dur[ee].clear();
		for (double v : raw) dur[ee].push_back((v / total) * T);
		bool a[QTOS_NEE]; decode(codes[0], a);
		contact0[ee] = a[ee];
	}
	return true;
}

/* ------------------------------------------------------------------ splines */

struct Spl {
	std::vector<double> dur;
	std::vector<int> phase, is_const;
	std::vector<int> opt;            /* [n_nodes*6] set-local index or -1 */
	int n_vars = 0;
	int n_polys() const { return (int)dur.size(); }
	int n_nodes() const { return (int)dur.size() + 1; }
	int &o(int node, int deriv, int dim) { return opt[node * 6 + deriv * 3 + dim]; }
	bool const_node(int node) const
	{
		if (node == 0) return is_const.front();
		if (node == n_nodes() - 1) return is_const.back();
		return is_const[node - 1] || is_const[node];
	}
};

Spl make_phase_spline(const std::vector<double> &phase_dur, bool first_const, int polys_changing)
{
	Spl s; bool c = first_const;
	for (size_t i = 0; i < phase_dur.size(); ++i, c = !c) {
		int np = c ? 1 : polys_changing;
		for (int j = 0; j < np; ++j) { s.dur.push_back(phase_dur[i] / np); s.phase.push_back((int)i); s.is_const.push_back(c); }
	}
	s.opt.assign(s.n_nodes() * 6, -1);
	return s;
}

void locate(const std::vector<double> &dur, double t, int *id, double *tl)
{
	const double eps = 1e-10;
	double acc = 0.0; int found = (int)dur.size() - 1;
	for (int i = 0; i < (int)dur.size(); ++i) { acc += dur[i]; if (acc >= t - eps) { found = i; break; } }
	double loc = t;
	for (int i = 0; i < found; ++i) loc -= dur[i];
	*id = found; *tl = loc;
}

/* d{p,v,a}(t)/d{p0,v0,p1,v1} of a cubic Hermite polynomial of duration T */
void hermite_weights(double T, double t, int deriv, double w[4])
{
	const double t2 = std::pow(t, 2), t3 = std::pow(t, 3), T2 = std::pow(T, 2), T3 = std::pow(T, 3);
	if (deriv == 0) {
		w[0] = (2 * t3) / T3 - (3 * t2) / T2 + 1; w[1] = t - (2 * t2) / T + t3 / T2;
		w[2] = (3 * t2) / T2 - (2 * t3) / T3;     w[3] = t3 / T2 - t2 / T;
	} else if (deriv == 1) {
		w[0] = (6 * t2) / T3 - (6 * t) / T2;      w[1] = (3 * t2) / T2 - (4 * t) / T + 1;
		w[2] = (6 * t) / T2 - (6 * t2) / T3;      w[3] = (3 * t2) / T2 - (2 * t) / T;
	} else {
		w[0] = (12 * t) / T3 - 6 / T2;            w[1] = (6 * t) / T2 - 4 / T;
		w[2] = 6 / T2 - (12 * t) / T3;            w[3] = (6 * t) / T2 - 2 / T;
	}
}

std::vector<double> sample_times(double T, double dt)
{
	std::vector<double> t; double acc = 0.0;
	t.push_back(acc);
	for (int i = 0; i < (int)std::floor(T / dt); ++i) { acc += dt; t.push_back(acc); }
	t.push_back(T);
	return t;
}

/* ------------------------------------------------------------------ ordering */

/* reverse Cuthill-McKee; start < 0: pseudo-peripheral start node of every component (the classic choice),
 * start >= 0: that node starts the first component */
std::vector<int> rcm(int n, const std::vector<std::set<int>> &adj, int start = -1)
{
	std::vector<int> order; order.reserve(n);
	std::vector<char> seen(n, 0);
	auto bfs_far = [&](int s) {
		std::vector<int> lvl(n, -1); std::queue<int> q; q.push(s); lvl[s] = 0; int far = s;
		while (!q.empty()) {
			int u = q.front(); q.pop();
			if (lvl[u] > lvl[far] || (lvl[u] == lvl[far] && adj[u].size() < adj[far].size())) far = u;
			for (int v : adj[u]) if (lvl[v] < 0 && !seen[v]) { lvl[v] = lvl[u] + 1; q.push(v); }
		}
		return far;
	};
	while ((int)order.size() < n) {
		int s = -1;
		for (int i = 0; i < n; ++i) if (!seen[i] && (s < 0 || adj[i].size() < adj[s].size())) s = i;
		if (start >= 0 && order.empty()) s = start;
		else for (int rep = 0; rep < 4; ++rep) { int f = bfs_far(s); if (f == s) break; s = f; }
		size_t head = order.size();
		order.push_back(s); seen[s] = 1;
		while (head < order.size()) {
			int u = order[head++];
			std::vector<int> nb;
			for (int v : adj[u]) if (!seen[v]) { nb.push_back(v); seen[v] = 1; }
			std::stable_sort(nb.begin(), nb.end(), [&](int a, int b) { return adj[a].size() < adj[b].size(); });
			order.insert(order.end(), nb.begin(), nb.end());
		}
	}
	std::reverse(order.begin(), order.end());
	return order;
}

struct LinRow { int row; std::vector<std::pair<int, double>> terms; };   /* (full var, coef), merged */

void lin_add(LinRow &r, int var, double coef)
{
	if (var < 0) return;
	for (auto &t : r.terms) if (t.first == var) { t.second += coef; return; }
	r.terms.push_back({var, coef});
}

}  // namespace

int qtos_asm_rows_dealt = -1;

int qtos_compile_shape(const qtos_shape *shape, HostTables *H, char *err, int errlen)
{
	auto fail = [&](const char *msg) { std::snprintf(err, errlen, "%s", msg); return (int)QTOS_ESHAPE; };
	const qtos_shape &sh = *shape;
	H->shape = sh;
	if (!(sh.duration > 0) || !(sh.dt_base_poly > 0) || !(sh.dt_dynamic > 0) || !(sh.dt_rom > 0) || !(sh.mass > 0))
		return fail("shape: duration, dt_* and mass must be positive");
	if (sh.force_polys_per_stance < 1 || sh.ee_polys_per_swing != 2)
		return fail("shape: force_polys_per_stance >= 1 and ee_polys_per_swing == 2 required");

	std::vector<double> ph[QTOS_NEE]; bool contact0[QTOS_NEE];
	if (!phase_durations(sh.combo, sh.duration, ph, contact0)) return fail("shape: unknown gait combo");
	for (int ee = 0; ee < QTOS_NEE; ++ee)
		if (!contact0[ee]) return fail("shape: every foot must start in contact");
	double T = 0.0;
	for (double v : ph[0]) T += v;
	H->T = T;

	/* ---- splines and variable layout (ifopt order) ---- */
	Spl spl[10];
	{
		std::vector<double> bd; double left = T;
		while (left > 1e-10) { bd.push_back(left > sh.dt_base_poly ? sh.dt_base_poly : left); left -= sh.dt_base_poly; }
		for (int s = 0; s < 2; ++s) {
			spl[s].dur = bd; spl[s].phase.assign(bd.size(), 0); spl[s].is_const.assign(bd.size(), 0);
			spl[s].opt.resize(spl[s].n_nodes() * 6);
			for (int i = 0; i < spl[s].n_nodes() * 6; ++i) spl[s].opt[i] = i;
			spl[s].n_vars = spl[s].n_nodes() * 6;
		}
	}
	for (int ee = 0; ee < QTOS_NEE; ++ee) {
		Spl &mo = spl[2 + ee];
		mo = make_phase_spline(ph[ee], contact0[ee], sh.ee_polys_per_swing);
		int idx = 0;
		for (int node = 0; node < mo.n_nodes(); ++node) {
			if (!mo.const_node(node)) {
				for (int d = 0; d < 3; ++d) { mo.o(node, 0, d) = idx++; if (d != 2) mo.o(node, 1, d) = idx++; }
			} else {
				for (int d = 0; d < 3; ++d) { mo.o(node, 0, d) = idx; mo.o(node + 1, 0, d) = idx; idx++; }
				node++;
			}
		}
		mo.n_vars = idx;
		Spl &fo = spl[6 + ee];
		fo = make_phase_spline(ph[ee], !contact0[ee], sh.force_polys_per_stance);
		idx = 0;
		for (int node = 0; node < fo.n_nodes(); ++node) {
			if (!fo.const_node(node)) { for (int d = 0; d < 3; ++d) { fo.o(node, 0, d) = idx++; fo.o(node, 1, d) = idx++; } }
			else node++;
		}
		fo.n_vars = idx;
	}
	int off = 0, max_nodes = 0;
	for (int s = 0; s < 10; ++s) {
		H->var_off[s] = off; off += spl[s].n_vars;
		H->n_nodes[s] = spl[s].n_nodes(); H->n_polys[s] = spl[s].n_polys();
		H->dur[s] = spl[s].dur;
		max_nodes = std::max(max_nodes, spl[s].n_nodes());
		if (spl[s].n_polys() > 250) return fail("shape: too many polynomials (horizon too long)");
	}
	H->var_off[10] = off;
	const int n_all = off;
	if (n_all > 32000) return fail("shape: too many variables");
	H->n_all = n_all; H->max_nodes = max_nodes;
	H->node_var.assign((size_t)10 * max_nodes * 6, -1);
	auto nv = [&](int s, int node, int q) -> int16_t & { return H->node_var[((size_t)s * max_nodes + node) * 6 + q]; };
	H->x0_spline.assign(n_all, 0); H->x0_deriv.assign(n_all, 0); H->x0_dim.assign(n_all, 0); H->x0_node.assign(n_all, 0);
	for (int s = 0; s < 10; ++s)
		for (int node = 0; node < spl[s].n_nodes(); ++node)
			for (int q = 0; q < 6; ++q) {
				int o = spl[s].opt[node * 6 + q];
				if (o < 0) continue;
				int v = H->var_off[s] + o;
				nv(s, node, q) = (int16_t)v;
				/* later node overwrites: GetValues() reports the last node mapped to an index */
				H->x0_spline[v] = (uint8_t)s; H->x0_node[v] = (int16_t)node; H->x0_deriv[v] = (uint8_t)(q / 3); H->x0_dim[v] = (uint8_t)(q % 3);
			}
	/* fixed variables = equal bounds (ref: nlp_formulation.cc:110-121,151; parameters.cc:66-69) */
	H->fix_src.assign(n_all, -1);
	auto fix = [&](int s, int node, int deriv, int dim, int src) { int v = nv(s, node, deriv * 3 + dim); if (v >= 0) H->fix_src[v] = (int8_t)src; };
	{
		const int last = spl[0].n_nodes() - 1;
		for (int d = 0; d < 3; ++d) {
			fix(0, 0, 0, d, QP_START_POS + d); fix(0, 0, 1, d, QP_START_VEL + d); fix(0, last, 1, d, QP_ZERO);
			fix(1, 0, 0, d, QP_START_ANG + d); fix(1, 0, 1, d, QP_START_ANGVEL + d);
			fix(1, last, 0, d, QP_ZERO);       fix(1, last, 1, d, QP_ZERO);
		}
		fix(0, last, 0, 0, QP_GOAL + 0); fix(0, last, 0, 1, QP_GOAL + 1);
		for (int ee = 0; ee < QTOS_NEE; ++ee) for (int d = 0; d < 3; ++d) fix(2 + ee, 0, 0, d, QP_EE + 3 * ee + d);
	}
	/* optional cost terms (ref: nlp_formulation.cc:343-376, node_cost.cc:53-83): weight * value^2 per NODE, so a variable
	 * shared by several nodes collects the weight once per node */
	H->cost_c.clear();
	if (sh.cost_force_z != 0.0 || sh.cost_ee_vel_xy != 0.0) {
		H->cost_c.assign(n_all, 0.0);
		for (int ee = 0; ee < QTOS_NEE; ++ee) {
			for (int node = 0; node < spl[6 + ee].n_nodes(); ++node) { const int v = nv(6 + ee, node, 0 * 3 + 2); if (v >= 0) H->cost_c[v] += sh.cost_force_z; }
			for (int node = 0; node < spl[2 + ee].n_nodes(); ++node)
				for (int d = 0; d < 2; ++d) { const int v = nv(2 + ee, node, 1 * 3 + d); if (v >= 0) H->cost_c[v] += sh.cost_ee_vel_xy; }
		}
	}
	std::vector<int> free_of(n_all, -1), var_of_free;
	for (int v = 0; v < n_all; ++v) if (H->fix_src[v] < 0) { free_of[v] = (int)var_of_free.size(); var_of_free.push_back(v); }
	const int n_free = (int)var_of_free.size();
	H->n_free = n_free;

	/* ---- rows (ifopt order: Terrain, Dynamic, BaseAcc, EndeffectorRom, Force, Swing) ---- */
	std::vector<double> t_dyn = sample_times(T, sh.dt_dynamic), t_rom = sample_times(T, sh.dt_rom);
	H->n_dyn = (int)t_dyn.size(); H->n_rom = (int)t_rom.size();
	int row = 0, ro = 0;
	int row_ter[QTOS_NEE], row_rom[QTOS_NEE], row_force[QTOS_NEE], row_swing[QTOS_NEE];
	for (int ee = 0; ee < QTOS_NEE; ++ee) { row_ter[ee] = row; H->row_off[ro++] = row; row += spl[2 + ee].n_nodes() - 1; }
	const int row_dyn = row; H->row_off[ro++] = row; row += 6 * H->n_dyn;
	const int row_acc[2] = {row, row + 3 * (spl[0].n_polys() - 1)};
	H->row_off[ro++] = row_acc[0]; H->row_off[ro++] = row_acc[1];
	row += 6 * (spl[0].n_polys() - 1);
	for (int ee = 0; ee < QTOS_NEE; ++ee) { row_rom[ee] = row; H->row_off[ro++] = row; row += 3 * H->n_rom; }
	for (int ee = 0; ee < QTOS_NEE; ++ee) {
		int cnt = 0; for (int nd = 0; nd < spl[6 + ee].n_nodes(); ++nd) cnt += !spl[6 + ee].const_node(nd);
		row_force[ee] = row; H->row_off[ro++] = row; row += 5 * cnt;
	}
	for (int ee = 0; ee < QTOS_NEE; ++ee) {
		int cnt = 0; for (int nd = 0; nd < spl[2 + ee].n_nodes(); ++nd) cnt += !spl[2 + ee].const_node(nd);
		row_swing[ee] = row; H->row_off[ro++] = row; row += 4 * cnt;
	}
	/* optional BaseMotionConstraint (Parameters::BaseRom; ref: base_motion_constraint.cc:38-93), appended like
	 * constraints_.push_back(BaseRom) would: six rows per sample, AX AY AZ LX LY LZ */
	std::vector<double> t_brom;
	const int row_brom = row;
	if (sh.base_rom) {
		if (!(sh.dt_base_rom > 0.0)) return fail("dt_base_rom must be positive");
		t_brom = sample_times(T, sh.dt_base_rom);
		H->row_off[ro++] = row; row += 6 * (int)t_brom.size();
	}
	H->row_off[ro] = row;
	const int m = row;
	H->m = m;
	const double INF = 1e20;
	H->gl.assign(m, 0.0); H->gu.assign(m, 0.0); H->row_elem.assign(m, -1);

	/* ---- elements ---- */
	std::vector<std::vector<int>> ecols;      /* element -> free (unpermuted) columns */
	std::vector<std::vector<double>> econst;  /* element -> col-major constant values (empty for dyn/rom) */
	auto new_elem = [&](int type, int row0, int nrows, const std::vector<int> &cols) {
		Element e; e.type = type; e.row0 = row0; e.nrows = nrows; e.ncols = (int)cols.size(); e.valoff = 0; e.coloff = 0; e.ld = (nrows + 1) & ~1; e.pad_ = 0;
		H->elems.push_back(e); ecols.push_back(cols); econst.push_back({});
		for (int r = 0; r < nrows; ++r) H->row_elem[row0 + r] = (int)H->elems.size() - 1;
		return (int)H->elems.size() - 1;
	};
	std::vector<LinRow> lin;
	/* const element from a group of consecutive linear rows */
	auto const_elem = [&](const std::vector<LinRow> &rows) {
		std::vector<int> cols;
		for (auto &r : rows) for (auto &t : r.terms) {
			int f = free_of[t.first];
			if (f >= 0 && std::find(cols.begin(), cols.end(), f) == cols.end()) cols.push_back(f);
		}
		int e = new_elem(EL_CONST, rows[0].row, (int)rows.size(), cols);
		econst[e].assign(cols.size() * rows.size(), 0.0);
		for (size_t r = 0; r < rows.size(); ++r) for (auto &t : rows[r].terms) {
			int f = free_of[t.first]; if (f < 0) continue;
			size_t a = std::find(cols.begin(), cols.end(), f) - cols.begin();
			econst[e][a * rows.size() + r] += t.second;
		}
	};
	/* terrain (ref: terrain_constraint.cc:59-108): g = z - h(x,y); dh/dx = dh/dy = 0 on this path */
	for (int ee = 0; ee < QTOS_NEE; ++ee) {
		const Spl &s = spl[2 + ee];
		for (int nd = 1; nd < s.n_nodes(); ++nd) {
			int r = row_ter[ee] + nd - 1;
			if (!s.const_node(nd)) H->gu[r] = INF;
			int vx = nv(2 + ee, nd, 0), vy = nv(2 + ee, nd, 1), vz = nv(2 + ee, nd, 2);
			H->ter_row.push_back(r);
			H->ter_var.push_back((int16_t)vx); H->ter_var.push_back((int16_t)vy); H->ter_var.push_back((int16_t)vz);
			if (sh.terrain_gradients) {
				/* dg/d(x, y) = -dh/d(x, y) at the node (ref: terrain_constraint.cc:90-108 with the derivatives of custom_terrain.cpp:96-156) */
				std::vector<int> cols; int slot[3];
				const int vv[3] = {vx, vy, vz};
				for (int d = 0; d < 3; ++d) { const int f = vv[d] >= 0 ? free_of[vv[d]] : -1; slot[d] = -1; if (f >= 0) { slot[d] = (int)cols.size(); cols.push_back(f); } }
				const int e = new_elem(EL_TG, r, 1, cols);
				const int rec[8] = {r, vx, vy, vz, slot[0], slot[1], slot[2], e};
				H->tg_ter.insert(H->tg_ter.end(), rec, rec + 8);
				continue;
			}
			LinRow lr; lr.row = r; lin_add(lr, vz, 1.0);
			const_elem({lr});              /* J only; g comes from the terrain table */
		}
	}
	/* dynamics (ref: dynamic_constraint.cc:37-137) */
	H->dyn.resize(H->n_dyn);
	for (int k = 0; k < H->n_dyn; ++k) {
		DynSample &D = H->dyn[k];
		double tl; locate(spl[0].dur, t_dyn[k], &D.base_id, &tl);
		for (int d = 0; d < 3; ++d) hermite_weights(spl[0].dur[D.base_id], tl, d, D.W[d]);
		std::vector<int> cols; std::vector<int> canon_var(QTOS_DYN_CANON, -1);
		for (int q = 0; q < 12; ++q) {
			canon_var[q] = H->var_off[0] + (D.base_id + q / 6) * 6 + q % 6;
			canon_var[12 + q] = H->var_off[1] + (D.base_id + q / 6) * 6 + q % 6;
		}
		for (int ee = 0; ee < QTOS_NEE; ++ee) {
			locate(spl[2 + ee].dur, t_dyn[k], &D.mo_id[ee], &tl);
			hermite_weights(spl[2 + ee].dur[D.mo_id[ee]], tl, 0, D.mo_w[ee]);
			locate(spl[6 + ee].dur, t_dyn[k], &D.fo_id[ee], &tl);
			hermite_weights(spl[6 + ee].dur[D.fo_id[ee]], tl, 0, D.fo_w[ee]);
			for (int q = 0; q < 12; ++q) {
				canon_var[24 + ee * 24 + q] = nv(2 + ee, D.mo_id[ee] + q / 6, q % 6);
				canon_var[24 + ee * 24 + 12 + q] = nv(6 + ee, D.fo_id[ee] + q / 6, q % 6);
			}
		}
		for (int c = 0; c < QTOS_DYN_CANON; ++c) {
			D.slot[c] = -1;
			int f = canon_var[c] >= 0 ? free_of[canon_var[c]] : -1;
			if (f < 0) continue;
			size_t a = std::find(cols.begin(), cols.end(), f) - cols.begin();
			if (a == cols.size()) cols.push_back(f);
			D.slot[c] = (int8_t)a;
		}
		if (cols.size() > 127) return fail("dyn element too wide");
		D.col0 = (int)H->jcols.size();
		H->jcols.resize(H->jcols.size() + cols.size());
		for (size_t a = 0; a < cols.size(); ++a) { JCol &C = H->jcols[D.col0 + a]; memset(&C, 0, sizeof(C)); C.kind = 255; }
		for (int c = 0; c < QTOS_DYN_CANON; ++c) {
			if (D.slot[c] < 0) continue;
			JCol &C = H->jcols[D.col0 + D.slot[c]];
			const int grp = c < 12 ? 0 : (c < 24 ? 1 : (((c - 24) % 24) < 12 ? 2 : 3));
			const int foot = c < 24 ? 0 : (c - 24) / 24, q12 = c < 24 ? c % 12 : (c - 24) % 12;
			const int q = (q12 / 6) * 2 + (q12 % 6) / 3, dim = q12 % 3;
			if (C.kind != 255 && (C.kind != grp || C.foot != foot || C.dim != dim)) return fail("internal: inconsistent column merge");
			C.kind = (uint8_t)grp; C.foot = (uint8_t)foot; C.dim = (uint8_t)dim;
			if (grp < 2) { C.w[0] += D.W[0][q]; C.w[1] += D.W[1][q]; C.w[2] += D.W[2][q]; }
			else if (grp == 2) C.w[0] += D.mo_w[foot][q];
			else C.w[0] += D.fo_w[foot][q];
		}
		D.elem = new_elem(EL_DYN, row_dyn + 6 * k, 6, cols);
	}
	/* base acceleration continuity (ref: spline_acc_constraint.cc:48-80) */
	for (int w = 0; w < 2; ++w)
		for (int j = 0; j + 1 < spl[w].n_polys(); ++j)
			for (int d = 0; d < 3; ++d) {
				LinRow lr; lr.row = row_acc[w] + 3 * j + d;
				double a0[4], a1[4];
				hermite_weights(spl[w].dur[j], spl[w].dur[j], 2, a0);
				hermite_weights(spl[w].dur[j + 1], 0.0, 2, a1);
				for (int q = 0; q < 4; ++q) {
					lin_add(lr, H->var_off[w] + (j + q / 2) * 6 + (q % 2) * 3 + d, a0[q]);
					lin_add(lr, H->var_off[w] + (j + 1 + q / 2) * 6 + (q % 2) * 3 + d, -a1[q]);
				}
				lin.push_back(lr); const_elem({lr});
			}
	/* range of motion (ref: range_of_motion_constraint.cc:59-109) */
	for (int ee = 0; ee < QTOS_NEE; ++ee)
		for (int k = 0; k < H->n_rom; ++k) {
			RomSample R; double tl;
			R.ee = ee;
			locate(spl[0].dur, t_rom[k], &R.base_id, &tl);
			hermite_weights(spl[0].dur[R.base_id], tl, 0, R.wp);
			locate(spl[2 + ee].dur, t_rom[k], &R.mo_id, &tl);
			hermite_weights(spl[2 + ee].dur[R.mo_id], tl, 0, R.mo_w);
			std::vector<int> cols;
			for (int c = 0; c < QTOS_ROM_CANON; ++c) {
				int q = c % 12, var;
				if (c < 12) var = H->var_off[0] + (R.base_id + q / 6) * 6 + q % 6;
				else if (c < 24) var = H->var_off[1] + (R.base_id + q / 6) * 6 + q % 6;
				else var = nv(2 + ee, R.mo_id + q / 6, q % 6);
				/* only position weights act: velocity nodes still enter through the Hermite basis */
				R.slot[c] = -1;
				int f = var >= 0 ? free_of[var] : -1;
				if (f < 0) continue;
				size_t a = std::find(cols.begin(), cols.end(), f) - cols.begin();
				if (a == cols.size()) cols.push_back(f);
				R.slot[c] = (int8_t)a;
			}
			const int r0 = row_rom[ee] + 3 * k;
			for (int d = 0; d < 3; ++d) { H->gl[r0 + d] = sh.nominal[ee][d] - sh.max_dev[d]; H->gu[r0 + d] = sh.nominal[ee][d] + sh.max_dev[d]; }
			R.col0 = (int)H->jcols.size();
			H->jcols.resize(H->jcols.size() + cols.size());
			for (size_t a = 0; a < cols.size(); ++a) { JCol &C = H->jcols[R.col0 + a]; memset(&C, 0, sizeof(C)); C.kind = 255; }
			for (int c = 0; c < QTOS_ROM_CANON; ++c) {
				if (R.slot[c] < 0) continue;
				JCol &C = H->jcols[R.col0 + R.slot[c]];
				const int grp = c / 12, q12 = c % 12, q = (q12 / 6) * 2 + (q12 % 6) / 3, dim = q12 % 3;
				if (C.kind != 255 && (C.kind != grp || C.dim != dim)) return fail("internal: inconsistent column merge");
				C.kind = (uint8_t)grp; C.foot = (uint8_t)ee; C.dim = (uint8_t)dim;
				C.w[0] += grp < 2 ? R.wp[q] : R.mo_w[q];
			}
			R.elem = new_elem(EL_ROM, r0, 3, cols);
			H->rom.push_back(R);
		}
	/* force (ref: force_constraint.cc:67-135); terrain basis is n=ez, t1=ex, t2=ey on this path */
	for (int ee = 0; ee < QTOS_NEE; ++ee) {
		int r = row_force[ee];
		for (int nd = 0; nd < spl[6 + ee].n_nodes(); ++nd) {
			if (spl[6 + ee].const_node(nd)) continue;
			const int fx = nv(6 + ee, nd, 0), fy = nv(6 + ee, nd, 1), fz = nv(6 + ee, nd, 2);
			if (sh.terrain_gradients) {
				/* the five rows in the contact basis of the foothold = the motion node at the start of the force node's phase
				 * (ref: force_constraint.cc:67-135, nodes_variables_phase_based.cc:113-141); f . d(basis)/d(foothold) is zero
				 * because the reference's second derivatives are, so the columns stay the three force components */
				const int phase = spl[6 + ee].phase[nd == 0 ? 0 : nd - 1];
				int en = 0;
				for (int i = 0; i < spl[2 + ee].n_polys(); ++i) if (spl[2 + ee].phase[i] == phase) { en = i; break; }
				const int ex = nv(2 + ee, en, 0), ey = nv(2 + ee, en, 1);
				std::vector<int> cols; int slot[3];
				const int vv[3] = {fx, fy, fz};
				for (int d = 0; d < 3; ++d) { const int f = vv[d] >= 0 ? free_of[vv[d]] : -1; slot[d] = -1; if (f >= 0) { slot[d] = (int)cols.size(); cols.push_back(f); } }
				const int e = new_elem(EL_TG, r, 5, cols);
				const int rec[10] = {r, fx, fy, fz, ex, ey, slot[0], slot[1], slot[2], e};
				H->tg_frc.insert(H->tg_frc.end(), rec, rec + 10);
				H->gl[r] = 0.0;      H->gu[r] = sh.force_limit;
				H->gl[r + 1] = -INF; H->gu[r + 1] = 0.0;
				H->gl[r + 2] = 0.0;  H->gu[r + 2] = INF;
				H->gl[r + 3] = -INF; H->gu[r + 3] = 0.0;
				H->gl[r + 4] = 0.0;  H->gu[r + 4] = INF;
				r += 5;
				continue;
			}
			std::vector<LinRow> rows(5);
			for (int q = 0; q < 5; ++q) rows[q].row = r + q;
			lin_add(rows[0], fz, 1.0);
			lin_add(rows[1], fx, 1.0); lin_add(rows[1], fz, -sh.mu);
			lin_add(rows[2], fx, 1.0); lin_add(rows[2], fz, sh.mu);
			lin_add(rows[3], fy, 1.0); lin_add(rows[3], fz, -sh.mu);
			lin_add(rows[4], fy, 1.0); lin_add(rows[4], fz, sh.mu);
			H->gl[r] = 0.0;      H->gu[r] = sh.force_limit;
			H->gl[r + 1] = -INF; H->gu[r + 1] = 0.0;
			H->gl[r + 2] = 0.0;  H->gu[r + 2] = INF;
			H->gl[r + 3] = -INF; H->gu[r + 3] = 0.0;
			H->gl[r + 4] = 0.0;  H->gu[r + 4] = INF;
			for (auto &lr : rows) lin.push_back(lr);
			const_elem(rows);
			r += 5;
		}
	}
	/* swing (ref: swing_constraint.cc:59-108) */
	for (int ee = 0; ee < QTOS_NEE; ++ee) {
		int r = row_swing[ee];
		for (int nd = 0; nd < spl[2 + ee].n_nodes(); ++nd) {
			if (spl[2 + ee].const_node(nd)) continue;
			std::vector<LinRow> rows(4);
			for (int d = 0; d < 2; ++d) {
				const int cur_p = nv(2 + ee, nd, d), cur_v = nv(2 + ee, nd, 3 + d);
				const int prev = nv(2 + ee, nd - 1, d), next = nv(2 + ee, nd + 1, d);
				LinRow &rp = rows[2 * d], &rv = rows[2 * d + 1];
				rp.row = r + 2 * d; rv.row = r + 2 * d + 1;
				lin_add(rp, cur_p, 1.0); lin_add(rp, next, -0.5); lin_add(rp, prev, -0.5);
				lin_add(rv, cur_v, 1.0); lin_add(rv, next, -1.0 / sh.t_swing_avg); lin_add(rv, prev, 1.0 / sh.t_swing_avg);
			}
			for (auto &lr : rows) lin.push_back(lr);
			const_elem(rows);
			r += 4;
		}
	}
	/* base motion (ref: base_motion_constraint.cc:47-86): rows AX, AY = roll, pitch within +-0.01 rad; AZ, LX, LY unbounded; LZ within
	 * [z_init - 0.02, z_init + 0.1] with z_init the base spline's initial height.  The bounds of a shape are shared by all its
	 * problems, so the LZ row is stated as z(t) - z(0) in [-0.02, 0.1]: z(0) is the (fixed) start-height variable */
	for (size_t k = 0; k < t_brom.size(); ++k)
		for (int w = 0; w < 2; ++w) {               /* w = 0: angular rows first (AX AY AZ), then linear */
			const int sp = w ? 0 : 1;
			int id; double tl, wt[4];
			locate(spl[sp].dur, t_brom[k], &id, &tl);
			hermite_weights(spl[sp].dur[id], tl, 0, wt);
			for (int d = 0; d < 3; ++d) {
				LinRow lr; lr.row = row_brom + 6 * (int)k + 3 * w + d;
				for (int q = 0; q < 4; ++q) lin_add(lr, H->var_off[sp] + (id + q / 2) * 6 + (q % 2) * 3 + d, wt[q]);
				if (w == 1 && d == 2) { lin_add(lr, H->var_off[0] + 2, -1.0); H->gl[lr.row] = -0.02; H->gu[lr.row] = 0.1; }
				else if (w == 0 && d < 2) { H->gl[lr.row] = -0.01; H->gu[lr.row] = 0.01; }
				else { H->gl[lr.row] = -INF; H->gu[lr.row] = INF; }
				lin.push_back(lr); const_elem({lr});
			}
		}
	for (int r = 0; r < m; ++r) if (H->row_elem[r] < 0) return fail("internal: row without element");
	/* row flags */
	H->row_flags.assign(m, 0); H->n_eq = H->n_ineq = H->n_bounds = 0;
	for (int r = 0; r < m; ++r) {
		if (H->gl[r] == H->gu[r]) { H->row_flags[r] = ROW_EQ; H->n_eq++; continue; }
		H->n_ineq++;
		if (H->gl[r] > -1e19) { H->row_flags[r] |= ROW_HASL; H->n_bounds++; }
		if (H->gu[r] < 1e19) { H->row_flags[r] |= ROW_HASU; H->n_bounds++; }
	}
	/* linear-row table */
	H->lin_ptr.push_back(0);
	for (auto &lr : lin) {
		H->lin_row.push_back(lr.row);
		for (auto &t : lr.terms) { H->lin_col.push_back((int16_t)t.first); H->lin_val.push_back(t.second); }
		H->lin_ptr.push_back((int)H->lin_col.size());
	}

	/* ---- ordering of the condensed KKT matrix ---- */
	std::vector<std::set<int>> adj(n_free);
	for (auto &c : ecols) for (int a : c) for (int b : c) if (a != b) adj[a].insert(b);
	const int NB = QTOS_NB;
	const int npad = ((n_free + NB - 1) / NB) * NB, nb = npad / NB;
	/* the factorization pays per 16x16 block (barriers, second product) and per block product of the left-looking
	 * sweep, not per scalar of the envelope: among the RCM orderings from every start node keep the one with the
	 * cheapest BLOCK skyline (ties: the classic pseudo-peripheral start, then the lowest node) */
	auto block_cost = [&](const std::vector<int> &ord) {
		std::vector<int> ps(n_free), fst(npad);
		for (int i = 0; i < n_free; ++i) ps[ord[i]] = i;
		for (int i = 0; i < npad; ++i) fst[i] = i;
		for (auto &c : ecols) {
			int mn = npad;
			for (int a : c) mn = std::min(mn, ps[a]);
			for (int a : c) fst[ps[a]] = std::min(fst[ps[a]], mn);
		}
		std::vector<int> fbk(nb);
		long blocks = 0, macs = 0;
		for (int I = 0; I < nb; ++I) {
			int mn = npad;
			for (int i = I * NB; i < (I + 1) * NB; ++i) mn = std::min(mn, fst[i]);
			fbk[I] = mn / NB; blocks += I - fbk[I] + 1;
			for (int J = fbk[I]; J <= I; ++J) macs += J - std::max(fbk[I], fbk[J]);
		}
		return macs + 2 * blocks;
	};
	std::vector<int> order = rcm(n_free, adj);           /* order[new] = old free index */
	{
		long best = block_cost(order);
		/* neighbours pre-sorted by (degree, index): a Cuthill-McKee sweep from a given start is then a plain BFS, the
		 * same visiting order rcm() produces.  Every fourth node is tried (the nodes of one spline node share their
		 * neighbourhood and give the same skyline) */
		std::vector<std::vector<int>> nbrs(n_free);
		for (int u = 0; u < n_free; ++u) {
			nbrs[u].assign(adj[u].begin(), adj[u].end());
			std::stable_sort(nbrs[u].begin(), nbrs[u].end(), [&](int a, int b) { return adj[a].size() < adj[b].size(); });
		}
		std::vector<int> cand; std::vector<char> seen;
		for (int st = 0; st < n_free; st += 4) {
			cand.clear(); seen.assign(n_free, 0);
			cand.push_back(st); seen[st] = 1;
			for (size_t head = 0; head < cand.size(); ++head)
				for (int v : nbrs[cand[head]]) if (!seen[v]) { seen[v] = 1; cand.push_back(v); }
			if ((int)cand.size() != n_free) continue;            /* disconnected pattern: keep the default ordering */
			std::reverse(cand.begin(), cand.end());
			const long c = block_cost(cand);
			if (c < best) { best = c; order = cand; }
		}
	}
	/* refinement at block granularity: the nodes of two neighbouring blocks are re-dealt so that those whose leftmost
	 * connection lies further left come first (only moves ACROSS a block boundary change the block skyline); a re-deal
	 * is kept when the block cost drops.  Deterministic, a few passes. */
	auto refine = [&](std::vector<int> &ord) {
		long best = block_cost(ord);
		std::vector<int> ps(n_free), left(n_free);
		auto leftmost = [&]() {                                  /* leftmost connection of every node under ord */
			for (int i = 0; i < n_free; ++i) ps[ord[i]] = i;
			std::fill(left.begin(), left.end(), n_free);
			for (auto &c : ecols) {
				int mn = n_free;
				for (int a : c) mn = std::min(mn, ps[a]);
				for (int a : c) left[a] = std::min(left[a], mn);
			}
		};
		leftmost();
		for (int pass = 0; pass < 8; ++pass) {
			bool improved = false;
			for (int I = 0; I + 1 < nb; ++I) {
				const int lo = I * NB, hi = std::min(n_free, (I + 2) * NB);
				if (hi - lo <= NB) continue;
				std::vector<int> win(ord.begin() + lo, ord.begin() + hi), cand = ord;
				std::stable_sort(win.begin(), win.end(), [&](int a, int b) { return std::min(left[a], lo) < std::min(left[b], lo); });
				std::copy(win.begin(), win.end(), cand.begin() + lo);
				const long c = block_cost(cand);
				if (c < best) { best = c; ord.swap(cand); improved = true; leftmost(); }
			}
			if (!improved) break;
		}
		return best;
	};
	{
		long best = refine(order);
		/* second family: a spectral ordering.  Neighbour averaging x <- D^-1 A x with the stationary component removed
		 * is power iteration towards the Fiedler direction of the graph; started from the positions of the RCM ordering
		 * it settles within a few dozen sweeps (cheap: one pass over the pattern each).  Both directions are refined and
		 * the cheapest block skyline of the three wins (S2: spectral, S5: RCM). */
		std::vector<double> xs(n_free), ys(n_free);
		for (int i = 0; i < n_free; ++i) xs[order[i]] = (double)i;
		double degsum = 0.0;
		for (int u = 0; u < n_free; ++u) degsum += (double)adj[u].size();
		std::vector<std::vector<double>> snaps;              /* the embedding after 50 (settled), 1 and 3 sweeps */
		std::vector<double> snap1, snap3;
		for (int it = 0; it < 50 && degsum > 0.0; ++it) {
			double wm = 0.0, mx = 0.0;
			for (int u = 0; u < n_free; ++u) {
				double acc = 0.0;
				for (int v : adj[u]) acc += xs[v];
				ys[u] = adj[u].empty() ? xs[u] : acc / (double)adj[u].size();
				wm += (double)adj[u].size() * ys[u];
			}
			wm /= degsum;
			for (int u = 0; u < n_free; ++u) { ys[u] -= wm; mx = std::max(mx, std::fabs(ys[u])); }
			if (!(mx > 0.0)) break;
			for (int u = 0; u < n_free; ++u) xs[u] = ys[u] / mx;
			if (it == 0) snap1 = xs;
			if (it == 2) snap3 = xs;
		}
		snaps.push_back(xs); snaps.push_back(snap1); snaps.push_back(snap3);
		const long rcm_best = best;
		for (size_t k = 0; k < snaps.size(); ++k) {
			if (snaps[k].empty()) continue;
			if (k > 0 && best == rcm_best) break;              /* the settled embedding lost to RCM (wide shapes): the half-settled ones are not tried */
			const std::vector<double> &xv = snaps[k];
			for (int dir = 0; dir < 2; ++dir) {
				std::vector<int> cand(n_free);
				for (int i = 0; i < n_free; ++i) cand[i] = i;
				std::stable_sort(cand.begin(), cand.end(), [&](int a, int b) { return dir ? xv[a] > xv[b] : xv[a] < xv[b]; });
				if (block_cost(cand) * 100 > best * 115) continue;   /* hopeless before refinement (wide shapes): not refined */
				const long c = refine(cand);
				if (c < best) { best = c; order = cand; }
			}
		}
	}
	std::vector<int> pos(n_free);
	for (int i = 0; i < n_free; ++i) pos[order[i]] = i;
	H->npad = npad; H->nb = nb;
	H->perm_of_var.assign(n_all, -1); H->var_of_perm.assign(npad, -1);
	for (int f = 0; f < n_free; ++f) { H->perm_of_var[var_of_free[f]] = (int16_t)pos[f]; H->var_of_perm[pos[f]] = (int16_t)var_of_free[f]; }
	std::vector<int> first(npad);
	for (int i = 0; i < npad; ++i) first[i] = i;
	for (auto &c : ecols) {
		int mn = npad;
		for (int a : c) mn = std::min(mn, pos[a]);
		for (int a : c) first[pos[a]] = std::min(first[pos[a]], mn);
	}
	H->flops_factor = 0;
	for (int i = 0; i < n_free; ++i) { double w = i - first[i]; H->flops_factor += w * w; }
	H->fb.assign(nb, 0); H->blkptr.assign(nb + 1, 0);
	for (int I = 0; I < nb; ++I) {
		int mn = npad;
		for (int i = I * NB; i < (I + 1) * NB; ++i) mn = std::min(mn, first[i]);
		H->fb[I] = mn / NB;
		H->blkptr[I + 1] = H->blkptr[I] + (I - H->fb[I] + 1);
	}
	H->nM = H->blkptr[nb] * NB * NB;
	auto m_off = [&](int i, int j) {    /* i >= j, inside the block skyline */
		int I = i / NB, J = j / NB;
		return (H->blkptr[I] + J - H->fb[I]) * NB * NB + (i % NB) * NB + (j % NB);
	};
	H->diag_off.resize(npad);
	for (int i = 0; i < npad; ++i) H->diag_off[i] = m_off(i, i);

	/* ---- element columns in permuted order (assembly visits a prefix of an element's columns) ---- */
	{
		std::vector<int> jc0(ecols.size(), -1);
		for (auto &D : H->dyn) jc0[D.elem] = D.col0;
		for (auto &R : H->rom) jc0[R.elem] = R.col0;
		for (size_t e = 0; e < ecols.size(); ++e) {
			const int nc = (int)ecols[e].size(), nr = H->elems[e].nrows;
			std::vector<int> sig(nc), inv(nc);
			for (int a = 0; a < nc; ++a) sig[a] = a;
			std::stable_sort(sig.begin(), sig.end(), [&](int a, int b) { return pos[ecols[e][a]] < pos[ecols[e][b]]; });
			for (int a = 0; a < nc; ++a) inv[sig[a]] = a;
			std::vector<int> nc_(nc);
			for (int a = 0; a < nc; ++a) nc_[a] = ecols[e][sig[a]];
			ecols[e] = nc_;
			if (!econst[e].empty()) {
				std::vector<double> v(econst[e].size());
				for (int a = 0; a < nc; ++a) for (int r = 0; r < nr; ++r) v[(size_t)a * nr + r] = econst[e][(size_t)sig[a] * nr + r];
				econst[e] = v;
			}
			if (jc0[e] >= 0) {
				std::vector<JCol> v(nc);
				for (int a = 0; a < nc; ++a) v[a] = H->jcols[jc0[e] + sig[a]];
				for (int a = 0; a < nc; ++a) H->jcols[jc0[e] + a] = v[a];
			}
			if (H->elems[e].type == EL_DYN) { for (auto &D : H->dyn) if (D.elem == (int)e) for (auto &sl : D.slot) if (sl >= 0) sl = (int8_t)inv[sl]; }
			if (H->elems[e].type == EL_ROM) { for (auto &R : H->rom) if (R.elem == (int)e) for (auto &sl : R.slot) if (sl >= 0) sl = (int8_t)inv[sl]; }
			if (H->elems[e].type == EL_TG) {
				for (size_t k = 0; k + 8 <= H->tg_ter.size(); k += 8) if (H->tg_ter[k + 7] == (int)e) for (int d = 4; d < 7; ++d) if (H->tg_ter[k + d] >= 0) H->tg_ter[k + d] = inv[H->tg_ter[k + d]];
				for (size_t k = 0; k + 10 <= H->tg_frc.size(); k += 10) if (H->tg_frc[k + 9] == (int)e) for (int d = 6; d < 9; ++d) if (H->tg_frc[k + d] >= 0) H->tg_frc[k + d] = inv[H->tg_frc[k + d]];
			}
		}
	}

	/* ---- element storage, gather lists ---- */
	int valoff = 0; H->nnz_jac = 0;
	for (size_t e = 0; e < H->elems.size(); ++e) {
		Element &E = H->elems[e];
		E.valoff = valoff; E.coloff = (int)H->elem_cols.size();
		for (int a : ecols[e]) H->elem_cols.push_back((int16_t)pos[a]);
		valoff += E.ld * E.ncols;
		H->nnz_jac += E.nrows * E.ncols;
		if (E.ncols > 255 || e > 65535) return fail("element table overflow");
	}
	H->nJ = valoff;
	H->Jconst.assign(H->nJ > 0 ? H->nJ : 1, 0.0);
	for (size_t e = 0; e < H->elems.size(); ++e) {
		const Element &E = H->elems[e];
		if (econst[e].empty()) continue;
		for (int a = 0; a < E.ncols; ++a) for (int r = 0; r < E.nrows; ++r) H->Jconst[E.valoff + a * E.ld + r] = econst[e][(size_t)a * E.nrows + r];
	}
	H->jrow.assign(H->nJ > 0 ? H->nJ : 1, -1);
	for (size_t e = 0; e < H->elems.size(); ++e) {
		const Element &E = H->elems[e];
		for (int a = 0; a < E.ncols; ++a) for (int r = 0; r < E.nrows; ++r) H->jrow[E.valoff + a * E.ld + r] = (int16_t)(E.row0 + r);
	}
	if (H->nJ >= (1 << 20)) return fail("Jacobian value array too large for packed assembly terms");
	{
		std::vector<std::vector<uint32_t>> jt(npad);
		for (size_t e = 0; e < H->elems.size(); ++e)
			for (size_t a = 0; a < ecols[e].size(); ++a) jt[pos[ecols[e][a]]].push_back((uint32_t)(e << 8) | (uint32_t)a);
		/* assembly lists: warp w of the assembly kernel owns panel rows 4 w .. 4 w + 3 of every block row and walks a
		 * flat stream of terms (element e, row column a, column b <= a in permuted order), 32 per step, one per
		 * lane.  Inside a step all targets are distinct (steps are closed early with no-op terms otherwise),
		 * so panel[i][perm(b)] += A_a . J_b is race-free and its summation order is fixed. */
		int max_w = 0;
		for (int I = 0; I < nb; ++I) max_w = std::max(max_w, I - H->fb[I] + 1);
		H->max_w = max_w; H->rp_ld = max_w * NB + 8;    /* rows 64 bytes apart modulo 128: conflict-free fragment loads and stores (SWZ in qtos_kernels.cu) */
		if (16 * H->rp_ld >= (1 << 13)) return fail("assembly: panel too wide for packed terms");
		H->as_ptr.assign(1, 0); H->at_ptr.assign(1, 0); H->as_col.clear(); H->at.clear();
		H->as_max = 0; H->asm_terms_total = 0;
		long long asm_cost = 0, asm_cost_even = 0;
		for (int I = 0; I < nb; ++I) {
			const int s0 = (int)H->as_col.size();
			/* the CTA of a block row ends with its slowest warp (ncu had 15 % of the kernel's stall samples on that barrier when warp w
			 * simply took rows 4 w .. 4 w + 3), so the panel rows are DEALT to the warps by their term counts, longest first to the
			 * least loaded warp, in units of two adjacent rows (adjacent rows are mostly components of one spline node: the same
			 * elements touch them, an element's terms for the warp stay a (columns b) x (rows a) rectangle and fewer steps are cut
			 * short).  Slowest warp's stream against rows 4 w .. 4 w + 3, with the step packing below: S2 -24 %, S5 -27 %.
			 * Measured on the B200 (k_asm per bench step): S2 20.3 -> 19.6 ms, S5 (2048 windows) 32.5 -> 29.1 ms.  Which warp owns
			 * a row changes neither the terms of a target nor their order (qtos_assembly_table_stats, tests/test_host_logic.py). */
			const bool deal_rows = qtos_asm_rows_dealt >= 0 ? qtos_asm_rows_dealt != 0 : !getenv("QTOS_ASM_ROWS_EVEN");
			int owner[NB];
			{
				long long cnt[NB] = {0}, load[4] = {0, 0, 0, 0}, even[4] = {0, 0, 0, 0};
				for (size_t e = 0; e < H->elems.size(); ++e)
					for (size_t a = 0; a < ecols[e].size(); ++a) {
						const int pa = pos[ecols[e][a]];
						if (pa / NB == I) cnt[pa % NB] += (long long)a + 1;
					}
				const int unit = getenv("QTOS_ASM_DEAL_UNIT") ? atoi(getenv("QTOS_ASM_DEAL_UNIT")) : 2, nu = NB / unit;
				int order[NB]; long long ucnt[NB] = {0};
				for (int i = 0; i < NB; ++i) { ucnt[i / unit] += cnt[i]; even[i / 4] += cnt[i]; }
				for (int u = 0; u < nu; ++u) order[u] = u;
				std::stable_sort(order, order + nu, [&](int x, int y) { return ucnt[x] > ucnt[y]; });
				for (int q = 0; q < nu; ++q) {
					int best = 0;
					for (int w = 1; w < 4; ++w) if (load[w] < load[best]) best = w;
					for (int i = order[q] * unit; i < (order[q] + 1) * unit; ++i) owner[i] = best;
					load[best] += ucnt[order[q]];
				}
				if (!deal_rows) for (int i = 0; i < NB; ++i) owner[i] = i / 4;
				asm_cost += *std::max_element(load, load + 4); asm_cost_even += *std::max_element(even, even + 4);
			}
			/* a warp's steps are packed from the next pack_window pending terms: a term whose target the step already holds waits for
			 * the next step and keeps its place in the queue (so every target still gets its terms in element order) instead of
			 * cutting the step short -- 19 % of the slots were padding that way, 10 % are with a window of 96 terms (wider windows mix
			 * more elements into a step and lose again: 160 measured no better) */
			const int pack_window = getenv("QTOS_ASM_PACK_WINDOW") ? atoi(getenv("QTOS_ASM_PACK_WINDOW")) : 96;
			for (int w = 0; w < 4; ++w) {
				std::set<int> in_step;
				std::vector<std::pair<uint32_t, int>> wt;          /* (descriptor, target) of the warp's terms in canonical order (packed below) */
				auto close_step = [&]() { while (H->at.size() % 32) H->at.push_back(0u); in_step.clear(); };
				for (size_t e = 0; e < H->elems.size(); ++e) {
					const Element &E = H->elems[e];
					const std::vector<int> &c = ecols[e];
					/* the columns of this element that land in the warp's four panel rows, staged in order */
					std::vector<int> own, own_k;
					for (size_t a = 0; a < c.size(); ++a) {
						const int pa = pos[c[a]];
						if (pa / NB != I || owner[pa % NB] != w) continue;
						const int k = (int)H->as_col.size() - s0;
						if (k > 511) return fail("assembly: too many staged columns in a block row");
						AsmCol A; A.voff = E.valoff + (int)a * E.ld; A.row0 = (int16_t)E.row0; A.nrows = (uint8_t)E.nrows; A.i = (uint8_t)(pa % NB);
						H->as_col.push_back(A);
						own.push_back((int)a); own_k.push_back(k);
					}
					if (own.empty()) continue;
					/* terms in (b, a) order: 32 consecutive lanes cover ~32/|own| consecutive columns b times all own
					 * columns a, so a step touches few distinct J columns (B operand, global) and few staged columns */
					for (int b = 0; b <= own.back(); ++b)
						for (size_t j = 0; j < own.size(); ++j) {
							const int a = own[j];
							if (b > a) continue;                         /* columns are sorted: pos[c[b]] <= pos[c[a]] */
							const int pa = pos[c[a]];
							const int off = (pa % NB) * H->rp_ld + pos[c[b]] - H->fb[I] * NB;
							if (a - b > 127) return fail("assembly: element too wide for packed terms");
							const uint32_t desc = (uint32_t)(E.ld / 2) | ((uint32_t)own_k[j] << 2) | ((uint32_t)(a - b) << 11) | ((uint32_t)off << 18);
							H->asm_terms_total++;
							if (pack_window > 0) { wt.push_back({desc, off}); continue; }
							if (in_step.count(off)) close_step();
							in_step.insert(off);
							H->at.push_back(desc);
							if (H->at.size() % 32 == 0) in_step.clear();
						}
				}
				if (pack_window > 0) {
					std::vector<char> taken(wt.size(), 0);
					size_t head = 0;
					while (head < wt.size()) {
						in_step.clear();
						int n_in = 0;
						for (size_t q = head; q < wt.size() && q < head + (size_t)pack_window && n_in < 32; ++q) {
							if (taken[q] || in_step.count(wt[q].second)) { if (!taken[q]) in_step.insert(wt[q].second); continue; }
							in_step.insert(wt[q].second);
							taken[q] = 1; H->at.push_back(wt[q].first); ++n_in;
						}
						while (head < wt.size() && taken[head]) ++head;
						while (H->at.size() % 32) H->at.push_back(0u);
					}
				}
				close_step();
				H->at_ptr.push_back((int)H->at.size());
			}
			H->as_ptr.push_back((int)H->as_col.size());
			H->as_max = std::max(H->as_max, (int)H->as_col.size() - s0);
		}
		if (getenv("QTOS_COMPILE_STATS")) fprintf(stderr, "assembly: %lld terms, slowest-warp terms summed over block rows %lld (rows 4 w .. 4 w + 3: %lld), a quarter of the terms %lld\n",
		                                          (long long)H->asm_terms_total, asm_cost, asm_cost_even, (long long)H->asm_terms_total / 4);
		for (int pass = 0; pass < 2; ++pass) {             /* 0: all terms (J' v), 1: elements with equality rows only (Jc' v) */
			std::vector<int> &ptr = pass ? H->jgc_ptr : H->jg_ptr;
			std::vector<uint2_t> &tab = pass ? H->jgc : H->jg;
			ptr.assign(1, 0); tab.clear();
			std::vector<std::vector<uint32_t>> sel(npad);
			for (int i = 0; i < npad; ++i)
				for (uint32_t t : jt[i]) {
					const Element &E = H->elems[t >> 8];
					bool has_eq = false;
					for (int r = 0; r < E.nrows; ++r) has_eq = has_eq || (H->row_flags[E.row0 + r] & ROW_EQ);
					if (!pass || has_eq) sel[i].push_back(t);
				}
			for (int g = 0; g * 32 < npad; ++g) {
				size_t ns = 0;
				for (int l = 0; l < 32 && g * 32 + l < npad; ++l) ns = std::max(ns, sel[g * 32 + l].size());
				for (size_t st = 0; st < ns; ++st)
					for (int l = 0; l < 32; ++l) {
						const int i = g * 32 + l;
						uint2_t d; d.x = d.y = 0;
						if (i < npad && st < sel[i].size()) {
							const Element &E = H->elems[sel[i][st] >> 8];
							d.x = (uint32_t)(E.valoff + (int)(sel[i][st] & 255u) * E.ld) | ((uint32_t)E.nrows << 20);
							d.y = (uint32_t)E.row0;
						}
						tab.push_back(d);
					}
				ptr.push_back((int)tab.size());
			}
		}
		for (int i = 0; i < H->m; ++i) ((H->row_flags[i] & ROW_EQ) ? H->eq_rows : H->iq_rows).push_back((int16_t)i);
	}

	/* ---- 1 kHz sampler (ref: main.cpp:92-131): t accumulates += 0.001 while t <= T + 1e-4 ---- */
	{
		double Tb = 0.0; for (double d : spl[0].dur) Tb += d;
		double t = 0.0;
		while (t <= Tb + 1e-4) { H->csv_t.push_back(t); t += 0.001; }
		H->csv_rows = (int)H->csv_t.size();
		H->csv_id.resize((size_t)H->csv_rows * 10); H->csv_tl.resize((size_t)H->csv_rows * 10);
		for (int r = 0; r < H->csv_rows; ++r)
			for (int s = 0; s < 10; ++s) {
				int id; double tl; locate(spl[s].dur, H->csv_t[r], &id, &tl);
				H->csv_id[(size_t)r * 10 + s] = (uint8_t)id; H->csv_tl[(size_t)r * 10 + s] = tl;
			}
	}
	return QTOS_OK;
}
