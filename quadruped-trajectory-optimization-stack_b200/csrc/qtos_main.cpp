/*
 * qtos_main.cpp -- native drop-in for the reference's `./main` (solver/towr/src/main.cpp:133-471):
 * same flags, same heightfield path relative to the working directory, same traj.csv, exit code =
 * solver status.  Host C++ over the C ABI only (include/qtos_b200.h); all numerics run on the GPU.
 *
 *   flags      -g -s -s_ang -s_vel -s_ang_vel -e1..-e4 x y z | -t -r -resolution -duration v | -n t|f
 *              (ref: main.cpp:163-306); unknown tokens are ignored; -n absent => start velocity zeroed
 *   terrain    ../data/heightfields/from_pybullet/towr_heightfield.txt   (ref: main.cpp:364)
 *   output     traj.csv, 37 columns at 1 kHz                             (ref: main.cpp:15,92-131,470)
 */
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/qtos_b200.h"

namespace {

/* the (up to) three tokens after the first occurrence of `opt`; empty when absent or last */
std::vector<std::string> grab(int argc, char **argv, const std::string &opt)
{
	std::vector<std::string> out;
	char **end = argv + argc, **it = std::find(argv, end, opt);
	if (it != end && ++it != end)
		for (int i = 0; i < 3 && it != end; ++i, ++it) out.push_back(*it);
	return out;
}

bool vec3(int argc, char **argv, const char *opt, double *dst)
{
	std::vector<std::string> t = grab(argc, argv, opt);
	if (t.empty()) return false;
	for (int i = 0; i < 3; ++i) dst[i] = std::stod(t.at(i));
	return true;
}

bool scalar(int argc, char **argv, const char *opt, double *dst)
{
	std::vector<std::string> t = grab(argc, argv, opt);
	if (t.empty()) return false;
	*dst = std::stod(t[0]);
	return true;
}

bool read_grid(const std::string &path, std::vector<double> &h, int &nx, int &ny)
{
	std::ifstream f(path);
	if (!f) return false;
	std::string line; nx = 0; ny = -1;
	while (std::getline(f, line)) {
		std::stringstream ss(line);
		double v; int n = 0;
		while (ss >> v) { h.push_back(v); ++n; if (ss.peek() == ',') ss.ignore(); }
		if (n == 0) continue;
		if (ny < 0) ny = n; else if (n != ny) return false;
		++nx;
	}
	return nx > 0 && ny > 0;
}

}  // namespace

int main(int argc, char **argv)
{
	const bool timing = getenv("QTOS_TIMING") != nullptr;      /* development: phase times on stderr */
	auto t_prev = std::chrono::steady_clock::now();
	auto lap = [&](const char *what) {
		if (!timing) return;
		auto t = std::chrono::steady_clock::now();
		std::cerr << "[qtos timing] " << what << " " << std::chrono::duration<double, std::milli>(t - t_prev).count() << " ms" << std::endl;
		t_prev = t;
	};
	qtos_shape shape; qtos_default_shape(&shape);
	qtos_problem p; memset(&p, 0, sizeof(p));
	double goal[3] = {0.5, 0.0, 0.24}, start[3] = {0.0, 0.0, 0.24}, runtime = 15.0, res = 0.1, duration = 5.0, t0 = 0.0;
	double ee[4][3];
	for (int e = 0; e < 4; ++e) { ee[e][0] = shape.nominal[e][0]; ee[e][1] = shape.nominal[e][1]; ee[e][2] = 0.0; }
	bool normalize = false, default_gait = false;
	try {
		vec3(argc, argv, "-g", goal); scalar(argc, argv, "-r", &runtime); vec3(argc, argv, "-s", start);
		vec3(argc, argv, "-s_ang", p.start_ang); vec3(argc, argv, "-s_ang_vel", p.start_ang_vel); vec3(argc, argv, "-s_vel", p.start_vel);
		std::vector<std::string> n = grab(argc, argv, "-n");
		if (!n.empty()) {
			/* the reference compares the JOINED (up to three) tokens after -n with "t" (main.cpp:66,228-232): `-n t` only
			 * normalises when it is the last thing on the command line */
			std::string joined = n[0];
			for (size_t i = 1; i < n.size(); ++i) joined += " " + n[i];
			normalize = joined == "t";
		}
		else p.start_vel[0] = p.start_vel[1] = p.start_vel[2] = 0.0;
		const char *eopt[4] = {"-e1", "-e2", "-e3", "-e4"};
		for (int e = 0; e < 4; ++e) vec3(argc, argv, eopt[e], ee[e]);
		scalar(argc, argv, "-t", &t0); scalar(argc, argv, "-resolution", &res);
		if (scalar(argc, argv, "-duration", &duration)) default_gait = true;
	} catch (const std::exception &e) {
		std::cerr << "Argument input error" << std::endl << "Error: " << e.what() << std::endl;
	}
	if (normalize) { goal[0] -= start[0]; goal[1] -= start[1]; start[0] = start[1] = 0.0; }
	memcpy(p.start_pos, start, sizeof(start)); memcpy(p.goal, goal, sizeof(goal)); memcpy(p.ee, ee, sizeof(ee));
	p.t_start = t0;
	shape.combo = default_gait ? QTOS_C0 : QTOS_CUSTOM;
	shape.duration = duration;
	if (const char *m = getenv("QTOS_MASS")) shape.mass = atof(m);

	std::vector<double> grid; int nx = 0, ny = 0;
	const std::string hf_path = "../data/heightfields/from_pybullet/towr_heightfield.txt";
	if (!read_grid(hf_path, grid, nx, ny)) {
		std::cerr << "Could not open file " << hf_path << std::endl;   /* the reference carries on into UB here */
		return 2;
	}
	lap("flags + heightfield file");
	qtos_ctx *ctx = nullptr;
	if (qtos_create(0, &shape, 1, &ctx) != QTOS_OK) { std::cerr << "qtos_create: " << qtos_last_error(nullptr) << std::endl; return 3; }
	lap("qtos_create (CUDA context, shape compile, workspace)");
	int rc = qtos_upload_heightfield(ctx, grid.data(), nx, ny, res, &p.hf_id);
	qtos_dims d; qtos_get_dims(ctx, &d);
	std::vector<double> x(d.n_vars), rows((size_t)d.csv_rows * QTOS_CSV_COLS);
	qtos_result r; memset(&r, 0, sizeof(r));
	qtos_options o; qtos_default_options(&o);
	o.max_cpu_time = runtime;   /* -r: counted in reference iterations (0.1 s each), see qtos_options.max_cpu_time */
	if (rc == QTOS_OK) rc = qtos_solve_batch(ctx, &p, 1, &o, &r, x.data(), rows.data());
	if (rc != QTOS_OK) { std::cerr << "qtos: " << qtos_last_error(ctx) << std::endl; qtos_destroy(ctx); return 3; }
	lap("upload + solve + 1 kHz sampling");
	std::cout << "Number of Iterations....: " << r.iters << std::endl;
	std::cout << "Constraint violation....: " << r.constr_viol << std::endl;
	std::cout << "status -> " << r.status << std::endl;
	if (qtos_write_csv(rows.data(), d.csv_rows, "traj.csv") != QTOS_OK) {
		std::cerr << "qtos: cannot write traj.csv" << std::endl;
		qtos_destroy(ctx);
		return 3;
	}
	lap("traj.csv");
	qtos_destroy(ctx);
	lap("qtos_destroy");
	return r.status;
}
