"""CPU tier: host logic of the product (no compute calls): the C-ABI library loads and exports every
symbol include/qtos_b200.h declares, the ./main command-line mirror, the heightfield producer mirror."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

import qtos_b200 as Q
from qtos_b200 import heightfield as HF
from qtos_b200 import towr_cli, workloads
from conftest import GOLDEN, ROOT


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "qtos_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(qtos_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 19
    L = Q.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(Q.EXPORTS) == declared


@pytest.mark.parametrize("combo,duration", [("C1", 2.0), ("Custom", 5.0)])
def test_assembly_rows_dealt_to_warps_keep_every_sum(combo, duration):
    """k_asm's panel rows are dealt to its four warps by term count (units of two adjacent rows) instead of rows 4 w .. 4 w + 3:
    both deals hold the same terms, give every target its terms in the same order (same floating-point sums) and never put a
    target twice into one 32-lane step; the deal shortens the slowest warp's stream, and the packed steps hold little padding."""
    sh = Q.default_shape(combo, duration)
    even, dealt, auto = (Q.assembly_table_stats(sh, m) for m in (0, 1, -1))
    assert even["rows_dealt"] == 0 and dealt["rows_dealt"] == 1 and auto == dealt
    assert even["terms"] == dealt["terms"] == {"C1": 87894, "Custom": 217649}[combo]
    assert even["order_hash"] == dealt["order_hash"]
    assert even["duplicate_targets"] == dealt["duplicate_targets"] == 0
    assert dealt["slowest_warp_slots"] < 0.88 * even["slowest_warp_slots"]
    assert dealt["slots"] < 1.02 * even["slots"] and dealt["slots"] < 1.13 * dealt["terms"]


def test_struct_layouts_match_the_header():
    assert ctypes.sizeof(Q.Shape) == 8 * (1 + 9 + 12 + 3 + 3 + 1) + 8 + 8 + 8 + 8 + 8 + 8 + 8 + 16 + 8   # ints padded to 8; base_rom, dt_base_rom; two cost weights; terrain_gradients
    assert Q.PROBLEM_DTYPE.itemsize == 232 and Q.RESULT_DTYPE.itemsize == 56
    s = Q.default_shape()
    assert s.mass == 1.5 and s.combo == 5 and s.duration == 5.0 and s.force_polys_per_stance == 3
    assert abs(s.I_b[1] + 0.01938108) < 1e-15 and s.I_b[4] == 0.0          # F5 quirk tensor
    o = Q.default_options()
    assert o.tol == 1e-3 and o.max_iter == 200 and o.constr_viol_tol == 1e-4 and o.feas_exit == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(Q.QtosError, match="no CUDA device"):
        Q.Solver()


def test_cmd_args_matches_reference_outputs():
    gold = json.load(open(os.path.join(GOLDEN, "cmd_args.json")))
    assert towr_cli._flags == gold["flags"]
    for case in gold["cases"]:
        assert towr_cli.cmd_args(case["args"]) == case["out"]
    # numpy >= 2 scalars must not leak their repr into the command line
    out = towr_cli.cmd_args({"-s": [np.float64(0.25), np.float64(-1.0), 0.24], "-t": np.float64(1.5)})
    assert out == "-s 0.25 -1.0 0.24 -t 1.5 "


def test_parse_main_argv_semantics():
    a = towr_cli.parse_main_argv([])
    assert a["goal"] == [0.5, 0.0, 0.24] and a["start"] == [0.0, 0.0, 0.24] and a["combo"] == "Custom" and a["duration"] == 5.0
    assert a["ee"][0] == [0.21, 0.18, 0.0] and a["resolution"] == 0.1 and a["runtime"] == 15.0
    argv = "-s 0.335266 -0.0123145 0.221551 -g 0.91 0.0 0.24 -t 3.756 -resolution 0.01 -s_ang -0.04 -0.04 0.008 s_vel 1 2 3 -e1 0.5 0.2 0.0".split()
    a = towr_cli.parse_main_argv(argv)
    assert a["start"] == [0.335266, -0.0123145, 0.221551] and a["t_start"] == 3.756 and a["resolution"] == 0.01
    assert a["start_vel"] == [0.0, 0.0, 0.0]              # bare `s_vel` is ignored (F8)
    assert a["ee"][0] == [0.5, 0.2, 0.0] and a["ee"][1] == [0.21, -0.18, 0.0]
    a = towr_cli.parse_main_argv("-s_vel 1 2 3 -g 1 0 0.24".split())
    assert a["start_vel"] == [0.0, 0.0, 0.0]              # -n absent => zeroed (main.cpp:237-242)
    a = towr_cli.parse_main_argv("-s_vel 1 2 3 -n f".split())
    assert a["start_vel"] == [1.0, 2.0, 3.0]
    a = towr_cli.parse_main_argv("-duration 2.0 -s 1 1 0.24 -g 1.5 1 0.24 -n t".split())
    assert a["combo"] == "C0" and a["duration"] == 2.0 and a["start"][:2] == [0.0, 0.0] and a["goal"][:2] == [0.5, 0.0]
    assert towr_cli.exit_code(0) == 0 and towr_cli.exit_code(-1) == 255 and towr_cli.exit_code(-2) == 254


def test_heightfield_mirror_matches_reference_generator(golden_hf, tmp_path):
    g = golden_hf
    assert np.array_equal(HF.scale_map(g["tile_climb_1"], 3), g["scale_map_3"])
    for name in ("exp_1", "exp_3", "exp_5"):
        world, towr = g[name + "_world"], g[name + "_towr"]
        assert np.array_equal(HF.towr_grid(world), towr)
        assert HF.resolution(world) == float(g[name + "_res"])
    world5 = HF.combine_tiles([HF.scale_map(g["tile_climb_1"] * 0 + 0, 1)] * 2)
    assert world5.shape == (20, 40)
    # file format round trip: "v, v, ...," rows, no trailing newline
    path = str(tmp_path / "towr_heightfield.txt")
    HF.write_heightfield(path, g["exp_3_towr"])
    txt = open(path).read()
    assert txt.endswith(",") and txt.count("\n") == 59
    assert np.array_equal(HF.read_towr_heightfield(path), g["exp_3_towr"])
    with pytest.raises(ValueError):
        open(path, "w").write("")
        HF.read_towr_heightfield(path)


def test_host_get_height_is_bit_exact_with_oracle(oracle, golden_hf):
    rng = np.random.default_rng(3)
    for name in ("exp_3", "exp_5"):
        grid, res = golden_hf[name + "_towr"], float(golden_hf[name + "_res"])
        ter = oracle.Terrain(grid, res)
        pts = np.concatenate([rng.uniform(-1.3, 4.5, (400, 2)), [[-1.0, -1.0], [0.0, 0.0], [-1.5, 0.2], [9.0, 9.0], [0.5, -1.0 + res * 7]]])
        h = HF.get_height(grid, res, pts[:, 0], pts[:, 1])
        ix0, iy0, ix1, iy1 = HF.cell_indices(grid.shape, res, pts[:, 0], pts[:, 1])
        for i, (x, y) in enumerate(pts):
            assert h[i] == ter.height(x, y)
            assert (ix0[i], iy0[i], ix1[i], iy1[i]) == ter.cell(x, y)


def test_workload_generator_is_deterministic():
    grid, res = HF.rough_terrain(1234)
    assert grid.shape == (256, 256) and res == 0.02 and 0 <= grid.min() and grid.max() <= 0.075
    a = workloads.multistart_problems(64, grid, res, seed=1234, group_size=8)
    b = workloads.multistart_problems(64, grid, res, seed=1234, group_size=8)
    assert a.tobytes() == b.tobytes() and a["group"].max() == 7
    assert np.all(a["goal"][:, 0] - a["start_pos"][:, 0] >= 0.2) and np.all(a["goal"][:, 0] - a["start_pos"][:, 0] <= 0.6)
    assert np.allclose(a["start_pos"][:, 2] - 0.24, HF.get_height(grid, res, a["start_pos"][:, 0], a["start_pos"][:, 1]))


def test_csv_free_handoff_equals_the_text_detour(tmp_path):
    """arrays handed to the Combiner-side readers carry exactly the values the CSV text would."""
    import pandas as pd
    from qtos_b200 import handoff
    rng = np.random.default_rng(5)
    rows = rng.normal(size=(200, 37)) * 10.0 ** rng.integers(-7, 4, size=(200, 37))
    rows[3, 5] = 0.0; rows[4, 6] = -0.0; rows[5, 7] = 1e-5; rows[6, 8] = 123456.5; rows[7, 9] = 0.1 + 0.2
    path = str(tmp_path / "traj.csv")
    Q.write_csv(rows, path)                                         # C-ABI writer, "%g" like ofstream << double
    text = np.loadtxt(path, delimiter=",")
    assert np.array_equal(handoff.as_csv_values(rows), text)
    assert np.array_equal(handoff.read_csv_frame(rows), pd.read_csv(path).to_numpy())       # first row eaten as header
    import csv
    with open(path, newline="") as f:
        r7 = list(csv.reader(f))[7]
    st = handoff.state_of_row(handoff.as_csv_values(rows)[7])
    assert st["CoM"] == [float(v) for v in r7[1:4]] and st["HR_FOOT"] == [float(v) for v in r7[16:19]]
    assert st["CoM_vel_ang"] == [float(v) for v in r7[22:25]] and sorted(st) == sorted(
        ["CoM", "orientation", "FL_FOOT", "FR_FOOT", "HL_FOOT", "HR_FOOT", "CoM_vel", "CoM_vel_ang"])


def test_heightfield_file_edge_cases(tmp_path):
    """parser of towr_heightfield.txt (ref: custom_terrain.cpp:22-49): trailing commas and blank tail lines are fine,
    ragged or empty files fail loudly instead of the reference's undefined behaviour."""
    good = tmp_path / "ok.txt"; good.write_text("0.0, 0.5, 1.0,\n0.25, 0.0, 0.0,\n\n")
    g = HF.read_towr_heightfield(str(good))
    assert g.shape == (2, 3) and g[0, 1] == 0.5 and g[1, 0] == 0.25
    ragged = tmp_path / "ragged.txt"; ragged.write_text("0, 1, 2,\n0, 1,\n")
    with pytest.raises(ValueError, match="ragged"):
        HF.read_towr_heightfield(str(ragged))
    empty = tmp_path / "empty.txt"; empty.write_text("\n\n")
    with pytest.raises(ValueError, match="empty"):
        HF.read_towr_heightfield(str(empty))
    # write -> read round trip of the reference's file format, incl. the missing final newline
    m = np.round(np.random.default_rng(3).uniform(0, 0.2, (7, 5)), 4)
    HF.write_heightfield(str(tmp_path / "rt.txt"), m)
    assert not open(tmp_path / "rt.txt").read().endswith("\n")
    assert np.array_equal(HF.read_towr_heightfield(str(tmp_path / "rt.txt")), m)


def test_n_flag_compares_the_joined_tokens():
    """main.cpp joins the (up to three) tokens after -n and compares the result with "t" (main.cpp:66,228-232): `-n t`
    normalises only as the LAST thing on the command line; followed by another flag it reads "t -g 0.5" != "t"."""
    from qtos_b200 import towr_cli
    a = towr_cli.parse_main_argv(["-s", "1.0", "2.0", "0.24", "-g", "1.5", "2.0", "0.24", "-n", "t"])
    assert a["normalize"] and a["start"][:2] == [0.0, 0.0] and a["goal"][:2] == [0.5, 0.0]
    b = towr_cli.parse_main_argv(["-n", "t", "-s", "1.0", "2.0", "0.24", "-g", "1.5", "2.0", "0.24"])
    assert not b["normalize"] and b["start"][:2] == [1.0, 2.0]


def test_python_structs_follow_the_header_field_by_field():
    """the ctypes / numpy mirrors of the C ABI's structs name the header's fields in the header's order (sizes alone would not
    notice two swapped doubles), and so does the stub INTEGRATION.md shows a maintainer"""
    hdr = open(os.path.join(ROOT, "include", "qtos_b200.h")).read()

    def fields(name):
        body = re.search(r"typedef struct\s*\{([^{}]*)\}\s*%s;" % name, hdr).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            out.append(re.match(r"(?:const\s+)?\w+\s+\**(\w+)", decl).group(1))
        return out

    assert fields("qtos_shape") == [f[0] for f in Q.Shape._fields_]
    assert fields("qtos_options") == [f[0] for f in Q.Options._fields_]
    assert fields("qtos_problem") == list(Q.PROBLEM_DTYPE.names)
    assert fields("qtos_result") == list(Q.RESULT_DTYPE.names)
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    stub = re.search(r"class Shape\(C\.Structure\):.*?_fields_ = \[(.*?)\][^\n]*\nclass Problem", doc, re.S).group(1)
    assert re.findall(r'\("(\w+)"', stub) == fields("qtos_shape")
