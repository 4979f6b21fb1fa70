import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """a plain `pytest tests` on a box without CUDA skips the GPU tier instead of failing in it"""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device (the solver has no CPU fallback)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_hf():
    return np.load(os.path.join(GOLDEN, "heightfields.npz"))


@pytest.fixture(scope="session")
def golden_csv():
    return np.load(os.path.join(GOLDEN, "gait_csv.npz"))


@pytest.fixture(scope="session")
def towr_log():
    return json.load(open(os.path.join(GOLDEN, "towr_log.json")))


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.lib()
    return O


FEET_19 = [(0.21, 0.19, 0.0), (0.21, -0.19, 0.0), (-0.21, 0.19, 0.0), (-0.21, -0.19, 0.0)]


def oracle_problem(O, shape_o, pr, grid, res):
    """oracle problem for one qtos_problem record (numpy structured scalar)."""
    inst = O.make_instance(start_pos=pr["start_pos"], start_ang=pr["start_ang"], goal=pr["goal"], ee=pr["ee"],
                           t_start=float(pr["t_start"]), start_vel=pr["start_vel"], start_ang_vel=pr["start_ang_vel"])
    return O.Problem(shape_o, inst, O.Terrain(grid, res))
