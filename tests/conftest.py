import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Every test session starts from a built tree (no-op when the artefacts are newer than the sources)."""
    import __graft_entry__

    __graft_entry__.build()


def rel_l2(a, b):
    """Per-row relative L2 error of a against the reference b."""
    return np.linalg.norm(np.asarray(a) - np.asarray(b), axis=-1) / np.maximum(np.linalg.norm(b, axis=-1), 1e-300)


def load_golden(name: str):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_pair_inputs(g):
    """Rebuilds the float64 inputs of a golden pair from its seeds + the stored float32 normals."""
    from shot_fpfh_b200 import synthetic

    scan, _ = synthetic.bumpy_sphere(int(g["n_points"]), int(g["seed"]))
    normals = g["scan_normals_f32"].astype(np.float64)
    ref, ref_normals, perm, rot, trans = synthetic.rigid_pair(scan, normals)
    return {"scan": (scan, normals), "ref": (ref, ref_normals)}, float(g["radius"])


def edge_case_inputs(g):
    from shot_fpfh_b200 import synthetic

    n = int(g["edge_n"])
    pts, dirs = synthetic.bumpy_sphere(n, seed=int(g["edge_seed"]))
    pts = np.concatenate([pts, pts[:40]])
    nrm = np.concatenate([dirs, dirs[:40]])
    return pts, nrm, g["edge_queries"], float(g["edge_radius"])
