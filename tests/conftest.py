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


def registration_case(n=4000, seed=31):
    """
    A scan, its rigidly moved, permuted and slightly noisy copy with normals, and 600 matches, a third wrong.
    The scan is three mutually orthogonal unit faces (a box corner) with a ripple on each: point-to-plane ICP is
    well conditioned on it (on the near-spherical benchmark cloud it slides around the centre and does not
    converge — in the reference just the same).
    """
    from shot_fpfh_b200 import synthetic

    rng = np.random.default_rng(seed)
    face = rng.integers(0, 3, n)
    uv = rng.uniform(0.0, 1.0, size=(n, 2))
    ripple = 0.03 * np.sin(7.0 * uv[:, 0]) * np.cos(5.0 * uv[:, 1])
    scan = np.zeros((n, 3))
    normals = np.zeros((n, 3))
    for k in range(3):
        on = face == k
        a, b = (k + 1) % 3, (k + 2) % 3
        scan[on, a], scan[on, b], scan[on, k] = uv[on, 0], uv[on, 1], ripple[on]
        # normal of the rippled face: (-dh/da, -dh/db, 1) normalised
        da = 0.21 * np.cos(7.0 * uv[on, 0]) * np.cos(5.0 * uv[on, 1])
        db = -0.15 * np.sin(7.0 * uv[on, 0]) * np.sin(5.0 * uv[on, 1])
        nrm = np.stack([-da, -db, np.ones(on.sum())], axis=1)
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        normals[on, a], normals[on, b], normals[on, k] = nrm[:, 0], nrm[:, 1], nrm[:, 2]
    ref, ref_normals, perm, _, _ = synthetic.rigid_pair(scan, normals)
    ref = ref + 0.002 * rng.normal(size=ref.shape)
    true_ref = np.empty(n, dtype=np.int64)
    true_ref[perm] = np.arange(n)  # ref[k] is scan point perm[k]
    scan_idx = rng.choice(n, 600, replace=False)
    ref_idx = true_ref[scan_idx].copy()
    wrong = rng.random(600) < 0.33
    ref_idx[wrong] = rng.integers(0, n, wrong.sum())
    return scan, ref, ref_normals, scan_idx, ref_idx
