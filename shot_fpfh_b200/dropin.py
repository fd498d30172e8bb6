"""
Drop-in installation: rebinds the hot-path names of an importable reference `shot_fpfh` package to the B200
implementations, so that its `RegistrationPipeline` (pipeline.py) and `scripts/register_point_clouds.py` run
UNCHANGED on top of the CUDA kernels.

    import shot_fpfh_b200.dropin as dropin
    dropin.install()                      # before or after `import shot_fpfh`
    from shot_fpfh import RegistrationPipeline   # the reference's own orchestration, untouched

What is replaced (SURVEY.md §8b — exactly what pipeline.py imports at :15 and :24-30, plus the deeper helpers):
    shot_fpfh.descriptors.ShotMultiprocessor / compute_fpfh_descriptor   (and in shot_fpfh.pipeline's namespace)
    shot_fpfh.descriptors.shot.{get_local_rf, compute_single_shot_descriptor, compute_shot_descriptor}
    shot_fpfh.matching.{basic_matching, match_descriptors, double_matching_with_rejects}
    shot_fpfh.descriptors.compute_normals (also re-exported by `shot_fpfh`), shot_fpfh.core.grid_subsampling,
    shot_fpfh.keypoint_selection.{select_keypoints_subsampling, select_keypoints_with_density_threshold}
        — the two "next" rows immediately upstream of the hot path
    shot_fpfh.matching.ransac_on_matches, shot_fpfh.icp.icp_point_to_plane   (they return this
        package's RigidTransform, which has the reference class's interface)
    shot_fpfh.icp.icp_point_to_point — the reference's raises a TypeError on every input with more than one pair (icp.py:118-120,
        SURVEY.md D-8); the replacement is what it intends (shot_fpfh_b200/icp.py), so `run_icp("point_to_point")` works
Everything else (iterative / random keypoint selection, the sampling ICP, I/O, configuration, analysis) stays the
reference's code.
(`scripts/register_point_clouds.py` does `from shot_fpfh import compute_normals` at import time: call install()
before importing the script for the GPU normals to be picked up there.)
"""

from __future__ import annotations

import importlib
import sys

_REPLACED: dict[tuple[str, str], object] = {}


def _bind(module_name: str, attr: str, value) -> None:
    module = importlib.import_module(module_name)
    key = (module_name, attr)
    if key not in _REPLACED:
        _REPLACED[key] = getattr(module, attr, None)
    setattr(module, attr, value)


def install() -> list[str]:
    """Rebinds the names; returns the list of `module.attr` that were replaced. Needs the built CUDA library."""
    from . import descriptors as d
    from . import icp
    from . import keypoint_selection as k
    from . import matching as m
    from .descriptors import shot as s
    from .subsampling import grid_subsampling_gpu as _grid_subsampling

    importlib.import_module("shot_fpfh")  # raises ImportError if the reference package is not importable
    plan = [
        ("shot_fpfh.descriptors.shot_parallelization", "ShotMultiprocessor", d.ShotMultiprocessor),
        ("shot_fpfh.descriptors", "ShotMultiprocessor", d.ShotMultiprocessor),
        ("shot_fpfh.descriptors.fpfh", "compute_fpfh_descriptor", d.compute_fpfh_descriptor),
        ("shot_fpfh.descriptors", "compute_fpfh_descriptor", d.compute_fpfh_descriptor),
        ("shot_fpfh.descriptors.pca_based_descriptors", "compute_normals", d.compute_normals),
        ("shot_fpfh.descriptors", "compute_normals", d.compute_normals),
        ("shot_fpfh", "compute_normals", d.compute_normals),
        ("shot_fpfh.core.subsampling", "grid_subsampling", _grid_subsampling),
        ("shot_fpfh.core", "grid_subsampling", _grid_subsampling),
        ("shot_fpfh.keypoint_selection", "grid_subsampling", _grid_subsampling),
        ("shot_fpfh.keypoint_selection", "select_keypoints_subsampling", k.select_keypoints_subsampling),
        ("shot_fpfh.keypoint_selection", "select_keypoints_with_density_threshold", k.select_keypoints_with_density_threshold),
        ("shot_fpfh.pipeline", "select_keypoints_subsampling", k.select_keypoints_subsampling),
        ("shot_fpfh.pipeline", "select_keypoints_with_density_threshold", k.select_keypoints_with_density_threshold),
        ("shot_fpfh.descriptors.shot", "get_local_rf", s.get_local_rf),
        ("shot_fpfh.descriptors.shot", "compute_single_shot_descriptor", s.compute_single_shot_descriptor),
        ("shot_fpfh.descriptors.shot", "compute_shot_descriptor", s.compute_shot_descriptor),
        ("shot_fpfh.matching.matching", "basic_matching", m.basic_matching),
        ("shot_fpfh.matching.matching", "match_descriptors", m.match_descriptors),
        ("shot_fpfh.matching.matching", "double_matching_with_rejects", m.double_matching_with_rejects),
        ("shot_fpfh.matching", "basic_matching", m.basic_matching),
        ("shot_fpfh.matching", "match_descriptors", m.match_descriptors),
        ("shot_fpfh.matching", "double_matching_with_rejects", m.double_matching_with_rejects),
        ("shot_fpfh.matching.ransac", "ransac_on_matches", m.ransac_on_matches),
        ("shot_fpfh.matching", "ransac_on_matches", m.ransac_on_matches),
        ("shot_fpfh.pipeline", "ransac_on_matches", m.ransac_on_matches),
        ("shot_fpfh.icp", "icp_point_to_plane", icp.icp_point_to_plane),
        ("shot_fpfh.pipeline", "icp_point_to_plane", icp.icp_point_to_plane),
        ("shot_fpfh.icp", "icp_point_to_point", icp.icp_point_to_point),
        ("shot_fpfh.pipeline", "icp_point_to_point", icp.icp_point_to_point),
        # pipeline.py did `from shot_fpfh.descriptors import ...` / `from shot_fpfh.matching import ...`
        ("shot_fpfh.pipeline", "ShotMultiprocessor", d.ShotMultiprocessor),
        ("shot_fpfh.pipeline", "compute_fpfh_descriptor", d.compute_fpfh_descriptor),
        ("shot_fpfh.pipeline", "basic_matching", m.basic_matching),
        ("shot_fpfh.pipeline", "match_descriptors", m.match_descriptors),
        ("shot_fpfh.pipeline", "double_matching_with_rejects", m.double_matching_with_rejects),
    ]
    done = []
    for module_name, attr, value in plan:
        _bind(module_name, attr, value)
        done.append(f"{module_name}.{attr}")
    return done


def uninstall() -> None:
    """Restores the reference's own functions."""
    for (module_name, attr), value in _REPLACED.items():
        if module_name in sys.modules and value is not None:
            setattr(sys.modules[module_name], attr, value)
    _REPLACED.clear()
