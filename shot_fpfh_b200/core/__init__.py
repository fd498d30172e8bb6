"""Host-side helpers with the reference's names (shot_fpfh/core/__init__.py): the SE(3) wrapper and the two solvers."""
from ..subsampling import grid_subsampling_gpu as grid_subsampling
from .rigid_transform import RigidTransform
from .solvers import solver_point_to_plane, solver_point_to_point

__all__ = ["RigidTransform", "solver_point_to_point", "solver_point_to_plane", "grid_subsampling"]
