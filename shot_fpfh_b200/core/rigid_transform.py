"""
`RigidTransform` with the reference's interface (shot_fpfh/core/rigid_transform.py:10-106): a rotation matrix and a
translation with composition (`@`), application to points (`t[points]`, `t.transform(points)`), inversion (`~t`,
`t.inv()`) and quaternion re-normalisation. Host NumPy/SciPy: a 3x3 object, nothing to accelerate — it is here
because the device RANSAC / ICP return it.
"""

from __future__ import annotations

import numpy as np
import numpy.typing as npt
from scipy.spatial.transform import Rotation


class RigidTransform:
    def __init__(self, rotation: npt.NDArray[np.float64] = np.eye(3), translation: npt.NDArray[np.float64] = np.zeros(3)):
        self.rotation = rotation
        self.translation = translation

    def __repr__(self) -> str:
        """The 4x4 matrix without scientific notation or brackets (pastes into CloudCompare), as the reference prints it."""
        top = np.hstack((self.rotation, self.translation[:, None]))
        with np.printoptions(suppress=True):
            return str(np.vstack((top, np.array([0, 0, 0, 1])))).replace("[", "").replace("]", "")

    def normalize_rotation(self) -> None:
        """Through a unit quaternion and back (rigid_transform.py:45-52)."""
        quat = Rotation.from_matrix(self.rotation).as_quat()
        self.rotation = Rotation.from_quat(quat / np.linalg.norm(quat)).as_matrix()

    def __matmul__(self, other_transformation: "RigidTransform") -> "RigidTransform":
        """self after the other one; the product's rotation is re-normalised (rigid_transform.py:54-70)."""
        out = RigidTransform(
            self.rotation @ other_transformation.rotation,
            self.rotation @ other_transformation.translation + self.translation,
        )
        out.normalize_rotation()
        return out

    def __invert__(self) -> "RigidTransform":
        # as in the reference (rigid_transform.py:72-79): transposed rotation, negated translation
        return RigidTransform(self.rotation.T, -self.translation)

    def __getitem__(self, points: npt.NDArray[np.float64]) -> npt.NDArray[np.float64]:
        return points.dot(self.rotation.T) + self.translation

    def transform(self, points: npt.NDArray[np.float64]) -> npt.NDArray[np.float64]:
        return self[points]

    def inv(self) -> "RigidTransform":
        return ~self

    def as_row(self) -> npt.NDArray[np.float64]:
        """(12,) = rotation row-major then translation: the layout the C ABI takes."""
        return np.concatenate([np.asarray(self.rotation, dtype=np.float64).ravel(),
                               np.asarray(self.translation, dtype=np.float64).ravel()])
