"""
TEST INFRASTRUCTURE ONLY. NumPy restatement of the reference's PCA normals
(shot_fpfh/descriptors/pca_based_descriptors.py:15-26 `pca`, :29-59 `compute_normals`).

Neighbourhoods: `KDTree(cloud).query(queries, k)` (k nearest, the query itself included when it belongs to the cloud)
or `query_radius(queries, radius)`; normal = eigenvector of the smallest eigenvalue of the covariance of the
neighbourhood about its barycentre, as `np.linalg.eigh` returns it (LAPACK's sign), flipped when it points against
`pre_computed_normals[i]`.
"""

from __future__ import annotations

import numpy as np
from sklearn.neighbors import KDTree


def pca_normal(points):
    centred = points - points.mean(axis=0)
    cov = centred.T @ centred / points.shape[0]
    return np.linalg.eigh(cov)[1][:, 0]


def compute_normals(query_points, cloud_points, k=None, radius=None, pre_computed_normals=None):
    tree = KDTree(cloud_points)
    nbh = tree.query(query_points, k=k, return_distance=False) if k is not None else tree.query_radius(query_points, radius)
    out = np.zeros((query_points.shape[0], 3))
    for i in range(query_points.shape[0]):
        out[i] = pca_normal(cloud_points[nbh[i]])
        if pre_computed_normals is not None and out[i].dot(pre_computed_normals[i]) < 0:
            out[i] *= -1
    return out


def knn_sets(query_points, cloud_points, k):
    return KDTree(cloud_points).query(query_points, k=k, return_distance=False)
