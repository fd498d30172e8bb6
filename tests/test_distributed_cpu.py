"""
CPU, world_size 2 and 3 over gloo: the multi-GPU host logic of shot_fpfh_b200/distributed.py (block sharding,
padded all-gather, SPFH exchange, merge of per-shard nearest neighbours) with the NumPy oracle standing in for the
CUDA kernels — sharded results must equal the unsharded ones exactly.
"""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fpfh_oracle, matching_oracle, shot_oracle
from shot_fpfh_b200 import distributed as sfd
from shot_fpfh_b200 import synthetic


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, size: int, port: int, fn_name: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(size))
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        globals()[fn_name](rank, size)
    finally:
        dist.destroy_process_group()


def _spawn(fn_name: str, size: int):
    mp.spawn(_worker, args=(size, _free_port(), fn_name), nprocs=size, join=True)


def _cloud():
    pts, normals = synthetic.bumpy_sphere(1500, seed=41)
    return pts, normals, 6.0 * synthetic.mean_spacing(1500)


def _check_shot(rank, size):
    pts, normals, radius = _cloud()
    kp = pts[::7]  # 215 queries: not divisible by 2 or 3
    full = shot_oracle.shot_single_scale(pts, normals, kp, radius, True, 5)

    def block(lo, hi):
        return torch.from_numpy(shot_oracle.shot_single_scale(pts, normals, kp[lo:hi], radius, True, 5))

    got = sfd.sharded_rows(kp.shape[0], block, gather=True)
    assert np.array_equal(got.numpy(), full)
    local = sfd.sharded_rows(kp.shape[0], block, gather=False)
    lo, hi = sfd.block_bounds(kp.shape[0], size, rank)
    assert np.array_equal(local.numpy(), full[lo:hi])


def _check_fpfh(rank, size):
    from sklearn.neighbors import KDTree

    pts, normals, radius = _cloud()
    n = pts.shape[0]
    kp = np.arange(0, n, 3)
    full, spfh_full = fpfh_oracle.fpfh(kp, pts, normals, radius, 11, True, return_spfh=True)
    tree = KDTree(pts)
    edges = fpfh_oracle.bin_edges(11)

    def spfh_block(first, end):  # "cell-sorted order" = original order here
        rows = np.zeros((end - first, 33))
        nbh = tree.query_radius(pts[first:end], radius)
        for r, i in enumerate(range(first, end)):
            a, p, t = fpfh_oracle.pair_features(pts[i], normals[i], pts[nbh[r]], normals[nbh[r]])
            rows[r] = fpfh_oracle.spfh_row(a, p, t, nbh[r].shape[0], 11, True, edges)
        return torch.from_numpy(rows)

    def fpfh_block(spfh_all, lo, hi):
        s = spfh_all.numpy()
        assert np.array_equal(s, spfh_full)  # the exchange reproduced the whole table on every rank
        nbh, d = tree.query_radius(pts[kp[lo:hi]], radius, return_distance=True)
        out = np.zeros((hi - lo, 33))
        for r, i in enumerate(kp[lo:hi]):
            far = d[r] > 0
            out[r] = s[i] + (s[nbh[r][far]] / d[r][far][:, None]).sum(axis=0) / nbh[r].shape[0]
        return torch.from_numpy(out)

    got = sfd.sharded_fpfh(n, kp.shape[0], spfh_block, fpfh_block, gather=True)
    assert np.array_equal(got.numpy(), full)


def _check_fpfh_by_position(rank, size):
    """Both FPFH stages on the same blocks of the (permuted) "cell-sorted" order; keypoints scattered over the blocks."""
    from sklearn.neighbors import KDTree

    pts, normals, radius = _cloud()
    n = pts.shape[0]
    kp = np.random.default_rng(8).permutation(n)[:701]  # arbitrary order, not all points
    full, spfh_full = fpfh_oracle.fpfh(kp, pts, normals, radius, 11, True, return_spfh=True)
    perm = np.random.default_rng(9).permutation(n)  # sorted position -> original index
    inv_perm = np.empty(n, dtype=np.int64)
    inv_perm[perm] = np.arange(n)
    tree = KDTree(pts)
    edges = fpfh_oracle.bin_edges(11)

    def spfh_block(first, end):
        rows = np.zeros((end - first, 33))
        nbh = tree.query_radius(pts[perm[first:end]], radius)
        for r, i in enumerate(perm[first:end]):
            a, p, t = fpfh_oracle.pair_features(pts[i], normals[i], pts[nbh[r]], normals[nbh[r]])
            rows[r] = fpfh_oracle.spfh_row(a, p, t, nbh[r].shape[0], 11, True, edges)
        return torch.from_numpy(rows)

    def fpfh_rows(spfh_all, mine):
        s = spfh_all.numpy()  # rows in sorted order
        assert np.array_equal(s, spfh_full[perm])
        first, end = sfd.block_bounds(n, size, rank)
        ids = kp[mine.numpy()]
        assert ((inv_perm[ids] >= first) & (inv_perm[ids] < end)).all()
        nbh, d = tree.query_radius(pts[ids], radius, return_distance=True)
        out = np.zeros((ids.shape[0], 33))
        for r, i in enumerate(ids):
            far = d[r] > 0
            out[r] = s[inv_perm[i]] + (s[inv_perm[nbh[r][far]]] / d[r][far][:, None]).sum(axis=0) / nbh[r].shape[0]
        return torch.from_numpy(out)

    positions = torch.from_numpy(inv_perm[kp])
    got = sfd.sharded_fpfh_by_position(n, positions, 33, spfh_block, fpfh_rows, gather=True)
    assert np.array_equal(got.numpy(), full)
    mine, local = sfd.sharded_fpfh_by_position(n, positions, 33, spfh_block, fpfh_rows, gather=False)
    assert np.array_equal(local.numpy(), full[mine.numpy()])


def _check_upload_replicated(rank, size):
    """An N-th of the rows per rank + one all-gather == the whole array on every rank (equal and ragged blocks)."""
    rng = np.random.default_rng(17)
    for n in (12, 13, 1, 601):
        rows = rng.standard_normal((n, 5))
        seen = []

        def to_cpu(a, dtype):
            seen.append(a.shape[0])
            return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)

        got = sfd.upload_replicated(rows, torch.float64, to_device=to_cpu)
        assert np.array_equal(got.numpy(), rows)
        lo, hi = sfd.block_bounds(n, size, rank)
        assert seen == ([hi - lo] if n >= size else [n])  # only this rank's block went "through PCIe"


def _check_matching(rank, size):
    rng = np.random.default_rng(3)
    a = synthetic.sparse_unit_rows(90, 64, seed=1).astype(np.float64)
    b = synthetic.sparse_unit_rows(131, 64, seed=2).astype(np.float64)
    b[40] = b[7]  # duplicated target rows in different shards: the lowest index must win
    b[100] = b[7]
    a[3] = b[7]
    a[::9] = 0.0
    sa, sb, nn, d1, dmat = matching_oracle.nearest(a, b)
    d2 = np.partition(dmat, 1, axis=1)[:, 1]

    def shard(lo, hi):
        sub = dmat[:, lo:hi]
        if sub.shape[1] == 0:
            inf = torch.full((sa.shape[0],), float("inf"), dtype=torch.float64)
            return torch.full((sa.shape[0],), -1, dtype=torch.int64), inf, inf.clone()
        local_nn = sub.argmin(axis=1)
        local_d1 = sub[np.arange(sub.shape[0]), local_nn]
        local_d2 = np.partition(sub, 1, axis=1)[:, 1] if sub.shape[1] > 1 else np.full(sub.shape[0], np.inf)
        return torch.from_numpy(local_nn + lo), torch.from_numpy(local_d1), torch.from_numpy(local_d2)

    got_nn, got_d1, got_d2 = sfd.sharded_nearest(sb.shape[0], shard)
    assert np.array_equal(got_nn.numpy(), nn)
    assert np.array_equal(got_d1.numpy(), d1) and np.array_equal(got_d2.numpy(), d2)
    del rng


def _check_root_and_workers(rank, size):
    """One process (rank 0) runs the caller's code, the others serve: the request, the two broadcasts of the operands
    and the merged answer — with a brute-force search standing in for the CUDA one."""
    a = torch.from_numpy(synthetic.sparse_unit_rows(70, 32, seed=11).astype(np.float64))
    b = torch.from_numpy(synthetic.sparse_unit_rows(101, 32, seed=12).astype(np.float64))
    a[::8] = 0.0
    b[50] = b[3]

    def brute(a_rows, ref_block, n_ref, k, group, mark):
        rows_a = torch.nonzero(a_rows.abs().sum(dim=1) > 0).squeeze(1)

        def shard(lo, hi):
            blk = ref_block(lo, hi)
            live = torch.nonzero(blk.abs().sum(dim=1) > 0).squeeze(1)
            if live.shape[0] == 0:
                inf = torch.full((rows_a.shape[0],), float("inf"), dtype=torch.float64)
                return torch.full((rows_a.shape[0],), -1, dtype=torch.int64), inf, inf.clone()
            d = torch.cdist(a_rows[rows_a], blk[live], compute_mode="donot_use_mm_for_euclid_dist")
            order = torch.sort(d, dim=1, stable=True)
            second = order.values[:, 1] if live.shape[0] > 1 else torch.full_like(order.values[:, 0], float("inf"))
            return live[order.indices[:, 0]] + lo, order.values[:, 0], second

        return (rows_a,) + tuple(sfd.sharded_nearest(n_ref, shard, group))

    sfd._HANDLERS["nearest_core"] = brute
    try:
        if rank == 0:
            sfd.start_root_service(warm=False)
            assert sfd.root_service_active() == (size > 1)
            for x, y in ((a, b), (b, a)):
                rows, nn, d1, d2 = sfd.nearest_neighbors_from_root(x, y, 8) if size > 1 else brute(
                    x, lambda lo, hi: y[lo:hi], y.shape[0], 8, None, None)
                live_y = torch.nonzero(y.abs().sum(dim=1) > 0).squeeze(1)
                d = torch.cdist(x[rows], y[live_y], compute_mode="donot_use_mm_for_euclid_dist")
                assert torch.equal(rows, torch.nonzero(x.abs().sum(dim=1) > 0).squeeze(1))
                assert torch.equal(nn, live_y[d.argmin(dim=1)]) and torch.equal(d1, d.min(dim=1).values)
                assert torch.equal(d2, torch.sort(d, dim=1).values[:, 1])
            sfd.stop_root_service()
            assert not sfd.root_service_active()
        else:
            assert sfd.serve() == 2
    finally:
        sfd._HANDLERS.pop("nearest_core", None)


def _check_slabs(rank, size):
    """Halo partition: every keypoint belongs to exactly one slab, and the points a rank sorts (slab + halo) contain
    EVERY neighbour of its keypoints — checked with rows that are exact functions of the neighbour set (count, sum and
    maximum of the neighbours' original indices, the SHOT oracle's rows on top), so a missing neighbour cannot hide."""
    from sklearn.neighbors import KDTree

    pts, normals, radius = _cloud()
    pts = pts * np.array([1.0, 1.7, 0.6])  # the longest axis is not the first one
    kp_idx = np.arange(0, pts.shape[0], 5)
    kp = np.ascontiguousarray(pts[kp_idx])
    kp[-3:] += 10.0 * radius  # keypoints away from the cloud: no neighbours, zero rows
    tree = KDTree(pts)

    def rows_from(points_idx, keypoints):
        sub = pts[points_idx]
        nbh = KDTree(sub).query_radius(keypoints, radius)
        rows = np.zeros((keypoints.shape[0], 3 + 352))
        for r, nb in enumerate(nbh):
            orig = np.sort(points_idx[nb])
            rows[r, :3] = (orig.shape[0], orig.sum(), orig.max() if orig.shape[0] else -1)
        rows[:, 3:] = shot_oracle.shot_single_scale(sub, normals[points_idx], keypoints, radius, True, 5)
        return rows

    full = rows_from(np.arange(pts.shape[0]), kp)
    assert len(tree.query_radius(kp[-1:], radius)[0]) == 0
    seen = {}

    def geometry(lo, hi, r):
        return {"cell": r * 1.001}

    def rows_of(point_idx, keypoint_idx, box):
        seen["points"] = int(point_idx.shape[0])
        return torch.from_numpy(rows_from(point_idx.numpy(), kp[keypoint_idx.numpy()]))

    got = sfd.sharded_rows_by_slab(torch.from_numpy(pts), torch.from_numpy(kp), radius, geometry, rows_of, 355,
                                   gather=True, out_dtype=torch.float64)
    assert np.array_equal(got.numpy()[:, :3], full[:, :3])  # the same neighbour sets
    assert np.allclose(got.numpy()[:, 3:], full[:, 3:], rtol=0, atol=1e-12)  # (KDTree order differs: summation order)
    mine, local = sfd.sharded_rows_by_slab(torch.from_numpy(pts), torch.from_numpy(kp), radius, geometry, rows_of, 355,
                                           gather=False, out_dtype=torch.float64)
    counts = torch.zeros(kp.shape[0], dtype=torch.int64)
    counts[mine] = 1
    dist.all_reduce(counts)
    assert bool((counts == 1).all())  # a partition of the keypoints
    if size > 1 and "points" in seen:
        assert seen["points"] < pts.shape[0]  # a rank does not sort the whole cloud


@pytest.mark.parametrize("size", [2, 3])
@pytest.mark.parametrize("fn", ["_check_shot", "_check_fpfh", "_check_fpfh_by_position", "_check_upload_replicated",
                                "_check_matching", "_check_slabs", "_check_root_and_workers"])
def test_sharded_equals_unsharded(fn, size):
    _spawn(fn, size)


def test_slab_bounds_single_process():
    layers = torch.tensor([5, 5, 5, 5, 2, 9, 9, 3, 3, 3, 7, 1])
    for parts in (1, 2, 3, 5, 16):
        b = sfd.slab_bounds(layers, parts)
        assert len(b) == parts + 1 and b[0] == 1 and b[-1] == 10 and all(x <= y for x, y in zip(b[:-1], b[1:]))
        owned = sum(int(((layers >= b[r]) & (layers < b[r + 1])).sum()) for r in range(parts))
        assert owned == layers.shape[0]
    assert sfd.slab_bounds(torch.empty(0, dtype=torch.long), 3) == [0, 0, 0, 0]


def test_block_bounds_and_merge_single_process():
    for n in (0, 1, 7, 100):
        for parts in (1, 2, 3, 8):
            b = [sfd.block_bounds(n, parts, p) for p in range(parts)]
            assert b[0][0] == 0 and b[-1][1] == n and all(x[1] == y[0] for x, y in zip(b[:-1], b[1:]))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1
    d1 = torch.tensor([[0.5, 0.2, 0.3], [0.5, 0.1, 0.4]], dtype=torch.float64)
    nn = torch.tensor([[1, 2, 3], [11, 12, 13]])
    d2 = torch.tensor([[0.6, 0.25, 0.9], [0.7, 0.15, 0.45]], dtype=torch.float64)
    m_nn, m_d1, m_d2 = sfd.merge_nearest(d1, nn, d2)
    assert m_nn.tolist() == [1, 12, 3] and m_d1.tolist() == [0.5, 0.1, 0.3] and m_d2.tolist() == [0.5, 0.15, 0.4]
