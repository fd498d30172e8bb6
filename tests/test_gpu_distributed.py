"""
GPU: the sharded drivers of shot_fpfh_b200/distributed.py.
  * "virtual ranks" on one device: every block a rank would compute is computed in turn and the concatenation is
    compared BIT-EXACTLY with the unsharded result (no collective involved — that part is covered under gloo in
    tests/test_distributed_cpu.py);
  * when the box has >= 2 GPUs: a real 2-rank NCCL run of the three drivers.
"""

import os
import socket

import numpy as np
import pytest
import torch

from shot_fpfh_b200 import synthetic

pytestmark = pytest.mark.gpu


def _inputs():
    n = 40000
    pts, normals = synthetic.bumpy_sphere(n, seed=51)
    return pts, normals, 5.0 * synthetic.mean_spacing(n)


def test_virtual_ranks_shot_and_fpfh_bit_exact():
    from shot_fpfh_b200 import distributed as sfd
    from shot_fpfh_b200 import ops
    from shot_fpfh_b200.device import Grid, upload

    pts, normals, radius = _inputs()
    p, nr = upload(pts), upload(normals)
    grid = Grid().build(p, nr, radius)
    kp_idx = torch.arange(0, pts.shape[0], 3, device=p.device)
    kp = p[kp_idx].contiguous()

    # SHOT: unsharded vs 3 virtual ranks
    def shot_block(lo, hi):
        q = kp[lo:hi].contiguous()
        offsets, nbr, _, _ = ops.radius_csr(grid, q, radius)
        lrf = ops.shot_lrf(grid, q, radius, offsets, nbr)
        return ops.shot_descriptor(grid, q, radius, offsets, nbr, lrf, 10, True, out_dtype=torch.float32)

    full = shot_block(0, kp.shape[0])
    parts = torch.cat([shot_block(*sfd.block_bounds(kp.shape[0], 3, r)) for r in range(3)])
    assert torch.equal(full, parts)

    # FPFH: whole-cloud self-CSR path vs (SPFH by cell-sorted blocks) + (FPFH by keypoint blocks, CSR by keypoint)
    offsets, nbr, _, dist = ops.radius_csr(grid, None, radius, want_dist=True)
    spfh_full = ops.spfh(grid, offsets, nbr, 11, True)
    want = ops.fpfh(grid, offsets, nbr, dist, spfh_full, kp_idx, out_dtype=torch.float32)
    blocks = []
    for r in range(4):
        first, end = sfd.block_bounds(grid.n, 4, r)
        o, nb, _, _ = ops.radius_csr(grid, None, radius, self_range=(first, end - first))
        blocks.append(ops.spfh(grid, o, nb, 11, True, self_range=(first, end - first)))
    spfh_cat = torch.cat(blocks)
    assert torch.equal(spfh_cat, spfh_full)
    rows = []
    for r in range(4):
        lo, hi = sfd.block_bounds(kp_idx.shape[0], 4, r)
        mine = kp_idx[lo:hi].contiguous()
        o, nb, _, d = ops.radius_csr(grid, p[mine].contiguous(), radius, want_dist=True)
        rows.append(ops.fpfh(grid, o, nb, d, spfh_cat, mine, out_dtype=torch.float32, csr_by_keypoint=True))
    assert torch.equal(torch.cat(rows), want)

    # FPFH, the fused driver by blocks (what distributed.fpfh runs): both stages on the same 4 blocks of the
    # cell-sorted cloud == the fused driver on the whole cloud, bit for bit; keypoints in arbitrary order
    kp_any = kp_idx[torch.randperm(kp_idx.shape[0], device=kp_idx.device)].contiguous()
    want_fused, _ = ops.fpfh_cloud(grid, radius, 11, True, kp_any, out_dtype=torch.float32)
    _, inv_perm = ops.grid_permutation(grid)
    positions = inv_perm[kp_any].long()
    blocks4 = [ops.FpfhBlock(grid, radius, 11, True, *(lambda b: (b[0], b[1] - b[0]))(sfd.block_bounds(grid.n, 4, r)), p.device)
               for r in range(4)]
    spfh_all = torch.cat([b.spfh() for b in blocks4]).contiguous()
    got_fused = torch.zeros_like(want_fused)
    for r, b in enumerate(blocks4):
        first, end = sfd.block_bounds(grid.n, 4, r)
        mine = torch.nonzero((positions >= first) & (positions < end)).squeeze(1)
        got_fused[mine] = b.rows(spfh_all, kp_any[mine].contiguous())
    assert torch.equal(got_fused, want_fused)
    assert torch.allclose(got_fused, ops.fpfh(grid, offsets, nbr, dist, spfh_full, kp_any, out_dtype=torch.float32),
                          rtol=2e-5, atol=1e-7)
    grid.close()


def test_virtual_ranks_halo_slabs_bit_exact():
    """The halo partition (distributed.shot_single_scale_slabs) rank by rank on one device: a grid over a slab + halo
    of the cloud, built in the WHOLE cloud's box (sf_grid_build_in_box), gives the rows of the slab's keypoints bit for
    bit as the grid of the whole cloud does, and sorts a fraction of the points."""
    from shot_fpfh_b200 import distributed as sfd
    from shot_fpfh_b200 import ops
    from shot_fpfh_b200.device import Grid, grid_geometry, upload

    pts, normals, radius = _inputs()
    p, nr = upload(pts), upload(normals)
    kp = p[torch.randperm(p.shape[0], device=p.device)[:9000]].contiguous()
    kp[-5:] += 7.0 * radius  # off the cloud: zero rows
    whole = Grid().build(p, nr, radius)
    want = ops.shot_single_scale(whole, kp, radius, 10, True, out_dtype=torch.float32)[0]
    info = whole.info()
    lo, hi = p.min(dim=0).values, p.max(dim=0).values
    box = (tuple(lo.tolist()), tuple(hi.tolist()))
    geo = grid_geometry(box[0], box[1], radius)
    assert geo["cell"] == info["cell"] and tuple(geo["dims"]) == tuple(info["dims"])
    axis = int(torch.argmax(hi - lo).item())
    k_layers = sfd.cell_layer(kp[:, axis], box[0][axis], geo["cell"])
    p_layers = sfd.cell_layer(p[:, axis], box[0][axis], geo["cell"])
    for parts in (2, 5):
        bounds = sfd.slab_bounds(k_layers, parts)
        got = torch.zeros_like(want)
        covered = torch.zeros(kp.shape[0], dtype=torch.int64, device=p.device)
        sorted_points = 0
        grid = Grid()
        for r in range(parts):
            mine, halo = sfd.slab_members(k_layers, p_layers, bounds, r)
            covered[mine] += 1
            if mine.shape[0] == 0 or halo.shape[0] == 0:
                continue
            sorted_points += int(halo.shape[0])
            grid.build(p[halo].contiguous(), nr[halo].contiguous(), radius, box=box)
            got[mine] = ops.shot_single_scale(grid, kp[mine].contiguous(), radius, 10, True, out_dtype=torch.float32)[0]
            torch.cuda.synchronize()
            assert grid.poll() == 0
        assert bool((covered == 1).all())
        assert torch.equal(got, want)
        assert sorted_points < (1.0 + 0.35 * parts) * p.shape[0]  # slabs + halos, not `parts` whole clouds
        grid.close()
    # a point outside the given box is reported
    g2 = Grid().build(p[:1000].contiguous(), nr[:1000].contiguous(), radius,
                      box=((0.0, 0.0, 0.0), (0.01, 0.01, 0.01)))
    torch.cuda.synchronize()
    assert g2.poll() == 1
    g2.close()
    whole.close()


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, size, port, out_dir):
    import torch.distributed as dist

    from oracle import matching_oracle
    from shot_fpfh_b200 import distributed as sfd
    from shot_fpfh_b200.descriptors import ShotMultiprocessor, compute_fpfh_descriptor

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(size))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=size, device_id=torch.device("cuda", rank))
    try:
        pts, normals, radius = _inputs()
        kp_idx = np.arange(0, pts.shape[0], 5)
        got = sfd.shot_single_scale(pts, normals, pts[kp_idx], radius, True, 10, gather=True).cpu().numpy()
        with ShotMultiprocessor(min_neighborhood_size=10, verbose=False) as shot:
            want = shot.compute_descriptor_single_scale(pts, normals, pts[kp_idx], radius)
        assert np.array_equal(got.astype(np.float64), want)
        got_h = sfd.shot_single_scale(pts, normals, pts[kp_idx], radius, True, 10, gather=True, partition="slabs")
        assert np.array_equal(got_h.cpu().numpy().astype(np.float64), want)  # halo partition: the same rows
        got_f = sfd.fpfh(kp_idx, pts, normals, radius, 11, True, gather=True, out_dtype=torch.float64).cpu().numpy()
        want_f = compute_fpfh_descriptor(kp_idx, pts, normals, radius, 11, True, verbose=False)
        assert np.array_equal(got_f, want_f)
        a = synthetic.sparse_unit_rows(3000, 352, seed=5).astype(np.float64)
        b = synthetic.sparse_unit_rows(5001, 352, seed=6).astype(np.float64)
        b[4000] = b[17]
        sa, sb_nn, d1, d2 = sfd.nearest_neighbors(a, b)
        o_sa, o_sb, o_nn, o_d1, dmat = matching_oracle.nearest(a, b)
        assert np.array_equal(sa, o_sa) and np.array_equal(sb_nn, o_sb[o_nn]) and np.array_equal(d1, o_d1)
        assert np.array_equal(d2, np.partition(dmat, 1, axis=1)[:, 1])
        # root + workers: rank 0 calls the package's ordinary matchers, rank 1 lends its GPU; the same matches as alone
        from shot_fpfh_b200 import matching as m

        if rank == 0:
            a[::11] = 0.0
            alone = (m.basic_matching(a, b), m.match_descriptors(a, b, m.threshold_filter, True, False, 100,
                                                                  threshold_multiplier=1.2),
                     m.double_matching_with_rejects(a, b, 0.9, verbose=False))
            sfd.start_root_service()
            shared = (m.basic_matching(a, b), m.match_descriptors(a, b, m.threshold_filter, True, False, 100,
                                                                   threshold_multiplier=1.2),
                      m.double_matching_with_rejects(a, b, 0.9, verbose=False))
            sfd.stop_root_service()
            for x, y in zip(alone, shared):
                assert np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]) and x[0].shape[0] > 50
        else:
            assert sfd.serve() == 5  # the warm-up, basic 1, reciprocal 2, ratio 1
        with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_nccl(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run under `gpurun --gpus 2`)")
    import torch.multiprocessing as mp

    mp.spawn(_nccl_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
