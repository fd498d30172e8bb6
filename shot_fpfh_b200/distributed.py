"""
Multi-GPU drivers: one process per GPU (torchrun), `torch.distributed` for the plumbing (NCCL over NVLink on the
box, gloo in the CPU tests). The hot path shards without any collective in the data path (BASELINE.json north_star,
SURVEY.md §8e):

  * SHOT:     query points by contiguous blocks, cloud replicated; rows are only all-gathered if the caller asks.
  * FPFH:     SPFH of the cloud points by contiguous blocks of the cell-sorted order -> ONE all-gather of the SPFH
              rows (they are needed for every neighbour) -> FPFH of the keypoints that live in the rank's block
              (same blocks for both stages: the neighbour lists are built once per point).
  * matching: the TARGET (reference) descriptor set by contiguous blocks, every rank sees all queries; each rank
              produces its exact (nearest index, d1, d2) against its shard -> ONE all-gather -> merge.

The sharding / gathering / merging logic is written against callables that compute one block, so that the same
code runs the CUDA kernels on the box and the NumPy oracle under gloo in tests/test_distributed_cpu.py.
"""

from __future__ import annotations

from typing import Callable

import torch
import torch.distributed as dist


def _marker(timings: dict):
    """mark(name): records a CUDA event under `name` when the caller asked for timings (bench.py)."""
    events = timings.setdefault("_events", [])

    def mark(name):
        if timings.get("enabled"):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            events.append((name, ev))

    return mark


def _elapsed(timings: dict) -> None:
    """Turns the recorded events into `timings["ms"][name]` = milliseconds since the previous mark."""
    events = timings.pop("_events", [])
    if timings.get("enabled") and len(events) > 1:
        timings["ms"] = {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(events[:-1], events[1:])}


def world(group=None) -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def block_bounds(n: int, parts: int, part: int) -> tuple[int, int]:
    """Contiguous blocks whose sizes differ by at most one: [lo, hi) of block `part`."""
    base, extra = divmod(n, parts)
    lo = part * base + min(part, extra)
    return lo, lo + base + (1 if part < extra else 0)


def all_gather_blocks(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """
    Concatenation over ranks of per-rank row blocks of sizes `block_bounds(n_total, world, r)`; one collective
    (`all_gather_into_tensor`) on blocks padded to the largest size.
    """
    rank, size = world(group)
    if size == 1:
        return local
    lo, hi = block_bounds(n_total, size, rank)
    assert local.shape[0] == hi - lo, (local.shape, lo, hi)
    pad = block_bounds(n_total, size, 0)[1]
    if hi - lo == pad:
        padded = local.contiguous()
    else:
        padded = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded[: hi - lo] = local
    out = torch.empty((size * pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    if n_total % size == 0:  # equal blocks: the gathered tensor IS the concatenation
        return out
    pieces = []
    for r in range(size):
        rlo, rhi = block_bounds(n_total, size, r)
        pieces.append(out[r * pad : r * pad + (rhi - rlo)])
    return torch.cat(pieces, dim=0)


def upload_replicated(host_rows, dtype=torch.float64, group=None, to_device=None) -> torch.Tensor:
    """
    A host array that every rank holds -> the same device tensor on every rank, without every rank pushing all of it
    through PCIe: rank r uploads only block r of the rows and ONE all-gather over NVLink completes the tensor (on an
    8-GPU box the ranks share the host's memory and PCIe switches: eight full uploads of the 563 MB scan descriptors
    took 25 ms, an eighth each plus the all-gather a few).
    """
    if to_device is None:  # (the gloo tests pass a CPU stand-in)
        from .device import upload as to_device

    rank, size = world(group)
    n = int(host_rows.shape[0])
    if size == 1 or n < size:
        return to_device(host_rows, dtype)
    lo, hi = block_bounds(n, size, rank)
    return all_gather_blocks(to_device(host_rows[lo:hi], dtype), n, group)


def sharded_rows(n_items: int, compute_block: Callable[[int, int], torch.Tensor], gather: bool = True, group=None):
    """Rows of items [0, n_items): every rank computes its block; optionally all-gathered to every rank."""
    rank, size = world(group)
    lo, hi = block_bounds(n_items, size, rank)
    local = compute_block(lo, hi)
    return all_gather_blocks(local, n_items, group) if gather else local


def sharded_fpfh(
    n_points: int,
    n_keypoints: int,
    spfh_block: Callable[[int, int], torch.Tensor],
    fpfh_block: Callable[[torch.Tensor, int, int], torch.Tensor],
    gather: bool = True,
    group=None,
):
    """
    `spfh_block(first, end)` -> SPFH rows of the cell-sorted points [first, end); one all-gather; then
    `fpfh_block(spfh_all, lo, hi)` -> FPFH rows of the keypoints [lo, hi).
    """
    rank, size = world(group)
    s0, s1 = block_bounds(n_points, size, rank)
    spfh_all = all_gather_blocks(spfh_block(s0, s1), n_points, group)
    lo, hi = block_bounds(n_keypoints, size, rank)
    local = fpfh_block(spfh_all, lo, hi)
    return all_gather_blocks(local, n_keypoints, group) if gather else local


def merge_nearest(d1: torch.Tensor, nn: torch.Tensor, d2: torch.Tensor):
    """
    (parts, Q) exact results against disjoint target shards (shard p holds smaller target indices than shard
    p + 1; nn are GLOBAL indices, -1 / +inf where a shard is empty) -> (nn, d1, d2) against the union:
    nearest = smallest d1, lowest index on ties (= first shard among equals); second = second smallest of the
    multiset of all shards' d1 and d2.
    """
    best, part = torch.min(d1, dim=0)  # first minimal value along dim 0 on ties
    nearest = torch.gather(nn, 0, part.unsqueeze(0)).squeeze(0)
    both = torch.cat([d1, d2], dim=0)
    second = torch.sort(both, dim=0).values[1] if both.shape[0] > 1 else torch.full_like(best, float("inf"))
    return nearest, best, second


def sharded_nearest(
    n_targets: int,
    nearest_in_shard: Callable[[int, int], tuple[torch.Tensor, torch.Tensor, torch.Tensor]],
    group=None,
):
    """
    `nearest_in_shard(lo, hi)` -> (nn global int64 (Q,), d1 float64 (Q,), d2 float64 (Q,)) against the target rows
    [lo, hi). One all-gather of a packed (Q, 3) float64 tensor per rank (indices < 2**53 are exact in float64).
    """
    rank, size = world(group)
    lo, hi = block_bounds(n_targets, size, rank)
    nn, d1, d2 = nearest_in_shard(lo, hi)
    if size == 1:
        return nn, d1, d2
    packed = torch.stack([d1.double(), nn.double(), d2.double()], dim=1).contiguous()
    flat = torch.empty((size * packed.shape[0], 3), dtype=torch.float64, device=packed.device)
    dist.all_gather_into_tensor(flat, packed, group=group)  # concatenated along dim 0 (gloo and NCCL agree on this)
    out = flat.view(size, packed.shape[0], 3)
    if out.is_cuda:  # one kernel (sf_nearest_merge); merge_nearest states the same rule in torch for the gloo tests
        from . import ops

        return ops.nearest_merge(out)
    nn_m, d1_m, d2_m = merge_nearest(out[:, :, 0], out[:, :, 1].long(), out[:, :, 2])
    return nn_m, d1_m, d2_m


# ---------------------------------------------------------------------------------------------------------------
# CUDA drivers (one process per GPU)
# ---------------------------------------------------------------------------------------------------------------
def shot_single_scale(point_cloud, normals, keypoints, radius, normalize=True, min_neighborhood_size=100, gather=True,
                      out_dtype=torch.float32, group=None, partition: str = "blocks"):
    """
    SHOT rows (device tensor) of the keypoint COORDINATES over the ranks.
    partition="blocks": queries by contiguous block, every rank grid-sorts the whole (replicated) cloud.
    partition="slabs" : the halo scheme — see `shot_single_scale_slabs`.
    """
    if partition == "slabs":
        return shot_single_scale_slabs(point_cloud, normals, keypoints, radius, normalize, min_neighborhood_size, gather,
                                       out_dtype, group)
    from . import ops
    from .descriptors.fpfh import _cached_grid
    from .device import upload

    pts, nrm, kp = upload(point_cloud), upload(normals), upload(keypoints)
    grid = _cached_grid().build(pts, nrm, radius)  # (the handle keeps its buffers between calls)

    def block(lo, hi):
        q = kp[lo:hi].contiguous()
        return ops.shot_single_scale(grid, q, radius, min_neighborhood_size, normalize, out_dtype=out_dtype)[0]

    out = sharded_rows(int(kp.shape[0]), block, gather, group)
    torch.cuda.synchronize()
    return out


# ---- halo partition ---------------------------------------------------------------------------------------------
# A rank owns a SLAB of the cloud's bounding box along its longest axis — whole layers of grid cells, cut so that the
# slabs hold about the same number of keypoints — and sorts only the points of its slab and of the layers of cells next
# to it (the halo: a neighbour within the radius lies at most one cell away), in the cell geometry of the WHOLE cloud
# (Grid.build(box=...)). Every one of the 27 cells around a keypoint of the slab then holds exactly the points it
# holds in the grid of the whole cloud, in the same order, so each rank's rows are bit-identical to the unsharded ones.
def cell_layer(x: torch.Tensor, origin: float, cell: float) -> torch.Tensor:
    """Layer of grid cells a coordinate falls in: csrc/sf_common.cuh cell_coord without its clamp."""
    return torch.floor((x - origin) * (1.0 / cell)).long()


def slab_bounds(keypoint_layers: torch.Tensor, parts: int) -> list[int]:
    """parts + 1 layer numbers: part r owns the layers [b[r], b[r + 1]); cut at the quantiles of the keypoints' layers
    (a layer is never split, so a crowded layer makes its slab larger; slabs may be empty)."""
    q = int(keypoint_layers.shape[0])
    if q == 0:
        return [0] * (parts + 1)
    ordered = torch.sort(keypoint_layers).values
    cuts = [int(ordered[min(q - 1, (q * r) // parts)].item()) for r in range(1, parts)]
    bounds = [int(ordered[0].item())] + cuts + [int(ordered[-1].item()) + 1]
    for r in range(1, parts + 1):
        bounds[r] = max(bounds[r], bounds[r - 1])
    return bounds


def slab_members(keypoint_layers, point_layers, bounds, part: int):
    """(indices of the part's keypoints, indices of the points it must sort: its slab plus two layers of halo — one is
    needed, the second makes a one-off disagreement in a layer number harmless), both ascending."""
    lo, hi = bounds[part], bounds[part + 1]
    mine = torch.nonzero((keypoint_layers >= lo) & (keypoint_layers < hi)).squeeze(1)
    halo = torch.nonzero((point_layers >= lo - 2) & (point_layers <= hi + 1)).squeeze(1)
    return mine, halo


def sharded_rows_by_slab(points, keypoints, radius: float, geometry: Callable, rows_of: Callable, width: int, gather=True,
                         group=None, out_dtype=torch.float32):
    """
    `geometry(lo, hi, radius)` -> {"cell": edge}; `rows_of(point_idx, keypoint_idx, (lo, hi))` -> rows of those
    keypoints from a grid over those points built in the box (lo, hi). Returns the (Q, width) rows on every rank
    (each row is written by exactly one rank: a sum-reduction of otherwise zero rows moves them exactly), or
    (keypoint indices, rows) of this rank when `gather` is false.
    """
    rank, size = world(group)
    lo, hi = points.min(dim=0).values, points.max(dim=0).values  # (what sf_grid_build's bounding-box pass finds)
    axis = int(torch.argmax(hi - lo).item())
    box = (tuple(float(v) for v in lo.tolist()), tuple(float(v) for v in hi.tolist()))
    cell = float(geometry(box[0], box[1], radius)["cell"])
    k_layers = cell_layer(keypoints[:, axis], box[0][axis], cell)
    p_layers = cell_layer(points[:, axis], box[0][axis], cell)
    bounds = slab_bounds(k_layers, size)
    mine, halo = slab_members(k_layers, p_layers, bounds, rank)
    if int(mine.shape[0]) > 0 and int(halo.shape[0]) > 0:
        local = rows_of(halo, mine, box)
    else:  # no keypoint here, or keypoints with nothing around them: zero rows, as the whole-cloud run gives
        local = torch.zeros((int(mine.shape[0]), width), dtype=out_dtype, device=points.device)
    if not gather:
        return mine, local
    full = torch.zeros((int(keypoints.shape[0]), width), dtype=local.dtype, device=local.device)
    full[mine] = local
    if size > 1:
        dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)
    return full


def shot_single_scale_slabs(point_cloud, normals, keypoints, radius, normalize=True, min_neighborhood_size=100,
                            gather=True, out_dtype=torch.float32, group=None):
    """The halo scheme on the GPUs: the raw cloud reaches every rank once (an N-th over PCIe each, the rest over
    NVLink), each rank sorts its slab + halo only and computes the rows of its slab's keypoints."""
    from . import ops
    from .descriptors.fpfh import _cached_grid
    from .device import grid_geometry, upload

    pts, nrm = upload_replicated(point_cloud, group=group), upload_replicated(normals, group=group)
    kp = upload(keypoints)
    grid = _cached_grid()  # (the handle keeps its buffers between calls)

    def rows_of(point_idx, keypoint_idx, box):
        grid.build(pts[point_idx].contiguous(), nrm[point_idx].contiguous(), radius, box=box)
        rows = ops.shot_single_scale(grid, kp[keypoint_idx].contiguous(), radius, min_neighborhood_size, normalize,
                                     out_dtype=out_dtype)[0]
        torch.cuda.synchronize()
        assert grid.poll() == 0, "a point of the slab lies outside the cloud's bounding box"
        return rows

    out = sharded_rows_by_slab(pts, kp, float(radius), grid_geometry, rows_of, 352, gather, group, out_dtype)
    torch.cuda.synchronize()
    return out


def sharded_fpfh_by_position(
    n_points: int,
    positions: torch.Tensor,
    width: int,
    spfh_block: Callable[[int, int], torch.Tensor],
    fpfh_rows: Callable[[torch.Tensor, torch.Tensor], torch.Tensor],
    gather: bool = True,
    group=None,
):
    """
    FPFH with BOTH stages sharded by the same blocks of the cell-sorted cloud, so that the neighbour lists a rank
    builds for its SPFH rows also serve its FPFH rows: `positions[q]` = cell-sorted position of keypoint q;
    `spfh_block(first, end)` -> SPFH rows of the block; ONE all-gather; `fpfh_rows(spfh_all, mine)` -> rows of the
    keypoints `mine` (ascending indices into `positions`) whose position lies in this rank's block.
    Returns (mine, rows) when `gather` is false; else the (Q, width) rows in keypoint order on every rank (each
    row is written by exactly one rank: a sum-reduction of otherwise zero rows moves them exactly).
    """
    rank, size = world(group)
    first, end = block_bounds(n_points, size, rank)
    spfh_all = all_gather_blocks(spfh_block(first, end), n_points, group)
    mine = torch.nonzero((positions >= first) & (positions < end)).squeeze(1)
    local = fpfh_rows(spfh_all, mine)
    if not gather:
        return mine, local
    full = torch.zeros((positions.shape[0], width), dtype=local.dtype, device=local.device)
    full[mine] = local
    if size > 1:
        dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)
    return full


def fpfh(keypoints_indices, cloud_points, normals, radius, n_bins, decorrelated=False, gather=True,
         out_dtype=torch.float32, group=None, timings: dict | None = None):
    """
    FPFH rows (device tensor) of the keypoint INDICES over the ranks: every rank builds the grid of the (replicated)
    cloud and runs the fused driver on ITS block of the cell-sorted points — one scan of the candidates for the
    lists, SPFH rows of the block — then ONE all-gather of the SPFH rows, then the FPFH rows of the keypoints that
    live in the block, from the same lists. `gather`: (Q, width) rows on every rank, else (keypoint ordinals, rows)
    of this rank.
    """
    from . import ops
    from .descriptors.fpfh import _cached_grid
    from .device import upload

    t = timings if timings is not None else {}
    mark = _marker(t)
    mark("start")
    pts, nrm = upload_replicated(cloud_points, group=group), upload_replicated(normals, group=group)
    kp = upload(keypoints_indices, torch.int64)
    mark("cloud_on_every_rank")
    grid = _cached_grid().build(pts, nrm, radius)  # the handle keeps its buffers between calls
    _, inv_perm = ops.grid_permutation(grid)
    positions = inv_perm[kp].long()
    mark("grid")
    state = {}

    def spfh_block(first, end):
        state["block"] = ops.FpfhBlock(grid, radius, n_bins, decorrelated, first, end - first, pts.device)
        rows = state["block"].spfh()
        mark("search_and_spfh_of_block")
        return rows

    def fpfh_rows(spfh_all, mine):
        mark("spfh_all_gather")
        rows = state["block"].rows(spfh_all.contiguous(), kp[mine].contiguous(), out_dtype=out_dtype)
        mark("fpfh_rows_of_block")
        return rows

    width = 3 * n_bins if decorrelated else n_bins**3
    out = sharded_fpfh_by_position(grid.n, positions, width, spfh_block, fpfh_rows, gather, group)
    torch.cuda.synchronize()
    _elapsed(t)
    return out


def _nearest_of_device_rows(a, ref_block: Callable[[int, int], torch.Tensor], n_ref: int, k: int, group, mark):
    """The sharded search once the scan rows `a` (float64, device) are on every rank; `ref_block(lo, hi)` -> this rank's
    block of the reference rows on the device."""
    from . import ops
    from .matching.matching import exact_nearest, largest, pack_scale

    rows_a, a_top = ops.nonempty_rows(a, want_absmax=True)
    qa = int(rows_a.shape[0])

    def shard(lo, hi):
        b = ref_block(lo, hi)
        rows_b, b_top = (ops.nonempty_rows(b, want_absmax=True) if hi > lo
                         else (torch.empty(0, dtype=torch.int64, device=a.device), 0.0))
        top = torch.tensor([largest(a_top, b_top)], dtype=torch.float64, device=a.device)
        if world(group)[1] > 1:
            dist.all_reduce(top, op=dist.ReduceOp.MAX, group=group)  # the same float16 scale on every rank
        scale = pack_scale(float(top.item()))
        mark("shard_uploaded")
        if int(rows_b.shape[0]) == 0 or qa == 0:
            inf = torch.full((qa,), float("inf"), dtype=torch.float64, device=a.device)
            return torch.full((qa,), -1, dtype=torch.int64, device=a.device), inf, inf.clone()
        nn, d1, d2, _ = exact_nearest(a, rows_a, b, rows_b, scale, k, want_second=True)
        mark("shard_searched")
        return rows_b[nn.long()].long() + lo, d1, d2  # original reference row ids

    nn, d1, d2 = sharded_nearest(n_ref, shard, group)
    mark("gathered_and_merged")
    return rows_a, nn, d1, d2


def nearest_neighbors(scan_descriptors, ref_descriptors, k: int = 8, group=None, timings: dict | None = None):
    """
    Exact nearest / second-nearest reference row of every non-empty scan row, the reference set sharded over the
    ranks by contiguous blocks of rows: a rank uploads ITS block of the scan rows (one all-gather over NVLink gives
    every rank all of them) and ITS block of reference rows, emits its exact (nearest, d1, d2) against that block,
    and ONE all-gather + merge gives the result against the union
    (lowest reference index on ties, as `cdist().argmin()`). The float16 shortlist uses the same scale on every rank
    (a MAX all-reduce of one scalar). Returns host arrays (scan row ids, ref row ids of the nearest, d1, d2),
    identical on every rank.
    """
    from .device import upload

    ref = ref_descriptors
    t = timings if timings is not None else {}
    mark = _marker(t)
    mark("start")
    a = upload_replicated(scan_descriptors, group=group)  # an N-th over PCIe per rank, the rest over NVLink
    mark("scan_rows_on_every_rank")
    rows_a, nn, d1, d2 = _nearest_of_device_rows(a, lambda lo, hi: upload(ref[lo:hi]), int(ref.shape[0]), k, group, mark)
    torch.cuda.synchronize()
    _elapsed(t)
    return rows_a.cpu().numpy(), nn.cpu().numpy(), d1.cpu().numpy(), d2.cpu().numpy()


# ---------------------------------------------------------------------------------------------------------------
# Root + workers: ONE process runs the caller's code, the other ranks lend their GPUs
# ---------------------------------------------------------------------------------------------------------------
# The drivers above are SPMD: every rank holds the inputs on the host and calls the same function. A program that is not
# written that way — the reference's `scripts/register_point_clouds.py` is a single process — runs on rank 0 only, and
# the other ranks sit in `serve()`: rank 0 announces a request (a small pickled dict), sends the operands over NVLink
# from ITS device (they are usually there already: the rows a descriptor call left behind), every rank works on its
# share, and the merged result lands on rank 0. Matching is served this way (the one stage of a 10M-point registration
# that is GPU-bound: two 1M x 1M x 352 searches); descriptors of one cloud are not worth sharding across processes for
# this API — their time is the host-side construction of the float64 result, which only rank 0 wants.
_SERVICE = {"group": None, "active": False}
_HANDLERS: dict[str, Callable] = {}


def root_service_active() -> bool:
    return bool(_SERVICE["active"]) and world(_SERVICE["group"])[1] > 1


def start_root_service(group=None, warm: bool = True) -> None:
    """Rank 0: from now on the matchers of this package (and of a reference rebound by dropin) use every rank.
    `warm`: one small search is served at once, so that every rank has loaded the kernels, sized its allocator and
    opened its NCCL channels before the first real request (a service is up before the work arrives)."""
    assert world(group)[0] == 0, "the root service runs on rank 0"
    _SERVICE.update(group=group, active=True)
    from . import device

    device.RANKS_SHARING_HOST = 1  # the workers build no host arrays: rank 0 keeps the host's cores for its results
    if warm and root_service_active() and torch.cuda.is_available():
        g = torch.Generator(device="cuda").manual_seed(1)
        rows = torch.rand((8192, 352), generator=g, device="cuda", dtype=torch.float64)
        nearest_neighbors_from_root(rows[:4096].contiguous(), rows, 8)
        torch.cuda.synchronize()


def stop_root_service() -> None:
    """Rank 0: releases the workers from serve()."""
    if root_service_active():
        dist.broadcast_object_list([{"op": "stop"}], src=0, group=_SERVICE["group"])
    _SERVICE.update(active=False)
    from . import device

    device.RANKS_SHARING_HOST = None


def serve(group=None) -> int:
    """Ranks other than 0: answer rank 0's requests until it stops the service. Returns the number served."""
    served = 0
    while True:
        box = [None]
        dist.broadcast_object_list(box, src=0, group=group)
        request = box[0]
        if request["op"] == "stop":
            return served
        _HANDLERS[request["op"]](None, None, request, group)
        served += 1


def _broadcast_rows(rows, shape, dtype, device, group):
    t = rows.contiguous() if rows is not None else torch.empty(shape, dtype=dtype, device=device)
    dist.broadcast(t, src=0, group=group)
    return t


def _serve_nearest(a, b, request, group):
    """Both row sets travel from rank 0's device to every rank (two broadcasts over NVLink); the reference set is then
    searched by blocks as in `nearest_neighbors`. SF_TRACE_SERVICE=1: rank 0 prints where the time of a request goes."""
    import os
    import time

    on_cpu = request["device"] == "cpu"
    device = torch.device("cpu") if on_cpu else torch.device("cuda", torch.cuda.current_device())
    trace = os.environ.get("SF_TRACE_SERVICE") == "1" and world(group)[0] == 0 and not on_cpu
    stamps = []

    def mark(name):
        if trace:
            torch.cuda.synchronize()
            stamps.append((name, time.perf_counter()))

    mark("start")
    a = _broadcast_rows(a, (request["qa"], request["width"]), torch.float64, device, group)
    b = _broadcast_rows(b, (request["qb"], request["width"]), torch.float64, device, group)
    mark("operands_on_every_rank")
    search = _HANDLERS.get("nearest_core", _nearest_of_device_rows)
    out = search(a, lambda lo, hi: b[lo:hi], request["qb"], request["k"], group, mark)
    if trace:
        print("  service: " + ", ".join(f"{n} {1e3 * (t1 - t0):.1f} ms" for (_, t0), (n, t1) in zip(stamps[:-1], stamps[1:])),
              flush=True)
    return out


_HANDLERS["nearest"] = _serve_nearest


def nearest_neighbors_from_root(a: torch.Tensor, b: torch.Tensor, k: int = 8):
    """Rank 0 (after start_root_service): exact nearest / second nearest row of `b` for every non-empty row of `a`
    (float64 rows on rank 0's device), searched by all ranks. Device tensors (rows_a, nearest row of b, d1, d2)."""
    group = _SERVICE["group"]
    request = {"op": "nearest", "qa": int(a.shape[0]), "qb": int(b.shape[0]), "width": int(a.shape[1]), "k": int(k),
               "device": "cpu" if a.device.type == "cpu" else "cuda"}
    dist.broadcast_object_list([request], src=0, group=group)
    return _serve_nearest(a, b, request, group)
