"""
Descriptor matching with the reference's signatures (shot_fpfh/matching/matching.py:9-146, :149-169, :172-221).

The reference materialises the full float64 `cdist` matrix and takes `argmin`. Here (csrc/match.cu, match_tc.cu):
non-empty rows are compacted on the device, a tensor-core distance GEMM on float16 copies of the rows keeps a
k-candidate shortlist per query without ever writing the matrix, the shortlist is re-ranked with the exact float64
distance accumulated in SciPy's order, and a CERTIFICATE decides per query whether the shortlist provably contained
the nearest (second-nearest) neighbour: the k-th shortlist score bounds the distance to everything outside it, up to
the float16 rounding of the operands and the float32 accumulation (csrc/match.cu::certify_kernel). Queries that are
not certified — adversarial near-ties, rows that quantise to zero next to a large one — are redone exhaustively in
float64, so the returned indices and distances are `cdist(...).argmin()`'s in every case; `LAST_STATS` counts them.

Parity limits, stated: (1) descriptors that this package produced are float32 values widened to float64 (1e-7
relative against the reference's float64 rows): matching is exact on the rows it is given, near-ties between the two
versions of a row can resolve differently; (2) NaN or infinite entries raise ValueError (the reference's argmin over
NaN distances silently returns the first NaN column).
"""

from __future__ import annotations

import logging
from typing import Callable

import numpy as np
import numpy.typing as npt
import torch

from .. import ops
from ..device import device_rows_of, upload

DEFAULT_SHORTLIST = 8
USE_TENSOR_CORES = True  # the tests switch it off to cross-check the two shortlist kernels
LAST_STATS = {"queries": 0, "fallback_rows": 0, "handoff": 0}  # of the last public call (measurement)


class DeviceMatch:
    """Result of one directed nearest-neighbour search, on the host."""

    def __init__(self, rows_a, rows_b, nn, d1, d2):
        self.rows_a, self.rows_b, self.nn, self.d1, self.d2 = rows_a, rows_b, nn, d1, d2


def _to_device(desc) -> torch.Tensor:
    """float64 device rows of `desc`; the float32 rows a descriptor call left on the device when `desc` is the very
    array that call returned (device.remember_device_rows) — no PCIe transfer then."""
    cached = device_rows_of(desc)
    if cached is not None:
        LAST_STATS["handoff"] += 1
        return cached.double()
    return upload(desc)


def _prepare(desc) -> tuple[torch.Tensor, torch.Tensor, float]:
    dev = _to_device(desc)
    if dev.dim() != 2:
        raise ValueError("descriptors must be a 2-D (n, width) array")
    rows, top = ops.nonempty_rows(dev, want_absmax=True)
    return dev, rows, top


def largest(*tops: float) -> float:
    """max that keeps a NaN / inf (Python's max drops a NaN that is not first)."""
    return float(np.max(np.asarray(tops, dtype=np.float64)))  # np.max propagates NaN


def pack_scale(top: float) -> float:
    """Power-of-two scale that brings the largest |entry| just below 1 (float16 operands)."""
    if not np.isfinite(top):
        raise ValueError("descriptors contain NaN or infinite values")
    return float(2.0 ** np.floor(np.log2(1.0 / max(top, 1e-300))))


def exact_nearest(a_dev, rows_a, b_dev, rows_b, scale: float, k: int = DEFAULT_SHORTLIST, want_second: bool = False,
                  packed_b=None):
    """
    For every row rows_a of a: the nearest and second-nearest rows rows_b of b (positions in rows_b, float64
    distances, lowest index on ties) — shortlist GEMM, exact re-rank, certificate, exhaustive redo of the queries
    the certificate rejects. `want_second`: the second distance is certified too (the ratio test needs it).
    Returns (nn int32, d1, d2, packed_b) device tensors; packed_b = (float16 rows, squared norms, largest norm) of b.
    """
    width = int(a_dev.shape[1])
    # the tcgen05 kernel keeps a 128-row tile of <= 384 columns resident; wider rows (multi-scale SHOT, FPFH with
    # n_bins >= 8) take the CUDA-core shortlist kernel, which loops over any width
    tensor_cores = ops.padded_width(width) <= 384 and USE_TENSOR_CORES
    a_packed, a_sqnorm = ops.match_pack(a_dev, rows_a, scale)
    if packed_b is None:
        b_packed, b_sqnorm = ops.match_pack(b_dev, rows_b, scale)
        packed_b = (b_packed, b_sqnorm, float(b_sqnorm.max().sqrt().item()))
    b_packed, b_sqnorm, b_norm_max = packed_b
    score, cand = ops.match_topk(a_packed, b_packed, b_sqnorm, k, 0, tensor_cores)
    nn, d1, d2 = ops.match_rerank(a_dev, rows_a, b_dev, rows_b, cand)
    flags = ops.match_certify(score, a_sqnorm, d1, d2, scale, b_norm_max, width, int(rows_b.shape[0]), want_second)
    which = torch.nonzero(flags).squeeze(1)
    LAST_STATS["queries"] += int(rows_a.shape[0])
    if int(which.shape[0]):
        LAST_STATS["fallback_rows"] += int(which.shape[0])
        # (within the current SECOND distance: the redo then returns both neighbours exactly)
        # |a| + |b| over the SCALED rows, from the float16 operands' norms (a margin for their rounding and underflow)
        norm_bound = 1.01 * (float(a_sqnorm.max().sqrt().item()) + b_norm_max) + float(np.sqrt(width)) * 2.0**-23
        nn[which], d1[which], d2[which] = exhaustive_redo(a_dev, rows_a, which, b_dev, rows_b, d2[which], scale, norm_bound)
    return nn, d1, d2, packed_b


def exhaustive_redo(a_dev, rows_a, which, b_dev, rows_b, limit, scale: float, norm_bound: float):
    """Exact (nn, d1, d2) of the flagged queries `which`: one pass over all the targets lists every target within `limit`
    (the exact distance the re-rank found: an upper bound on the true one) — in float32 on the rows times `scale`, with a
    proven slack, `norm_bound` bounding |a| + |b| of the scaled rows —, the float64 re-rank decides among the listed; the rare query with more than 16 of them gets a
    full float64 scan of its own."""
    finite = torch.isfinite(limit)
    limit = torch.where(finite, limit, torch.full_like(limit, 1e300))  # (fewer than two candidates so far: everything)
    cand16 = ops.match_exhaustive(a_dev, rows_a, which, limit, b_dev, rows_b, scale, norm_bound)
    crowded = torch.nonzero(cand16[:, 0] == -2).squeeze(1)
    if int(crowded.shape[0]):
        full = torch.full((int(rows_a.shape[0]), 16), -1, dtype=torch.int32, device=a_dev.device)
        ops.match_exhaustive_topk(a_dev, rows_a, which[crowded].contiguous(), b_dev, rows_b, full)
        cand16[crowded] = full[which[crowded]]
    return ops.match_rerank(a_dev, rows_a[which].contiguous(), b_dev, rows_b, cand16)


def nearest_neighbors_device(a_dev, rows_a, b_dev, rows_b, k: int = DEFAULT_SHORTLIST, want_second: bool = False,
                             top: float | None = None):
    """
    For every non-empty row of a: the nearest and second-nearest non-empty rows of b (float64 distances, lowest
    index on ties). Returns device tensors (nn positions in rows_b, d1, d2).
    """
    qa, qb = int(rows_a.shape[0]), int(rows_b.shape[0])
    if qa == 0:
        e = torch.empty(0, device=a_dev.device)
        return e.to(torch.int32), e.to(torch.float64), e.to(torch.float64)
    if qb == 0:
        raise ValueError("attempt to get argmin of an empty sequence")  # what NumPy raises in the reference
    if top is None:
        top = largest(ops.nonempty_rows(a_dev, want_absmax=True)[1], ops.nonempty_rows(b_dev, want_absmax=True)[1])
    nn, d1, d2, _ = exact_nearest(a_dev, rows_a, b_dev, rows_b, pack_scale(top), k, want_second)
    return nn, d1, d2


_PIPELINE_MIN_ROWS = 65536  # scan sets at least this large cross PCIe in chunks under the shortlist GEMM
_PIPELINE_CHUNK_ROWS = 3 * 148 * 128  # three full waves of the shortlist kernel's 128-query CTAs on 148 SMs


def _upload_pipelined(scan: np.ndarray, ref: np.ndarray, k: int, want_second: bool):
    """
    Host arrays in: the reference rows are copied first, then the scan rows in chunks of `_PIPELINE_CHUNK_ROWS` on a
    side stream; the shortlist GEMM + exact re-rank of chunk c (against all the reference rows) run while chunk c + 1
    is still crossing PCIe — at 200k x 200k x 352 the copy of the scan rows (563 MB, 10 ms) disappears under the GEMM.
    The chunks are QUERY rows, whose results are independent: cutting the reference rows instead restarts every
    query's top-k list per chunk, and the shortlist kernel pays ~1.5 ms per restart (measured: 4 chunks 27.3 ms against
    21.6 ms in one piece), which ate the overlap. Each chunk is packed with its own power-of-two scale; the packed
    reference rows are reused when the scale repeats (it does unless the chunks differ in magnitude).
    Returns (a_dev, rows_a, b_dev, rows_b, nn, d1, d2) with device tensors.
    """
    dev = torch.device("cuda", torch.cuda.current_device())
    a_host = torch.from_numpy(np.ascontiguousarray(scan, dtype=np.float64))
    b_host = torch.from_numpy(np.ascontiguousarray(ref, dtype=np.float64))
    na = int(a_host.shape[0])
    a_dev = torch.empty(a_host.shape, dtype=torch.float64, device=dev)
    b_dev = torch.empty(b_host.shape, dtype=torch.float64, device=dev)
    bounds = list(range(0, na, _PIPELINE_CHUNK_ROWS)) + [na]
    cur, side = torch.cuda.current_stream(), torch.cuda.Stream()
    start = torch.cuda.Event()
    start.record(cur)  # the destination buffers exist on the compute stream before the side stream writes them
    arrived = []
    with torch.cuda.stream(side):
        side.wait_event(start)
        b_dev.copy_(b_host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(side)
        arrived.append(ev)
        for c in range(len(bounds) - 1):
            a_dev[bounds[c]:bounds[c + 1]].copy_(a_host[bounds[c]:bounds[c + 1]], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(side)
            arrived.append(ev)
    cur.wait_event(arrived[0])
    rows_b, b_top = ops.nonempty_rows(b_dev, want_absmax=True)
    qb = int(rows_b.shape[0])
    packed_b: dict[float, tuple] = {}
    rows_a_parts, nn_parts, d1_parts, d2_parts = [], [], [], []
    for c in range(len(bounds) - 1):
        cur.wait_event(arrived[c + 1])
        chunk = a_dev[bounds[c]:bounds[c + 1]]
        rows_c, a_top = ops.nonempty_rows(chunk, want_absmax=True)
        if int(rows_c.shape[0]) == 0:
            continue
        if qb == 0:
            raise ValueError("attempt to get argmin of an empty sequence")  # what NumPy raises in the reference
        scale = pack_scale(largest(a_top, b_top))
        nn, d1, d2, packed_b[scale] = exact_nearest(chunk, rows_c, b_dev, rows_b, scale, k, want_second, packed_b.get(scale))
        rows_a_parts.append(rows_c + bounds[c])
        nn_parts.append(nn)
        d1_parts.append(d1)
        d2_parts.append(d2)
    if not rows_a_parts:
        e = torch.empty(0, device=dev)
        return (a_dev, torch.empty(0, dtype=torch.int64, device=dev), b_dev, rows_b, e.to(torch.int32), e.to(torch.float64),
                e.to(torch.float64))
    return a_dev, torch.cat(rows_a_parts), b_dev, rows_b, torch.cat(nn_parts), torch.cat(d1_parts), torch.cat(d2_parts)


def _match(scan, ref, k=DEFAULT_SHORTLIST, reverse=False, want_second=False, tensor_cores=True):
    global USE_TENSOR_CORES
    previous, USE_TENSOR_CORES = USE_TENSOR_CORES, bool(tensor_cores)
    try:
        return _match_impl(scan, ref, k, reverse, want_second)
    finally:
        USE_TENSOR_CORES = previous


def _match_with_workers(scan, ref, k, reverse):
    """The same result with the reference rows searched by every rank of the process group: this process is rank 0 of a
    root + workers job (distributed.start_root_service; the other ranks sit in distributed.serve)."""
    from .. import distributed

    a_dev, b_dev = _to_device(scan), _to_device(ref)
    if a_dev.dim() != 2 or b_dev.dim() != 2:
        raise ValueError("descriptors must be a 2-D (n, width) array")
    if a_dev.shape[1] != b_dev.shape[1]:
        raise ValueError("XA and XB must have the same number of columns (i.e. feature dimension.)")
    rows_a, nn, d1, d2 = distributed.nearest_neighbors_from_root(a_dev, b_dev, k)
    # nn are row numbers of `ref` itself: rows_b is the identity here
    fwd = DeviceMatch(rows_a.cpu().numpy(), np.arange(int(b_dev.shape[0]), dtype=np.int64), nn.cpu().numpy().astype(np.int64),
                      d1.cpu().numpy(), d2.cpu().numpy())
    if not reverse:
        return fwd, None
    rows_b, nn_r, _, _ = distributed.nearest_neighbors_from_root(b_dev, a_dev, k)
    position_in_rows_a = np.full(int(a_dev.shape[0]), -1, dtype=np.int64)
    position_in_rows_a[fwd.rows_a] = np.arange(fwd.rows_a.shape[0])
    nn_reverse = np.full(int(b_dev.shape[0]), -1, dtype=np.int64)
    nn_reverse[rows_b.cpu().numpy()] = position_in_rows_a[nn_r.cpu().numpy()]
    return fwd, nn_reverse


def _match_impl(scan, ref, k, reverse, want_second):
    LAST_STATS.update(queries=0, fallback_rows=0, handoff=0)
    from .. import distributed

    if distributed.root_service_active() and np.ndim(scan) == 2 and np.ndim(ref) == 2:
        return _match_with_workers(scan, ref, k, reverse)
    pipelined = (
        isinstance(scan, np.ndarray) and isinstance(ref, np.ndarray) and scan.ndim == 2 and ref.ndim == 2
        and scan.shape[1] == ref.shape[1] and scan.shape[0] >= _PIPELINE_MIN_ROWS
        and device_rows_of(scan) is None  # rows that are already on the device do not cross PCIe at all
    )
    if pipelined:
        a_dev, rows_a, b_dev, rows_b, nn, d1, d2 = _upload_pipelined(scan, ref, k, want_second)
        top = None
    else:
        a_dev, rows_a, a_top = _prepare(scan)
        b_dev, rows_b, b_top = _prepare(ref)
        if a_dev.shape[1] != b_dev.shape[1]:
            raise ValueError("XA and XB must have the same number of columns (i.e. feature dimension.)")
        top = largest(a_top, b_top)
        nn, d1, d2 = nearest_neighbors_device(a_dev, rows_a, b_dev, rows_b, k, want_second, top)
    fwd = DeviceMatch(rows_a.cpu().numpy(), rows_b.cpu().numpy(), nn.cpu().numpy().astype(np.int64), d1.cpu().numpy(),
                      d2.cpu().numpy())
    if not reverse:
        return fwd, None
    nn_r, _, _ = nearest_neighbors_device(b_dev, rows_b, a_dev, rows_a, k, False, top)
    return fwd, nn_r.cpu().numpy().astype(np.int64)


def basic_matching(
    scan_descriptors: npt.NDArray[np.float64], ref_descriptors: npt.NDArray[np.float64]
) -> tuple[npt.NDArray[np.int64], npt.NDArray[np.int64]]:
    """Each non-empty scan descriptor with its nearest non-empty reference descriptor (matching.py:149-169)."""
    m, _ = _match(scan_descriptors, ref_descriptors)
    return m.rows_a, m.rows_b[m.nn]


def match_descriptors(
    scan_descriptors: npt.NDArray[np.float64],
    ref_descriptors: npt.NDArray[np.float64],
    filter_callback: Callable[..., npt.NDArray[np.bool_]] | None = None,
    filter_nonreciprocal: bool = False,
    verbose: bool = True,
    n_min_matches: int = 100,
    **kwargs: bool | int | float | tuple[float, float],
) -> tuple[npt.NDArray[np.int64], npt.NDArray[np.int64]]:
    """
    Nearest-neighbour matching with an optional filter on the nearest-neighbour distances and an optional
    reciprocity filter (matching.py:9-74, :138-146). The reciprocity check `D.argmin(0)[idx] == arange` is a second
    nearest-neighbour search with the roles swapped. (n_scales, n_points, width) inputs take the multi-scale
    "infinite-norm" branch (matching.py:76-136), see `_match_multiscale`.
    """
    if np.ndim(scan_descriptors) != 2:
        return _match_multiscale(
            scan_descriptors, ref_descriptors, filter_callback, filter_nonreciprocal, verbose, n_min_matches, **kwargs
        )
    logging.info("")
    logging.info("-- Matching descriptors based on Euclidian-norm proximity --")
    m, nn_reverse = _match(scan_descriptors, ref_descriptors, reverse=filter_nonreciprocal)
    distances = m.d1
    filtered = (
        filter_callback(distances, **kwargs) if filter_callback is not None else np.ones(distances.shape[0], dtype=bool)
    )
    if filter_nonreciprocal:
        reciprocal = nn_reverse[m.nn] == np.arange(m.nn.shape[0])
        if (final_mask := filtered & reciprocal).sum() >= n_min_matches:
            filtered = final_mask
        elif verbose:
            logging.warning("Too few reciprocal matches, keeping non-reciprocal matches.")
    if verbose:
        logging.info(f"Kept {filtered.sum()} matches out of {np.shape(scan_descriptors)[-2]} descriptors.")
    return m.rows_a[filtered], m.rows_b[m.nn[filtered]]


def _match_multiscale(scan, ref, filter_callback, filter_nonreciprocal, verbose, n_min_matches, **kwargs):
    """
    The 3-D branch of the reference's `match_descriptors` (matching.py:76-136): descriptors given as
    (n_scales, n_points, width); the distance between two points is the MINIMUM over the scales of their Euclidean
    descriptor distances, empty rows counting as `max_val = 1000`. The reference builds n_scales dense matrices; here
    each scale is one exact nearest-neighbour search on the GPU and the per-scale winners are merged on the host —
    `argmin_j min_s D_s[i, j]` is attained by the nearest neighbour of some scale (lowest j on ties).
    Kept literally: the reciprocity filter of this branch assigns into a temporary and is a no-op in the reference
    (matching.py:106-108); only its "too few matches" fallback has an effect.
    """
    logging.info("")
    logging.info("-- Matching descriptors based on infinite-norm proximity --")
    max_val = 1000.0
    scan, ref = np.asarray(scan), np.asarray(ref)
    n_scales, n_points, _ = scan.shape
    n_points_ref = ref.shape[1]
    distances = np.full(n_points, max_val)
    indices = np.zeros(n_points, dtype=np.int64)
    for scale in range(n_scales):
        if not ref[scale].any() or not scan[scale].any():
            continue
        m, _ = _match(scan[scale], ref[scale])
        d = np.minimum(m.d1, max_val)
        j = m.rows_b[m.nn]
        cur_d, cur_j = distances[m.rows_a], indices[m.rows_a]
        better = (d < cur_d) | ((d == cur_d) & (j < cur_j) & (d < max_val))
        distances[m.rows_a] = np.where(better, d, cur_d)
        indices[m.rows_a] = np.where(better, j, cur_j)
    filtered = (
        filter_callback(distances, **kwargs) if filter_callback is not None else np.ones(n_points, dtype=bool)
    ) & (distances < max_val)
    if filtered.sum() < n_min_matches and filter_nonreciprocal:
        logging.warning("Too few reciprocal matches, keeping non-reciprocal matches.")
        return match_descriptors(scan, ref, filter_callback, filter_nonreciprocal=False, verbose=verbose, **kwargs)
    if verbose:
        logging.info(f"Kept {filtered.sum()} matches out of {n_points} descriptors.")
    return np.arange(n_points)[filtered], np.arange(n_points_ref)[indices[filtered]]


def double_matching_with_rejects(
    scan_descriptors: npt.NDArray[np.float64],
    ref_descriptors: npt.NDArray[np.float64],
    threshold: float,
    verbose: bool = True,
) -> tuple[npt.NDArray[np.int64], npt.NDArray[np.int64]]:
    """
    Ratio test (matching.py:172-221). The reference raises on every input (three independent indexing bugs,
    SURVEY.md F3); this is what its code and docstring intend: with d1 <= d2 the distances to the nearest and
    second-nearest reference descriptors, keep the scan descriptors whose ratio d1 / d2 (1 where d2 == 0) is
    `>= threshold` — the comparison as written at matching.py:203-211 — and match them to their nearest neighbour.
    """
    m, _ = _match(scan_descriptors, ref_descriptors, want_second=True)
    d2 = np.where(np.isfinite(m.d2), m.d2, 0.0)  # a single candidate: no second neighbour -> ratio 1
    ratio = np.divide(m.d1, d2, out=np.ones_like(m.d1), where=d2 != 0)
    mask = ratio >= threshold
    if verbose:
        logging.info(f"Kept {mask.sum()} matches out of {np.shape(scan_descriptors)[0]} descriptors.")
    return m.rows_a[mask], m.rows_b[m.nn[mask]]
