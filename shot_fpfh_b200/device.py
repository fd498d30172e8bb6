"""
Device plumbing: PyTorch owns device memory and streams, the C ABI (see _lib.py) does the work.
Everything here fails loudly when no CUDA device is present — there is no CPU path.
"""

from __future__ import annotations

import os

import numpy as np
import torch

from ._lib import check, lib


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "shot_fpfh_b200 needs a CUDA device (built for NVIDIA B200, sm_100a); there is no CPU fallback."
        )
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr() -> int:
    return int(torch.cuda.current_stream().cuda_stream)


def ptr(t: torch.Tensor | None) -> int | None:
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous()
    return int(t.data_ptr())


def upload(a, dtype=torch.float64) -> torch.Tensor:
    """Host array (or tensor already on the device) -> contiguous device tensor of `dtype`."""
    dev = require_cuda()
    if isinstance(a, torch.Tensor):
        return a.to(device=dev, dtype=dtype).contiguous()
    np_dtype = {torch.float64: np.float64, torch.float32: np.float32, torch.int64: np.int64, torch.int32: np.int32}[dtype]
    host = np.ascontiguousarray(a, dtype=np_dtype)
    source = torch.from_numpy(host)
    if host.nbytes >= _STAGED_UPLOAD_BYTES and not source.is_pinned():
        # Pageable memory: the driver would stage the copy itself, synchronously and on one thread (50 MB of cloud:
        # 9 ms). The library's host threads copy it into a page-locked block of PyTorch's caching host allocator (the
        # allocator keeps the block until the copy queued below has run), and the DMA engine takes it from there.
        staging = torch.empty(host.shape, dtype=dtype, pin_memory=True)
        check(lib.sf_host_copy_begin(host.ctypes.data, staging.data_ptr(), host.nbytes, host_threads()))
        check(lib.sf_host_wait())
        return staging.to(dev, non_blocking=True)
    return source.to(dev, non_blocking=True)


_STAGED_UPLOAD_BYTES = 4 << 20


def download(t: torch.Tensor) -> np.ndarray:
    """
    Device tensor -> fresh host NumPy array, through page-locked memory: a pageable destination makes the driver
    stage the copy and costs ~10x (measured: 288 MB of SHOT rows, 100+ ms pageable vs ~12 ms pinned). The array
    is backed by a block of PyTorch's caching pinned allocator; the block goes back to the cache when the array
    (and every view of it) is released, so repeated calls do not re-pin memory.
    """
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host.numpy()


RANKS_SHARING_HOST: int | None = None  # None: the ranks of the node (LOCAL_WORLD_SIZE); set by a root + workers job


def host_threads(requested: int | None = None) -> int:
    """
    Host threads that rebuild the dense float64 result. Two of the cores this process may use are left to the
    calling thread and the CUDA driver's: with every core spinning in the pool the thread that launches the next
    block's kernels gets descheduled (measured on the 16-core B200 box: 16 pool threads 5.3 ms per call, 8 threads
    4.0 ms; the expansion itself saturates the host's memory bandwidth at 8 threads).
    """
    import os

    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        avail = os.cpu_count() or 1
    # one process per GPU (torchrun): the ranks of a node share its cores
    ranks = RANKS_SHARING_HOST if RANKS_SHARING_HOST else max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
    share = avail // ranks - 2
    return max(1, min(share, 8 if requested is None else int(requested)))


# ---- host result buffers ------------------------------------------------------------------------------------------
# The dense float64 arrays the reference API returns are WRITTEN BY HOST THREADS (never by DMA), so they are plain
# NumPy memory. What a fresh 288 MB buffer costs is its first touch (70 000 page faults, ~20 ms), so a buffer the caller
# has dropped is handed out again instead. Measured on the B200 hosts (same box, alternating runs): the pool's threads
# fill memory from cudaHostAlloc in 3.4 ms per call at C2, pageable memory in 4.4-5.7 ms (bimodal), pageable memory
# page-locked in place (cudaHostRegister) in 5.6-5.8 ms, huge-page-advised memory in 7.0 ms; but cudaHostAlloc costs
# ~100-200 ms per buffer. So: a FRESH buffer is pageable (a pipeline that keeps its results never pays the page-locked
# allocation), and the first time a dropped buffer would be handed out again — a caller that loops — it is replaced,
# once, by a cudaHostAlloc block that is then recycled.
# The pool owns the base arrays; the caller gets a VIEW, so the base's reference count tells when every reference to
# a result (views and slices included) is gone.
_RESULT_POOL: list[list] = []  # [base array, from cudaHostAlloc?]
_RESULT_POOL_BYTES = 4 << 30
_HUGE_PAGE_HINT_BYTES = 512 << 20


def result_buffer(shape) -> tuple[np.ndarray, np.ndarray]:
    """(float64 base array of `shape`, the view of it that the API returns — return THAT view or views of it)."""
    import sys

    shape = tuple(int(x) for x in shape)
    count = int(np.prod(shape)) if shape else 1
    entry = None
    for i in range(len(_RESULT_POOL)):
        # references: the pool's entry and getrefcount's argument -> 2 when no caller holds a view of it any more
        if _RESULT_POOL[i][0].size == count and sys.getrefcount(_RESULT_POOL[i][0]) == 2:
            entry = _RESULT_POOL.pop(i)
            if not entry[1] and entry[0].nbytes >= (8 << 20) and torch.cuda.is_available():
                entry = [torch.empty(count, dtype=torch.float64, pin_memory=True).numpy(), True]
            break
    if entry is None:
        entry = [np.empty(count, dtype=np.float64), False]
        if entry[0].nbytes >= _HUGE_PAGE_HINT_BYTES:
            # first touch of a LARGE fresh buffer (1M keypoints: 2.8 GB = 700 000 page faults, taken by eight writer
            # threads at once): ask for 2 MB pages. Measured through the reference's pipeline on the 10M-point pair,
            # alternating runs on three boxes (THP "madvise"): descriptor stage 0.71-0.88 s with the hint, every time;
            # without it 0.61 s on a good run but 1.2-3.3 s on every second one. At 288 MB (C2) it makes no difference.
            lib.sf_host_advise_huge(entry[0].ctypes.data, entry[0].nbytes)
    _RESULT_POOL.append(entry)
    held = sum(e[0].nbytes for e in _RESULT_POOL)
    while held > _RESULT_POOL_BYTES and len(_RESULT_POOL) > 1:  # the pool forgets its oldest arrays (their views live on)
        held -= _RESULT_POOL.pop(0)[0].nbytes
    return entry[0], entry[0].reshape(shape)


def download_widened(t: torch.Tensor, threads: int | None = None, blocks: int = 8) -> np.ndarray:
    """
    float32 device tensor -> fresh float64 host array holding the same values (float64(float32 x) is exact): the
    array a float64 kernel output + `download` would deliver, for half the PCIe bytes. The tensor crosses in
    `blocks` pieces through page-locked float32 staging; the library's host threads (csrc/host_io.cpp) widen piece
    b into the result while piece b + 1 is in flight. Measured on the B200 box, 1M x 33 rows: 4.05 ms against 4.8 ms
    for a float64 copy (the widening runs at the host's memory bandwidth: 132 MB read + 264 MB written while the
    DMA writes another 132 MB).
    """
    assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
    n = t.numel()
    result, result_arr = result_buffer(t.shape)
    if n == 0:
        return result_arr
    staging = torch.empty(n, dtype=torch.float32, pin_memory=True)
    flat = t.reshape(-1)
    blocks = max(1, min(int(blocks), n // (1 << 20)))
    bounds = [n * b // blocks for b in range(blocks + 1)]
    stream = torch.cuda.current_stream()
    events = []
    for b in range(blocks):  # all the copies are queued at once; the host follows them event by event
        staging[bounds[b]:bounds[b + 1]].copy_(flat[bounds[b]:bounds[b + 1]], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(stream)
        events.append(ev)
    n_threads = host_threads(threads)
    try:
        for b in range(blocks):
            events[b].synchronize()
            check(lib.sf_host_widen_begin(staging.data_ptr() + 4 * bounds[b], bounds[b + 1] - bounds[b],
                                          result.ctypes.data + 8 * bounds[b], n_threads))
    finally:
        check(lib.sf_host_wait())  # the staging buffer is read by the pool until here
    return result_arr


class DenseRowsDownload:
    """
    Consecutive blocks of float32 device rows -> the dense float64 host array the reference API returns:

        job = DenseRowsDownload(n_rows, width)
        job.push(block)    # queued on a side stream behind the block's kernels: the copy of block b runs while the
                           # kernels of block b + 1 do
        arr = job.finish() # follows the copies event by event; the library's host threads widen block b while
                           # block b + 1 is still crossing PCIe

    The values are the kernels' float32 values widened exactly.
    """

    def __init__(self, n_rows: int, width: int, threads: int | None = None) -> None:
        require_cuda()
        self.n_rows, self.width = int(n_rows), int(width)
        self.threads = host_threads(threads)
        self.result, self.result_array = result_buffer((self.n_rows, self.width))
        self.staging = torch.empty((self.n_rows, self.width), dtype=torch.float32, pin_memory=True)
        self.side = torch.cuda.Stream()
        self.filled = 0
        self._blocks: list[tuple[int, int, torch.cuda.Event, torch.Tensor]] = []

    def push(self, rows: torch.Tensor) -> None:
        assert rows.is_cuda and rows.dtype == torch.float32 and rows.is_contiguous()
        n = int(rows.shape[0])
        assert rows.dim() == 2 and rows.shape[1] == self.width and self.filled + n <= self.n_rows
        if n == 0:
            return
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        done = torch.cuda.Event()
        with torch.cuda.stream(self.side):
            self.side.wait_event(ready)
            self.staging[self.filled:self.filled + n].copy_(rows, non_blocking=True)
            done.record(self.side)
        self._blocks.append((self.filled, n, done, rows))  # `rows` stays alive until its copy has run
        self.filled += n

    def finish(self) -> np.ndarray:
        assert self.filled == self.n_rows, "DenseRowsDownload.finish before every row was pushed"
        try:
            for lo, n, done, _ in self._blocks:
                done.synchronize()
                check(lib.sf_host_widen_begin(self.staging.data_ptr() + 4 * lo * self.width, n * self.width,
                                              self.result.ctypes.data + 8 * lo * self.width, self.threads))
        finally:
            check(lib.sf_host_wait())  # the staging buffer is read by the pool until here
            self._blocks = []
        return self.result_array


class SparseRowsDownload:
    """
    Dense float32 device rows that are mostly zeros (SHOT: ~86 %) -> the dense float64 host array the reference API
    returns, without sending the zeros over PCIe:

        job = SparseRowsDownload(n_rows, width, threads)
        job.push(block)    # consecutive blocks of rows: compacted on the device (csrc/transport.cu), the non-zeros
                           # copied (~6 bytes each), then the library's host threads START rebuilding the block's
                           # dense float64 rows (csrc/host_io.cpp) and push returns: the next block is computed
                           # and copied meanwhile
        arr = job.finish() # waits for the host threads
        job.abandon()      # error paths: waits, so that no buffer is released under the threads

    The values are the kernels' float32 values widened exactly, i.e. the array a dense float64 copy would deliver.
    """

    def __init__(self, n_rows: int, width: int, threads: int | None = None) -> None:
        require_cuda()
        assert 0 < width <= 4096
        self.n_rows, self.width = int(n_rows), int(width)
        self.threads = host_threads(threads)
        self.result, self.result_array = result_buffer((self.n_rows, self.width))
        self.filled = 0
        self.bytes_copied = 0
        self._keep: list[torch.Tensor] = []
        self._pending = False

    def push(self, rows: torch.Tensor) -> None:
        import ctypes

        assert rows.is_cuda and rows.dtype == torch.float32 and rows.is_contiguous()
        n = int(rows.shape[0])
        assert rows.dim() == 2 and rows.shape[1] == self.width and self.filled + n <= self.n_rows
        if n == 0:
            return
        offsets = torch.empty(n + 1, dtype=torch.int64, device=rows.device)
        total = ctypes.c_int64()
        check(lib.sf_rows_compact_count(ptr(rows), n, self.width, ptr(offsets), ctypes.byref(total), stream_ptr()))
        nnz = max(int(total.value), 1)
        cols = torch.empty(nnz, dtype=torch.int16, device=rows.device)  # uint16 bit patterns
        vals = torch.empty(nnz, dtype=torch.float32, device=rows.device)
        check(lib.sf_rows_compact_fill(ptr(rows), n, self.width, ptr(offsets), ptr(cols), ptr(vals), stream_ptr()))
        h_offsets = torch.empty(n + 1, dtype=torch.int64, pin_memory=True)
        h_cols = torch.empty(nnz, dtype=torch.int16, pin_memory=True)
        h_vals = torch.empty(nnz, dtype=torch.float32, pin_memory=True)
        h_offsets.copy_(offsets, non_blocking=True)
        h_cols.copy_(cols, non_blocking=True)
        h_vals.copy_(vals, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self.bytes_copied += 8 * (n + 1) + 6 * int(total.value)
        self._keep += [h_offsets, h_cols, h_vals]  # read by the host threads until finish()
        self._pending = True
        check(lib.sf_host_expand_rows_begin(h_offsets.data_ptr(), h_cols.data_ptr(), h_vals.data_ptr(), n, self.width,
                                            self.result.ctypes.data + 8 * self.filled * self.width, self.threads))
        self.filled += n

    def finish(self) -> np.ndarray:
        assert self.filled == self.n_rows, "SparseRowsDownload.finish before every row was pushed"
        self._pending = False
        check(lib.sf_host_wait())
        self._keep = []
        return self.result_array

    def abandon(self) -> None:
        if self._pending:
            self._pending = False
            lib.sf_host_wait()
        self._keep = []

    def __del__(self):  # noqa: D105
        try:
            self.abandon()
        except Exception:  # noqa: BLE001
            pass


def download_sparse_rows(rows: torch.Tensor, threads: int | None = None) -> np.ndarray:
    """float32 device rows, mostly zeros -> dense float64 host array (see SparseRowsDownload)."""
    if rows.shape[0] == 0:
        return np.zeros(tuple(rows.shape), dtype=np.float64)
    job = SparseRowsDownload(rows.shape[0], rows.shape[1], threads)
    try:
        job.push(rows)
        return job.finish()
    finally:
        job.abandon()


# ---- device-resident hand-off between stages ----------------------------------------------------------------------
# The reference API returns host float64 arrays and takes them back a moment later (pipeline.py:154-174 computes the
# descriptors, :376-399 hands them to the matcher). The float32 device rows a descriptor call produced are remembered
# under the host array it returned (weak reference: dropped with the array); a later call that receives THAT array
# takes the device rows instead of sending 8 bytes per element back over PCIe. The array is the caller's and may
# have been modified since: a sample of its rows is compared with the device rows before they are trusted.
_HANDOFF: dict[int, tuple[object, torch.Tensor]] = {}
_HANDOFF_BUDGET_FRACTION = 0.25  # of the device's memory
_HANDOFF_SAMPLE_ROWS = 96


def remember_device_rows(host_array: np.ndarray, rows: torch.Tensor) -> None:
    """`rows` (float32, device) are the values of `host_array` (float64, host) — see above."""
    import weakref

    if not (isinstance(host_array, np.ndarray) and rows.is_cuda and tuple(rows.shape) == host_array.shape):
        return
    budget = _HANDOFF_BUDGET_FRACTION * torch.cuda.get_device_properties(rows.device).total_memory
    held = sum(t.numel() * t.element_size() for _, t in _HANDOFF.values())
    for key in list(_HANDOFF):  # oldest first
        if held + rows.numel() * rows.element_size() <= budget:
            break
        held -= _HANDOFF[key][1].numel() * _HANDOFF[key][1].element_size()
        del _HANDOFF[key]
    key = id(host_array)
    try:
        ref = weakref.ref(host_array, lambda _r, k=key: _HANDOFF.pop(k, None))
    except TypeError:  # pragma: no cover
        return
    _HANDOFF[key] = (ref, rows)


def device_rows_of(host_array) -> torch.Tensor | None:
    """The device rows remembered for exactly this array object, if a sample of its rows still equals them."""
    if not isinstance(host_array, np.ndarray):
        return None
    entry = _HANDOFF.get(id(host_array))
    if entry is None or entry[0]() is not host_array:
        return None
    rows = entry[1]
    if tuple(rows.shape) != host_array.shape or rows.device.index != torch.cuda.current_device():
        return None
    n = host_array.shape[0]
    if n:
        pick = np.unique(np.linspace(0, n - 1, min(n, _HANDOFF_SAMPLE_ROWS)).astype(np.int64))
        on_device = rows[torch.from_numpy(pick).to(rows.device)].double().cpu().numpy()
        if not np.array_equal(on_device, host_array[pick]):
            _HANDOFF.pop(id(host_array), None)
            return None
    return rows


def grid_geometry(lo, hi, radius: float) -> dict:
    """Cell edge and table dimensions the grid derives from a bounding box and a radius (sf_grid_geometry)."""
    import ctypes

    cell, ncells = ctypes.c_double(), ctypes.c_int64()
    dims = (ctypes.c_int32 * 3)()
    lo_c = (ctypes.c_double * 3)(*[float(v) for v in lo])
    hi_c = (ctypes.c_double * 3)(*[float(v) for v in hi])
    check(lib.sf_grid_geometry(lo_c, hi_c, float(radius), ctypes.byref(cell), dims, ctypes.byref(ncells)))
    return {"cell": cell.value, "dims": tuple(dims), "ncells": ncells.value}


class Grid:
    """Owner of one `sf_grid` handle (uniform grid over a cloud, see csrc/grid.cu)."""

    def __init__(self) -> None:
        import ctypes

        require_cuda()
        self._h = ctypes.c_void_p()
        check(lib.sf_grid_create(ctypes.byref(self._h)))
        self.n = 0
        self.radius = 0.0
        self.has_normals = False
        self._keep = ()

    def build(self, xyz: torch.Tensor, normals: torch.Tensor | None, radius: float, box=None) -> "Grid":
        """`box` = (lo, hi), three floats each: build in that bounding box instead of the cloud's own (the cells of a
        cloud that CONTAINS these points, see sf_grid_build_in_box); `poll()` then tells whether a point lay outside."""
        import ctypes

        assert xyz.dtype == torch.float64 and xyz.dim() == 2 and xyz.shape[1] == 3
        if normals is not None:
            assert normals.dtype == torch.float64 and normals.shape == xyz.shape
        if box is None:
            check(lib.sf_grid_build(self._h, ptr(xyz), ptr(normals), xyz.shape[0], float(radius), stream_ptr()))
        else:
            lo = (ctypes.c_double * 3)(*[float(v) for v in box[0]])
            hi = (ctypes.c_double * 3)(*[float(v) for v in box[1]])
            check(lib.sf_grid_build_in_box(self._h, ptr(xyz), ptr(normals), xyz.shape[0], float(radius), lo, hi, stream_ptr()))
        self.n, self.radius, self.has_normals = int(xyz.shape[0]), float(radius), normals is not None
        return self

    @property
    def handle(self):
        return self._h

    def set_speculative(self, builds: bool = False, shot_lists: bool = True) -> "Grid":
        """Repeated builds of the same cloud / fused SHOT calls on this handle skip their host synchronisations (see
        sf_grid_set_speculative in include/shotfpfh_b200.h); `poll()` after synchronising tells whether what they
        assumed held."""
        check(lib.sf_grid_set_speculative(self._h, int(bool(builds)) | (int(bool(shot_lists)) << 1)))
        return self

    def poll(self) -> int:
        """0, or the reason the speculative calls since the last poll did nothing (then repeat them)."""
        import ctypes

        status = ctypes.c_int32(0)
        check(lib.sf_grid_poll(self._h, ctypes.byref(status)))
        return int(status.value)

    def info(self) -> dict:
        import ctypes

        n, ncells, cell = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_double()
        dims = (ctypes.c_int32 * 3)()
        check(lib.sf_grid_info(self._h, ctypes.byref(n), ctypes.byref(ncells), ctypes.byref(cell), dims))
        return {"n": n.value, "ncells": ncells.value, "cell": cell.value, "dims": tuple(dims)}

    def close(self) -> None:
        if self._h:
            lib.sf_grid_destroy(self._h)
            self._h = None

    def __del__(self):  # noqa: D105
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
