"""
TEST INFRASTRUCTURE ONLY — the CPU oracle for the hot path (neighbour search -> SHOT / FPFH -> matching).

This package restates, in NumPy, what the reference `aubin-tchoi/shot-fpfh` computes on the hot path, each
function citing the reference file:line it follows. It exists to CHECK the CUDA path; it is never the
thing shipped or measured. Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it. The product package `shot_fpfh_b200` never imports it and has no CPU
fallback: it raises if its CUDA library is missing.

How the oracle is pinned (the reference ships no tests, no golden vectors and no fixtures — SURVEY.md §4):
  * `oracle/make_golden.py` imports the UNMODIFIED reference from /root/reference (this container only,
    with empty stub modules for matplotlib / coloredlogs), runs it on seeded synthetic inputs, checks every
    function below against it, and commits the reference's outputs as `tests/golden/*.npz`.
  * `tests/test_oracle_golden.py` (CPU, `-m "not gpu"`) re-checks the oracle against those fixtures anywhere.
Two paths are NOT pinned by the unmodified reference because the reference raises on them:
  * FPFH `decorrelated=True` (33-d): pinned against the reference with the one-token fix at fpfh.py:78
    (`).T` -> `).ravel()`), applied in memory by make_golden.py — "parity pinned to a patched reference".
  * `double_matching_with_rejects` (ratio test): "parity unpinned by the reference; pinned to the documented
    restatement" in `matching_oracle.ratio_matching`.

Third-party natives the reference calls and that are not vendored in it: scikit-learn `KDTree.query_radius`
(pinned 1.5.1 in poetry.lock:916-917; 1.9.0 here) and scipy `cdist` (pinned 1.14.0, poetry.lock:962-963;
1.18.1 here). The oracle calls the same libraries, and `neighbors_oracle.brute_force_radius` restates the
published fp64 predicate (`sum((a-b)^2) <= r^2`, sequential, inclusive) independently of the tree.
"""
