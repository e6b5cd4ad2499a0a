"""Shared parity checks for top-k results (used by the -m gpu tests).

Policy ("Exact top-k index sets vs fp32 rounding"): the GPU sums in a
different order than NumPy, so two rows whose true scores differ by less than the fp32 rounding
of the reduction may legitimately swap around a rank boundary.  A result row is accepted when
  (1) every returned score equals the fp64 truth of the returned index within `rtol`/`atol`
      (the tolerance north_star states: 1e-5 relative for fp32),
  (2) the returned scores are sorted best-first, ties by ascending index,
  (3) no excluded row is better than the worst returned row by more than the tolerance, and
  (4) where the returned index differs from the oracle's at a rank, the two rows' true scores
      are within the tolerance of each other (a "disputed" near-tie).
The number of disputed ranks is returned so tests can bound it; on integer-valued data the
tests demand exact equality instead.
"""
import numpy as np


def check_topk(scores, idx, truth, k, descending, rtol=1e-5, atol=2e-6, score_of_truth=None):
    scores = np.asarray(scores, dtype=np.float64)
    idx = np.asarray(idx)
    nq, N = truth.shape
    kk = min(k, N)
    sgn = -1.0 if descending else 1.0
    disputed = 0
    order = np.argsort(sgn * truth, axis=1, kind="stable")
    for r in range(nq):
        got_i = idx[r, :kk]
        assert np.all(got_i >= 0) and np.all(got_i < N), f"row {r}: index out of range {got_i}"
        assert len(set(got_i.tolist())) == kk, f"row {r}: duplicate indices"
        t = truth[r, got_i]
        tol = atol + rtol * np.abs(t)
        want_scores = t if score_of_truth is None else score_of_truth(t)
        assert np.all(np.abs(scores[r, :kk] - want_scores) <= atol + rtol * np.abs(want_scores)), \
            f"row {r}: score mismatch {scores[r, :kk]} vs {want_scores}"
        key = sgn * scores[r, :kk]
        assert np.all(np.diff(key) >= 0), f"row {r}: not sorted"
        ties = np.diff(key) == 0
        assert np.all(np.diff(got_i)[ties] > 0), f"row {r}: ties not in index order"
        worst = np.max(sgn * t)
        mask = np.ones(N, bool)
        mask[got_i] = False
        if mask.any():
            best_excluded = np.min(sgn * truth[r, mask])
            assert best_excluded >= worst - tol.max(), f"row {r}: missed a better row ({best_excluded} < {worst})"
        want_i = order[r, :kk]
        diff = got_i != want_i
        if diff.any():
            gap = np.abs(truth[r, got_i[diff]] - truth[r, want_i[diff]])
            assert np.all(gap <= 2 * tol[diff]), f"row {r}: wrong index beyond rounding: {got_i} vs {want_i}"
            disputed += int(diff.sum())
    # padding slots
    if k > kk:
        assert np.all(idx[:, kk:] == -1)
        assert np.all(np.isinf(scores[:, kk:]))
    return disputed
