"""Evaluation-side helpers of the reference's drivers that sit right behind ``NJODE.get_pred`` / ``forward(return_path=True)``
(SURVEY.md 8f rank 3).  Pure NumPy: O((n + m) log n) instead of the drivers' O(n m) Python loops."""
import numpy as np


def get_comparison_times_ind(path_t, times_val):
    """NJODE/physionet_train.py:478-510: for every entry of ``times_val`` the index of the entry of ``path_t`` (the
    nondecreasing record times of a return_path call, with the duplicate stamp after every jump) closest to it -- the
    first index within 1e-10 if there is one (i.e. the record BEFORE the jump at an observation time), else the nearer
    of the two neighbours, ties to the left one.  Same asserts, same result list as the reference's double loop."""
    path_t = np.asarray(path_t, dtype=np.float64)
    times_val = np.asarray(times_val, dtype=np.float64)
    assert np.min(path_t) < np.min(times_val) and np.max(path_t) + 1e-10 > np.max(times_val), \
        "mins: {}, {}, max: {}, {}".format(np.min(path_t), np.min(times_val), np.max(path_t), np.max(times_val))
    n = len(path_t)
    # first index whose stamp is within 1e-10 of t
    near = np.searchsorted(path_t, times_val - 1e-10, side="right")
    near_ok = (near < n - 1) & (np.abs(path_t[np.minimum(near, n - 1)] - times_val) < 1e-10)
    # otherwise: i = last index with path_t[i] <= t (at most n - 2), then the nearer of i and i + 1
    i = np.clip(np.searchsorted(path_t, times_val, side="right") - 1, 0, n - 2)
    inside = (path_t[i] <= times_val) & (times_val < path_t[i + 1])
    left = np.abs(times_val - path_t[i]) <= np.abs(times_val - path_t[i + 1])
    idx = np.where(inside, np.where(left, i, i + 1), n - 1)
    # a near-equal stamp is met by the loop before (or as the right end of) that interval; both routes agree on it
    idx = np.where(near_ok, near, idx)
    return [int(v) for v in idx]
