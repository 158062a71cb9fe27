"""njode_b200.eval_utils against a literal restatement of the reference's loops."""
import numpy as np

from njode_b200 import eval_utils


def ref_get_comparison_times_ind(path_t, times_val):
    """NJODE/physionet_train.py:478-510, line by line"""
    indices = []
    for t in times_val:
        for i in range(len(path_t) - 1):
            if abs(path_t[i] - t) < 1e-10 or path_t[i] <= t < path_t[i + 1]:
                if abs(t - path_t[i]) <= abs(t - path_t[i + 1]):
                    indices.append(i)
                else:
                    indices.append(i + 1)
                break
            elif i == len(path_t) - 2:
                indices.append(i + 1)
                break
    return indices


def test_comparison_times_indices_match_the_reference_loop():
    rng = np.random.default_rng(0)
    for trial in range(30):
        # record times of a return_path call: a grid plus duplicate stamps at the observation times
        grid = np.cumsum(np.full(60, 1.0 / 60))
        obs = np.sort(rng.choice(grid, size=rng.integers(1, 12), replace=False))
        if trial % 3 == 0:
            obs = np.concatenate((obs, rng.random(3) * 0.9 + 0.05))        # off-grid observation times
        path_t = np.sort(np.concatenate(([0.0], grid, obs, obs)))
        tv = np.concatenate((rng.random(25) * 0.98 + 0.01, obs[:3], grid[rng.integers(0, 60, 5)],
                             grid[rng.integers(0, 59, 3)] + 0.5 / 60, [path_t[-1]]))
        assert eval_utils.get_comparison_times_ind(path_t, tv) == ref_get_comparison_times_ind(path_t, tv)
