"""The reference's collate contract (NJODE/data_utils.py) on the host, vectorised.

``custom_collate_fn`` / ``CustomCollateFnGen`` accept the items the reference's
``IrregularDataset.__getitem__`` produces (NJODE/data_utils.py:269-275) and return the same dict
(keys, dtypes, row order: time ascending, then batch position ascending; NJODE/data_utils.py:311-315)
without the reference's Python double loop over (time step, path).
"""
import numpy as np
import torch

hyperparam_default = {                      # NJODE/data_utils.py:25-31
    'drift': 2., 'volatility': 0.3, 'mean': 4,
    'speed': 2., 'correlation': 0.5, 'nb_paths': 10000, 'nb_steps': 100,
    'S0': 1, 'maturity': 1., 'dimension': 1,
    'obs_perc': 0.1,
    'scheme': 'euler', 'return_vol': False, 'v0': 1,
}


def _get_func(name):
    """NJODE/data_utils.py:319-334"""
    if name in ['exp', 'exponential']:
        return np.exp
    if 'power-' in name:
        x = float(name.split('-')[1])
        return lambda inp: np.power(inp, x)
    return None


def _get_X_with_func_appl(X, functions, axis):
    """NJODE/data_utils.py:336-349"""
    Y = X
    for f in functions:
        Y = np.concatenate([Y, f(X)], axis=axis)
    return Y


def collate_paths(stock_paths, observed_dates, nb_obs, dt, functions=()):
    """stock_paths f64 [B, d, steps+1], observed_dates int [B, steps+1] -> contract dict.
    Restates the loop of NJODE/data_utils.py:292-315: ``current_time += dt`` accumulated in
    float64 once per grid step; a time enters ``times`` iff at least one path observes it."""
    B, d, n1 = stock_paths.shape
    obs = np.asarray(observed_dates)[:, 1:] == 1                  # column 0 is never an observation
    grid_t = np.cumsum(np.full(n1 - 1, dt, dtype=np.float64))     # same left-to-right additions
    any_obs = obs.any(axis=0)
    times = grid_t[any_obs]
    t_idx, p_idx = np.nonzero(obs.T)                              # time-major, path ascending
    counts = obs.sum(axis=0)[any_obs]
    time_ptr = np.concatenate(([0], np.cumsum(counts)))
    Xp = stock_paths[p_idx, :, t_idx + 1]
    start = stock_paths[:, :, 0]
    if functions:
        Xp = _get_X_with_func_appl(Xp, functions, axis=1)
        start = _get_X_with_func_appl(start, functions, axis=1)
    assert len(p_idx) == obs.sum()
    return {'times': times, 'time_ptr': time_ptr,
            'obs_idx': torch.from_numpy(p_idx.astype(np.int64)),
            'start_X': torch.tensor(start, dtype=torch.float32),
            'n_obs_ot': torch.as_tensor(np.asarray(nb_obs)),
            'X': torch.tensor(Xp, dtype=torch.float32).reshape(len(p_idx), -1),
            'true_paths': stock_paths, 'observed_dates': observed_dates}


def _cat(batch, key):
    return np.concatenate([b[key] for b in batch], axis=0)


def custom_collate_fn(batch):
    """NJODE/data_utils.py:278-316"""
    return collate_paths(_cat(batch, 'stock_path'), _cat(batch, 'observed_dates'),
                         _cat(batch, 'nb_obs'), batch[0]['dt'])


def CustomCollateFnGen(func_names=None):
    """NJODE/data_utils.py:352-416: returns (collate_fn, dimension multiplier)"""
    functions = []
    if func_names is not None:
        for func_name in func_names:
            f = _get_func(func_name)
            if f is not None:
                functions.append(f)
    mult = len(functions) + 1

    def collate(batch):
        return collate_paths(_cat(batch, 'stock_path'), _cat(batch, 'observed_dates'),
                             _cat(batch, 'nb_obs'), batch[0]['dt'], functions)
    return collate, mult
