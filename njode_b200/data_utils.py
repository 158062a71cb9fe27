"""The reference's collate contract (NJODE/data_utils.py) on the host, vectorised.

``custom_collate_fn`` / ``CustomCollateFnGen`` accept the items the reference's
``IrregularDataset.__getitem__`` produces (NJODE/data_utils.py:269-275) and return the same dict
(keys, dtypes, row order: time ascending, then batch position ascending; NJODE/data_utils.py:311-315)
without the reference's Python double loop over (time step, path).
"""
import json
import os
import time

import numpy as np
import torch

from . import stock_model

hyperparam_default = {                      # NJODE/data_utils.py:25-31
    'drift': 2., 'volatility': 0.3, 'mean': 4,
    'speed': 2., 'correlation': 0.5, 'nb_paths': 10000, 'nb_steps': 100,
    'S0': 1, 'maturity': 1., 'dimension': 1,
    'obs_perc': 0.1,
    'scheme': 'euler', 'return_vol': False, 'v0': 1,
}


_STOCK_MODELS = stock_model.STOCK_MODELS

data_path = '../data/'                                            # NJODE/data_utils.py:37-38 (relative to the CWD)
training_data_path = '{}training_data/'.format(data_path)


# ------------------------------------------------------------------------------------------------
# datasets on disk: same layout as the reference (NJODE/data_utils.py:42-275), so datasets written by
# either implementation load in the other:
#   <training_data_path>/<name>-<time_id>/data.npy      three consecutive np.save records: stock_paths
#                                                       f64 [nb_paths, dim, steps+1], observed_dates int64
#                                                       [nb_paths, steps+1] (0/1), nb_obs int64 [nb_paths]
#   <training_data_path>/<name>-<time_id>/metadata.txt  JSON of the hyper-parameters (+ 'dt', 'model_name')
#   <training_data_path>/dataset_overview.csv           columns name, id, description
# The paths and the observation mask come from the CUDA generators (njode_sde_generate): same law as the
# reference's NumPy loop, Philox stream keyed by (seed, path id) instead of the global MT state.
# ------------------------------------------------------------------------------------------------
def makedirs(dirname):
    if not os.path.exists(dirname):
        os.makedirs(dirname)


def get_dataset_overview():
    """NJODE/data_utils.py:47-56 -> (DataFrame[name, id, description], csv path)"""
    import pandas as pd
    data_overview = '{}dataset_overview.csv'.format(training_data_path)
    makedirs(training_data_path)
    if os.path.exists(data_overview):
        df_overview = pd.read_csv(data_overview, index_col=0)
    else:
        df_overview = pd.DataFrame(data=None, columns=['name', 'id', 'description'])
    return df_overview, data_overview


def _register_and_write(name, desc, metadata, stock_paths, observed_dates, nb_obs):
    import pandas as pd
    df_overview, data_overview = get_dataset_overview()
    time_id = int(time.time())
    path = '{}{}-{}/'.format(training_data_path, name, time_id)
    if os.path.exists(path):
        print('Path already exists - abort')
        raise ValueError
    df_app = pd.DataFrame(data=[[name, time_id, desc]], columns=['name', 'id', 'description'])
    pd.concat([df_overview, df_app], ignore_index=True).to_csv(data_overview)
    os.makedirs(path)
    with open('{}data.npy'.format(path), 'wb') as f:
        np.save(f, np.asarray(stock_paths, dtype=np.float64))
        np.save(f, np.asarray(observed_dates, dtype=np.int64))
        np.save(f, np.asarray(nb_obs, dtype=np.int64))
    with open('{}metadata.txt'.format(path), 'w') as f:
        json.dump(metadata, f, sort_keys=True)
    return path, time_id


def create_dataset(stock_model_name="BlackScholes", hyperparam_dict=hyperparam_default, seed=0):
    """NJODE/data_utils.py:59-108: generates nb_paths paths and the Bernoulli(obs_perc) observation mask on the
    device and writes them in the reference's format -> (path, time_id).  Like the reference it adds
    'model_name' and 'dt' to ``hyperparam_dict``."""
    hyperparam_dict['model_name'] = stock_model_name
    obs_perc = hyperparam_dict['obs_perc']
    kw = {k: v for k, v in hyperparam_dict.items() if k != 'seed'}          # a 'seed' entry would collide with seed=
    stockmodel = _STOCK_MODELS[stock_model_name](**kw, seed=seed)
    paths, observed, nb_obs, dt = stockmodel.generate_paths_device(obs_perc=obs_perc)
    desc = json.dumps(hyperparam_dict, sort_keys=True)          # the overview row has no 'dt' (as in the reference)
    hyperparam_dict['dt'] = dt
    return _register_and_write(stock_model_name, desc, hyperparam_dict, paths.cpu().numpy(),
                               observed.cpu().numpy(), nb_obs.cpu().numpy())


def create_combined_dataset(stock_model_names=("BlackScholes", "OrnsteinUhlenbeck"),
                            hyperparam_dicts=(hyperparam_default, hyperparam_default), seed=0):
    """NJODE/data_utils.py:111-195: every further model continues the paths from the last value of the previous
    one; the observation mask is drawn once over the concatenated grid with the FIRST model's obs_perc."""
    assert len(stock_model_names) == len(hyperparam_dicts)
    filename = 'combined_{}'.format(stock_model_names[0])
    maturity = hyperparam_dicts[0]['maturity']
    hyperparam_dicts[0]['model_name'] = stock_model_names[0]
    obs_perc = hyperparam_dicts[0]['obs_perc']
    sm = _STOCK_MODELS[stock_model_names[0]](**hyperparam_dicts[0], seed=seed)
    paths, _, _, dt = sm.generate_paths_device()
    pieces = [paths]
    for i in range(1, len(stock_model_names)):
        assert hyperparam_dicts[i]['dimension'] == hyperparam_dicts[i - 1]['dimension']
        assert hyperparam_dicts[i]['nb_paths'] == hyperparam_dicts[i - 1]['nb_paths']
        filename += '_{}'.format(stock_model_names[i])
        maturity += hyperparam_dicts[i]['maturity']
        hyperparam_dicts[i]['model_name'] = stock_model_names[i]
        sm = _STOCK_MODELS[stock_model_names[i]](**hyperparam_dicts[i], seed=seed + 7919 * i)
        nxt, _, _, dt_i = sm.generate_paths_device(start_X=pieces[-1][:, :, -1].cpu().numpy())
        assert dt_i == dt
        pieces.append(nxt[:, :, 1:])
    stock_paths = torch.cat(pieces, dim=2)
    n_paths, _, n1 = stock_paths.shape
    # the mask of the concatenated grid: one Philox stream per path, disjoint from the path streams
    gen = torch.Generator(device=stock_paths.device)
    gen.manual_seed(int(seed) * 1000003 + 17)
    observed = (torch.rand(n_paths, n1, generator=gen, device=stock_paths.device, dtype=torch.float64) < obs_perc).to(torch.int64)
    nb_obs = observed[:, 1:].sum(dim=1)
    metadata = {'dt': dt, 'maturity': maturity, 'dimension': hyperparam_dicts[0]['dimension'],
                'nb_paths': hyperparam_dicts[0]['nb_paths'], 'model_name': 'combined',
                'stock_model_names': list(stock_model_names), 'hyperparam_dicts': list(hyperparam_dicts)}
    desc = json.dumps(metadata, sort_keys=True)
    return _register_and_write(filename, desc, metadata, stock_paths.cpu().numpy(), observed.cpu().numpy(),
                               nb_obs.cpu().numpy())


def _get_time_id(stock_model_name="BlackScholes", time_id=None):
    """NJODE/data_utils.py:198-216: newest dataset of that name when time_id is None"""
    if time_id is None:
        ids = [int(entry.split('-')[1]) for entry in os.listdir(training_data_path)
               if entry.split('-')[0] == stock_model_name]
        time_id = max(ids) if ids else None
    return time_id


def _dataset_dir(stock_model_name, time_id):
    time_id = _get_time_id(stock_model_name=stock_model_name, time_id=time_id)
    return '{}{}-{}/'.format(training_data_path, stock_model_name, int(time_id))


def load_metadata(stock_model_name="BlackScholes", time_id=None):
    """NJODE/data_utils.py:219-228"""
    with open('{}metadata.txt'.format(_dataset_dir(stock_model_name, time_id)), 'r') as f:
        return json.load(f)


def load_dataset(stock_model_name="BlackScholes", time_id=None):
    """NJODE/data_utils.py:231-249 -> (stock_paths, observed_dates, nb_obs, hyperparam_dict)"""
    path = _dataset_dir(stock_model_name, time_id)
    with open('{}data.npy'.format(path), 'rb') as f:
        stock_paths = np.load(f)
        observed_dates = np.load(f)
        nb_obs = np.load(f)
    with open('{}metadata.txt'.format(path), 'r') as f:
        hyperparam_dict = json.load(f)
    return stock_paths, observed_dates, nb_obs, hyperparam_dict


class IrregularDataset(torch.utils.data.Dataset):
    """NJODE/data_utils.py:252-275: items are dicts of [1, ...] slices that custom_collate_fn concatenates"""

    def __init__(self, model_name, time_id=None, idx=None):
        stock_paths, observed_dates, nb_obs, hyperparam_dict = load_dataset(stock_model_name=model_name, time_id=time_id)
        if idx is None:
            idx = np.arange(hyperparam_dict['nb_paths'])
        self.metadata = hyperparam_dict
        self.stock_paths = stock_paths[idx]
        self.observed_dates = observed_dates[idx]
        self.nb_obs = nb_obs[idx]

    def __len__(self):
        return len(self.nb_obs)

    def __getitem__(self, idx):
        if type(idx) == int:
            idx = [idx]
        return {"idx": idx, "stock_path": self.stock_paths[idx], "observed_dates": self.observed_dates[idx],
                "nb_obs": self.nb_obs[idx], "dt": self.metadata['dt']}


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
            'X': torch.tensor(Xp, dtype=torch.float32).reshape(len(p_idx), start.shape[1]),   # [0, d] when nothing is observed
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
