"""Seeded synthetic inputs in the reference's collate contract (NJODE/data_utils.py:278-316 and
latent_ODE/physionet_LODE.py:428-544 for the masked flavour) plus the model configurations used
across the golden fixtures and the parity tests.  Pure numpy/torch, no reference import."""
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

DEMO_NN = [[50, "tanh"], [50, "tanh"]]


def demo_cfg(**over):
    """demo.py / model_overview.csv:2 architecture (d=1, H=10, 2x50 tanh)."""
    cfg = dict(input_size=1, hidden_size=10, output_size=1, ode_nn=DEMO_NN, readout_nn=DEMO_NN,
               enc_nn=DEMO_NN, use_rnn=False, bias=True, dropout_rate=0.0, solver="euler",
               weight=0.5, weight_decay=1.0, options={})
    cfg.update(over)
    return cfg


CONFIGS = {
    "demo": demo_cfg(),
    "easy_w07_nores": demo_cfg(input_size=2, output_size=2, hidden_size=6, weight=0.7,
                               ode_nn=[[12, "tanh"], [9, "tanh"]], enc_nn=[[7, "tanh"]],
                               readout_nn=[[8, "tanh"], [5, "tanh"], [11, "tanh"]],
                               options={"which_loss": "easy", "residual_enc_dec": False}),
    "curt_nobias_relu": demo_cfg(input_size=3, output_size=3, hidden_size=9, bias=False,
                                 ode_nn=[[20, "relu"], [16, "tanh"]], enc_nn=None,
                                 readout_nn=[[14, "relu"]],
                                 options={"input_current_t": True}),
    "res_case2": demo_cfg(input_size=4, output_size=4, hidden_size=2,
                          ode_nn=[[10, "tanh"]], enc_nn=[[10, "tanh"]], readout_nn=None,
                          options={}),
    "masked_small": demo_cfg(input_size=5, output_size=5, hidden_size=10,
                             ode_nn=[[24, "tanh"], [24, "tanh"]], enc_nn=[[24, "tanh"], [24, "tanh"]],
                             readout_nn=[[24, "tanh"], [24, "tanh"]], options={"masked": True}),
    "masked_physio": demo_cfg(input_size=41, output_size=41, hidden_size=41, options={"masked": True}),
    # use_rnn=True: GRU jump (NJODE/models.py:202-217); never enabled in a shipped config, built for coverage
    "gru_demo": demo_cfg(use_rnn=True, hidden_size=10),
    "gru_d3_nores": demo_cfg(input_size=3, output_size=3, hidden_size=7, use_rnn=True, weight=0.6,
                             ode_nn=[[18, "tanh"]], enc_nn=[[12, "relu"]], readout_nn=[[16, "tanh"], [9, "tanh"]],
                             options={"residual_enc_dec": False, "which_loss": "easy"}),
    "gru_masked": demo_cfg(input_size=4, output_size=4, hidden_size=8, use_rnn=True,
                           ode_nn=[[20, "tanh"]], enc_nn=[[20, "tanh"]], readout_nn=[[20, "tanh"]],
                           options={"masked": True}),
    "wide_small": demo_cfg(input_size=16, output_size=16, hidden_size=64,
                           ode_nn=[[64, "tanh"]] * 4, enc_nn=[[64, "tanh"]] * 4,
                           readout_nn=[[64, "tanh"]] * 4),
}


def grid_batch(B, d, n_steps, obs_perc, seed, dt=None, S0=1.0, vol=0.3, drift=2.0):
    """Black-Scholes-like paths on a regular grid, Bernoulli(obs_perc) observation mask, collated
    exactly like custom_collate_fn (time-major rows, path index ascending inside a time, float64
    ``current_time += dt`` accumulation, times only where at least one path observes)."""
    rng = np.random.default_rng(seed)
    dt = 1.0 / n_steps if dt is None else dt
    paths = np.empty((B, d, n_steps + 1))
    paths[:, :, 0] = S0
    for k in range(1, n_steps + 1):
        z = rng.standard_normal((B, d))
        paths[:, :, k] = paths[:, :, k - 1] * (1 + drift * dt + vol * np.sqrt(dt) * z)
    observed = (rng.random((B, n_steps + 1)) < obs_perc) * 1
    return collate_arrays(paths, observed, dt)


def collate_arrays(paths, observed, dt):
    """vectorised restatement of NJODE/data_utils.py:285-315."""
    B, d, n1 = paths.shape
    times, time_ptr, obs_idx, X = [], [0], [], []
    current_time = 0.0
    for t in range(1, n1):
        current_time += dt
        idx = np.nonzero(observed[:, t] == 1)[0]
        if idx.size > 0:
            times.append(current_time)
            obs_idx.append(idx)
            X.append(paths[idx, :, t])
            time_ptr.append(time_ptr[-1] + idx.size)
    obs_idx = np.concatenate(obs_idx) if obs_idx else np.zeros(0, dtype=np.int64)
    X = np.concatenate(X, axis=0) if X else np.zeros((0, d))
    n_obs_ot = observed[:, 1:].sum(axis=1)
    return {"times": np.array(times), "time_ptr": np.array(time_ptr),
            "obs_idx": torch.tensor(obs_idx, dtype=torch.long),
            "start_X": torch.tensor(paths[:, :, 0], dtype=torch.float32),
            "n_obs_ot": torch.tensor(n_obs_ot),
            "X": torch.tensor(X, dtype=torch.float32).reshape(-1, d),
            "true_paths": paths, "observed_dates": observed}


def irregular_batch(B, d, n_times, seed, masked=False, times_f32=False, obs_at_zero=False,
                    empty_slot=True, zero_obs_path=True, t_max=0.97, row_prob=0.35, feat_prob=0.4):
    """irregular (off-grid) observation times; optional feature masks (PhysioNet flavour: float32
    ``times``, start_X = 0, rows = (time, path) pairs with any feature observed)."""
    rng = np.random.default_rng(seed)
    ts = np.sort(rng.random(n_times)) * t_max
    if obs_at_zero:
        ts[0] = 0.0
    ts = np.unique(ts.astype(np.float32) if times_f32 else ts)
    times = ts.astype(np.float32) if times_f32 else ts.astype(np.float64)
    time_ptr, obs_idx, X, M = [0], [], [], []
    for i in range(len(times)):
        rows = np.nonzero(rng.random(B) < row_prob)[0]
        if zero_obs_path:
            rows = rows[rows != B - 1]
        if empty_slot and i == len(times) // 2:
            rows = rows[:0]
        for p in rows:
            m = (rng.random(d) < feat_prob).astype(np.float32) if masked else np.ones(d, np.float32)
            if m.sum() == 0:
                m[rng.integers(d)] = 1.0
            x = rng.random(d).astype(np.float32) * (m if masked else 1.0) + (0.0 if masked else 0.5)
            obs_idx.append(p)
            X.append(x)
            M.append(m)
        time_ptr.append(len(obs_idx))
    obs_idx = np.array(obs_idx, dtype=np.int64)
    n_obs_ot = np.bincount(obs_idx, minlength=B).astype(np.int64)
    if masked:
        start_X = torch.zeros(B, d, dtype=torch.float32)
    else:
        start_X = torch.tensor(rng.random((B, d)) + 0.5, dtype=torch.float32)
    out = {"times": times, "time_ptr": np.array(time_ptr),
           "obs_idx": torch.tensor(obs_idx, dtype=torch.long),
           "start_X": start_X, "n_obs_ot": torch.tensor(n_obs_ot),
           "X": torch.tensor(np.array(X, dtype=np.float32).reshape(-1, d))}
    if masked:
        out["M"] = torch.tensor(np.array(M, dtype=np.float32).reshape(-1, d))
    return out


# ----------------------------------------------------------------------------------------------
# golden fixture I/O (npz: 'cfg' json, 'meta' json, 'sd/<name>', 'in/<name>', 'out/<name>')
# ----------------------------------------------------------------------------------------------
def save_case(path, cfg, meta, sd, batch, outs):
    arrs = {"cfg": np.array(json.dumps(cfg)), "meta": np.array(json.dumps(meta))}
    for k, v in sd.items():
        arrs["sd/" + k] = v.detach().numpy()
    for k, v in batch.items():
        if k in ("true_paths", "observed_dates"):
            continue
        arrs["in/" + k] = v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)
    for k, v in outs.items():
        arrs["out/" + k] = v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)
    np.savez_compressed(path, **arrs)


def load_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    cfg = json.loads(str(z["cfg"]))
    meta = json.loads(str(z["meta"]))
    sd = {k[3:]: torch.tensor(z[k]) for k in z.files if k.startswith("sd/")}
    batch = {}
    for k in z.files:
        if k.startswith("in/"):
            a = z[k]
            batch[k[3:]] = a if k[3:] in ("times", "time_ptr") else torch.tensor(a)
    outs = {k[4:]: z[k] for k in z.files if k.startswith("out/")}
    return cfg, meta, sd, batch, outs


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and not f.startswith(("sde_", "condexp_")))
