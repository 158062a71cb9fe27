"""Generates tests/golden/condexp_ref.npz by running the REAL reference's ``compute_cond_exp`` / ``get_optimal_loss``
(NJODE/stock_model.py:50-158, 426-468; imported from /root/reference, build container only) on seeded batches of the
collate contract: the four stock models (HestonWOFeller also with return_vol), the regime-switching ``Combined`` model,
a restart at ``start_time`` and two horizons that leave a tail after the last processed observation.

The reference's tail loop calls ``next_cond_exp`` without the current time (stock_model.py:139) and raises TypeError, so
the two tail cases run a subclass whose only change is a default for that argument (constant coefficient: the value is
not used).  Run:  python tests/golden/make_condexp_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _reference  # noqa: E402
import cases  # noqa: E402

ref = _reference.load_reference()
assert ref is not None, "needs /root/reference"
RSM = ref.stock_model

HP = dict(drift=2., volatility=0.3, mean=4, speed=2., correlation=0.5, nb_paths=24, nb_steps=40, S0=1.,
          maturity=1., dimension=1, v0=0.5)
STEPS, DT = 40, 1.0 / 40


def batch_for(d, seed, vol=False):
    b = cases.grid_batch(24, d, STEPS, 0.15, seed=seed)
    X, sx = b["X"].numpy(), b["start_X"].numpy()
    if vol:                                   # [spot, variance] coordinates (stock_model.py:329-330)
        rng = np.random.default_rng(seed + 100)
        X = np.concatenate([X, (0.2 + rng.random(X.shape)).astype(np.float32)], axis=1)
        sx = np.concatenate([sx, np.full_like(sx, 0.5)], axis=1)
    return dict(times=np.asarray(b["times"]), time_ptr=np.asarray(b["time_ptr"]), X=X, obs_idx=b["obs_idx"].numpy(),
                start_X=sx, n_obs_ot=b["n_obs_ot"].numpy())


def tail_safe(cls):
    class TailSafe(cls):
        def next_cond_exp(self, y, delta_t, current_t=0.0):
            return super().next_cond_exp(y, delta_t, current_t)
    return TailSafe


CASES = [
    # name, model, hyper-parameter overrides, data dim, T, kwargs, tail-safe subclass
    ("bs", "BlackScholes", {}, 1, None, {}, False),
    ("bs_d2_sine", "BlackScholes", {"sine_coeff": 3.0, "S0": [1., 1.], "dimension": 2}, 2, None, {}, False),
    ("ou_sine", "OrnsteinUhlenbeck", {"sine_coeff": 2.0}, 1, None, {}, False),
    ("heston", "Heston", {}, 1, None, {}, False),
    ("hwof", "HestonWOFeller", {}, 1, None, {}, False),
    ("hwof_vol", "HestonWOFeller", {"return_vol": True}, 2, None, {}, False),
    ("bs_start_time", "BlackScholes", {}, 1, None, {"start_time": 0.4}, False),
    ("ou_tail", "OrnsteinUhlenbeck", {}, 1, 0.615, {}, True),
    ("bs_tail_T2", "BlackScholes", {}, 1, 1.3, {}, True),
]
out = {}
meta = {}
for i, (name, model, over, d, T, kw, safe) in enumerate(CASES):
    hp = dict(HP, **over)
    b = batch_for(1 if (over.get("return_vol")) else d, seed=40 + i, vol=bool(over.get("return_vol")))
    T_ = float(b["times"][-1]) if T is None else T
    cls = RSM.STOCK_MODELS[model]
    m = (tail_safe(cls) if safe else cls)(**hp)
    args = (b["times"], b["time_ptr"], b["X"], b["obs_idx"], DT, T_, b["start_X"], b["n_obs_ot"])
    loss, path_t, path_y = m.compute_cond_exp(*args, return_path=True, get_loss=True, weight=0.5, **kw)
    meta[name] = dict(model=model, hp=hp, T=T_, delta_t=DT, kwargs=kw, d=int(b["X"].shape[1]))
    for k, v in b.items():
        out["%s/%s" % (name, k)] = v
    out[name + "/loss"], out[name + "/path_t"], out[name + "/path_y"] = np.float64(loss), path_t, path_y
    if not kw and not safe:
        out[name + "/optimal_loss_w07"] = np.float64(m.get_optimal_loss(*args, weight=0.7))

# the regime-switching model of parallel_train.py:588-605 (two regimes of maturity 0.5 on one 40-step grid)
hps = [dict(HP, maturity=0.5, nb_steps=20, mean=10, speed=2.), dict(HP, maturity=0.5, nb_steps=20)]
names = ["OrnsteinUhlenbeck", "BlackScholes"]
b = batch_for(1, seed=77)
comb = RSM.STOCK_MODELS["combined"](stock_model_names=names, hyperparam_dicts=hps)
args = (b["times"], b["time_ptr"], b["X"], b["obs_idx"], DT, 1.0, b["start_X"], b["n_obs_ot"])
loss, path_t, path_y = comb.compute_cond_exp(*args, return_path=True, get_loss=True)
meta["combined"] = dict(model="combined", names=names, hps=hps, T=1.0, delta_t=DT, kwargs={}, d=1)
for k, v in b.items():
    out["combined/%s" % k] = v
out["combined/loss"], out["combined/path_t"], out["combined/path_y"] = np.float64(loss), path_t, path_y
out["combined/optimal_loss_w07"] = np.float64(comb.get_optimal_loss(*args, weight=0.7))
out["meta"] = np.array(json.dumps(meta))
np.savez_compressed(os.path.join(HERE, "condexp_ref.npz"), **out)
print("wrote condexp_ref.npz:", {k: (float(out[k + "/loss"]), out[k + "/path_y"].shape) for k in meta})
