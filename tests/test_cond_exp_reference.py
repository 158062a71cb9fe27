"""a15: ``njode_b200.stock_model.{compute_cond_exp, get_optimal_loss}`` pinned to the REAL reference
(NJODE/stock_model.py:50-158, 426-468): against the committed outputs of the reference itself
(tests/golden/condexp_ref.npz, written by tests/golden/make_condexp_golden.py) and, when /root/reference is present
(build container), against the reference imported live on fresh seeded batches.  The GPU kernel ``njode_cond_exp`` is
checked against the same fixtures in tests/test_gpu_sde.py."""
import json
import os

import numpy as np
import pytest

import _reference
import cases
from njode_b200 import stock_model as sm

Z = np.load(os.path.join(cases.GOLDEN_DIR, "condexp_ref.npz"))
META = json.loads(str(Z["meta"]))
RTOL = 1e-12          # float64 closed form vs the reference's float64 stepping
# the reference keeps y in the dtype of start_X (float32 from the collate) until the first Euler step promotes it, so the
# loss terms of observations that precede the first step are evaluated in float32 there and in float64 here
RTOL_LOSS = 1e-8


def _model(meta):
    if meta["model"] == "combined":
        return sm.STOCK_MODELS["combined"](stock_model_names=meta["names"], hyperparam_dicts=meta["hps"])
    return sm.STOCK_MODELS[meta["model"]](**meta["hp"])


def _args(name):
    m = META[name]
    g = lambda k: Z["%s/%s" % (name, k)]
    return (g("times"), g("time_ptr"), g("X"), g("obs_idx"), m["delta_t"], m["T"], g("start_X"), g("n_obs_ot"))


@pytest.mark.parametrize("name", sorted(META))
def test_compute_cond_exp_matches_the_reference_fixture(name):
    meta = META[name]
    loss, path_t, path_y = _model(meta).compute_cond_exp(*_args(name), return_path=True, get_loss=True, weight=0.5,
                                                         **meta["kwargs"])
    assert np.array_equal(np.asarray(path_t, dtype=np.float64), Z[name + "/path_t"])           # the event list: exact
    assert path_y.shape == Z[name + "/path_y"].shape
    np.testing.assert_allclose(path_y, Z[name + "/path_y"], rtol=RTOL, atol=0)
    np.testing.assert_allclose(loss, Z[name + "/loss"], rtol=RTOL_LOSS)
    if name + "/optimal_loss_w07" in Z.files:
        got = _model(meta).get_optimal_loss(*_args(name), weight=0.7)
        np.testing.assert_allclose(got, Z[name + "/optimal_loss_w07"], rtol=RTOL)
    # loss only / path only calls return what the reference returns
    only = _model(meta).compute_cond_exp(*_args(name), return_path=False, get_loss=True, **meta["kwargs"])
    np.testing.assert_allclose(only, Z[name + "/loss"], rtol=RTOL_LOSS)
    l0, _, _ = _model(meta).compute_cond_exp(*_args(name), return_path=True, get_loss=False, **meta["kwargs"])
    assert l0 == 0


def test_observed_values_are_reproduced_exactly():
    """after a jump the observed paths carry X_obs itself (stock_model.py:119-121), not a rounded recomputation"""
    name = "ou_sine"
    _, path_t, path_y = _model(META[name]).compute_cond_exp(*_args(name))
    times, time_ptr, X, obs_idx = _args(name)[:4]
    for i, t in enumerate(times):
        e = int(np.nonzero(path_t == t)[0][-1])                   # the record after the jump (duplicate time stamp)
        rows = slice(time_ptr[i], time_ptr[i + 1])
        assert np.array_equal(path_y[e, obs_idx[rows]], X[rows].astype(np.float64))


REF = _reference.load_reference()


@pytest.mark.skipif(REF is None, reason="needs /root/reference (build container)")
@pytest.mark.parametrize("model,extra,d", [("BlackScholes", {}, 3), ("OrnsteinUhlenbeck", {"sine_coeff": 1.5}, 1),
                                           ("Heston", {"sine_coeff": 2.5}, 2), ("HestonWOFeller", {"v0": 0.3}, 1)])
@pytest.mark.parametrize("seed", [0, 1])
def test_live_reference_on_fresh_batches(model, extra, d, seed):
    hp = dict(drift=1.3, volatility=0.3, mean=2.5, speed=1.7, correlation=0.5, nb_paths=1, nb_steps=1, S0=1.,
              maturity=1., dimension=d, **extra)
    b = cases.grid_batch(17, d, 25, 0.3, seed=90 + seed)
    T = float(b["times"][-1])
    args = (b["times"], b["time_ptr"], b["X"].numpy(), b["obs_idx"].numpy(), 1.0 / 25, T, b["start_X"].numpy(),
            b["n_obs_ot"].numpy())
    ours, theirs = sm.STOCK_MODELS[model](**hp), REF.stock_model.STOCK_MODELS[model](**hp)
    la, ta, ya = ours.compute_cond_exp(*args, return_path=True, get_loss=True, weight=0.3)
    lr, tr, yr = theirs.compute_cond_exp(*args, return_path=True, get_loss=True, weight=0.3)
    assert np.array_equal(ta, tr)
    np.testing.assert_allclose(ya, yr, rtol=RTOL)
    np.testing.assert_allclose(la, lr, rtol=RTOL)
    np.testing.assert_allclose(ours.get_optimal_loss(*args), theirs.get_optimal_loss(*args), rtol=RTOL)
    y0 = np.abs(b["start_X"].numpy()) + 0.5
    np.testing.assert_allclose(ours.next_cond_exp(y0, 0.04, 0.3), theirs.next_cond_exp(y0, 0.04, 0.3), rtol=1e-14)


@pytest.mark.skipif(REF is None, reason="needs /root/reference (build container)")
def test_live_reference_combined_model():
    hps = [dict(drift=2., volatility=0.3, mean=10, speed=2., correlation=0.5, nb_paths=1, nb_steps=10, S0=1., maturity=0.5,
                dimension=2), dict(drift=2., volatility=0.3, mean=4, speed=2., correlation=0.5, nb_paths=1, nb_steps=10,
                                   S0=1., maturity=0.5, dimension=2)]
    for names in (["OrnsteinUhlenbeck", "BlackScholes"], ["BlackScholes", "OrnsteinUhlenbeck"]):
        b = cases.grid_batch(11, 2, 20, 0.3, seed=5)
        args = (b["times"], b["time_ptr"], b["X"].numpy(), b["obs_idx"].numpy(), 0.05, 1.0, b["start_X"].numpy(),
                b["n_obs_ot"].numpy())
        ours = sm.STOCK_MODELS["combined"](stock_model_names=names, hyperparam_dicts=hps)
        theirs = REF.stock_model.STOCK_MODELS["combined"](stock_model_names=names, hyperparam_dicts=hps)
        la, ta, ya = ours.compute_cond_exp(*args, return_path=True, get_loss=True)
        lr, tr, yr = theirs.compute_cond_exp(*args, return_path=True, get_loss=True)
        assert np.array_equal(ta, tr)
        np.testing.assert_allclose(ya, yr, rtol=RTOL)
        np.testing.assert_allclose(la, lr, rtol=RTOL)
        np.testing.assert_allclose(ours.get_optimal_loss(*args), theirs.get_optimal_loss(*args), rtol=RTOL)
