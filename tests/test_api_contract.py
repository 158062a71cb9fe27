"""The drop-in boundary of SURVEY.md 8b as the reference's drivers use it: constructor argument handling, forward's
return types, module state, ownership of the inputs and the error behaviour (same exception types as
NJODE/models.py).  CPU-only: the kernels run as their host simulation."""
import numpy as np
import pytest
import torch

import cases
import hostsim_util
from njode_b200 import models


@pytest.fixture(autouse=True)
def sim_runner():
    hostsim_util.install()
    yield
    hostsim_util.uninstall()


def call(m, b, dt=0.05, T=1.0, **kw):
    return m(b["times"], b["time_ptr"], b["X"], b["obs_idx"], dt, T, b["start_X"], b["n_obs_ot"], **kw)


def test_constructor_takes_the_params_dict_of_train_py():
    """train.py:290-299 passes epochs, batch_size, dataset, ... next to the model arguments; `options` is required"""
    cfg = dict(cases.demo_cfg(), epochs=3, batch_size=100, dataset="BlackScholes", dataset_id=7, learning_rate=1e-3,
               test_size=0.2, seed=398, optimal_eval_loss=0.5)
    m = models.NJODE(**cfg)
    assert m.epoch == 1 and m.weight == 0.5
    no_opt = {k: v for k, v in cases.demo_cfg().items() if k != "options"}
    with pytest.raises(KeyError):                     # NJODE/models.py:321: options['options']
        models.NJODE(**no_opt)
    with pytest.raises(AssertionError):               # NJODE/models.py:326
        models.NJODE(**cases.demo_cfg(options={"which_loss": "nope"}))
    with pytest.raises(ValueError):                   # NJODE/models.py:248,256: residual sizes must be multiples
        models.NJODE(**cases.demo_cfg(input_size=3, output_size=3, hidden_size=10))
    models.NJODE(**cases.demo_cfg(input_size=3, output_size=3, hidden_size=10, options={"residual_enc_dec": False}))


def test_state_dict_keys_and_weight_schedule():
    m = models.NJODE(**cases.demo_cfg(use_rnn=True, weight=0.9, weight_decay=0.5))
    keys = set(m.state_dict().keys())
    want = {"%s.%d.%s" % (n, i, p) for n in ("ode_f.f", "encoder_map.ffnn", "readout_map.ffnn") for i in (0, 3, 6)
            for p in ("weight", "bias")} | {"obs_c.gru_d.%s_%s" % (p, q) for p in ("weight", "bias") for q in ("ih", "hh")}
    assert keys == want
    assert abs(m.weight_decay_step() - 0.7) < 1e-12 and abs(m.weight - 0.7) < 1e-12      # 0.5 + (0.9 - 0.5) * 0.5


def test_forward_returns_and_input_ownership():
    b = cases.grid_batch(12, 1, 20, 0.3, seed=3)
    m = models.NJODE(**cases.demo_cfg())
    snap = {k: (v.clone() if torch.is_tensor(v) else np.array(v, copy=True)) for k, v in b.items() if k in
            ("times", "time_ptr", "X", "obs_idx", "start_X", "n_obs_ot")}
    hT, loss = call(m, b)
    assert hT.shape == (12, 10) and loss.dim() == 0 and loss.device.type == "cpu" and loss.requires_grad
    loss.backward()
    assert all(p.grad is not None for p in m.parameters())
    float(loss.detach().numpy())                                     # train.py:558
    hT2, zero = call(m, b, get_loss=False)
    assert zero == 0 and isinstance(zero, int)                       # NJODE/models.py:420,513: python int when no loss
    out = call(m, b, return_path=True, until_T=True)
    assert len(out) == 5 and isinstance(out[2], np.ndarray) and out[3].shape[0] == len(out[2]) == out[4].shape[0]
    pred = m.get_pred(b["times"], b["time_ptr"], b["X"], b["obs_idx"], 0.05, 1.0, b["start_X"])
    assert set(pred) == {"pred", "pred_t"} and pred["pred"].device.type == "cpu"
    pred["pred"].detach().numpy()                                    # train.py:726
    for k, v in snap.items():                                        # inputs are never mutated (models.py:463,481)
        assert (torch.equal(b[k], v) if torch.is_tensor(v) else np.array_equal(b[k], v)), k


def test_error_behaviour_matches_the_reference():
    b = cases.grid_batch(8, 1, 20, 0.3, seed=4)
    m = models.NJODE(**cases.demo_cfg())
    with pytest.raises(AssertionError):                              # NJODE/models.py:428
        m(b["times"][:-1], b["time_ptr"], b["X"], b["obs_idx"], 0.05, 1.0, b["start_X"], b["n_obs_ot"])
    bad = models.NJODE(**cases.demo_cfg(solver="rk4"))
    with pytest.raises(ValueError):                                  # NJODE/models.py:374
        call(bad, b)
    masked = models.NJODE(**cases.CONFIGS["masked_small"])
    bm = cases.irregular_batch(6, 5, 10, seed=5, masked=True, times_f32=True)
    with pytest.raises(AssertionError):                              # NJODE/models.py:263: masked model needs M
        call(masked, bm)
    call(masked, bm, M=bm["M"])
    with pytest.raises(ValueError):
        m(b["times"], b["time_ptr"], b["X"], b["obs_idx"], 0.05, 1.0, b["start_X"], None)     # loss without n_obs_ot
    dup = dict(b)
    oi = b["obs_idx"].clone()
    s, e = int(b["time_ptr"][0]), int(b["time_ptr"][1])
    if e - s >= 2:
        oi[s + 1] = oi[s]                                            # the same path twice at one observation time
        dup["obs_idx"] = oi
        with pytest.raises(ValueError):
            call(m, dup)
    oob = dict(b)
    oi = b["obs_idx"].clone()
    oi[0] = 99
    oob["obs_idx"] = oi
    with pytest.raises(IndexError):
        call(m, oob)


def test_checkpoint_round_trip(tmp_path):
    m = models.NJODE(**cases.demo_cfg())
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    b = cases.grid_batch(8, 1, 20, 0.3, seed=6)
    hT, loss = call(m, b)
    loss.backward()
    opt.step()
    m.epoch, m.weight = 5, 0.61
    models.save_checkpoint(m, opt, str(tmp_path) + "/", m.epoch)
    m2 = models.NJODE(**cases.demo_cfg())
    opt2 = torch.optim.Adam(m2.parameters(), lr=1e-3)
    models.get_ckpt_model(str(tmp_path) + "/", m2, opt2, "cpu")
    assert m2.epoch == 5 and m2.weight == 0.61
    h1, l1 = call(m, b)
    h2, l2 = call(m2, b)
    assert float(l1) == float(l2) and torch.equal(h1, h2)
    with pytest.raises(Exception):
        models.get_ckpt_model(str(tmp_path) + "/missing/", m2, opt2, "cpu")
