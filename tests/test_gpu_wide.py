"""GPU tests of the tcgen05 tensor-core path (njode_wide_forward: bf16 operands, fp32 accumulation and
fp32 hidden state) against the oracle run live in fp32 on the same seeded inputs.

Stated bf16 tolerance (BASELINE.json north_star: "... or a stated bf16 tolerance"): operands are rounded to
bf16 (8 mantissa bits, relative 2^-9) before every contraction and tanh is MUFU.TANH (abs. error 2^-11), so
    loss                      |rel. error| < 5e-3
    hT, parameter gradients   max |a - b| / max |b| < 2e-2
Index handling (units, rows, Euler schedule) is shared with the fp32 path and stays exact."""
import numpy as np
import pytest
import torch

import cases
import parity_util
import oracle.njode_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
LOSS_RTOL, STATE_RTOL = 5e-3, 2e-2


def _cfg(d, H, W, L, **over):
    nn = [[W, "tanh"]] * L
    return cases.demo_cfg(input_size=d, output_size=d, hidden_size=H, ode_nn=nn, enc_nn=nn, readout_nn=nn, **over)


def _check(cfg, batch, dt, seed, train=False, grads=True):
    ocfg = orc.Config(**cfg)
    sd = orc.init_state_dict(ocfg, seed=seed)
    m = parity_util.build_model(cfg, sd, DEV, tensor_cores="on")
    drop_seed = None
    if train:
        m.train()
        torch.manual_seed(77)
        drop_seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        torch.manual_seed(77)
    else:
        m.eval()
    hT, loss = parity_util.call(m, batch, {"delta_t": dt, "T": 1.0}, DEV)
    assert m.last_forward_path == "tcgen05"
    if grads:
        loss.backward()
    o_hT, o_loss, o_g = orc.loss_and_grads(ocfg, sd, batch, dt, 1.0, dropout_seed=drop_seed)
    assert abs(float(loss) - float(o_loss)) < LOSS_RTOL * abs(float(o_loss))
    assert parity_util.rel_err(hT.detach().cpu().numpy(), o_hT.numpy()) < STATE_RTOL
    if grads:
        for n, p in m.named_parameters():
            assert parity_util.rel_err(p.grad.cpu().numpy(), o_g[n].numpy()) < STATE_RTOL, n


def test_config5_architecture_eval():
    """BASELINE config 5 nets (d=16, H=256, 4x256 tanh) on a batch the oracle finishes in seconds"""
    _check(_cfg(16, 256, 256, 4), cases.grid_batch(300, 16, 20, 0.2, seed=31), 0.05, seed=3)


def test_config5_architecture_train_dropout():
    """dropout on: the oracle replays the device's counter-based keep masks"""
    _check(_cfg(16, 256, 256, 4, dropout_rate=0.1), cases.grid_batch(200, 16, 16, 0.25, seed=32), 1.0 / 16, seed=4, train=True)


@pytest.mark.parametrize("d,H,W,L", [(16, 64, 64, 1), (4, 128, 192, 2), (1, 256, 128, 3), (8, 96, 80, 2)])
def test_other_shapes(d, H, W, L):
    """narrower layers (padded to the UMMA granularity), one to three hidden layers, every residual fold"""
    _check(_cfg(d, H, W, L), cases.grid_batch(150, d, 12, 0.3, seed=33 + d), 1.0 / 12, seed=5)


def test_options_easy_loss_current_t_relu_no_residual():
    nn = [[128, "relu"], [128, "tanh"]]
    cfg = cases.demo_cfg(input_size=2, output_size=2, hidden_size=128, ode_nn=nn, enc_nn=nn, readout_nn=nn, weight=0.7,
                         options={"which_loss": "easy", "input_current_t": True, "residual_enc_dec": False})
    _check(cfg, cases.grid_batch(130, 2, 10, 0.3, seed=41), 0.1, seed=6)


def test_irregular_times_zero_obs_paths_and_observation_at_t0():
    cfg = _cfg(4, 128, 128, 2)
    batch = cases.irregular_batch(37, 4, 9, seed=51, obs_at_zero=True)
    _check(cfg, batch, 0.07, seed=7)


def test_auto_mode_keeps_narrow_models_on_fp32():
    cfg = cases.CONFIGS["demo"]
    sd = orc.init_state_dict(orc.Config(**cfg), seed=1)
    m = parity_util.build_model(cfg, sd, DEV, tensor_cores="auto").eval()
    batch = cases.grid_batch(20, 1, 10, 0.3, seed=2)
    with torch.no_grad():
        parity_util.call(m, batch, {"delta_t": 0.1, "T": 1.0}, DEV)
    assert m.last_forward_path == "fp32"
