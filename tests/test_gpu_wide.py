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


def _check(cfg, batch, dt, seed, train=False, grads=True, grad_hT=False, until_T=False):
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
    hT, loss = parity_util.call(m, batch, {"delta_t": dt, "T": 1.0}, DEV, until_T=until_T)
    assert m.last_forward_path == "tcgen05"
    G = None
    if grads:
        obj = loss
        if grad_hT:
            G = torch.randn(hT.shape, generator=torch.Generator().manual_seed(seed)) * 0.05
            obj = loss + (hT * G.to(hT.device)).sum().cpu()
        obj.backward()
    o_hT, o_loss, o_g = orc.loss_and_grads(ocfg, sd, batch, dt, 1.0, dropout_seed=drop_seed, grad_hT=G, until_T=until_T)
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


def test_gradient_into_hT_and_until_T_tail():
    """a caller gradient into the returned hidden state flows through the tail units (until_T: every path is
    integrated to T after its last observation)"""
    _check(_cfg(8, 128, 128, 2), cases.grid_batch(140, 8, 10, 0.3, seed=61), 0.1, seed=8, grad_hT=True, until_T=True)


def test_no_observations_at_all():
    """empty times / obs_idx: only tail units, loss 0, hT = Euler integration of enc(start_X) to T"""
    cfg = _cfg(4, 128, 128, 2)
    ocfg = orc.Config(**cfg)
    sd = orc.init_state_dict(ocfg, seed=9)
    m = parity_util.build_model(cfg, sd, DEV, tensor_cores="on").eval()
    B = 37
    batch = {"times": np.zeros(0), "time_ptr": np.array([0]), "X": torch.zeros(0, 4),
             "obs_idx": torch.zeros(0, dtype=torch.long), "start_X": torch.rand(B, 4) + 0.5,
             "n_obs_ot": torch.zeros(B, dtype=torch.long)}
    with torch.no_grad():
        hT, loss = parity_util.call(m, batch, {"delta_t": 0.125, "T": 1.0}, DEV, until_T=True)
        o_hT, o_loss = orc.forward(ocfg, sd, batch["times"], batch["time_ptr"], batch["X"], batch["obs_idx"], 0.125, 1.0,
                                   batch["start_X"], batch["n_obs_ot"], until_T=True)
    assert m.last_forward_path == "tcgen05"
    assert float(loss) == 0.0
    assert parity_util.rel_err(hT.cpu().numpy(), o_hT.numpy()) < STATE_RTOL


def test_forward_without_loss():
    cfg = _cfg(2, 128, 128, 2)
    ocfg = orc.Config(**cfg)
    sd = orc.init_state_dict(ocfg, seed=10)
    m = parity_util.build_model(cfg, sd, DEV, tensor_cores="on").eval()
    batch = cases.grid_batch(50, 2, 10, 0.3, seed=62)
    with torch.no_grad():
        hT, loss = parity_util.call(m, batch, {"delta_t": 0.1, "T": 1.0}, DEV, get_loss=False)
        o_hT, _ = orc.forward(ocfg, sd, batch["times"], batch["time_ptr"], batch["X"], batch["obs_idx"], 0.1, 1.0,
                              batch["start_X"], batch["n_obs_ot"])
    assert loss == 0 and m.last_forward_path == "tcgen05"
    assert parity_util.rel_err(hT.cpu().numpy(), o_hT.numpy()) < STATE_RTOL


def test_fp32_backward_after_tensor_core_forward():
    """model.tensor_core_backward = "fp32": the FMA backward kernels re-read the h history the tcgen05 forward wrote"""
    cfg = _cfg(16, 256, 256, 2)
    ocfg = orc.Config(**cfg)
    sd = orc.init_state_dict(ocfg, seed=11)
    m = parity_util.build_model(cfg, sd, DEV, tensor_cores="on").eval()
    m.tensor_core_backward = "fp32"
    batch = cases.grid_batch(60, 16, 8, 0.3, seed=63)
    hT, loss = parity_util.call(m, batch, {"delta_t": 0.125, "T": 1.0}, DEV)
    loss.backward()
    o_hT, o_loss, o_g = orc.loss_and_grads(ocfg, sd, batch, 0.125, 1.0)
    assert abs(float(loss) - float(o_loss)) < LOSS_RTOL * abs(float(o_loss))
    for n, p in m.named_parameters():
        assert parity_util.rel_err(p.grad.cpu().numpy(), o_g[n].numpy()) < STATE_RTOL, n
