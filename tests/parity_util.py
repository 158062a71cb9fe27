"""shared checks: run njode_b200.models.NJODE (CUDA kernels, or their host simulation when a test
runner is injected) on a golden case and compare with the reference's committed outputs."""
import numpy as np
import torch

import cases
import oracle.njode_oracle as orc
from njode_b200 import models

RTOL = 1e-4          # BASELINE.json north_star: loss, predictions, parameter gradients (fp32)


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def build_model(cfg, sd, device, tensor_cores="off"):
    """``tensor_cores="off"``: these are the fp32 parity checks (rtol 1e-4); the bf16 tensor-core path has
    its own tests and tolerance (tests/test_gpu_wide.py)."""
    m = models.NJODE(**cfg)
    m.load_state_dict(sd)
    m.tensor_cores = tensor_cores
    return m.to(device)


def call(m, batch, meta, device, **kw):
    to = lambda t: t.to(device)
    return m(batch["times"], batch["time_ptr"], to(batch["X"]), batch["obs_idx"], meta["delta_t"],
             meta["T"], to(batch["start_X"]), batch["n_obs_ot"],
             M=to(batch["M"]) if "M" in batch else None, **kw)


def check_training_call(name, device, with_hT_grad=False):
    cfg, meta, sd, batch, outs = cases.load_case(name)
    m = build_model(cfg, sd, device)
    m.eval()
    hT, loss = call(m, batch, meta, device)
    assert loss.device.type == "cpu" and loss.dim() == 0
    assert rel_err(loss.detach().numpy(), outs["loss"]) < RTOL
    assert rel_err(hT.detach().cpu().numpy(), outs["hT"]) < RTOL
    key = "grad/"
    obj = loss
    if with_hT_grad:
        obj = loss + (hT * torch.tensor(outs["G"]).to(hT.device)).sum().cpu()
        key = "gradG/"
    obj.backward()
    for n, p in m.named_parameters():
        assert p.grad is not None, n
        assert rel_err(p.grad.cpu().numpy(), outs[key + n]) < RTOL, n


def check_path_call(name, device):
    cfg, meta, sd, batch, outs = cases.load_case(name)
    m = build_model(cfg, sd, device)
    m.eval()
    with torch.no_grad():
        hT, loss, path_t, path_h, path_y = call(m, batch, meta, device, return_path=True,
                                                get_loss=True, until_T=True)
    assert np.array_equal(np.asarray(path_t, dtype=np.float64), outs["path_t"])     # exact
    assert path_h.shape == outs["path_h"].shape and path_y.shape == outs["path_y"].shape
    assert rel_err(loss.numpy(), outs["loss_T"]) < RTOL
    assert rel_err(hT.cpu().numpy(), outs["hT_T"]) < RTOL
    assert rel_err(path_h.numpy(), outs["path_h"]) < RTOL
    assert rel_err(path_y.numpy(), outs["path_y"]) < RTOL


def check_against_oracle(cfg, batch, dt, T, seed, device, train=False, rtol=RTOL, grad_hT=False):
    """fresh seeded inputs: product vs the oracle run live (fp32).  train=True: dropout on, the
    oracle replays the device's counter-based keep-masks from the same seed."""
    ocfg = orc.Config(**cfg)
    sd = orc.init_state_dict(ocfg, seed=seed)
    m = build_model(cfg, sd, device)
    drop_seed = None
    if train:
        m.train()
        torch.manual_seed(1234)
        drop_seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        torch.manual_seed(1234)      # NJODE.forward draws the same seed
    else:
        m.eval()
    hT, loss = call(m, batch, {"delta_t": dt, "T": T}, device)
    G = None
    obj = loss
    if grad_hT:
        G = torch.randn(hT.shape, generator=torch.Generator().manual_seed(seed)) * 0.05
        obj = loss + (hT * G.to(hT.device)).sum().cpu()
    obj.backward()
    o_hT, o_loss, o_g = orc.loss_and_grads(ocfg, sd, batch, dt, T, dropout_seed=drop_seed, grad_hT=G)
    assert rel_err(loss.detach().numpy(), o_loss.numpy()) < rtol
    assert rel_err(hT.detach().cpu().numpy(), o_hT.numpy()) < rtol
    for n, p in m.named_parameters():
        assert rel_err(p.grad.cpu().numpy(), o_g[n].numpy()) < rtol, n
