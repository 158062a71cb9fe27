"""shared checks: run njode_b200.models.NJODE (CUDA kernels, or their host simulation when a test
runner is injected) on a golden case and compare with the reference's committed outputs."""
import os

import numpy as np
import torch

import cases
import oracle.njode_oracle as orc
from njode_b200 import models

RTOL = 1e-4          # BASELINE.json north_star: loss, predictions, parameter gradients (fp32)
NOISE_MULT = 8.0


def rel_err(a, b):
    """max-norm relative error (reporting only; the parity assertions are element-wise, see assert_close)"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def noise_floor(ref32, truth64, kern32=None):
    """the fp32 rounding noise of the REFERENCE computation itself on one tensor: the largest element-wise deviation of
    an fp32 evaluation from the same function evaluated in fp64 (oracle, dtype=float64).  An element whose magnitude is
    below that noise carries no information at rtol 1e-4 -- the reference would not reproduce it against itself -- so it
    is the absolute floor (atol) of the element-wise comparison, per tensor, measured, not chosen.  Two fp32 evaluations
    are measured: ``ref32`` (ATen on the CPU, libm tanh) and ``kern32`` = the same oracle with tanh evaluated by the
    kernels' formula 1 - 2 / (exp2(2 x log2 e) + 1) (absolute error ~1e-7, the accuracy class of CUDA's own tanhf, which
    uses that expression above |x| = 0.55): sums that cancel (the gradient of a trained model) see that noise directly.
    NOISE_MULT covers that a measured maximum is ONE draw of the noise, other summation orders, and that the emulation
    rounds exp2 correctly (<= 1 ulp on the CPU) where the SFU's ex2.approx.ftz has a relative error of up to 2^-22 (4 ulp):
    measured on the B200, the worst element of the golden cases sits at 5.6x the emulated floor (gru_masked,
    grad ode_f.f.0.weight, |diff| 2.4e-7 = 5.5e-6 of the tensor's max)."""
    t = np.asarray(truth64, dtype=np.float64)
    f = np.abs(np.asarray(ref32, dtype=np.float64) - t).max()
    if kern32 is not None:
        f = max(f, np.abs(np.asarray(kern32, dtype=np.float64) - t).max())
    return NOISE_MULT * float(f)


def assert_close(got, want, atol, what, rtol=RTOL):
    """element-wise |got - want| <= rtol |want| + atol (north_star: "within rtol 1e-4"), worst element reported"""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    err = np.abs(got - want)
    lim = rtol * np.abs(want) + atol
    if os.environ.get("NJODE_PARITY_STATS"):
        ex = np.maximum(err - rtol * np.abs(want), 0).max()
        with open(os.environ["NJODE_PARITY_STATS"], "a") as f:
            f.write("%s\t%.3g\t%.3g\t%.3g\n" % (what, ex / (atol / NOISE_MULT + 1e-300), ex, np.abs(want).max()))
    if np.any(err > lim):
        i = int(np.argmax(err - lim))
        raise AssertionError("%s: element %d got %.9g want %.9g |diff| %.3g > rtol %.1e * |want| + atol %.3g "
                             "(max-norm rel. err %.3g)" % (what, i, got.flat[i], want.flat[i], err.flat[i], rtol, atol,
                                                           rel_err(got, want)))


def _t64(batch, meta, cfg, sd, grad_hT=None, dropout_seed=None):
    """the fp64 evaluation of the same call: (hT, loss, grads)"""
    return orc.loss_and_grads(orc.Config(**cfg), sd, batch, meta["delta_t"], meta["T"], dtype=torch.float64,
                              grad_hT=grad_hT, dropout_seed=dropout_seed)


class kernel_tanh:
    """context: the oracle evaluates tanh with the kernels' formula (fp32 noise measurement only)"""

    def __enter__(self):
        orc.TANH = orc.tanh_kernel_formula

    def __exit__(self, *a):
        orc.TANH = torch.tanh


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
    G = torch.tensor(outs["G"]) if with_hT_grad else None
    t_hT, t_loss, t_g = _t64(batch, meta, cfg, sd, grad_hT=G)
    with kernel_tanh():
        k_hT, k_loss, k_g = orc.loss_and_grads(orc.Config(**cfg), sd, batch, meta["delta_t"], meta["T"], grad_hT=G)
    assert_close(loss.detach().numpy(), outs["loss"], noise_floor(outs["loss"], t_loss.detach().numpy(), k_loss.detach().numpy()),
                 name + " loss")
    assert_close(hT.detach().cpu().numpy(), outs["hT"], noise_floor(outs["hT"], t_hT.detach().numpy(), k_hT.detach().numpy()),
                 name + " hT")
    key = "grad/"
    obj = loss
    if with_hT_grad:
        obj = loss + (hT * G.to(hT.device)).sum().cpu()
        key = "gradG/"
    obj.backward()
    for n, p in m.named_parameters():
        assert p.grad is not None, n
        assert_close(p.grad.cpu().numpy(), outs[key + n], noise_floor(outs[key + n], t_g[n].numpy(), k_g[n].numpy()),
                     name + " " + key + n)


def check_path_call(name, device):
    cfg, meta, sd, batch, outs = cases.load_case(name)
    m = build_model(cfg, sd, device)
    m.eval()
    with torch.no_grad():
        hT, loss, path_t, path_h, path_y = call(m, batch, meta, device, return_path=True,
                                                get_loss=True, until_T=True)
        sd64 = {k: v.double() for k, v in sd.items()}
        M = batch.get("M")
        t = orc.forward(orc.Config(**cfg), sd64, batch["times"], batch["time_ptr"], batch["X"].double(), batch["obs_idx"],
                        meta["delta_t"], meta["T"], batch["start_X"].double(), batch["n_obs_ot"], return_path=True,
                        get_loss=True, until_T=True, M=None if M is None else M.double())
    assert np.array_equal(np.asarray(path_t, dtype=np.float64), outs["path_t"])     # exact
    assert_close(loss.numpy(), outs["loss_T"], noise_floor(outs["loss_T"], t[1].numpy()), name + " loss_T")
    assert_close(hT.cpu().numpy(), outs["hT_T"], noise_floor(outs["hT_T"], t[0].numpy()), name + " hT_T")
    assert_close(path_h.numpy(), outs["path_h"], noise_floor(outs["path_h"], t[3].numpy()), name + " path_h")
    assert_close(path_y.numpy(), outs["path_y"], noise_floor(outs["path_y"], t[4].numpy()), name + " path_y")


def check_against_oracle(cfg, batch, dt, T, seed, device, train=False, rtol=RTOL, grad_hT=False, report=None):
    """fresh seeded inputs: product vs the oracle run live (fp32).  train=True: dropout on, the
    oracle replays the device's counter-based keep-masks from the same seed.  ``report``: dict that receives the
    max-norm and worst element-wise errors per output (for tests that print them)."""
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
    t_hT, t_loss, t_g = orc.loss_and_grads(ocfg, sd, batch, dt, T, dropout_seed=drop_seed, grad_hT=G, dtype=torch.float64)
    with kernel_tanh():
        k_hT, k_loss, k_g = orc.loss_and_grads(ocfg, sd, batch, dt, T, dropout_seed=drop_seed, grad_hT=G)
    pairs = [("loss", loss.detach().numpy(), o_loss.numpy(), t_loss.numpy(), k_loss.numpy()),
             ("hT", hT.detach().cpu().numpy(), o_hT.numpy(), t_hT.numpy(), k_hT.numpy())]
    pairs += [("grad " + n, p.grad.cpu().numpy(), o_g[n].numpy(), t_g[n].numpy(), k_g[n].numpy()) for n, p in m.named_parameters()]
    for what, got, want, truth, kern in pairs:
        floor = noise_floor(want, truth, kern)
        if report is not None:
            report[what] = dict(max_norm_rel=rel_err(got, want), fp32_noise_floor=floor,
                                worst_abs=float(np.abs(np.asarray(got, dtype=np.float64) - want).max()))
        assert_close(got, want, floor, what, rtol=rtol)
