"""GPU: njode_sde_generate / njode_collate (njode_b200/csrc/njode_sde.cu) through the C ABI against
oracle/sde_oracle.py (same Philox stream: fp64 agreement to libm rounding, integer outputs exact), in
distribution against the reference generator (committed moments) and the closed-form moments of the
Euler scheme, and end to end: a device-collated batch drives the model to the same loss as the host path."""
import json
import os

import numpy as np
import pytest
import torch

import cases
from oracle import sde_oracle as so
from njode_b200 import models, stock_model

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "sde_ref.npz"))
HP = dict(drift=2., volatility=0.3, mean=4, speed=2., correlation=0.5, nb_paths=300, nb_steps=37, S0=[1., 1.5, 0.7],
          maturity=1., dimension=3, sine_coeff=None, v0=0.5, return_vol=False, obs_perc=0.2)


@pytest.mark.parametrize("sine", [None, 2.5])
@pytest.mark.parametrize("name", ["BlackScholes", "OrnsteinUhlenbeck", "Heston", "HestonWOFeller", "HestonWOFeller_vol"])
def test_generator_matches_oracle_on_the_same_stream(name, sine):
    model = name.split("_")[0]
    hp = dict(HP, return_vol=name.endswith("_vol"), sine_coeff=sine)
    m = stock_model.STOCK_MODELS[model](**hp, seed=1234567890123, first_path=1000)
    paths, obs, nb, dt = m.generate_paths_device(obs_perc=0.2)
    o_paths, o_obs, o_nb = so.generate(model, hp, 1234567890123, 1000, 300, obs_perc=0.2)
    assert paths.shape == o_paths.shape
    np.testing.assert_allclose(paths.cpu().numpy(), o_paths, rtol=1e-10, atol=1e-12)      # fp64, libm differences only
    assert np.array_equal(obs.cpu().numpy(), o_obs) and np.array_equal(nb.cpu().numpy(), o_nb)   # integer outputs: exact


def test_generator_is_sharding_invariant_and_honours_start_X():
    hp = dict(HP, dimension=1, S0=1.0, nb_paths=64)
    a, *_ = stock_model.Heston(**hp, seed=7, first_path=0).generate_paths_device(nb_paths=64)
    b, *_ = stock_model.Heston(**hp, seed=7, first_path=40).generate_paths_device(nb_paths=24)
    assert torch.equal(a[40:], b)
    sx = np.linspace(0.5, 2.0, 64).reshape(64, 1)
    c, *_ = stock_model.BlackScholes(**hp, seed=7).generate_paths_device(start_X=sx)
    assert np.allclose(c[:, :, 0].cpu().numpy(), sx)
    p, dt = stock_model.BlackScholes(**hp, seed=7).generate_paths()
    assert isinstance(p, np.ndarray) and p.dtype == np.float64 and p.shape == (64, 1, 38) and dt == 1.0 / 37


def test_moments_against_closed_forms_and_reference_generator():
    demo = json.loads(str(GOLD["demo_hp"]))
    n = 400000
    k = np.arange(101)
    dt = 0.01
    for name in ("BlackScholes", "OrnsteinUhlenbeck", "Heston"):
        paths, *_ = stock_model.STOCK_MODELS[name](**dict(demo, nb_paths=n), seed=99).generate_paths_device()
        x = paths[:, 0, :]
        mean, var = x.mean(0).cpu().numpy(), x.var(0).cpu().numpy()
        # reference generator, 20 000 paths: agreement within its own Monte-Carlo error (5 sigma)
        r_mean, r_var = GOLD["moments/%s/mean" % name], GOLD["moments/%s/var" % name]
        assert np.all(np.abs(mean - r_mean) <= 5 * np.sqrt(r_var / 20000) + 1e-12), name
        if name != "Heston":                # Heston spot (vol ~ 200 %) has no stable sample variance at 20 000 paths
            assert np.all(np.abs(var - r_var) <= 0.08 * r_var + 1e-12), name
        # quartiles: robust for every model; sampling error of a quantile ~ sqrt(p(1-p)/n) / density
        q = torch.quantile(x[:100000].float(), torch.tensor([0.25, 0.5, 0.75], device=x.device), dim=0).cpu().numpy()
        r_q = GOLD["moments/%s/quantiles" % name]
        iqr = r_q[2] - r_q[0]
        assert np.all(np.abs(q - r_q) <= 0.05 * iqr[None, :] + 1e-9), name
        if name == "BlackScholes":          # E[S_k] = S0 (1 + mu dt)^k ; E[S_k^2] = S0^2 ((1 + mu dt)^2 + sigma^2 dt)^k
            em = (1 + 2.0 * dt) ** k
            e2 = ((1 + 2.0 * dt) ** 2 + 0.09 * dt) ** k
            assert np.all(np.abs(mean - em) <= 5 * np.sqrt((e2 - em ** 2) / n) + 1e-12)
            assert np.all(np.abs(var - (e2 - em ** 2)) <= 0.03 * (e2 - em ** 2) + 1e-12)
        if name == "OrnsteinUhlenbeck":     # E[S_k] = m + (S0 - m)(1 - kappa dt)^k
            em = 4 + (1 - 4) * (1 - 2.0 * dt) ** k
            assert np.all(np.abs(mean - em) <= 5 * np.sqrt(var / n) + 1e-12)


def test_mask_rate():
    m = stock_model.BlackScholes(**dict(HP, dimension=1, S0=1.0, nb_steps=100, nb_paths=20000), seed=3)
    _, obs, nb, _ = m.generate_paths_device(obs_perc=0.1)
    assert abs(float(obs[:, 0].float().mean()) - float(obs[:, 1:].float().mean())) < 0.02      # column 0 is drawn like the others (NJODE/data_utils.py:79-80)
    rate = obs[:, 1:].float().mean().item()
    assert abs(rate - 0.1) < 0.002
    assert torch.equal(nb, obs[:, 1:].sum(1).to(torch.int32))


@pytest.mark.parametrize("B", [1, 31, 200, 1000])
def test_device_collate_matches_oracle_exactly(B):
    hp = dict(HP, nb_paths=1500, dimension=2, S0=[1.0, 2.0], nb_steps=25, obs_perc=0.15)
    ds = stock_model.DeviceDataset("Heston", hp, seed=11)
    sel = np.random.default_rng(B).permutation(1500)[:B]
    got = ds.collate(sel)
    paths, obs, nb = ds.paths.cpu().numpy()[sel], ds.observed.cpu().numpy()[sel], ds.nb_obs.cpu().numpy()[sel]
    want = so.collate(paths, obs, nb, ds.dt)
    assert np.array_equal(got["times"], want["times"]) and np.array_equal(got["time_ptr"], want["time_ptr"])
    assert np.array_equal(got["obs_idx"].numpy(), want["obs_idx"])
    assert np.array_equal(got["X"].cpu().numpy(), want["X"]) and np.array_equal(got["start_X"].cpu().numpy(), want["start_X"])
    assert np.array_equal(got["n_obs_ot"].numpy(), want["n_obs_ot"])


def test_device_dataset_drives_the_model_like_the_host_collate():
    from njode_b200 import data_utils
    hp = dict(HP, nb_paths=600, dimension=1, S0=1.0, nb_steps=50, obs_perc=0.1)
    ds = stock_model.DeviceDataset("BlackScholes", hp, seed=5)
    sel = np.arange(100, 300)
    b_dev = ds.collate(sel)
    b_host = data_utils.collate_paths(ds.paths.cpu().numpy()[sel], ds.observed.cpu().numpy()[sel],
                                      ds.nb_obs.cpu().numpy()[sel], ds.dt)
    torch.manual_seed(0)
    m = models.NJODE(**cases.CONFIGS["demo"]).to("cuda:0").eval()
    out = []
    for b in (b_dev, b_host):
        with torch.no_grad():
            hT, loss = m(b["times"], b["time_ptr"], b["X"], b["obs_idx"], ds.dt, 1.0, b["start_X"], b["n_obs_ot"])
        out.append((hT.cpu().numpy(), float(loss)))
    assert out[0][1] == out[1][1] and np.array_equal(out[0][0], out[1][0])


def test_device_collate_with_func_appl_X_matches_the_host_collate_generator():
    """'func_appl_X': ["power-2"] (NJODE/data_utils.py:352-416): X and start_X gain the squared coordinates"""
    from njode_b200 import data_utils
    hp = dict(HP, nb_paths=400, dimension=2, S0=[1.0, 2.0], nb_steps=20, obs_perc=0.2)
    ds = stock_model.DeviceDataset("BlackScholes", hp, seed=7)
    sel = np.arange(50, 150)
    got = ds.collate(sel, func_names=["power-2"])
    want = data_utils.collate_paths(ds.paths.cpu().numpy()[sel], ds.observed.cpu().numpy()[sel], ds.nb_obs.cpu().numpy()[sel],
                                    ds.dt, functions=[data_utils._get_func("power-2")])
    assert got["X"].shape[1] == 4 and got["start_X"].shape[1] == 4
    np.testing.assert_allclose(got["X"].cpu().numpy(), want["X"].numpy(), rtol=1e-6)
    np.testing.assert_allclose(got["start_X"].cpu().numpy(), want["start_X"].numpy(), rtol=1e-6)
    assert np.array_equal(got["obs_idx"].numpy(), want["obs_idx"].numpy())


@pytest.mark.parametrize("name,extra", [("BlackScholes", {}), ("OrnsteinUhlenbeck", {"sine_coeff": 3.0}), ("Heston", {}),
                                        ("HestonWOFeller", {"return_vol": True, "v0": 0.5})])
def test_cond_exp_kernel_matches_the_numpy_event_loop(name, extra):
    """njode_cond_exp against StockModel.compute_cond_exp (the restatement of NJODE/stock_model.py:50-151) on the records
    of a return_path call; then NJODE.evaluate: device route == NumPy route"""
    hp = dict(HP, nb_paths=300, dimension=1, S0=1.0, nb_steps=40, obs_perc=0.15, **extra)
    ds = stock_model.DeviceDataset(name, hp, seed=9)
    b = ds.collate(np.arange(300))
    d = ds.paths.shape[1]
    cfg = cases.demo_cfg(input_size=d, output_size=d, hidden_size=10)
    torch.manual_seed(1)
    m = models.NJODE(**cfg).to("cuda:0").eval()
    sm = ds.model
    assert sm.supports_cond_exp_device(d)
    pb = m.prepare_batch(b["times"], b["time_ptr"], b["X"], b["obs_idx"], ds.dt, 1.0, b["start_X"], None,
                         return_path=True, get_loss=False, until_T=True)
    got = sm.compute_cond_exp_device(pb, d).cpu().numpy()
    _, want_t, want = sm.compute_cond_exp(b["times"], b["time_ptr"], b["X"].cpu().numpy(), b["obs_idx"].numpy(), ds.dt, 1.0,
                                          b["start_X"].cpu().numpy(), b["n_obs_ot"].numpy(), return_path=True, get_loss=False)
    assert got.shape == want.shape and np.array_equal(np.asarray(pb.sched.path_t, dtype=np.float64), np.asarray(want_t, dtype=np.float64))
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-7)
    args = (b["times"], b["time_ptr"], b["X"], b["obs_idx"], ds.dt, 1.0, b["start_X"], b["n_obs_ot"], sm)
    on_device = m.evaluate(*args)
    on_host = m.evaluate(*args, diff_fun=lambda x, y: np.mean((x - y) ** 2))
    assert abs(on_device - on_host) <= 1e-5 * abs(on_host)


_CE = np.load(os.path.join(cases.GOLDEN_DIR, "condexp_ref.npz"))
_CE_META = json.loads(str(_CE["meta"]))


@pytest.mark.parametrize("name", ["bs", "bs_d2_sine", "ou_sine", "heston", "hwof", "hwof_vol", "bs_tail_T2"])
def test_cond_exp_kernel_matches_the_reference_fixture(name):
    """njode_cond_exp against the outputs of the REAL reference's compute_cond_exp (NJODE/stock_model.py:50-151;
    tests/golden/make_condexp_golden.py) on the same batch: same records in the same order as NJODE.forward's path_t"""
    meta = _CE_META[name]
    g = lambda k: _CE["%s/%s" % (name, k)]
    d = meta["d"]
    m = models.NJODE(**cases.demo_cfg(input_size=d, output_size=d, hidden_size=10)).to("cuda:0").eval()
    smodel = stock_model.STOCK_MODELS[meta["model"]](**meta["hp"])
    assert smodel.supports_cond_exp_device(d)
    pb = m.prepare_batch(g("times"), g("time_ptr"), torch.from_numpy(g("X")), torch.from_numpy(g("obs_idx")),
                         meta["delta_t"], meta["T"], torch.from_numpy(g("start_X")), None, return_path=True,
                         get_loss=False, until_T=True)
    got = smodel.compute_cond_exp_device(pb, d).cpu().numpy()
    assert np.array_equal(np.asarray(pb.sched.path_t, dtype=np.float64), g("path_t"))
    np.testing.assert_allclose(got, g("path_y"), rtol=2e-6, atol=1e-7)


def test_device_collated_batch_needs_no_host_round_trip():
    """f1: DeviceDataset.collate(on_device=True) hands device index arrays (time_ptr, obs_idx, n_obs_ot) straight to
    NJODE.forward; same loss / hT as the host-index route bit for bit, same gradients up to the summation order (the
    two index builders order units of equal length differently)"""
    hp = dict(HP, nb_paths=3000, dimension=2, S0=[1.0, 1.5], nb_steps=60, obs_perc=0.1)
    ds = stock_model.DeviceDataset("BlackScholes", hp, seed=11)
    sel = np.arange(500, 2500)
    outs = []
    for on_device in (False, True):
        b = ds.collate(sel, on_device=on_device)
        if on_device:
            assert all(b[k].device.type == "cuda" for k in ("time_ptr", "obs_idx", "n_obs_ot", "X", "start_X"))
        torch.manual_seed(0)
        m = models.NJODE(**cases.demo_cfg(input_size=2, output_size=2)).to("cuda:0").eval()
        hT, loss = m(b["times"], b["time_ptr"], b["X"], b["obs_idx"], ds.dt, 1.0, b["start_X"], b["n_obs_ot"])
        loss.backward()
        outs.append((float(loss.detach()), hT.detach().cpu().numpy(), torch.cat([p.grad.reshape(-1) for p in m.parameters()]).cpu().numpy()))
    assert outs[0][0] == outs[1][0] and np.array_equal(outs[0][1], outs[1][1])
    np.testing.assert_allclose(outs[0][2], outs[1][2], rtol=1e-5, atol=1e-6 * np.abs(outs[1][2]).max())
