"""The oracle restatement (oracle/njode_oracle.py) against the committed outputs of the REAL
reference (tests/golden/*.npz, produced by tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

import cases
import oracle.njode_oracle as orc

NAMES = cases.golden_names()


def test_fixtures_present():
    assert len(NAMES) >= 8


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_training_call(name):
    cfg, meta, sd, batch, outs = cases.load_case(name)
    ocfg = orc.Config(**cfg)
    hT, loss, g = orc.loss_and_grads(ocfg, sd, batch, meta["delta_t"], meta["T"])
    np.testing.assert_allclose(float(loss), float(outs["loss"]), rtol=1e-6)
    np.testing.assert_allclose(hT.numpy(), outs["hT"], rtol=1e-5, atol=1e-6)
    for n in sd:
        np.testing.assert_allclose(g[n].numpy(), outs["grad/" + n], rtol=1e-4,
                                   atol=1e-6 * (np.abs(outs["grad/" + n]).max() + 1e-30))
    # objective with a gradient flowing into hT as well
    _, _, gG = orc.loss_and_grads(ocfg, sd, batch, meta["delta_t"], meta["T"],
                                  grad_hT=torch.tensor(outs["G"]))
    for n in sd:
        np.testing.assert_allclose(gG[n].numpy(), outs["gradG/" + n], rtol=1e-4,
                                   atol=1e-6 * (np.abs(outs["gradG/" + n]).max() + 1e-30))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_path_call(name):
    cfg, meta, sd, batch, outs = cases.load_case(name)
    ocfg = orc.Config(**cfg)
    M = batch.get("M")
    with torch.no_grad():
        hT, loss, path_t, path_h, path_y = orc.forward(
            ocfg, sd, batch["times"], batch["time_ptr"], batch["X"], batch["obs_idx"],
            meta["delta_t"], meta["T"], batch["start_X"], batch["n_obs_ot"], return_path=True,
            get_loss=True, until_T=True, M=M)
    assert np.array_equal(np.asarray(path_t, dtype=np.float64), outs["path_t"])   # schedule: exact
    np.testing.assert_allclose(float(loss), float(outs["loss_T"]), rtol=1e-6)
    np.testing.assert_allclose(hT.numpy(), outs["hT_T"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(path_h.numpy(), outs["path_h"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(path_y.numpy(), outs["path_y"], rtol=1e-5, atol=1e-6)


def test_fp64_oracle_close_to_fp32_reference():
    """the fp64 evaluation of the same function bounds the fp32 rounding noise of the reference
    itself -- this is what the rtol=1e-4 parity budget of the CUDA path is measured against."""
    cfg, meta, sd, batch, outs = cases.load_case("bs_ckpt1")
    ocfg = orc.Config(**cfg)
    hT, loss, g = orc.loss_and_grads(ocfg, sd, batch, meta["delta_t"], meta["T"], dtype=torch.float64)
    assert abs(float(loss) - float(outs["loss"])) / float(outs["loss"]) < 1e-5
    for n in sd:
        ref = outs["grad/" + n]
        assert np.abs(g[n].numpy() - ref).max() <= 2e-4 * np.abs(ref).max() + 1e-9
