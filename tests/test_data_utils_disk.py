"""On-disk dataset format of njode_b200.data_utils (NJODE/data_utils.py:42-275): datasets written here load with
the reference's own loader and collate to the same batch (CPU; the reference is imported when present), and the
GPU test drives create_dataset -> IrregularDataset -> DataLoader -> NJODE training steps end to end."""
import json
import os

import numpy as np
import pytest
import torch

import _reference
import cases
import hostsim_util
from njode_b200 import data_utils as du


def _fake_dataset(tmp, name="BlackScholes", n=12, d=2, steps=9, seed=5):
    rng = np.random.default_rng(seed)
    paths = 1.0 + 0.1 * rng.standard_normal((n, d, steps + 1)).cumsum(axis=2)
    obs = (rng.random((n, steps + 1)) < 0.4) * 1
    nb = obs[:, 1:].sum(axis=1)
    meta = dict(du.hyperparam_default, nb_paths=n, nb_steps=steps, dimension=d, model_name=name, dt=1.0 / steps)
    return du._register_and_write(name, json.dumps(meta, sort_keys=True), meta, paths, obs, nb), (paths, obs, nb, meta)


@pytest.fixture
def data_dir(tmp_path, monkeypatch):
    p = str(tmp_path) + "/training_data/"
    monkeypatch.setattr(du, "training_data_path", p)
    return p


def test_write_load_roundtrip_and_overview(data_dir):
    (path, time_id), (paths, obs, nb, meta) = _fake_dataset(data_dir)
    assert os.path.isfile(path + "data.npy") and os.path.isfile(path + "metadata.txt")
    sp, od, no, hp = du.load_dataset("BlackScholes", time_id)
    assert sp.dtype == np.float64 and od.dtype == np.int64 and no.dtype == np.int64
    assert np.array_equal(sp, paths) and np.array_equal(od, obs) and np.array_equal(no, nb)
    assert hp == json.loads(json.dumps(meta)) and du.load_metadata("BlackScholes") == hp
    assert du._get_time_id("BlackScholes") == time_id and du._get_time_id("Heston") is None
    df, csv = du.get_dataset_overview()
    assert list(df.columns) == ["name", "id", "description"] and int(df["id"].iloc[0]) == time_id
    ds = du.IrregularDataset("BlackScholes", idx=np.array([1, 3, 4, 7]))
    assert len(ds) == 4
    item = ds[2]
    assert item["stock_path"].shape == (1, 2, 10) and item["dt"] == meta["dt"] and item["idx"] == [2]


def test_dataset_is_read_identically_by_the_reference(data_dir, monkeypatch):
    ref = _reference.load_reference()
    if ref is None:
        pytest.skip("reference tree not present")
    (path, time_id), _ = _fake_dataset(data_dir, n=10, d=1, steps=14, seed=8)
    monkeypatch.setattr(ref.data_utils, "training_data_path", data_dir)
    a, b = du.load_dataset("BlackScholes"), ref.data_utils.load_dataset("BlackScholes")
    for x, y in zip(a[:3], b[:3]):
        assert np.array_equal(x, y)
    assert a[3] == b[3]
    idx = np.arange(10)
    ours = du.IrregularDataset("BlackScholes", idx=idx)
    theirs = ref.data_utils.IrregularDataset("BlackScholes", idx=idx)
    dl_a = torch.utils.data.DataLoader(ours, collate_fn=du.custom_collate_fn, batch_size=5, shuffle=False)
    dl_b = torch.utils.data.DataLoader(theirs, collate_fn=ref.data_utils.custom_collate_fn, batch_size=5, shuffle=False)
    for ba, bb in zip(dl_a, dl_b):
        assert np.array_equal(ba["times"], bb["times"]) and np.array_equal(ba["time_ptr"], bb["time_ptr"])
        for k in ("obs_idx", "X", "start_X", "n_obs_ot"):
            assert torch.equal(ba[k], bb[k]), k


@pytest.mark.gpu
def test_create_dataset_and_train_steps_on_device(data_dir):
    """demo.py / train.py flow (NJODE/demo.py:64-81, train.py:243-264,493-522) on the B200 modules"""
    from njode_b200 import models
    hostsim_util.uninstall()
    hp = dict(du.hyperparam_default, nb_paths=600, nb_steps=50, obs_perc=0.2)
    path, time_id = du.create_dataset("OrnsteinUhlenbeck", hp, seed=3)
    sp, od, no, meta = du.load_dataset("OrnsteinUhlenbeck", time_id)
    assert sp.shape == (600, 1, 51) and od.shape == (600, 51) and meta["dt"] == pytest.approx(1.0 / 50)
    assert np.array_equal(no, od[:, 1:].sum(axis=1))
    assert abs(od[:, 1:].mean() - 0.2) < 0.02
    assert abs(sp[:, 0, -1].mean() - (4 + (1 - 4) * (1 - 2.0 / 50) ** 50)) < 0.05      # Euler mean of the OU scheme
    ds = du.IrregularDataset("OrnsteinUhlenbeck", time_id=time_id, idx=np.arange(500))
    dl = torch.utils.data.DataLoader(ds, collate_fn=du.custom_collate_fn, batch_size=100, shuffle=True)
    torch.manual_seed(0)
    model = models.NJODE(**cases.demo_cfg(dropout_rate=0.1)).to("cuda:0")
    opt = torch.optim.Adam(model.parameters(), lr=5e-3, weight_decay=5e-4)
    losses = []
    for epoch in range(6):
        model.train()
        tot = 0.0
        for b in dl:
            opt.zero_grad()
            hT, loss = model(b["times"], b["time_ptr"], b["X"].to("cuda:0"), b["obs_idx"], meta["dt"], meta["maturity"],
                             b["start_X"].to("cuda:0"), b["n_obs_ot"])
            loss.backward()
            opt.step()
            tot += float(loss.detach())
        losses.append(tot)
    assert losses[-1] < 0.7 * losses[0]


@pytest.mark.gpu
def test_create_combined_dataset_on_device(data_dir):
    hp1 = dict(du.hyperparam_default, nb_paths=200, nb_steps=20, maturity=0.5, mean=10)
    hp2 = dict(du.hyperparam_default, nb_paths=200, nb_steps=20, maturity=0.5, mean=10)
    path, time_id = du.create_combined_dataset(["OrnsteinUhlenbeck", "BlackScholes"], [hp1, hp2], seed=1)
    name = "combined_OrnsteinUhlenbeck_BlackScholes"
    sp, od, no, meta = du.load_dataset(name, time_id)
    assert sp.shape == (200, 1, 41) and od.shape == (200, 41)
    assert meta["model_name"] == "combined" and meta["maturity"] == 1.0 and meta["stock_model_names"] == ["OrnsteinUhlenbeck", "BlackScholes"]
    assert np.array_equal(no, od[:, 1:].sum(axis=1))
