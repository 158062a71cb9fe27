"""Size-independent properties of the CUDA path at BASELINE.json's FULL sizes, where the oracle is too slow to run:
  * shard invariance (what data parallelism relies on): loss and gradients of a batch = sums over its path shards when
    every shard normalises by the global batch size and keys dropout by global path ids;
  * path-permutation invariance (the reference's loss is a sum over rows: NJODE/models.py:105-106);
  * linearity of the backward pass in the incoming gradient;
  * a path's own numbers do not depend on what else is in the batch (hT rows of a sub-batch);
  * determinism of the forward pass (fixed-order reductions).
Tolerances: fp32 kernels rtol 1e-4 (BASELINE.json); tcgen05 path at its stated bf16 tolerance."""
import numpy as np
import pytest
import torch

import cases
from njode_b200 import models
from njode_b200.dist import shard_batch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def run(model, batch, dt, T, grad_scale=1.0, until_T=False):
    for p in model.parameters():
        p.grad = None
    kw = {"M": batch["M"]} if "M" in batch else {}
    hT, loss = model(batch["times"], batch["time_ptr"], batch["X"], batch["obs_idx"], dt, T, batch["start_X"], batch["n_obs_ot"],
                     until_T=until_T, **kw)
    (loss * grad_scale).backward()
    g = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).double().cpu().numpy()
    return float(loss.detach()), hT.detach().cpu().numpy(), g


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / (np.abs(np.asarray(b)).max() + 1e-30))


def make_model(cfg, train, seed=0, tensor_cores="off"):
    torch.manual_seed(seed)
    m = models.NJODE(**cfg).to(DEV)
    m.tensor_cores = tensor_cores
    m.train() if train else m.eval()
    return m


@pytest.mark.parametrize("train", [False, True])
def test_config2_full_size_shard_invariance(train):
    """BASELINE configs[1] size: 20 000 paths x 100 steps, demo nets; 4 shards with global normalisation and path ids"""
    B = 20000
    batch = cases.grid_batch(B, 1, 100, 0.1, seed=41)
    m = make_model(cases.demo_cfg(dropout_rate=0.1), train)
    torch.manual_seed(99)
    loss, hT, g = run(m, batch, 0.01, 1.0)
    loss_s, g_s, hT_s = 0.0, 0.0, []
    for r in range(4):
        sb, lo = shard_batch(batch, r, 4)
        m.batch_size_norm, m.path_id_offset = B, lo
        torch.manual_seed(99)                      # same dropout seed as the full batch
        l, h, gg = run(m, sb, 0.01, 1.0)
        loss_s += l; g_s = g_s + gg; hT_s.append(h)
    m.batch_size_norm, m.path_id_offset = None, 0
    assert abs(loss_s - loss) < 1e-4 * abs(loss)
    assert rel(np.concatenate(hT_s), hT) < 1e-5
    assert rel(g_s, g) < 1e-4


def test_config2_full_size_permutation_invariance_and_linearity():
    B = 20000
    batch = cases.grid_batch(B, 1, 100, 0.1, seed=42)
    m = make_model(cases.demo_cfg(), False)
    loss, hT, g = run(m, batch, 0.01, 1.0)
    # determinism of the forward pass
    loss2, hT2, g2 = run(m, batch, 0.01, 1.0)
    assert loss == loss2 and np.array_equal(hT, hT2) and rel(g2, g) < 1e-5
    # backward is linear in the incoming gradient
    _, _, g3 = run(m, batch, 0.01, 1.0, grad_scale=-2.5)
    assert rel(g3, -2.5 * g) < 1e-5
    # relabel the paths: reversed order (rows inside a time slot re-sorted by the new path index, as the collate would)
    perm = np.arange(B)[::-1].copy()                      # new index of old path p is B - 1 - p
    paths = batch["true_paths"][perm]
    obs = batch["observed_dates"][perm]
    pb = cases.collate_arrays(paths, obs, 0.01)
    lossp, hTp, gp = run(m, pb, 0.01, 1.0)
    assert abs(lossp - loss) < 1e-5 * abs(loss)
    assert rel(hTp[::-1], hT) < 1e-5
    assert rel(gp, g) < 1e-4


def test_paths_do_not_interact_masked_model_physionet_shape():
    """config 4 shape (masked, d = H = 41): hT of a path is the same in a batch of 64 as in a sub-batch of 16"""
    batch = cases.irregular_batch(64, 41, 300, seed=43, masked=True, times_f32=True, obs_at_zero=True, row_prob=0.05,
                                  feat_prob=0.12)
    m = make_model(cases.CONFIGS["masked_physio"], False)
    dt, T = 0.016 / 48 * 10, 1 + 1e-12
    loss, hT, g = run(m, batch, dt, T, until_T=True)
    sb, lo = shard_batch(batch, 1, 4)
    m.batch_size_norm = 64
    l1, h1, g1 = run(m, sb, dt, T, until_T=True)
    m.batch_size_norm = None
    assert rel(h1, hT[16:32]) < 1e-5
    parts = []
    m.batch_size_norm = 64
    for r in range(4):
        parts.append(run(m, shard_batch(batch, r, 4)[0], dt, T, until_T=True))
    m.batch_size_norm = None
    assert abs(sum(p[0] for p in parts) - loss) < 1e-4 * abs(loss)
    assert rel(sum(p[2] for p in parts), g) < 1e-4


def test_gru_model_shard_invariance():
    batch = cases.grid_batch(2000, 1, 100, 0.1, seed=44)
    m = make_model(cases.demo_cfg(use_rnn=True, dropout_rate=0.1), True)
    torch.manual_seed(5)
    loss, hT, g = run(m, batch, 0.01, 1.0)
    tot_l, tot_g = 0.0, 0.0
    for r in range(2):
        sb, lo = shard_batch(batch, r, 2)
        m.batch_size_norm, m.path_id_offset = 2000, lo
        torch.manual_seed(5)
        l, h, gg = run(m, sb, 0.01, 1.0)
        tot_l += l; tot_g = tot_g + gg
    assert abs(tot_l - loss) < 1e-4 * abs(loss) and rel(tot_g, g) < 1e-4


def test_config5_shape_shard_invariance_on_tensor_cores():
    """config 5 architecture (d=16, H=256, 4x256) on the tcgen05 kernels, 2048 paths x 200 steps: two shards sum to the
    whole batch (bf16 operands: the same rounding in both runs, so the agreement is far inside the bf16 tolerance)"""
    nn = [[256, "tanh"]] * 4
    cfg = cases.demo_cfg(input_size=16, output_size=16, hidden_size=256, ode_nn=nn, enc_nn=nn, readout_nn=nn, dropout_rate=0.1)
    B = 2048
    batch = cases.grid_batch(B, 16, 200, 0.1, seed=45)
    m = make_model(cfg, True, tensor_cores="on")
    torch.manual_seed(6)
    loss, hT, g = run(m, batch, 0.005, 1.0)
    assert m.last_forward_path == "tcgen05"
    tot_l, tot_g, hs = 0.0, 0.0, []
    for r in range(2):
        sb, lo = shard_batch(batch, r, 2)
        m.batch_size_norm, m.path_id_offset = B, lo
        torch.manual_seed(6)
        l, h, gg = run(m, sb, 0.005, 1.0)
        tot_l += l; tot_g = tot_g + gg; hs.append(h)
    assert abs(tot_l - loss) < 1e-4 * abs(loss)
    assert rel(np.concatenate(hs), hT) < 1e-4
    assert rel(tot_g, g) < 2e-3


def test_training_trajectory_matches_the_oracle_under_adam():
    """30 Adam steps on the CUDA kernels vs 30 Adam steps on the oracle (CPU, autograd), same data / weights / optimiser
    settings as train.py (lr 1e-3, weight_decay 5e-4), dropout off: the per-step losses stay together (the difference is
    the accumulated effect of 1e-5-level gradient differences)"""
    import oracle.njode_oracle as orc
    cfg = cases.demo_cfg()
    ocfg = orc.Config(**cfg)
    sd0 = orc.init_state_dict(ocfg, seed=21)
    batches = [cases.grid_batch(64, 1, 25, 0.2, seed=50 + i) for i in range(3)]
    m = models.NJODE(**cfg)
    m.load_state_dict(sd0)
    m.to(DEV).train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=0.0005)
    osd = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
    oopt = torch.optim.Adam(list(osd.values()), lr=1e-3, weight_decay=0.0005)
    ours, theirs = [], []
    for it in range(30):
        b = batches[it % 3]
        opt.zero_grad()
        hT, loss = m(b["times"], b["time_ptr"], b["X"], b["obs_idx"], 0.04, 1.0, b["start_X"], b["n_obs_ot"])
        loss.backward()
        opt.step()
        ours.append(float(loss.detach()))
        oopt.zero_grad()
        _, oloss = orc.forward(ocfg, osd, b["times"], b["time_ptr"], b["X"], b["obs_idx"], 0.04, 1.0, b["start_X"], b["n_obs_ot"])
        oloss.backward()
        oopt.step()
        theirs.append(float(oloss.detach()))
    ours, theirs = np.array(ours), np.array(theirs)
    assert theirs[-1] < theirs[0]                                   # it trains
    np.testing.assert_allclose(ours, theirs, rtol=2e-3)
    for (n, p) in m.named_parameters():
        assert rel(p.detach().cpu().numpy(), osd[n].detach().numpy()) < 5e-3, n
