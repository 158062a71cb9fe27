"""CPU-only: the kernel source (njode_b200/csrc/njode_core.cuh) compiled as a sequential host
simulation, driven through the product's Python layer (schedule, staging, autograd bridge), against
the golden outputs of the real reference.  Exercises the exact device logic without a GPU; the GPU
parity tests proper are tests/test_gpu_parity.py."""
import os

import numpy as np
import pytest
import torch

import cases
import hostsim_util
import parity_util
from njode_b200 import models

NAMES = cases.golden_names()


@pytest.fixture(autouse=True)
def sim_runner():
    hostsim_util.install()
    # the golden cases are tiny: left alone the planner would give every segment batch to the thread-per-neuron kernels
    # of small batches; the tests below choose (NJODE_SEG_TPN = 1 / unset) where those are the subject
    os.environ["NJODE_SEG_TPN"] = "0"
    yield
    os.environ.pop("NJODE_SEG_TPN", None)
    hostsim_util.uninstall()
    os.environ.pop("NJODE_FORCE_TILE", None)
    os.environ.pop("NJODE_FORCE_TR", None)
    os.environ.pop("NJODE_NO_SEG", None)
    os.environ.pop("NJODE_INDEX", None)
    os.environ.pop("NJODE_FORCE_NW", None)
    os.environ.pop("NJODE_FORCE_DW", None)
    os.environ.pop("NJODE_SEG_HELPERS", None)
    os.environ.pop("NJODE_SEG_LOW", None)
    os.environ.pop("NJODE_NO_PATH", None)
    os.environ.pop("NJODE_PATH_R", None)
    os.environ.pop("NJODE_NO_STAT", None)
    os.environ.pop("NJODE_NO_TPN", None)
    os.environ.pop("NJODE_FORCE_TPN", None)
    os.environ.pop("NJODE_FORCE_STAT", None)
    os.environ.pop("NJODE_SAVE_ACTIVATIONS", None)
    os.environ.pop("NJODE_NO_PIPE", None)
    os.environ.pop("NJODE_FORCE_PIPE", None)
    os.environ.pop("NJODE_SIM_SMS", None)


@pytest.mark.parametrize("name", NAMES)
def test_training_call(name):
    parity_util.check_training_call(name, "cpu")


@pytest.mark.parametrize("name", NAMES)
def test_training_call_with_hT_gradient(name):
    parity_util.check_training_call(name, "cpu", with_hT_grad=True)


@pytest.mark.parametrize("name", NAMES)
def test_path_call(name):
    parity_util.check_path_call(name, "cpu")


@pytest.mark.parametrize("tile", [8, 16, 64])
@pytest.mark.parametrize("name", ["bs_ckpt1", "masked_small", "curt_nobias_relu", "gru_masked"])
def test_tile_size_invariance(name, tile):
    os.environ["NJODE_FORCE_TILE"] = str(tile)
    os.environ["NJODE_NO_PATH"] = "1"            # the generic CTA-cooperative kernels (njode_core.cuh)
    parity_util.check_training_call(name, "cpu", with_hT_grad=True)
    parity_util.check_path_call(name, "cpu")


@pytest.mark.parametrize("dw", [0, 2])
@pytest.mark.parametrize("name", ["masked_small", "gru_d3_nores", "res_case2"])
def test_gradient_image_residency(name, dw):
    """gradient image: whole in shared memory (default for small nets), ODE-network part only (2), or the per-CTA
    partial in global memory (0) -- same gradients"""
    os.environ["NJODE_FORCE_DW"] = str(dw)
    os.environ["NJODE_NO_SEG"] = "1"
    os.environ["NJODE_NO_PATH"] = "1"
    parity_util.check_training_call(name, "cpu", with_hT_grad=True)


@pytest.mark.parametrize("tile", [1, 3, 5])
@pytest.mark.parametrize("name", ["heston_ckpt2", "masked_small", "gru_demo"])
def test_small_tiles(name, tile):
    """tile heights the planner picks for small whole-path batches (PhysioNet batch of 50 -> one path per CTA)"""
    os.environ["NJODE_FORCE_TILE"] = str(tile)
    os.environ["NJODE_NO_SEG"] = "1"
    os.environ["NJODE_NO_PATH"] = "1"
    parity_util.check_training_call(name, "cpu", with_hT_grad=True)
    parity_util.check_path_call(name, "cpu")


def test_no_cpu_fallback_without_test_runner():
    hostsim_util.uninstall()
    cfg, meta, sd, batch, outs = cases.load_case("bs_ckpt1")
    m = parity_util.build_model(cfg, sd, "cpu")
    with pytest.raises(Exception) as ei:
        parity_util.call(m, batch, meta, "cpu")
    assert "CUDA" in str(ei.value)


def test_train_mode_dropout_masks_replayed_by_oracle():
    cfg = cases.demo_cfg(dropout_rate=0.1)
    batch = cases.grid_batch(24, 1, 20, 0.2, seed=6)
    parity_util.check_against_oracle(cfg, batch, 0.05, 1.0, seed=2, device="cpu", train=True, grad_hT=True)


def test_train_mode_dropout_masked_model():
    cfg = dict(cases.CONFIGS["masked_small"], dropout_rate=0.25)
    batch = cases.irregular_batch(6, 5, 10, seed=21, masked=True, times_f32=True)
    parity_util.check_against_oracle(cfg, batch, 0.05, 1 + 1e-12, seed=3, device="cpu", train=True)


def test_gru_jump_train_mode_dropout():
    """use_rnn=True: the GRU cell replaces the encoder at the jumps (NJODE/models.py:202-217,460-461)"""
    cfg = cases.demo_cfg(use_rnn=True, dropout_rate=0.2, bias=False, hidden_size=6)
    batch = cases.grid_batch(30, 1, 20, 0.25, seed=26)
    parity_util.check_against_oracle(cfg, batch, 0.05, 1.0, seed=5, device="cpu", train=True, grad_hT=True)


def test_physionet_shape_masked():
    batch = cases.irregular_batch(5, 41, 12, seed=7, masked=True, times_f32=True, obs_at_zero=True,
                                  row_prob=0.3, feat_prob=0.12)
    parity_util.check_against_oracle(cases.CONFIGS["masked_physio"], batch, 0.02, 1 + 1e-12, seed=3, device="cpu")


def test_global_weight_image_path():
    """weights too large for shared memory -> parameter / gradient images stay in global memory"""
    cfg = cases.demo_cfg(input_size=4, output_size=4, hidden_size=128,
                         ode_nn=[[256, "tanh"], [256, "tanh"]], enc_nn=[[256, "tanh"]],
                         readout_nn=[[256, "tanh"]])
    batch = cases.grid_batch(6, 4, 5, 0.4, seed=8)
    parity_util.check_against_oracle(cfg, batch, 0.2, 1.0, seed=4, device="cpu")


# ---- segment fast path (njode_b200/csrc/njode_seg.cuh): every tile height, train and eval ----
SEG_NAMES = [n for n in NAMES if "masked" not in n and "gru" not in n]


@pytest.mark.parametrize("tr", [1, 2, 4])
@pytest.mark.parametrize("name", SEG_NAMES)
def test_segment_path_tile_heights(name, tr):
    os.environ["NJODE_FORCE_TR"] = str(tr)
    parity_util.check_training_call(name, "cpu", with_hT_grad=True)
    parity_util.check_training_call(name, "cpu")


@pytest.mark.parametrize("name", ["bs_ckpt1", "curt_nobias_relu"])
def test_generic_kernels_still_serve_segment_units(name):
    os.environ["NJODE_NO_SEG"] = "1"
    parity_util.check_training_call(name, "cpu", with_hT_grad=True)


@pytest.mark.parametrize("tr", [1, 2, 4])
def test_segment_path_dropout_masks_replayed_by_oracle(tr):
    os.environ["NJODE_FORCE_TR"] = str(tr)
    cfg = cases.demo_cfg(dropout_rate=0.2)
    batch = cases.grid_batch(40, 1, 25, 0.2, seed=16)
    parity_util.check_against_oracle(cfg, batch, 0.04, 1.0, seed=12, device="cpu", train=True, grad_hT=True)


@pytest.mark.parametrize("helpers", [0, 1])
def test_segment_path_wide_layers_use_output_chunks(helpers):
    """hidden width 100 > 64: two output chunks per layer in the warp GEMM (64 + 40 outputs); with / without the dW
    helper warps of the backward (threads that own no rows)"""
    os.environ["NJODE_SEG_HELPERS"] = str(helpers)
    cfg = cases.demo_cfg(input_size=2, output_size=2, hidden_size=6, dropout_rate=0.1,
                         ode_nn=[[100, "tanh"], [70, "relu"]], enc_nn=[[100, "tanh"]], readout_nn=[[33, "tanh"]])
    batch = cases.grid_batch(20, 2, 10, 0.3, seed=18)
    parity_util.check_against_oracle(cfg, batch, 0.1, 1.0, seed=14, device="cpu", train=True, grad_hT=True)
    parity_util.check_against_oracle(cfg, batch, 0.1, 1.0, seed=14, device="cpu", train=False)


@pytest.mark.parametrize("name", ["bs_ckpt1", "easy_w07_nores", "curt_nobias_relu"])
def test_segment_forward_low_regions(name):
    """8-row warp regions / tiles of at most 8 rows (the layout big nets get when the image leaves little room)"""
    os.environ["NJODE_SEG_LOW"] = "1"
    parity_util.check_training_call(name, "cpu", with_hT_grad=True)
    cfg = cases.demo_cfg(dropout_rate=0.2)
    batch = cases.grid_batch(40, 1, 25, 0.2, seed=16)
    parity_util.check_against_oracle(cfg, batch, 0.04, 1.0, seed=12, device="cpu", train=True, grad_hT=True)


@pytest.mark.parametrize("mode", ["host", "device"])
@pytest.mark.parametrize("name", ["bs_ckpt1", "irregular_demo", "masked_small"])
def test_both_index_builders(name, mode):
    """per-path CSR + work units built by NumPy on the host (small batches) or by tensor ops where the
    batch lives (large batches): same results"""
    os.environ["NJODE_INDEX"] = mode
    parity_util.check_training_call(name, "cpu", with_hT_grad=True)
    parity_util.check_path_call(name, "cpu")


def test_index_builders_agree_exactly():
    from njode_b200 import schedule
    batch = cases.grid_batch(57, 1, 30, 0.2, seed=31)
    sched = schedule.build_schedule(batch["times"], 1.0 / 30, 1.0, False, False)
    B = 57
    pp, pr, rj = schedule.build_csr(batch["time_ptr"], batch["obs_idx"].numpy(), B)
    for segments in (True, False):
        units, n_loss = schedule.build_units(sched, pp, pr, rj, B, segments)
        t = schedule.build_index_torch(batch["obs_idx"], torch.tensor(batch["time_ptr"]), torch.tensor(sched.jump_step),
                                       B, sched.S, segments, 8, 4)
        assert np.array_equal(t[0].numpy(), pp) and np.array_equal(t[1].numpy(), pr) and np.array_equal(t[2].numpy(), rj)
        assert np.array_equal(t[3].numpy().reshape(-1, 6), units) and t[4] == n_loss
        if segments:
            lens = units[:, 2] - units[:, 1]
            assert [int(v) for v in t[5][:4]] == [int((lens[:n_loss] >= 8).sum()), int((lens[:n_loss] >= 4).sum()),
                                                   int((lens[n_loss:] >= 8).sum()), int((lens[n_loss:] >= 4).sum())]


@pytest.mark.parametrize("nw", [2, 5])
@pytest.mark.parametrize("name", ["bs_ckpt1", "easy_w07_nores"])
def test_segment_path_small_ctas_spill_dw_tiles_to_the_partial_image(name, nw):
    """few warps per CTA (small batches): dW tiles beyond the register capacity accumulate through global memory"""
    os.environ["NJODE_FORCE_NW"] = str(nw)
    parity_util.check_training_call(name, "cpu", with_hT_grad=True)


def test_segment_path_small_ctas_with_dropout():
    os.environ["NJODE_FORCE_NW"] = "3"
    cfg = cases.demo_cfg(dropout_rate=0.2)
    batch = cases.grid_batch(40, 1, 25, 0.2, seed=16)
    parity_util.check_against_oracle(cfg, batch, 0.04, 1.0, seed=12, device="cpu", train=True, grad_hT=True)


@pytest.mark.parametrize("name", ["bs_ckpt1", "masked_small"])
def test_backward_without_hT_gradient_skips_the_tail_units(name):
    """loss.backward() hands the Function ``None`` for the unused hT (set_materialize_grads(False)): the backward then runs
    on the loss units only (PreparedBatch.bwd_loss) and must give the gradients of the all-units backward fed with zeros"""
    cfg, meta, sd, batch, outs = cases.load_case(name)
    grads = []
    for with_zero_hT in (False, True):
        m = parity_util.build_model(cfg, sd, "cpu")
        m.eval()
        hT, loss = parity_util.call(m, batch, meta, "cpu")
        obj = loss + (hT * 0.0).sum() if with_zero_hT else loss
        obj.backward()
        grads.append({n: p.grad.clone() for n, p in m.named_parameters()})
    for n in grads[0]:
        # other tiles -> other fp32 summation orders, nothing else
        np.testing.assert_allclose(grads[0][n].numpy(), grads[1][n].numpy(), rtol=2e-5, atol=1e-6 * float(grads[1][n].abs().max()))


def test_backward_after_a_parameter_update_raises():
    """the kernels' backward re-reads the parameters: changing them between forward and backward must fail loudly, as
    autograd's saved-tensor version check does for the reference"""
    cfg, meta, sd, batch, outs = cases.load_case("bs_ckpt1")
    m = parity_util.build_model(cfg, sd, "cpu")
    m.eval()
    hT, loss = parity_util.call(m, batch, meta, "cpu")
    with torch.no_grad():
        next(m.parameters()).mul_(1.5)
    with pytest.raises(RuntimeError, match="modified"):
        loss.backward()


# ---- whole-path units on the warp GEMMs (njode_path.cuh): every tile shape the planner can pick ----
STAT = pytest.mark.parametrize("stat", ["warp-gemm", "warp-gemm-pipelined", "weight-stationary", "thread-per-neuron"])


def _pick(stat):
    """small batches take the weight-stationary Euler steps (ODE weights in registers, all warps of a CTA on one tile);
    NJODE_NO_STAT keeps them on the warp-GEMM path kernels that serve the larger batches, whose backward runs the ODE
    network's dW phase on helper warps concurrently with the row warps' next step unless NJODE_NO_PIPE is set"""
    if stat == "thread-per-neuron":
        # (njode_tpn.cuh: tiles of 1 or 4 paths; networks outside its dimension classes fall to the kernels below)
        os.environ["NJODE_FORCE_TPN"] = "1"
        if os.environ.get("NJODE_PATH_R") in ("2", "8"):
            pytest.skip("thread-per-neuron tiles have 1 or 4 rows")
        return
    os.environ["NJODE_NO_TPN"] = "1"
    if stat == "weight-stationary":
        os.environ["NJODE_FORCE_STAT"] = "1"      # (the planner itself takes them for one path per CTA only)
    else:
        os.environ["NJODE_NO_STAT"] = "1"
    if stat == "warp-gemm":
        os.environ["NJODE_NO_PIPE"] = "1"
    if stat == "warp-gemm-pipelined":
        os.environ["NJODE_FORCE_PIPE"] = "1"


@STAT
@pytest.mark.parametrize("rows", [1, 2, 4, 8])
@pytest.mark.parametrize("name", ["masked_small", "gru_demo", "gru_masked", "gru_d3_nores"])
def test_path_kernels_every_tile_shape(name, rows, stat):
    """rows per warp 1 / 2 (split reduction dimension, partial sums meet in shuffles), 4 and 8: same loss, hT, gradients
    (with a gradient flowing into hT) and recorded paths as the reference"""
    os.environ["NJODE_PATH_R"] = str(rows)
    _pick(stat)
    parity_util.check_training_call(name, "cpu", with_hT_grad=True)
    parity_util.check_training_call(name, "cpu")
    parity_util.check_path_call(name, "cpu")


@STAT
@pytest.mark.parametrize("rows", [1, 2, 4, 8])
@pytest.mark.parametrize("name", ["bs_ckpt1", "curt_nobias_relu", "res_case2", "easy_w07_nores"])
def test_path_kernels_record_paths_of_the_non_masked_model(name, rows, stat):
    """return_path / until_T calls of the non-masked model (evaluate, get_pred) are whole-path units too"""
    os.environ["NJODE_PATH_R"] = str(rows)
    _pick(stat)
    parity_util.check_path_call(name, "cpu")


@STAT
@pytest.mark.parametrize("rows", [1, 2, 4, 8])
def test_path_kernels_train_mode_dropout(rows, stat):
    os.environ["NJODE_PATH_R"] = str(rows)
    _pick(stat)
    cfg = dict(cases.CONFIGS["masked_small"], dropout_rate=0.25)
    batch = cases.irregular_batch(11, 5, 10, seed=21, masked=True, times_f32=True)
    parity_util.check_against_oracle(cfg, batch, 0.05, 1 + 1e-12, seed=3, device="cpu", train=True, grad_hT=True)
    cfg = cases.demo_cfg(use_rnn=True, dropout_rate=0.2, bias=False, hidden_size=6)
    batch = cases.grid_batch(13, 1, 16, 0.3, seed=9)
    parity_util.check_against_oracle(cfg, batch, 1.0 / 16, 1.0, seed=4, device="cpu", train=True, grad_hT=True)


def test_path_kernels_physionet_shape():
    """d = H = 41 masked, 2x50 nets, float32 times, several waves of tiles per CTA on the 4-SM simulation"""
    batch = cases.irregular_batch(70, 41, 30, seed=7, masked=True, times_f32=True, obs_at_zero=True, row_prob=0.2, feat_prob=0.12)
    parity_util.check_against_oracle(cases.CONFIGS["masked_physio"], batch, 0.01, 1 + 1e-12, seed=3, device="cpu", grad_hT=True)


@pytest.mark.parametrize("B", [50, 300])
def test_thread_per_neuron_kernels_physionet_shape_with_the_b200_launch_plan(B):
    """the reference's PhysioNet batch of 50 records (one path per CTA: the planner's own choice) and 300 records (tiles of
    4, forced: the planner gives such batches to the pipelined warp kernels) with the launch plan of a 148-SM device:
    dimension class B (84 / 52 / 44), dropout on, gradient into hT"""
    os.environ["NJODE_SIM_SMS"] = "148"
    if B > 148:
        os.environ["NJODE_FORCE_TPN"] = "1"
    batch = cases.irregular_batch(B, 41, 12, seed=17, masked=True, times_f32=True, obs_at_zero=True, row_prob=0.25, feat_prob=0.12)
    cfg = dict(cases.CONFIGS["masked_physio"], dropout_rate=0.2)
    m = models.NJODE(**cfg)
    pb = m.prepare_batch(batch["times"], batch["time_ptr"], batch["X"], batch["obs_idx"], 1.0 / 12, 1 + 1e-12, batch["start_X"], batch["n_obs_ot"], M=batch["M"])
    assert "tpn" in hostsim_util.plan_kind(m, pb, "fwd") and "tpn" in hostsim_util.plan_kind(m, pb, "bwd_all")
    parity_util.check_against_oracle(cfg, batch, 1.0 / 12, 1 + 1e-12, seed=5, device="cpu", train=True, grad_hT=True)
    parity_util.check_against_oracle(cfg, batch, 1.0 / 12, 1 + 1e-12, seed=5, device="cpu", train=True)


@pytest.mark.parametrize("B", [50, 300])
def test_weight_stationary_kernels_physionet_shape_with_the_b200_launch_plan(B):
    """the reference's PhysioNet batch of 50 records (one path per CTA) and 300 records (tiles of 4 rows... on 148 SMs:
    2 rows per CTA) with the launch plan of a 148-SM device: d = H = 41 masked, 2x50 nets -> 13 warps per CTA"""
    os.environ["NJODE_SIM_SMS"] = "148"
    os.environ["NJODE_NO_TPN"] = "1"
    os.environ["NJODE_FORCE_STAT"] = "1"
    batch = cases.irregular_batch(B, 41, 12, seed=17, masked=True, times_f32=True, obs_at_zero=True, row_prob=0.25, feat_prob=0.12)
    cfg = dict(cases.CONFIGS["masked_physio"], dropout_rate=0.2)
    parity_util.check_against_oracle(cfg, batch, 1.0 / 12, 1 + 1e-12, seed=5, device="cpu", train=True, grad_hT=True)


# ---- segment backward in recompute mode: nothing saved by the forward pass (north_star: "the backward pass recomputes
# forward segments from checkpointed h at observation times rather than storing every step") ----
@pytest.fixture
def recompute_on():
    os.environ["NJODE_RECOMPUTE"] = "on"
    yield
    os.environ.pop("NJODE_RECOMPUTE", None)
    os.environ.pop("NJODE_FORCE_TR", None)


@pytest.mark.parametrize("name", ["bs_ckpt1", "heston_ckpt2", "ou_ckpt3", "irregular_demo", "curt_nobias_relu", "res_case2", "easy_w07_nores"])
def test_segment_backward_recomputes_from_checkpoints(name, recompute_on):
    parity_util.check_training_call(name, "cpu", with_hT_grad=True)
    parity_util.check_training_call(name, "cpu")


@pytest.mark.parametrize("tr", [1, 2])
def test_segment_backward_recompute_train_mode_dropout(tr, recompute_on):
    """the recomputed forward replays the forward kernel's dropout masks (same counter-based keys)"""
    os.environ["NJODE_FORCE_TR"] = str(tr)
    cfg = cases.demo_cfg(dropout_rate=0.15, input_size=2, output_size=2)
    batch = cases.grid_batch(150, 2, 30, 0.2, seed=12)
    parity_util.check_against_oracle(cfg, batch, 1.0 / 30, 1.0, seed=6, device="cpu", train=True, grad_hT=True)


def test_recompute_mode_saves_nothing(recompute_on):
    from njode_b200 import _ext
    cfg, meta, sd, batch, outs = cases.load_case("bs_ckpt1")
    m = parity_util.build_model(cfg, sd, "cpu")
    assert m.recompute == "on"
    m.eval()
    hT, loss = parity_util.call(m, batch, meta, "cpu")
    assert loss.grad_fn is not None and loss.grad_fn.saved == ()


# ---- segment units of small batches on the thread-per-neuron kernels (njode_tpn.cuh, nj_segtpn_*) ----
SEG_TPN_NAMES = ["bs_ckpt1", "heston_ckpt2", "ou_ckpt3", "irregular_demo", "res_case2", "easy_w07_nores"]


@pytest.mark.parametrize("name", SEG_NAMES)
def test_segment_thread_per_neuron_kernels(name):
    """every non-masked golden case whose ODE network fits a dimension class (the others keep the warp kernels)"""
    os.environ["NJODE_SEG_TPN"] = "1"
    parity_util.check_training_call(name, "cpu", with_hT_grad=True)
    parity_util.check_training_call(name, "cpu")


@pytest.mark.parametrize("d", [1, 2])
def test_segment_thread_per_neuron_kernels_train_mode_dropout(d):
    """the demo networks (class A), several CTAs, dropout masks replayed; with and without a gradient into hT"""
    os.environ["NJODE_SEG_TPN"] = "1"
    cfg = cases.demo_cfg(dropout_rate=0.2, input_size=d, output_size=d)
    batch = cases.grid_batch(40, d, 25, 0.2, seed=16)
    m = models.NJODE(**cfg)
    pb = m.prepare_batch(batch["times"], batch["time_ptr"], batch["X"], batch["obs_idx"], 0.04, 1.0, batch["start_X"], batch["n_obs_ot"])
    assert "segtpn" in hostsim_util.plan_kind(m, pb, "fwd") and "segtpn" in hostsim_util.plan_kind(m, pb, "bwd_loss")
    parity_util.check_against_oracle(cfg, batch, 0.04, 1.0, seed=12, device="cpu", train=True, grad_hT=True)
    parity_util.check_against_oracle(cfg, batch, 0.04, 1.0, seed=12, device="cpu", train=True, grad_hT=False)


def test_segment_thread_per_neuron_kernels_class_b():
    """d = 20, H = 40 non-masked: the wider dimension class (84 / 52 / 44)"""
    os.environ["NJODE_SEG_TPN"] = "1"
    cfg = cases.demo_cfg(dropout_rate=0.1, input_size=20, output_size=20, hidden_size=40)
    batch = cases.grid_batch(30, 20, 12, 0.3, seed=19)
    parity_util.check_against_oracle(cfg, batch, 1.0 / 12, 1.0, seed=13, device="cpu", train=True, grad_hT=True)


def test_segment_thread_per_neuron_kernels_recompute(recompute_on):
    os.environ["NJODE_SEG_TPN"] = "1"
    cfg = cases.demo_cfg(dropout_rate=0.15)
    batch = cases.grid_batch(60, 1, 30, 0.2, seed=12)
    parity_util.check_against_oracle(cfg, batch, 1.0 / 30, 1.0, seed=6, device="cpu", train=True, grad_hT=True)


def test_planner_gives_the_reference_batch_to_the_thread_per_neuron_kernels():
    """B200 launch plan (148 SMs): the reference's own batch of 200 paths x 100 steps (~2 200 segments) takes the
    thread-per-neuron kernels, a batch of 20 000 paths the 12-warp tile kernels; ODE networks outside the dimension
    classes (2 x 100) never do"""
    os.environ.pop("NJODE_SEG_TPN", None)
    os.environ["NJODE_SIM_SMS"] = "148"
    for B, layers, want in ((200, [[50, "tanh"]] * 2, True), (20000, [[50, "tanh"]] * 2, False), (200, [[100, "tanh"]] * 2, False)):
        cfg = cases.demo_cfg(ode_nn=layers)
        m = models.NJODE(**cfg)
        batch = cases.grid_batch(B, 1, 100, 0.1, seed=3)
        pb = m.prepare_batch(batch["times"], batch["time_ptr"], batch["X"], batch["obs_idx"], 0.01, 1.0, batch["start_X"], batch["n_obs_ot"])
        for which in ("fwd", "bwd_all", "bwd_loss"):
            kind = hostsim_util.plan_kind(m, pb, which)
            assert "seg" in kind and ("segtpn" in kind) == want, (B, layers, which, kind)


# ---- saved hidden activations (njode_plan_t.act_bytes / njode_saved_t.act_hist): the default; without them the segment
# backward recomputes the hidden layers of every step ----
@pytest.mark.parametrize("tr", [1, 2])
@pytest.mark.parametrize("name", ["bs_ckpt1", "heston_ckpt2", "curt_nobias_relu", "easy_w07_nores"])
def test_segment_backward_recomputes_hidden_layers_without_saved_activations(name, tr):
    os.environ["NJODE_SAVE_ACTIVATIONS"] = "0"
    os.environ["NJODE_FORCE_TR"] = str(tr)
    parity_util.check_training_call(name, "cpu", with_hT_grad=True)
    parity_util.check_training_call(name, "cpu")


def test_saved_activations_train_mode_dropout_marks_survive():
    """the saved records carry the dropout marks (-0.0f) the backward needs; one- and two-hidden-layer ODE networks"""
    for layers in (1, 2):
        cfg = cases.demo_cfg(dropout_rate=0.25, input_size=2, output_size=2, ode_nn=[[50, "tanh"]] * layers)
        batch = cases.grid_batch(40, 2, 25, 0.2, seed=16)
        for save in ("1", "0"):
            os.environ["NJODE_SAVE_ACTIVATIONS"] = save
            parity_util.check_against_oracle(cfg, batch, 0.04, 1.0, seed=12, device="cpu", train=True, grad_hT=True)


def test_planner_kernel_families_for_whole_path_batches():
    """B200 launch plan (measured thresholds, profiles/r2af_*): up to 5 paths per SM -> thread per neuron, one path per tile,
    forward and backward; up to 16 -> thread per neuron backward, warp-GEMM forward; more -> the pipelined warp kernels;
    ODE networks outside the dimension classes: K-split stationary kernels up to one path per SM"""
    os.environ["NJODE_SIM_SMS"] = "148"
    for B, layers, want in ((50, 2, {"tpn", "tpn_fwd"}), (600, 2, {"tpn", "tpn_fwd"}), (2000, 2, {"tpn"}), (3000, 2, {"pipe"}),
                            (50, 1, {"pathstat"})):
        cfg = dict(cases.CONFIGS["masked_physio"], ode_nn=[[50, "tanh"]] * layers)
        m = models.NJODE(**cfg)
        batch = cases.irregular_batch(B, 41, 6, seed=17, masked=True, times_f32=True, obs_at_zero=True, row_prob=0.25, feat_prob=0.12)
        pb = m.prepare_batch(batch["times"], batch["time_ptr"], batch["X"], batch["obs_idx"], 1.0 / 6, 1 + 1e-12, batch["start_X"],
                             batch["n_obs_ot"], M=batch["M"])
        kind = hostsim_util.plan_kind(m, pb, "bwd_all")
        assert "path" in kind and kind & {"tpn", "tpn_fwd", "pipe", "pathstat"} == want, (B, layers, kind)


def test_thread_per_neuron_backward_after_a_warp_gemm_forward():
    """mid-size batches: the forward runs on the warp kernels, the backward on the thread-per-neuron kernels (same saved
    history); several one-path tiles per CTA"""
    os.environ["NJODE_TPN_WAVES_FWD"] = "0"
    os.environ["NJODE_TPN_WAVES"] = "16"
    try:
        for name in ("masked_small", "gru_demo"):
            parity_util.check_training_call(name, "cpu", with_hT_grad=True)
        cfg = dict(cases.CONFIGS["masked_physio"], dropout_rate=0.2)
        batch = cases.irregular_batch(30, 41, 12, seed=17, masked=True, times_f32=True, obs_at_zero=True, row_prob=0.25, feat_prob=0.12)
        parity_util.check_against_oracle(cfg, batch, 1.0 / 12, 1 + 1e-12, seed=5, device="cpu", train=True, grad_hT=True)
    finally:
        os.environ.pop("NJODE_TPN_WAVES_FWD", None)
        os.environ.pop("NJODE_TPN_WAVES", None)
