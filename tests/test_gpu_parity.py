"""GPU parity tests proper: njode_b200.models.NJODE on cuda:0 (libnjode_b200.so through the C ABI)
against the committed outputs of the real reference and against the oracle on fresh seeded inputs.
Tolerance: rtol 1e-4 (fp32) as stated in BASELINE.json; path_t / schedule handling exact."""
import os

import numpy as np
import pytest
import torch

import cases
import hostsim_util
import parity_util
import oracle.njode_oracle as orc
from njode_b200 import models

pytestmark = pytest.mark.gpu
NAMES = cases.golden_names()
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def no_test_runner():
    hostsim_util.uninstall()
    yield
    os.environ.pop("NJODE_FORCE_TILE", None)
    os.environ.pop("NJODE_SEG_TPN", None)
    os.environ.pop("NJODE_FORCE_TR", None)
    os.environ.pop("NJODE_NO_TPN", None)
    os.environ.pop("NJODE_NO_STAT", None)
    os.environ.pop("NJODE_FORCE_TPN", None)
    os.environ.pop("NJODE_FORCE_STAT", None)


@pytest.mark.parametrize("name", NAMES)
def test_training_call(name):
    parity_util.check_training_call(name, DEV)


@pytest.mark.parametrize("name", NAMES)
def test_training_call_with_hT_gradient(name):
    parity_util.check_training_call(name, DEV, with_hT_grad=True)


@pytest.mark.parametrize("name", NAMES)
def test_path_call(name):
    parity_util.check_path_call(name, DEV)


@pytest.mark.parametrize("tile", [8, 16, 32, 64])
@pytest.mark.parametrize("name", ["bs_ckpt1", "masked_small", "gru_demo"])
def test_tile_size_invariance(name, tile):
    os.environ["NJODE_FORCE_TILE"] = str(tile)
    parity_util.check_training_call(name, DEV, with_hT_grad=True)
    parity_util.check_path_call(name, DEV)


def _last_kernels():
    import ctypes as C
    from njode_b200 import _ext
    dll = _ext.cuda_lib().dll
    dll.njode_last_kernel.argtypes, dll.njode_last_kernel.restype = [C.c_int], C.c_char_p
    return dll.njode_last_kernel(0).decode(), dll.njode_last_kernel(1).decode()


def _oracle_vs_cuda(cfg, batch, dt, T, seed, train=False, rtol=1e-4):
    parity_util.check_against_oracle(cfg, batch, dt, T, seed, DEV, train=train, rtol=rtol)


def test_demo_batch_200_against_oracle():
    """BASELINE config 1 shape: d=1, H=10, 2x50, 100 steps, B=200, obs_perc 0.1"""
    batch = cases.grid_batch(200, 1, 100, 0.1, seed=5)
    _oracle_vs_cuda(cases.CONFIGS["demo"], batch, 0.01, 1.0, seed=1)


def test_demo_batch_train_mode_dropout_against_oracle():
    """train mode: the oracle replays the device's counter-based keep-masks"""
    cfg = cases.demo_cfg(dropout_rate=0.1)
    batch = cases.grid_batch(64, 1, 50, 0.15, seed=6)
    _oracle_vs_cuda(cfg, batch, 0.02, 1.0, seed=2, train=True)


def test_masked_physionet_shape_against_oracle():
    batch = cases.irregular_batch(12, 41, 40, seed=7, masked=True, times_f32=True, obs_at_zero=True,
                                  row_prob=0.2, feat_prob=0.12)
    _oracle_vs_cuda(cases.CONFIGS["masked_physio"], batch, 0.01, 1 + 1e-12, seed=3)


def test_wide_model_global_weights_against_oracle():
    """weights too large for shared memory -> global-memory parameter image path"""
    cfg = cases.demo_cfg(input_size=4, output_size=4, hidden_size=128,
                         ode_nn=[[256, "tanh"], [256, "tanh"]], enc_nn=[[256, "tanh"]],
                         readout_nn=[[256, "tanh"]])
    batch = cases.grid_batch(48, 4, 12, 0.3, seed=8)
    _oracle_vs_cuda(cfg, batch, 1.0 / 12, 1.0, seed=4)


def test_config3_combined_2x100_nets_against_oracle():
    """BASELINE config 3 (i): 2x100 tanh nets, H=10 (parallel_train.py:609-632) on a regular grid"""
    nn100 = [[100, "tanh"], [100, "tanh"]]
    cfg = cases.demo_cfg(ode_nn=nn100, enc_nn=nn100, readout_nn=nn100)
    batch = cases.grid_batch(96, 1, 40, 0.1, seed=21)
    _oracle_vs_cuda(cfg, batch, 1.0 / 40, 1.0, seed=5)
    _oracle_vs_cuda(dict(cfg, dropout_rate=0.1), batch, 1.0 / 40, 1.0, seed=5, train=True)


@pytest.mark.parametrize("helpers", ["0", "1"])
def test_segment_backward_helper_warps(helpers):
    """the backward instantiation with dW helper warps (nj_seg_bwd_kernel_h) and the plain one give the oracle's
    gradients whichever the planner would pick: demo nets at a batch whose CTAs have 4 row warps, dropout on"""
    os.environ["NJODE_SEG_HELPERS"] = helpers
    try:
        cfg = cases.demo_cfg(dropout_rate=0.1)
        batch = cases.grid_batch(300, 1, 40, 0.15, seed=27)
        parity_util.check_against_oracle(cfg, batch, 1.0 / 40, 1.0, seed=10, device=DEV, train=True, grad_hT=True)
    finally:
        os.environ.pop("NJODE_SEG_HELPERS", None)


def test_config3_heston_wo_feller_two_coordinates_against_oracle():
    """BASELINE config 3 (ii): HestonWOFeller with return_vol -> input_size 2, 2x50 nets"""
    cfg = cases.demo_cfg(input_size=2, output_size=2)
    batch = cases.grid_batch(300, 2, 50, 0.1, seed=22)
    _oracle_vs_cuda(cfg, batch, 0.02, 1.0, seed=6)
    _oracle_vs_cuda(dict(cfg, dropout_rate=0.1), batch, 0.02, 1.0, seed=6, train=True)


def test_config5_scaled_architecture_against_oracle():
    """BASELINE config 5 architecture (d=16, H=256, 4x256 tanh) at a size the oracle finishes in seconds"""
    nn = [[256, "tanh"]] * 4
    cfg = cases.demo_cfg(input_size=16, output_size=16, hidden_size=256, ode_nn=nn, enc_nn=nn, readout_nn=nn)
    batch = cases.grid_batch(24, 16, 8, 0.25, seed=23)
    _oracle_vs_cuda(cfg, batch, 0.125, 1.0, seed=7)


def test_gru_jump_train_mode_dropout_against_oracle():
    """use_rnn=True (GRU jump, NJODE/models.py:202-217) on the demo nets, dropout on, batch 200"""
    cfg = cases.demo_cfg(use_rnn=True, dropout_rate=0.1)
    batch = cases.grid_batch(200, 1, 50, 0.1, seed=25)
    parity_util.check_against_oracle(cfg, batch, 0.02, 1.0, seed=9, device=DEV, train=True, grad_hT=True)
    parity_util.check_against_oracle(cfg, batch, 0.02, 1.0, seed=9, device=DEV, train=False)


@pytest.mark.parametrize("B", [200, 5000])
def test_demo_batch_sweep_train_mode(B):
    """batch sweep of config 3 on the demo nets, dropout on: segment fast path with every tile height"""
    cfg = cases.demo_cfg(dropout_rate=0.1)
    batch = cases.grid_batch(B, 1, 100, 0.1, seed=24)
    _oracle_vs_cuda(cfg, batch, 0.01, 1.0, seed=8, train=True)


def test_empty_batch_edges():
    """no observation rows at all / zero Euler steps"""
    cfg = cases.CONFIGS["demo"]
    sd = orc.init_state_dict(orc.Config(**cfg), seed=9)
    m = parity_util.build_model(cfg, sd, DEV).eval()
    B = 5
    batch = {"times": np.zeros(0), "time_ptr": np.array([0]), "X": torch.zeros(0, 1),
             "obs_idx": torch.zeros(0, dtype=torch.long), "start_X": torch.ones(B, 1),
             "n_obs_ot": torch.zeros(B, dtype=torch.long)}
    with torch.no_grad():
        hT, loss = parity_util.call(m, batch, {"delta_t": 0.1, "T": 1.0}, DEV, until_T=True)
        o_hT, o_loss = orc.forward(orc.Config(**cfg), sd, batch["times"], batch["time_ptr"], batch["X"],
                                   batch["obs_idx"], 0.1, 1.0, batch["start_X"], batch["n_obs_ot"], until_T=True)
    assert float(loss) == 0.0
    assert parity_util.rel_err(hT.cpu().numpy(), o_hT.numpy()) < 1e-4


def test_native_library_is_loaded():
    from njode_b200 import _ext
    with open("/proc/self/maps") as f:
        assert "libnjode_b200.so" in f.read()
    assert _ext.cuda_lib().dll.njode_abi_version() == 6


@pytest.mark.parametrize("width", [50, 200])
def test_config4_physionet_full_size_against_oracle(width):
    """BASELINE config 4 at its FULL size: the reference's PhysioNet batch of 50 records
    (parallel_train.py:656), d = H = 41 masked, ~3 700 dependent Euler steps on the 3000-tick grid with float32 times
    (physionet_train.py:192-193), 2x50 and 2x200 tanh nets -- where fp32 error has thousands of steps to accumulate.
    Element-wise rtol 1e-4 + the measured fp32 noise floor (parity_util.noise_floor); the per-output errors go to
    gpurun_out/parity_physionet_b50_<width>.json."""
    import json
    import bench
    wl = dict(bench.WORKLOADS["physionet_synth_b50"], width=width)
    batch, dt = bench.synth_batch_physio(wl, 1234, 0, wl["paths"])
    cfg = dict(bench.model_cfg(wl), dropout_rate=0.0)
    report = {}
    try:
        parity_util.check_against_oracle(cfg, batch, dt, bench.horizon(wl), seed=11, device=DEV, report=report)
    finally:
        out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_physionet_b50_%d.json" % width), "w") as f:
            json.dump(report, f, indent=1)


# ---- segment backward in recompute mode (north_star: "recomputes forward segments from checkpointed h at observation
# times rather than storing every step"): nothing is saved by the forward pass ----
@pytest.fixture
def recompute_on():
    os.environ["NJODE_RECOMPUTE"] = "on"
    yield
    os.environ.pop("NJODE_RECOMPUTE", None)


@pytest.mark.parametrize("name", ["bs_ckpt1", "heston_ckpt2", "ou_ckpt3", "irregular_demo", "curt_nobias_relu", "res_case2", "easy_w07_nores"])
def test_recompute_mode_golden_cases(name, recompute_on):
    parity_util.check_training_call(name, DEV, with_hT_grad=True)
    parity_util.check_training_call(name, DEV)


def test_recompute_mode_train_dropout_and_big_nets(recompute_on):
    cfg = cases.demo_cfg(dropout_rate=0.1)
    batch = cases.grid_batch(600, 1, 60, 0.12, seed=31)
    parity_util.check_against_oracle(cfg, batch, 1.0 / 60, 1.0, seed=12, device=DEV, train=True, grad_hT=True)
    nn100 = [[100, "tanh"], [100, "tanh"]]
    cfg = cases.demo_cfg(ode_nn=nn100, enc_nn=nn100, readout_nn=nn100, dropout_rate=0.1)
    batch = cases.grid_batch(96, 1, 40, 0.1, seed=21)
    parity_util.check_against_oracle(cfg, batch, 1.0 / 40, 1.0, seed=5, device=DEV, train=True)


def test_recompute_mode_allocates_no_history(recompute_on):
    """20 000 paths x 100 steps: the default mode keeps 80 MB of h history per step of training, recompute mode none"""
    batch = cases.grid_batch(20000, 1, 100, 0.1, seed=41)
    peaks = {}
    for mode in ("off", "on"):
        torch.manual_seed(0)
        m = models.NJODE(**cases.demo_cfg()).to(DEV)
        m.recompute = mode
        m.output_device = "cuda"
        pb = m.prepare_batch(batch["times"], batch["time_ptr"], batch["X"], batch["obs_idx"], 0.01, 1.0, batch["start_X"], batch["n_obs_ot"])
        hT, loss = m.forward_prepared(pb)
        loss.backward()                            # warm-up: workspaces exist
        torch.cuda.synchronize()
        del hT, loss                               # the warm-up graph (and what it kept) is gone before the baseline is read
        for p in m.parameters():
            p.grad = None
        torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        hT, loss = m.forward_prepared(pb)
        held = torch.cuda.memory_allocated() - base           # what the graph keeps alive between forward and backward
        loss.backward()
        torch.cuda.synchronize()
        peaks[mode] = held
        del hT, loss
    assert peaks["off"] >= 4 * 100 * 20000 * 10, peaks    # [S, B, H] fp32
    assert peaks["on"] < 2 * 1024 * 1024, peaks           # hT + scalars only


# ---- small segment batches: the planner gives them to the thread-per-neuron kernels (nj_segtpn_*); the 12-warp tile
# kernels of big batches are kept covered on the same small cases (NJODE_SEG_TPN=0) ----
SEG_NAMES = [n for n in NAMES if "masked" not in n and "gru" not in n]


@pytest.mark.parametrize("tpn", ["0", "1"])
@pytest.mark.parametrize("name", SEG_NAMES)
def test_segment_units_both_kernel_families(name, tpn):
    os.environ["NJODE_SEG_TPN"] = tpn
    parity_util.check_training_call(name, DEV, with_hT_grad=True)
    parity_util.check_training_call(name, DEV)


@pytest.mark.parametrize("B", [40, 200, 1500])
def test_segment_thread_per_neuron_kernels_train_mode(B):
    """the reference's batch of 200 (and a smaller / larger one) in train mode, with and without a gradient into hT"""
    os.environ["NJODE_SEG_TPN"] = "1"
    cfg = cases.demo_cfg(dropout_rate=0.1)
    batch = cases.grid_batch(B, 1, 100, 0.1, seed=27)
    parity_util.check_against_oracle(cfg, batch, 0.01, 1.0, seed=8, device=DEV, train=True, grad_hT=True)
    parity_util.check_against_oracle(cfg, batch, 0.01, 1.0, seed=8, device=DEV, train=True)
    kf, kb = _last_kernels()
    assert "segtpn" in kf and "segtpn" in kb, (kf, kb)


def test_segment_thread_per_neuron_kernels_class_b_and_recompute(recompute_on):
    os.environ["NJODE_SEG_TPN"] = "1"
    cfg = cases.demo_cfg(dropout_rate=0.1, input_size=20, output_size=20, hidden_size=40)
    batch = cases.grid_batch(64, 20, 40, 0.2, seed=28)
    parity_util.check_against_oracle(cfg, batch, 0.025, 1.0, seed=9, device=DEV, train=True, grad_hT=True)
    cfg = cases.demo_cfg(dropout_rate=0.1)
    batch = cases.grid_batch(300, 1, 100, 0.1, seed=28)
    parity_util.check_against_oracle(cfg, batch, 0.01, 1.0, seed=9, device=DEV, train=True, grad_hT=True)


# ---- small whole-path batches: thread-per-neuron kernels (njode_tpn.cuh), next to the kernels they replace ----
@pytest.mark.parametrize("family", ["tpn", "stat", "warp"])
@pytest.mark.parametrize("B", [50, 300])
def test_small_physionet_batches_every_kernel_family(B, family):
    """the reference's PhysioNet batch (50 records; 300: tiles of 4) in train mode, d = H = 41 masked, 2x50 nets, dropout
    0.2, gradient into hT: thread-per-neuron (the planner's choice), K-split weight-stationary, warp GEMMs"""
    if family != "tpn":
        os.environ["NJODE_NO_TPN"] = "1"
    if family == "warp":
        os.environ["NJODE_NO_STAT"] = "1"
    if B > 148:                        # (more than one path per SM: the planner's own choice is the pipelined warp kernels)
        os.environ["NJODE_FORCE_TPN" if family == "tpn" else "NJODE_FORCE_STAT"] = "1"
    batch = cases.irregular_batch(B, 41, 60, seed=17, masked=True, times_f32=True, obs_at_zero=True, row_prob=0.2, feat_prob=0.12)
    cfg = dict(cases.CONFIGS["masked_physio"], dropout_rate=0.2)
    parity_util.check_against_oracle(cfg, batch, 1.0 / 60, 1 + 1e-12, seed=5, device=DEV, train=True, grad_hT=True)
    parity_util.check_against_oracle(cfg, batch, 1.0 / 60, 1 + 1e-12, seed=5, device=DEV, train=True)
    kf, kb = _last_kernels()
    want = {"tpn": "nj_tpn_", "stat": "nj_stat_", "warp": "nj_path_"}[family]
    assert want in kf and want in kb, (kf, kb)


@pytest.mark.parametrize("B", [60, 400])
def test_thread_per_neuron_kernels_demo_nets_gru_jump(B):
    """dimension class A (d = 1, H = 10, 2x50 nets) on whole-path units: the GRU-jump variant of the demo model"""
    if B > 148:
        os.environ["NJODE_FORCE_TPN"] = "1"
    cfg = cases.demo_cfg(use_rnn=True, dropout_rate=0.1)
    batch = cases.grid_batch(B, 1, 100, 0.1, seed=31)
    parity_util.check_against_oracle(cfg, batch, 0.01, 1.0, seed=6, device=DEV, train=True, grad_hT=True)
    kf, kb = _last_kernels()
    assert "nj_tpn_" in kf and "nj_tpn_" in kb, (kf, kb)


def test_thread_per_neuron_kernels_record_paths():
    """return_path call (evaluation / plotting) of the non-masked demo model: a readout record after every step"""
    for name in ("bs_ckpt1", "masked_small"):
        parity_util.check_path_call(name, DEV)
