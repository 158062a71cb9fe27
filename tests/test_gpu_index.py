"""njode_build_index (njode_b200/csrc/njode_index.cu) against the NumPy builder (njode_b200/schedule.py::build_csr +
build_units, itself exercised against the reference through every parity test): bit-exact index arrays, unit order and
tile-class statistics; error flags for duplicate (time, path) rows and out-of-range path indices."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import cases
from njode_b200 import _ext, schedule, models

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def native_index(obs, time_ptr, jump_step, B, S, segments, T1, T2):
    lib = _ext.cuda_lib()
    dev = torch.device(DEV)
    N, K = len(obs), len(time_ptr) - 1
    t = lambda a: torch.tensor(np.asarray(a, dtype=np.int32), device=dev)
    obs_t, tp_t, js_t = t(obs), t(time_ptr), t(jump_step)
    n_u = (N + B) if segments else B
    path_ptr = torch.full((B + 1,), -7, dtype=torch.int32, device=dev)
    path_rows = torch.full((max(N, 1),), -7, dtype=torch.int32, device=dev)
    row_jump = torch.full((max(N, 1),), -7, dtype=torch.int32, device=dev)
    units = torch.full((n_u * 6,), -7, dtype=torch.int32, device=dev)
    stats = torch.full((6,), -7, dtype=torch.int32, device=dev)
    wsb = int(lib.dll.njode_index_workspace_bytes(N, B))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    p = lambda x: C.c_void_p(x.data_ptr())
    rc = lib.dll.njode_build_index(p(obs_t), N, p(tp_t), K, p(js_t), B, S, int(segments), T1, T2, p(path_ptr), p(path_rows),
                                   p(row_jump), p(units), p(stats), p(ws), wsb,
                                   C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    lib.check(rc, "njode_build_index")
    torch.cuda.synchronize()
    return (path_ptr.cpu().numpy(), path_rows.cpu().numpy()[:N], row_jump.cpu().numpy()[:N],
            units.cpu().numpy().reshape(-1, 6), stats.cpu().numpy())


def check(batch, dt, T, B, T1=8, T2=4):
    sched = schedule.build_schedule(batch["times"], dt, T, False, False)
    obs = batch["obs_idx"].numpy()
    pp, pr, rj = schedule.build_csr(np.asarray(batch["time_ptr"]), obs, B)
    for segments in (True, False):
        units, n_loss = schedule.build_units(sched, pp, pr, rj, B, segments)
        g = native_index(obs, batch["time_ptr"], sched.jump_step, B, sched.S, segments, T1, T2)
        assert np.array_equal(g[0], pp) and np.array_equal(g[1], pr) and np.array_equal(g[2], rj)
        assert np.array_equal(g[3], units)
        lens = units[:, 2] - units[:, 1]
        want = [0, 0, 0, 0]
        if segments:
            want = [int((lens[:n_loss] >= T1).sum()), int((lens[:n_loss] >= T2).sum()),
                    int((lens[n_loss:] >= T1).sum()), int((lens[n_loss:] >= T2).sum())]
        assert list(g[4][:4]) == want and g[4][4] == 0 and g[4][5] == 0


@pytest.mark.parametrize("B,steps,obs_perc", [(57, 30, 0.2), (200, 100, 0.1), (5000, 100, 0.1), (3, 10, 0.9), (1, 7, 0.5)])
def test_grid_batches(B, steps, obs_perc):
    batch = cases.grid_batch(B, 1, steps, obs_perc, seed=31 + B)
    check(batch, 1.0 / steps, 1.0, B)


def test_irregular_batch_with_empty_slot_and_unobserved_path():
    batch = cases.irregular_batch(40, 3, 25, seed=5)
    check(batch, 0.02, 1.0, 40)
    batch = cases.irregular_batch(300, 41, 120, seed=6, masked=True, times_f32=True, obs_at_zero=True, row_prob=0.1)
    check(batch, 0.016 / 48 * 30, 1 + 1e-12, 300, T1=20, T2=6)


def test_no_rows_at_all():
    batch = {"times": np.zeros(0), "time_ptr": np.array([0]), "obs_idx": torch.zeros(0, dtype=torch.long)}
    check(batch, 0.1, 1.0, 5)


def test_error_flags():
    # two rows of path 1 at the same observation time; path index out of range
    g = native_index([0, 1, 1, 2], [0, 4], [3], 3, 5, True, 8, 4)
    assert g[4][4] == 1 and g[4][5] == 0
    g = native_index([0, 7, 2], [0, 3], [3], 3, 5, False, 8, 4)
    assert g[4][5] == 1


def test_model_call_uses_the_native_builder_and_matches_the_host_builder():
    """same loss / gradients whichever builder staged the batch (NJODE_INDEX=host: NumPy; device: njode_build_index)"""
    cfg = cases.demo_cfg(dropout_rate=0.0)
    batch = cases.grid_batch(300, 1, 40, 0.15, seed=77)
    res = {}
    for mode in ("host", "device"):
        os.environ["NJODE_INDEX"] = mode
        try:
            torch.manual_seed(3)
            m = models.NJODE(**cfg).to(DEV).eval()
            hT, loss = m(batch["times"], batch["time_ptr"], batch["X"], batch["obs_idx"], 1.0 / 40, 1.0, batch["start_X"],
                         batch["n_obs_ot"])
            loss.backward()
            res[mode] = (float(loss.detach()), hT.detach().cpu().numpy(), torch.cat([p.grad.reshape(-1) for p in m.parameters()]).cpu().numpy())
        finally:
            os.environ.pop("NJODE_INDEX", None)
    # identical index arrays -> identical per-unit arithmetic; the gradient partials are summed per CTA and the tiles are
    # handed out dynamically, so gradients agree to rounding, not bitwise
    assert abs(res["host"][0] - res["device"][0]) <= 1e-6 * abs(res["host"][0])
    np.testing.assert_allclose(res["host"][1], res["device"][1], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(res["host"][2], res["device"][2], rtol=1e-4, atol=1e-6 * np.abs(res["host"][2]).max())
