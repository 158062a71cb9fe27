"""CPU-only: pins oracle/sde_oracle.py (update rules of the Euler-Maruyama generators, the collate)
against outputs of the REAL reference committed in tests/golden/sde_ref.npz (made by
tests/golden/make_sde_golden.py), checks the host collate of njode_b200.data_utils against it, and
checks that libnjode_b200.so exports every symbol include/njode_b200.h declares."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

import _reference
from oracle import sde_oracle as so
from njode_b200 import data_utils

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "sde_ref.npz"))
HP = json.loads(str(GOLD["hp"]))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["BlackScholes", "OrnsteinUhlenbeck", "Heston", "HestonWOFeller", "HestonWOFeller_vol"])
def test_update_rules_bit_exact_against_reference_fixture(name):
    h = dict(HP, return_vol=name.endswith("_vol"))
    p = so.euler_paths(name.split("_")[0], h, GOLD[name + "/n1"], GOLD[name + "/n2"])
    assert np.array_equal(p, GOLD[name + "/paths"])                  # integer-exact float64 equality


def test_philox_known_answer():
    """Random123 known-answer vectors for Philox-4x32-10"""
    x = so.philox4x32_10(np.uint32(0), np.uint32(0), np.uint32(0), np.uint32(0), 0, 0)
    assert [int(v) for v in x] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    x = so.philox4x32_10(np.uint32(0xffffffff), np.uint32(0xffffffff), np.uint32(0xffffffff), np.uint32(0xffffffff),
                         0xffffffff, 0xffffffff)
    assert [int(v) for v in x] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    x = so.philox4x32_10(np.uint32(0x243f6a88), np.uint32(0x85a308d3), np.uint32(0x13198a2e), np.uint32(0x03707344),
                         0xa4093822, 0x299f31d0)
    assert [int(v) for v in x] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_mask_and_normals_are_sharding_invariant():
    a1, a2 = so.philox_normals(5, np.arange(0, 40), 12, 2)
    b1, b2 = so.philox_normals(5, np.arange(17, 29), 12, 2)
    assert np.array_equal(a1[17:29], b1) and np.array_equal(a2[17:29], b2)
    m = so.philox_mask(5, np.arange(0, 40), 30, 0.3)
    assert np.array_equal(m[17:29], so.philox_mask(5, np.arange(17, 29), 30, 0.3))
    assert 0.2 < m[:, 1:].mean() < 0.4 and 0.1 < m[:, 0].mean() < 0.5


def test_normals_are_standard():
    n1, n2 = so.philox_normals(1, np.arange(4000), 50, 1)
    for n in (n1, n2):
        assert abs(n.mean()) < 0.01 and abs(n.var() - 1) < 0.02
    assert abs(np.mean(n1 * n2)) < 0.01


def _batch(seed=0, B=23, d=2, steps=15, p=0.3):
    rng = np.random.default_rng(seed)
    paths = rng.random((B, d, steps + 1)) + 0.5
    obs = (rng.random((B, steps + 1)) < p).astype(np.int64)
    obs[:, 0] = 1
    obs[:, 7] = 0                      # a grid time nobody observes -> absent from `times`
    obs[3, 1:] = 0                     # a path without observations
    return paths, obs, obs[:, 1:].sum(1), 1.0 / steps


def test_host_collate_matches_oracle_loops():
    paths, obs, nb, dt = _batch()
    a = so.collate(paths, obs, nb, dt)
    b = data_utils.collate_paths(paths, obs, nb, dt)
    assert np.array_equal(a["times"], b["times"]) and np.array_equal(a["time_ptr"], b["time_ptr"])
    assert np.array_equal(a["obs_idx"], b["obs_idx"].numpy())
    assert np.array_equal(a["X"], b["X"].numpy()) and np.array_equal(a["start_X"], b["start_X"].numpy())


def test_collate_oracle_matches_live_reference():
    ref = _reference.load_reference()
    if ref is None:
        pytest.skip("reference tree not present (GPU box)")
    paths, obs, nb, dt = _batch(seed=4)
    items = [{"idx": i, "stock_path": paths[i:i + 1], "observed_dates": obs[i:i + 1], "nb_obs": nb[i:i + 1], "dt": dt}
             for i in range(len(paths))]
    r = ref.data_utils.custom_collate_fn(items)
    a = so.collate(paths, obs, nb, dt)
    assert np.array_equal(r["times"], a["times"]) and np.array_equal(r["time_ptr"], a["time_ptr"])
    assert np.array_equal(r["obs_idx"].numpy(), a["obs_idx"])
    assert np.array_equal(r["X"].numpy(), a["X"]) and np.array_equal(r["start_X"].numpy(), a["start_X"])
    assert np.array_equal(r["n_obs_ot"].numpy(), a["n_obs_ot"])


def test_shared_library_exports_every_declared_symbol():
    """no compute call: only dlopen + symbol lookup (works without a GPU)"""
    hdr = open(os.path.join(ROOT, "include", "njode_b200.h")).read()
    names = set(re.findall(r"\b(njode_[a-z0-9_]+)\s*\(", hdr))
    assert {"njode_forward", "njode_backward", "njode_plan", "njode_sde_generate", "njode_collate"} <= names
    lib = ctypes.CDLL(os.path.join(ROOT, "njode_b200", "libnjode_b200.so"))
    for n in sorted(names):
        assert hasattr(lib, n), "libnjode_b200.so does not export " + n
    lib.njode_abi_version.restype = ctypes.c_int
    assert lib.njode_abi_version() == 6
