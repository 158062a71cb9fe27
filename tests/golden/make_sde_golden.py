"""Generates tests/golden/sde_ref.npz by running the REAL reference generators (imported from
/root/reference, build container only): for every model the standard normals the reference drew
(replayed from the same np.random seed) and the paths it produced, plus sample moments of the
reference generator at 20 000 paths for the distributional checks of the CUDA generator.
Run:  python tests/golden/make_sde_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _reference  # noqa: E402

ref = _reference.load_reference()
assert ref is not None, "needs /root/reference"
HP = dict(drift=2., volatility=0.3, mean=4, speed=2., correlation=0.5, nb_paths=7, nb_steps=20, S0=[1., 1.5],
          maturity=1., dimension=2, sine_coeff=3.0, v0=0.5, return_vol=False)
out = {"hp": np.array(json.dumps(HP))}
for name in ("BlackScholes", "OrnsteinUhlenbeck", "Heston", "HestonWOFeller", "HestonWOFeller_vol"):
    model = name.split("_")[0]
    h = dict(HP, return_vol=name.endswith("_vol"))
    np.random.seed(3)
    paths, dt = ref.stock_model.STOCK_MODELS[model](**h).generate_paths()
    np.random.seed(3)
    two = model.startswith("Heston")
    n1, n2 = np.empty((7, 2, 20)), np.zeros((7, 2, 20))
    for i in range(7):
        for k in range(20):
            n1[i, :, k] = np.random.normal(0, 1, 2)
            if two:
                n2[i, :, k] = np.random.normal(0, 1, 2)
    out[name + "/n1"], out[name + "/n2"], out[name + "/paths"] = n1, n2, paths
# moments of the reference generators with the demo hyper-parameters (NJODE/data_utils.py:25-31)
DEMO = dict(drift=2., volatility=0.3, mean=4, speed=2., correlation=0.5, nb_paths=20000, nb_steps=100, S0=1,
            maturity=1., dimension=1)
for model in ("BlackScholes", "OrnsteinUhlenbeck", "Heston"):
    np.random.seed(0)
    paths, dt = ref.stock_model.STOCK_MODELS[model](**DEMO).generate_paths()
    out["moments/%s/mean" % model] = paths[:, 0, :].mean(axis=0)
    out["moments/%s/var" % model] = paths[:, 0, :].var(axis=0)
    out["moments/%s/quantiles" % model] = np.quantile(paths[:, 0, :], [0.25, 0.5, 0.75], axis=0)
out["demo_hp"] = np.array(json.dumps(DEMO))
np.savez_compressed(os.path.join(HERE, "sde_ref.npz"), **out)
print("wrote sde_ref.npz")
