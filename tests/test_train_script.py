"""scripts/train_demo.py: the on-device dropout keep-rate measurement, on the host simulation (CPU) -- the training run
itself is a GPU test (tests/test_gpu_train.py)."""
import os
import sys

import hostsim_util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def test_keep_rate_measurement_on_the_host_simulation():
    import train_demo
    hostsim_util.install()
    try:
        kr = train_demo.keep_rate_on_device(0.25, paths=96, steps=40, dev="cpu")
    finally:
        hostsim_util.uninstall()
    # 96 x 40 x 50 = 192 000 Bernoulli(0.75) draws: sigma = 0.001
    assert abs(kr - 0.75) < 0.005, kr
