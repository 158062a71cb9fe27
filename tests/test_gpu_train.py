"""GPU: end-to-end training on the B200 path (scripts/train_demo.py restates NJODE/train.py:488-579): the eval loss of the
demo model falls towards the analytic optimum like the reference's published curve, the dropout keep-rate measured on the
device is 1 - p, and the bf16 tensor-core path trains like the fp32 path."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
pytestmark = pytest.mark.gpu


def test_dropout_keep_rate_on_device():
    import train_demo
    kr = train_demo.keep_rate_on_device(0.1)          # 2048 x 100 x 50 = 10.2 M draws: sigma = 1e-4
    assert abs(kr - 0.9) < 1e-3, kr


def test_demo_model_trains_towards_the_optimal_loss():
    """a short run (6 epochs of 40 batches of 200 paths): the reference goes from 2.23x the optimum after epoch 1 to 1.5x
    after epoch 6 with 80 batches per epoch; half the data here, so only the direction and a loose level are asserted --
    the full 20 000-path curve is profiles/r2_train_bs_demo.json"""
    import train_demo
    out = train_demo.run(epochs=6, paths=10000, log=lambda s: None)
    r = [c["ratio"] for c in out["curve"]]
    assert r[-1] < r[0] and r[-1] < 2.2, r
    assert all(c["eval_loss"] > 0.9 * out["optimal_eval_loss"] for c in out["curve"])


def test_tensor_core_path_trains_like_fp32():
    import train_demo
    a = train_demo.run(epochs=2, paths=2560, steps=50, batch=512, d=16, H=256, width=256, layers=4, tensor_cores="on", log=lambda s: None)
    b = train_demo.run(epochs=2, paths=2560, steps=50, batch=512, d=16, H=256, width=256, layers=4, tensor_cores="off", log=lambda s: None)
    assert a["curve"][0]["path"] == "tcgen05" and b["curve"][0]["path"] == "fp32"
    for x, y in zip(a["curve"], b["curve"]):
        assert abs(x["eval_loss"] - y["eval_loss"]) <= 0.05 * abs(y["eval_loss"]), (x, y)
