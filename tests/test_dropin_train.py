"""Drop-in check of the module boundary: the reference's UNMODIFIED training driver
(NJODE/train.py::train -- DataLoader + custom_collate_fn, NJODE(**params_dict), Adam, loss.backward(), the eval
loop, model.evaluate, save_checkpoint / get_ckpt_model, the metric CSV) is run twice on the same seeded dataset:
once with the reference's own `models` module and once with `njode_b200.models` swapped in.  Training loss, eval
loss and the evaluation mean-square difference of every epoch must agree.

Only possible where the reference tree exists (the build container); skipped elsewhere.  No GPU here, so the
kernels run as their host simulation (tests/hostsim) -- the swap itself (names, signatures, state, return types,
checkpoint format) is what this test pins; numerical parity of the CUDA path is tests/test_gpu_parity.py."""
import os
import sys

import numpy as np
import pytest
import torch

import hostsim_util
from _reference import load_reference, REFERENCE_ROOT

ref = load_reference()
pytestmark = pytest.mark.skipif(ref is None, reason="reference tree not available")

NN = [[16, "tanh"], [16, "tanh"]]


def _run(ref_train, tmp, model_id, use_rnn=False, resume_epochs=None):
    torch.manual_seed(7)
    np.random.seed(7)
    ref_train.train(model_id=model_id, epochs=resume_epochs or 3, batch_size=10, save_every=1, learning_rate=0.01,
                    hidden_size=6, bias=True, dropout_rate=0.0, ode_nn=NN, readout_nn=NN, enc_nn=NN, use_rnn=use_rnn,
                    dataset="BlackScholes", dataset_id=None, plot=False,
                    saved_models_path=os.path.join(tmp, "saved_models") + "/", evaluate=True)
    import pandas as pd
    return pd.read_csv(os.path.join(tmp, "saved_models", "id-%d" % model_id, "metric_id-%d.csv" % model_id), index_col=0)


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    """scratch tree shaped like the reference checkout: <tmp>/NJODE is the CWD, datasets land in <tmp>/data"""
    root = tmp_path_factory.mktemp("dropin")
    (root / "NJODE").mkdir()
    cwd = os.getcwd()
    os.chdir(root / "NJODE")
    hp = dict(ref.data_utils.hyperparam_default)
    hp.update(nb_paths=40, nb_steps=20, obs_perc=0.25)
    ref.data_utils.create_dataset("BlackScholes", hp, seed=3)
    yield str(root / "data")
    os.chdir(cwd)


@pytest.mark.parametrize("use_rnn", [False, True])
def test_reference_train_loop_drives_the_b200_models_module(workdir, use_rnn):
    if os.path.join(REFERENCE_ROOT) not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import NJODE.train as ref_train
    from njode_b200 import models as b200_models
    base = 10 if use_rnn else 0
    ref_models = ref_train.models
    try:
        want = _run(ref_train, workdir, base + 1, use_rnn)
        ref_train.models = b200_models                          # the swap a user makes
        hostsim_util.install()        # no GPU here: host simulation of the kernels
        got = _run(ref_train, workdir, base + 2, use_rnn)
        # resume from the checkpoint our save_checkpoint wrote, with the reference's loader logic (get_ckpt_model)
        more = _run(ref_train, workdir, base + 2, use_rnn, resume_epochs=4)
    finally:
        ref_train.models = ref_models
        hostsim_util.uninstall()
    assert list(got.columns) == list(want.columns) and len(got) == len(want) == 3
    for col in ("train_loss", "eval_loss", "optimal_eval_loss", "evaluation_mean_diff"):
        np.testing.assert_allclose(got[col].values.astype(float), want[col].values.astype(float), rtol=2e-3, err_msg=col)
    assert len(more) == 4 and float(more["eval_loss"].values[-1]) > 0
    # the checkpoint written through our module loads into the REFERENCE model class (same state_dict keys)
    ck = torch.load(os.path.join(workdir, "saved_models", "id-%d" % (base + 2), "last_checkpoint", "checkpt.tar"), weights_only=False)
    m = ref_models.NJODE(input_size=1, hidden_size=6, output_size=1, ode_nn=NN, readout_nn=NN, enc_nn=NN, use_rnn=use_rnn,
                         bias=True, dropout_rate=0.0, options={})
    m.load_state_dict(ck["model_state_dict"])
