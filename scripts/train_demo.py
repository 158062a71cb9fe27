#!/usr/bin/env python
"""End-to-end training of the demo model on the B200 path, next to the reference's published curve.

Restates the loop of NJODE/train.py:488-579 (Adam lr 1e-3, weight_decay 5e-4, batch 200, dropout 0.1, one validation
batch = the whole validation set, `eval_loss` and `optimal_eval_loss` per epoch) on a device-generated dataset
(njode_b200.stock_model.DeviceDataset: Philox Euler-Maruyama paths + device collate), 80/20 split.  Reports per epoch
train loss, eval loss, eval_loss / optimal_eval_loss, wall time of the epoch including collate and optimizer, and the
reference's own numbers at the same epoch (data/saved_models/id-1/metric_id-1.csv of the reference repository: BlackScholes,
batch 200; copied into REF_CURVE below as data, not code).  Also measures the dropout keep-rate of the kernels on the
device (see keep_rate_on_device).

    python scripts/train_demo.py [--epochs 20] [--paths 20000] [--batch 200] [--sde BlackScholes] [--out file.json]
    python scripts/train_demo.py --compare-tensor-cores      # short run of the d=16 / H=256 model, tcgen05 (bf16) vs fp32
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from njode_b200 import models, stock_model  # noqa: E402

# epoch -> (eval_loss, optimal_eval_loss) of the reference's shipped BlackScholes run (metric_id-1.csv, 200 epochs on CPU,
# 72-84 s per epoch)
REF_CURVE = {1: (0.31442, 0.1408), 2: (0.29342, 0.1408), 3: (0.26811, 0.1408), 4: (0.21733, 0.1408), 10: (0.19563, 0.1408),
             20: (0.18179, 0.1408), 26: (0.15392, 0.1408), 30: (0.15323, 0.1408), 40: (0.15017, 0.1408), 100: (0.14813, 0.1408),
             200: (0.14757, 0.1408)}
HP = dict(drift=2., volatility=0.3, mean=4, speed=2., correlation=0.5, S0=1, maturity=1., dimension=1, sine_coeff=None,
          scheme='euler', return_vol=False, v0=1)             # NJODE/data_utils.py:25-31


def demo_cfg(d=1, H=10, width=50, layers=2, dropout=0.1):
    nn_desc = [[width, "tanh"]] * layers
    return dict(input_size=d, hidden_size=H, output_size=d, ode_nn=nn_desc, readout_nn=nn_desc, enc_nn=nn_desc, use_rnn=False,
                bias=True, dropout_rate=dropout, solver="euler", weight=0.5, weight_decay=1.0,
                options={"which_loss": "standard", "residual_enc_dec": True})


def keep_rate_on_device(p=0.1, paths=2048, steps=100, dev="cuda:0"):
    """fraction of hidden units the kernels keep under dropout rate p, measured on the device: an ODE network whose
    second hidden layer outputs tanh(20) = 1 whatever its input and whose last layer averages it makes
    f = mean_j keep_j / (1 - p), so hT of a batch without observations integrates to T * keep_rate / (1 - p)
    (encoder zeroed, no residual).  paths x steps x 50 Bernoulli draws."""
    cfg = demo_cfg(dropout=p)
    cfg["options"] = {"residual_enc_dec": False}
    m = models.NJODE(**cfg).to(dev)
    with torch.no_grad():
        for prm in m.parameters():
            prm.zero_()
        lin = [x for x in m.ode_f.f if isinstance(x, torch.nn.Linear)]
        lin[1].bias.fill_(20.0)
        lin[2].weight.fill_(1.0 / 50)
    m.train()
    B = paths
    with torch.no_grad():
        hT, _ = m(np.zeros(0), np.zeros(1, dtype=np.int64), torch.zeros(0, 1), torch.zeros(0, dtype=torch.int64), 1.0 / steps, 1.0,
                  torch.ones(B, 1), None, get_loss=False, until_T=True)
    return float(hT.double().mean()) * (1.0 - p)


def run(sde="BlackScholes", epochs=20, paths=20000, steps=100, batch=200, seed=0, d=1, H=10, width=50, layers=2,
        tensor_cores="auto", dev="cuda:0", log=print, lr=1e-3):
    hp = dict(HP, nb_paths=paths, nb_steps=steps, obs_perc=0.1, dimension=d, S0=[1.0] * d if d > 1 else 1)
    t0 = time.perf_counter()
    ds = stock_model.DeviceDataset(sde, hp, seed=seed, device=dev)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    rng = np.random.default_rng(398)                              # the reference splits with random_state=398
    perm = rng.permutation(paths)
    n_val = paths // 5
    val_idx, train_idx = np.sort(perm[:n_val]), perm[n_val:]
    torch.manual_seed(seed)
    model = models.NJODE(**demo_cfg(d, H, width, layers)).to(dev)
    model.tensor_cores = tensor_cores
    opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=0.0005)        # NJODE/train.py:397-398
    T, dt = hp["maturity"], ds.dt
    vb = ds.collate(val_idx)
    opt_loss = float(ds.model.get_optimal_loss(vb["times"], vb["time_ptr"], vb["X"].cpu().numpy(), vb["obs_idx"].numpy(), dt, T,
                                               vb["start_X"].cpu().numpy(), vb["n_obs_ot"].numpy()))
    curve = []
    for epoch in range(1, epochs + 1):
        model.train()
        order = train_idx[np.random.default_rng(1000 + epoch).permutation(len(train_idx))]
        t0 = time.perf_counter()
        for i in range(0, len(order), batch):
            b = ds.collate(np.sort(order[i:i + batch]), on_device=True)
            opt.zero_grad()
            hT, loss = model(b["times"], b["time_ptr"], b["X"], b["obs_idx"], dt, T, b["start_X"], b["n_obs_ot"])
            loss.backward()
            opt.step()
        torch.cuda.synchronize()
        train_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        model.eval()
        with torch.no_grad():
            _, vloss = model(vb["times"], vb["time_ptr"], vb["X"], vb["obs_idx"], dt, T, vb["start_X"], vb["n_obs_ot"])
        eval_s = time.perf_counter() - t0
        model.weight_decay_step()
        rec = {"epoch": epoch, "train_time_s": train_s, "eval_time_s": eval_s, "train_loss": float(loss.detach()), "eval_loss": float(vloss),
               "optimal_eval_loss": opt_loss, "ratio": float(vloss) / opt_loss, "path": model.last_forward_path}
        if sde == "BlackScholes" and d == 1 and epoch in REF_CURVE:
            rec["reference_eval_loss"], rec["reference_optimal"] = REF_CURVE[epoch]
            rec["reference_ratio"] = REF_CURVE[epoch][0] / REF_CURVE[epoch][1]
        curve.append(rec)
        log("epoch %3d  train %.3fs  eval %.3fs  train-loss %.5f  eval-loss %.5f  optimal %.5f  ratio %.4f%s" % (
            epoch, train_s, eval_s, rec["train_loss"], rec["eval_loss"], opt_loss, rec["ratio"],
            ("   [reference: %.5f / %.4f = %.4f]" % (rec["reference_eval_loss"], rec["reference_optimal"], rec["reference_ratio"]))
            if "reference_ratio" in rec else ""))
    return {"sde": sde, "paths": paths, "steps": steps, "batch": batch, "epochs": epochs, "d": d, "H": H,
            "mlp": "%dx%d tanh" % (layers, width), "tensor_cores": tensor_cores, "dataset_generation_s": gen_s,
            "batches_per_epoch": (len(train_idx) + batch - 1) // batch, "optimal_eval_loss": opt_loss, "curve": curve}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sde", default="BlackScholes")
    ap.add_argument("--epochs", type=int, default=20)
    ap.add_argument("--paths", type=int, default=20000)
    ap.add_argument("--batch", type=int, default=200)
    ap.add_argument("--out", default=None)
    ap.add_argument("--compare-tensor-cores", action="store_true")
    args = ap.parse_args()
    out = {}
    if args.compare_tensor_cores:
        # the scaled architecture (BASELINE config 5: d = 16, H = 256, 4x256) at a size that trains in seconds:
        # same data, same seeds, tcgen05 (bf16 operands) vs the fp32 FMA kernels
        for tc in ("on", "off"):
            print("--- tensor_cores = %s" % tc)
            out["tensor_cores_" + tc] = run(epochs=args.epochs, paths=5120, steps=100, batch=512, d=16, H=256, width=256, layers=4,
                                            tensor_cores=tc)
        a, b = out["tensor_cores_on"]["curve"], out["tensor_cores_off"]["curve"]
        out["max_rel_diff_eval_loss_bf16_vs_fp32"] = max(abs(x["eval_loss"] - y["eval_loss"]) / y["eval_loss"] for x, y in zip(a, b))
        print("max relative difference of the eval loss, bf16 tcgen05 vs fp32: %.3e" % out["max_rel_diff_eval_loss_bf16_vs_fp32"])
    else:
        kr = keep_rate_on_device(0.1)
        print("dropout keep-rate measured on the device (p = 0.1, 10.2 M draws): %.5f" % kr)
        out = run(sde=args.sde, epochs=args.epochs, paths=args.paths, batch=args.batch)
        out["dropout_keep_rate_measured"] = kr
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
