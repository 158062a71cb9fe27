"""Generates the golden fixtures in this directory by running the REAL reference
(/root/reference, imported with the shims in tests/_reference.py; only available in the build
container).  Run:  python tests/golden/make_golden.py

For every case the reference NJODE (eval mode, fp32, CPU) produces
  loss, hT, d(loss + <hT,G>)/d(params)           (training-shaped call, until_T=False)
  path_t, path_h, path_y, loss_T, hT_T           (return_path=True, until_T=True, get_loss=True)
on seeded inputs; inputs, weights and outputs are stored as npz.  The script finally checks the
oracle restatement (oracle/njode_oracle.py) against what it just stored and prints the deviations.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from _reference import load_reference, REFERENCE_ROOT   # noqa: E402
import cases                                            # noqa: E402


def ref_generated_batch(ref, model_name, B, seed, **over):
    """paths from the reference generator + the reference collate function."""
    hp = dict(ref.data_utils.hyperparam_default)
    hp.update(nb_paths=B, **over)
    np.random.seed(seed)
    sm = ref.stock_model.STOCK_MODELS[model_name](**hp)
    paths, dt = sm.generate_paths()
    observed = (np.random.random((B, paths.shape[2])) < hp["obs_perc"]) * 1
    nb_obs = observed[:, 1:].sum(axis=1)
    items = [{"idx": [i], "stock_path": paths[[i]], "observed_dates": observed[[i]],
              "nb_obs": nb_obs[[i]], "dt": dt} for i in range(B)]
    b = ref.data_utils.custom_collate_fn(items)
    return b, dt, hp["maturity"]


def run_reference(ref, cfg, sd, batch, delta_t, T, seed):
    model = ref.models.NJODE(**cfg)
    if sd is None:
        torch.manual_seed(seed)
        model = ref.models.NJODE(**cfg)
        g = torch.Generator().manual_seed(seed + 1)
        for n, p in model.named_parameters():
            if n.endswith("bias"):
                p.data = (torch.rand(p.shape, generator=g) * 2 - 1) * 0.1
    else:
        model.load_state_dict(sd)
    model.eval()
    M = batch.get("M")
    g = torch.Generator().manual_seed(seed + 2)
    G = torch.randn(batch["start_X"].shape[0], cfg["hidden_size"], generator=g) * 0.05
    hT, loss = model(batch["times"], batch["time_ptr"], batch["X"], batch["obs_idx"], delta_t, T,
                     batch["start_X"], batch["n_obs_ot"], return_path=False, get_loss=True, M=M)
    names = [n for n, _ in model.named_parameters()]
    grads = torch.autograd.grad(loss, list(model.parameters()), retain_graph=True)
    grads_G = torch.autograd.grad(loss + (hT * G).sum(), list(model.parameters()))
    with torch.no_grad():
        hT_T, loss_T, path_t, path_h, path_y = model(
            batch["times"], batch["time_ptr"], batch["X"], batch["obs_idx"], delta_t, T,
            batch["start_X"], batch["n_obs_ot"], return_path=True, get_loss=True, until_T=True, M=M)
    outs = {"loss": loss.detach(), "hT": hT.detach(), "G": G, "loss_T": loss_T, "hT_T": hT_T,
            "path_t": np.asarray(path_t, dtype=np.float64), "path_h": path_h, "path_y": path_y}
    for n, a, b in zip(names, grads, grads_G):
        outs["grad/" + n] = a
        outs["gradG/" + n] = b
    return {k: v.detach().clone() for k, v in model.state_dict().items()}, outs


def main():
    ref = load_reference()
    assert ref is not None, "reference tree not found at " + REFERENCE_ROOT
    todo = []
    # (name, cfg, batch, delta_t, T, state_dict or None)
    for name, mid, sm, B in (("bs_ckpt1", 1, "BlackScholes", 24), ("heston_ckpt2", 2, "Heston", 16),
                             ("ou_ckpt3", 3, "OrnsteinUhlenbeck", 16)):
        ck = torch.load(os.path.join(REFERENCE_ROOT, "data/saved_models/id-%d/last_checkpoint/checkpt.tar" % mid),
                        weights_only=False)
        b, dt, T = ref_generated_batch(ref, sm, B, seed=mid)
        todo.append((name, cases.CONFIGS["demo"], b, dt, T, ck["model_state_dict"]))
    b = cases.grid_batch(12, 2, 20, 0.25, seed=11)
    todo.append(("easy_w07_nores", cases.CONFIGS["easy_w07_nores"], b, 0.05, 1.0, None))
    b = cases.irregular_batch(9, 3, 14, seed=12)
    todo.append(("curt_nobias_relu", cases.CONFIGS["curt_nobias_relu"], b, 0.04, 1.0, None))
    b = cases.grid_batch(10, 4, 16, 0.3, seed=13)
    todo.append(("res_case2", cases.CONFIGS["res_case2"], b, 1.0 / 16, 1.0, None))
    b = cases.irregular_batch(7, 5, 18, seed=14, masked=True, times_f32=True, obs_at_zero=True)
    todo.append(("masked_small", cases.CONFIGS["masked_small"], b, 0.03, 1 + 1e-12, None))
    b = cases.irregular_batch(8, 1, 12, seed=15)
    todo.append(("irregular_demo", cases.CONFIGS["demo"], b, 0.05, 1.0, None))

    b = cases.grid_batch(14, 1, 20, 0.25, seed=16)
    todo.append(("gru_demo", cases.CONFIGS["gru_demo"], b, 0.05, 1.0, None))
    b = cases.irregular_batch(9, 3, 14, seed=17)
    todo.append(("gru_d3_nores", cases.CONFIGS["gru_d3_nores"], b, 0.04, 1.0, None))
    b = cases.irregular_batch(7, 4, 15, seed=18, masked=True, times_f32=True, obs_at_zero=True)
    todo.append(("gru_masked", cases.CONFIGS["gru_masked"], b, 0.03, 1 + 1e-12, None))

    import oracle.njode_oracle as orc
    only = set(sys.argv[1:])            # optional: regenerate only the named cases
    for i, (name, cfg, batch, dt, T, sd) in enumerate(todo):
        if only and name not in only:
            continue
        sd, outs = run_reference(ref, cfg, sd, batch, dt, T, seed=100 + i)
        meta = {"delta_t": dt, "T": T}
        cases.save_case(os.path.join(HERE, name + ".npz"), cfg, meta, sd, batch, outs)
        # immediate cross-check of the restatement
        ocfg = orc.Config(**cfg)
        hT, loss, g = orc.loss_and_grads(ocfg, sd, batch, dt, T)
        dl = abs(float(loss) - float(outs["loss"])) / max(abs(float(outs["loss"])), 1e-30)
        dg = max(float((g[n] - outs["grad/" + n]).abs().max() / (outs["grad/" + n].abs().max() + 1e-30))
                 for n in sd)
        print("%-18s B=%-3d K=%-3d N=%-4d loss=%.6g  oracle rel.dev loss %.2e grads %.2e" % (
            name, batch["start_X"].shape[0], len(batch["times"]), len(batch["obs_idx"]),
            float(outs["loss"]), dl, dg))


if __name__ == "__main__":
    main()
