"""debug driver: one physionet-shaped masked batch through the path kernels (train or eval mode), compared with the oracle"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import cases, parity_util
import oracle.njode_oracle as orc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 12
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
batch = cases.irregular_batch(B, 41, steps, seed=7, masked=True, times_f32=True, obs_at_zero=True, row_prob=0.2, feat_prob=0.12)
cfg = dict(cases.CONFIGS["masked_physio"], dropout_rate=0.0)
dt, T = (0.01 if steps == 40 else 1.0 / steps), 1 + 1e-12
ocfg = orc.Config(**cfg)
sd = orc.init_state_dict(ocfg, seed=3)
m = parity_util.build_model(cfg, sd, "cuda:0"); m.eval()
hT, loss = parity_util.call(m, batch, {"delta_t": dt, "T": T}, "cuda:0")
loss.backward()
o_hT, o_loss, o_g = orc.loss_and_grads(ocfg, sd, batch, dt, T)
print("R=%s loss %.3e hT %.3e" % (os.environ.get("NJODE_PATH_R"), parity_util.rel_err(loss.detach().numpy(), o_loss.numpy()), parity_util.rel_err(hT.detach().cpu().numpy(), o_hT.numpy())),
      " ".join("%s=%.1e" % (n.replace("encoder_map.ffnn", "enc").replace("readout_map.ffnn", "ro").replace("ode_f.f", "ode").replace("weight", "w").replace("bias", "b"),
                            parity_util.rel_err(p.grad.cpu().numpy(), o_g[n].numpy())) for n, p in m.named_parameters()))
