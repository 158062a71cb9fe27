import os, sys, time, cProfile, pstats
ROOT = os.getcwd(); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, bench
from njode_b200 import models
wl = bench.WORKLOADS["heston_demo_20k"]
dev = torch.device("cuda:0")
batch, dt = bench.synth_batch(wl, 4321, 0, wl["paths"])
m = models.NJODE(**bench.model_cfg(wl)).to(dev).train()
args = (batch["times"], batch["time_ptr"], batch["X"], batch["obs_idx"], dt, 1.0, batch["start_X"], batch["n_obs_ot"])
for _ in range(5): m.prepare_batch(*args); torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(50): m.prepare_batch(*args)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
