"""where the end-to-end step of bench.py spends its time (host staging vs kernels), heston_demo_20k"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from njode_b200 import models, schedule
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "heston_demo_20k"]
dev = torch.device("cuda:0")
batch, dt = bench.synth_batch(wl, 4321, 0, wl["paths"])
T = bench.horizon(wl)
torch.manual_seed(0)
model = models.NJODE(**bench.model_cfg(wl)).to(dev).train()
args = (batch["times"], batch["time_ptr"], batch["X"], batch["obs_idx"], dt, T, batch["start_X"], batch["n_obs_ot"])
kw = {"M": batch["M"]} if "M" in batch else {}
def sync(): torch.cuda.synchronize()
for _ in range(3):
    pb = model.prepare_batch(*args, **kw); hT, loss = model.forward_prepared(pb); loss.backward(); float(loss)
res = {}
def tm(name, f, n=20):
    sync(); t0 = time.perf_counter()
    for _ in range(n): f()
    sync(); res[name] = (time.perf_counter() - t0) / n * 1e3
tm("build_schedule", lambda: schedule.build_schedule(batch["times"], dt, T, False, False))
tm("prepare_batch (host + H2D + index build, synced)", lambda: (model.prepare_batch(*args, **kw), sync()))
tm("prepare_batch (host side only, no sync)", lambda: model.prepare_batch(*args, **kw))
pb = model.prepare_batch(*args, **kw)
def fb():
    for p in model.parameters(): p.grad = None
    hT, loss = model.forward_prepared(pb); loss.backward(); return float(loss.detach())
tm("forward_prepared + backward + loss D2H", fb)
def full():
    for p in model.parameters(): p.grad = None
    hT, loss = model(*args, **kw); loss.backward(); return float(loss.detach())
tm("full public-API step", full)
for k, v in res.items(): print("%-55s %8.3f ms" % (k, v))
