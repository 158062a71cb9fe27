"""the evaluation path (NJODE.evaluate, NJODE/models.py:521-562 = forward(return_path=True, until_T=True) + the analytic
conditional expectation of the stock model + MSE) on a validation-set sized batch: wall time per call, kernels vs host"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from njode_b200 import models, stock_model
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
wl = dict(bench.WORKLOADS["heston_demo_20k"], sde="BlackScholes", paths=B)
dev = torch.device("cuda:0")
batch, dt = bench.synth_batch(wl, 77, 0, B)
torch.manual_seed(0)
model = models.NJODE(**bench.model_cfg(wl)).to(dev)
sm = stock_model.BlackScholes(drift=2.0, volatility=0.3, nb_paths=B, nb_steps=100, S0=1.0, maturity=1.0, dimension=1)
args = (batch["times"], batch["time_ptr"], batch["X"], batch["obs_idx"], dt, 1.0, batch["start_X"], batch["n_obs_ot"])
def sync(): torch.cuda.synchronize()
res = {}
def tm(name, f, n=5):
    f(); sync(); t0 = time.perf_counter()
    for _ in range(n): f()
    sync(); res[name] = (time.perf_counter() - t0) / n * 1e3
model.eval()
def fwd_path():
    with torch.no_grad():
        return model(*args[:7], None, return_path=True, get_loss=False, until_T=True)
model.output_device = "cuda"
tm("forward(return_path=True), outputs stay on the device", fwd_path)
model.output_device = "cpu"
tm("forward(return_path=True), path_h / path_y copied to the host", fwd_path)
tm("stockmodel.compute_cond_exp (host NumPy)", lambda: sm.compute_cond_exp(
    args[0], args[1], batch["X"].numpy(), batch["obs_idx"].numpy(), dt, 1.0, batch["start_X"].numpy(), batch["n_obs_ot"].numpy(),
    return_path=True, get_loss=False))
tm("model.evaluate (all of it)", lambda: model.evaluate(*args, sm))
print("eval batch: %d paths x 100 steps, demo nets" % B)
for k, v in res.items(): print("%-62s %9.3f ms" % (k, v))
