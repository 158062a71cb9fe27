#!/usr/bin/env python
"""bench.py -- NJ-ODE training hot path throughput (train paths x Euler-steps / s, fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" = one forward + backward of the NJ-ODE over one batch (+ the gradient all-reduce when
N > 1).  Headline workload at every N (weak scaling, per-GPU work fixed): BASELINE.json configs[1] -- the
demo architecture (d=1, H=10, 2x50 tanh ode/enc/readout, dropout 0.1, train mode) on Heston paths,
100 Euler steps, obs_perc 0.1, 20 000 paths per GPU in one batch.  Data are synthetic (seeded
Euler-Maruyama Heston, NJODE/stock_model.py:181-221 restated), weights random-init (Xavier, seed 0).

The same command also measures the two configurations BASELINE.json's targets are quoted on and reports them in
``target_configs`` of the same JSON line ("next" #1):
  bs_demo_200          the reference's own batch (model_overview.csv:2: BlackScholes, batch 200): value / e2e on the B200
                       next to the CPU arm on the SAME 200 paths, one thread (the reference's server setting) and all host
                       threads, and the resulting ratios (target: >= 50x); N = 1 only
  bs_scaled_d16_h256   d = 16, H = 256, 4x256 nets, 1000 steps, device-generated paths, tcgen05 kernels: value / e2e /
                       roofline at every N, so the 1 -> 8 curve of the 3.5 MB all-reduce is computable from the scaling run
and, when N > 1, ``dp_check``: the all-reduced flat gradient is bit-identical on every rank and equals the single-GPU
gradient of the concatenated batch (rtol 1e-5), checked outside the timed regions.

Printed JSON (rank 0, one line):
  value        whole-job paths*steps/s with the collated batch already resident in HBM
               (CUDA events around every step, max over ranks, L2 flushed between steps)
  e2e          same metric through the public API `model(times, time_ptr, X, obs_idx, ...)` with
               HOST tensors: host schedule/CSR build + one pinned H2D copy + fwd + bwd + D2H of the
               loss inside the timed region
  roofline     dominant kernel (its name comes from the library: njode_last_kernel): algorithmic fp32 flops /
               CUDA-event duration over the measured fp32-FMA peak of this GPU (narrow nets are FMA-pipe bound, not
               HBM / tensor bound: SURVEY.md 8d); the HBM view of the same launch is in roofline.hbm
  cpu_baseline the oracle port of the reference (oracle/njode_oracle.py, same ATen op sequence as
               NJODE/models.py) timed on this box's host cores on a bounded sample (N=1 only)
`--impl reference` times only that CPU port with all host threads (one step = the bounded sample).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "train_paths_x_euler_steps_per_sec_fwd_bwd"
UNIT = "paths*steps/s"

WORKLOADS = {
    # BASELINE.json configs[1]: Heston demo config, same architecture as the BS demo, 20k paths
    "heston_demo_20k": dict(sde="Heston", paths=20000, steps=100, d=1, H=10, width=50, layers=2,
                            obs_perc=0.1, dropout=0.1, cpu_sample_paths=2000),
    "ou_demo_20k": dict(sde="OrnsteinUhlenbeck", paths=20000, steps=100, d=1, H=10, width=50,
                        layers=2, obs_perc=0.1, dropout=0.1, cpu_sample_paths=2000),
    # BASELINE.json configs[0]: the reference's own CPU-runnable demo batch
    "bs_demo_200": dict(sde="BlackScholes", paths=200, steps=100, d=1, H=10, width=50, layers=2,
                        obs_perc=0.1, dropout=0.1, cpu_sample_paths=200),
    # batch sweep around it (BASELINE.json configs[2]: batch in {200 .. 20 000})
    "bs_demo_1k": dict(sde="BlackScholes", paths=1000, steps=100, d=1, H=10, width=50, layers=2,
                       obs_perc=0.1, dropout=0.1, cpu_sample_paths=1000),
    "bs_demo_5k": dict(sde="BlackScholes", paths=5000, steps=100, d=1, H=10, width=50, layers=2,
                       obs_perc=0.1, dropout=0.1, cpu_sample_paths=1000),
    # BASELINE.json configs[4]: scaled BlackScholes, d=16, H=256, 4x256 nets, 1000 Euler steps, paths generated
    # on the device (Philox Euler-Maruyama + device collate), tcgen05 tensor-core kernels (bf16 operands)
    # The dataset is BASELINE's "1M on-device-generated paths" sharded over 8 GPUs: every rank generates ITS 131 072 paths
    # (global path ids rank * 131072 ..., 16.8 GB of fp64 paths per GPU) and steps through batches of 8 192 of them.
    "bs_scaled_d16_h256": dict(sde="BlackScholes", paths=8192, steps=1000, d=16, H=256, width=256,
                               layers=4, obs_perc=0.1, dropout=0.1, cpu_sample_paths=128,
                               cpu_sample_steps=100, device_data=True, dataset_paths=131072),
    "bs_scaled_d16_h256_small": dict(sde="BlackScholes", paths=4096, steps=100, d=16, H=256, width=256,
                                     layers=4, obs_perc=0.1, dropout=0.1, cpu_sample_paths=128, device_data=True),
    # BASELINE.json configs[3]: PhysioNet-shaped synthetic irregular series: 41 masked features,
    # 3000-tick grid (delta_t = 0.016/48, T = 1 + 1e-12), 40..110 stamps per record, Bernoulli(0.12) feature masks,
    # float32 times, start_X = 0; masked model d = H = 41, 2x50 nets.  B = 50 is the reference's batch size
    # (parallel_train.py:656); the 2000-record variant shows the throughput of the generic masked kernels.
    "physionet_synth_b50": dict(sde="physionet_synth", paths=50, steps=3000, d=41, H=41, width=50, layers=2,
                                obs_perc=None, dropout=0.1, masked=True, cpu_sample_paths=50),
    "physionet_synth_b50_2x200": dict(sde="physionet_synth", paths=50, steps=3000, d=41, H=41, width=200, layers=2,
                                      obs_perc=None, dropout=0.1, masked=True, cpu_sample_paths=50),
    "physionet_synth_b300": dict(sde="physionet_synth", paths=300, steps=3000, d=41, H=41, width=50, layers=2,
                                 obs_perc=None, dropout=0.1, masked=True, cpu_sample_paths=50),
    "physionet_synth_b600": dict(sde="physionet_synth", paths=600, steps=3000, d=41, H=41, width=50, layers=2,
                                 obs_perc=None, dropout=0.1, masked=True, cpu_sample_paths=50),
    "bs_demo_gru_500": dict(sde="BlackScholes", paths=500, steps=100, d=1, H=10, width=50, layers=2,
                            obs_perc=0.1, dropout=0.1, use_rnn=True, cpu_sample_paths=500),
    "bs_demo_gru_1k": dict(sde="BlackScholes", paths=1000, steps=100, d=1, H=10, width=50, layers=2,
                           obs_perc=0.1, dropout=0.1, use_rnn=True, cpu_sample_paths=500),
    "bs_demo_600": dict(sde="BlackScholes", paths=600, steps=100, d=1, H=10, width=50, layers=2,
                        obs_perc=0.1, dropout=0.1, cpu_sample_paths=600),
    "bs_demo_400": dict(sde="BlackScholes", paths=400, steps=100, d=1, H=10, width=50, layers=2,
                        obs_perc=0.1, dropout=0.1, cpu_sample_paths=400),
    "physionet_synth_b2000": dict(sde="physionet_synth", paths=2000, steps=3000, d=41, H=41, width=50, layers=2,
                                  obs_perc=None, dropout=0.1, masked=True, cpu_sample_paths=50),
    # BASELINE.json configs[2] (ii): Heston without Feller condition (parallel_train.py:525-537), demo nets, batch sweep
    "hestonwof_demo_1k": dict(sde="HestonWOFeller", paths=1000, steps=100, d=1, H=10, width=50, layers=2,
                              obs_perc=0.1, dropout=0.1, cpu_sample_paths=1000),
    "hestonwof_demo_20k": dict(sde="HestonWOFeller", paths=20000, steps=100, d=1, H=10, width=50, layers=2,
                               obs_perc=0.1, dropout=0.1, cpu_sample_paths=2000),
    # use_rnn=True (GRU jump, NJODE/models.py:202-217): whole-path units on the generic kernels
    "bs_demo_gru_5k": dict(sde="BlackScholes", paths=5000, steps=100, d=1, H=10, width=50, layers=2,
                           obs_perc=0.1, dropout=0.1, use_rnn=True, cpu_sample_paths=1000),
    # BASELINE.json configs[2] (i): combined-dataset nets (2x100 tanh), batch 5000
    "bs_2x100_5k": dict(sde="BlackScholes", paths=5000, steps=100, d=1, H=10, width=100, layers=2,
                        obs_perc=0.1, dropout=0.1, cpu_sample_paths=1000),
    "bs_2x100_200": dict(sde="BlackScholes", paths=200, steps=100, d=1, H=10, width=100, layers=2,
                         obs_perc=0.1, dropout=0.1, cpu_sample_paths=200),
    "bs_2x100_20k": dict(sde="BlackScholes", paths=20000, steps=100, d=1, H=10, width=100, layers=2,
                         obs_perc=0.1, dropout=0.1, cpu_sample_paths=1000),
}

SDE_PARAMS = dict(drift=2.0, volatility=0.3, mean=4.0, speed=2.0, correlation=0.5, S0=1.0,
                  maturity=1.0)          # NJODE/data_utils.py:25-31 (hyperparam_default)


# ------------------------------------------------------------------------------------------------
# synthetic data (host, seeded, vectorised over paths): restates the generators' update rules
# ------------------------------------------------------------------------------------------------
def synth_paths(sde, n_paths, steps, d, seed, first_path=0):
    """f64 [n_paths, d, steps+1]; path i depends only on (seed, first_path + i)."""
    p = SDE_PARAMS
    dt = p["maturity"] / steps
    out = np.empty((n_paths, d, steps + 1))
    out[:, :, 0] = p["S0"]
    # one generator per block of 1024 global path ids: identical data for any sharding
    blk = 1024
    for b0 in range((first_path // blk) * blk, first_path + n_paths, blk):
        rng = np.random.default_rng([seed, b0 // blk])
        z = rng.standard_normal((blk, steps, 2, d))
        lo, hi = max(b0, first_path), min(b0 + blk, first_path + n_paths)
        zz = z[lo - b0:hi - b0]
        S = np.full((hi - lo, d), p["S0"])
        v = np.full((hi - lo, d), p["mean"])
        sl = slice(lo - first_path, hi - first_path)
        for k in range(steps):
            n1, n2 = zz[:, k, 0], zz[:, k, 1]
            if sde == "BlackScholes":            # NJODE/stock_model.py:356-375
                S = S + p["drift"] * S * dt + p["volatility"] * S * n1 * np.sqrt(dt)
            elif sde == "OrnsteinUhlenbeck":     # NJODE/stock_model.py:397-418
                S = S - p["speed"] * (S - p["mean"]) * dt + p["volatility"] * n1 * np.sqrt(dt)
            elif sde == "Heston":                # NJODE/stock_model.py:181-221 (spot uses the NEW variance)
                dW = n1 * np.sqrt(dt)
                dZ = (p["correlation"] * n1 + np.sqrt(1 - p["correlation"] ** 2) * n2) * np.sqrt(dt)
                v = v - p["speed"] * (v - p["mean"]) * dt + p["volatility"] * np.sqrt(np.abs(v)) * dZ
                S = S + p["drift"] * S * dt + np.sqrt(np.abs(v)) * S * dW
            elif sde == "HestonWOFeller":        # NJODE/stock_model.py:288-335 (drift 2, volatility 3, mean 1, speed 2, v0 0.5)
                if k == 0:
                    v = np.full((hi - lo, d), 0.5)
                dW = n1 * np.sqrt(dt)
                dZ = (p["correlation"] * n1 + np.sqrt(1 - p["correlation"] ** 2) * n2) * np.sqrt(dt)
                vp = np.maximum(v, 0.0)
                S = S * np.exp((p["drift"] - 0.5 * vp) * dt + np.sqrt(vp) * dW)
                v = v + 2.0 * (1.0 - v) * dt + 3.0 * np.sqrt(vp) * dZ
            else:
                raise ValueError(sde)
            out[sl, :, k + 1] = S
    return out, dt


def synth_batch(wl, seed, first_path, n_paths):
    from njode_b200 import data_utils
    if wl["sde"] == "physionet_synth":
        return synth_batch_physio(wl, seed, first_path, n_paths)
    paths, dt = synth_paths(wl["sde"], n_paths, wl["steps"], wl["d"], seed, first_path)
    blk = 1024
    obs = np.empty((n_paths, wl["steps"] + 1), dtype=np.int64)
    for b0 in range((first_path // blk) * blk, first_path + n_paths, blk):
        rng = np.random.default_rng([seed + 7919, b0 // blk])
        u = rng.random((blk, wl["steps"] + 1))
        lo, hi = max(b0, first_path), min(b0 + blk, first_path + n_paths)
        obs[lo - first_path:hi - first_path] = (u[lo - b0:hi - b0] < wl["obs_perc"]) * 1
    obs[:, 0] = 1                                     # NJODE/data_utils.py:80
    nb_obs = obs[:, 1:].sum(axis=1)
    b = data_utils.collate_paths(paths, obs, nb_obs, dt)
    return b, dt


def synth_batch_physio(wl, seed, first_path, n_paths):
    """PhysioNet-shaped synthetic records in the masked collate contract (latent_ODE/physionet_LODE.py:428-544 flavour):
    float32 ``times`` = union of the records' ticks, rows = (time, record) pairs ordered time-major, ``M`` = feature masks."""
    import torch
    d, grid = wl["d"], wl["steps"]
    rows = []                                          # (tick, record, values, mask)
    for p_ in range(n_paths):
        rng = np.random.default_rng([seed, first_path + p_])
        n_st = int(rng.integers(40, 111))
        ticks = np.concatenate(([0], 1 + np.sort(rng.choice(grid - 1, size=n_st, replace=False))))
        m = rng.random((len(ticks), d)) < 0.12
        m[0, :5] = True
        empty = ~m.any(axis=1)
        m[empty, rng.integers(d, size=int(empty.sum()))] = True
        x = rng.random((len(ticks), d)) * m
        rows.append((ticks, np.full(len(ticks), p_), x, m))
    ticks = np.concatenate([r[0] for r in rows]); rec = np.concatenate([r[1] for r in rows])
    X = np.concatenate([r[2] for r in rows]).astype(np.float32); M = np.concatenate([r[3] for r in rows]).astype(np.float32)
    order = np.lexsort((rec, ticks))
    ticks, rec, X, M = ticks[order], rec[order], X[order], M[order]
    ut, counts = np.unique(ticks, return_counts=True)
    times = (ut.astype(np.float32) * np.float32(0.016)) / np.float32(48.0)
    batch = {"times": times.astype(np.float32), "time_ptr": np.concatenate(([0], np.cumsum(counts))),
             "obs_idx": torch.from_numpy(rec.astype(np.int64)), "X": torch.from_numpy(X), "M": torch.from_numpy(M),
             "start_X": torch.zeros(n_paths, d), "n_obs_ot": torch.from_numpy(np.bincount(rec, minlength=n_paths).astype(np.int64))}
    return batch, 0.016 / 48


def device_dataset(wl, seed, first_path, n_paths, dev):
    """paths + observation mask generated on the device (njode_sde_generate, one Philox subsequence per global
    path id: identical data for any sharding), kept there; -> (dataset, generation ms by CUDA events, bytes written)"""
    import torch
    from njode_b200 import stock_model
    p = SDE_PARAMS
    hp = dict(drift=p["drift"], volatility=p["volatility"], mean=p["mean"], speed=p["speed"],
              correlation=p["correlation"], S0=[p["S0"]] * wl["d"], nb_paths=n_paths, nb_steps=wl["steps"],
              maturity=p["maturity"], sine_coeff=None, obs_perc=wl["obs_perc"])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    ds = stock_model.DeviceDataset(wl["sde"], hp, seed=seed, first_path=first_path, device=dev)
    e1.record()
    torch.cuda.synchronize()
    nbytes = ds.paths.numel() * 8 + ds.observed.numel() * 4 + ds.nb_obs.numel() * 4
    return ds, e0.elapsed_time(e1), nbytes


def synth_batch_device(wl, seed, first_path, n_paths, dev, on_device=False):
    """one batch of a device-generated dataset, collated on the device (njode_collate).  on_device: the index arrays stay
    device tensors too (the resident arm); otherwise they are host arrays as in the reference's contract."""
    import torch
    ds, _, _ = device_dataset(wl, seed, first_path, n_paths, dev)
    b = ds.collate(torch.arange(n_paths), on_device=on_device)
    return b, ds.dt


def model_cfg(wl):
    nn_desc = [[wl["width"], "tanh"]] * wl["layers"]
    return dict(input_size=wl["d"], hidden_size=wl["H"], output_size=wl["d"], ode_nn=nn_desc,
                readout_nn=nn_desc, enc_nn=nn_desc, use_rnn=bool(wl.get("use_rnn")), bias=True,
                dropout_rate=wl["dropout"], solver="euler", weight=0.5, weight_decay=1.0,
                options={"masked": True} if wl.get("masked") else {})


def horizon(wl):
    """T passed to NJODE.forward (physionet_train.py:192-193 uses 1 + 1e-12)"""
    return 1.0 + 1e-12 if wl.get("masked") else SDE_PARAMS["maturity"]


def flops_per_unit(wl):
    """algorithmic fp32 flops: F_ode per path-step, F_enc / F_ro per row; fwd+bwd
    = 3x fwd (recompute not counted)."""
    d, H, W, L = wl["d"], wl["H"], wl["width"], wl["layers"]
    inf = d + H + 2
    F_ode = 2 * (inf * W + (L - 1) * W * W + W * H)
    F_enc = 2 * ((2 * d if wl.get("masked") else d) * W + (L - 1) * W * W + W * H)
    F_ro = 2 * (H * W + (L - 1) * W * W + W * d)
    return F_ode, F_enc, F_ro


# ------------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# the CPU arm: oracle port of the reference (test infrastructure, used here only as the baseline)
# ------------------------------------------------------------------------------------------------
def cpu_port_step_fn(wl, seed, threads, n=None):
    import torch
    import oracle.njode_oracle as orc
    torch.set_num_threads(threads)
    n = wl["cpu_sample_paths"] if n is None else n
    wl_s = dict(wl)
    if "cpu_sample_steps" in wl:
        wl_s["steps"] = wl["cpu_sample_steps"]
    batch, dt = synth_batch(wl_s, seed, 0, n)
    cfg = model_cfg(wl)
    ocfg = orc.Config(**cfg)
    sd = orc.init_state_dict(ocfg, seed=0)
    S = len(__import__("njode_b200.schedule", fromlist=["x"]).build_schedule(
        batch["times"], dt, horizon(wl), False, False).step_dt)

    def step():
        orc.loss_and_grads(ocfg, sd, batch, dt, horizon(wl), dropout_seed="native")
    whole = n == wl["paths"] and "cpu_sample_steps" not in wl
    sample = "%s%d paths x %d Euler steps, fwd+bwd, train mode (aten dropout), %d thread%s" % (
        "the workload's batch: " if whole else "a sample of the workload: ", n, S, threads, "" if threads == 1 else "s")
    return step, n * S, sample


def time_cpu(step, budget_s, min_reps, max_reps):
    step()
    ts = []
    t_end = time.perf_counter() + budget_s
    while len(ts) < min_reps or (time.perf_counter() < t_end and len(ts) < max_reps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


def cpu_arms(wl, budget_all=20.0, budget_one=8.0):
    """(all host threads, one thread) baselines of one workload; the one-thread arm (the reference's own server setting,
    NJODE/train.py:44,209: N_CPUS = 1) runs a quarter of the sample unless the sample is the whole batch"""
    import torch
    threads = os.cpu_count() or 1
    step, units, sample = cpu_port_step_fn(wl, 1234, threads)
    all_ = {"value": units / time_cpu(step, budget_all, 3, 8), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
    whole = wl["cpu_sample_paths"] == wl["paths"]
    n1 = wl["cpu_sample_paths"] if whole else max(1, wl["cpu_sample_paths"] // 4)
    step1, units1, sample1 = cpu_port_step_fn(wl, 1234, 1, n=n1)
    one = {"value": units1 / time_cpu(step1, budget_one, 2, 5), "unit": UNIT, "cores": 1, "kind": "port", "sample": sample1}
    torch.set_num_threads(threads)
    return all_, one


def config_of(wl_name, wl, **extra):
    c = {"workload": wl_name, "sde": wl["sde"], "paths_per_gpu": wl["paths"], "euler_steps": wl["steps"],
         "input_size": wl["d"], "hidden_size": wl["H"], "mlp": "%dx%d tanh" % (wl["layers"], wl["width"]),
         "dropout": wl["dropout"], "mode": "train"}
    c.update(extra)
    return c


def run_reference(args, wl_name, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    step, units, sample = cpu_port_step_fn(wl, 1234, threads)
    for _ in range(args.warmup):
        step()
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(ts))
    val = units / (ms * 1e-3)
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": config_of(wl_name, wl, parallelism="cpu"),
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "torch_threads": torch.get_num_threads(),
           "note": "ONE CPU process on rank 0 whatever --gpus says (n_gpus only echoes the launch)"}
    if args.targets:
        # the reference's own batch (model_overview.csv:2): the full 200 paths, all threads and one thread
        tw = WORKLOADS["bs_demo_200"]
        all_, one = cpu_arms(tw, budget_all=6.0, budget_one=6.0)
        out["target_configs"] = {"bs_demo_200": {"config": config_of("bs_demo_200", tw, parallelism="cpu"),
                                                 "cpu_baseline": all_, "cpu_baseline_1thread": one}}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------
class Ctx:
    """per-process state shared by the workloads of one bench command"""

    def __init__(self):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from njode_b200 import _ext
        self.C, self.torch, self.dist = C, torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            # the driver reads ONE JSON line from stdout, and NCCL prints its version banner (NCCL_DEBUG >= VERSION) to
            # stdout when the first communicator comes up: point fd 1 at stderr until that has happened
            sys.stdout.flush()
            keep = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=self.dev)
                dist.barrier()
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(keep, 1)
                os.close(keep)
        lib = self.lib = _ext.cuda_lib().dll
        lib.njode_launch_count.restype = C.c_longlong
        lib.njode_get_timing.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float)]
        lib.njode_fma_peak_launch.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_void_p]
        lib.njode_l2_flush.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        lib.njode_last_kernel.argtypes = [C.c_int]
        lib.njode_last_kernel.restype = C.c_char_p
        self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)      # > 126 MB L2
        self.peaks = {}
        try:
            self.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        self.traffic = {}
        try:
            self.traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        except Exception:
            pass
        self._fma_peak = None

    def stream(self):
        return self.C.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)

    def flush(self):
        self.lib.njode_l2_flush(self.C.c_void_p(self.flush_buf.data_ptr()), self.flush_buf.numel(), self.stream())

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def fma_peak(self):
        """measured fp32 FMA peak of this GPU in TFLOP/s (dependent FFMA chains on every SM)"""
        if self._fma_peak is None:
            torch, C = self.torch, self.C
            scratch = torch.empty(148 * 8 * 256 * 2, dtype=torch.float32, device=self.dev)
            fm = C.c_double()
            peak = 0.0
            for it in range(4):
                a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                self.lib.njode_fma_peak_launch(C.c_void_p(scratch.data_ptr()), 4096, C.byref(fm), self.stream())
                b_.record()
                torch.cuda.synchronize()
                peak = max(peak, 2.0 * fm.value / (a.elapsed_time(b_) * 1e-3) / 1e12)
            self._fma_peak = peak
        return self._fma_peak


def measure_b200(ctx, wl_name, wl, steps, warmup, sample_clocks=True):
    """resident-input arm, end-to-end arm and the roofline of one workload on this rank's GPU; returns the JSON fields
    (rank 0) -- every rank must call it (collectives inside)."""
    torch, C, lib = ctx.torch, ctx.C, ctx.lib
    from njode_b200 import models
    from njode_b200 import dist as njdist
    world, rank, dev = ctx.world, ctx.rank, ctx.dev

    B = wl["paths"]                      # per GPU (weak scaling)
    first = rank * B
    dev_data = bool(wl.get("device_data"))
    generator = None
    if dev_data:
        # this rank's shard of the dataset, generated where it is used; batches are collated from it on the device
        n_ds = int(wl.get("dataset_paths", B))
        ds, gen_ms, gen_bytes = device_dataset(wl, 1234, rank * n_ds, n_ds, dev)
        dt = ds.dt
        batch = ds.collate(torch.arange(B), on_device=True)
        hbm = float(ctx.peaks.get("hbm_gbs", 6650.0))
        generator = {"paths_per_gpu": n_ds, "paths_total": n_ds * world, "bytes_written_per_gpu": gen_bytes, "ms": gen_ms,
                     "gb_per_s": gen_bytes / (gen_ms * 1e-3) / 1e9, "frac_of_hbm_peak": gen_bytes / (gen_ms * 1e-3) / 1e9 / hbm,
                     "note": "njode_sde_generate: fp64 paths [paths, d, steps+1] + int32 observation mask, CUDA events"}
    else:
        batch, dt = synth_batch(wl, 1234, first, B)
    # the Euler grid is batch-global (NJODE/models.py:430-439): all ranks use the union of times.
    # On the regular grid with >= 20k paths per rank every grid time is observed on every rank.
    T = horizon(wl)
    torch.manual_seed(0)
    model = models.NJODE(**model_cfg(wl)).to(dev)
    model.train()
    if world > 1:
        dp = njdist.DataParallel(model, global_batch_size=B * world)
        dp.set_batch(B * world, first)
    params = [p for p in model.parameters()]

    def args_of(b):
        return (b["times"], b["time_ptr"], b["X"], b["obs_idx"], dt, T, b["start_X"], b["n_obs_ot"])

    # ---- resident-input arm -----------------------------------------------------------------
    model.output_device = "cuda"
    kw_of = lambda b: ({"M": b["M"]} if "M" in b else {})
    pb = model.prepare_batch(*args_of(batch), **kw_of(batch))
    S = pb.sched.S
    units_per_step = B * S * world

    def step_resident():
        for p in params:
            p.grad = None
        hT, loss = model.forward_prepared(pb)
        loss.backward()
        return loss

    lib.njode_set_timing(1)
    for _ in range(max(warmup, 3)):
        step_resident()
    torch.cuda.synchronize()
    sampler = ClockSampler(ctx.local)
    if rank == 0 and sample_clocks:
        sampler.start()
    ctx.barrier()
    torch.cuda.synchronize()
    launches0 = lib.njode_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    kf, kb = [], []
    wall0 = time.perf_counter()
    for i in range(steps):
        ctx.flush()
        ev[i][0].record()
        step_resident()
        ev[i][1].record()
        if i < 8:                       # per-kernel CUDA-event durations (C ABI, launching stream)
            f, b_ = C.c_float(), C.c_float()
            lib.njode_get_timing(C.byref(f), C.byref(b_))
            kf.append(f.value); kb.append(b_.value)
    torch.cuda.synchronize()
    ctx.barrier()
    wall = time.perf_counter() - wall0
    launches = lib.njode_launch_count() - launches0
    total_ms = ctx.max_over_ranks(float(sum(a.elapsed_time(b_) for a, b_ in ev)))
    ms_per_step = total_ms / steps
    value = units_per_step / (ms_per_step * 1e-3)
    kname_f = (lib.njode_last_kernel(0) or b"").decode()
    kname_b = (lib.njode_last_kernel(1) or b"").decode()

    # ---- the same step with the backward in recompute mode (segment kernels): nothing saved by the forward pass -----
    recompute = None
    if kname_b.startswith("nj_seg_bwd"):
        keep_mode = model.recompute
        model.recompute = "on"
        for _ in range(3):
            step_resident()
        torch.cuda.synchronize()
        ctx.barrier()
        n_rc = max(3, min(steps, 10))
        evr = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_rc)]
        for i in range(n_rc):
            ctx.flush()
            evr[i][0].record()
            step_resident()
            evr[i][1].record()
        torch.cuda.synchronize()
        rc_ms = ctx.max_over_ranks(float(sum(a.elapsed_time(b_) for a, b_ in evr))) / n_rc
        recompute = {"ms_per_step": rc_ms, "value": units_per_step / (rc_ms * 1e-3), "saved_bytes_per_step": 0,
                     "history_bytes_per_step_without": 4 * (S * B * wl["H"] + pb.N * (wl["H"] + wl["d"])),
                     "note": "NJODE.recompute = 'on': the backward recomputes every segment from its checkpoint at the "
                             "observation time; default 'auto' keeps the [S, B, H] history up to 256 MiB"}
        model.recompute = keep_mode
        step_resident()

    # ---- end-to-end arm: public API, host tensors ---------------------------------------------
    model.output_device = "cpu"
    nb = 3
    if dev_data:
        # the e2e arm still hands HOST tensors to the public API: other batches of the same dataset, copied to the host here
        host_batches = []
        for j in range(nb):
            lo = ((j + 1) * B) % max(1, len(ds) - B + 1)
            hb = ds.collate(torch.arange(lo, lo + B))
            hb["X"], hb["start_X"] = hb["X"].cpu(), hb["start_X"].cpu()
            host_batches.append(hb)
    else:
        host_batches = [synth_batch(wl, 4321 + j, first, B)[0] for j in range(nb)]

    def step_e2e(b):
        for p in params:
            p.grad = None
        hT, loss = model(*args_of(b), **kw_of(b))
        loss.backward()
        return float(loss.detach())       # D2H of the step's result

    for j in range(2):
        step_e2e(host_batches[j % nb])
    torch.cuda.synchronize()
    ctx.barrier()
    n_e2e = max(3, min(steps, 10))
    t0 = time.perf_counter()
    h2d = 0
    for j in range(n_e2e):
        step_e2e(host_batches[j % nb])
        h2d += model.last_h2d_bytes
    torch.cuda.synchronize()
    e2e_s = ctx.max_over_ranks(time.perf_counter() - t0)
    e2e_val = units_per_step * n_e2e / e2e_s
    clocks = sampler.stop() if (rank == 0 and sample_clocks) else None

    # ---- roofline of the dominant kernel ------------------------------------------------------
    N_rows = pb.N
    F_ode, F_enc, F_ro = flops_per_unit(wl)
    fwd_flops = B * S * F_ode + N_rows * (F_enc + 2 * F_ro) + B * F_enc
    bwd_flops = 2 * fwd_flops
    peak = ctx.fma_peak()
    bwd_ms = float(np.median(kb)) if kb else float("nan")
    fwd_ms = float(np.median(kf)) if kf else float("nan")
    achieved = bwd_flops / (bwd_ms * 1e-3) / 1e12
    H = wl["H"]
    tensor_path = model.last_forward_path == "tcgen05"
    alg_bytes_bwd = 4 * (N_rows * (H + 2 * wl["d"]) + B * wl["d"]) + 4 * model._flat.numel()
    peaks = ctx.peaks
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    # DRAM bytes per launch of the dominant kernel: the committed `ncu --set full` capture of this workload
    # (dram__bytes_read.sum + dram__bytes_write.sum); profiles/ncu_traffic.json names the capture and the git commit it
    # was taken at -- null when there is no capture of this workload / kernel
    tr = ctx.traffic.get(wl_name, {}).get(kname_b, {})
    roofline = {"kernel": kname_b, "fwd_kernel": kname_f, "bound": "fp32_fma", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                "traffic": tr.get("dram_bytes"), "traffic_source": tr.get("source"), "traffic_commit": tr.get("commit"),
                "algorithmic_bytes": alg_bytes_bwd,
                "peak_source": "measured in this run (njode_fma_peak_launch: FFMA chains on all SMs)",
                "kernel_ms": bwd_ms, "fwd_kernel_ms": fwd_ms,
                "fwd_achieved": fwd_flops / (fwd_ms * 1e-3) / 1e12,
                "step_frac": 3.0 * fwd_flops / (ms_per_step * 1e-3) / 1e12 / peak if peak else None,
                "flops_per_launch": bwd_flops,
                "hbm": {"bound": "hbm", "achieved": alg_bytes_bwd / (bwd_ms * 1e-3) / 1e9,
                        "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg_bytes_bwd / (bwd_ms * 1e-3) / 1e9 / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}

    if tensor_path:
        # wide nets: the step runs on the tcgen05 kernels (njode_wide_*): bf16 operands, fp32 accumulation.
        # Roofline = all MLP flops of fwd+bwd (3x forward) over the summed durations of the
        # tensor-core kernels (CUDA events on the launching stream, last timed step), against the measured
        # sustained bf16 GEMM peak (the kernels run inside a long step).
        lib.njode_wide_get_timing.argtypes = [C.POINTER(C.c_float)] * 3
        lib.njode_wide_get_timing_bwd.argtypes = [C.POINTER(C.c_float)] * 2
        te, to, tr_, tc_, td = (C.c_float() for _ in range(5))
        lib.njode_wide_get_timing(C.byref(te), C.byref(to), C.byref(tr_))
        lib.njode_wide_get_timing_bwd(C.byref(tc_), C.byref(td))
        k_ms = te.value + to.value + tr_.value + tc_.value + td.value
        tpeak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        ach = 3.0 * fwd_flops / (k_ms * 1e-3) / 1e12
        trw = ctx.traffic.get(wl_name, {}).get("nj_wide", {})
        roofline = {"kernel": "nj_wide_kernel + nj_wide_bwd_kernel + nj_wide_dw_kernel (tcgen05)", "bound": "tensor",
                    "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak,
                    "traffic": trw.get("dram_bytes"), "traffic_source": trw.get("source"), "traffic_commit": trw.get("commit"),
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback",
                    "kernel_ms": k_ms, "fwd_enc_ms": te.value, "fwd_ode_ms": to.value, "fwd_ro_ms": tr_.value,
                    "bwd_chain_ms": tc_.value, "bwd_dw_ms": td.value, "flops_per_launch": 3.0 * fwd_flops}
    del model, pb, host_batches, batch
    if dev_data:
        del ds
    torch.cuda.empty_cache()
    res_extra = {"generator": generator} if generator else {}
    if recompute:
        res_extra["recompute"] = recompute
    return {**res_extra, "value": value, "ms_per_step": ms_per_step, "steps": steps, "warmup": max(warmup, 3),
            "dtype": "bf16" if tensor_path else "f32",
            "config": config_of(wl_name, wl, euler_steps=S, obs_rows_per_gpu=N_rows, parallelism="dp%d" % world,
                                l2="flushed between timed steps (256 MiB write)"),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d / n_e2e),
                    "d2h_bytes_per_step": 4, "steps": n_e2e},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "wall_s_timed_region": wall}


def dp_check(ctx):
    """N > 1, outside every timed region: the NCCL-reduced flat gradient (a) is bit-identical on all ranks and (b) equals
    the gradient rank 0 computes alone on the concatenated batch (element-wise rtol 1e-5 + 1e-7 of the tensor's max).
    Demo nets (fp32 kernels), dropout on (keys use global path ids: rank-count invariant), 96 paths per rank."""
    torch, dist = ctx.torch, ctx.dist
    from njode_b200 import models
    from njode_b200 import dist as njdist
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    wl = dict(WORKLOADS["bs_demo_200"], paths=96 * world)
    full, dt = synth_batch(wl, 777, 0, wl["paths"])
    shard, first = njdist.shard_batch(full, rank, world)

    def grad_of(batch, bs_norm, first_id, sync):
        torch.manual_seed(0)
        m = models.NJODE(**model_cfg(wl)).to(dev)
        m.train()
        m.batch_size_norm, m.path_id_offset = bs_norm, first_id
        if sync:
            njdist.DataParallel(m, global_batch_size=bs_norm).set_batch(bs_norm, first_id)
        torch.manual_seed(99)                     # the same dropout seed on every rank and in the single-GPU run
        hT, loss = m(batch["times"], batch["time_ptr"], batch["X"], batch["obs_idx"], dt, 1.0, batch["start_X"], batch["n_obs_ot"])
        loss.backward()
        return torch.cat([p.grad.reshape(-1) for p in m.parameters()]), float(loss)

    g, part_loss = grad_of(shard, wl["paths"], first, True)
    gathered = [torch.empty_like(g) for _ in range(world)]
    dist.all_gather(gathered, g)
    identical = all(bool(torch.equal(gathered[0], x)) for x in gathered[1:])
    lsum = torch.tensor([part_loss], device=dev, dtype=torch.float64)
    dist.all_reduce(lsum)
    out = None
    if rank == 0:
        g1, loss1 = grad_of(full, wl["paths"], 0, False)
        err = (g - g1).abs()
        lim = 1e-5 * g1.abs() + 1e-7 * float(g1.abs().max())
        out = {"ranks": world, "paths": wl["paths"], "grad_floats": int(g.numel()),
               "identical_across_ranks": identical,
               "max_abs_err_vs_single_gpu": float(err.max()), "max_norm_rel_err": float(err.max() / g1.abs().max()),
               "within_rtol_1e-5": bool((err <= lim).all()),
               "loss_sum_over_ranks": float(lsum.item()), "loss_single_gpu": loss1}
        assert identical, "dp_check: the all-reduced gradient differs between ranks"
        assert out["within_rtol_1e-5"], "dp_check: data-parallel gradient != single-GPU gradient: %r" % out
    dist.barrier()
    return out


def run_b200(args, wl_name, wl):
    ctx = Ctx()
    head = measure_b200(ctx, wl_name, wl, args.steps, args.warmup)
    targets = {}
    if args.targets:
        if ctx.world == 1 and wl_name != "bs_demo_200":
            targets["bs_demo_200"] = measure_b200(ctx, "bs_demo_200", WORKLOADS["bs_demo_200"], max(args.steps, 20), args.warmup)
        if wl_name != "bs_scaled_d16_h256":
            targets["bs_scaled_d16_h256"] = measure_b200(ctx, "bs_scaled_d16_h256", WORKLOADS["bs_scaled_d16_h256"],
                                                         max(3, min(args.steps, 6)), args.warmup)
    check = dp_check(ctx) if ctx.world > 1 else None
    if ctx.rank != 0:
        if ctx.world > 1:
            ctx.dist.destroy_process_group()
        return
    out = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": ctx.world, "steps": args.steps,
           "warmup": head["warmup"], "ms_per_step": head["ms_per_step"], "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": head["dtype"], "data": "synthetic",
           "config": head["config"], "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "clocks": head["clocks"],
           "roofline": head["roofline"], "wall_s_timed_region": head["wall_s_timed_region"]}
    for k in ("recompute", "generator"):
        if k in head:
            out[k] = head[k]
    if check is not None:
        out["dp_check"] = check
    if ctx.world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"], out["cpu_baseline_1thread"] = cpu_arms(wl)
    for name, t in targets.items():
        t.update({"metric": METRIC, "unit": UNIT, "n_gpus": ctx.world, "scaling": "weak"})
        if ctx.world == 1 and not args.no_cpu_baseline:
            tw = WORKLOADS[name]
            t["cpu_baseline"], t["cpu_baseline_1thread"] = cpu_arms(tw, budget_all=6.0, budget_one=6.0)
            t["same_config_as_cpu_arm"] = tw["cpu_sample_paths"] == tw["paths"] and "cpu_sample_steps" not in tw
            for k, cb in (("speedup_e2e_vs_cpu_all_threads", "cpu_baseline"), ("speedup_e2e_vs_cpu_1thread", "cpu_baseline_1thread")):
                t[k] = t["e2e"]["value"] / t[cb]["value"]
    if targets:
        out["target_configs"] = targets
    print(json.dumps(out), flush=True)
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="heston_demo_20k", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-targets", dest="targets", action="store_false",
                    help="skip the target_configs sub-records (bs_demo_200, bs_scaled_d16_h256)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, args.workload, wl)
    else:
        run_b200(args, args.workload, wl)


if __name__ == "__main__":
    main()
